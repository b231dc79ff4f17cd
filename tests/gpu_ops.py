"""numpy-in / numpy-out wrappers that call the CUDA kernels through the C ABI (uzu_b200.binding).
Same argument names as oracle/oracle.py so the parity tests read symmetrically."""
import ctypes as C

import numpy as np

from uzu_b200 import binding as B

PROLOGUE = {0: B.B_FULL_PRECISION, 1: B.B_SCALE_BIAS, 2: B.B_SCALE_ZERO_POINT, 3: B.B_SCALE_SYMMETRIC}  # oracle QM_* -> B_*


def _up(ctx, a):
    return ctx.upload(a) if a is not None else None


def _ptr(b):
    return b.ptr if b is not None else 0


def matmul(ctx, a, w, *, m, n, k, scales=None, zero_points=None, biases=None, method=0, bits=4, group_size=64,
           signed_codes=False, b_transpose=True, ld=0, d=None, d_f32=False, gather=None, ab_scale=1.0, accumulate=False,
           bias=None, soft_cap=None, a_f32=False, w_f32=False, return_launches=False, rht_factors=None):
    if d is None:
        d = np.zeros((m, n), np.float32 if d_f32 else np.uint16)
    if rht_factors is not None:
        rht_factors = np.ascontiguousarray(rht_factors, dtype=np.int32)
    bufs = [_up(ctx, x) for x in (a, w, scales, zero_points, biases, d, bias, gather, rht_factors)]
    ba, bw, bs, bz, bb, bd, bbias, bg, brht = bufs
    mask = B.D_RHT if rht_factors is not None else 0
    if ab_scale != 1.0:
        mask |= B.D_SCALE
    if accumulate:
        mask |= B.D_ACCUMULATE
    if bias is not None:
        mask |= B.D_BIAS
    if soft_cap is not None:
        mask |= B.D_SOFT_CAP
    args = B.MatmulArgs(a=_ptr(ba), b=_ptr(bw), b_scales=_ptr(bs), b_zero_points=_ptr(bz), b_biases=_ptr(bb), d=_ptr(bd),
                        bias=_ptr(bbias), gather_indices=_ptr(bg), rht_factors=_ptr(brht), b_prologue=PROLOGUE[method],
                        b_mode=B.QMODE_U4 if bits == 4 else B.QMODE_U8, b_group_size=group_size,
                        b_signed_codes=int(signed_codes), b_leading_dimension=ld, b_transpose=int(b_transpose),
                        d_transform=mask, ab_scale=ab_scale, soft_cap=soft_cap or 0.0, m=m, n=n, k=k,
                        weights_dt=B.DT_F32 if w_f32 else B.DT_BF16, input_dt=B.DT_F32 if a_f32 else B.DT_BF16,
                        output_dt=B.DT_F32 if d.dtype == np.float32 else B.DT_BF16)
    with ctx.command_buffer("matmul") as cmd:
        cmd.encode("uzu_matmul_encode", C.byref(args))
    out = bd.numpy(d.dtype, d.shape)
    for b in bufs:
        if b is not None:
            b.close()
    if return_launches:
        return out, cmd.launches
    return out


def activation_transform(ctx, x, factors, *, op=0, in_place=False, activation_group_size=0, sum_group_size=0):
    """uzu_activation_transform_encode; same returns as oracle.activation_transform."""
    rows, cols = x.shape
    factors = np.ascontiguousarray(factors, dtype=np.int32)
    bx, bf = ctx.upload(np.ascontiguousarray(x)), ctx.upload(factors)
    dt = B.DT_F32 if x.dtype == np.float32 else B.DT_BF16
    if op in (0, 1):
        bo = bx if in_place else ctx.upload(np.zeros_like(x))
        args = B.ActivationTransformArgs(input=0 if in_place else bx.ptr, fp_out=bo.ptr, rht_factors=bf.ptr, batch_size=rows,
                                         element_count=cols, ops=op, in_place=int(in_place), data_type=dt)
        with ctx.command_buffer("rht") as cmd:
            cmd.encode("uzu_activation_transform_encode", C.byref(args))
        return bo.numpy(x.dtype, x.shape)
    bq = ctx.upload(np.zeros((rows, cols), np.int8))
    bs = ctx.upload(np.zeros((rows, cols // activation_group_size), np.float32))
    bg = ctx.upload(np.zeros((rows, cols // sum_group_size), np.int32)) if op == 3 else None
    args = B.ActivationTransformArgs(input=bx.ptr, q_out=bq.ptr, scales_out=bs.ptr, group_sums_out=_ptr(bg), rht_factors=bf.ptr,
                                     batch_size=rows, element_count=cols, ops=op, in_place=0,
                                     activation_scale_group_size=activation_group_size, sum_group_size=sum_group_size, data_type=dt)
    with ctx.command_buffer("rht-quantize") as cmd:
        cmd.encode("uzu_activation_transform_encode", C.byref(args))
    return (bq.numpy(np.int8, (rows, cols)), bs.numpy(np.float32, (rows, cols // activation_group_size)),
            bg.numpy(np.int32, (rows, cols // sum_group_size)) if bg is not None else None)


def normalization(ctx, inp, scales, *, shortcut=None, residual_add=False, epsilon=1e-5, scale_offset=0.0, full_layer=False,
                  subtract_mean=False):
    rows, n = inp.shape
    bi, bs, bo = ctx.upload(inp), _up(ctx, scales), ctx.upload(np.zeros((rows, n), np.uint16))
    bsc = _up(ctx, shortcut)
    args = B.NormalizationArgs(input=bi.ptr, scales=_ptr(bs), output=bo.ptr, shortcut=_ptr(bsc), batch_size=rows, element_count=n,
                               epsilon=epsilon, scale_offset=scale_offset, post_layer_scalar=1.0, subtract_mean=int(subtract_mean),
                               full_layer=int(full_layer), copy_to_shortcut=int(shortcut is not None), residual_add=int(residual_add),
                               has_scales=int(scales is not None))
    with ctx.command_buffer("norm") as cmd:
        cmd.encode("uzu_normalization_encode", C.byref(args))
    out = bo.numpy(np.uint16, (rows, n))
    sc = bsc.numpy(np.uint16, (rows, n)) if bsc is not None else None
    return out, sc


def qkv_norm(ctx, qkv, scales, *, total_heads, head_dim, epsilon, scale_offset, head_offset, head_count, full_layer):
    rows = qkv.shape[0]
    bq, bs = ctx.upload(qkv), _up(ctx, scales)
    args = B.QkvNormArgs(scales=_ptr(bs), qkv_output=bq.ptr, batch_size=rows, total_heads=total_heads, head_dim=head_dim,
                         epsilon=epsilon, scale_offset=scale_offset, head_offset=head_offset, head_count=head_count,
                         full_layer=int(full_layer), in_place=1, has_scales=int(scales is not None))
    with ctx.command_buffer("qkvnorm") as cmd:
        cmd.encode("uzu_qkv_norm_encode", C.byref(args))
    return bq.numpy(np.uint16, qkv.shape)


def attention_prepare(ctx, qkv, keys, values, cos, sin, *, num_q_heads, num_kv_heads, head_dim, rope_dim, kv_token_offset):
    m = qkv.shape[0]
    bq, bk, bv = ctx.upload(qkv), ctx.upload(keys), ctx.upload(values)
    bc, bs = _up(ctx, cos), _up(ctx, sin)
    bqo = ctx.upload(np.zeros((num_q_heads, m, head_dim), np.uint16))
    args = B.AttentionPrepareArgs(qkv=bq.ptr, queries=bqo.ptr, keys=bk.ptr, values=bv.ptr, cosines=_ptr(bc), sines=_ptr(bs),
                                  num_q_heads=num_q_heads, num_kv_heads=num_kv_heads, head_dim=head_dim, rope_dim=rope_dim or 0,
                                  kv_token_offset=kv_token_offset, batch_dim=m, has_kv=1, has_rope=int(cos is not None))
    with ctx.command_buffer("prepare") as cmd:
        cmd.encode("uzu_attention_prepare_encode", C.byref(args))
    return bqo.numpy(np.uint16, (num_q_heads, m, head_dim)), bk.numpy(np.uint16, keys.shape), bv.numpy(np.uint16, values.shape)


def _attn_args(ctx, queries, keys, values, *, head_dim, gqa_factor, sequence_length, k_head_stride, k_seq_stride, v_head_stride,
               v_seq_stride, scale, num_heads, suffix_length, is_causal=True, sinks=None, ring=None, sliding_window=None, trie=None):
    bq, bk, bv = ctx.upload(queries), ctx.upload(keys), ctx.upload(values)
    bsink, btrie = _up(ctx, sinks), _up(ctx, trie)
    args = B.AttentionArgs(queries=bq.ptr, keys=bk.ptr, values=bv.ptr, gqa_factor=gqa_factor, sequence_length=sequence_length,
                           k_head_stride=k_head_stride, k_seq_stride=k_seq_stride, v_head_stride=v_head_stride,
                           v_seq_stride=v_seq_stride, ring_params=B.RingParams(*(ring or (0, 0))), scale=scale, trie=_ptr(btrie),
                           sliding_window_size=sliding_window or 0, sinks=_ptr(bsink), num_heads=num_heads,
                           suffix_length=suffix_length, head_dim=head_dim, has_sinks=int(sinks is not None),
                           is_kv_cache_ring=int(ring is not None), is_causal=int(is_causal), is_trie=int(trie is not None),
                           is_sliding_window=int(sliding_window is not None))
    return args, (bq, bk, bv, bsink, btrie)


def attention_single_pass(ctx, queries, keys, values, **kw):
    args, keep = _attn_args(ctx, queries, keys, values, **kw)
    S, H, D = args.suffix_length, args.num_heads, args.head_dim
    bo = ctx.upload(np.zeros((S, H, D), np.uint16))
    args.out = bo.ptr
    with ctx.command_buffer("attn1") as cmd:
        cmd.encode("uzu_attention_single_pass_encode", C.byref(args))
    return bo.numpy(np.uint16, (S, H, D))


def attention_two_pass(ctx, queries, keys, values, **kw):
    args, keep = _attn_args(ctx, queries, keys, values, **kw)
    S, H, D = args.suffix_length, args.num_heads, args.head_dim
    bp = ctx.upload(np.full((S, H, 32, D), 777.0, np.float32))
    bs = ctx.upload(np.full((S, H, 32), 777.0, np.float32))
    bm = ctx.upload(np.full((S, H, 32), 777.0, np.float32))
    bo = ctx.upload(np.zeros((S, H, D), np.uint16))
    args.out, args.sums, args.maxs = bp.ptr, bs.ptr, bm.ptr
    a2 = B.AttentionTwoPass2Args(partials=bp.ptr, sums=bs.ptr, maxs=bm.ptr, out=bo.ptr, num_heads=H, suffix_length=S, head_dim=D)
    with ctx.command_buffer("attn2") as cmd:
        cmd.encode("uzu_attention_two_pass1_encode", C.byref(args))
        cmd.encode("uzu_attention_two_pass2_encode", C.byref(a2))
    return bo.numpy(np.uint16, (S, H, D))


def kv_cache_update(ctx, keys, values, copies, element_dim):
    bk, bv = ctx.upload(keys), ctx.upload(values)
    arr = (B.KvCopy * len(copies))(*[B.KvCopy(int(s), int(d)) for s, d in copies])
    args = B.KvCacheUpdateArgs(in_place_keys=bk.ptr, in_place_values=bv.ptr, copies=arr, copy_count=len(copies), element_dim=element_dim)
    with ctx.command_buffer("kvupd") as cmd:
        cmd.encode("uzu_kv_cache_update_encode", C.byref(args))
    return bk.numpy(np.uint16, keys.shape), bv.numpy(np.uint16, values.shape)


def sigmoid_gate(ctx, gate, output):
    bg, bo = ctx.upload(gate), ctx.upload(output)
    with ctx.command_buffer("gate") as cmd:
        cmd.encode("uzu_sigmoid_gate_encode", bg.ptr, bo.ptr, output.size)
    return bo.numpy(np.uint16, output.shape)


def gated_act_mul(ctx, fused_up, gated_dim, act=0):
    rows = fused_up.shape[0]
    bi, bo = ctx.upload(fused_up), ctx.upload(np.zeros((rows, gated_dim), np.uint16))
    args = B.GatedActMulArgs(act_operand=bi.ptr, fp_out=bo.ptr, gated_dim=gated_dim, batch_dim=rows, act_type=act, interleaved=1)
    with ctx.command_buffer("gam") as cmd:
        cmd.encode("uzu_gated_act_mul_encode", C.byref(args))
    return bo.numpy(np.uint16, (rows, gated_dim))


def quant_embedding_lookup(ctx, token_ids, weights, scales, *, zero_points=None, biases=None, vocab_size, model_dim, input_scale=1.0,
                           group_size=64, mode=0, method=2):
    token_ids = np.ascontiguousarray(token_ids, np.uint32)
    bt, bw, bs, bz, bb = ctx.upload(token_ids), ctx.upload(weights), ctx.upload(scales), _up(ctx, zero_points), _up(ctx, biases)
    bo = ctx.upload(np.zeros((len(token_ids), model_dim), np.uint16))
    qmeth = {1: B.QMETHOD_SCALE_BIAS, 2: B.QMETHOD_SCALE_ZERO_POINT, 3: B.QMETHOD_SCALE_SYMMETRIC}[method]
    args = B.QuantizedEmbeddingLookupArgs(token_ids=bt.ptr, weights=bw.ptr, scales=bs.ptr, zero_points=_ptr(bz), biases=_ptr(bb),
                                          output=bo.ptr, batch_size=len(token_ids), vocab_size=vocab_size, model_dim=model_dim,
                                          input_scale=input_scale, group_size=group_size, quantization_mode=mode,
                                          quantization_method=qmeth)
    with ctx.command_buffer("emb") as cmd:
        cmd.encode("uzu_quantized_embedding_lookup_encode", C.byref(args))
    return bo.numpy(np.uint16, (len(token_ids), model_dim))


def fp_embedding_lookup(ctx, token_ids, weights, *, vocab_size, model_dim, input_scale=1.0):
    token_ids = np.ascontiguousarray(token_ids, np.uint32)
    bt, bw = ctx.upload(token_ids), ctx.upload(weights)
    bo = ctx.upload(np.zeros((len(token_ids), model_dim), np.uint16))
    with ctx.command_buffer("emb") as cmd:
        cmd.encode("uzu_full_precision_embedding_lookup_encode", bt.ptr, bw.ptr, bo.ptr, len(token_ids), vocab_size, model_dim, input_scale)
    return bo.numpy(np.uint16, (len(token_ids), model_dim))


def logit_transform(ctx, logits, scale, soft_cap=None):
    b = ctx.upload(logits)
    with ctx.command_buffer("lt") as cmd:
        cmd.encode("uzu_logit_transform_encode", b.ptr, logits.size, scale, soft_cap or 0.0, int(soft_cap is not None))
    return b.numpy(np.uint16, logits.shape)


def tensor_add_scale(ctx, inp, bias, num_cols, scale):
    bi, bb, bo = ctx.upload(inp), ctx.upload(bias), ctx.upload(np.zeros_like(inp))
    with ctx.command_buffer("tas") as cmd:
        cmd.encode("uzu_tensor_add_scale_encode", bi.ptr, bb.ptr, bo.ptr, num_cols, inp.size, scale)
    return bo.numpy(np.uint16, inp.shape)


def tensor_add_bias(ctx, inp, bias, num_cols):
    bi, bb, bo = ctx.upload(inp), ctx.upload(bias), ctx.upload(np.zeros_like(inp))
    with ctx.command_buffer("tab") as cmd:
        cmd.encode("uzu_tensor_add_bias_encode", bi.ptr, bb.ptr, bo.ptr, num_cols, inp.size)
    return bo.numpy(np.uint16, inp.shape)


def tensor_add_swap(ctx, skip, main):
    bs, bm = ctx.upload(skip), ctx.upload(main)
    with ctx.command_buffer("swap") as cmd:
        cmd.encode("uzu_tensor_add_swap_encode", bs.ptr, bm.ptr, skip.size)
    return bs.numpy(np.uint16, skip.shape), bm.numpy(np.uint16, main.shape)


def tensor_copy(ctx, src):
    bs, bd = ctx.upload(src), ctx.upload(np.zeros_like(src))
    with ctx.command_buffer("copy") as cmd:
        cmd.encode("uzu_tensor_copy_encode", bs.ptr, bd.ptr, src.size)
    return bd.numpy(np.uint16, src.shape)


def unified_sampling(ctx, logits, *, seeds=None, bitmask=None, temperature=None, top_k=None, top_p=None, min_p=None):
    rows, V = logits.shape
    bl, bo = ctx.upload(logits), ctx.upload(np.zeros(rows, np.uint32))
    bs = _up(ctx, np.ascontiguousarray(seeds, np.uint64)) if seeds is not None else None
    bm = _up(ctx, bitmask)
    args = B.UnifiedSamplingArgs(logits=bl.ptr, output=bo.ptr, seeds=_ptr(bs), bitmask=_ptr(bm), temperature=temperature or 0.0,
                                 top_k=top_k or 0, top_p=top_p or 0.0, min_p=min_p or 0.0, vocab_size=V, batch_size=rows,
                                 is_stochastic=int(seeds is not None), has_bitmask=int(bitmask is not None),
                                 has_temperature=int(temperature is not None), has_top_k=int(top_k is not None),
                                 has_top_p=int(top_p is not None), has_min_p=int(min_p is not None))
    with ctx.command_buffer("sampling") as cmd:
        cmd.encode("uzu_unified_sampling_encode", C.byref(args))
    return bo.numpy(np.uint32, (rows,))


def delta_net_conv_update(ctx, conv_weight, bias, in_out, state, kernel_size, conv_dim):
    bw, bb, bio, bs = ctx.upload(conv_weight), _up(ctx, bias), ctx.upload(in_out), ctx.upload(state)
    args = B.DeltaNetConvUpdateArgs(conv_weight=bw.ptr, bias=_ptr(bb), in_out=bio.ptr, state=bs.ptr, kernel_size=kernel_size,
                                    conv_dim=conv_dim, state_stride=kernel_size - 1, has_bias=int(bias is not None))
    with ctx.command_buffer("conv") as cmd:
        cmd.encode("uzu_delta_net_conv_update_encode", C.byref(args))
    return bio.numpy(np.uint16, in_out.shape), bs.numpy(np.float32, state.shape)


def delta_net_update(ctx, in_proj, a_log, dt_bias, norm_weight, state, *, num_v_heads, num_k_heads, head_k_dim, head_v_dim, key_dim,
                     value_dim, norm_epsilon):
    bi, ba, bd, bn, bs = (ctx.upload(x) for x in (in_proj, a_log, dt_bias, norm_weight, state))
    bo = ctx.upload(np.zeros(value_dim, np.uint16))
    args = B.DeltaNetUpdateArgs(in_proj=bi.ptr, a_log=ba.ptr, dt_bias=bd.ptr, norm_weight=bn.ptr, state=bs.ptr, out=bo.ptr,
                                num_v_heads=num_v_heads, num_k_heads=num_k_heads, head_v_dim=head_v_dim, key_dim=key_dim,
                                value_dim=value_dim, norm_epsilon=norm_epsilon, head_k_dim=head_k_dim)
    with ctx.command_buffer("dn") as cmd:
        cmd.encode("uzu_delta_net_update_encode", C.byref(args))
    return bo.numpy(np.uint16, (value_dim,)), bs.numpy(np.float32, state.shape)
