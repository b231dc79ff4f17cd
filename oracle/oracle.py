"""ctypes wrapper over oracle/_build/liboracle.so (the C restatement of uzu's CPU backend).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs. The product package (uzu_b200) never imports this.

bf16 tensors are numpy uint16 arrays holding the raw bits.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "liboracle.so"

DT_BF16, DT_F32 = 0, 1
QM_NONE, QM_SCALE_BIAS, QM_ZERO_POINT, QM_SYMMETRIC = 0, 1, 2, 3
ACT_SILU, ACT_GELU_APPROX, ACT_GELU_EXACT, ACT_IDENTITY, ACT_SOFTPLUS = 0, 1, 2, 3, 4
MODE_U4, MODE_I8, MODE_U8 = 0, 1, 2
TWO_PASS_BLOCKS = 32


def build(force: bool = False) -> Path:
    """Compile the oracle with the committed recipe (oracle/Makefile)."""
    src = _HERE / "uzu_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.oracle_bf16_to_f32.restype = C.c_float
        _lib.oracle_bf16_to_f32.argtypes = [C.c_uint16]
        _lib.oracle_f32_to_bf16.restype = C.c_uint16
        _lib.oracle_f32_to_bf16.argtypes = [C.c_float]
        _lib.oracle_unit_interval.restype = C.c_float
        _lib.oracle_unit_interval.argtypes = [C.c_uint32]
        _lib.oracle_gumbel_float.restype = C.c_float
        _lib.oracle_gumbel_float.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32]
        _lib.oracle_prng_derive.restype = C.c_uint64
        _lib.oracle_prng_derive.argtypes = [C.c_uint64, C.c_uint64]
        _lib.oracle_max_threads.restype = C.c_int
    return _lib


# ---------------------------------------------------------------------------------------------
# bf16 helpers (numpy, vectorised; identical to `half` RNE; checked against the C functions
# in tests/test_oracle_pins.py)
# ---------------------------------------------------------------------------------------------
def f32_to_bf16(x) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32)
    nan = (u & 0x7FFFFFFF) > 0x7F800000
    round_bit = np.uint32(0x8000)
    up = ((u & round_bit) != 0) & ((u & np.uint32(0x17FFF)) != 0)
    r = (u >> 16).astype(np.uint32) + up.astype(np.uint32)
    r = np.where(nan, (u >> 16) | 0x40, r)
    return r.astype(np.uint16)


def bf16_to_f32(h) -> np.ndarray:
    h = np.ascontiguousarray(h, dtype=np.uint16)
    return (h.astype(np.uint32) << 16).view(np.float32)


def _p(a, ty=C.c_void_p):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "oracle inputs must be C-contiguous"
    return a.ctypes.data_as(ty)


class _MatmulArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("a_dt", C.c_int),
        ("w", C.c_void_p), ("scales", C.c_void_p), ("zero_points", C.c_void_p), ("biases", C.c_void_p),
        ("w_dt", C.c_int), ("method", C.c_int), ("bits", C.c_int), ("group_size", C.c_int),
        ("signed_codes", C.c_int), ("b_transpose", C.c_int), ("ld", C.c_int),
        ("d", C.c_void_p), ("d_dt", C.c_int), ("gather", C.c_void_p),
        ("ab_scale", C.c_float), ("accumulate", C.c_int), ("bias", C.c_void_p),
        ("has_soft_cap", C.c_int), ("soft_cap", C.c_float),
        ("m", C.c_int), ("n", C.c_int), ("k", C.c_int),
    ]


def matmul(a, w, *, m, n, k, scales=None, zero_points=None, biases=None, method=QM_NONE, bits=4, group_size=64,
           signed_codes=False, b_transpose=True, ld=0, d=None, d_f32=False, gather=None, ab_scale=1.0,
           accumulate=False, bias=None, soft_cap=None, w_f32=False, a_f32=False, threads=1, rht_factors=None):
    """D = epilogue(A @ dequant(W)^T); returns D ([m,n] uint16 bf16 bits, or float32 if d_f32).
    rht_factors (MatmulDOps::rht_factors, i32 [n]): the matmul runs WITHOUT its bias, then OutputRht in place over D, then the bias
    (cpu/kernel/matmul/kernel.rs:64,162,285,297-303: bias_after_rht)."""
    if rht_factors is not None:
        assert not d_f32 and gather is None
        out = matmul(a, w, m=m, n=n, k=k, scales=scales, zero_points=zero_points, biases=biases, method=method, bits=bits,
                     group_size=group_size, signed_codes=signed_codes, b_transpose=b_transpose, ld=ld, d=d, ab_scale=ab_scale,
                     accumulate=accumulate, bias=None, soft_cap=soft_cap, w_f32=w_f32, a_f32=a_f32, threads=threads)
        activation_transform(out, rht_factors, op=RHT_OUTPUT, in_place=True)
        if bias is not None:
            out[...] = tensor_add_bias(out, bias, n)
        return out
    if d is None:
        d = np.zeros((m, n), dtype=np.float32 if d_f32 else np.uint16)
    args = _MatmulArgs(
        a=_p(a).value, a_dt=DT_F32 if a_f32 else DT_BF16, w=_p(w).value,
        scales=_p(scales).value if scales is not None else None,
        zero_points=_p(zero_points).value if zero_points is not None else None,
        biases=_p(biases).value if biases is not None else None,
        w_dt=DT_F32 if w_f32 else DT_BF16, method=method, bits=bits, group_size=group_size,
        signed_codes=int(signed_codes), b_transpose=int(b_transpose), ld=ld,
        d=_p(d).value, d_dt=DT_F32 if d.dtype == np.float32 else DT_BF16,
        gather=_p(gather).value if gather is not None else None,
        ab_scale=ab_scale, accumulate=int(accumulate), bias=_p(bias).value if bias is not None else None,
        has_soft_cap=int(soft_cap is not None), soft_cap=soft_cap or 0.0, m=m, n=n, k=k,
    )
    lib().oracle_matmul(C.byref(args), C.c_int(threads))
    return d


class _NormArgs(C.Structure):
    _fields_ = [
        ("input", C.c_void_p), ("scales", C.c_void_p), ("biases", C.c_void_p), ("output", C.c_void_p),
        ("shortcut", C.c_void_p), ("batch_size", C.c_int), ("element_count", C.c_int),
        ("epsilon", C.c_float), ("scale_offset", C.c_float), ("post_layer_scalar", C.c_float),
        ("in_place", C.c_int), ("subtract_mean", C.c_int), ("full_layer", C.c_int),
        ("copy_to_shortcut", C.c_int), ("residual_add", C.c_int), ("scale_residual_sum", C.c_int),
        ("scale_output", C.c_int),
    ]


def normalization(inp, scales, *, shortcut=None, residual_add=False, epsilon=1e-5, scale_offset=0.0,
                  full_layer=False, subtract_mean=False, biases=None, post_layer_scalar=1.0,
                  scale_residual_sum=False, scale_output=False):
    """Returns output [rows, n] (bf16 bits); `shortcut` is updated in place when given."""
    rows, n = inp.shape
    out = np.zeros((rows, n), dtype=np.uint16)
    args = _NormArgs(
        input=_p(inp).value, scales=_p(scales).value if scales is not None else None,
        biases=_p(biases).value if biases is not None else None, output=_p(out).value,
        shortcut=_p(shortcut).value if shortcut is not None else None, batch_size=rows, element_count=n,
        epsilon=epsilon, scale_offset=scale_offset, post_layer_scalar=post_layer_scalar, in_place=0,
        subtract_mean=int(subtract_mean), full_layer=int(full_layer),
        copy_to_shortcut=int(shortcut is not None), residual_add=int(residual_add),
        scale_residual_sum=int(scale_residual_sum), scale_output=int(scale_output),
    )
    lib().oracle_normalization(C.byref(args))
    return out


def qkv_norm(qkv, scales, *, total_heads, head_dim, epsilon, scale_offset, head_offset, head_count, full_layer):
    rows = qkv.shape[0]
    lib().oracle_qkv_norm(_p(qkv), _p(scales), C.c_int(rows), C.c_int(total_heads), C.c_int(head_dim),
                          C.c_float(epsilon), C.c_float(scale_offset), C.c_int(head_offset), C.c_int(head_count),
                          C.c_int(int(full_layer)))
    return qkv


class _RopeCfg(C.Structure):
    _fields_ = [("kind", C.c_int), ("base", C.c_float), ("head_dim", C.c_int), ("scaling_factor", C.c_float),
                ("original_context_length", C.c_int), ("low_frequency_factor", C.c_float),
                ("high_frequency_factor", C.c_float)]


ROPE_KINDS = {"UnscaledRoPEConfig": 0, "LinearScalingRoPEConfig": 1, "LlamaRoPEConfig": 2}


def rope_tables(cfg: dict, positions):
    positions = np.ascontiguousarray(positions, dtype=np.uint32)
    hd = cfg["head_dim"]
    cos = np.zeros((len(positions), hd), dtype=np.float32)
    sin = np.zeros((len(positions), hd), dtype=np.float32)
    c = _RopeCfg(kind=ROPE_KINDS[cfg["type"]], base=cfg["base"], head_dim=hd,
                 scaling_factor=cfg.get("scaling_factor", 1.0),
                 original_context_length=cfg.get("original_context_length", 0),
                 low_frequency_factor=cfg.get("low_frequency_factor", 1.0),
                 high_frequency_factor=cfg.get("high_frequency_factor", 1.0))
    lib().oracle_rope_tables(C.byref(c), _p(positions), C.c_int(len(positions)), _p(cos), _p(sin))
    return cos, sin


def attention_prepare(qkv, keys, values, cos, sin, *, num_q_heads, num_kv_heads, head_dim, rope_dim,
                      kv_token_offset, has_kv=True):
    """qkv [m, total_heads*D]; writes K/V rows into `keys`/`values` ([T, Hkv*D]); returns queries [Hq, m, D]."""
    m = qkv.shape[0]
    queries = np.zeros((max(num_q_heads, 1), m, head_dim), dtype=np.uint16)
    has_rope = cos is not None
    lib().oracle_attention_prepare(_p(qkv), _p(queries), _p(keys), _p(values), _p(cos), _p(sin),
                                   C.c_int(num_q_heads), C.c_int(num_kv_heads), C.c_int(head_dim),
                                   C.c_int(rope_dim or 0), C.c_int(kv_token_offset), C.c_int(m),
                                   C.c_int(int(has_kv)), C.c_int(int(has_rope)))
    return queries


class _Mask(C.Structure):
    _fields_ = [("has_ring", C.c_int), ("ring_offset", C.c_uint32), ("ring_length", C.c_uint32),
                ("trie", C.c_void_p), ("has_sliding_window", C.c_int), ("sliding_window_size", C.c_uint32),
                ("is_causal", C.c_int)]


class _AttnArgs(C.Structure):
    _fields_ = [("queries", C.c_void_p), ("keys", C.c_void_p), ("values", C.c_void_p), ("head_dim", C.c_int),
                ("gqa_factor", C.c_int), ("sequence_length", C.c_int), ("k_head_stride", C.c_int),
                ("k_seq_stride", C.c_int), ("v_head_stride", C.c_int), ("v_seq_stride", C.c_int),
                ("scale", C.c_float), ("sinks", C.c_void_p), ("num_heads", C.c_int), ("suffix_length", C.c_int),
                ("mask", _Mask)]


def _attn_args(queries, keys, values, *, head_dim, gqa_factor, sequence_length, k_head_stride, k_seq_stride,
               v_head_stride, v_seq_stride, scale, num_heads, suffix_length, is_causal=True, sinks=None, ring=None,
               sliding_window=None, trie=None):
    mask = _Mask(has_ring=int(ring is not None), ring_offset=ring[0] if ring else 0,
                 ring_length=ring[1] if ring else 0, trie=_p(trie).value if trie is not None else None,
                 has_sliding_window=int(sliding_window is not None), sliding_window_size=sliding_window or 0,
                 is_causal=int(is_causal))
    return _AttnArgs(queries=_p(queries).value, keys=_p(keys).value, values=_p(values).value, head_dim=head_dim,
                     gqa_factor=gqa_factor, sequence_length=sequence_length, k_head_stride=k_head_stride,
                     k_seq_stride=k_seq_stride, v_head_stride=v_head_stride, v_seq_stride=v_seq_stride, scale=scale,
                     sinks=_p(sinks).value if sinks is not None else None, num_heads=num_heads,
                     suffix_length=suffix_length, mask=mask)


def attention_single_pass(queries, keys, values, **kw):
    a = _attn_args(queries, keys, values, **kw)
    out = np.zeros((a.suffix_length, a.num_heads, a.head_dim), dtype=np.uint16)
    lib().oracle_attention_single_pass(C.byref(a), _p(out))
    return out


def attention_two_pass(queries, keys, values, return_partials=False, **kw):
    a = _attn_args(queries, keys, values, **kw)
    S, H, D = a.suffix_length, a.num_heads, a.head_dim
    partials = np.zeros((S, H, TWO_PASS_BLOCKS, D), dtype=np.float32)
    sums = np.zeros((S, H, TWO_PASS_BLOCKS), dtype=np.float32)
    maxs = np.zeros((S, H, TWO_PASS_BLOCKS), dtype=np.float32)
    lib().oracle_attention_two_pass1(C.byref(a), _p(partials), _p(sums), _p(maxs))
    out = np.zeros((S, H, D), dtype=np.uint16)
    lib().oracle_attention_two_pass2(_p(partials), _p(sums), _p(maxs), _p(out), C.c_int(D), C.c_int(H), C.c_int(S))
    if return_partials:
        return out, partials, sums, maxs
    return out


def kv_cache_update(keys, values, copies, element_dim):
    copies = np.ascontiguousarray(copies, dtype=np.uint32).reshape(-1, 2)
    lib().oracle_kv_cache_update(_p(keys), _p(values), _p(copies), C.c_int(len(copies)), C.c_int(element_dim))


def sigmoid_gate(gate, output):
    lib().oracle_sigmoid_gate(_p(gate), _p(output), C.c_int(output.size))
    return output


def gated_act_mul(fused_up, gated_dim, act=ACT_SILU):
    rows = fused_up.shape[0]
    out = np.zeros((rows, gated_dim), dtype=np.uint16)
    lib().oracle_gated_act_mul(_p(fused_up), _p(out), C.c_int(gated_dim), C.c_int(rows), C.c_int(act))
    return out


def quant_embedding_lookup(token_ids, weights, scales, *, zero_points=None, biases=None, vocab_size, model_dim,
                           input_scale=1.0, group_size=64, mode=MODE_U4, method=QM_ZERO_POINT):
    token_ids = np.ascontiguousarray(token_ids, dtype=np.uint32)
    out = np.zeros((len(token_ids), model_dim), dtype=np.uint16)
    lib().oracle_quant_embedding_lookup(_p(token_ids), _p(weights), _p(scales), _p(zero_points), _p(biases), _p(out),
                                        C.c_int(len(token_ids)), C.c_uint32(vocab_size), C.c_int(model_dim),
                                        C.c_float(input_scale), C.c_int(group_size), C.c_int(mode), C.c_int(method))
    return out


def fp_embedding_lookup(token_ids, weights, *, vocab_size, model_dim, input_scale=1.0):
    token_ids = np.ascontiguousarray(token_ids, dtype=np.uint32)
    out = np.zeros((len(token_ids), model_dim), dtype=np.uint16)
    lib().oracle_fp_embedding_lookup(_p(token_ids), _p(weights), _p(out), C.c_int(len(token_ids)),
                                     C.c_uint32(vocab_size), C.c_int(model_dim), C.c_float(input_scale))
    return out


def logit_transform(logits, scale, soft_cap=None):
    lib().oracle_logit_transform(_p(logits), C.c_int(logits.size), C.c_float(scale), C.c_float(soft_cap or 0.0),
                                 C.c_int(int(soft_cap is not None)))
    return logits


def tensor_add_scale(inp, bias, num_cols, scale):
    out = np.zeros_like(inp)
    lib().oracle_tensor_add_scale(_p(inp), _p(bias), _p(out), C.c_int(num_cols), C.c_int(inp.size), C.c_float(scale))
    return out


def tensor_add_bias(inp, bias, num_cols):
    out = np.zeros_like(inp)
    lib().oracle_tensor_add_bias(_p(inp), _p(bias), _p(out), C.c_int(num_cols), C.c_int(inp.size))
    return out


RHT_INPUT, RHT_OUTPUT, RHT_QUANTIZE, RHT_QUANTIZE_WITH_GROUP_SUMS = 0, 1, 2, 3   # gpu_types ActivationTransformOp
HADAMARD_TRANSFORM_BLOCK_SIZE = 32


def activation_transform(x, factors, *, op=RHT_INPUT, in_place=False, activation_group_size=0, sum_group_size=0):
    """ActivationTransformKernel (cpu/kernel/activation_transform/activation_transform.rs:44-136). x: [rows, cols] bf16 bits (uint16) or
    float32. InputRht / OutputRht return the transformed array (x itself when in_place); the quantize ops return (codes i8, scales f32
    [rows, cols/activation_group_size], group sums i32 [rows, cols/sum_group_size] or None)."""
    x = np.ascontiguousarray(x) if not in_place else x
    rows, cols = x.shape
    assert cols % HADAMARD_TRANSFORM_BLOCK_SIZE == 0
    factors = np.ascontiguousarray(factors, dtype=np.int32)
    assert factors.shape == (cols,)
    is_f32 = int(x.dtype == np.float32)
    if op in (RHT_INPUT, RHT_OUTPUT):
        out = x if in_place else np.zeros_like(x)
        lib().oracle_activation_transform(_p(x), C.c_int(is_f32), _p(out), None, None, None, _p(factors), C.c_int(rows), C.c_int(cols),
                                          C.c_int(op), C.c_int(0), C.c_int(0))
        return out
    q = np.zeros((rows, cols), dtype=np.int8)
    sc = np.zeros((rows, cols // activation_group_size), dtype=np.float32)
    gs = np.zeros((rows, cols // sum_group_size), dtype=np.int32) if op == RHT_QUANTIZE_WITH_GROUP_SUMS else None
    lib().oracle_activation_transform(_p(x), C.c_int(is_f32), None, _p(q), _p(sc), _p(gs) if gs is not None else None, _p(factors),
                                      C.c_int(rows), C.c_int(cols), C.c_int(op), C.c_int(activation_group_size), C.c_int(sum_group_size))
    return q, sc, gs


def tensor_add_swap(skip, main):
    lib().oracle_tensor_add_swap(_p(skip), _p(main), C.c_int(skip.size))


def philox4x32_10(ctr, key):
    ctr = np.ascontiguousarray(ctr, dtype=np.uint32)
    key = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().oracle_philox4x32_10(_p(ctr), _p(key), _p(out))
    return out


def revidx(i, vocab):
    off, w = C.c_uint32(), C.c_uint32()
    lib().oracle_revidx(C.c_uint32(i), C.c_uint32(vocab), C.byref(off), C.byref(w))
    return off.value, w.value


class _SamplingArgs(C.Structure):
    _fields_ = [("logits", C.c_void_p), ("output", C.c_void_p), ("seeds", C.c_void_p), ("bitmask", C.c_void_p),
                ("has_temperature", C.c_int), ("temperature", C.c_float), ("has_top_k", C.c_int),
                ("top_k", C.c_uint32), ("has_top_p", C.c_int), ("top_p", C.c_float), ("has_min_p", C.c_int),
                ("min_p", C.c_float), ("vocab_size", C.c_uint32), ("batch_size", C.c_uint32)]


def unified_sampling(logits, *, seeds=None, bitmask=None, temperature=None, top_k=None, top_p=None, min_p=None):
    rows, V = logits.shape
    out = np.zeros(rows, dtype=np.uint32)
    if seeds is not None:
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
    a = _SamplingArgs(logits=_p(logits).value, output=_p(out).value,
                      seeds=_p(seeds).value if seeds is not None else None,
                      bitmask=_p(bitmask).value if bitmask is not None else None,
                      has_temperature=int(temperature is not None), temperature=temperature or 0.0,
                      has_top_k=int(top_k is not None), top_k=top_k or 0, has_top_p=int(top_p is not None),
                      top_p=top_p or 0.0, has_min_p=int(min_p is not None), min_p=min_p or 0.0, vocab_size=V,
                      batch_size=rows)
    lib().oracle_unified_sampling(C.byref(a))
    return out


def delta_net_conv_update(conv_weight, bias, in_out, state, kernel_size, conv_dim):
    lib().oracle_delta_net_conv_update(_p(conv_weight), _p(bias), _p(in_out), _p(state), C.c_int(kernel_size),
                                       C.c_int(conv_dim), C.c_int(kernel_size - 1))


def delta_net_update(in_proj, a_log, dt_bias, norm_weight, state, *, num_v_heads, num_k_heads, head_k_dim,
                     head_v_dim, key_dim, value_dim, norm_epsilon):
    out = np.zeros(value_dim, dtype=np.uint16)
    lib().oracle_delta_net_update(_p(in_proj), _p(a_log), _p(dt_bias), _p(norm_weight), _p(state), _p(out),
                                  C.c_int(num_v_heads), C.c_int(num_k_heads), C.c_int(head_k_dim),
                                  C.c_int(head_v_dim), C.c_int(key_dim), C.c_int(value_dim), C.c_float(norm_epsilon))
    return out
