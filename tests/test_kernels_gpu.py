"""Parity of every CUDA kernel (called through the C ABI) against the CPU oracle on the same seeded inputs.
Shapes follow the reference's own kernel tests (tests/unit/backends/common/kernel/**) plus the BASELINE
layer shapes. Integer / index outputs are compared bit-exactly; bf16 outputs of kernels whose f32 summation
order necessarily differs from the sequential CPU loop are held to <= 1 bf16 ulp with >= 97 % identical bits,
and to rtol 1e-3 / atol 1e-4 on f32 outputs (BASELINE.json north_star tolerance)."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import gpu_ops as G
from tests.test_oracle_pins import attention_inputs, random_quant
from tests.util import assert_bf16_close, assert_f32_close, bf16_to_f32, f32_to_bf16

pytestmark = pytest.mark.gpu

METHODS = {"zp": O.QM_ZERO_POINT, "mlx": O.QM_SCALE_BIAS, "sym": O.QM_SYMMETRIC}


def _quant_case(seed, m, n, k, bits, gs, method, scale_lo=0.01, scale_hi=0.3):
    rng = np.random.default_rng(seed)
    packed, scales, zp, biases = random_quant(rng, n, k, bits, gs, method)
    x = f32_to_bf16(rng.uniform(-0.3, 0.3, size=(m, k)).astype(np.float32))
    return x, packed, dict(scales=scales, zero_points=zp, biases=biases, method=method, bits=bits, group_size=gs)


# quant_dispatch_test.rs:102-167 matrix (shape x group x method) + decode shapes of the BASELINE models
QUANT_CASES = [
    (1, 64, 128, 4, 64, "zp"), (1, 48, 256, 4, 32, "mlx"), (1, 2048, 4096, 4, 64, "zp"), (4, 64, 4096, 4, 128, "sym"),
    (8, 2048, 128, 4, 64, "mlx"), (1, 11, 512, 8, 64, "zp"), (3, 100, 1024, 8, 32, "mlx"), (2, 64, 2048, 8, 128, "sym"),
    (1, 3072, 1024, 4, 64, "zp"),      # Qwen3.5-0.8B qkv
    (1, 1024, 3584, 4, 64, "zp"),      # Qwen3.5-0.8B down (split-K)
    (1, 4096, 14336, 4, 64, "zp"),     # Llama-3-8B down (split-K)
    (5, 1024, 2048, 4, 64, "zp"), (8, 256, 4096, 4, 64, "zp"), (16, 128, 1024, 4, 64, "zp"), (40, 96, 512, 4, 64, "zp"),
    (7, 80, 512, 8, 64, "zp"), (1, 8224, 1024, 4, 64, "zp"),
]


@pytest.mark.parametrize("m,n,k,bits,gs,method", QUANT_CASES)
def test_quantized_matmul_f32_out(ctx, m, n, k, bits, gs, method):
    x, w, kw = _quant_case(100 + n + k, m, n, k, bits, gs, METHODS[method])
    ref = O.matmul(x, w, m=m, n=n, k=k, d_f32=True, **kw)
    got = G.matmul(ctx, x, w, m=m, n=n, k=k, d_f32=True, **kw)
    assert_f32_close(got, ref, rtol=1e-3, atol=1e-4, what=f"matmul {m}x{n}x{k} int{bits} gs{gs} {method}")


# shapes the round-1 review found untested against the oracle: the readout matrices of the BASELINE models and int8 gs64 at Llama-3-8B K
BASELINE_SHAPES = [
    (1, 128256, 4096, 4, 64, "zp"),    # Llama-3-8B readout
    (1, 248320, 1024, 4, 64, "zp"),    # Qwen3.5-0.8B readout
    (1, 6144, 4096, 8, 64, "zp"),      # Llama-3-8B int8 qkv
    (1, 4096, 14336, 8, 64, "zp"),     # Llama-3-8B int8 down (split-K)
    (8, 4096, 4096, 8, 64, "zp"),      # config 4: eight sequences share the weight pass
]


@pytest.mark.parametrize("m,n,k,bits,gs,method", BASELINE_SHAPES)
def test_quantized_matmul_baseline_shapes(ctx, m, n, k, bits, gs, method):
    """Full matrices on the GPU; the oracle (scalar C) checks a seeded sample of 2048 output rows -- restating the loop on a row subset is the
    same function on those rows."""
    x, w, kw = _quant_case(300 + n % 1000 + k, m, n, k, bits, gs, METHODS[method])
    got = G.matmul(ctx, x, w, m=m, n=n, k=k, d_f32=True, **kw)
    rows = np.sort(np.random.default_rng(n).choice(n, size=min(n, 2048), replace=False))
    sub = dict(kw)
    sub["scales"] = np.ascontiguousarray(kw["scales"][rows])
    if kw.get("zero_points") is not None:
        sub["zero_points"] = np.ascontiguousarray(kw["zero_points"][rows])
    if kw.get("biases") is not None:
        sub["biases"] = np.ascontiguousarray(kw["biases"][rows])
    ref = O.matmul(x, np.ascontiguousarray(w[rows]), m=m, n=len(rows), k=k, d_f32=True, **sub)
    assert_f32_close(got[:, rows], ref, rtol=1e-3, atol=1e-4, what=f"matmul {m}x{n}x{k} int{bits} gs{gs} {method} (sampled rows)")


# rows kernel (qmv_rows_kernel, 2..16 activation rows per weight pass: speculation passes, multi-sequence decode): every quantisation
# method it instantiates, one and two MMA column blocks, k-slices through the split-k workspace, ragged n, several launches (m > 16)
ROWS_CASES = [
    (2, 64, 512, 4, 64, "zp"), (3, 4096, 4096, 4, 64, "zp"), (8, 6144, 4096, 4, 64, "zp"), (9, 4096, 4096, 4, 64, "zp"),
    (16, 4096, 14336, 4, 64, "zp"), (16, 28672, 4096, 4, 64, "zp"), (5, 1000, 1024, 4, 64, "zp"), (16, 72, 2048, 4, 64, "mlx"),
    (12, 512, 2048, 4, 128, "sym"), (7, 1024, 1536, 4, 128, "zp"), (8, 4096, 4096, 8, 64, "zp"), (16, 1024, 14336, 8, 64, "zp"),
    (4, 256, 1024, 8, 64, "mlx"), (13, 320, 2048, 8, 64, "sym"), (40, 96, 512, 4, 64, "zp"), (33, 2048, 1024, 4, 64, "zp"),
]


@pytest.mark.parametrize("m,n,k,bits,gs,method", ROWS_CASES)
def test_rows_kernel_matches_oracle(ctx, m, n, k, bits, gs, method):
    x, w, kw = _quant_case(500 + n + k + m, m, n, k, bits, gs, METHODS[method])
    cap = 1536                                    # oracle on a row sample for the big shapes (same function on those rows)
    rows = np.arange(n) if n <= cap else np.sort(np.random.default_rng(n + m).choice(n, size=cap, replace=False))
    sub = dict(kw)
    sub["scales"] = np.ascontiguousarray(kw["scales"][rows])
    for key in ("zero_points", "biases"):
        if kw.get(key) is not None:
            sub[key] = np.ascontiguousarray(kw[key][rows])
    got, launches = G.matmul(ctx, x, w, m=m, n=n, k=k, d_f32=True, return_launches=True, **kw)
    assert launches == (m + 15) // 16, "one launch per 16 activation rows"
    ref = O.matmul(x, np.ascontiguousarray(w[rows]), m=m, n=len(rows), k=k, d_f32=True, threads=8, **sub)
    assert_f32_close(got[:, rows], ref, rtol=1e-3, atol=1e-4, what=f"rows kernel {m}x{n}x{k} int{bits} gs{gs} {method}")
    # every row of the batch is what the m = 1 decode kernel computes for it
    one = G.matmul(ctx, x[m - 1:m], w, m=1, n=n, k=k, d_f32=True, **kw)
    assert_f32_close(got[m - 1:m], one, rtol=1e-3, atol=1e-4, what="rows kernel vs decode kernel")


def test_rows_kernel_epilogue_variants(ctx):
    m, n, k = 11, 1024, 2048
    x, w, kw = _quant_case(77, m, n, k, 4, 64, METHODS["zp"])
    rng = np.random.default_rng(5)
    bias = f32_to_bf16(rng.uniform(-0.5, 0.5, n).astype(np.float32))
    d0 = f32_to_bf16(rng.uniform(-1, 1, (m, n)).astype(np.float32))
    for extra in (dict(bias=bias), dict(ab_scale=0.37), dict(soft_cap=20.0), dict(accumulate=True, d=d0.copy()), dict(bias=bias, ab_scale=1.7, soft_cap=30.0)):
        ref = O.matmul(x, w, m=m, n=n, k=k, threads=8, **kw, **{k_: (v.copy() if isinstance(v, np.ndarray) else v) for k_, v in extra.items()})
        got = G.matmul(ctx, x, w, m=m, n=n, k=k, **kw, **{k_: (v.copy() if isinstance(v, np.ndarray) else v) for k_, v in extra.items()})
        assert_bf16_close(got, ref, what=f"rows kernel epilogue {sorted(extra)}")
    sx, sw, skw = _quant_case(78, 6, 512, 1024, 4, 64, METHODS["zp"])
    ref = O.matmul(sx, sw, m=6, n=512, k=1024, signed_codes=True, threads=4, **skw)
    got = G.matmul(ctx, sx, sw, m=6, n=512, k=1024, signed_codes=True, **skw)
    assert_bf16_close(got, ref, what="rows kernel signed codes")


@pytest.mark.parametrize("m,n,k,bits,gs,method", QUANT_CASES[:9])
def test_quantized_matmul_bf16_out(ctx, m, n, k, bits, gs, method):
    x, w, kw = _quant_case(200 + n + k, m, n, k, bits, gs, METHODS[method])
    ref = O.matmul(x, w, m=m, n=n, k=k, **kw)
    got = G.matmul(ctx, x, w, m=m, n=n, k=k, **kw)
    assert_bf16_close(got, ref, max_ulp=1, min_exact=0.97, what="matmul bf16")


def test_matmul_epilogue_signed_codes_and_generic(ctx):
    m, n, k = 2, 80, 256
    x, w, kw = _quant_case(7, m, n, k, 4, 64, O.QM_ZERO_POINT)
    rng = np.random.default_rng(8)
    bias = f32_to_bf16(rng.uniform(-1, 1, n).astype(np.float32))
    d0 = rng.uniform(-1, 1, (m, n)).astype(np.float32)
    ep = dict(ab_scale=1.7, accumulate=True, bias=bias, soft_cap=2.5)
    ref = O.matmul(x, w, m=m, n=n, k=k, d=d0.copy(), **kw, **ep)
    got = G.matmul(ctx, x, w, m=m, n=n, k=k, d=d0.copy(), **kw, **ep)
    assert_f32_close(got, ref, what="epilogue")
    # signed codes: flipped storage + signed_codes flag == original
    a = G.matmul(ctx, x, w, m=m, n=n, k=k, d_f32=True, **kw)
    b = G.matmul(ctx, x, w ^ np.uint8(0x88), m=m, n=n, k=k, d_f32=True, signed_codes=True, **kw)
    assert (a == b).all()
    # gather (sparse readout) == dense columns, generic kernel path
    gather = rng.integers(0, n, size=(m, 9)).astype(np.uint32)
    sp = G.matmul(ctx, x, w, m=m, n=9, k=k, gather=gather, **kw)
    dense = G.matmul(ctx, x, w, m=m, n=n, k=k, **kw)
    ref_sp = O.matmul(x, w, m=m, n=9, k=k, gather=gather, **kw)
    assert_bf16_close(sp, ref_sp, what="gather vs oracle")
    for r in range(m):
        assert_bf16_close(sp[r], dense[r][gather[r]], min_exact=0.9, what="gather vs dense")


@pytest.mark.parametrize("m,k,n", [(1, 33, 3), (4, 128, 11), (8, 4096, 64), (1, 128, 2048)])
def test_full_precision_matmul(ctx, m, k, n):
    # gemv_test.rs:38-58,135-167 closed-form inputs
    a = f32_to_bf16(((np.arange(m * k) % 13) * 0.1 - 0.6).astype(np.float32)).reshape(m, k)
    b = f32_to_bf16(((np.arange(n * k) % 17) * 0.1 - 0.8).astype(np.float32)).reshape(n, k)
    ref = O.matmul(a, b, m=m, n=n, k=k, method=O.QM_NONE, d_f32=True)
    got = G.matmul(ctx, a, b, m=m, n=n, k=k, method=O.QM_NONE, d_f32=True)
    assert_f32_close(got, ref, rtol=1e-3, atol=1e-3, what="fp matmul")


def test_matmul_invalid_arguments_are_sticky_errors(ctx):
    from uzu_b200 import binding as B
    x, w, kw = _quant_case(9, 1, 16, 128, 4, 64, O.QM_ZERO_POINT)
    kw["scales"] = None
    with pytest.raises(B.UzuError, match="scales"):
        G.matmul(ctx, x, w, m=1, n=16, k=128, **kw)


def test_matmul_linearity_full_size(ctx):
    """Size-independent property at a BASELINE shape (Llama-3-8B up projection 28672 x 4096 int4):
    f32 outputs are linear in the activation, A(x1 + x2) == A x1 + A x2 up to f32 rounding, and the
    zero vector maps to exactly zero."""
    n, k = 28672, 4096
    rng = np.random.default_rng(10)
    packed, scales, zp, _ = random_quant(rng, n, k, 4, 64, O.QM_ZERO_POINT)
    kw = dict(scales=scales, zero_points=zp, method=O.QM_ZERO_POINT, bits=4, group_size=64)
    x1 = f32_to_bf16((rng.integers(-8, 8, (1, k)) / 16).astype(np.float32))      # sums stay exact in bf16
    x2 = f32_to_bf16((rng.integers(-8, 8, (1, k)) / 16).astype(np.float32))
    x12 = f32_to_bf16(bf16_to_f32(x1) + bf16_to_f32(x2))
    y1 = G.matmul(ctx, x1, packed, m=1, n=n, k=k, d_f32=True, **kw)
    y2 = G.matmul(ctx, x2, packed, m=1, n=n, k=k, d_f32=True, **kw)
    y12 = G.matmul(ctx, x12, packed, m=1, n=n, k=k, d_f32=True, **kw)
    assert_f32_close(y12, y1 + y2, rtol=1e-4, atol=1e-3, what="linearity")
    y0 = G.matmul(ctx, np.zeros((1, k), np.uint16), packed, m=1, n=n, k=k, d_f32=True, **kw)
    assert (y0 == 0).all()
    # a sample of rows against the oracle
    rows = rng.choice(n, 64, replace=False)
    ref = O.matmul(x1, np.ascontiguousarray(packed[rows]), m=1, n=64, k=k, d_f32=True, scales=np.ascontiguousarray(scales[rows]),
                   zero_points=np.ascontiguousarray(zp[rows]), method=O.QM_ZERO_POINT, bits=4, group_size=64)
    assert_f32_close(y1[:, rows], ref, what="sampled rows")


# ---- normalisation -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,n,full_layer,offset", [(3, 256, False, 0.0), (1, 4096, False, 0.0), (2, 1024, True, 1.0), (5, 8192, False, 1.0)])
def test_normalization(ctx, rows, n, full_layer, offset):
    x = f32_to_bf16((0.5 + np.arange(rows * n, dtype=np.float32) * np.float32(0.01)).reshape(rows, n) % np.float32(7.0))
    scales = (1.0 - offset + 0.001 * np.arange(n)).astype(np.float32)
    for mode in ("none", "copy", "add"):
        sc = f32_to_bf16(np.sin(np.arange(rows * n, dtype=np.float32)).reshape(rows, n)) if mode != "none" else None
        sc_ref = sc.copy() if sc is not None else None
        ref = O.normalization(x, scales, shortcut=sc_ref, residual_add=mode == "add", epsilon=1e-5, scale_offset=offset, full_layer=full_layer)
        got, sc_got = G.normalization(ctx, x, scales, shortcut=sc, residual_add=mode == "add", epsilon=1e-5, scale_offset=offset, full_layer=full_layer)
        # OnlyNormalization multiplies two bf16-rounded factors: a 1-ulp flip of bf16(norm) (rms_inv differs in its last f32
        # bit between the sequential and the tree sum) can land 2 ulp away after the product is rounded again
        assert_bf16_close(got, ref, max_ulp=1 if full_layer else 2, min_exact=0.99, what=f"norm {mode}")
        if sc is not None:
            assert (sc_got == sc_ref).all(), "shortcut (bf16 residual) must be bit-exact"


def test_qkv_norm(ctx):
    rng = np.random.default_rng(11)
    rows, Hq, Hkv, D = 3, 8, 2, 128
    qkv = f32_to_bf16(rng.standard_normal((rows, (Hq + 2 * Hkv) * D)).astype(np.float32))
    scales = (rng.standard_normal(D) * 0.05).astype(np.float32)
    for off, cnt, full in [(0, Hq, False), (Hq, Hkv, False), (0, Hq, True)]:
        kw = dict(total_heads=Hq + 2 * Hkv, head_dim=D, epsilon=1e-6, scale_offset=1.0, head_offset=off, head_count=cnt, full_layer=full)
        ref = O.qkv_norm(qkv.copy(), scales, **kw)
        got = G.qkv_norm(ctx, qkv, scales, **kw)
        assert_bf16_close(got, ref, max_ulp=1, min_exact=0.99, what="qkv_norm")


# ---- attention -------------------------------------------------------------------------------------------------
def test_attention_prepare_bit_exact(ctx):
    rng = np.random.default_rng(12)
    for Hq, Hkv, D, rope_dim, m, off in [(4, 2, 128, 128, 3, 5), (8, 2, 256, 64, 1, 17), (4, 4, 64, None, 2, 0)]:
        qkv = f32_to_bf16(rng.standard_normal((m, (Hq + 2 * Hkv) * D)).astype(np.float32))
        keys = f32_to_bf16(rng.standard_normal((32, Hkv * D)).astype(np.float32))
        values = f32_to_bf16(rng.standard_normal((32, Hkv * D)).astype(np.float32))
        cos = sin = None
        if rope_dim:
            cos, sin = O.rope_tables({"type": "UnscaledRoPEConfig", "base": 10000.0, "head_dim": rope_dim}, np.arange(off, off + m))
        kr, vr = keys.copy(), values.copy()
        q_ref = O.attention_prepare(qkv, kr, vr, cos, sin, num_q_heads=Hq, num_kv_heads=Hkv, head_dim=D, rope_dim=rope_dim, kv_token_offset=off)
        q, k, v = G.attention_prepare(ctx, qkv, keys, values, cos, sin, num_q_heads=Hq, num_kv_heads=Hkv, head_dim=D, rope_dim=rope_dim, kv_token_offset=off)
        assert (q == q_ref).all() and (k == kr).all() and (v == vr).all()   # same f32 ops, same rounding -> bit-exact


ATTN_CASES = [  # H, Hkv, seq, suffix, D, causal
    (4, 4, 16, 1, 64, False), (8, 2, 40, 4, 128, True), (4, 1, 64, 8, 64, True), (32, 8, 700, 1, 128, True),
    (8, 2, 1500, 1, 256, True), (64, 8, 300, 2, 128, True), (6, 2, 90, 3, 128, True), (3, 1, 50, 1, 64, True),
    (32, 8, 2304, 1, 128, True),      # Llama-3-8B decode at the end of BASELINE config 3 (prefill 2048 + decode 256)
]


@pytest.mark.parametrize("H,Hkv,seq,suffix,D,causal", ATTN_CASES)
def test_attention_single_and_two_pass(ctx, H, Hkv, seq, suffix, D, causal):
    q, k, v = attention_inputs(H, Hkv, seq, suffix, D)   # attention_single_pass_test.rs:33-78 inputs
    kw = dict(head_dim=D, gqa_factor=H // Hkv, sequence_length=seq, k_head_stride=seq * D, k_seq_stride=D, v_head_stride=seq * D,
              v_seq_stride=D, scale=float(np.float32(1.0) / np.sqrt(np.float32(D))), num_heads=H, suffix_length=suffix, is_causal=causal)
    ref = O.attention_single_pass(q, k, v, **kw)
    got1 = G.attention_single_pass(ctx, q, k, v, **kw)
    assert_bf16_close(got1, ref, max_ulp=1, min_exact=0.95, what="single pass")
    got2 = G.attention_two_pass(ctx, q, k, v, **kw)
    assert_bf16_close(got2, O.attention_two_pass(q, k, v, **kw), max_ulp=1, min_exact=0.95, what="two pass")


def test_attention_token_major_cache_ring_window_sinks(ctx):
    rng = np.random.default_rng(13)
    H, Hkv, D, prefix, suffix = 4, 2, 64, 24, 2
    q = f32_to_bf16(rng.standard_normal((H, suffix, D)).astype(np.float32))
    k = f32_to_bf16(rng.standard_normal((prefix + suffix, Hkv * D)).astype(np.float32))
    v = f32_to_bf16(rng.standard_normal((prefix + suffix, Hkv * D)).astype(np.float32))
    sinks = f32_to_bf16(rng.standard_normal(H).astype(np.float32))
    base = dict(head_dim=D, gqa_factor=2, sequence_length=prefix + suffix, k_head_stride=D, k_seq_stride=Hkv * D, v_head_stride=D,
                v_seq_stride=Hkv * D, scale=0.125, num_heads=H, suffix_length=suffix)
    for extra in [dict(is_causal=True), dict(is_causal=True, ring=(5, 17), sliding_window=8), dict(is_causal=False, sliding_window=10),
                  dict(is_causal=True, sinks=sinks)]:
        ref = O.attention_single_pass(q, k, v, **base, **extra)
        got = G.attention_single_pass(ctx, q, k, v, **base, **extra)
        assert_bf16_close(got, ref, max_ulp=1, min_exact=0.95, what=f"attention {sorted(extra)}")
    trie = np.array([[0, 1, 0], [1, 1, 1]], np.uint32)   # flat 2-token chain as trie nodes
    ref = O.attention_single_pass(q, k, v, **base, is_causal=True, trie=trie)
    got = G.attention_single_pass(ctx, q, k, v, **base, is_causal=True, trie=trie)
    assert_bf16_close(got, ref, max_ulp=1, min_exact=0.95, what="trie")


def test_kv_cache_update_and_sigmoid_gate(ctx):
    rng = np.random.default_rng(14)
    keys = rng.integers(0, 65535, (10, 96)).astype(np.uint16); values = rng.integers(0, 65535, (10, 96)).astype(np.uint16)
    copies = [(7, 2), (2, 3), (9, 0)]     # chained: row 3 must receive the *updated* row 2
    kr, vr = keys.copy(), values.copy()
    O.kv_cache_update(kr, vr, copies, 96)
    k, v = G.kv_cache_update(ctx, keys, values, copies, 96)
    assert (k == kr).all() and (v == vr).all()
    gate = f32_to_bf16(rng.standard_normal(1000).astype(np.float32) * 3); out = f32_to_bf16(rng.standard_normal(1000).astype(np.float32))
    ref = O.sigmoid_gate(gate, out.copy())
    assert_bf16_close(G.sigmoid_gate(ctx, gate, out), ref, max_ulp=1, min_exact=0.995, what="sigmoid gate")


# ---- elementwise / embedding -----------------------------------------------------------------------------------
def test_gated_act_mul(ctx):
    rng = np.random.default_rng(15)
    up = f32_to_bf16(rng.standard_normal((3, 2 * 3584)).astype(np.float32) * 3)
    # SiLU: expf differs by <= 1 f32 ulp between CUDA and glibc. GELU: 1 + tanh / 1 + erf cancel for negative gates, so the
    # libm difference is amplified before the bf16 rounding (transcendental parity is unpinned in the reference)
    for act, min_exact in ((O.ACT_SILU, 0.995), (O.ACT_GELU_APPROX, 0.95), (O.ACT_GELU_EXACT, 0.95)):
        assert_bf16_close(G.gated_act_mul(ctx, up, 3584, act), O.gated_act_mul(up, 3584, act), max_ulp=1, min_exact=min_exact,
                          what=f"gated_act_mul act={act}")


def test_embedding_lookups_bit_exact(ctx):
    rng = np.random.default_rng(16)
    V, H = 300, 256
    toks = np.array([3, 299, 300, 0, 17], np.uint32)
    for bits, method in [(4, O.QM_ZERO_POINT), (8, O.QM_SCALE_BIAS), (4, O.QM_SYMMETRIC), (8, O.QM_ZERO_POINT)]:
        packed, scales, zp, biases = random_quant(rng, V, H, bits, 64, method)
        kw = dict(zero_points=zp, biases=biases, vocab_size=V, model_dim=H, input_scale=1.5, group_size=64,
                  mode=O.MODE_U4 if bits == 4 else O.MODE_U8, method=method)
        assert (G.quant_embedding_lookup(ctx, toks, packed, scales, **kw) == O.quant_embedding_lookup(toks, packed, scales, **kw)).all()
    w = f32_to_bf16(rng.standard_normal((V, H)).astype(np.float32))
    assert (G.fp_embedding_lookup(ctx, toks, w, vocab_size=V, model_dim=H, input_scale=0.5) ==
            O.fp_embedding_lookup(toks, w, vocab_size=V, model_dim=H, input_scale=0.5)).all()


def test_elementwise_glue(ctx):
    rng = np.random.default_rng(17)
    x = f32_to_bf16(rng.standard_normal((4, 100)).astype(np.float32) * 5)
    bias = f32_to_bf16(rng.standard_normal(100).astype(np.float32))
    assert_bf16_close(G.logit_transform(ctx, x, 0.7, 3.0), O.logit_transform(x.copy(), 0.7, 3.0), min_exact=0.995, what="logit_transform")
    assert (G.logit_transform(ctx, x, 0.5) == O.logit_transform(x.copy(), 0.5)).all()
    assert (G.tensor_add_scale(ctx, x, bias, 100, 0.25) == O.tensor_add_scale(x, bias, 100, 0.25)).all()
    assert (G.tensor_add_bias(ctx, x, bias, 100) == O.tensor_add_bias(x, bias, 100)).all()
    assert (G.tensor_copy(ctx, x) == x).all()
    s, m_ = x.copy(), f32_to_bf16(rng.standard_normal((4, 100)).astype(np.float32))
    sr, mr = s.copy(), m_.copy()
    O.tensor_add_swap(sr, mr)
    gs, gm = G.tensor_add_swap(ctx, s, m_)
    assert (gs == sr).all() and (gm == mr).all()


# ---- sampling (token ids are bit-exact) -----------------------------------------------------------------------
@pytest.mark.parametrize("V,rows", [(1000, 4), (128256, 2), (248320, 1), (5000, 300)])
def test_sampling_greedy_bit_exact(ctx, V, rows):
    rng = np.random.default_rng(18)
    logits = f32_to_bf16(rng.standard_normal((rows, V)).astype(np.float32) * 2)
    logits[0, V // 3] = logits[0, V - 1] = f32_to_bf16(np.array([50.0], np.float32))[0]   # tie -> lowest index
    assert (G.unified_sampling(ctx, logits) == O.unified_sampling(logits)).all()
    bm = rng.integers(0, 2**32, (rows, (V + 31) // 32), dtype=np.uint64).astype(np.uint32)
    assert (G.unified_sampling(ctx, logits, bitmask=bm, temperature=0.8) == O.unified_sampling(logits, bitmask=bm, temperature=0.8)).all()


def test_sampling_stochastic_bit_exact(ctx):
    rng = np.random.default_rng(19)
    for V in (1000, 32000, 128256):
        rows = 6
        logits = f32_to_bf16(rng.standard_normal((rows, V)).astype(np.float32) * 2)
        seeds = rng.integers(0, 2**63, rows).astype(np.uint64)
        assert (G.unified_sampling(ctx, logits, seeds=seeds) == O.unified_sampling(logits, seeds=seeds)).all()
        assert (G.unified_sampling(ctx, logits, seeds=seeds, temperature=0.6) == O.unified_sampling(logits, seeds=seeds, temperature=0.6)).all()


def test_sampling_filters_bit_exact(ctx):
    rng = np.random.default_rng(20)
    V, rows = 4096, 8
    logits = f32_to_bf16(rng.standard_normal((rows, V)).astype(np.float32) * 3)
    seeds = rng.integers(0, 2**63, rows).astype(np.uint64)
    for kw in [dict(top_k=1), dict(top_k=40), dict(top_p=0.9), dict(min_p=0.05), dict(top_k=50, top_p=0.95, min_p=0.02, temperature=0.7),
               dict(top_k=0)]:
        assert (G.unified_sampling(ctx, logits, seeds=seeds, **kw) == O.unified_sampling(logits, seeds=seeds, **kw)).all(), kw
        assert (G.unified_sampling(ctx, logits, **kw) == O.unified_sampling(logits, **kw)).all(), kw


# ---- DeltaNet decode ---------------------------------------------------------------------------------------------
def test_delta_net_kernels(ctx):
    rng = np.random.default_rng(21)
    Hv, Hk, Dk, Dv, ks = 16, 16, 128, 128, 4
    kd, vd = Hk * Dk, Hv * Dv
    conv_dim = 2 * kd + vd
    total = conv_dim + vd + 2 * Hv
    x = f32_to_bf16(rng.standard_normal(total).astype(np.float32))
    cw = (rng.standard_normal((conv_dim, ks)) * 0.5).astype(np.float32)
    cstate = rng.standard_normal((conv_dim, ks - 1)).astype(np.float32)
    xr, sr = x.copy(), cstate.copy()
    O.delta_net_conv_update(cw, None, xr, sr, ks, conv_dim)
    xg, sg = G.delta_net_conv_update(ctx, cw, None, x, cstate, ks, conv_dim)
    assert (sg == sr).all()
    assert_bf16_close(xg, xr, max_ulp=1, min_exact=0.995, what="conv_update")
    a_log = rng.uniform(-1, 1, Hv).astype(np.float32); dt = rng.uniform(-1, 1, Hv).astype(np.float32)
    nw = (1 + 0.1 * rng.standard_normal(Dv)).astype(np.float32)
    state = (rng.standard_normal((Hv, Dv, Dk)) * 0.1).astype(np.float32)
    st_ref = state.copy()
    kw = dict(num_v_heads=Hv, num_k_heads=Hk, head_k_dim=Dk, head_v_dim=Dv, key_dim=kd, value_dim=vd, norm_epsilon=1e-6)
    out_ref = O.delta_net_update(xr, a_log, dt, nw, st_ref, **kw)
    out, st = G.delta_net_update(ctx, xr, a_log, dt, nw, state, **kw)
    assert_f32_close(st, st_ref, rtol=1e-4, atol=1e-5, what="ssm state")
    assert_bf16_close(out, out_ref, max_ulp=1, min_exact=0.97, what="delta_net_update")


# ---- runtime ---------------------------------------------------------------------------------------------------------
def test_runtime_buffers_copy_fill_sparse(ctx):
    import ctypes as C
    from uzu_b200 import binding as B
    src = ctx.upload(np.arange(1000, dtype=np.uint32))
    dst = ctx.buffer(4000)
    with ctx.command_buffer("copyfill") as cmd:
        cmd.encode("uzu_command_buffer_encode_copy", src.ptr, dst.ptr, 4000)
        cmd.encode("uzu_command_buffer_encode_fill", dst.ptr, 400, 0)
    out = dst.numpy(np.uint32)
    assert (out[:100] == 0).all() and (out[100:] == np.arange(100, 1000)).all()
    assert cmd.gpu_seconds >= 0
    pinned = ctx.buffer(64, B.BUFFER_PINNED_HOST)
    assert pinned.cpu_ptr and pinned.ptr
    if ctx.capabilities() & 1:
        h = C.c_void_p()
        assert ctx.lib.uzu_sparse_buffer_create(ctx.h, 64 << 20, C.byref(h)) == 0
        page = ctx.lib.uzu_sparse_buffer_page_size_bytes(h)
        pages = (B.u32 * 2)(0, 3)
        assert ctx.lib.uzu_sparse_buffer_map(h, pages, 2) == 0
        base = ctx.lib.uzu_sparse_buffer_gpu_ptr(h)
        with ctx.command_buffer("sparse") as cmd:
            cmd.encode("uzu_command_buffer_encode_fill", base + 3 * page, 4096, 0xAB)
            cmd.encode("uzu_command_buffer_encode_copy", base + 3 * page, dst.ptr, 4000)
        assert (dst.numpy(np.uint8) == 0xAB).all()
        assert ctx.lib.uzu_sparse_buffer_unmap(h, pages, 2) == 0
        ctx.lib.uzu_sparse_buffer_destroy(h)
