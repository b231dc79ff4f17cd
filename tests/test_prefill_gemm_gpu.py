"""Parity of the tcgen05 prefill GEMM (uzu_b200/csrc/prefill_gemm.cu, reached through uzu_matmul_encode for m >= 64) against the
CPU oracle's MatmulKernel restatement (backends/cpu/kernel/matmul/kernel.rs:164-295). Shapes: the reference's quant dispatch matrix
(quant_dispatch_test.rs:102-167: bits x group x method) at prefill row counts, ragged m / n tails, and the BASELINE layer shapes
through size-independent properties. Tolerance: BASELINE rtol 1e-3 / atol 1e-4 on f32 outputs, <= 1 bf16 ulp on bf16 outputs."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import gpu_ops as G
from tests.test_kernels_gpu import METHODS, _quant_case
from tests.test_oracle_pins import random_quant
from tests.util import assert_bf16_close, assert_f32_close, bf16_to_f32, f32_to_bf16

pytestmark = pytest.mark.gpu

# (m, n, k, bits, group, method)
CASES = [
    (64, 128, 64, 4, 64, "zp"),          # one K block, one tile
    (128, 128, 256, 4, 64, "zp"),
    (256, 256, 512, 4, 64, "zp"),        # two token sub-tiles
    (200, 136, 320, 4, 64, "zp"),        # ragged m and n (n % 8 == 0), 5 K blocks > stage count wraps the ring
    (130, 100, 1024, 4, 32, "mlx"),      # scalar-store tail (n % 8 != 0), two groups per K block
    (96, 384, 512, 4, 128, "sym"),       # group spans two K blocks
    (64, 256, 256, 8, 64, "zp"),
    (300, 200, 512, 8, 32, "mlx"),
    (128, 128, 384, 8, 128, "sym"),
    (512, 3072, 1024, 4, 64, "zp"),      # Qwen3.5-0.8B qkv at the prefill=512 config
]


@pytest.mark.parametrize("m,n,k,bits,gs,method", CASES)
def test_prefill_gemm_f32_out(ctx, m, n, k, bits, gs, method):
    x, w, kw = _quant_case(300 + m + n + k, m, n, k, bits, gs, METHODS[method])
    ref = O.matmul(x, w, m=m, n=n, k=k, d_f32=True, **kw)
    got, launches = G.matmul(ctx, x, w, m=m, n=n, k=k, d_f32=True, return_launches=True, **kw)
    assert launches == 1, "prefill shapes must run as ONE tensor-core GEMM launch, not per-16-row GEMV passes"
    assert_f32_close(got, ref, rtol=1e-3, atol=1e-4, what=f"prefill gemm {m}x{n}x{k} int{bits} gs{gs} {method}")


@pytest.mark.parametrize("m,n,k,bits,gs,method", CASES[:6])
def test_prefill_gemm_bf16_out(ctx, m, n, k, bits, gs, method):
    x, w, kw = _quant_case(400 + m + n + k, m, n, k, bits, gs, METHODS[method])
    ref = O.matmul(x, w, m=m, n=n, k=k, **kw)
    got = G.matmul(ctx, x, w, m=m, n=n, k=k, **kw)
    assert_bf16_close(got, ref, max_ulp=1, min_exact=0.97, what="prefill gemm bf16")


@pytest.mark.parametrize("mt", [1, 2])
def test_prefill_gemm_token_tile_variants_agree(ctx, mt):
    """128- and 256-token CTAs (one / two accumulator tiles sharing the dequantised weight tile) give the same bits."""
    m, n, k = 384, 256, 512
    x, w, kw = _quant_case(11, m, n, k, 4, 64, O.QM_ZERO_POINT)
    base = G.matmul(ctx, x, w, m=m, n=n, k=k, d_f32=True, **kw)
    ctx.lib.uzu_debug_set_umma(-1, 0, 0, 0, 0, mt)
    try:
        got = G.matmul(ctx, x, w, m=m, n=n, k=k, d_f32=True, **kw)
    finally:
        ctx.lib.uzu_debug_set_umma(-1, 0, 0, 0, 0, 0)
    assert (got == base).all()


def test_prefill_gemm_epilogue_and_signed_codes(ctx):
    m, n, k = 160, 144, 256
    x, w, kw = _quant_case(12, m, n, k, 4, 64, O.QM_ZERO_POINT)
    rng = np.random.default_rng(13)
    bias = f32_to_bf16(rng.uniform(-1, 1, n).astype(np.float32))
    d0 = rng.uniform(-1, 1, (m, n)).astype(np.float32)
    ep = dict(ab_scale=1.7, accumulate=True, bias=bias, soft_cap=2.5)
    ref = O.matmul(x, w, m=m, n=n, k=k, d=d0.copy(), **kw, **ep)
    got = G.matmul(ctx, x, w, m=m, n=n, k=k, d=d0.copy(), **kw, **ep)
    assert_f32_close(got, ref, what="prefill epilogue")
    a = G.matmul(ctx, x, w, m=m, n=n, k=k, d_f32=True, **kw)
    b = G.matmul(ctx, x, w ^ np.uint8(0x88), m=m, n=n, k=k, d_f32=True, signed_codes=True, **kw)
    assert (a == b).all()


def test_prefill_gemm_rows_match_decode_gemv(ctx):
    """The same activation row through the m = 1 decode GEMV and through the 128-token tensor-core tile: both are f32-accumulated
    exact products, so they agree to f32 summation-order noise (prefill and decode of the same token see the same logits)."""
    m, n, k = 128, 512, 1024
    x, w, kw = _quant_case(14, m, n, k, 4, 64, O.QM_ZERO_POINT)
    full = G.matmul(ctx, x, w, m=m, n=n, k=k, d_f32=True, **kw)
    for r in (0, 77, 127):
        one = G.matmul(ctx, np.ascontiguousarray(x[r:r + 1]), w, m=1, n=n, k=k, d_f32=True, **kw)
        assert_f32_close(full[r:r + 1], one, rtol=1e-4, atol=1e-5, what=f"row {r}")


def test_prefill_gemm_linearity_full_size(ctx):
    """Size-independent properties at a BASELINE shape (Llama-3-8B qkv 6144 x 4096 int4, m = 512): zero rows map to exact zeros,
    A(x1 + x2) == A x1 + A x2, and a sample of outputs matches the oracle."""
    m, n, k = 512, 6144, 4096
    rng = np.random.default_rng(15)
    packed, scales, zp, _ = random_quant(rng, n, k, 4, 64, O.QM_ZERO_POINT)
    kw = dict(scales=scales, zero_points=zp, method=O.QM_ZERO_POINT, bits=4, group_size=64)
    x1 = f32_to_bf16((rng.integers(-8, 8, (m, k)) / 16).astype(np.float32))
    x2 = f32_to_bf16((rng.integers(-8, 8, (m, k)) / 16).astype(np.float32))
    x1[5] = 0
    x2[5] = 0
    x12 = f32_to_bf16(bf16_to_f32(x1) + bf16_to_f32(x2))
    y1 = G.matmul(ctx, x1, packed, m=m, n=n, k=k, d_f32=True, **kw)
    y2 = G.matmul(ctx, x2, packed, m=m, n=n, k=k, d_f32=True, **kw)
    y12 = G.matmul(ctx, x12, packed, m=m, n=n, k=k, d_f32=True, **kw)
    assert (y1[5] == 0).all()
    assert_f32_close(y12, y1 + y2, rtol=1e-4, atol=1e-3, what="linearity")
    rows = rng.choice(m, 4, replace=False)
    cols = rng.choice(n, 64, replace=False)
    ref = O.matmul(np.ascontiguousarray(x1[rows]), np.ascontiguousarray(packed[cols]), m=4, n=64, k=k, d_f32=True,
                   scales=np.ascontiguousarray(scales[cols]), zero_points=np.ascontiguousarray(zp[cols]), method=O.QM_ZERO_POINT,
                   bits=4, group_size=64)
    assert_f32_close(y1[np.ix_(rows, cols)], ref, what="sampled outputs")
