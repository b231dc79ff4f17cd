"""The arithmetic claims behind the tensor-core prefill GEMM's dequant stage (uzu_b200/csrc/prefill_gemm.cu), checked
exhaustively in numpy (float64 holds every intermediate exactly, so this is a proof over the enumerated domain, not a sample):

  1. ZeroPoint / Symmetric weights: w = scale*(code - zp) (what the reference's `scale*code + corr`, corr = -scale*zp, evaluates to in f32,
     cpu/kernel/matmul/kernel.rs:254-276) splits EXACTLY into two bf16 planes hi = RNE_bf16(w), lo = w - hi: lo is representable, hi + lo == w.
  2. the kernel's evaluation order gives the same value: fmaf(scale, BASE + code, -scale*(zp + BASE)) with BASE = 128 (4-bit) / 256 (8-bit),
     both products exact in f32.
  3. the packed 4-bit path (d = (128+q) - (128+zp) in bf16, hi = RNE(scale*d), -lo = fma(-scale, d, hi)) produces the same two planes.
"""
import numpy as np

from tests.util import bf16_to_f32, f32_to_bf16


def _normal_bf16_scales():
    bits = np.arange(0x0080, 0x7F80, dtype=np.uint32).astype(np.uint16)          # every positive normal bf16
    s = bf16_to_f32(bits)
    return s[(s > 2.0 ** -100) & (s < 2.0 ** 100)]                                # scales whose products stay far from f32 under/overflow


def _rne_bf16(x64):
    """float64 (exactly representable in f32) -> nearest bf16 as float64"""
    return bf16_to_f32(f32_to_bf16(x64.astype(np.float32))).astype(np.float64)


def _check(bits):
    s = _normal_bf16_scales().astype(np.float64)[:, None]
    base = 128.0 if bits == 4 else 256.0
    d = np.arange(-(2 ** bits - 1), 2 ** bits, dtype=np.float64)[None, :]         # code - zp over all (code, zp) pairs
    w = s * d                                                                     # exact: <= 8 + 9 significant bits
    assert (w.astype(np.float32).astype(np.float64) == w).all(), "w is an f32 value (the reference's f32 arithmetic is exact here)"
    hi = _rne_bf16(w)
    lo = w - hi
    assert (_rne_bf16(lo) == lo).all(), "the remainder is a bf16 value"
    assert (hi + lo == w).all()
    return s, w, hi, lo, base


def test_int4_planes_are_exact_and_match_the_kernel_order():
    s, w, hi, lo, base = _check(4)
    # (2) fmaf(s, BASE + q, -s*(z + BASE)) for every (q, z): both products are exact f32 values, the fused sum is exact
    q = np.arange(16, dtype=np.float64)
    for z in range(16):
        c = -(s * (z + base))
        assert (c.astype(np.float32).astype(np.float64) == c).all()
        fused = s * (base + q)[None, :] + c                                        # float64: exact
        assert (fused == s * (q - z)[None, :]).all()
    # (3) packed path: d exact in bf16, hi = RNE(s*d), -lo = fma(-s, d, hi) exact
    dq = np.arange(-15, 16, dtype=np.float64)[None, :]
    assert (_rne_bf16(dq) == dq).all() and (_rne_bf16(np.float64(128.0) + np.arange(16)) == 128.0 + np.arange(16)).all()
    h2 = _rne_bf16(s * dq)
    nl = -(s * dq) + h2                                                            # fma(-s, d, hi): single rounding of this exact value
    assert (_rne_bf16(nl) == nl).all() and (h2 - nl == s * dq).all()


def test_int8_planes_are_exact_and_match_the_kernel_order():
    s, w, hi, lo, base = _check(8)
    rng = np.random.default_rng(0)
    q = np.arange(256, dtype=np.float64)
    for z in rng.integers(0, 256, 24):
        c = -(s * (float(z) + base))
        assert (c.astype(np.float32).astype(np.float64) == c).all()                # 8 x 10 significant bits
        p = s * (base + q)[None, :]
        assert (p.astype(np.float32).astype(np.float64) == p).all()
        assert (p + c == s * (q - float(z))[None, :]).all()


def test_single_plane_would_miss_the_parity_budget():
    """Why two planes: one rounded bf16 plane carries up to 2^-9 relative error per weight, i.e. ~1e-3 of a dot product's rms -- the
    whole rtol the north_star allows."""
    s, w, hi, lo, _ = _check(4)
    rel = np.abs(lo[w != 0] / w[w != 0])
    assert rel.max() > 1.5e-3 and rel.max() <= 2.0 ** -8
