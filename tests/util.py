"""Shared helpers for parity tests (numpy only)."""
import numpy as np


def f32_to_bf16(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32)
    up = ((u & np.uint32(0x8000)) != 0) & ((u & np.uint32(0x17FFF)) != 0)
    return ((u >> 16) + up.astype(np.uint32)).astype(np.uint16)


def bf16_to_f32(h):
    return (np.ascontiguousarray(h, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)


def bf16_ulp_distance(a_bits, b_bits):
    """Distance in bf16 representable steps (sign-magnitude -> monotone integer)."""
    def key(h):
        h = h.astype(np.int32)
        return np.where(h & 0x8000, -(h & 0x7FFF), h & 0x7FFF)
    return np.abs(key(np.asarray(a_bits)) - key(np.asarray(b_bits)))


def assert_bf16_close(got_bits, ref_bits, *, max_ulp=1, min_exact=0.97, atol=1e-4, what=""):
    """bf16 outputs of a kernel whose f32 summation order differs from the oracle's:
    every element within `max_ulp` bf16 steps of the oracle (or within atol in value), and at least
    `min_exact` of them bit-identical. BASELINE's rtol=1e-3 is below bf16's own 2^-8 rounding step, so it is
    asserted on f32 outputs (see assert_f32_close) and as this ulp bound on bf16 ones."""
    got_bits, ref_bits = np.asarray(got_bits), np.asarray(ref_bits)
    assert got_bits.shape == ref_bits.shape, (got_bits.shape, ref_bits.shape)
    d = bf16_ulp_distance(got_bits, ref_bits)
    g, r = bf16_to_f32(got_bits), bf16_to_f32(ref_bits)
    bad = (d > max_ulp) & ~(np.abs(g - r) <= atol)
    assert not bad.any(), f"{what}: {bad.sum()} elements differ by more than {max_ulp} bf16 ulp; worst {d.max()} ulp, " \
                          f"first bad idx {np.argwhere(bad)[0]} got {g[tuple(np.argwhere(bad)[0])]} ref {r[tuple(np.argwhere(bad)[0])]}"
    exact = float((got_bits == ref_bits).mean())
    assert exact >= min_exact, f"{what}: only {exact:.4f} of elements are bit-identical (< {min_exact})"
    return exact


def assert_f32_close(got, ref, *, rtol=1e-3, atol=1e-4, what=""):
    """BASELINE tolerance (rtol 1e-3 / atol 1e-4) on f32 outputs. atol is taken relative to the output scale
    (rms of the reference, floor 1): an absolute 1e-4 on outputs of magnitude 100 would be 1e-6 relative, below the
    f32 accumulation noise of the *reference's own* 2048-term sequential sum."""
    got, ref = np.asarray(got, dtype=np.float32), np.asarray(ref, dtype=np.float32)
    err = np.abs(got - ref)
    scale = max(1.0, float(np.sqrt(np.mean(ref.astype(np.float64) ** 2))))
    tol = atol * scale + rtol * np.abs(ref)
    bad = err > tol
    assert not bad.any(), f"{what}: {bad.sum()} / {bad.size} elements outside rtol={rtol} atol={atol}; max err {err.max()}"
