"""Sweeps the UMMA shared-memory descriptor / layout encodings of the tcgen05 prefill GEMM on a real B200 (uzu_debug_set_umma) and
reports, per variant, the error against a float64 numpy contraction of the dequantised weights. Every variant runs in its own
process under a timeout: a wrong descriptor can trap the kernel (the mbarrier watchdog) and poison the CUDA context.

    python tools/umma_probe.py            # all variants
    python tools/umma_probe.py one <layout> <desc_hi> <lbo> <k_step> <idesc> <mt> <bits>   # one variant, in-process
"""
import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

VARIANTS = [
    # name, layout, desc_hi, lbo, k_step, idesc, mt, bits
    ("default (sw128, sbo=1024, lbo=1, v1)", -1, 0, 0, 0, 0, 1, 4),
    ("default mt=2", -1, 0, 0, 0, 0, 2, 4),
    ("default int8", -1, 0, 0, 0, 0, 1, 8),
    ("sw128 lbo=0", 0, 0x40004040, 0, 2, 0, 1, 4),
    ("sw128 version=0", 0, 0x40000040, 1, 2, 0, 1, 4),
    # (round-1 history: the no-swizzle core-matrix layout -- 8 rows x 16 B, LBO = 128, SBO = 1024, k step 256 B -- also matched the
    #  reference and LBO/SBO swapped did not; the kernel now keeps the SWIZZLE_128B writer only, so those variants are gone)
    ("sw128 wrong sbo=512 (must FAIL: proves the probe can tell)", 0, 0x40004020, 1, 2, 0, 1, 4),
]


def bf16(x):
    u = np.ascontiguousarray(x, np.float32).view(np.uint32)
    up = ((u & 0x8000) != 0) & ((u & 0x17FFF) != 0)
    return ((u >> 16) + up).astype(np.uint16)


def f32(h):
    return (h.astype(np.uint32) << 16).view(np.float32)


def one(layout, desc_hi, lbo, k_step, idesc, mt, bits, m=256, n=256, k=256):
    from uzu_b200 import binding as B
    rng = np.random.default_rng(0)
    gs = 64
    codes = rng.integers(0, 1 << bits, (n, k), dtype=np.uint8)
    packed = (codes[:, 0::2] | (codes[:, 1::2] << 4)).astype(np.uint8) if bits == 4 else codes
    scales = bf16(rng.uniform(0.01, 0.3, (n, k // gs)))
    zpv = rng.integers(0, 1 << bits, (n, k // gs), dtype=np.uint8)
    zp = (zpv[:, 0::2] | (zpv[:, 1::2] << 4)).astype(np.uint8) if bits == 4 else zpv
    x = bf16(rng.uniform(-0.3, 0.3, (m, k)))
    w = np.repeat(f32(scales), gs, axis=1).astype(np.float64) * (codes.astype(np.float64) - np.repeat(zpv, gs, axis=1))
    ref = f32(x).astype(np.float64) @ w.T
    with B.Context(0) as ctx:
        ctx.lib.uzu_debug_set_umma(layout, desc_hi, lbo, k_step, idesc, mt)
        bufs = [ctx.upload(a) for a in (x, packed, scales, zp, np.zeros((m, n), np.float32))]
        args = B.MatmulArgs(a=bufs[0].ptr, b=bufs[1].ptr, b_scales=bufs[2].ptr, b_zero_points=bufs[3].ptr, d=bufs[4].ptr,
                            b_prologue=B.B_SCALE_ZERO_POINT, b_mode=B.QMODE_U4 if bits == 4 else B.QMODE_U8, b_group_size=gs,
                            b_transpose=1, ab_scale=1.0, m=m, n=n, k=k, weights_dt=B.DT_BF16, input_dt=B.DT_BF16, output_dt=B.DT_F32)
        with ctx.command_buffer("probe") as cmd:
            cmd.encode("uzu_matmul_encode", C.byref(args))
        got = bufs[4].numpy(np.float32, (m, n)).astype(np.float64)
    err = np.abs(got - ref)
    rms = float(np.sqrt((ref ** 2).mean()))
    ok_rows = int((err.max(axis=1) < 1e-3 * rms).sum())
    ok_cols = int((err.max(axis=0) < 1e-3 * rms).sum())
    print(f"max_err {err.max():.3e} rms_ref {rms:.3e} rel {err.max() / rms:.3e} ok_rows {ok_rows}/{m} ok_cols {ok_cols}/{n} "
          f"nonzero {int((got != 0).sum())}/{got.size} launches {cmd.launches}")
    if err.max() > 1e-3 * rms:
        # which 8x8 (row block, col block) pattern is right helps telling a layout error from a descriptor error
        blk = (err.reshape(m // 32, 32, n // 32, 32).max(axis=(1, 3)) < 1e-3 * rms).astype(int)
        print("ok map of 32x32 blocks:\n" + "\n".join("".join(map(str, r)) for r in blk))
        print("got[0,:6]", got[0, :6], "ref[0,:6]", ref[0, :6])
    return err.max() <= 1e-3 * rms


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        a = sys.argv[2:]
        ok = one(int(a[0]), int(a[1], 0), int(a[2], 0), int(a[3], 0), int(a[4], 0), int(a[5]), int(a[6]))
        sys.exit(0 if ok else 1)
    for name, layout, hi, lbo, ks, idesc, mt, bits in VARIANTS:
        cmd = [sys.executable, __file__, "one", str(layout), hex(hi), str(lbo), str(ks), hex(idesc), str(mt), str(bits)]
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=90)
            tail = (r.stdout + r.stderr).strip().splitlines()[-14:]
            print(f"== {name}: exit {r.returncode}\n   " + "\n   ".join(tail), flush=True)
        except subprocess.TimeoutExpired:
            print(f"== {name}: TIMEOUT", flush=True)
