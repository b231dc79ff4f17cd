"""Micro-benchmark of the fused dequant+GEMV through the C ABI: GB/s per shape, cold-cache (rotates through
enough weight copies to exceed the 126 MB L2, like the reference's tests/cold_pool.rs), device vs managed memory.
usage: python tools/gemv_bench.py [--kind device|managed] [--shapes n,k;n,k] [--m 1] [--iters 20]"""
import argparse
import ctypes as C
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from uzu_b200 import binding as B  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--kind", default="device")
ap.add_argument("--shapes", default="6144,4096;4096,4096;28672,4096;4096,14336;128256,4096;3072,1024;1024,2048;7168,1024;1024,3584;8224,1024;248320,1024")
ap.add_argument("--m", type=int, default=1)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--bits", type=int, default=4)
args = ap.parse_args()

ctx = B.Context(0)
rng = np.random.default_rng(0)
kind = B.BUFFER_DEVICE if args.kind == "device" else B.BUFFER_MANAGED


def put(arr):
    arr = np.ascontiguousarray(arr)
    if kind == B.BUFFER_MANAGED:
        return ctx.upload(arr)
    staging = ctx.upload(arr, B.BUFFER_PINNED_HOST)
    dev = ctx.buffer(max(arr.nbytes, 16), B.BUFFER_DEVICE)
    with ctx.command_buffer("up") as cmd:
        cmd.encode("uzu_command_buffer_encode_copy", staging.ptr, dev.ptr, arr.nbytes)
    staging.close()
    return dev


for shape in args.shapes.split(";"):
    n, k = map(int, shape.split(","))
    gs = 64
    groups = k // gs
    wbytes = n * k * args.bits // 8 + n * groups * 2 + n * (groups // 2 if args.bits == 4 else groups)
    copies = max(2, int(300e6 // wbytes) + 1)
    copies = min(copies, 64)
    ws = [(put(rng.integers(0, 256, (n, k * args.bits // 8), dtype=np.uint8)),
           put((rng.integers(0x3C00, 0x3E00, (n, groups))).astype(np.uint16)),
           put(rng.integers(0, 256, (n, groups // 2 if args.bits == 4 else groups), dtype=np.uint8))) for _ in range(copies)]
    x = put((rng.integers(0x3C00, 0x3F00, (args.m, k))).astype(np.uint16))
    d = put(np.zeros((args.m, n), np.uint16))

    def enc(cmd, i):
        w, s, z = ws[i % copies]
        a = B.MatmulArgs(a=x.ptr, b=w.ptr, b_scales=s.ptr, b_zero_points=z.ptr, d=d.ptr, b_prologue=B.B_SCALE_ZERO_POINT,
                         b_mode=B.QMODE_U4 if args.bits == 4 else B.QMODE_U8, b_group_size=gs, b_transpose=1, ab_scale=1.0, m=args.m, n=n, k=k,
                         weights_dt=B.DT_BF16, input_dt=B.DT_BF16, output_dt=B.DT_BF16)
        cmd.encode("uzu_matmul_encode", C.byref(a))

    with ctx.command_buffer("warm") as cmd:
        for i in range(3):
            enc(cmd, i)
    with ctx.command_buffer("timed") as cmd:
        for i in range(args.iters):
            enc(cmd, i)
    us = cmd.gpu_seconds / args.iters * 1e6
    print(f"{args.kind:8s} m={args.m} n={n:6d} k={k:5d} bytes={wbytes/1e6:8.2f} MB  {us:8.2f} us/launch  {wbytes/us/1e3:8.1f} GB/s  ({copies} copies)", flush=True)
    for w, s, z in ws:
        w.close(); s.close(); z.close()
ctx.close()
