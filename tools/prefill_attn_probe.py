"""Quick hardware check of the opt-in tensor-core prefill attention against a float64 numpy softmax (no oracle, a few seconds):
python tools/prefill_attn_probe.py"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from uzu_b200 import binding as B  # noqa: E402


def bf16(x):
    u = np.ascontiguousarray(x, np.float32).view(np.uint32)
    up = ((u & 0x8000) != 0) & ((u & 0x17FFF) != 0)
    return ((u >> 16) + up).astype(np.uint16)


def f32(h):
    return (h.astype(np.uint32) << 16).view(np.float32)


def run(ctx, H, Hkv, prefix, suffix, D, causal, mode):
    rng = np.random.default_rng(1)
    seq = prefix + suffix
    q = bf16(rng.standard_normal((H, suffix, D)))
    k = bf16(rng.standard_normal((seq, Hkv * D)))
    v = bf16(rng.standard_normal((seq, Hkv * D)))
    ctx.lib.uzu_debug_set_prefill_attention(mode)
    bq, bk, bv, bo = ctx.upload(q), ctx.upload(k), ctx.upload(v), ctx.upload(np.zeros((suffix, H, D), np.uint16))
    a = B.AttentionArgs(queries=bq.ptr, keys=bk.ptr, values=bv.ptr, out=bo.ptr, gqa_factor=H // Hkv, sequence_length=seq, k_head_stride=D,
                        k_seq_stride=Hkv * D, v_head_stride=D, v_seq_stride=Hkv * D, scale=float(1 / np.sqrt(D)), num_heads=H,
                        suffix_length=suffix, head_dim=D, is_causal=int(causal))
    with ctx.command_buffer("probe") as cmd:
        cmd.encode("uzu_attention_single_pass_encode", C.byref(a))
    got = f32(bo.numpy(np.uint16, (suffix, H, D))).astype(np.float64)
    qf, kf, vf = f32(q).astype(np.float64), f32(k).astype(np.float64).reshape(seq, Hkv, D), f32(v).astype(np.float64).reshape(seq, Hkv, D)
    ref = np.zeros_like(got)
    g = H // Hkv
    for h in range(H):
        s = qf[h] @ kf[:, h // g].T / np.sqrt(D)
        if causal:
            for t in range(suffix):
                s[t, prefix + t + 1:] = -np.inf
        p = np.exp(s - s.max(axis=1, keepdims=True))
        ref[:, h] = (p / p.sum(axis=1, keepdims=True)) @ vf[:, h // g]
    err = np.abs(got - ref)
    print(f"mode {mode} H{H} Hkv{Hkv} prefix{prefix} suffix{suffix} D{D} causal{int(causal)}: max err {err.max():.3e} (bf16 step ~4e-3) "
          f"bad rows {int((err.max(axis=(1, 2)) > 2e-2).sum())}/{suffix} bad heads {int((err.max(axis=(0, 2)) > 2e-2).sum())}/{H} nan {int(np.isnan(got).sum())}", flush=True)
    if err.max() > 2e-2:
        print("  per-row max err (first 20):", np.round(err.max(axis=(1, 2))[:20], 3))
        print("  got[0,0,:6]", got[0, 0, :6], "ref", ref[0, 0, :6])


if __name__ == "__main__":
    with B.Context(0) as ctx:
        run(ctx, 8, 2, 0, 64, 128, True, 1)
        run(ctx, 8, 2, 77, 45, 128, True, 1)
        run(ctx, 4, 4, 10, 33, 64, False, 1)
        run(ctx, 8, 2, 77, 45, 128, True, 0)      # the split-KV kernel on the same inputs, for reference
