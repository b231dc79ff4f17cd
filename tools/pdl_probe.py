"""Per-kernel-family timing inside the engine process (CUDA events around back-to-back launches over all layers, PDL on / off).
usage: python tools/pdl_probe.py <workload> [prefill_tokens]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from uzu_b200 import binding as B
wl = sys.argv[1] if len(sys.argv) > 1 else "llama3-8b-int4"
prefill = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
ctx = B.Context(); eng = B.Engine(ctx, bench.model_dir_for(wl), max_context_length=prefill + 64)
info = eng.info
eng.prefill(np.random.default_rng(0).integers(0, info.vocab_size, prefill, dtype=np.uint32))
sels = {1: "mixer_in", 2: "mixer_out", 4: "up", 32: "up+gated", 8: "down", 16: "readout", 31: "all linears", 64: f"attention mix @ctx {prefill}"}
if info.num_delta_net_layers:
    sels[128] = "deltanet conv+update"
for sel, name in sels.items():
    r = []
    for nopdl in (0, 1):
        t, n = eng.time_linears(10, sel | (0x80000000 if nopdl else 0))
        r.append((t / max(n, 1) * 1e6, n // 10))
    print(f"{wl} {name}: pdl {r[0][0]:.2f} us/launch, no-pdl {r[1][0]:.2f} us/launch ({r[0][1]} launches/pass, {r[0][0] * r[0][1]:.0f} us/pass)", flush=True)
