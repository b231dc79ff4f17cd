"""Run one kernel family of the decode step back to back (for ncu): python tools/one_family.py <workload> <select> [iters] [prefill]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from uzu_b200 import binding as B
wl, sel = sys.argv[1], int(sys.argv[2])
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
prefill = int(sys.argv[4]) if len(sys.argv) > 4 else 0
ctx = B.Context(); eng = B.Engine(ctx, bench.model_dir_for(wl), max_context_length=max(prefill, 64) + 64)
if prefill:
    eng.prefill(np.random.default_rng(0).integers(0, eng.info.vocab_size, prefill, dtype=np.uint32))
t, n = eng.time_linears(iters, sel)
print(f"{wl} select {sel}: {t / n * 1e6:.2f} us/launch over {n} launches")
