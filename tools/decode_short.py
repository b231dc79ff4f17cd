"""A few decode steps of a BASELINE workload on a chosen decode path (ncu target): python tools/decode_short.py <workload> <prefill> <steps> kernels|persistent"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import bench
from uzu_b200 import binding as B
workload, prefill, steps, mode = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
ctx = B.Context(0)
eng = B.Engine(ctx, bench.model_dir_for(workload), max_context_length=max(1024, prefill + steps + 64))
eng.set_persistent_decode(mode == "persistent")
rng = np.random.default_rng(0)
tok = eng.prefill(rng.integers(0, eng.info.vocab_size, prefill).astype(np.uint32))
for _ in range(steps):
    tok = eng.step_host(tok)
print("ok", tok, "persistent" if eng.persistent_decode else "kernels")
eng.close(); ctx.close()
