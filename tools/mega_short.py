"""A few persistent decode steps of a BASELINE workload (the target of ncu captures): python tools/mega_short.py <workload> [prefill] [steps]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import bench
from uzu_b200 import binding as B
workload = sys.argv[1]
prefill = int(sys.argv[2]) if len(sys.argv) > 2 else bench.WORKLOADS[workload][2]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
ctx = B.Context(0)
eng = B.Engine(ctx, bench.model_dir_for(workload), max_context_length=max(1024, prefill + steps + 64))
eng.set_persistent_decode(True)
assert eng.persistent_decode, eng.persistent_decode_reason
rng = np.random.default_rng(0)
tok = eng.prefill(rng.integers(0, eng.info.vocab_size, prefill).astype(np.uint32))
for _ in range(steps):
    tok = eng.step_host(tok)
print("ok", tok)
eng.close(); ctx.close()
