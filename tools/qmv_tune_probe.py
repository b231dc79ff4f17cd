"""A/B of the decode GEMV's work-split overrides (uzu_debug_set_qmv_tuning) on the per-kernel decode path of a BASELINE workload:
    python tools/qmv_tune_probe.py <workload> [steps]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import bench
from uzu_b200 import binding as B
workload = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 64
prefill = bench.WORKLOADS[workload][2]
ctx = B.Context(0)
lib = ctx.lib
rng = np.random.default_rng(0)
for (wpt, dks, per_sm, stages) in [(0, 0, 0, 0), (0, 0, 0, 2), (0, 0, 0, 3), (0, 0, 0, 4), (0, 0, 2, 4), (0, 0, 3, 3), (0, 0, 2, 3)]:
    lib.uzu_debug_set_qmv_tuning(wpt, dks, per_sm, stages)
    eng = B.Engine(ctx, bench.model_dir_for(workload), max_context_length=max(1024, prefill + 4 * steps + 64))
    eng.set_persistent_decode(False)
    eng.prefill(rng.integers(0, eng.info.vocab_size, prefill).astype(np.uint32))
    eng.snapshot()
    eng.decode_timed(4)
    eng.restore()
    s = eng.decode_timed(steps)
    print(f"wpt={wpt} dks={dks} per_sm={per_sm} stages={stages}: {1e3 * s / steps:.4f} ms/step = {steps / s:.1f} tok/s", flush=True)
    eng.close()
lib.uzu_debug_set_qmv_tuning(0, 0, 0, 0)
ctx.close()
