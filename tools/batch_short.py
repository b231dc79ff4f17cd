"""A few batched decode steps (ncu target): python tools/batch_short.py <workload> <prefill> <steps> <batch>"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import bench
from uzu_b200 import binding as B
workload, prefill, steps, nseq = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
ctx = B.Context(0)
eng = B.Engine(ctx, bench.model_dir_for(workload), max_context_length=max(1024, prefill + steps + 64), use_cuda_graph=False)
rng = np.random.default_rng(0)
eng.batch_begin(nseq)
toks = [eng.batch_prefill(b, rng.integers(0, eng.info.vocab_size, prefill).astype(np.uint32)) for b in range(nseq)]
for _ in range(steps):
    toks = eng.batch_step(toks)
print("ok", toks)
eng.close(); ctx.close()
