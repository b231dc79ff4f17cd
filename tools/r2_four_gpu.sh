#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 4 > gpurun_out/r2q_bench_n4.json 2> gpurun_out/r2q_bench_n4.err; echo "bench n4 rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2q_bench_n4.json').read().strip().splitlines()[-1])
print(d["value"], d["config"]["per_gpu_value"], d["ms_per_step"], d["e2e"]["value"]); print(json.dumps(d["config"].get("tp"), indent=1)); print(d["roofline"]["frac"], d["roofline"]["achieved"])
PY
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r2q_bench_n4.err | tail -n 6
