#!/bin/bash
mkdir -p gpurun_out
for W in 8 16; do
export UZU_QMV_ROWS_WARPS=$W
timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 200 --timeout-method thread -k "rows_kernel" > gpurun_out/r2z_rows_tests_w$W.log 2>&1; echo "W=$W rows kernel tests rc=$?"
tail -n 3 gpurun_out/r2z_rows_tests_w$W.log
timeout -s KILL 300 python -u tools/trie_probe.py llama3-8b-int4 2048 > gpurun_out/r2z_trie_probe_w$W.json 2> gpurun_out/r2z_trie_probe_w$W.err; echo "probe rc=$?"
python - <<PY
import json; d=json.load(open('gpurun_out/r2z_trie_probe_w$W.json')); print('W=$W', d['decode_timed_ms'], [(r['nodes'], round(r['pass_ms'],2)) for r in d['trie_pass']])
PY
timeout -s KILL 500 python bench.py --workload llama3-8b-int8 --batch 8 --steps 128 --no-cpu-baseline > gpurun_out/r2z_bench_int8_batch8_w$W.json 2> gpurun_out/r2z_bench_int8_batch8_w$W.err; echo "batch8 rc=$?"
python - <<PY
import json; d=json.loads(open('gpurun_out/r2z_bench_int8_batch8_w$W.json').read().strip().splitlines()[-1]); print('W=$W batch8', d['value'], d['ms_per_step'])
PY
done
