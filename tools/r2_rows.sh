#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 200 --timeout-method thread -k "rows_kernel or quantized_matmul" > gpurun_out/r2u_rows_tests.log 2>&1; echo "rows kernel tests rc=$?"
tail -n 25 gpurun_out/r2u_rows_tests.log
timeout -s KILL 500 python -m pytest tests/test_trie_gpu.py tests/test_batch_decode_gpu.py tests/test_rht_gpu.py -q --timeout 200 --timeout-method thread > gpurun_out/r2u_engine_tests.log 2>&1; echo "trie/batch/rht tests rc=$?"
tail -n 8 gpurun_out/r2u_engine_tests.log
timeout -s KILL 300 python -u tools/trie_probe.py llama3-8b-int4 2048 > gpurun_out/r2u_trie_probe.json 2> gpurun_out/r2u_trie_probe.err; echo "probe rc=$?"
python - <<'PY'
import json; d=json.load(open('gpurun_out/r2u_trie_probe.json')); print(d['decode_timed_ms'], [(r['nodes'], round(r['pass_ms'],2)) for r in d['trie_pass']])
PY
timeout -s KILL 500 python bench.py --workload llama3-8b-int8 --batch 8 --steps 128 --no-cpu-baseline > gpurun_out/r2u_bench_int8_batch8.json 2> gpurun_out/r2u_bench_int8_batch8.err; echo "batch8 rc=$?"
python - <<'PY'
import json; d=json.loads(open('gpurun_out/r2u_bench_int8_batch8.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['config'].get('workload'))
PY
