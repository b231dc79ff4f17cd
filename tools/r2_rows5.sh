#!/bin/bash
mkdir -p gpurun_out
for F in 1 0; do
export UZU_QMV_ROWS_FEED=$F
timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 200 --timeout-method thread -k "rows_kernel" > gpurun_out/r2y_rows_tests_f$F.log 2>&1; echo "FEED=$F rows kernel tests rc=$?"
tail -n 3 gpurun_out/r2y_rows_tests_f$F.log
timeout -s KILL 300 python -u tools/trie_probe.py llama3-8b-int4 2048 > gpurun_out/r2y_trie_probe_f$F.json 2> gpurun_out/r2y_trie_probe_f$F.err; echo "probe rc=$?"
python - <<PY
import json; d=json.load(open('gpurun_out/r2y_trie_probe_f$F.json')); print('FEED=$F', d['decode_timed_ms'], [(r['nodes'], round(r['pass_ms'],2)) for r in d['trie_pass']])
PY
done
export UZU_QMV_ROWS_FEED=1
timeout -s KILL 500 python bench.py --workload llama3-8b-int8 --batch 8 --steps 128 --no-cpu-baseline > gpurun_out/r2y_bench_int8_batch8_f1.json 2> gpurun_out/r2y_bench_int8_batch8_f1.err; echo "batch8 rc=$?"
python - <<PY
import json; d=json.loads(open('gpurun_out/r2y_bench_int8_batch8_f1.json').read().strip().splitlines()[-1]); print('FEED=1 batch8', d['value'], d['ms_per_step'])
PY
