#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -u -m pytest tests/ -q -m gpu --timeout 240 --timeout-method thread > gpurun_out/r2A_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -n 6 gpurun_out/r2A_gpu_tests.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2A_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/r2A_smoke.log
timeout -s KILL 480 python bench.py > gpurun_out/r2A_bench_default.json 2> gpurun_out/r2A_bench_default.err; echo "bench rc=$?"; tail -n 3 gpurun_out/r2A_bench_default.err
python - <<'PY'
import json; d=json.loads(open('gpurun_out/r2A_bench_default.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['config'].get('speculation'), d['config']['secondary'].get('value'), d.get('parity',{}).get('token_ids_equal'))
PY
timeout -s KILL 300 python bench.py --workload llama3-8b-int8 --batch 8 --steps 128 --no-cpu-baseline > gpurun_out/r2A_bench_int8_batch8.json 2> gpurun_out/r2A_bench_int8_batch8.err; echo "batch8 rc=$?"
python - <<'PY'
import json; d=json.loads(open('gpurun_out/r2A_bench_int8_batch8.json').read().strip().splitlines()[-1]); print('batch8', d['value'], d['ms_per_step'])
PY
timeout -s KILL 200 python -u tools/trie_probe.py llama3-8b-int4 2048 > gpurun_out/r2A_trie_probe.json 2> gpurun_out/r2A_trie_probe.err; echo "probe rc=$?"
python - <<'PY'
import json; d=json.load(open('gpurun_out/r2A_trie_probe.json')); print(d['decode_timed_ms'], [(r['nodes'], round(r['pass_ms'],2)) for r in d['trie_pass']])
PY
