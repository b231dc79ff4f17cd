#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 200 --timeout-method thread -k "rows_kernel" > gpurun_out/r2w_rows_tests.log 2>&1; echo "rows kernel tests rc=$?"
tail -n 5 gpurun_out/r2w_rows_tests.log
timeout -s KILL 300 python -u tools/trie_probe.py llama3-8b-int4 2048 > gpurun_out/r2w_trie_probe.json 2> gpurun_out/r2w_trie_probe.err; echo "probe rc=$?"
python - <<'PY'
import json; d=json.load(open('gpurun_out/r2w_trie_probe.json')); print(d['decode_timed_ms'], [(r['nodes'], round(r['pass_ms'],2)) for r in d['trie_pass']])
PY
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:qmv_rows_kernel -c 4 -o gpurun_out/r2w_rows_full python tools/trie_short.py llama3-8b-int4 64 16 > gpurun_out/r2w_rows_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/r2w_rows_full.ncu-rep
