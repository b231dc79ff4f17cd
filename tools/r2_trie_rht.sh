#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 500 python -m pytest tests/test_trie_gpu.py tests/test_rht_gpu.py -q --timeout 150 --timeout-method thread > gpurun_out/r2s_trie_rht_tests.log 2>&1; echo "trie+rht tests rc=$?"
tail -n 40 gpurun_out/r2s_trie_rht_tests.log
timeout -s KILL 400 python -u tools/trie_probe.py llama3-8b-int4 2048 > gpurun_out/r2s_trie_probe.json 2> gpurun_out/r2s_trie_probe.err; echo "probe rc=$?"
tail -n 3 gpurun_out/r2s_trie_probe.err; cat gpurun_out/r2s_trie_probe.json
