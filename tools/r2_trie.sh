#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_trie_gpu.py -x -q --timeout 120 --timeout-method thread > gpurun_out/r2r_trie_tests.log 2>&1; echo "trie tests rc=$?"
tail -n 15 gpurun_out/r2r_trie_tests.log
timeout -s KILL 400 python -u tools/trie_probe.py llama3-8b-int4 2048 > gpurun_out/r2r_trie_probe.json 2> gpurun_out/r2r_trie_probe.err; echo "probe rc=$?"
tail -n 3 gpurun_out/r2r_trie_probe.err; cat gpurun_out/r2r_trie_probe.json
