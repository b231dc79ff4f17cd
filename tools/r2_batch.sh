#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -u -m pytest tests/test_batch_decode_gpu.py -q --timeout 120 --timeout-method thread > gpurun_out/r2l_batch_tests.log 2>&1; echo "batch tests rc=$?"; tail -n 6 gpurun_out/r2l_batch_tests.log
timeout -s KILL 400 python bench.py --workload llama3-8b-int8 --batch 8 --steps 128 > gpurun_out/r2l_bench_llama_int8_batch8.json 2> gpurun_out/r2l_bench_llama_int8_batch8.err; echo "batch8 rc=$?"; cat gpurun_out/r2l_bench_llama_int8_batch8.json; tail -n 3 gpurun_out/r2l_bench_llama_int8_batch8.err
