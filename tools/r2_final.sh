#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -u -m pytest tests/ -q -m gpu --timeout 240 --timeout-method thread > gpurun_out/r2o_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -n 6 gpurun_out/r2o_gpu_tests.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2o_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/r2o_smoke.log
timeout -s KILL 480 python bench.py > gpurun_out/r2o_bench_default.json 2> gpurun_out/r2o_bench_default.err; echo "bench rc=$?"; cut -c1-700 gpurun_out/r2o_bench_default.json; tail -n 4 gpurun_out/r2o_bench_default.err
