#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -u -m pytest tests/ -q -m gpu --timeout 240 --timeout-method thread -x > gpurun_out/r2i_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -n 15 gpurun_out/r2i_gpu_tests.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2i_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/r2i_smoke.log
