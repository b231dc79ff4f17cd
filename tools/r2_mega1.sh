#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -u -m pytest tests/test_mega_gpu.py -x -v --timeout 90 --timeout-method thread > gpurun_out/r2b_mega_test.log 2>&1; echo "mega test rc=$?"
tail -n 60 gpurun_out/r2b_mega_test.log
