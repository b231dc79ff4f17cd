#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 python -u -m pytest tests/test_engine_gpu.py tests/test_mega_gpu.py tests/test_batch_decode_gpu.py -q -x --timeout 200 --timeout-method thread > gpurun_out/r2n_tests.log 2>&1; echo "tests rc=$?"; tail -n 6 gpurun_out/r2n_tests.log
timeout -s KILL 120 python -u tools/mega_probe.py llama3-8b-int4 > gpurun_out/r2n_probe_llama.log 2>&1; echo "llama rc=$?"; grep -E "64 steps" gpurun_out/r2n_probe_llama.log
UZU_QMV_TMA=0 timeout -s KILL 120 python -u tools/mega_probe.py llama3-8b-int4 > gpurun_out/r2n_probe_llama_notma.log 2>&1; echo "llama no-tma rc=$?"; grep -E "64 steps" gpurun_out/r2n_probe_llama_notma.log
timeout -s KILL 100 python -u tools/mega_probe.py qwen3.5-0.8b-int4 > gpurun_out/r2n_probe_qwen.log 2>&1; echo "qwen rc=$?"; grep -E "64 steps" gpurun_out/r2n_probe_qwen.log
