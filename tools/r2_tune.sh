#!/bin/bash
mkdir -p gpurun_out
UZU_DECODE_PATH=kernels timeout -s KILL 300 python -u tools/qmv_tune_probe.py llama3-8b-int4 > gpurun_out/r2p_tune_llama.log 2>&1; echo "llama rc=$?"; grep "ms/step" gpurun_out/r2p_tune_llama.log
UZU_DECODE_PATH=kernels timeout -s KILL 200 python -u tools/qmv_tune_probe.py qwen3.5-0.8b-int4 > gpurun_out/r2p_tune_qwen.log 2>&1; echo "qwen rc=$?"; grep "ms/step" gpurun_out/r2p_tune_qwen.log
