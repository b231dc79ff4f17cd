#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 150 python -u tools/mega_probe.py qwen3.5-0.8b-int4 > gpurun_out/r2d_probe_qwen.log 2>&1; echo "qwen rc=$?"; cat gpurun_out/r2d_probe_qwen.log | tail -20
timeout -s KILL 200 python -u tools/mega_probe.py llama3-8b-int4 > gpurun_out/r2d_probe_llama.log 2>&1; echo "llama rc=$?"; cat gpurun_out/r2d_probe_llama.log | tail -20
