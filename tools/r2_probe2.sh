#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 120 python -u tools/mega_probe.py qwen3.5-0.8b-int4 > gpurun_out/r2e_probe_qwen.log 2>&1; echo "qwen rc=$?"; tail -n 16 gpurun_out/r2e_probe_qwen.log
timeout -s KILL 150 python -u tools/mega_probe.py llama3-8b-int4 > gpurun_out/r2e_probe_llama.log 2>&1; echo "llama rc=$?"; tail -n 14 gpurun_out/r2e_probe_llama.log
