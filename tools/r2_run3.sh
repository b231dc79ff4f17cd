#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 100 python -u tools/mega_probe.py qwen3.5-0.8b-int4 > gpurun_out/r2g_probe_qwen.log 2>&1; echo "qwen rc=$?"; grep -E "steps |trace cta 0|gemv|attn|dnupd|act " gpurun_out/r2g_probe_qwen.log | head -12
timeout -s KILL 120 python -u tools/mega_probe.py llama3-8b-int4 > gpurun_out/r2g_probe_llama.log 2>&1; echo "llama rc=$?"; grep -E "steps |trace cta 0|gemv|attn|act " gpurun_out/r2g_probe_llama.log | head -12
timeout -s KILL 480 python bench.py > gpurun_out/r2g_bench_default.json 2> gpurun_out/r2g_bench_default.err; echo "bench rc=$?"
cat gpurun_out/r2g_bench_default.json; tail -n 12 gpurun_out/r2g_bench_default.err
