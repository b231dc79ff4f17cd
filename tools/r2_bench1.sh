#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python bench.py > gpurun_out/r2c_bench_default.json 2> gpurun_out/r2c_bench_default.err; echo "bench rc=$?"
timeout -s KILL 200 python bench.py --workload qwen3.5-0.8b-int4 --no-persistent --no-cpu-baseline > gpurun_out/r2c_bench_qwen_perkernel.json 2> gpurun_out/r2c_bench_qwen_perkernel.err; echo "qwen per-kernel rc=$?"
cat gpurun_out/r2c_bench_default.json; tail -n 5 gpurun_out/r2c_bench_default.err; cat gpurun_out/r2c_bench_qwen_perkernel.json
