#!/bin/bash
mkdir -p gpurun_out
# (1) launch list + DRAM bytes of one Llama-3-8B decode step on the default (per-kernel) path at context ~2300
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1200 -c 600 --csv --log-file gpurun_out/r2k_launches_llama_ctx2300.csv python tools/decode_short.py llama3-8b-int4 2296 6 kernels > gpurun_out/r2k_ncu_llama.log 2>&1; echo "ncu llama rc=$?"; tail -n 2 gpurun_out/r2k_ncu_llama.log
# (2) the same for the persistent kernel (one launch = one token)
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:decode_mega -s 2 -c 2 --csv --log-file gpurun_out/r2k_launches_llama_persistent.csv python tools/decode_short.py llama3-8b-int4 2296 5 persistent > gpurun_out/r2k_ncu_llama_mega.log 2>&1; echo "ncu mega rc=$?"
# (3) config 4: Llama-3-8B int8, 8 sequences sharing the weight pass
timeout -s KILL 400 python bench.py --workload llama3-8b-int8 --batch 8 --steps 128 > gpurun_out/r2k_bench_llama_int8_batch8.json 2> gpurun_out/r2k_bench_llama_int8_batch8.err; echo "batch8 rc=$?"; cat gpurun_out/r2k_bench_llama_int8_batch8.json; tail -n 3 gpurun_out/r2k_bench_llama_int8_batch8.err
