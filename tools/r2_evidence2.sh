#!/bin/bash
mkdir -p gpurun_out
UZU_DECODE_PATH=persistent timeout -s KILL 300 python bench.py --no-secondary --no-cpu-baseline > gpurun_out/r2m_bench_llama_persistent.json 2> gpurun_out/r2m_bench_llama_persistent.err; echo "persistent bench rc=$?"; cut -c1-900 gpurun_out/r2m_bench_llama_persistent.json
timeout -s KILL 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none --csv --log-file gpurun_out/r2m_launches_batch8_int8.csv python tools/batch_short.py llama3-8b-int8 1024 2 8 > gpurun_out/r2m_ncu_batch.log 2>&1; echo "ncu batch rc=$?"; tail -n 2 gpurun_out/r2m_ncu_batch.log; wc -l gpurun_out/r2m_launches_batch8_int8.csv
