#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:decode_mega -s 3 -c 1 -f -o gpurun_out/r2_mega_qwen_v2 python tools/mega_short.py qwen3.5-0.8b-int4 512 6 > gpurun_out/r2_ncu1.log 2>&1; echo "ncu rc=$?"; tail -n 5 gpurun_out/r2_ncu1.log; ls -la gpurun_out/*.ncu-rep
