#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 500 ncu --section WarpStateStats --section SourceCounters --section SpeedOfLight --section MemoryWorkloadAnalysis --clock-control none --import-source on -k regex:decode_mega -s 3 -c 1 -f -o gpurun_out/r2_mega_llama_v3 python tools/mega_short.py llama3-8b-int4 2048 5 > gpurun_out/r2_ncu2.log 2>&1; echo "ncu rc=$?"; tail -n 4 gpurun_out/r2_ncu2.log; ls -la gpurun_out/r2_mega_llama_v3.ncu-rep
