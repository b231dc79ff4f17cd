#!/bin/bash
mkdir -p gpurun_out
for v in "" _sm2 _sm0; do
  UZU_B200_LIB=$PWD/uzu_b200/lib/libuzu_b200$v.so timeout -s KILL 120 python -u tools/mega_probe.py llama3-8b-int4 > gpurun_out/r2h_probe_llama$v.log 2>&1; echo "llama$v rc=$?"; grep -E "64 steps|trace cta 0|gemv|attn " gpurun_out/r2h_probe_llama$v.log | head -6
done
