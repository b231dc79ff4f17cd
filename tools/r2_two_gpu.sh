#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_tp_gpu.py -q > gpurun_out/r2j_tp_tests.log 2>&1; echo "tp tests rc=$?"; tail -n 3 gpurun_out/r2j_tp_tests.log
timeout -s KILL 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 > gpurun_out/r2j_bench_n2.json 2> gpurun_out/r2j_bench_n2.err; echo "bench n2 rc=$?"
cat gpurun_out/r2j_bench_n2.json; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r2j_bench_n2.err | tail -n 8
timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --impl reference --steps 4 --warmup 1 > gpurun_out/r2j_ref_n2.json 2> gpurun_out/r2j_ref_n2.err; echo "ref n2 rc=$?"; cat gpurun_out/r2j_ref_n2.json
