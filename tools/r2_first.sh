#!/bin/bash
# round-2 first hardware run: everything written at the end of round 1 that never ran on a GPU
mkdir -p gpurun_out
export UZU_TEST_PREFILL_ATTN=1 UZU_TEST_DELTA_PREFILL=1 UZU_TEST_BATCH_DECODE=1 UZU_TEST_TP_NCCL=1
nvidia-smi -L > gpurun_out/r2a_gpus.txt
nvidia-smi topo -m >> gpurun_out/r2a_gpus.txt 2>&1
timeout 400 python -m pytest tests/test_tp_gpu.py -q -x -k "nccl" > gpurun_out/r2a_tp_nccl.log 2>&1; echo "tp nccl rc=$?"
timeout 300 python -m pytest tests/test_tp_gpu.py -q -x -k "p2p" > gpurun_out/r2a_tp_p2p.log 2>&1; echo "tp p2p rc=$?"
timeout 300 python -m pytest tests/test_prefill_attention_gpu.py -q > gpurun_out/r2a_pattn.log 2>&1; echo "pattn rc=$?"
timeout 300 python -m pytest tests/test_delta_prefill_gpu.py -q > gpurun_out/r2a_dprefill.log 2>&1; echo "dprefill rc=$?"
timeout 300 python -m pytest tests/test_batch_decode_gpu.py -q > gpurun_out/r2a_batch.log 2>&1; echo "batch rc=$?"
timeout 500 python bench.py --workload llama3-8b-int4 --no-cpu-baseline > gpurun_out/r2a_llama_1gpu.json 2> gpurun_out/r2a_llama_1gpu.err; echo "llama rc=$?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 2 --tp 2 --workload llama3-8b-int4 --no-cpu-baseline > gpurun_out/r2a_llama_tp2.json 2> gpurun_out/r2a_llama_tp2.err; echo "tp2 rc=$?"
UZU_TP_P2P=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --tp 2 --workload llama3-8b-int4 --no-cpu-baseline > gpurun_out/r2a_llama_tp2_p2p.json 2> gpurun_out/r2a_llama_tp2_p2p.err; echo "tp2 p2p rc=$?"
tail -3 gpurun_out/r2a_*.log; cat gpurun_out/r2a_llama*.json
