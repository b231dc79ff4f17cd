"""Tensor-parallel path of the engine on the GPU. With one GPU the whole TP data path except the NCCL calls is exercised
through a size-1 shard: row-parallel projections emit f32 partials that uzu_tp_all_reduce rounds to bf16 once, the readout goes through
uzu_tp_all_gather. With >= 2 GPUs two ranks (one process per GPU, NCCL) run the sharded model against the unsharded oracle."""
import ctypes as C
import json
import os
import subprocess
import sys
import textwrap
from pathlib import Path

import numpy as np
import pytest

from oracle.model import OracleModel
from tests.test_engine_gpu import _logit_check
from tests.util import bf16_to_f32, f32_to_bf16
from uzu_b200 import binding as B
from uzu_b200 import synth, tp

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def test_collective_encodes_on_a_single_rank_context(ctx):
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(1000 + 3) * 3).astype(np.float32)            # not a multiple of 4: scalar tail
    src, dst = ctx.upload(x), ctx.upload(np.zeros(x.size, np.uint16))
    with ctx.command_buffer("ar") as cmd:
        cmd.encode("uzu_tp_all_reduce_encode", src.ptr, x.size, dst.ptr)
    assert (dst.numpy(np.uint16, x.shape) == f32_to_bf16(x)).all()        # one RNE rounding, no sum on one rank
    lg = rng.integers(0, 65535, (3, 40)).astype(np.uint16)
    a, b, s = ctx.upload(lg), ctx.upload(np.zeros_like(lg)), ctx.upload(np.zeros_like(lg))
    args = B.TpAllGatherArgs(src=a.ptr, dst=b.ptr, scratch=s.ptr, rows=3, cols_local=40)
    with ctx.command_buffer("ag") as cmd:
        cmd.encode("uzu_tp_all_gather_encode", C.byref(args))
    assert (b.numpy(np.uint16, lg.shape) == lg).all()
    assert ctx.lib.uzu_context_tp_size(ctx.h) == 1 and ctx.lib.uzu_context_tp_rank(ctx.h) == 0


@pytest.mark.parametrize("fused", [False, True])
def test_size1_shard_runs_the_tp_data_path(ctx, tmp_path, fused):
    spec = synth.tiny("llama-512")
    full = synth.write_model(spec, tmp_path / "full", seed=21)
    shard = tp.shard_checkpoint(full, tmp_path / "r0", 0, 1)
    ref = OracleModel(full, max_context=128)
    rng = np.random.default_rng(5)
    prompt = rng.integers(0, spec.vocab_size, 70)                          # >= 64 rows: the prefill goes through the tensor-core GEMM
    with B.Engine(ctx, shard, max_context_length=128, use_cuda_graph=False, fused_decode=fused) as eng:
        _logit_check(eng.forward(prompt), ref.forward(prompt), "size-1 shard prefill")
        tok = 3
        for step in range(4):
            lr, lg = ref.forward([tok]), eng.forward([tok])
            _logit_check(lg, lr, f"size-1 shard decode {step}")
            tok = int(np.argmax(bf16_to_f32(lr[0])))
    # a shard needs the matching communicator: a 2-rank shard on this 1-rank context is rejected, not mis-run
    two = tp.shard_checkpoint(full, tmp_path / "r0of2", 0, 2)
    with pytest.raises(B.UzuError, match="same rank / size"):
        B.Engine(ctx, two, max_context_length=128, tp_rank=0, tp_size=2)


def _gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:
        return 0


@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("exchange", ["nccl", "p2p"])
def test_two_rank_nccl_engine_matches_unsharded_oracle(tmp_path, exchange):
    spec = synth.tiny("llama-512")
    full = synth.write_model(spec, tmp_path / "full", seed=22)
    for r in range(2):
        tp.shard_checkpoint(full, tmp_path / f"rank{r}", r, 2)
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import sys, os, json
        sys.path.insert(0, {str(ROOT)!r})
        import numpy as np, torch, torch.distributed as dist
        from uzu_b200 import binding as B
        rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
        torch.cuda.set_device(lr)
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.tensor(list(B.tp_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        ctx = B.Context(lr)
        ctx.tp_init(rank, world, bytes(idt.cpu().tolist()))
        if os.environ.get("UZU_TP_P2P"):          # second run of this test: decode all-reduces over peer memory instead of NCCL
            mine = torch.tensor(list(ctx.tp_p2p_export(16 * {spec.model_dim})), dtype=torch.uint8, device="cuda")
            allh = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allh, mine)
            ctx.tp_p2p_import([bytes(h.cpu().tolist()) for h in allh])
            dist.barrier()
        out = {{}}
        for graph in (False, True):
            with B.Engine(ctx, {str(tmp_path)!r} + f"/rank{{rank}}", max_context_length=256, use_cuda_graph=graph, tp_rank=rank, tp_size=world) as eng:
                prompt = (np.arange(80) * 37 % {spec.vocab_size}).astype(np.uint32)
                lg = eng.forward(prompt)
                lg2 = eng.forward([5])
                eng.reset()
                toks = eng.generate(prompt, 12)
                out[graph] = (lg, lg2, toks)
        assert out[False][2] == out[True][2], "graph replay with captured NCCL collectives == eager"
        gathered = [None] * world
        dist.all_gather_object(gathered, out[True][2])
        assert gathered[0] == gathered[1], "every rank samples the same tokens"
        if rank == 0:
            from oracle.model import OracleModel
            ref = OracleModel({str(full)!r}, max_context=256)
            f = lambda h: (np.asarray(h).astype(np.uint32) << 16).view(np.float32)
            r1 = ref.forward(prompt); r2 = ref.forward([5])
            err = max(float(np.abs(f(out[False][0]) - f(r1)).max()), float(np.abs(f(out[False][1]) - f(r2)).max()))
            print(json.dumps({{"err": err, "scale": float(np.abs(f(r1)).max()), "tokens": [int(t) for t in out[True][2]]}}))
        ctx.close()
        dist.barrier(); dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29621")
    env.pop("UZU_TP_P2P", None)
    if exchange == "p2p":
        env["UZU_TP_P2P"] = "1"
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29621", str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-4000:])
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert d["err"] <= 0.02 * d["scale"] + 1e-3 and len(d["tokens"]) == 12, d


@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_contexts_on_two_devices_in_one_process():
    """cudaFuncAttributeMaxDynamicSharedMemorySize is per (function, device): a second context on another GPU of the same process must be able to
    launch the > 48 KB shared-memory GEMV too (round 1 kept a per-process flag: the second device's launches failed)."""
    from oracle import oracle as O
    from tests import gpu_ops as G
    rng = np.random.default_rng(3)
    n, k = 2048, 4096
    w = rng.integers(0, 256, (n, k // 2), dtype=np.uint8)
    sc = O.f32_to_bf16(rng.uniform(0.01, 0.3, (n, k // 64)).astype(np.float32))
    zp = rng.integers(0, 256, (n, k // 128), dtype=np.uint8)
    x = O.f32_to_bf16(rng.uniform(-0.3, 0.3, (1, k)).astype(np.float32))
    ref = O.matmul(x, w, m=1, n=n, k=k, scales=sc, zero_points=zp, method=O.QM_ZERO_POINT, d_f32=True)
    outs = []
    ctxs = [B.Context(0), B.Context(1)]
    try:
        for c in ctxs + ctxs[::-1]:              # interleave the devices on one host thread
            outs.append(G.matmul(c, x, w, m=1, n=n, k=k, scales=sc, zero_points=zp, method=O.QM_ZERO_POINT, d_f32=True))
    finally:
        for c in ctxs:
            c.close()
    for got in outs:
        assert np.allclose(got, ref, rtol=1e-3, atol=1e-3)
