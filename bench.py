#!/usr/bin/env python
"""bench.py -- decode tokens/s of the uzu transformer decode hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

A "step" is one decoded token: one full forward pass (all layers + readout + sampling) at batch 1 over the
synthetic uzu-format checkpoint of the workload, starting right after a `prefill`-token prompt. Default (N = 1)
workload = BASELINE.json configs[1]: Qwen3.5-0.8B int4, prefill 512, decode 128.

  value      decode tokens/s with everything resident in HBM: K device-chained steps (CUDA-graph replay, sampled token fed
             back on the device) between two CUDA events on the engine's stream; max over ranks; whole-job (sum of replicas).
  e2e        the same metric through the host-facing call: every step copies the input token host->device from pinned
             memory, runs the pass, and reads the sampled token device->host (uzu_engine_step_host), wall-clock timed.
  roofline   the weight-streaming fused dequant+GEMV (the dominant kernel): all GEMV launches of one token replayed back
             to back between CUDA events; achieved = algorithmic weight bytes per token / that time.
  cpu_baseline  the CPU oracle (C restatement of the reference CPU backend) decoding a few tokens of the same model on
             this box's host cores, single worker thread like the reference (rank 0, N = 1 only).
  --impl reference   times the same CPU restatement with all host threads (the reference itself is Rust and cannot be
             built here: no rustc/cargo in the image).
Multi-GPU: the BASELINE models that fit one GPU run as independent replicas (weak scaling, no collective on the path).
`--tp P` (with --gpus P) runs ONE model sharded over the P ranks instead (uzu_b200/tp.py shards by attention head / FFN column /
vocabulary row; two NCCL all-reduces per layer + one all-gather of the logits): value = tokens/s of that single replica, "strong".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (preset, kwargs, prefill, decode)
    "qwen3.5-0.8b-int4": ("qwen3.5-0.8b", dict(bits=4, group_size=64, hybrid=True), 512, 128),
    "qwen3.5-0.8b-int4-dense": ("qwen3.5-0.8b", dict(bits=4, group_size=64, hybrid=False), 512, 128),
    "llama3-8b-int4": ("llama3-8b", dict(bits=4, group_size=64), 2048, 256),
    "llama3-8b-int8": ("llama3-8b", dict(bits=8, group_size=64), 4096, 512),
    "llama3-70b-int4": ("llama3-70b", dict(bits=4, group_size=64), 2048, 256),      # needs --tp 8 (37 GB of weights, 4.7 GB per rank)
    "tiny": (None, {}, 32, 16),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def model_dir_for(workload: str, seed: int = 0) -> Path:
    from uzu_b200 import synth
    preset, kw, _, _ = WORKLOADS[workload]
    spec = synth.tiny("llama") if preset is None else synth.PRESETS[preset](**kw)
    base = Path(os.environ.get("UZU_MODEL_CACHE", "/dev/shm/uzu_b200_models"))
    d = base / f"{spec.name}-seed{seed}"
    done = d / ".done"
    if done.exists():
        return d
    base.mkdir(parents=True, exist_ok=True)
    lock = base / f"{spec.name}-seed{seed}.lock"
    try:
        fd = os.open(lock, os.O_CREAT | os.O_EXCL | os.O_WRONLY)
    except FileExistsError:
        t0 = time.time()
        while not done.exists():          # another rank is generating
            time.sleep(0.5)
            if time.time() - t0 > 1800:
                raise RuntimeError("timed out waiting for the synthetic checkpoint")
        return d
    try:
        t0 = time.time()
        synth.write_model(spec, d, seed=seed)
        done.write_text("ok")
        log(f"[bench] wrote synthetic checkpoint {d} in {time.time() - t0:.1f}s")
    finally:
        os.close(fd)
        os.unlink(lock)
    return d


def shard_dir_for(full_dir: Path, rank: int, size: int) -> Path:
    """The rank's tensor-parallel shard of a synthetic checkpoint (written once per rank, next to the full model)."""
    from uzu_b200 import tp as tpmod
    d = full_dir.parent / f"{full_dir.name}-tp{size}-rank{rank}"
    done = d / ".done"
    if not done.exists():
        t0 = time.time()
        tpmod.shard_checkpoint(full_dir, d, rank, size)
        done.write_text("ok")
        log(f"[bench] wrote tensor-parallel shard {d} in {time.time() - t0:.1f}s")
    return d


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device: int):
        self.device, self.proc = device, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, smax, reasons, power = [], [], set(), []
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def max_over_ranks(dist, seconds: float, device: str) -> float:
    """Timing contract: every number reported is the MAX over ranks of the device-timed region."""
    if dist is None:
        return seconds
    import torch
    t = torch.tensor([seconds], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def whole_job_value(world: int, steps: int, seconds_max: float) -> float:
    """Replicas (weak scaling): every rank decodes `steps` tokens; whole-job throughput = all tokens / slowest rank."""
    return world * steps / seconds_max


def hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json copy bandwidth)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def tensor_peak():
    """Dense bf16 TFLOP/s to divide the prefill GEMM by: the sustained cuBLAS figure of MEASURED_PEAKS.json (a GEMM timed inside a pass)."""
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json cuBLAS bf16, sustained)"
        except Exception:
            pass
    return 1500.0, "fallback (B200_PROFILING.md)"


def cpu_baseline(model_dir: Path, threads: int, tokens: int, prompt_len: int = 4):
    """Decode `tokens` tokens with the CPU oracle; returns (tokens/s, description)."""
    from oracle.model import OracleModel
    m = OracleModel(model_dir, threads=threads, max_context=256)
    rng = np.random.default_rng(0)
    prompt = rng.integers(0, m.V, prompt_len)
    m.prefill(prompt)               # untimed: fills a few context positions
    tok = 1                         # any valid token id; the arithmetic per token does not depend on it
    t0 = time.perf_counter()
    for _ in range(tokens):
        logits = m.forward([tok])
    dt = time.perf_counter() - t0
    return tokens / dt, f"{tokens} decode tokens of the full model at context ~{prompt_len + tokens} (prompt {prompt_len} tokens untimed)"


def run_reference(args, rank: int):
    """--impl reference: the reference's CPU path (C restatement; no Rust toolchain here) on all host threads."""
    if rank != 0:
        return
    from oracle import oracle as O
    workload = args.workload
    mdir = model_dir_for(workload)
    threads = O.lib().oracle_max_threads()
    from oracle.model import OracleModel
    m = OracleModel(mdir, threads=threads, max_context=256)
    rng = np.random.default_rng(0)
    prompt = rng.integers(0, m.V, 4)
    m.prefill(prompt)
    tok = 1
    t_probe0 = time.perf_counter()
    m.forward([tok])
    t_tok = time.perf_counter() - t_probe0
    budget = 150.0
    total = args.warmup + args.steps
    timed = args.steps if total * t_tok <= budget else max(1, int(budget / t_tok) - args.warmup)
    warm = args.warmup if total * t_tok <= budget else max(0, min(args.warmup, 1))
    for _ in range(warm):
        m.forward([tok])
    t0 = time.perf_counter()
    for _ in range(timed):
        m.forward([tok])
    dt = time.perf_counter() - t0
    value = timed / dt
    sample = f"{timed} of {args.steps} decode steps timed (bounded to ~{budget:.0f}s of CPU work), context ~{4 + warm + timed}, prompt 4 tokens"
    line = {
        "impl": "reference", "metric": "decode tokens/sec/GPU (int4)", "value": value, "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16 activations x int4 weights, f32 accumulate", "data": "synthetic",
        "config": {"workload": workload, "note": "reference CPU backend arithmetic restated in C (oracle/uzu_oracle.c); the Rust reference "
                   "cannot be built here (no rustc/cargo). OpenMP over output columns = not the reference's single-thread execution model."},
        "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args, rank: int, world: int, local_rank: int):
    from uzu_b200 import binding as B
    from uzu_b200 import build as ubuild
    if not B.LIB_PATH.exists():
        ubuild.build()
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod

    workload = args.workload
    _, _, prefill_default, decode_default = WORKLOADS[workload]
    prefill = args.prefill or prefill_default
    K, W = args.steps or decode_default, args.warmup
    mdir = model_dir_for(workload)
    max_ctx = max(1024, prefill + K + W + 64)

    ctx = B.Context(local_rank)
    tp = args.tp if args.tp > 1 else 1
    if tp > 1:
        # one tensor-parallel group over all ranks: rank 0's NCCL unique id travels over torch.distributed, every rank loads its shard
        if world != tp:
            raise SystemExit(f"--tp {tp} needs exactly {tp} ranks (torchrun --nproc-per-node {tp}); got WORLD_SIZE={world}")
        import torch
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.tensor(list(B.tp_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        ctx.tp_init(rank, world, bytes(idt.cpu().tolist()))
        mdir = shard_dir_for(mdir, rank, world)
        if os.environ.get("UZU_TP_P2P"):
            # opt-in: decode-sized all-reduces through the one-kernel peer-memory exchange (CUDA IPC handles travel over torch.distributed)
            model_dim = json.loads((mdir / "config.json").read_text())["decoder_config"]["transformer_config"]["model_dim"]
            mine = torch.tensor(list(ctx.tp_p2p_export(16 * model_dim)), dtype=torch.uint8, device="cuda")
            allh = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allh, mine)
            ctx.tp_p2p_import([bytes(h.cpu().tolist()) for h in allh])
            dist.barrier()
    eng = B.Engine(ctx, mdir, max_context_length=max_ctx, use_cuda_graph=not args.no_graph, fused_decode=not args.no_fused,
                   tp_rank=rank if tp > 1 else 0, tp_size=tp)
    replicas = 1 if tp > 1 else world
    if args.batch > 1:
        run_batched(args, eng, ctx, dist, rank, world, local_rank, workload, prefill, K, W)
        return
    info = eng.info
    rng = np.random.default_rng(0)
    prompt = rng.integers(0, info.vocab_size, prefill).astype(np.uint32)

    t0 = time.perf_counter()
    first = eng.prefill(prompt)
    prefill_s = time.perf_counter() - t0
    eng.snapshot()

    # ---- device-resident decode (value) ----
    eng.decode_timed(max(W, 3))
    eng.restore()
    sampler = ClockSampler(local_rank)
    sampler.start()
    if dist:
        import torch
        dist.barrier()
        torch.cuda.synchronize()
    launches0 = eng.launch_count
    seconds = eng.decode_timed(K)
    launches = eng.launch_count - launches0
    if dist:
        import torch
        torch.cuda.synchronize()
        seconds = max_over_ranks(dist, seconds, "cuda")
        dist.barrier()
    clocks = sampler.stop()
    value = whole_job_value(replicas, K, seconds)

    # ---- end to end through host buffers ----
    eng.restore()
    tok = first
    for _ in range(max(W, 3)):
        tok = eng.step_host(tok)
    eng.restore()
    tok = first
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        tok = eng.step_host(tok)
    e2e_s = time.perf_counter() - t0
    e2e_s = max_over_ranks(dist, e2e_s, "cuda")
    e2e_value = whole_job_value(replicas, K, e2e_s)

    # ---- roofline of the dominant kernel (fused dequant + GEMV) ----
    iters = 10
    lin_s, lin_launches = eng.time_linears(iters)
    gemv_s_per_token = lin_s / iters
    peak, peak_src = hbm_peak()
    achieved = info.weight_bytes_per_token / gemv_s_per_token / 1e9
    traffic = None
    tfile = ROOT / "profiles" / "roofline_traffic.json"
    if tfile.exists():
        try:
            traffic = json.loads(tfile.read_text()).get(workload)
        except Exception:
            traffic = None
    # ---- prefill GEMM (tcgen05 tensor cores, in-kernel dequant): every linear of one prefill pass, useful TFLOP/s ----
    prefill_gemm = None
    try:
        pm = min(prefill, 1024)
        if pm >= 64:
            p_s, p_flops = eng.time_prefill_linears(pm, 3)
            tpeak, tsrc = tensor_peak()
            prefill_gemm = {"rows": pm, "ms_per_pass": 1000.0 * p_s, "useful_tflops": p_flops / p_s / 1e12, "peak_tflops": tpeak,
                            "frac": p_flops / p_s / 1e12 / tpeak, "peak_source": tsrc,
                            "note": "useful flops = 2*m*N*K; the kernel issues 2x that (exact hi/lo bf16 weight planes)"}
    except Exception as ex:  # must not take the decode measurement down
        prefill_gemm = {"error": str(ex)[:200]}
    ctx_mid = prefill + K / 2
    bytes_per_token = info.weight_bytes_per_token + info.kv_bytes_per_token_per_ctx * ctx_mid + info.state_bytes_per_token

    line = {
        "metric": "decode tokens/sec/GPU (int4)" if "int4" in workload else "decode tokens/sec/GPU",
        "value": value, "unit": "tokens/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": 1000.0 * seconds / K,
        "higher_is_better": True, "scaling": "strong" if tp > 1 else "weak", "vs_baseline": None,
        "dtype": "bf16 activations x int4 weights, f32 accumulate" if "int4" in workload else "bf16 activations x int8 weights, f32 accumulate",
        "data": "synthetic",
        "config": {
            "workload": f"{workload}: prefill {prefill}, decode {K}, batch 1, greedy", "parallelism": f"tp{tp}" if tp > 1 else ("replicas" if world > 1 else "single"),
            "layers": info.num_layers, "attention_layers": info.num_attention_layers, "delta_net_layers": info.num_delta_net_layers,
            "model_dim": info.model_dim, "vocab": info.vocab_size, "weight_bytes_per_token": info.weight_bytes_per_token,
            "kv_bytes_per_token_at_mid_ctx": int(info.kv_bytes_per_token_per_ctx * ctx_mid), "state_bytes_per_token": info.state_bytes_per_token,
            "cache_policy": f"inputs larger than L2: {info.weight_bytes_per_token / 1e6:.0f} MB of weights streamed every step vs 126 MB L2",
            "cuda_graph": not args.no_graph, "fused_decode": not args.no_fused, "prefill_tokens_per_s": prefill / prefill_s,
            "whole_step_hbm_frac": bytes_per_token * (K / seconds) / 1e9 / peak,
            "gemv_launches_per_token": lin_launches // iters, "gemv_ms_per_token": 1000.0 * gemv_s_per_token,
            "prefill_gemm": prefill_gemm,
        },
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "kernel": "qmv_kernel (fused int4 dequant + GEMV, all linears + readout of one token)", "peak_source": peak_src},
        "e2e": {"value": e2e_value, "unit": "tokens/s", "h2d_bytes_per_step": 4, "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            v, sample = cpu_baseline(mdir, threads=1, tokens=3 if workload.startswith("qwen") else 1)
            line["cpu_baseline"] = {"value": v, "unit": "tokens/s", "cores": 1, "kind": "port", "sample": sample}
        except Exception as ex:  # the baseline must not take the measurement down
            line["cpu_baseline"] = {"value": None, "unit": "tokens/s", "cores": 1, "kind": "port", "sample": f"failed: {ex}"}
    eng.close()
    ctx.close()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


def run_batched(args, eng, ctx, dist, rank, world, local_rank, workload, prefill, K, W):
    """--batch B: B independent sequences decoded together (uzu_engine_batch_*: one pass over the weights per step for all of them).
    value = tokens/s summed over the B sequences (and over replicas); a "step" is one batched step = B tokens."""
    nseq = args.batch
    info = eng.info
    rng = np.random.default_rng(0)
    eng.batch_begin(nseq)
    firsts = [eng.batch_prefill(b, rng.integers(0, info.vocab_size, prefill).astype(np.uint32)) for b in range(nseq)]
    eng.batch_decode_timed(firsts, max(W, 3))
    sampler = ClockSampler(local_rank)
    sampler.start()
    if dist:
        import torch
        dist.barrier()
        torch.cuda.synchronize()
    launches0 = eng.launch_count
    seconds = eng.batch_decode_timed(firsts, K)
    launches = eng.launch_count - launches0
    seconds = max_over_ranks(dist, seconds, "cuda")
    clocks = sampler.stop()
    value = whole_job_value(world, K * nseq, seconds)
    toks = list(firsts)
    for _ in range(3):
        toks = eng.batch_step(toks)
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        toks = eng.batch_step(toks)              # H2D of the B input tokens, the step, D2H of the B sampled tokens
    e2e_s = max_over_ranks(dist, time.perf_counter() - t0, "cuda")
    peak, peak_src = hbm_peak()
    line = {
        "metric": "decode tokens/sec/GPU (int4)" if "int4" in workload else "decode tokens/sec/GPU",
        "value": value, "unit": "tokens/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": 1000.0 * seconds / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16 activations x int4 weights, f32 accumulate" if "int4" in workload else "bf16 activations x int8 weights, f32 accumulate",
        "data": "synthetic",
        "config": {"workload": f"{workload}: prefill {prefill}, decode {K}, batch {nseq} (independent sequences, one weight pass per step), greedy",
                   "parallelism": "replicas" if world > 1 else "single", "batch": nseq, "weight_bytes_per_token": info.weight_bytes_per_token,
                   "cache_policy": f"inputs larger than L2: {info.weight_bytes_per_token / 1e6:.0f} MB of weights streamed every step vs 126 MB L2",
                   "cuda_graph": False},
        "roofline": {"bound": "hbm", "achieved": info.weight_bytes_per_token * (K / seconds) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": info.weight_bytes_per_token * (K / seconds) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                     "kernel": "whole batched step (weights streamed once per step for all sequences)"},
        "e2e": {"value": whole_job_value(world, K * nseq, e2e_s), "unit": "tokens/s", "h2d_bytes_per_step": 4 * nseq, "d2h_bytes_per_step": 4 * nseq},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    eng.close()
    ctx.close()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="timed decode steps (default: the workload's decode length)")
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", default="qwen3.5-0.8b-int4", choices=sorted(WORKLOADS))
    ap.add_argument("--prefill", type=int, default=0)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--fused", action="store_true", help="(default) fold norm / gated-act / sigmoid-gate launches into the neighbouring GEMV")
    ap.add_argument("--no-fused", action="store_true", help="encode the reference's kernel sequence one launch per kernel")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch", type=int, default=1, help="decode this many independent sequences together (<= 16), one weight pass per step")
    ap.add_argument("--tp", type=int, default=1, help="tensor-parallel size: shard ONE model over this many ranks (= --gpus) instead of replicas")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if not args.steps:
            args.steps = WORKLOADS[args.workload][3]
        run_reference(args, rank)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
