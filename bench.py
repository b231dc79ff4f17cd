#!/usr/bin/env python
"""bench.py -- decode tokens/s of the uzu transformer decode hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference] [--tp P]

A "step" is one decoded token: one full forward pass (all layers + readout + sampling) at batch 1 over the synthetic uzu-format
checkpoint of the workload, starting right after a `prefill`-token prompt. Default workload = the configuration the "at 1/2/4/8 B200"
metric is quoted on: BASELINE.json configs[2], Llama-3-8B int4, prefill 2048, decode 256. The N = 1 run adds configs[1]
(Qwen3.5-0.8B int4, prefill 512, decode 128) as `config.secondary` with its own value / e2e / roofline / parity.

  value      decode tokens/s with everything resident in HBM: K device-chained steps (the sampled token is fed back on the device)
             between two CUDA events on the engine's stream; max over ranks; WHOLE JOB = summed over the N replicas
             (`config.per_gpu_value` = value / N).
  e2e        the same metric through the host-facing call: every step copies the input token host->device from pinned memory, runs
             the pass, and reads the sampled token device->host (uzu_engine_step_host), wall-clock timed.
  roofline   the dominant kernel. Default decode path = the persistent whole-token kernel (decode_mega_kernel, ONE launch per step):
             achieved = algorithmic bytes of a step (quantised weights + KV rows at the mid context + recurrent state) / the
             CUDA-event duration of a launch. Per-kernel path (--no-persistent, or models the persistent kernel does not cover): all
             fused dequant+GEMV launches of one token replayed back to back.
  parity     the timed engine against the CPU oracle on the same checkpoint (short prompt, 3 teacher-forced decode steps):
             greedy token ids and the last-row logit error (rank 0, N = 1).
  cpu_baseline  the CPU oracle (C restatement of the reference CPU backend) decoding a few tokens of the same model on this box's
             host cores, single worker thread like the reference (rank 0, N = 1 only).
  --impl reference   times the same CPU restatement on the host threads the cgroup really grants (the reference itself is Rust and
             cannot be built here: no rustc/cargo in the image).
Multi-GPU: the BASELINE models that fit one GPU run as independent replicas (weak scaling, no collective on the path). At N >= 2 the
run also shards ONE copy of the model over the N ranks (uzu_b200/tp.py: attention head / FFN column / vocabulary row; two all-reduces
per layer + one all-gather of the logits) and reports it under `config.tp` for both exchanges (NCCL, one-kernel peer memory).
`--tp P` (with --gpus P) makes that sharded model the headline instead: value = tokens/s of the single replica, "strong".
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


def effective_cpus() -> int:
    """Host threads this process may really use: the affinity mask capped by the cgroup CPU quota (a 128-way OpenMP team on an
    8-CPU cgroup runs slower than one thread)."""
    n = len(os.sched_getaffinity(0))
    quota = None
    try:
        txt = Path("/sys/fs/cgroup/cpu.max").read_text().split()
        if txt[0] != "max":
            quota = float(txt[0]) / float(txt[1])
    except Exception:
        try:
            q = float(Path("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read_text())
            per = float(Path("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read_text())
            if q > 0:
                quota = q / per
        except Exception:
            quota = None
    if quota:
        n = min(n, max(1, int(quota + 0.999)))
    return max(1, min(n, 64))


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
        m.forward([tok])
    dt = time.perf_counter() - t0
    return tokens / dt, f"{tokens} decode tokens of the full model at context ~{prompt_len + tokens} (prompt {prompt_len} tokens untimed)"


def bf16_to_f32(a):
    return (np.asarray(a, np.uint16).astype(np.uint32) << 16).view(np.float32)


def cpu_leg(eng, model_dir: Path, threads: int, budget_s: float, max_steps: int = 3):
    """ONE bounded CPU leg for both reported objects: the oracle decodes a 1-token prompt and up to `max_steps` more tokens of the
    full model (stops when `budget_s` of CPU time is spent; at least one decode step). Every oracle pass is timed (-> cpu_baseline,
    cores = the threads used) and the same passes are the parity oracle for the timed engine: teacher-forced with the oracle's tokens,
    greedy token ids and last-row logits compared step by step."""
    from oracle.model import OracleModel
    t_load = time.perf_counter()
    ref = OracleModel(model_dir, threads=threads, max_context=64)
    log(f"[bench] oracle loaded in {time.perf_counter() - t_load:.1f}s ({threads} threads)")
    rng = np.random.default_rng(1)
    prompt = rng.integers(0, ref.V, 1).astype(np.uint32)
    t0 = time.perf_counter()
    lr = ref.prefill(prompt)
    pass_times = [time.perf_counter() - t0]
    eng.reset()
    first = eng.prefill(prompt)
    want = int(np.argmax(bf16_to_f32(lr[0])))
    ids_gpu, ids_ref, worst, gaps = [int(first)], [want], 0.0, []
    tok = want
    for i in range(max_steps):
        if i > 0 and sum(pass_times) + pass_times[-1] > budget_s:
            break
        t1 = time.perf_counter()
        lr = ref.forward([tok])
        pass_times.append(time.perf_counter() - t1)
        got = eng.step_host(tok)
        lg = eng.last_logits()
        r, g = bf16_to_f32(lr[0]), bf16_to_f32(lg[0])
        scale = float(np.abs(r).max())
        worst = max(worst, float(np.abs(g - r).max()) / max(scale, 1e-30))
        top = np.sort(r)[::-1]
        gaps.append(float((top[0] - top[1]) / max(abs(top[0]), 1e-30)))
        tok = int(np.argmax(r))
        ids_gpu.append(int(got)); ids_ref.append(tok)
        log(f"[bench] oracle pass {i + 1}: {pass_times[-1]:.1f}s")
    eng.reset()
    decode_times = pass_times[1:]
    cpu = {"value": len(decode_times) / sum(decode_times), "unit": "tokens/s", "cores": threads, "kind": "port",
           "sample": f"{len(decode_times)} decode token(s) of the full model at context 1..{len(decode_times)} after a 1-token prompt "
                     f"(bounded to ~{budget_s:.0f}s of CPU work; OpenMP over output columns with {threads} thread(s), the reference runs 1 worker thread)"}
    parity = {"prompt_tokens": 1, "decode_steps_checked": len(decode_times), "token_ids_gpu": ids_gpu, "token_ids_oracle": ids_ref,
              "token_ids_equal": ids_gpu == ids_ref, "max_logit_err_over_range": worst, "top2_gap_over_range": gaps,
              "persistent_kernel": bool(eng.persistent_decode),
              "note": "teacher-forced with the oracle's tokens; logits are bf16, error = max|gpu - oracle| / max|oracle| over the vocabulary"}
    return cpu, parity


def run_reference(args, rank: int):
    """--impl reference: the reference's CPU path (C restatement; no Rust toolchain here) on the host threads this process may use."""
    if rank != 0:
        return
    workload = args.workload
    _, _, prefill_default, decode_default = WORKLOADS[workload]
    mdir = model_dir_for(workload)
    threads = effective_cpus()
    from oracle.model import OracleModel
    m = OracleModel(mdir, threads=threads, max_context=256)
    rng = np.random.default_rng(0)
    prompt = rng.integers(0, m.V, 4)
    m.prefill(prompt)
    tok = 1
    t_probe0 = time.perf_counter()
    m.forward([tok])
    t_tok = time.perf_counter() - t_probe0
    budget = 120.0
    total = args.warmup + args.steps
    timed = args.steps if total * t_tok <= budget else max(1, int(budget / t_tok) - 1)
    warm = args.warmup if total * t_tok <= budget else 1
    for _ in range(warm):
        m.forward([tok])
    t0 = time.perf_counter()
    for _ in range(timed):
        m.forward([tok])
    dt = time.perf_counter() - t0
    value = timed / dt
    sample = (f"{timed} of {args.steps} decode steps timed (bounded to ~{budget:.0f}s of CPU work) at context ~{5 + warm + timed} after a 4-token prompt: "
              f"the CPU path cannot prefill {args.prefill or prefill_default} tokens of this model in minutes; per-token arithmetic over the weights is "
              f"context-independent, only the (small) attention term is shorter than in the GPU arm")
    line = {
        "impl": "reference", "metric": metric_name(workload), "value": value, "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": dtype_name(workload), "data": "synthetic",
        "config": {"workload": workload_name(workload, args.prefill or prefill_default, args.steps, 1),
                   "note": "reference CPU backend arithmetic restated in C (oracle/uzu_oracle.c); the Rust reference cannot be built here "
                           "(no rustc/cargo). OpenMP over output columns = not the reference's single-thread execution model."},
        "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def metric_name(workload: str) -> str:
    return "decode tokens/sec (int4), summed over the N GPUs" if "int4" in workload else "decode tokens/sec (int8), summed over the N GPUs"


def dtype_name(workload: str) -> str:
    return "bf16 activations x int4 weights, f32 accumulate" if "int4" in workload else "bf16 activations x int8 weights, f32 accumulate"


def workload_name(workload: str, prefill: int, decode: int, batch: int) -> str:
    return f"{workload}: prefill {prefill}, decode {decode}, batch {batch}, greedy"


def tp_setup(ctx, dist, rank: int, world: int, full_dir: Path, p2p: bool):
    """One tensor-parallel group over all ranks: rank 0's NCCL unique id travels over torch.distributed; every rank loads its shard."""
    import torch
    from uzu_b200 import binding as B
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.tensor(list(B.tp_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    ctx.tp_init(rank, world, bytes(idt.cpu().tolist()))
    sdir = shard_dir_for(full_dir, rank, world)
    if p2p:
        model_dim = json.loads((sdir / "config.json").read_text())["decoder_config"]["transformer_config"]["model_dim"]
        mine = torch.tensor(list(ctx.tp_p2p_export(16 * model_dim)), dtype=torch.uint8, device="cuda")
        allh = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine)
        ctx.tp_p2p_import([bytes(h.cpu().tolist()) for h in allh])
        dist.barrier()
    return sdir


def measure_decode(eng, dist, world: int, replicas: int, local_rank: int, prompt, K: int, W: int):
    """Device-resident value, end-to-end value and launches of K decode steps after `prompt`; max over ranks."""
    t0 = time.perf_counter()
    first = eng.prefill(prompt)
    prefill_s = time.perf_counter() - t0
    eng.snapshot()
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
    # ---- end to end through host buffers: H2D of the input token, the step, D2H of the sampled token, every step ----
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
    e2e_s = max_over_ranks(dist, time.perf_counter() - t0, "cuda")
    eng.restore()
    return {"seconds": seconds, "launches": int(launches), "clocks": clocks, "prefill_s": prefill_s, "e2e_seconds": e2e_s,
            "value": whole_job_value(replicas, K, seconds), "e2e_value": whole_job_value(replicas, K, e2e_s)}


def both_paths(eng, steps: int = 48):
    """ms per decode step of both parity-tested decode paths on the current state (the engine auto-selects the faster one at load;
    this makes the choice visible): device-chained steps between CUDA events, state restored afterwards."""
    selected = bool(eng.persistent_decode)
    out = {"selected": "persistent kernel" if selected else "per-kernel path", "selection": eng.persistent_decode_reason}
    for name, mode in (("persistent_ms_per_step", True), ("per_kernel_ms_per_step", False)):
        try:
            eng.set_persistent_decode(mode)
            if bool(eng.persistent_decode) != mode:
                raise RuntimeError("mode not available")
            eng.restore()
            eng.decode_timed(4)
            eng.restore()
            out[name] = 1000.0 * eng.decode_timed(steps) / steps
        except Exception as ex:
            out[name] = None
            out[name + "_note"] = str(ex)[:120]
    try:
        eng.set_persistent_decode(selected)
    except Exception:
        pass
    eng.restore()
    return out


def roofline_of(eng, workload: str, prefill: int, K: int, seconds: float):
    """Dominant kernel of a decode step. Persistent mode: the step IS one kernel (decode_mega_kernel): algorithmic bytes per launch =
    quantised weights + KV rows read at the mid context + recurrent state read/write; duration = the CUDA-event time of the K launches / K.
    Per-kernel mode: all fused dequant+GEMV launches of one token replayed back to back."""
    info = eng.info
    peak, peak_src = hbm_peak()
    traffic = None
    tfile = ROOT / "profiles" / "roofline_traffic.json"
    if tfile.exists():
        try:
            traffic = json.loads(tfile.read_text()).get(f"{workload}:{'persistent' if eng.persistent_decode else 'per-kernel'}")
        except Exception:
            traffic = None
    ctx_mid = prefill + K / 2
    step_bytes = info.weight_bytes_per_token + info.kv_bytes_per_token_per_ctx * ctx_mid + info.state_bytes_per_token
    if eng.persistent_decode:
        achieved = step_bytes / (seconds / K) / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": "decode_mega_kernel (persistent whole-token kernel: every linear + attention + state update + argmax of one token in ONE launch)",
                "algorithmic_bytes_per_launch": int(step_bytes), "launch_us": 1e6 * seconds / K, "peak_source": peak_src}
        extra = {"gemv_launches_per_token": 0, "gemv_ms_per_token": None}
    else:
        iters = 10
        lin_s, lin_launches = eng.time_linears(iters)
        per_tok = lin_s / iters
        achieved = info.weight_bytes_per_token / per_tok / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": "qmv_decode_async_kernel (fused dequant + GEMV with TMA-bulk-fed rings from the decode-stream layout; all linears + readout of one token replayed back to back)",
                "algorithmic_bytes_per_launch": int(info.weight_bytes_per_token // max(1, lin_launches // iters)), "peak_source": peak_src}
        extra = {"gemv_launches_per_token": lin_launches // iters, "gemv_ms_per_token": 1000.0 * per_tok}
    extra["whole_step_hbm_frac"] = step_bytes * (K / seconds) / 1e9 / peak
    return roof, extra


def speculation_of(eng, step_ms: float, nodes: int = 16, passes: int = 6):
    """Cost of one speculation pass (uzu_engine_trie_pass over a linearized trie of `nodes` nodes: one sweep over the weights, logits + a
    sampled token for every node) next to the plain decode step of the timed run. Wall clock around the synchronous call,
    which ends with the D2H read of the sampled ids like the stream's speculation loop. `break_even_tokens_per_pass` = pass / step: a
    proposer has to get more than that many tokens accepted per pass for speculation to pay on this GPU."""
    if not eng.speculation_supported:
        return {"supported": False, "reason": "DeltaNet layers: no tree-verify core (Mixer::speculation_supported, delta_net.rs:442-444)"}
    from uzu_b200.trie import PRng, TrieNode
    rng = np.random.default_rng(1)
    V = eng.info.vocab_size
    eng.flush()
    root = TrieNode(int(rng.integers(0, V)), 0)
    node, made = root, 1
    while made < nodes:                      # a bushy trie: siblings and chains, the shape a draft model proposes
        child = TrieNode(int(rng.integers(0, V)), 0)
        if node.get(child.token) is not None:
            continue
        node.add(child)
        made += 1
        if made % 3 == 0:
            node = child
    flat = root.linearize()
    ms = []
    for _ in range(passes):
        t0 = time.perf_counter()
        sampled = eng.trie_pass(flat.token_ids(), flat.nodes())
        ms.append((time.perf_counter() - t0) * 1e3)
        eng.trie_accept([0], sampled[0])
    pass_ms = float(np.median(ms[1:]))
    return {"supported": True, "nodes": nodes, "pass_ms": pass_ms, "plain_step_ms": step_ms, "break_even_tokens_per_pass": pass_ms / step_ms,
            "note": "verify half only (no draft model in scope): the pass runs the unfused kernel sequence with the rows GEMV (2..16 rows per weight pass)"}


def prefill_gemm_of(eng, prefill: int):
    try:
        pm = min(prefill, 1024)
        if pm < 64:
            return None
        p_s, p_flops = eng.time_prefill_linears(pm, 3)
        tpeak, tsrc = tensor_peak()
        return {"rows": pm, "ms_per_pass": 1000.0 * p_s, "useful_tflops": p_flops / p_s / 1e12, "peak_tflops": tpeak,
                "frac": p_flops / p_s / 1e12 / tpeak, "peak_source": tsrc,
                "note": "useful flops = 2*m*N*K; the kernel issues 2x that (exact hi/lo bf16 weight planes)"}
    except Exception as ex:  # must not take the decode measurement down
        return {"error": str(ex)[:200]}


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
    full_dir = model_dir_for(workload)
    max_ctx = max(1024, prefill + K + W + 64)

    ctx = B.Context(local_rank)
    tp = args.tp if args.tp > 1 else 1
    mdir = full_dir
    if tp > 1:
        if world != tp:
            raise SystemExit(f"--tp {tp} needs exactly {tp} ranks (torchrun --nproc-per-node {tp}); got WORLD_SIZE={world}")
        mdir = tp_setup(ctx, dist, rank, world, full_dir, bool(os.environ.get("UZU_TP_P2P")))
    eng = B.Engine(ctx, mdir, max_context_length=max_ctx, use_cuda_graph=not args.no_graph, fused_decode=not args.no_fused,
                   tp_rank=rank if tp > 1 else 0, tp_size=tp)
    if args.no_persistent and eng.persistent_decode:
        eng.set_persistent_decode(False)
    replicas = 1 if tp > 1 else world
    if args.batch > 1:
        run_batched(args, eng, ctx, dist, rank, world, local_rank, workload, prefill, K, W)
        return
    info = eng.info
    rng = np.random.default_rng(0)
    prompt = rng.integers(0, info.vocab_size, prefill).astype(np.uint32)
    log(f"[bench] engine ready: {workload}, persistent decode = {eng.persistent_decode} {eng.persistent_decode_reason}")
    m = measure_decode(eng, dist, world, replicas, local_rank, prompt, K, W)
    log(f"[bench] {workload}: {m['value']:.1f} tok/s device-resident, {m['e2e_value']:.1f} end to end, {m['launches']} launches")
    roof, extra = roofline_of(eng, workload, prefill, K, m["seconds"])
    ctx_mid = prefill + K / 2
    line = {
        "metric": metric_name(workload), "value": m["value"], "unit": "tokens/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": 1000.0 * m["seconds"] / K, "higher_is_better": True, "scaling": "strong" if tp > 1 else "weak", "vs_baseline": None,
        "dtype": dtype_name(workload), "data": "synthetic",
        "config": {
            "workload": workload_name(workload, prefill, K, 1), "parallelism": f"tp{tp}" if tp > 1 else ("replicas" if world > 1 else "single"),
            "per_gpu_value": m["value"] / world,
            "layers": info.num_layers, "attention_layers": info.num_attention_layers, "delta_net_layers": info.num_delta_net_layers,
            "model_dim": info.model_dim, "vocab": info.vocab_size, "weight_bytes_per_token": info.weight_bytes_per_token,
            "kv_bytes_per_token_at_mid_ctx": int(info.kv_bytes_per_token_per_ctx * ctx_mid), "state_bytes_per_token": info.state_bytes_per_token,
            "cache_policy": f"inputs larger than L2: {info.weight_bytes_per_token / 1e6:.0f} MB of weights streamed every step vs 126 MB L2",
            "decode_path": "persistent kernel (1 launch per token)" if eng.persistent_decode else
                           f"per-kernel path ({'CUDA graph' if not args.no_graph else 'eager'}): {eng.persistent_decode_reason or 'persistent kernel disabled'}",
            "cuda_graph": not args.no_graph, "fused_decode": not args.no_fused, "prefill_tokens_per_s": prefill / m["prefill_s"],
            "prefill_gemm": prefill_gemm_of(eng, prefill), **extra,
            "decode_paths": both_paths(eng) if tp == 1 else None,
        },
        "roofline": roof,
        "e2e": {"value": m["e2e_value"], "unit": "tokens/s", "h2d_bytes_per_step": 4, "d2h_bytes_per_step": 4},
        "gpu_launches": m["launches"],
        "clocks": m["clocks"],
    }
    threads = effective_cpus()
    if rank == 0 and tp == 1:
        try:
            line["config"]["speculation"] = speculation_of(eng, 1000.0 * m["seconds"] / K)
        except Exception as ex:  # a side measurement must not take the line down
            line["config"]["speculation"] = {"error": str(ex)[:300]}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"], line["parity"] = cpu_leg(eng, full_dir, threads, budget_s=45.0)
        except Exception as ex:  # the baseline must not take the measurement down
            line["cpu_baseline"] = {"value": None, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": f"failed: {ex}"}
    eng.close()
    # ---- secondary line (N = 1): BASELINE.json configs[1], Qwen3.5-0.8B int4 prefill 512 / decode 128 ----
    if world == 1 and tp == 1 and not args.no_secondary and workload != SECONDARY:
        try:
            line["config"]["secondary"] = secondary_line(B, ctx, args, local_rank)
        except Exception as ex:
            line["config"]["secondary"] = {"error": str(ex)[:300]}
    # ---- tensor-parallel sub-measurement (N >= 2): the same model sharded over the N ranks ----
    if world > 1 and tp == 1 and not args.no_tp_sub:
        sub = {}
        # the one-kernel peer-memory exchange has run on hardware at 2 ranks only: larger groups need UZU_BENCH_TP_P2P=1
        modes = ("nccl", "p2p") if (world == 2 or os.environ.get("UZU_BENCH_TP_P2P")) else ("nccl",)
        for mode in modes:
            try:
                sub[mode] = tp_line(B, dist, args, rank, world, local_rank, full_dir, workload, prefill, max_ctx, mode == "p2p")
            except Exception as ex:
                sub[mode] = {"error": str(ex)[:300]}
        if rank == 0:
            line["config"]["tp"] = sub
    ctx.close()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


SECONDARY = "qwen3.5-0.8b-int4"


def secondary_line(B, ctx, args, local_rank: int):
    _, _, prefill, K = WORKLOADS[SECONDARY]
    mdir = model_dir_for(SECONDARY)
    eng = B.Engine(ctx, mdir, max_context_length=max(1024, prefill + K + 96), use_cuda_graph=not args.no_graph, fused_decode=not args.no_fused)
    try:
        if args.no_persistent and eng.persistent_decode:
            eng.set_persistent_decode(False)
        rng = np.random.default_rng(0)
        prompt = rng.integers(0, eng.info.vocab_size, prefill).astype(np.uint32)
        m = measure_decode(eng, None, 1, 1, local_rank, prompt, K, max(args.warmup, 3))
        log(f"[bench] {SECONDARY}: {m['value']:.1f} tok/s device-resident, {m['e2e_value']:.1f} end to end, {m['launches']} launches")
        roof, extra = roofline_of(eng, SECONDARY, prefill, K, m["seconds"])
        out = {"workload": workload_name(SECONDARY, prefill, K, 1), "value": m["value"], "unit": "tokens/s", "ms_per_step": 1000.0 * m["seconds"] / K,
               "e2e": m["e2e_value"], "gpu_launches": m["launches"], "roofline": roof, "prefill_tokens_per_s": prefill / m["prefill_s"],
               "decode_path": "persistent kernel (1 launch per token)" if eng.persistent_decode else "per-kernel path", "decode_paths": both_paths(eng), **extra}
        if not args.no_cpu_baseline:
            try:
                out["cpu_baseline"], out["parity"] = cpu_leg(eng, mdir, effective_cpus(), budget_s=20.0)
            except Exception as ex:
                out["parity"] = {"error": str(ex)[:300]}
        return out
    finally:
        eng.close()


def tp_line(B, dist, args, rank: int, world: int, local_rank: int, full_dir: Path, workload: str, prefill: int, max_ctx: int, p2p: bool):
    """ONE model sharded by attention head / FFN column / vocabulary row over all ranks; two all-reduces per layer + one all-gather."""
    ctx = B.Context(local_rank)
    try:
        sdir = tp_setup(ctx, dist, rank, world, full_dir, p2p)
        eng = B.Engine(ctx, sdir, max_context_length=max_ctx, use_cuda_graph=not args.no_graph, fused_decode=not args.no_fused, tp_rank=rank, tp_size=world)
        try:
            K = min(args.steps or 64, 64)
            rng = np.random.default_rng(0)
            prompt = rng.integers(0, eng.info.vocab_size, prefill).astype(np.uint32)
            eng.prefill(prompt)
            eng.snapshot()
            eng.decode_timed(3)
            eng.restore()
            dist.barrier()
            seconds = max_over_ranks(dist, eng.decode_timed(K), "cuda")
            layers = eng.info.num_layers
            return {"tokens_per_s": K / seconds, "ms_per_step": 1000.0 * seconds / K, "steps": K, "exchange": "one-kernel peer-memory all-reduce (CUDA IPC over NVLink)" if p2p else "ncclAllReduce",
                    "collectives_per_step": 2 * layers + 1, "comm_nranks": world,
                    "limiting_collective": f"{2 * layers} latency-bound all-reduces of [1, model_dim] f32 per step (2 per layer) + 1 all-gather of the logits"}
        finally:
            eng.close()
    finally:
        ctx.close()


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
        "metric": metric_name(workload),
        "value": value, "unit": "tokens/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": 1000.0 * seconds / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": dtype_name(workload),
        "data": "synthetic",
        "config": {"workload": f"{workload}: prefill {prefill}, decode {K}, batch {nseq} (independent sequences, one weight pass per step), greedy",
                   "parallelism": "replicas" if world > 1 else "single", "batch": nseq, "per_gpu_value": value / world, "weight_bytes_per_token": info.weight_bytes_per_token,
                   "cache_policy": f"inputs larger than L2: {info.weight_bytes_per_token / 1e6:.0f} MB of weights streamed every step vs 126 MB L2",
                   "cuda_graph": not args.no_graph, "launches_per_step": int(launches) // max(1, K)},
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
    ap.add_argument("--workload", default="llama3-8b-int4", choices=sorted(WORKLOADS))
    ap.add_argument("--no-secondary", action="store_true", help="skip the Qwen3.5-0.8B int4 line (BASELINE configs[1]) that the default N = 1 run adds under config.secondary")
    ap.add_argument("--no-tp-sub", action="store_true", help="N >= 2: skip the tensor-parallel sub-measurement reported under config.tp")
    ap.add_argument("--no-persistent", action="store_true", help="use the per-kernel decode path instead of the persistent whole-token kernel (A/B)")
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
