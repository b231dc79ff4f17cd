"""Tensor-parallel sharding (uzu_b200/tp.py) on CPU: the P-rank oracle (shards + all-reduce / all-gather hooks) against the
1-rank oracle on the same synthetic checkpoint. The reference has no tensor parallelism, so the unsharded model IS the oracle; the f32
partial sums are added in a different order than the sequential k loop, hence bf16-ulp (not bit) agreement. One test runs the ranks as
threads in lockstep, one as two real processes over torch.distributed gloo (the host-side path of `bench.py --tp 2`)."""
import json
import os
import subprocess
import sys
import textwrap
import threading
from pathlib import Path

import numpy as np
import pytest

from oracle.model import OracleModel
from tests.util import assert_bf16_close, bf16_to_f32
from uzu_b200 import safetensors_io as st
from uzu_b200 import synth, tp

ROOT = Path(__file__).resolve().parents[1]


def _write_shards(tmp_path, spec, size, seed=3):
    full = synth.write_model(spec, tmp_path / "full", seed=seed)
    return full, [tp.shard_checkpoint(full, tmp_path / f"rank{r}", r, size) for r in range(size)]


def test_shard_shapes_follow_the_partitioning(tmp_path):
    spec = synth.tiny("llama")          # H 256, F 512, V 1000, 4 q heads / 2 kv heads x 64
    full, shards = _write_shards(tmp_path, spec, 2)
    T, _, M = st.load(shards[1] / "model.safetensors")
    cfg = json.loads((shards[1] / "config.json").read_text())
    assert cfg["tensor_parallel"] == {"rank": 1, "size": 2, "vocab_size_local": 500, "vocab_offset": 500}
    mc = cfg["decoder_config"]["transformer_config"]["layer_configs"][0]["mixer_config"]
    assert (mc["num_heads"], mc["num_groups"]) == (2, 1) and cfg["decoder_config"]["transformer_config"]["hidden_dim"] == 256
    p = "decoder.transformer.layers.0."
    assert T[p + "mixer.qkv_projection.weights.weights"].shape == ((2 + 1 + 1) * 64, 256 // 2)
    assert T[p + "mixer.out_projection.weights.weights"].shape == (256, 128 // 2)
    assert T[p + "mixer.out_projection.weights.scales"].shape == (256, 2) and T[p + "mixer.out_projection.weights.zero_points"].shape == (256, 1)
    assert T[p + "mlp.up_projection.weights.weights"].shape == (512, 128) and T[p + "mlp.down_projection.weights.weights"].shape == (256, 128)
    assert T["decoder.embedding.input_embedding.weights"].shape[0] == 1000 and T["decoder.embedding.output_embedding.weights"].shape[0] == 500
    # rank 1's slices are the second halves of the full tensors
    F, _, _ = st.load(full / "model.safetensors")
    assert (T[p + "mlp.down_projection.weights.weights"] == F[p + "mlp.down_projection.weights.weights"][:, 128:]).all()
    up = F[p + "mlp.up_projection.weights.weights"]
    assert (T[p + "mlp.up_projection.weights.weights"] == np.concatenate([up[256:512], up[512 + 256:]])).all()
    qkv = F[p + "mixer.qkv_projection.weights.weights"]
    assert (T[p + "mixer.qkv_projection.weights.weights"] == np.concatenate([qkv[128:256], qkv[256 + 64:256 + 128], qkv[384 + 64:384 + 128]])).all()


def test_unshardable_checkpoints_are_rejected(tmp_path):
    hybrid = synth.write_model(synth.tiny("qwen-hybrid"), tmp_path / "h", seed=1)
    with pytest.raises(tp.TpError, match="attention mixers only"):
        tp.shard_checkpoint(hybrid, tmp_path / "h0", 0, 2)
    llama = synth.write_model(synth.tiny("llama"), tmp_path / "l", seed=1)
    with pytest.raises(tp.TpError, match="not divisible"):
        tp.shard_checkpoint(llama, tmp_path / "l0", 0, 4)          # 2 kv heads cannot feed 4 ranks
    g128 = synth.write_model(synth.tiny("llama", quant=synth.QuantSpec("int", 4, 128, False)), tmp_path / "g", seed=1)
    with pytest.raises(tp.TpError, match="two groups per byte"):
        tp.shard_checkpoint(g128, tmp_path / "g0", 0, 2)           # out-projection K shard = 1 group of 128: half a zero-point byte


class _LockstepExchange:
    """all-reduce / all-gather among rank threads of one process (sum in rank order, f32)."""

    def __init__(self, size):
        self.size, self.slots, self.barrier = size, [None] * size, threading.Barrier(size)

    def hooks(self, rank):
        def exchange(x, combine):
            self.slots[rank] = x
            self.barrier.wait()
            out = combine([self.slots[r] for r in range(self.size)])
            self.barrier.wait()
            return out

        def reduce(x):
            def comb(parts):
                acc = parts[0].astype(np.float32).copy()
                for q in parts[1:]:
                    acc += q
                return acc
            return exchange(x, comb)

        return reduce, lambda x: exchange(x, lambda parts: np.concatenate(parts, axis=1))


@pytest.mark.parametrize("kind,quant", [("llama", None), ("qwen-dense", None), ("llama", synth.QuantSpec("mlx", 4, 32)),
                                         ("llama", synth.QuantSpec("int", 8, 64, False))])
def test_two_rank_oracle_matches_unsharded(tmp_path, kind, quant):
    spec = synth.tiny(kind, quant=quant)
    full, shards = _write_shards(tmp_path, spec, 2)
    rng = np.random.default_rng(4)
    prompt = rng.integers(0, spec.vocab_size, 9)
    ref = OracleModel(full, max_context=64)
    want = [ref.forward(prompt)] + [ref.forward([t]) for t in (5, 17)]
    ex = _LockstepExchange(2)
    got = [None, None]

    def run(rank):
        red, gat = ex.hooks(rank)
        m = OracleModel(shards[rank], max_context=64, tp_reduce=red, tp_gather=gat)
        got[rank] = [m.forward(prompt)] + [m.forward([t]) for t in (5, 17)]

    threads = [threading.Thread(target=run, args=(r,)) for r in range(2)]
    [t.start() for t in threads]
    [t.join(timeout=120) for t in threads]
    assert got[0] is not None and got[1] is not None
    for a, b, w in zip(got[0], got[1], want):
        assert (a == b).all(), "ranks must hold identical logits after the all-gather"
        scale = float(np.abs(bf16_to_f32(w)).max())
        assert float(np.abs(bf16_to_f32(a) - bf16_to_f32(w)).max()) <= 0.02 * scale + 1e-3
        assert_bf16_close(a, w, max_ulp=4, min_exact=0.5, atol=2e-2 * scale, what=f"tp2 {kind}")
        assert int(np.argmax(bf16_to_f32(a[0]))) == int(np.argmax(bf16_to_f32(w[0])))


def test_eight_rank_oracle_with_llama3_head_geometry(tmp_path):
    """BASELINE config 5's split (TP = 8) on a small model with Llama-3's head geometry ratios: 8 kv heads -> one per rank, q heads 4 per
    kv head, FFN columns and vocabulary in 8 slices, every K shard an even number of quantisation groups. The 8 rank oracles (threads,
    lockstep exchange) must reproduce the unsharded logits."""
    from uzu_b200.synth import LLAMA3_ROPE, ModelSpec
    spec = ModelSpec(name="tiny-llama-8kv", model_dim=256, hidden_dim=1024, vocab_size=1024, layer_kinds=["attn"] * 2, num_heads=32, num_groups=8,
                     head_dim=32, rope=dict(LLAMA3_ROPE, head_dim=32), quant=synth.QuantSpec("int", 4, 64, False))
    full, shards = _write_shards(tmp_path, spec, 8)
    cfg = json.loads((shards[5] / "config.json").read_text())
    mc = cfg["decoder_config"]["transformer_config"]["layer_configs"][0]["mixer_config"]
    assert (mc["num_heads"], mc["num_groups"]) == (4, 1) and cfg["tensor_parallel"]["vocab_size_local"] == 128
    rng = np.random.default_rng(6)
    prompt = rng.integers(0, spec.vocab_size, 7)
    ref = OracleModel(full, max_context=64)
    want = [ref.forward(prompt), ref.forward([11])]
    ex = _LockstepExchange(8)
    got = [None] * 8

    def run(rank):
        red, gat = ex.hooks(rank)
        m = OracleModel(shards[rank], max_context=64, tp_reduce=red, tp_gather=gat)
        got[rank] = [m.forward(prompt), m.forward([11])]

    threads = [threading.Thread(target=run, args=(r,)) for r in range(8)]
    [t.start() for t in threads]
    [t.join(timeout=180) for t in threads]
    assert all(g is not None for g in got)
    for step, w in enumerate(want):
        for r in range(1, 8):
            assert (got[r][step] == got[0][step]).all()
        scale = float(np.abs(bf16_to_f32(w)).max())
        assert float(np.abs(bf16_to_f32(got[0][step]) - bf16_to_f32(w)).max()) <= 0.02 * scale + 1e-3


def test_two_rank_gloo_processes(tmp_path):
    spec = synth.tiny("llama")
    full, shards = _write_shards(tmp_path, spec, 2)
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import sys, json
        sys.path.insert(0, {str(ROOT)!r})
        import numpy as np, torch, torch.distributed as dist
        from oracle.model import OracleModel
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        def reduce(x):
            t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)); dist.all_reduce(t); return t.numpy()
        def gather(x):
            t = torch.from_numpy(np.ascontiguousarray(x).astype(np.int32)); outs = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(outs, t); return np.concatenate([o.numpy().astype(np.uint16) for o in outs], axis=1)
        m = OracleModel({str(tmp_path)!r} + f"/rank{{rank}}", max_context=64, tp_reduce=reduce, tp_gather=gather)
        lg = m.forward(np.arange(7) * 13 % {spec.vocab_size})
        lg2 = m.forward([3])
        if rank == 0:
            ref = OracleModel({str(full)!r}, max_context=64)
            r1 = ref.forward(np.arange(7) * 13 % {spec.vocab_size}); r2 = ref.forward([3])
            f = lambda h: (h.astype(np.uint32) << 16).view(np.float32)
            err = max(float(np.abs(f(lg) - f(r1)).max()), float(np.abs(f(lg2) - f(r2)).max()))
            print(json.dumps({{"err": err, "scale": float(np.abs(f(r1)).max()), "shape": list(lg.shape)}}))
        dist.barrier(); dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29617")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29617", str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert d["shape"] == [1, spec.vocab_size] and d["err"] <= 0.02 * d["scale"] + 1e-3, d


def test_nccl_is_resolved_at_run_time():
    """csrc/tp.cu dlopens NCCL on first use (no link-time dependency): the unique id call works without a GPU."""
    from uzu_b200 import binding as B
    try:
        a, b = B.tp_unique_id(), B.tp_unique_id()
    except B.UzuError as ex:
        if "cannot load NCCL" in str(ex):
            pytest.skip("no libnccl.so.2 on this host")
        raise
    assert len(a) == 128 and a != b


@pytest.mark.parametrize("preset,kw,sizes", [("llama3-8b", dict(bits=4), (2, 4, 8)), ("llama3-8b", dict(bits=8), (2, 4, 8)),
                                              ("llama3-70b", dict(bits=4), (2, 4, 8))])
def test_baseline_models_are_shardable_shape_only(preset, kw, sizes):
    """Llama-3-8B / 70B (BASELINE configs 3-5) at TP 2 / 4 / 8: whole heads, whole groups, whole zero-point bytes on every rank --
    checked from config + quantisation specs alone (no 35 GB of weights needed)."""
    spec = synth.PRESETS[preset](layers=2, **kw)
    cfg = synth.build_config(spec)
    q = spec.quant.spec_json("output_input")
    meta = {}
    for i in range(spec.num_layers):
        for name in ("mixer.qkv_projection", "mixer.out_projection", "mlp.up_projection", "mlp.down_projection"):
            meta[f"decoder.transformer.layers.{i}.{name}.weights.spec"] = json.dumps(q)
    for size in sizes:
        assert tp.check_shardable(cfg, meta, size) == [], (preset, size)
    assert tp.check_shardable(cfg, meta, 16)            # 8 kv heads cannot feed 16 ranks
    hybrid = synth.build_config(synth.qwen35_0p8b(layers=3))      # 3 DeltaNet layers
    assert any("DeltaNetConfig" in p for p in tp.check_shardable(hybrid, {}, 2))


def test_bench_shard_dir_is_written_once(tmp_path, monkeypatch):
    import bench
    full = synth.write_model(synth.tiny("llama"), tmp_path / "tiny-llama-seed0", seed=0)
    d0 = bench.shard_dir_for(full, 0, 2)
    assert (d0 / ".done").exists() and json.loads((d0 / "config.json").read_text())["tensor_parallel"]["rank"] == 0
    stamp = (d0 / "model.safetensors").stat().st_mtime_ns
    assert bench.shard_dir_for(full, 0, 2) == d0 and (d0 / "model.safetensors").stat().st_mtime_ns == stamp
    assert bench.shard_dir_for(full, 1, 2).name.endswith("tp2-rank1")
