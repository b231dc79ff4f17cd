"""Tensor-parallel sharding of a uzu-format checkpoint (nothing in the reference to mirror: uzu is single-device).

`shard_checkpoint(model_dir, out_dir, rank, size)` writes the checkpoint rank `rank` of `size` loads: an ordinary uzu checkpoint of a
*narrower* model (num_heads / num_groups / hidden_dim divided by `size`, model_dim unchanged) plus a `"tensor_parallel"` block in
config.json that tells the engine where the exchange steps are:

  * qkv_projection   rows: q heads [r*Hq/P, ..) | k heads [r*Hkv/P, ..) | v heads [r*Hkv/P, ..)   column-parallel, no exchange
  * gate_projection  rows: q heads [r*Hq/P, ..)                                                    (Qwen-style gated attention)
  * out_projection   K columns [r*Hq*D/P, ..)                                                      row-parallel  -> all-reduce [m, H]
  * up_projection    rows: value [r*F/P, ..) | gate [F + r*F/P, ..) of the fused [value | gate] matrix (gated_act_mul.rs:52-55)
  * down_projection  K columns [r*F/P, ..)                                                         row-parallel  -> all-reduce [m, H]
  * output embedding rows [r*V/P, ..) (vocab-parallel readout -> all-gather of [m, V/P] logits); the input embedding, every norm
    scale and the RoPE tables are replicated. A tied embedding is untied in the shard (full table for the lookup + its row slice).

Quantisation groups run along K, so K shards must be whole groups (and whole zero-point bytes for 4-bit): checked here.
The KV cache of a rank holds its Hkv/P heads only; attention is rank-local. DeltaNet layers are not sharded (the BASELINE config
that needs TP, Llama-3-70B, has none): such checkpoints are rejected.
"""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

from . import safetensors_io as st


class TpError(ValueError):
    pass


def _spec(meta, prefix):
    return json.loads(meta[prefix + ".spec"])


def _rows(tensors, dtypes, prefix, segments):
    """Row-slice every tensor of the weight matrix `prefix` (codes, scales, zero points, biases share the row axis)."""
    out = {}
    for suffix in ("weights", "scales", "zero_points", "biases"):
        name = f"{prefix}.{suffix}"
        if name in tensors:
            t = tensors[name]
            out[name] = np.concatenate([t[b:b + c] for b, c in segments], axis=0)
    return out


def _cols(tensors, meta, prefix, begin, count):
    """K-slice [begin, begin+count) (in weight elements) of the weight matrix `prefix`."""
    spec = _spec(meta, prefix)
    out = {}
    w = tensors[f"{prefix}.weights"]
    if spec["type"] == "FullPrecisionSpec":
        out[f"{prefix}.weights"] = w[:, begin:begin + count]
        return out
    bits, gs = spec["bits"], spec["group_size"]
    if begin % gs or count % gs:
        raise TpError(f"{prefix}: K shard [{begin}, {begin + count}) is not a whole number of quantisation groups of {gs}")
    per_byte = 8 // bits
    out[f"{prefix}.weights"] = w[:, begin // per_byte:(begin + count) // per_byte]
    g0, gc = begin // gs, count // gs
    out[f"{prefix}.scales"] = tensors[f"{prefix}.scales"][:, g0:g0 + gc]
    if f"{prefix}.biases" in tensors:
        out[f"{prefix}.biases"] = tensors[f"{prefix}.biases"][:, g0:g0 + gc]
    if f"{prefix}.zero_points" in tensors:
        zp = tensors[f"{prefix}.zero_points"]
        if bits == 4:
            if g0 % 2 or gc % 2:
                raise TpError(f"{prefix}: 4-bit zero points are packed two groups per byte; shard needs an even group offset / count")
            out[f"{prefix}.zero_points"] = zp[:, g0 // 2:(g0 + gc) // 2]
        else:
            out[f"{prefix}.zero_points"] = zp[:, g0:g0 + gc]
    return out


def shard_tensors(config: dict, tensors: dict, dtypes: dict, meta: dict, rank: int, size: int):
    """Returns (config_r, tensors_r, meta_r) for rank `rank` of `size`."""
    if size < 1 or not 0 <= rank < size:
        raise TpError(f"bad rank {rank} of {size}")
    cfg = json.loads(json.dumps(config))
    dec = cfg["decoder_config"]
    tr = dec["transformer_config"]
    H, F, V = tr["model_dim"], tr["hidden_dim"], dec["vocab_size"]
    if F % size or V % size:
        raise TpError(f"hidden_dim {F} and vocab_size {V} must be divisible by the tensor-parallel size {size}")
    T, M = {}, {}

    def keep(name):
        T[name] = tensors[name]

    def put(d, prefix):
        T.update(d)
        M[prefix + ".spec"] = meta[prefix + ".spec"]

    # embeddings: full table for the lookup, vocab slice for the readout
    Vl = V // size
    emb = dec["embedding_config"]
    tied = emb["type"] == "TiedEmbeddingConfig"
    src_in = "decoder.embedding.embedding" if tied else "decoder.embedding.input_embedding"
    src_out = "decoder.embedding.embedding" if tied else "decoder.embedding.output_embedding"
    for suffix in ("weights", "scales", "zero_points", "biases"):
        if f"{src_in}.{suffix}" in tensors:
            T[f"decoder.embedding.input_embedding.{suffix}"] = tensors[f"{src_in}.{suffix}"]
    M["decoder.embedding.input_embedding.spec"] = meta[src_in + ".spec"]
    sl = _rows(tensors, dtypes, src_out, [(rank * Vl, Vl)])
    for k, v in sl.items():
        T["decoder.embedding.output_embedding." + k.rsplit(".", 1)[1]] = v
    M["decoder.embedding.output_embedding.spec"] = meta[src_out + ".spec"]
    emb["type"] = "UntiedEmbeddingConfig"

    for i, lc in enumerate(tr["layer_configs"]):
        p = f"decoder.transformer.layers.{i}"
        keep(f"{p}.pre_mixer_norm.scales")
        keep(f"{p}.pre_mlp_norm.scales")
        mc = lc["mixer_config"]
        if mc["type"] != "AttentionConfig":
            raise TpError(f"layer {i}: tensor parallelism covers attention mixers only ({mc['type']} is not sharded)")
        Hq, Hkv, D = mc["num_heads"], mc["num_groups"], mc["head_dim"]
        if Hq % size or Hkv % size:
            raise TpError(f"layer {i}: {Hq} query / {Hkv} kv heads are not divisible by the tensor-parallel size {size}")
        hq, hkv = Hq // size, Hkv // size
        qd, kvd = Hq * D, Hkv * D
        put(_rows(tensors, dtypes, f"{p}.mixer.qkv_projection.weights",
                  [(rank * hq * D, hq * D), (qd + rank * hkv * D, hkv * D), (qd + kvd + rank * hkv * D, hkv * D)]),
            f"{p}.mixer.qkv_projection.weights")
        if mc.get("gate_projection_config") is not None:
            put(_rows(tensors, dtypes, f"{p}.mixer.gate_projection.weights", [(rank * hq * D, hq * D)]), f"{p}.mixer.gate_projection.weights")
        put(_cols(tensors, meta, f"{p}.mixer.out_projection.weights", rank * hq * D, hq * D), f"{p}.mixer.out_projection.weights")
        for nm in ("query_norm", "key_norm"):
            if f"{p}.mixer.{nm}.scales" in tensors:
                keep(f"{p}.mixer.{nm}.scales")
        mc["num_heads"], mc["num_groups"] = hq, hkv
        Fi = lc["hidden_dim"] if lc.get("hidden_dim") is not None else F
        if Fi % size:
            raise TpError(f"layer {i}: hidden_dim {Fi} is not divisible by {size}")
        fl = Fi // size
        put(_rows(tensors, dtypes, f"{p}.mlp.up_projection.weights", [(rank * fl, fl), (Fi + rank * fl, fl)]), f"{p}.mlp.up_projection.weights")
        put(_cols(tensors, meta, f"{p}.mlp.down_projection.weights", rank * fl, fl), f"{p}.mlp.down_projection.weights")
        if lc.get("hidden_dim") is not None:
            lc["hidden_dim"] = fl
    tr["hidden_dim"] = F // size
    keep("decoder.transformer.output_norm.scales")
    cfg["tensor_parallel"] = {"rank": rank, "size": size, "vocab_size_local": Vl, "vocab_offset": rank * Vl}
    # restore the BF16 marker on sliced arrays
    for name, t in list(T.items()):
        src = name
        if name.startswith("decoder.embedding.input_embedding.") or name.startswith("decoder.embedding.output_embedding."):
            src = (src_in if "input_embedding" in name else src_out) + "." + name.rsplit(".", 1)[1]
        if dtypes.get(src) == "BF16":
            T[name] = st.as_bf16(np.ascontiguousarray(t).view(np.uint16))
        else:
            T[name] = np.ascontiguousarray(t)
    return cfg, T, M


def check_shardable(config: dict, meta: dict, size: int) -> list:
    """Shape-only version of the checks shard_tensors() makes (no tensor data needed: usable on a 70B config before 35 GB of
    weights exist). Returns the list of problems; empty = every rank of `size` gets whole heads, whole quantisation groups and
    whole zero-point bytes."""
    problems = []
    dec = config["decoder_config"]
    tr = dec["transformer_config"]
    F, V = tr["hidden_dim"], dec["vocab_size"]
    if V % size:
        problems.append(f"vocab_size {V} is not divisible by {size}")

    def k_shard(prefix, k_total):
        spec = json.loads(meta[prefix + ".spec"])
        if spec["type"] == "FullPrecisionSpec":
            return
        gs, bits = spec["group_size"], spec["bits"]
        per = k_total // size
        if per % gs:
            problems.append(f"{prefix}: K shard of {per} is not a whole number of groups of {gs}")
        elif bits == 4 and spec["type"] == "IntSpec" and not spec.get("is_symmetric", False) and (per // gs) % 2:
            problems.append(f"{prefix}: {per // gs} groups per shard = half a 4-bit zero-point byte")

    for i, lc in enumerate(tr["layer_configs"]):
        p = f"decoder.transformer.layers.{i}"
        mc = lc["mixer_config"]
        if mc["type"] != "AttentionConfig":
            problems.append(f"layer {i}: {mc['type']} is not sharded")
            continue
        Hq, Hkv, D = mc["num_heads"], mc["num_groups"], mc["head_dim"]
        if Hq % size or Hkv % size:
            problems.append(f"layer {i}: {Hq} query / {Hkv} kv heads are not divisible by {size}")
        else:
            k_shard(f"{p}.mixer.out_projection.weights", Hq * D)
        Fi = lc["hidden_dim"] if lc.get("hidden_dim") is not None else F
        if Fi % size:
            problems.append(f"layer {i}: hidden_dim {Fi} is not divisible by {size}")
        else:
            k_shard(f"{p}.mlp.down_projection.weights", Fi)
    return problems


def shard_checkpoint(model_dir, out_dir, rank: int, size: int) -> Path:
    model_dir, out_dir = Path(model_dir), Path(out_dir)
    config = json.loads((model_dir / "config.json").read_text())
    tensors, dtypes, meta = st.load(model_dir / "model.safetensors")
    cfg, T, M = shard_tensors(config, tensors, dtypes, meta, rank, size)
    out_dir.mkdir(parents=True, exist_ok=True)
    (out_dir / "config.json").write_text(json.dumps(cfg, indent=1))
    st.save(out_dir / "model.safetensors", T, M)
    return out_dir
