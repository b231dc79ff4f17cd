"""Synthetic uzu-format checkpoints (config.json + model.safetensors) for tests and benchmarks.

There is no network in the build/bench environment, so real checkpoints cannot be fetched; this
module emits seeded random weights in exactly the on-disk layout the reference engine loads
(SURVEY.md Appendix A): config.json follows crates/backend-uzu/src/config/** (abstract-config
variants carry a "type" tag, every Option field is present as null), tensor keys follow
encodable_block/{embedding,transformer_layer,mixer/attention/mod,mixer/delta_net,mlp/mod}.rs and
per-matrix quantisation specs are JSON strings under `<prefix>.spec` in the safetensors
`__metadata__` (parameters/loader.rs:219-228).
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field, asdict
from pathlib import Path

import numpy as np

from . import safetensors_io as st


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even f32 -> bf16 bits (same as `half::bf16::from_f32`)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32)
    up = ((u & np.uint32(0x8000)) != 0) & ((u & np.uint32(0x17FFF)) != 0)
    return ((u >> 16) + up.astype(np.uint32)).astype(np.uint16)


@dataclass
class QuantSpec:
    kind: str = "int"          # "int" (IntSpec), "mlx" (MLXSpec), "fp" (FullPrecisionSpec)
    bits: int = 4
    group_size: int = 64
    symmetric: bool = False    # IntSpec.is_symmetric
    rht: bool = False          # HybridSpec: 32-wide input/output randomized Hadamard around this quantization (Mirai RHT)

    def spec_json(self, layout: str) -> dict:
        if self.rht:
            inner = QuantSpec(self.kind, self.bits, self.group_size, self.symmetric).spec_json(layout)
            return {"type": "HybridSpec", "quantization_spec": inner, "adapter_spec": None, "incoherence_block_size": 32,
                    "incoherence_processing_mode": "input_output"}
        if self.kind == "fp":
            return {"type": "FullPrecisionSpec", "layout": layout}
        if self.kind == "mlx":
            return {"type": "MLXSpec", "bits": self.bits, "group_size": self.group_size, "layout": layout}
        return {"type": "IntSpec", "bits": self.bits, "group_size": self.group_size,
                "is_symmetric": self.symmetric, "layout": layout}


@dataclass
class ModelSpec:
    name: str
    model_dim: int
    hidden_dim: int
    vocab_size: int
    # per layer: "attn" or "delta"
    layer_kinds: list = field(default_factory=list)
    num_heads: int = 8
    num_groups: int = 2
    head_dim: int = 128
    rope: dict = field(default_factory=dict)      # AnyRoPEConfig json (with "type")
    qk_norm: bool = False
    qk_norm_scale_offset: float | None = None
    attn_gate: bool = False
    norm_eps: float = 1e-5
    norm_scale_offset: float | None = None
    upcast_mode: str = "only_normalization"
    tied_embeddings: bool = False
    quant: QuantSpec = field(default_factory=QuantSpec)
    embedding_quant: QuantSpec | None = None      # default: same as `quant`
    # DeltaNet
    dn_num_heads: int = 16
    dn_num_groups: int = 16
    dn_head_dim: int = 128
    dn_value_head_dim: int = 128
    dn_kernel_size: int = 4
    max_sequence_length: int = 8192

    @property
    def num_layers(self) -> int:
        return len(self.layer_kinds)


def _norm_cfg(spec: ModelSpec, scale_offset=None, eps=None):
    return {"epsilon": eps if eps is not None else spec.norm_eps,
            "scale_offset": scale_offset if scale_offset is not None else spec.norm_scale_offset,
            "upcast_mode": spec.upcast_mode, "subtract_mean": False, "has_scale": True, "has_biases": False}


def build_config(spec: ModelSpec) -> dict:
    layers = []
    for kind in spec.layer_kinds:
        if kind == "attn":
            mixer = {
                "type": "AttentionConfig", "qkv_projection_config": {}, "out_projection_config": {},
                "query_norm_config": _norm_cfg(spec, spec.qk_norm_scale_offset, 1e-6) if spec.qk_norm else None,
                "key_norm_config": _norm_cfg(spec, spec.qk_norm_scale_offset, 1e-6) if spec.qk_norm else None,
                "num_heads": spec.num_heads, "num_groups": spec.num_groups, "head_dim": spec.head_dim,
                "is_causal": True, "scale": None, "sliding_window_size": None, "logit_soft_cap": None,
                "has_sinks": False, "has_qkv_biases": False, "has_out_biases": False,
                "gate_projection_config": {} if spec.attn_gate else None, "normalize_values": False,
                "is_kv_sharing": False,
            }
            rope = dict(spec.rope)
        else:
            mixer = {
                "type": "DeltaNetConfig", "in_proj_config": {}, "conv_config": {"has_biases": False},
                "out_proj_config": {}, "norm_config": _norm_cfg(spec, None, 1e-6),
                "num_heads": spec.dn_num_heads, "num_groups": spec.dn_num_groups, "head_dim": spec.dn_head_dim,
                "value_head_dim": spec.dn_value_head_dim, "kernel_size": spec.dn_kernel_size,
            }
            rope = None
        layers.append({
            "pre_mixer_norm_config": _norm_cfg(spec), "mixer_config": mixer, "post_mixer_norm_config": None,
            "pre_mlp_norm_config": _norm_cfg(spec),
            "mlp_config": {"type": "DenseMLPConfig", "linear_config": {}, "activation": {"type": "SiLU", "alpha": 1.0},
                           "has_up_biases": False, "has_down_biases": False, "gate_clipping": None,
                           "up_clipping": None},
            "post_mlp_norm_config": None, "hidden_dim": None, "ple_config": None, "has_post_layer_scalar": False,
            "kv_source_layer_index": None, "rope_config": rope,
        })
    emb_type = "TiedEmbeddingConfig" if spec.tied_embeddings else "UntiedEmbeddingConfig"
    return {
        "type": "LanguageModelConfig",
        "token_codec_config": {"type": "RawTextCodecConfig"},
        "decoder_config": {
            "embedding_config": {"type": emb_type, "input_scale": None, "logit_soft_cap": None, "logit_scale": None},
            "transformer_config": {"layer_configs": layers, "output_norm_config": _norm_cfg(spec),
                                   "model_dim": spec.model_dim, "hidden_dim": spec.hidden_dim},
            "vocab_size": spec.vocab_size, "ple_model_config": None, "embedding_norm_config": None,
        },
        "generation_config": {"stop_token_ids": [], "temperature": None, "top_k": None, "top_p": None, "min_p": None,
                              "banned_tokens": None, "repetition_penalty": None, "presence_penalty": None,
                              "frequency_penalty": None, "suffix_repetition_length": None},
    }


# ---------------------------------------------------------------------------------------------
# weights
# ---------------------------------------------------------------------------------------------
def quantized_matrix(rng: np.random.Generator, rows: int, cols: int, q: QuantSpec, target_std: float):
    """Random block-quantised [rows, cols] matrix in uzu layout.

    codes uniform in [0, 2^bits); packed little-endian nibbles / bytes row-major
    (cpu/kernel/matmul/kernel.rs:236-243); scales bf16 [rows, cols/gs]; zero points nibble-packed
    [rows, ceil(groups/2)] (4-bit) or [rows, groups] (8-bit) (weight_matrix.rs:136-151).
    """
    out = {}
    if q.kind == "fp":
        w = rng.standard_normal((rows, cols), dtype=np.float32) * target_std
        out["weights"] = st.as_bf16(f32_to_bf16_bits(w))
        return out
    levels = 1 << q.bits
    groups = -(-cols // q.group_size)
    if q.bits == 4:
        packed = rng.integers(0, 256, size=(rows, cols // 2), dtype=np.uint8)
    else:
        packed = rng.integers(0, 256, size=(rows, cols), dtype=np.uint8)
    out["weights"] = packed
    # (code - zp) has std ~ levels/sqrt(6); choose scales so dequantised weights have ~target_std
    base = target_std / (levels / np.sqrt(6.0))
    scales = (rng.random((rows, groups), dtype=np.float32) + 0.5) * base
    out["scales"] = st.as_bf16(f32_to_bf16_bits(scales))
    if q.kind == "mlx":
        sc = st_bits_to_f32(out["scales"])
        zp = rng.integers(0, levels, size=(rows, groups)).astype(np.float32)
        out["biases"] = st.as_bf16(f32_to_bf16_bits(-sc * zp))
    elif not q.symmetric:
        if q.bits == 4:
            out["zero_points"] = rng.integers(0, 256, size=(rows, -(-groups // 2)), dtype=np.uint8)
        else:
            out["zero_points"] = rng.integers(0, 256, size=(rows, groups), dtype=np.uint8)
    return out


def st_bits_to_f32(bits: np.ndarray) -> np.ndarray:
    return (np.asarray(bits, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)


def _add_matrix(tensors, meta, prefix, rng, rows, cols, q, layout, target_std):
    if q.rht:
        # RHTLinearWrapper::load_inner_with_output_rht (linear/rht_wrapper.rs:141-176): i32 +-1 signs + the inner matrix under `quantized`
        assert layout == "output_input" and q.kind != "fp" and rows % 32 == 0 and cols % 32 == 0
        inner = QuantSpec(q.kind, q.bits, q.group_size, q.symmetric)
        tensors[f"{prefix}.incoherence_signs.input_signs"] = rng.choice(np.array([-1, 1], np.int32), cols)
        tensors[f"{prefix}.incoherence_signs.output_signs"] = rng.choice(np.array([-1, 1], np.int32), rows)
        _add_matrix(tensors, meta, prefix + ".quantized", rng, rows, cols, inner, layout, target_std)
        meta[f"{prefix}.spec"] = json.dumps(q.spec_json(layout))
        return
    for k, v in quantized_matrix(rng, rows, cols, q, target_std).items():
        tensors[f"{prefix}.{k}"] = v
    meta[f"{prefix}.spec"] = json.dumps(q.spec_json(layout))


def build_weights(spec: ModelSpec, seed: int = 0):
    rng = np.random.default_rng(seed)
    T, M = {}, {}
    H, F, V = spec.model_dim, spec.hidden_dim, spec.vocab_size
    eq = spec.embedding_quant or spec.quant
    if eq.rht:     # quantized embeddings with a Hadamard are unimplemented on the reference CPU backend (quant_embedding.rs:32-34)
        eq = QuantSpec(eq.kind, eq.bits, eq.group_size, eq.symmetric)
    if spec.tied_embeddings:
        _add_matrix(T, M, "decoder.embedding.embedding", rng, V, H, eq, "input_output", 1.0 / np.sqrt(H) * 4.0)
    else:
        _add_matrix(T, M, "decoder.embedding.input_embedding", rng, V, H, eq, "input_output", 1.0)
        _add_matrix(T, M, "decoder.embedding.output_embedding", rng, V, H, eq, "input_output", 1.0 / np.sqrt(H) * 4.0)

    def norm_scales(n):
        base = 0.0 if spec.norm_scale_offset else 1.0
        return (base + rng.standard_normal(n, dtype=np.float32) * 0.02).astype(np.float32)

    for i, kind in enumerate(spec.layer_kinds):
        p = f"decoder.transformer.layers.{i}"
        T[f"{p}.pre_mixer_norm.scales"] = norm_scales(H)
        T[f"{p}.pre_mlp_norm.scales"] = norm_scales(H)
        if kind == "attn":
            qd = spec.num_heads * spec.head_dim
            kvd = spec.num_groups * spec.head_dim
            _add_matrix(T, M, f"{p}.mixer.qkv_projection.weights", rng, qd + 2 * kvd, H, spec.quant, "output_input",
                        1.0 / np.sqrt(H))
            if spec.qk_norm:
                off = 0.0 if spec.qk_norm_scale_offset else 1.0
                T[f"{p}.mixer.query_norm.scales"] = (off + rng.standard_normal(spec.head_dim, dtype=np.float32) * 0.05).astype(np.float32)
                T[f"{p}.mixer.key_norm.scales"] = (off + rng.standard_normal(spec.head_dim, dtype=np.float32) * 0.05).astype(np.float32)
            if spec.attn_gate:
                _add_matrix(T, M, f"{p}.mixer.gate_projection.weights", rng, qd, H, spec.quant, "output_input",
                            1.0 / np.sqrt(H))
            _add_matrix(T, M, f"{p}.mixer.out_projection.weights", rng, H, qd, spec.quant, "output_input",
                        1.0 / np.sqrt(qd))
        else:
            key_dim = spec.dn_num_groups * spec.dn_head_dim
            value_dim = spec.dn_num_heads * spec.dn_value_head_dim
            conv_dim = 2 * key_dim + value_dim
            total = conv_dim + value_dim + 2 * spec.dn_num_heads
            _add_matrix(T, M, f"{p}.mixer.in_proj.weights", rng, total, H, spec.quant, "output_input", 1.0 / np.sqrt(H))
            T[f"{p}.mixer.conv.weights"] = (rng.standard_normal((conv_dim, spec.dn_kernel_size), dtype=np.float32) * 0.5).astype(np.float32)
            T[f"{p}.mixer.a_log"] = rng.uniform(-1.0, 1.0, spec.dn_num_heads).astype(np.float32)
            T[f"{p}.mixer.dt_bias"] = rng.uniform(-1.0, 1.0, spec.dn_num_heads).astype(np.float32)
            T[f"{p}.mixer.norm.scales"] = (1.0 + rng.standard_normal(spec.dn_value_head_dim, dtype=np.float32) * 0.05).astype(np.float32)
            _add_matrix(T, M, f"{p}.mixer.out_proj.weights", rng, H, value_dim, spec.quant, "output_input",
                        1.0 / np.sqrt(value_dim))
        _add_matrix(T, M, f"{p}.mlp.up_projection.weights", rng, 2 * F, H, spec.quant, "output_input", 1.0 / np.sqrt(H))
        _add_matrix(T, M, f"{p}.mlp.down_projection.weights", rng, H, F, spec.quant, "output_input", 1.0 / np.sqrt(F))
    T["decoder.transformer.output_norm.scales"] = norm_scales(H)
    return T, M


def write_model(spec: ModelSpec, path, seed: int = 0) -> Path:
    path = Path(path)
    path.mkdir(parents=True, exist_ok=True)
    (path / "config.json").write_text(json.dumps(build_config(spec), indent=1))
    tensors, meta = build_weights(spec, seed)
    st.save(path / "model.safetensors", tensors, meta)
    (path / "synth_spec.json").write_text(json.dumps(asdict(spec), indent=1))
    return path


# ---------------------------------------------------------------------------------------------
# presets (BASELINE.json configs)
# ---------------------------------------------------------------------------------------------
LLAMA3_ROPE = {"type": "LlamaRoPEConfig", "base": 500000.0, "max_sequence_length": 8192, "head_dim": 128,
               "scaling_factor": 8.0, "original_context_length": 8192, "low_frequency_factor": 1.0,
               "high_frequency_factor": 4.0}


def llama3_8b(bits=4, group_size=64, layers=32) -> ModelSpec:
    return ModelSpec(name=f"llama3-8b-int{bits}", model_dim=4096, hidden_dim=14336, vocab_size=128256,
                     layer_kinds=["attn"] * layers, num_heads=32, num_groups=8, head_dim=128, rope=dict(LLAMA3_ROPE),
                     norm_eps=1e-5, quant=QuantSpec("int", bits, group_size, False))


def llama3_70b(bits=4, group_size=64, layers=80) -> ModelSpec:
    return ModelSpec(name=f"llama3-70b-int{bits}", model_dim=8192, hidden_dim=28672, vocab_size=128256,
                     layer_kinds=["attn"] * layers, num_heads=64, num_groups=8, head_dim=128, rope=dict(LLAMA3_ROPE),
                     norm_eps=1e-5, quant=QuantSpec("int", bits, group_size, False))


def qwen35_0p8b(bits=4, group_size=64, layers=24, hybrid=True, vocab=248320) -> ModelSpec:
    """Qwen3.5-0.8B: linear shapes pinned by the reference's bench table (tests/matmul/shape.rs:82-88);
    layer mix (3 DeltaNet : 1 attention), vocab and rotary fraction from the public HF config
    (not pinned in the reference). hybrid=False gives the dense-attention stand-in."""
    kinds = [("attn" if (i % 4 == 3) else "delta") for i in range(layers)] if hybrid else ["attn"] * layers
    rope = {"type": "UnscaledRoPEConfig", "base": 10000000.0, "max_sequence_length": 262144, "head_dim": 64}
    return ModelSpec(name=f"qwen3.5-0.8b-int{bits}" + ("" if hybrid else "-dense"), model_dim=1024, hidden_dim=3584,
                     vocab_size=vocab, layer_kinds=kinds, num_heads=8, num_groups=2, head_dim=256, rope=rope,
                     qk_norm=True, qk_norm_scale_offset=1.0, attn_gate=True, norm_eps=1e-6, norm_scale_offset=1.0,
                     tied_embeddings=True, quant=QuantSpec("int", bits, group_size, False),
                     dn_num_heads=16, dn_num_groups=16, dn_head_dim=128, dn_value_head_dim=128, dn_kernel_size=4,
                     max_sequence_length=262144)


def tiny(kind="llama", quant: QuantSpec | None = None, layers=2) -> ModelSpec:
    """Small models for tests: same structure, small dims."""
    quant = quant or QuantSpec("int", 4, 64, False)
    if kind == "llama":
        rope = dict(LLAMA3_ROPE, head_dim=64)
        return ModelSpec(name="tiny-llama", model_dim=256, hidden_dim=512, vocab_size=1000,
                         layer_kinds=["attn"] * layers, num_heads=4, num_groups=2, head_dim=64, rope=rope, quant=quant)
    if kind == "qwen-dense":
        rope = {"type": "UnscaledRoPEConfig", "base": 1e6, "max_sequence_length": 4096, "head_dim": 32}
        return ModelSpec(name="tiny-qwen-dense", model_dim=256, hidden_dim=512, vocab_size=1200,
                         layer_kinds=["attn"] * layers, num_heads=4, num_groups=2, head_dim=64, rope=rope,
                         qk_norm=True, qk_norm_scale_offset=1.0, attn_gate=True, norm_eps=1e-6, norm_scale_offset=1.0,
                         tied_embeddings=True, quant=quant)
    if kind == "qwen-hybrid":
        rope = {"type": "UnscaledRoPEConfig", "base": 1e6, "max_sequence_length": 4096, "head_dim": 32}
        kinds = [("attn" if (i % 2 == 1) else "delta") for i in range(layers)]
        return ModelSpec(name="tiny-qwen-hybrid", model_dim=256, hidden_dim=512, vocab_size=1200, layer_kinds=kinds,
                         num_heads=4, num_groups=2, head_dim=64, rope=rope, qk_norm=True, qk_norm_scale_offset=1.0,
                         attn_gate=True, norm_eps=1e-6, norm_scale_offset=1.0, tied_embeddings=True, quant=quant,
                         dn_num_heads=4, dn_num_groups=2, dn_head_dim=128, dn_value_head_dim=128, dn_kernel_size=4)
    if kind == "llama-512":        # dims that the fused decode kernels cover (k multiple of 512)
        rope = dict(LLAMA3_ROPE, head_dim=64)
        return ModelSpec(name="tiny-llama-512", model_dim=512, hidden_dim=1024, vocab_size=2048, layer_kinds=["attn"] * layers,
                         num_heads=8, num_groups=2, head_dim=64, rope=rope, quant=quant)
    if kind == "qwen-hybrid-512":
        rope = {"type": "UnscaledRoPEConfig", "base": 1e6, "max_sequence_length": 4096, "head_dim": 32}
        kinds = [("attn" if (i % 2 == 1) else "delta") for i in range(layers)]
        return ModelSpec(name="tiny-qwen-hybrid-512", model_dim=512, hidden_dim=1024, vocab_size=2048, layer_kinds=kinds,
                         num_heads=8, num_groups=2, head_dim=64, rope=rope, qk_norm=True, qk_norm_scale_offset=1.0, attn_gate=True,
                         norm_eps=1e-6, norm_scale_offset=1.0, tied_embeddings=True, quant=quant, dn_num_heads=4, dn_num_groups=4,
                         dn_head_dim=128, dn_value_head_dim=128, dn_kernel_size=4)
    raise ValueError(kind)


PRESETS = {"llama3-8b": llama3_8b, "llama3-70b": llama3_70b, "qwen3.5-0.8b": qwen35_0p8b}
