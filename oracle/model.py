"""Whole-model CPU oracle: the reference's per-token op schedule driven over the C oracle kernels.

TEST INFRASTRUCTURE ONLY (see oracle/uzu_oracle.c). Restates, in order:
  Decoder::encode            encodable_block/decoder.rs:138-203
  Transformer::encode        encodable_block/transformer.rs:226-329
  TransformerLayer::encode   encodable_block/transformer_layer.rs:194-238
  Attention::attend          encodable_block/mixer/attention/mode.rs:45-144
  AttentionCores::encode     encodable_block/mixer/attention/core/mod.rs:74-98 (CPU: never gemm)
  DeltaNet::encode (m == 1)  encodable_block/mixer/delta_net.rs:473-535
  DenseMlp::encode           encodable_block/mlp/dense.rs:32-48
  Embedding lookup/readout   encodable_block/embedding.rs:345-456
  Sampling::encode           encodable_block/sampling/mod.rs:83-195
"""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

from . import oracle as O

ATTENTION_SUFFIX_CAPACITY = 1024  # mixer/attention/state.rs:14


def _load_safetensors(path):
    import struct
    with open(path, "rb") as f:
        (hlen,) = struct.unpack("<Q", f.read(8))
        header = json.loads(f.read(hlen))
    data = np.memmap(path, dtype=np.uint8, mode="r", offset=8 + hlen)
    meta = header.pop("__metadata__", {})
    np_dt = {"F32": np.float32, "BF16": np.uint16, "U8": np.uint8, "I8": np.int8, "U32": np.uint32, "I32": np.int32}
    tensors = {}
    for name, info in header.items():
        b, e = info["data_offsets"]
        tensors[name] = np.ascontiguousarray(np.frombuffer(data[b:e], dtype=np_dt[info["dtype"]]).reshape(info["shape"]))
    return tensors, meta


class _Linear:
    """LinearMatmul::load + encode_with_a (encodable_block/linear/matmul.rs:60-148): b_transpose = true."""

    def __init__(self, tensors, meta, prefix, out_dim, in_dim, threads):
        spec = json.loads(meta[prefix + ".spec"])
        self.out_dim, self.in_dim, self.threads = out_dim, in_dim, threads
        self.rht = None
        if spec["type"] == "HybridSpec":
            # RHTLinearWrapper (encodable_block/linear/{mod.rs:128-143, rht_wrapper.rs:58-176}): 32-wide input/output randomized Hadamard
            # around an inner quantized LinearMatmul stored under `quantized`, signs under `incoherence_signs`
            assert spec["adapter_spec"] is None and spec["incoherence_block_size"] == O.HADAMARD_TRANSFORM_BLOCK_SIZE \
                and spec["incoherence_processing_mode"] == "input_output", f"unsupported HybridSpec {spec}"
            self.rht = (tensors[prefix + ".incoherence_signs.input_signs"], tensors[prefix + ".incoherence_signs.output_signs"])
            assert self.rht[0].shape == (in_dim,) and self.rht[1].shape == (out_dim,) and self.rht[0].dtype == np.int32
            prefix = prefix + ".quantized"
            spec = json.loads(meta[prefix + ".spec"])
            assert spec["type"] in ("IntSpec", "MLXSpec"), "fused output-hadamard factors require quantized weights (linear/matmul.rs:88-91)"
        self.w = tensors[prefix + ".weights"]
        self.kw = {}
        t = spec["type"]
        if t == "FullPrecisionSpec":
            self.kw = dict(method=O.QM_NONE)
        else:
            self.kw = dict(bits=spec["bits"], group_size=spec["group_size"], scales=tensors[prefix + ".scales"])
            if t == "MLXSpec":
                self.kw.update(method=O.QM_SCALE_BIAS, biases=tensors[prefix + ".biases"])
            elif spec["is_symmetric"]:
                self.kw.update(method=O.QM_SYMMETRIC)
            else:
                self.kw.update(method=O.QM_ZERO_POINT, zero_points=tensors[prefix + ".zero_points"])
        self.spec = spec

    def __call__(self, x, d_f32=False):
        m = x.shape[0]
        if self.rht is not None:
            # encode_input (rht_wrapper.rs:286-296): InputRht over the activation, inner matmul with the output factors as its epilogue
            assert not d_f32
            x = O.activation_transform(x, self.rht[0], op=O.RHT_INPUT)
            return O.matmul(x, self.w, m=m, n=self.out_dim, k=self.in_dim, threads=self.threads, rht_factors=self.rht[1], **self.kw)
        return O.matmul(x, self.w, m=m, n=self.out_dim, k=self.in_dim, d_f32=d_f32, threads=self.threads, **self.kw)

    def lookup(self, token_ids, vocab, input_scale):
        if self.spec["type"] == "FullPrecisionSpec":
            return O.fp_embedding_lookup(token_ids, self.w, vocab_size=vocab, model_dim=self.in_dim,
                                         input_scale=input_scale)
        mode = O.MODE_U4 if self.kw["bits"] == 4 else O.MODE_U8
        return O.quant_embedding_lookup(token_ids, self.w, self.kw["scales"], zero_points=self.kw.get("zero_points"),
                                        biases=self.kw.get("biases"), vocab_size=vocab, model_dim=self.in_dim,
                                        input_scale=input_scale, group_size=self.kw["group_size"], mode=mode,
                                        method=self.kw["method"])


class OracleModel:
    """`tp_reduce` / `tp_gather`: the two exchange steps of a tensor-parallel shard written by uzu_b200.tp.shard_checkpoint
    (config.json carries a "tensor_parallel" block): f32 partial sums [m, H] of the row-parallel out / down projections are summed
    over ranks (all-reduce) and rounded to bf16 once; the vocab-parallel readout's [rows, V/P] logits are concatenated (all-gather).
    The reference has no tensor parallelism; the 1-rank model is the oracle for the P-rank one."""

    def __init__(self, path, threads: int = 1, max_context: int = 4096, tp_reduce=None, tp_gather=None):
        path = Path(path)
        cfg = json.loads((path / "config.json").read_text())
        self.cfg = cfg
        dec = cfg["decoder_config"]
        tr = dec["transformer_config"]
        self.H, self.F, self.V = tr["model_dim"], tr["hidden_dim"], dec["vocab_size"]
        self.tp = cfg.get("tensor_parallel")
        self.tp_reduce, self.tp_gather = tp_reduce, tp_gather
        if self.tp is not None and self.tp["size"] > 1:
            assert tp_reduce is not None and tp_gather is not None, "a tensor-parallel shard needs the exchange hooks"
        self.layers_cfg = tr["layer_configs"]
        self.out_norm_cfg = tr["output_norm_config"]
        self.emb_cfg = dec["embedding_config"]
        self.threads = threads
        self.max_context = max_context
        T, M = _load_safetensors(path / "model.safetensors")
        self.T = T
        p = "decoder.embedding."
        if self.emb_cfg["type"] == "TiedEmbeddingConfig":
            self.in_emb = self.out_emb = _Linear(T, M, p + "embedding", self.V, self.H, threads)
        else:
            self.in_emb = _Linear(T, M, p + "input_embedding", self.V, self.H, threads)
            v_out = self.tp["vocab_size_local"] if self.tp is not None else self.V
            self.out_emb = _Linear(T, M, p + "output_embedding", v_out, self.H, threads)
        self.layers = []
        for i, lc in enumerate(self.layers_cfg):
            lp = f"decoder.transformer.layers.{i}."
            mc = lc["mixer_config"]
            L = {"cfg": lc, "mixer": mc, "prefix": lp}
            F = lc["hidden_dim"] or self.F
            if mc["type"] == "AttentionConfig":
                D, Hq, Hkv = mc["head_dim"], mc["num_heads"], mc["num_groups"]
                L["qkv"] = _Linear(T, M, lp + "mixer.qkv_projection.weights", (Hq + 2 * Hkv) * D, self.H, threads)
                L["out"] = _Linear(T, M, lp + "mixer.out_projection.weights", self.H, Hq * D, threads)
                if mc["gate_projection_config"] is not None:
                    L["gate"] = _Linear(T, M, lp + "mixer.gate_projection.weights", Hq * D, self.H, threads)
            else:
                kd = mc["num_groups"] * mc["head_dim"]
                vd = mc["num_heads"] * mc["value_head_dim"]
                total = 2 * kd + vd + vd + 2 * mc["num_heads"]
                L["in_proj"] = _Linear(T, M, lp + "mixer.in_proj.weights", total, self.H, threads)
                L["out_proj"] = _Linear(T, M, lp + "mixer.out_proj.weights", self.H, vd, threads)
            L["up"] = _Linear(T, M, lp + "mlp.up_projection.weights", 2 * F, self.H, threads)
            L["down"] = _Linear(T, M, lp + "mlp.down_projection.weights", self.H, F, threads)
            L["F"] = F
            self.layers.append(L)
        self.reset()

    # ---- state (LanguageModelState / TransformerState) ----
    def reset(self):
        self.context_length = 0
        self._pending = None
        self.state = []
        for L in self.layers:
            mc = L["mixer"]
            if mc["type"] == "AttentionConfig":
                rows = self.max_context + ATTENTION_SUFFIX_CAPACITY
                e = mc["num_groups"] * mc["head_dim"]
                self.state.append({"k": np.zeros((rows, e), np.uint16), "v": np.zeros((rows, e), np.uint16), "len": 0})
            else:
                kd = mc["num_groups"] * mc["head_dim"]
                vd = mc["num_heads"] * mc["value_head_dim"]
                conv_dim = 2 * kd + vd
                self.state.append({"conv": np.zeros((conv_dim, mc["kernel_size"] - 1), np.float32),
                                   "ssm": np.zeros((mc["num_heads"], mc["value_head_dim"], mc["head_dim"]), np.float32)})

    def _norm(self, x, key, ncfg, shortcut=None, residual_add=False):
        return O.normalization(x, self.T[key + ".scales"], shortcut=shortcut, residual_add=residual_add,
                               epsilon=ncfg["epsilon"], scale_offset=ncfg["scale_offset"] or 0.0,
                               full_layer=ncfg["upcast_mode"] == "full_layer", subtract_mean=ncfg["subtract_mean"])

    def _attention(self, L, st, hidden, positions, trie=None):
        mc = L["mixer"]
        m = hidden.shape[0]
        D, Hq, Hkv = mc["head_dim"], mc["num_heads"], mc["num_groups"]
        gate = None
        if "gate" in L:  # mode.rs:54-61
            gate = L["gate"](hidden)
        qkv = L["qkv"](hidden)
        total_heads = Hq + 2 * Hkv
        for name, off, cnt in (("query_norm", 0, Hq), ("key_norm", Hq, Hkv)):
            ncfg = mc[name + "_config"]
            if ncfg is not None:
                O.qkv_norm(qkv, self.T[L["prefix"] + "mixer." + name + ".scales"], total_heads=total_heads, head_dim=D,
                           epsilon=ncfg["epsilon"], scale_offset=ncfg["scale_offset"] or 0.0, head_offset=off,
                           head_count=cnt, full_layer=ncfg["upcast_mode"] == "full_layer")
        rope = L["cfg"]["rope_config"]
        cos = sin = None
        rope_dim = 0
        if rope is not None:
            cos, sin = O.rope_tables(rope, positions)
            rope_dim = rope["head_dim"]
        prefix = st["len"]
        queries = O.attention_prepare(qkv, st["k"], st["v"], cos, sin, num_q_heads=Hq, num_kv_heads=Hkv, head_dim=D,
                                      rope_dim=rope_dim, kv_token_offset=prefix)
        scale = mc["scale"] if mc["scale"] is not None else float(np.float32(1.0) / np.sqrt(np.float32(D)))
        kw = dict(head_dim=D, gqa_factor=Hq // Hkv, sequence_length=prefix + m, k_head_stride=D, k_seq_stride=Hkv * D,
                  v_head_stride=D, v_seq_stride=Hkv * D, scale=scale, num_heads=Hq, suffix_length=m,
                  is_causal=mc["is_causal"])
        if trie is not None:   # run_core picks trie_core + uploads the nodes when the topology is not flat (mode.rs:178-184)
            kw["trie"] = trie
        if prefix + m > 1024:  # core/mod.rs:88-92
            out = O.attention_two_pass(queries, st["k"], st["v"], **kw)
        else:
            out = O.attention_single_pass(queries, st["k"], st["v"], **kw)
        out = out.reshape(m, Hq * D)
        if gate is not None:
            O.sigmoid_gate(gate, out)
        if trie is None:
            st["len"] = prefix + m  # encode_accept(0..m): flat path, no copies (state.rs:174-237)
        return self._row_parallel(L["out"], out)

    def _row_parallel(self, linear, x):
        """out / down projection. Tensor-parallel shard: this rank's K slice gives an f32 partial; sum over ranks, round once."""
        if self.tp is None or self.tp["size"] == 1:
            return linear(x)
        return O.f32_to_bf16(self.tp_reduce(linear(x, d_f32=True)))

    def _delta_net(self, L, st, hidden):
        mc = L["mixer"]
        assert hidden.shape[0] == 1, "oracle DeltaNet covers the flat decode branch (m == 1) only"
        kd = mc["num_groups"] * mc["head_dim"]
        vd = mc["num_heads"] * mc["value_head_dim"]
        conv_dim = 2 * kd + vd
        pre = L["prefix"] + "mixer."
        in_proj = L["in_proj"](hidden)
        row = in_proj[0]
        O.delta_net_conv_update(self.T[pre + "conv.weights"], None, row, st["conv"], mc["kernel_size"], conv_dim)
        out = O.delta_net_update(row, self.T[pre + "a_log"], self.T[pre + "dt_bias"], self.T[pre + "norm.scales"],
                                 st["ssm"], num_v_heads=mc["num_heads"], num_k_heads=mc["num_groups"],
                                 head_k_dim=mc["head_dim"], head_v_dim=mc["value_head_dim"], key_dim=kd, value_dim=vd,
                                 norm_epsilon=mc["norm_config"]["epsilon"])
        return L["out_proj"](out.reshape(1, vd))

    def forward(self, token_ids, output_rows=None, return_hidden=False, trie=None):
        """One Decoder::encode over `token_ids` (m rows). Returns bf16 logits (bits) for `output_rows` (default: last row only,
        like the stream does). `trie` = u32 [m, 3] rows of (trie_start, trie_end, height) from FlatTrie::token_subtrie_ranges
        (trie.rs:211-222) for a speculation pass: positions are context + height (transformer.rs:248), attention masks suffix keys
        by subtrie range (mask.rs:21-29), and NOTHING is accepted -- call accept() with the indices the stream kept."""
        token_ids = np.ascontiguousarray(token_ids, dtype=np.uint32)
        m = len(token_ids)
        if trie is not None:
            trie = np.ascontiguousarray(trie, dtype=np.uint32).reshape(m, 3)
            assert self._pending is None, "a speculation pass is already pending: accept() first"
            assert all(L["mixer"]["type"] == "AttentionConfig" for L in self.layers), \
                "DeltaNet tree verification is not restated (Mixer::speculation_supported is false without it, delta_net.rs:442-444)"
            if output_rows is None:
                output_rows = (0, m)          # stream.rs:640: Some(0..batch_dim.size())
        if output_rows is None:
            output_rows = (m - 1, m)
        input_scale = self.emb_cfg["input_scale"] if self.emb_cfg["input_scale"] is not None else 1.0
        hidden = self.in_emb.lookup(token_ids, self.V, input_scale)
        shortcut = np.zeros_like(hidden)
        if trie is not None:
            positions = (self.context_length + trie[:, 2]).astype(np.uint32)
        else:
            positions = np.arange(self.context_length, self.context_length + m, dtype=np.uint32)
        for i, L in enumerate(self.layers):
            lc = L["cfg"]
            hidden = self._norm(hidden, L["prefix"] + "pre_mixer_norm", lc["pre_mixer_norm_config"], shortcut=shortcut,
                                residual_add=i > 0)
            if L["mixer"]["type"] == "AttentionConfig":
                hidden = self._attention(L, self.state[i], hidden, positions, trie)
            else:
                hidden = self._delta_net(L, self.state[i], hidden)
            hidden = self._norm(hidden, L["prefix"] + "pre_mlp_norm", lc["pre_mlp_norm_config"], shortcut=shortcut,
                                residual_add=True)
            up = L["up"](hidden)
            act = L["cfg"]["mlp_config"]["activation"]
            assert act["type"] == "SiLU"
            gated = O.gated_act_mul(up, L["F"], O.ACT_SILU)
            hidden = self._row_parallel(L["down"], gated)
        b, e = output_rows
        # output_norm over rows [b,e) with residual add into the same rows of shortcut (transformer.rs:317-323)
        sc_rows = np.ascontiguousarray(shortcut[b:e])
        normed = self._norm(np.ascontiguousarray(hidden[b:e]), "decoder.transformer.output_norm", self.out_norm_cfg,
                            shortcut=sc_rows, residual_add=True)
        logits = self.out_emb(normed)
        if self.tp is not None and self.tp["size"] > 1:
            logits = self.tp_gather(logits)          # [rows, V/P] per rank -> [rows, V]
        if self.emb_cfg["logit_scale"] is not None or self.emb_cfg["logit_soft_cap"] is not None:
            O.logit_transform(logits, self.emb_cfg["logit_scale"] or 1.0, self.emb_cfg["logit_soft_cap"])
        if trie is None:
            self.context_length += m
        else:
            self._pending = m
        if return_hidden:
            return logits, normed
        return logits

    def accept(self, accepted_indices):
        """TransformerState::encode_accept after a speculation pass (transformer.rs:56-75 -> mixer/attention/state.rs:174-237, Full
        state): row `length + accepted[i]` of every layer's K/V moves to row `length + i`; the context grows by len(accepted)."""
        idx = [int(i) for i in accepted_indices]
        assert self._pending is not None, "no speculation pass to accept"
        assert idx and all(a < b for a, b in zip(idx, idx[1:])) and idx[-1] < self._pending, "invalid accepted indices"
        for L, st in zip(self.layers, self.state):
            copies = [(st["len"] + a, st["len"] + i) for i, a in enumerate(idx) if a != i]
            if copies:
                O.kv_cache_update(st["k"], st["v"], copies, L["mixer"]["num_groups"] * L["mixer"]["head_dim"])
            st["len"] += len(idx)
        self.context_length += len(idx)
        self._pending = None

    def prefill(self, prompt):
        """Chunks of <= 1024 (stream.rs:194-195); hybrid (DeltaNet) models step token by token
        because the oracle restates only the m == 1 DeltaNet branch. Returns last-row logits."""
        prompt = np.asarray(prompt, dtype=np.uint32)
        has_delta = any(L["mixer"]["type"] != "AttentionConfig" for L in self.layers)
        step = 1 if has_delta else 1024
        logits = None
        for s in range(0, len(prompt), step):
            logits = self.forward(prompt[s:s + step])
        return logits

    def generate(self, prompt, steps, seed=None, **sampling):
        """Greedy (seed None) or seeded stochastic decode; returns (tokens, list of logits)."""
        logits = self.prefill(prompt)
        toks, all_logits = [], []
        for _ in range(steps):
            all_logits.append(logits.copy())
            seeds = None
            if seed is not None:
                # PRng::derive(position), position = absolute index of the sampled row (stream.rs:600)
                seeds = np.array([O.lib().oracle_prng_derive(seed, self.context_length - 1)], dtype=np.uint64)
            tok = int(O.unified_sampling(logits, seeds=seeds, **sampling)[0])
            toks.append(tok)
            logits = self.forward([tok])
        return toks, all_logits
