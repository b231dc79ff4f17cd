"""Parity of the tensor-core prefill attention (uzu_b200/csrc/attention_prefill.cu) against the CPU oracle's
AttentionSinglePass restatement (backends/cpu/kernel/attention/attention_single_pass.rs:49-126) and the independent float64 softmax of
tests/unit/encodable_block/attention_test.rs:26-124, on the reference tests' closed-form inputs (attention_single_pass_test.rs:33-78).

First hardware run of these tests: round 2 (7 passed, profiles/r2_first_hardware_run.txt); since then the kernel is the default prefill
attention for head_dim 64 / 128 (UZU_PREFILL_ATTN=0 restores the split-KV kernel applied per query token)."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests import gpu_ops as G
from tests.test_oracle_pins import attention_inputs, softmax_reference
from tests.util import assert_bf16_close, bf16_to_f32, f32_to_bf16

pytestmark = [pytest.mark.gpu]


@pytest.fixture
def prefill_attn(ctx):
    ctx.lib.uzu_debug_set_prefill_attention(1)
    yield ctx
    ctx.lib.uzu_debug_set_prefill_attention(-1)


# (heads, kv heads, sequence, suffix, head_dim, causal)
CASES = [(4, 4, 64, 64, 64, True), (8, 2, 128, 128, 128, True), (8, 1, 200, 37, 128, True), (4, 2, 96, 16, 64, False),
         (8, 2, 300, 100, 64, True), (32, 8, 512, 256, 128, True)]


@pytest.mark.parametrize("H,Hkv,seq,suffix,D,causal", CASES)
def test_prefill_attention_head_major_inputs(prefill_attn, H, Hkv, seq, suffix, D, causal):
    ctx = prefill_attn
    q, k, v = attention_inputs(H, Hkv, seq, suffix, D)        # K / V laid out [kv_head, seq, D] like the reference's kernel tests
    kw = dict(head_dim=D, gqa_factor=H // Hkv, sequence_length=seq, k_head_stride=seq * D, k_seq_stride=D, v_head_stride=seq * D,
              v_seq_stride=D, scale=float(np.float32(1.0) / np.sqrt(np.float32(D))), num_heads=H, suffix_length=suffix, is_causal=causal)
    ref = O.attention_single_pass(q, k, v, **kw)
    got = G.attention_single_pass(ctx, q, k, v, **kw)
    assert_bf16_close(got, ref, max_ulp=1, min_exact=0.95, what="prefill attention vs oracle")
    ind = softmax_reference(q, k, v, H, Hkv, seq, suffix, D, causal)
    assert np.abs(bf16_to_f32(got).astype(np.float64) - ind).max() < 4e-3


def test_prefill_attention_token_major_cache_matches_decode_kernel(prefill_attn):
    """The engine's layout (token-major cache, head stride = D) with a prefix: the tensor-core path and the split-KV decode kernel
    (parity-tested in test_kernels_gpu.py) agree to bf16 rounding on random inputs."""
    ctx = prefill_attn
    rng = np.random.default_rng(31)
    H, Hkv, D, prefix, suffix = 8, 2, 128, 77, 45
    q = f32_to_bf16(rng.standard_normal((H, suffix, D)).astype(np.float32))
    k = f32_to_bf16(rng.standard_normal((prefix + suffix, Hkv * D)).astype(np.float32))
    v = f32_to_bf16(rng.standard_normal((prefix + suffix, Hkv * D)).astype(np.float32))
    kw = dict(head_dim=D, gqa_factor=H // Hkv, sequence_length=prefix + suffix, k_head_stride=D, k_seq_stride=Hkv * D, v_head_stride=D,
              v_seq_stride=Hkv * D, scale=float(1.0 / np.sqrt(D)), num_heads=H, suffix_length=suffix, is_causal=True)
    got = G.attention_single_pass(ctx, q, k, v, **kw)
    ctx.lib.uzu_debug_set_prefill_attention(0)
    base = G.attention_single_pass(ctx, q, k, v, **kw)
    ctx.lib.uzu_debug_set_prefill_attention(1)
    assert_bf16_close(got, base, max_ulp=1, min_exact=0.9, what="tensor-core vs split-KV kernel")
    assert_bf16_close(got, O.attention_single_pass(q, k, v, **kw), max_ulp=1, min_exact=0.95, what="vs oracle")
