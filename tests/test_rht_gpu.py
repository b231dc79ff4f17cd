"""Mirai RHT path on the GPU: ActivationTransform (uzu_b200/csrc/activation_transform.cu), the matmul's output-RHT epilogue
and a HybridSpec checkpoint through the engine, against the oracle (oracle_activation_transform, pinned by float64 twins in
tests/test_oracle_pins.py). Reference: backends/cpu/kernel/activation_transform/*.rs, cpu/kernel/matmul/kernel.rs:297-303,
encodable_block/linear/rht_wrapper.rs; shapes and inputs of tests/unit/backends/common/kernel/activation_transform_test.rs."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as O
from oracle.model import OracleModel
from tests import gpu_ops as G
from tests.test_engine_gpu import _logit_check
from tests.util import bf16_to_f32
from uzu_b200 import binding as B
from uzu_b200 import synth

pytestmark = pytest.mark.gpu

f32_to_bf16 = O.f32_to_bf16


def _case(batch, channels):
    data = (np.sin(np.arange(batch * channels, dtype=np.float64) * 0.1) * 2.0).reshape(batch, channels).astype(np.float32)
    factors = np.where(np.arange(channels) % 3 == 0, -1, 1).astype(np.int32)
    return data, factors


@pytest.mark.parametrize("batch,channels", [(1, 32), (1, 64), (1, 128), (4, 32), (4, 256), (2, 2048), (16, 4096), (3, 14336)])
def test_input_and_output_rht_bit_exact(ctx, batch, channels):
    """One warp = one 32-wide stripe, butterflies as xor-shuffles in the reference's stage order, every operation a single IEEE f32 op:
    the GPU result is the CPU result bit for bit, f32 and bf16, in place and out of place."""
    data, factors = _case(batch, channels)
    for x in (data, f32_to_bf16(data)):
        for op in (O.RHT_INPUT, O.RHT_OUTPUT):
            want = O.activation_transform(x, factors, op=op)
            for in_place in (False, True):
                got = G.activation_transform(ctx, x.copy(), factors, op=op, in_place=in_place)
                assert got.dtype == want.dtype and (got.view(np.uint8) == want.view(np.uint8)).all(), (x.dtype, op, in_place)


@pytest.mark.parametrize("group,sum_group", [(32, 0), (64, 0), (128, 0), (256, 0), (32, 32), (64, 32), (32, 128), (128, 128), (64, 256)])
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_quantize_ops_bit_exact(ctx, group, sum_group, dtype):
    rng = np.random.default_rng(0x5EED0001 + group + sum_group)
    rows, cols = 5, 1024
    x = rng.uniform(-1, 1, (rows, cols)).astype(np.float32)
    x[1, 256:512] = 0.0                      # an all-zero span: divisor 1, codes 0
    x[2, 7] = 40.0                           # an outlier: its group saturates at 127, neighbours collapse
    if dtype == "bf16":
        x = f32_to_bf16(x)
    factors = rng.choice(np.array([-1, 1], np.int32), cols)
    op = O.RHT_QUANTIZE_WITH_GROUP_SUMS if sum_group else O.RHT_QUANTIZE
    wq, ws, wg = O.activation_transform(x, factors, op=op, activation_group_size=group, sum_group_size=sum_group)
    gq, gs, gg = G.activation_transform(ctx, x, factors, op=op, activation_group_size=group, sum_group_size=sum_group)
    assert (gs.view(np.uint32) == ws.view(np.uint32)).all(), "divisors"
    assert (gq == wq).all(), "codes"
    if sum_group:
        assert (gg == wg).all() and (gg == gq.astype(np.int32).reshape(rows, cols // sum_group, sum_group).sum(axis=2)).all()
    else:
        assert gg is None


def test_activation_transform_validation(ctx):
    lib = B.load()
    ok = B.ActivationTransformArgs(fp_out=256, rht_factors=256, batch_size=1, element_count=64, ops=0, in_place=1, data_type=B.DT_BF16)
    assert lib.uzu_activation_transform_validate(C.byref(ok)) == 0
    for bad in (dict(element_count=48), dict(rht_factors=0), dict(ops=7), dict(data_type=99), dict(in_place=0),
                dict(ops=2, in_place=0, input=256, q_out=256, scales_out=256, activation_scale_group_size=48),
                dict(ops=3, in_place=0, input=256, q_out=256, scales_out=256, activation_scale_group_size=32, sum_group_size=32),   # no sums buffer
                dict(ops=2, in_place=0, input=256, q_out=256, scales_out=256, activation_scale_group_size=512, element_count=1024)):
        a = B.ActivationTransformArgs(fp_out=256, rht_factors=256, batch_size=1, element_count=64, ops=0, in_place=1, data_type=B.DT_BF16)
        for k, v in bad.items():
            setattr(a, k, v)
        assert lib.uzu_activation_transform_validate(C.byref(a)) != 0, bad


@pytest.mark.parametrize("m,n,k,bits,gs", [(1, 4096, 4096, 4, 64), (1, 64, 128, 4, 64), (4, 1024, 2048, 8, 64), (16, 6144, 4096, 4, 128),
                                           (70, 512, 1024, 4, 64)])
def test_matmul_output_rht_epilogue(ctx, m, n, k, bits, gs):
    """MatmulDOps::rht_factors: D = bias + s o H(round_bf16(A W^T)) (kernel.rs:285,297-303) for the decode GEMV, the m <= 16 GEMV and the
    tensor-core prefill GEMM alike. The pre-transform product differs from the oracle's by summation order (<= 1 bf16 ulp on a few
    elements); the orthogonal transform spreads that over the stripe, so the bound is relative to the stripe's magnitude."""
    rng = np.random.default_rng(m * 7 + n)
    w = rng.integers(0, 256, (n, k * bits // 8), dtype=np.uint8)
    groups = k // gs
    scales = f32_to_bf16(rng.uniform(0.01, 0.3, (n, groups)).astype(np.float32))
    zp = rng.integers(0, 256, (n, (groups + 1) // 2 if bits == 4 else groups), dtype=np.uint8)
    x = f32_to_bf16(rng.uniform(-0.3, 0.3, (m, k)).astype(np.float32))
    bias = f32_to_bf16(rng.uniform(-0.5, 0.5, n).astype(np.float32))
    factors = rng.choice(np.array([-1, 1], np.int32), n)
    kw = dict(m=m, n=n, k=k, scales=scales, zero_points=zp, method=O.QM_ZERO_POINT, bits=bits, group_size=gs)
    for b in (None, bias):
        want = bf16_to_f32(O.matmul(x, w, bias=b, rht_factors=factors, threads=8, **kw))
        got = bf16_to_f32(G.matmul(ctx, x, w, bias=b, rht_factors=factors, **kw))
        stripe_rms = np.sqrt((want.reshape(m, n // 32, 32) ** 2).mean(axis=2, keepdims=True))
        err = np.abs(got - want).reshape(m, n // 32, 32)
        assert (err <= 0.02 * stripe_rms + 2 ** -8 * np.abs(want).reshape(m, n // 32, 32) + 1e-3).all(), (b is not None, float(err.max()))
        assert float((got == want).mean()) > 0.5
    # no transform requested, no transform applied; and a transform without factors is refused
    plain = G.matmul(ctx, x, w, **kw)
    assert (bf16_to_f32(plain) != got).any()
    a = B.MatmulArgs(a=256, b=256, b_scales=256, b_zero_points=256, d=256, b_prologue=B.B_SCALE_ZERO_POINT, b_mode=B.QMODE_U4, b_group_size=64,
                     b_transpose=1, d_transform=B.D_RHT, ab_scale=1.0, m=1, n=64, k=128, weights_dt=B.DT_BF16, input_dt=B.DT_BF16, output_dt=B.DT_BF16)
    assert B.load().uzu_matmul_validate(C.byref(a)) != 0


@pytest.mark.parametrize("kind,quant", [("llama", synth.QuantSpec("int", 4, 64, False, rht=True)),
                                         ("qwen-dense", synth.QuantSpec("int", 4, 128, True, rht=True)),
                                         ("llama-512", synth.QuantSpec("mlx", 4, 64, rht=True)),
                                         ("qwen-hybrid", synth.QuantSpec("int", 8, 64, False, rht=True))])
def test_hybrid_spec_checkpoint_through_the_engine(ctx, tmp_path, kind, quant):
    """A checkpoint whose layer linears are HybridSpec (input_output RHT, block 32) loads, takes the unfused per-kernel decode path
    (input transform -> GEMV -> output transform), and matches the oracle model under teacher forcing: prefill rows, decode steps,
    the CUDA-graph decode loop."""
    spec = synth.tiny(kind, quant=quant)
    if kind == "qwen-hybrid":     # Qwen3.5's DeltaNet geometry: in_proj rows = 2 * 2048 + 2 * 2048 + 32 = 8224, a multiple of the 32-wide block
        spec.dn_num_heads = spec.dn_num_groups = 16
    path = synth.write_model(spec, tmp_path / "m", seed=61)
    rng = np.random.default_rng(12)
    prompt = rng.integers(0, spec.vocab_size, 23)
    ref = OracleModel(path, max_context=128)
    lr = ref.prefill(prompt)
    with B.Engine(ctx, path, max_context_length=128) as eng:
        assert not eng.persistent_decode and "RHT" in eng.persistent_decode_reason
        first = eng.prefill(prompt)
        tok = int(np.argmax(bf16_to_f32(lr[0])))
        f = bf16_to_f32(lr[0]); top = np.sort(f)[::-1]
        if top[0] - top[1] > 0.05 * abs(top[0]):
            assert first == tok
        for step in range(5):
            lr = ref.forward([tok])
            eng.step_host(tok)
            _logit_check(eng.last_logits(), lr, f"{kind} RHT decode step {step}")
            tok = int(np.argmax(bf16_to_f32(lr[0])))
        assert eng.context_length == ref.context_length
        eng.reset()
        a = eng.generate(prompt, 10)
        eng.reset()
        b = [eng.prefill(prompt)]
        for _ in range(9):
            b.append(eng.step_host(b[-1]))
        assert a == b
