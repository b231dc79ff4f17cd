"""Multi-token DeltaNet prefill kernel (uzu_b200/csrc/deltanet_prefill.cu: the m-token recurrence of a layer in ONE launch, state
resident in shared memory) against the oracle's token-by-token DeltaNet (backends/cpu/kernel/gdn/{conv_update,update}.rs restated in C) and
against the default path (the parity-tested decode kernel launched once per token inside the batched pass).

First hardware run: round 2 (3 passed, profiles/r2_first_hardware_run.txt); the kernel is now the default hybrid prefill recurrence
(UZU_DELTA_PREFILL_KERNEL=0 restores one decode-kernel launch per token)."""
import os

import numpy as np
import pytest

from oracle.model import OracleModel
from tests.test_engine_gpu import _logit_check
from tests.util import bf16_to_f32
from uzu_b200 import binding as B
from uzu_b200 import synth

pytestmark = [pytest.mark.gpu]


@pytest.mark.parametrize("n", [2, 21, 90])
def test_one_launch_delta_prefill_matches_oracle_and_per_token_path(ctx, tmp_path, n):
    spec = synth.tiny("qwen-hybrid-512")               # 4 v heads == 4 k heads x 128: the geometry the kernel covers (like Qwen3.5-0.8B's 16 / 16)
    path = synth.write_model(spec, tmp_path / "m", seed=29)
    rng = np.random.default_rng(8)
    prompt = rng.integers(0, spec.vocab_size, n)
    ref = OracleModel(path, max_context=256)
    for t in prompt:
        lr = ref.forward([t])
    outs = {}
    for mode in (1, 0):
        ctx.lib.uzu_debug_set_delta_prefill(mode)
        try:
            with B.Engine(ctx, path, max_context_length=256, use_cuda_graph=False) as eng:
                launches0 = eng.launch_count
                lg = eng.forward(prompt)
                outs[mode] = (lg, eng.launch_count - launches0)
                _logit_check(lg, lr, f"delta prefill mode {mode}")
                tok = 3
                ref_d = OracleModel(path, max_context=256)
                for t in prompt:
                    ref_d.forward([t])
                for step in range(3):                  # the state left behind continues correctly through the decode kernels
                    _logit_check(eng.forward([tok]), ref_d.forward([tok]), f"decode after prefill mode {mode} step {step}")
        finally:
            ctx.lib.uzu_debug_set_delta_prefill(-1)
    assert outs[1][1] < outs[0][1], "one launch per DeltaNet layer instead of one per token"
    g1, g0 = bf16_to_f32(outs[1][0][0]), bf16_to_f32(outs[0][0][0])
    assert np.abs(g1 - g0).max() <= 0.02 * np.abs(g0).max() + 1e-3
