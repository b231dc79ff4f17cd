"""Multi-sequence batched decode (extension API uzu_engine_batch_*, BASELINE config 4 "batch = 8"): B independent sequences sharing one pass
over the weights per step, against B independent oracle models (the reference decodes one sequence at a time: N independent reference runs
are the oracle for an N-sequence batch).

First hardware run: round 2 (3 passed, profiles/r2_first_hardware_run.txt)."""
import os

import numpy as np
import pytest

from oracle.model import OracleModel
from tests.test_engine_gpu import _logit_check
from tests.util import bf16_to_f32
from uzu_b200 import binding as B
from uzu_b200 import synth

pytestmark = [pytest.mark.gpu]


@pytest.mark.parametrize("graph", [False, True])
@pytest.mark.parametrize("kind,nseq", [("llama", 3), ("qwen-hybrid", 2), ("llama-512", 8)])
def test_batched_decode_matches_independent_oracles(ctx, tmp_path, kind, nseq, graph):
    """graph=True: the batched step is captured once into a CUDA graph and replayed with per-sequence positions read from the device."""
    spec = synth.tiny(kind)
    path = synth.write_model(spec, tmp_path / "m", seed=41)
    rng = np.random.default_rng(9)
    prompts = [rng.integers(0, spec.vocab_size, 5 + 7 * b) for b in range(nseq)]        # different lengths -> different positions per row
    refs = [OracleModel(path, max_context=128) for _ in range(nseq)]
    ref_logits = [r.prefill(p) for r, p in zip(refs, prompts)]
    with B.Engine(ctx, path, max_context_length=128, use_cuda_graph=graph) as eng:
        eng.batch_begin(nseq)
        firsts = [eng.batch_prefill(b, prompts[b]) for b in range(nseq)]
        toks = []
        for b in range(nseq):
            l = np.sort(bf16_to_f32(ref_logits[b][0]))[::-1]
            want = int(np.argmax(bf16_to_f32(ref_logits[b][0])))
            if l[0] - l[1] > 0.05 * abs(l[0]):
                assert firsts[b] == want, (b, firsts[b], want)
            toks.append(want)                                                             # teacher forcing with the oracle's tokens
        for step in range(5):
            got = eng.batch_step(toks)
            lg = eng.batch_logits()
            nxt = []
            for b in range(nseq):
                lr = refs[b].forward([toks[b]])
                _logit_check(lg[b:b + 1], lr, f"{kind} seq {b} step {step}")
                want = int(np.argmax(bf16_to_f32(lr[0])))
                l = np.sort(bf16_to_f32(lr[0]))[::-1]
                if l[0] - l[1] > 0.05 * abs(l[0]):
                    assert got[b] == want, (b, step, got[b], want)
                nxt.append(want)
            toks = nxt
        for b in range(nseq):
            assert ctx.lib.uzu_engine_batch_context_length(eng.h, b) == len(prompts[b]) + 5
        assert eng.batch_decode_timed(toks, 4) > 0.0                                     # device-chained steps run


def test_batch_resize_frees_sequences_and_recaptures_the_graph(ctx, tmp_path):
    """batch_begin with a different sequence count: dropped sequences release their state, re-created ones get fresh buffers, and the captured
    batched step (which holds every sequence's buffer addresses) is re-captured -- the logits after 4 -> 2 -> 4 sequences equal those of a
    fresh 4-sequence engine bit for bit."""
    spec = synth.tiny("qwen-hybrid-512")
    path = synth.write_model(spec, tmp_path / "m", seed=43)
    rng = np.random.default_rng(10)
    prompts = [rng.integers(0, spec.vocab_size, 6 + 3 * b) for b in range(4)]

    def run(eng, nseq, steps=3):
        eng.batch_begin(nseq)
        toks = [eng.batch_prefill(b, prompts[b]) for b in range(nseq)]
        out = []
        for _ in range(steps):
            toks = eng.batch_step(toks)
            out.append(eng.batch_logits()[:nseq].copy())
        return out

    with B.Engine(ctx, path, max_context_length=128, use_cuda_graph=True) as eng:
        first = run(eng, 4)
        two = run(eng, 2)
        again = run(eng, 4)
    with B.Engine(ctx, path, max_context_length=128, use_cuda_graph=True) as fresh:
        want = run(fresh, 4)
    for a, b, c in zip(first, again, want):
        assert (a == c).all() and (b == c).all()
    with B.Engine(ctx, path, max_context_length=128, use_cuda_graph=True) as fresh2:
        want2 = run(fresh2, 2)
    for a, c in zip(two, want2):
        assert (a == c).all()
