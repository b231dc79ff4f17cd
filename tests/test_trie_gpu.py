"""Speculative (trie) decode on the GPU: uzu_engine_trie_pass / uzu_engine_trie_accept against the oracle's speculation pass
(oracle/model.py forward(trie=) + accept, itself checked against flat passes in tests/test_trie_cpu.py) and against the engine's own
plain decode: whatever a proposer suggests, the verified output is what token-by-token decode produces
(engine/language_model/stream/stream.rs:550-657, trie.rs:262-296, mixer/attention/state.rs:174-237, mask.rs:21-29)."""
import numpy as np
import pytest

from oracle.model import OracleModel
from tests.test_engine_gpu import _logit_check
from tests.util import bf16_to_f32
from uzu_b200 import binding as B
from uzu_b200 import synth
from uzu_b200.trie import PRng, TrieNode

pytestmark = pytest.mark.gpu


def _tree(tokens):
    """root t0 -> {t1 -> {t2, t3}, t4 -> {t5 -> {t6 -> {t7}}}, t8}: branching at two depths, a deep chain, a lone leaf."""
    t = [int(x) for x in tokens]
    root = TrieNode(t[0], 0)
    a = TrieNode(t[1], 0); a.add(TrieNode(t[2], 0)); a.add(TrieNode(t[3], 0))
    b = TrieNode(t[4], 0); c = TrieNode(t[5], 0); d = TrieNode(t[6], 0); d.add(TrieNode(t[7], 0)); c.add(d); b.add(c)
    root.add(a); root.add(b); root.add(TrieNode(t[8], 0))
    return root


@pytest.mark.parametrize("kind,quant,prompt_len", [("llama", None, 19), ("qwen-dense", None, 19), ("llama-512", None, 700),
                                                    ("llama", synth.QuantSpec("int", 8, 64, False), 33)])
def test_trie_pass_and_accept_match_oracle(ctx, tmp_path, kind, quant, prompt_len):
    spec = synth.tiny(kind, quant=quant)
    path = synth.write_model(spec, tmp_path / "m", seed=51)
    rng = np.random.default_rng(8)
    prompt = rng.integers(0, spec.vocab_size, prompt_len)
    toks = rng.choice(spec.vocab_size, 9, replace=False)
    flat = _tree(toks).linearize()
    assert len(flat) == 9 and not flat.is_flat()
    ref = OracleModel(path, max_context=1024)
    ref.prefill(prompt)
    want = ref.forward(flat.token_ids(), trie=flat.nodes())
    with B.Engine(ctx, path, max_context_length=1024) as eng:
        assert eng.speculation_supported
        eng.prefill(prompt)
        sampled, got = eng.trie_pass(flat.token_ids(), flat.nodes(), want_logits=True)
        assert eng.context_length == prompt_len, "a speculation pass accepts nothing"
        for i in range(len(flat)):
            _logit_check(got[i:i + 1], want[i:i + 1], f"{kind} trie node {i}")
            f = bf16_to_f32(got[i])
            assert sampled[i] == int(np.flatnonzero(f == f.max())[0]), "greedy id = lowest-index argmax of the node's own logits"
        # every other entry point refuses to run over an unaccepted suffix
        with pytest.raises(B.UzuError):
            eng.step_host(1)
        with pytest.raises(B.UzuError):
            eng.trie_accept([0, 5, 4], 0)           # not increasing
        # keep the deep chain root -> t4 -> t5 -> t6 (flat indices 0, 4, 5, 6): rows 4..6 move down to 1..3
        accepted = [0, 4, 5, 6]
        ref.accept(accepted)
        nxt = int(toks[7])
        eng.trie_accept(accepted, nxt)
        assert eng.context_length == ref.context_length == prompt_len + 4
        for step in range(3):
            lr = ref.forward([nxt])
            eng.step_host(nxt)
            _logit_check(eng.last_logits(), lr, f"{kind} decode step {step} after accept")
            nxt = int(np.argmax(bf16_to_f32(lr[0])))


def test_trie_branch_equals_flat_pass_on_the_engine(ctx, tmp_path):
    """Engine self-consistency, no oracle: the nodes on a root-to-leaf path of a trie see what a flat pass over that path sees."""
    spec = synth.tiny("llama-512")
    path = synth.write_model(spec, tmp_path / "m", seed=52)
    rng = np.random.default_rng(9)
    prompt = rng.integers(0, spec.vocab_size, 40)
    toks = rng.choice(spec.vocab_size, 9, replace=False)
    flat = _tree(toks).linearize()
    with B.Engine(ctx, path, max_context_length=256) as eng:
        eng.prefill(prompt)
        eng.snapshot()
        _, spec_logits = eng.trie_pass(flat.token_ids(), flat.nodes(), want_logits=True)
        parents = flat.parents()
        for leaf in (2, 3, 7, 8):
            idx, p = [], leaf
            while p >= 0:
                idx.append(p); p = parents[p]
            idx.reverse()
            eng.restore()
            branch = [flat.token_ids()[i] for i in idx]
            want = eng.forward(branch, 0, len(branch))
            for row, i in enumerate(idx):
                a, b = bf16_to_f32(spec_logits[i]), bf16_to_f32(want[row])
                assert float(np.abs(a - b).max()) <= 0.01 * float(np.abs(b).max()) + 1e-3, (leaf, i)
        # a flat trie is a flat pass
        eng.restore()
        chain = TrieNode.flat(eng.context_length, toks[:6], PRng(0)).linearize()
        _, lt = eng.trie_pass(chain.token_ids(), chain.nodes(), want_logits=True)
        eng.restore()
        lf = eng.forward(toks[:6], 0, 6)
        a, b = bf16_to_f32(lt), bf16_to_f32(lf)
        assert float(np.abs(a - b).max()) <= 0.01 * float(np.abs(b).max()) + 1e-3


def _common_prefix(a, b):
    return next((i for i, (x, y) in enumerate(zip(a, b)) if x != y), min(len(a), len(b)))


@pytest.mark.parametrize("kind", ["llama-512", "qwen-dense"])
def test_speculative_generation_equals_plain_decode(ctx, tmp_path, kind):
    spec = synth.tiny(kind)
    path = synth.write_model(spec, tmp_path / "m", seed=53)
    prompt = (np.arange(30) * 37 + 5) % spec.vocab_size
    steps = 40
    with B.Engine(ctx, path, max_context_length=256) as eng:
        plain = eng.generate(prompt, steps)
        truth = {}                                   # position in the output -> token plain decode produced there

        def oracle_proposer(history, root, budget):
            """Proposes the true continuation (known from the plain run) mixed with wrong siblings: a perfect draft model plus noise."""
            k = len(history) - len(prompt)           # index of `root` in the output
            node = root_node = TrieNode(root, 0, 0.0)
            for d in range(1, 6):
                if k + d >= len(plain):
                    break
                node.add(TrieNode((plain[k + d] + 1) % spec.vocab_size, 0, -3.0))     # wrong sibling first
                good = TrieNode(plain[k + d], 0, -0.1)
                node.add(good)
                node = good
            return root_node

        eng.reset()
        stats = {}
        spec_out = eng.generate_speculative(prompt, steps, oracle_proposer, stats=stats)
        # identical unless a near-tie flips an argmax between the m = 1 and the m <= 16 GEMV (different summation order)
        assert _common_prefix(plain, spec_out) >= 12, (plain, spec_out)
        assert stats["tokens_per_pass"] > 2.0, stats
        # a proposer that is always wrong costs nothing but the pass: one token per pass, same output
        eng.reset()
        stats = {}
        bad = eng.generate_speculative(prompt, 12, lambda h, r, b: _wrong(r, spec.vocab_size), stats=stats)
        assert _common_prefix(plain, bad) >= 8 and stats["tokens_per_pass"] == 1.0
        # no proposer: single-node tries == plain decode through the speculation entry points
        eng.reset()
        none = eng.generate_speculative(prompt, 12, None)
        assert _common_prefix(plain, none) >= 8
        # and the plain path still works afterwards from the accepted state
        tail = [eng.step_host(none[-1]) for _ in range(3)]
        assert len(tail) == 3 and eng.context_length == len(prompt) + 12 + 2


def _wrong(root, vocab):
    n = TrieNode(root, 0)
    n.add(TrieNode(vocab - 1, 0)); n.add(TrieNode(vocab - 2, 0))
    return n


def test_speculative_stochastic_sampling_uses_per_node_seeds(ctx, tmp_path):
    """Seeded sampling: node seeds = PRng::derive(context + height) (dflash_tfm.rs:267,304), the same seed plain decode uses at that
    position (stream.rs:600), so a verified speculative run reproduces the plain seeded run."""
    spec = synth.tiny("llama-512")
    path = synth.write_model(spec, tmp_path / "m", seed=54)
    prompt = (np.arange(25) * 11 + 3) % spec.vocab_size
    sm = B.Engine.sampling(seed=1234, temperature=0.8, top_k=20)
    with B.Engine(ctx, path, max_context_length=256) as eng:
        plain = eng.generate(prompt, 24, sm)

        def proposer(history, root, budget):
            k = len(history) - len(prompt)
            node = root_node = TrieNode(root, 0)
            for d in range(1, 4):
                if k + d >= len(plain):
                    break
                nxt = TrieNode(plain[k + d], 0)
                node.add(nxt)
                node = nxt
            return root_node

        eng.reset()
        stats = {}
        out = eng.generate_speculative(prompt, 24, proposer, sampling=sm, stats=stats)
        assert _common_prefix(plain, out) >= 10, (plain, out)
        assert stats["tokens_per_pass"] > 1.5


def test_speculation_is_refused_where_the_reference_refuses_it(ctx, tmp_path):
    spec = synth.tiny("qwen-hybrid")
    path = synth.write_model(spec, tmp_path / "m", seed=55)
    with B.Engine(ctx, path, max_context_length=128) as eng:
        assert not eng.speculation_supported          # Mixer::speculation_supported without a tree-verify core (delta_net.rs:442-444)
        eng.prefill([1, 2, 3])
        flat = TrieNode.flat(3, [4, 5], PRng(0)).linearize()
        with pytest.raises(B.UzuError):
            eng.trie_pass(flat.token_ids(), flat.nodes())
    spec = synth.tiny("llama")
    path = synth.write_model(spec, tmp_path / "m2", seed=56)
    with B.Engine(ctx, path, max_context_length=128) as eng:
        eng.prefill([1, 2, 3])
        with pytest.raises(B.UzuError):
            eng.trie_accept([0], 0)                   # nothing pending
        with pytest.raises(B.UzuError):               # not a linearized trie: height jumps by 2
            eng.trie_pass([4, 5], np.array([[0, 1, 0], [1, 1, 2]], np.uint32))
        with pytest.raises(B.UzuError):               # more than the stream's 16-node budget
            big = TrieNode.flat(3, range(17), PRng(0)).linearize()
            eng.trie_pass(big.token_ids(), big.nodes())
        assert eng.step_host(7) >= 0                  # engine still usable
