"""Speculation trie on the host (uzu_b200/trie.py) against the reference's own unit tests (crates/backend-uzu/tests/unit/trie_test.rs,
restated case by case: same trees, same expected orders / heights / seeds / pruned sets), and the oracle's speculation pass
(oracle/model.py forward(trie=) + accept) against its own flat passes: a branch of a trie must see exactly what a flat pass over that
branch sees (mask.rs:21-29, transformer.rs:248, mixer/attention/state.rs:174-237)."""
import numpy as np
import pytest

from oracle.model import OracleModel
from uzu_b200 import synth
from uzu_b200.trie import DuplicateTokenId, PRng, TrieNode


def test_prng_derive_matches_the_engine_formula():
    # prng.rs:12-23; the same three constants as the device-side derive in engine.cu (decode_step_begin_kernel)
    r = PRng(0)
    assert r.derive(0) == 0
    assert PRng(1).derive(0) == PRng(0).derive(1)
    h = 12345 + 7
    h ^= h >> 33; h = (h * 0xff51afd7ed558ccd) % 2**64
    h ^= h >> 33; h = (h * 0xc4ceb9fe1a85ec53) % 2**64
    h ^= h >> 33
    assert PRng(12345).derive(7) == h


def test_trie_manual_sprout():
    root = TrieNode(0, 0)
    flat = root.linearize()
    assert len(flat) == 1 and flat.index(root) == 0
    for other in (TrieNode(1, 0), TrieNode(0, 1), TrieNode(0, 0)):
        assert flat.index(other) is None       # identity, not equality
    assert flat.token_ids() == [0] and flat.heights() == [0] and flat.token_seeds() == [0]


def test_trie_manual_stick():
    rng = PRng(0)
    stick = TrieNode(9, rng.derive(9))
    for i in range(8, 0, -1):
        parent = TrieNode(i, rng.derive(i))
        parent.add(stick)
        stick = parent
    root = TrieNode(0, rng.derive(0))
    root.add(stick)
    flat = root.linearize()
    assert len(flat) == 10
    ids, heights, seeds = flat.token_ids(), flat.heights(), flat.token_seeds()
    cur = root
    assert (ids[0], heights[0], seeds[0]) == (0, 0, rng.derive(0))
    for i in range(1, 10):
        cur = cur.get(i)
        assert cur.token == i and cur.seed == rng.derive(i)
        p = flat.index(cur)
        assert (ids[p], heights[p], seeds[p]) == (i, i, rng.derive(i))
    assert flat.is_flat() and flat.parents() == [-1] + list(range(9))
    # TrieNode::flat builds the same chain (trie.rs:139-156)
    same = TrieNode.flat(0, range(10), rng).linearize()
    assert same.token_ids() == ids and same.token_seeds() == seeds and (same.nodes() == flat.nodes()).all()


def test_trie_manual_bush():
    rng = PRng(0)
    root = TrieNode(0, rng.derive(0))
    root.add(TrieNode(1, rng.derive(1)))
    with pytest.raises(DuplicateTokenId):
        root.add(TrieNode(1, rng.derive(1)))
    with pytest.raises(DuplicateTokenId):
        root.add(TrieNode(1, 10))
    root.add(TrieNode(2, rng.derive(1)))
    root.add(TrieNode(3, rng.derive(1)))
    flat = root.linearize()
    assert len(flat) == 4
    ids, heights, seeds = flat.token_ids(), flat.heights(), flat.token_seeds()
    assert (ids[0], heights[0], seeds[0]) == (0, 0, rng.derive(0))
    for leaf_token in (1, 2, 3):
        p = flat.index(root.get(leaf_token))
        assert (ids[p], heights[p], seeds[p]) == (leaf_token, 1, rng.derive(1))
    assert flat.nodes().tolist() == [[0, 3, 0], [1, 1, 1], [2, 2, 1], [3, 3, 1]]
    assert flat.parents() == [-1, 0, 0, 0] and not flat.is_flat()


def _tree(rng):
    root = TrieNode(0, rng.derive(0))
    root.add(TrieNode(1, rng.derive(1)))
    mid_b = TrieNode(2, rng.derive(1))
    mid_b.add(TrieNode(10, rng.derive(2)))
    mid_c = TrieNode(3, rng.derive(1))
    mid_c.add(TrieNode(20, rng.derive(2)))
    mid_c.add(TrieNode(21, rng.derive(2)))
    root.add(mid_b)
    root.add(mid_c)
    return root


def test_trie_manual_tree():
    rng = PRng(0)
    root = _tree(rng)
    flat = root.linearize()
    assert len(flat) == 7
    ids, heights, seeds = flat.token_ids(), flat.heights(), flat.token_seeds()
    assert (ids[0], heights[0], seeds[0]) == (0, 0, rng.derive(0))
    for mid in (1, 2, 3):
        p = flat.index(root.get(mid))
        assert (ids[p], heights[p], seeds[p]) == (mid, 1, rng.derive(1))
    for mid, leaf in ((2, 10), (3, 20), (3, 21)):
        p = flat.index(root.get(mid).get(leaf))
        assert (ids[p], heights[p], seeds[p]) == (leaf, 2, rng.derive(2))
    assert ids == [0, 1, 2, 10, 3, 20, 21]
    assert flat.nodes().tolist() == [[0, 6, 0], [1, 1, 1], [2, 3, 1], [3, 3, 2], [4, 6, 1], [5, 5, 2], [6, 6, 2]]
    assert flat.parents() == [-1, 0, 0, 2, 0, 4, 4]


def _sample_tree():
    root = TrieNode(0, 0, 0.0)
    a = TrieNode(1, 1, -0.1); a.add(TrieNode(4, 2, -0.4))
    b = TrieNode(2, 1, -0.2); b.add(TrieNode(5, 2, -2.8))
    root.add(a); root.add(b); root.add(TrieNode(3, 1, -0.3))
    return root


@pytest.mark.parametrize("budget,expected", [(4, [0, 1, 2, 3]), (2, [0, 1]), (6, [0, 1, 4, 2, 5, 3]), (100, [0, 1, 4, 2, 5, 3])])
def test_trie_prune_to_budget(budget, expected):
    t = _sample_tree()
    t.prune_to_budget(budget)
    assert t.node_count() == len(expected) and t.linearize().token_ids() == expected
    if budget == 4:
        assert [t.get(k).logprob for k in (1, 2, 3)] == [-0.1, -0.2, -0.3]


def test_trie_prune_to_budget_tie_keeps_parent():
    root = TrieNode(0, 0, 0.0)
    child = TrieNode(1, 1, 0.0); child.add(TrieNode(2, 2, 0.0))
    root.add(child); root.add(TrieNode(3, 1, 0.0))
    root.prune_to_budget(2)
    assert root.linearize().token_ids() == [0, 1]


def test_flat_trie_accept_walks_the_verified_path():
    # trie.rs:262-296: descend while the sampled token at a node is one of its proposed children
    flat = _tree(PRng(0)).linearize()          # order [0, 1, 2, 10, 3, 20, 21]
    sampled = [3, 99, 99, 99, 21, 99, 77]      # root -> 3 (proposed), node 3 -> 21 (proposed), node 21 -> 77 (fresh)
    assert flat.accept(sampled) == [(0, 0, 3), (4, 3, 21), (6, 21, 77)]
    assert flat.accept([5, 0, 0, 0, 0, 0, 0]) == [(0, 0, 5)]     # nothing proposed matched: one token, like plain decode


# ---- the oracle's speculation pass --------------------------------------------------------------------------------------------

@pytest.mark.parametrize("kind", ["llama", "qwen-dense"])
def test_oracle_trie_branch_equals_flat_pass_over_the_branch(tmp_path, kind):
    spec = synth.tiny(kind)
    path = synth.write_model(spec, tmp_path / "m", seed=41)
    rng = np.random.default_rng(5)
    prompt = rng.integers(0, spec.vocab_size, 11)
    # root 7 -> {8 -> {9, 12}, 30 -> {31 -> {32}}}
    root = TrieNode(7, 0)
    a = TrieNode(8, 0); a.add(TrieNode(9, 0)); a.add(TrieNode(12, 0))
    b = TrieNode(30, 0); c = TrieNode(31, 0); c.add(TrieNode(32, 0)); b.add(c)
    root.add(a); root.add(b)
    flat = root.linearize()
    assert flat.token_ids() == [7, 8, 9, 12, 30, 31, 32]
    m = OracleModel(path, max_context=64)
    m.prefill(prompt)
    spec_logits = m.forward(flat.token_ids(), trie=flat.nodes())
    assert spec_logits.shape[0] == 7 and m.context_length == len(prompt)
    paths = {2: [7, 8, 9], 3: [7, 8, 12], 6: [7, 30, 31, 32]}
    for leaf, toks in paths.items():
        ref = OracleModel(path, max_context=64)
        ref.prefill(prompt)
        want = ref.forward(toks, output_rows=(0, len(toks)))
        idx, p = [], leaf
        parents = flat.parents()
        while p >= 0:
            idx.append(p); p = parents[p]
        idx.reverse()
        assert (spec_logits[idx] == want).all(), f"branch to node {leaf}"
    # accept the deepest branch: the cache must now equal a flat pass over that branch, so the next token's logits agree bit for bit
    m.accept([0, 4, 5, 6])
    ref = OracleModel(path, max_context=64)
    ref.prefill(prompt)
    ref.forward([7, 30, 31, 32])
    assert m.context_length == ref.context_length
    assert (m.forward([5]) == ref.forward([5])).all()


def test_oracle_flat_trie_equals_flat_pass(tmp_path):
    spec = synth.tiny("llama")
    path = synth.write_model(spec, tmp_path / "m", seed=42)
    toks = [3, 1, 4, 1, 5]
    a, b = OracleModel(path, max_context=64), OracleModel(path, max_context=64)
    for mdl in (a, b):
        mdl.prefill([9, 2, 6])
    flat = TrieNode.flat(3, toks, PRng(0)).linearize()
    la = a.forward(toks, trie=flat.nodes())
    lb = b.forward(toks, output_rows=(0, len(toks)))
    assert (la == lb).all()
    a.accept(range(len(toks)))
    assert (a.forward([8]) == b.forward([8])).all()


def test_oracle_rejects_speculation_on_hybrid_models(tmp_path):
    spec = synth.tiny("qwen-hybrid")
    path = synth.write_model(spec, tmp_path / "m", seed=43)
    m = OracleModel(path, max_context=64)
    m.prefill([1, 2])
    with pytest.raises(AssertionError):
        m.forward([3, 4], trie=TrieNode.flat(2, [3, 4], PRng(0)).linearize().nodes())
