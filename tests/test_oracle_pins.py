"""Pins for the CPU oracle (no GPU). The reference stores no outputs for this path except the
`unit_interval` endpoints, so the oracle is pinned with: those endpoints, published Philox4x32-10
known answers, the reference tests' own equivalence properties evaluated on their closed-form inputs,
and independent float64 numpy restatements of the math (different code path, same definition)."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.util import bf16_to_f32, f32_to_bf16, assert_bf16_close


# ---- literal known answers ---------------------------------------------------------------------------
def test_unit_interval_endpoints():
    # tests/unit/encodable_block/sampling/gumbel_test.rs:7-12
    lib = O.lib()
    assert lib.oracle_unit_interval(0xFFFFFFFF) == np.float32(1.0) - np.float32(2.0) ** -24
    assert lib.oracle_unit_interval(0) == np.float32(2.0) ** -24
    assert lib.oracle_unit_interval(255) == np.float32(2.0) ** -24


PHILOX_KAT = [  # Random123 kat_vectors, philox4x32 with 10 rounds: counter, key -> output
    ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
    ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
    ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0], [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
]


@pytest.mark.parametrize("ctr,key,expected", PHILOX_KAT)
def test_philox_known_answers(ctr, key, expected):
    assert [int(x) for x in O.philox4x32_10(ctr, key)] == expected


def test_bf16_rounding_known_values():
    lib = O.lib()
    assert lib.oracle_f32_to_bf16(1.0) == 0x3F80
    # round to nearest even on exact ties
    assert lib.oracle_f32_to_bf16(np.array([0x3F808000], np.uint32).view(np.float32)[0]) == 0x3F80
    assert lib.oracle_f32_to_bf16(np.array([0x3F818000], np.uint32).view(np.float32)[0]) == 0x3F82
    assert lib.oracle_f32_to_bf16(np.array([0x3F808001], np.uint32).view(np.float32)[0]) == 0x3F81
    x = np.random.default_rng(0).standard_normal(10000).astype(np.float32) * 100
    c = np.array([lib.oracle_f32_to_bf16(float(v)) for v in x[:2000]], np.uint16)
    assert (c == f32_to_bf16(x[:2000])).all() and (c == O.f32_to_bf16(x[:2000])).all()


def test_prng_derive_is_murmur_fmix64():
    def fmix(h):
        m = (1 << 64) - 1
        h ^= h >> 33; h = (h * 0xFF51AFD7ED558CCD) & m
        h ^= h >> 33; h = (h * 0xC4CEB9FE1A85EC53) & m
        h ^= h >> 33
        return h
    lib = O.lib()
    for seed, idx in [(0, 0), (42, 7), (2**63, 12345), (0xDEADBEEF, 2**40)]:
        assert lib.oracle_prng_derive(seed, idx) == fmix((seed + idx) & ((1 << 64) - 1))
    assert lib.oracle_prng_derive(0, 0) == 0


# ---- matmul ----------------------------------------------------------------------------------------------
def dequant_numpy(packed, scales, zero_points, biases, *, n, k, bits, gs, method):
    """Independent float64 restatement of the weight layout (cpu/kernel/matmul/kernel.rs:236-277)."""
    if bits == 4:
        b = packed.reshape(n, k // 2)
        codes = np.empty((n, k), np.float64)
        codes[:, 0::2] = b & 0xF
        codes[:, 1::2] = b >> 4
    else:
        codes = packed.reshape(n, k).astype(np.float64)
    groups = -(-k // gs)
    s = bf16_to_f32(scales).astype(np.float64).reshape(n, groups)
    s_full = np.repeat(s, gs, axis=1)[:, :k]
    if method == O.QM_ZERO_POINT:
        if bits == 4:
            zb = zero_points.reshape(n, -(-groups // 2))
            zp = np.empty((n, 2 * zb.shape[1]), np.float64)
            zp[:, 0::2] = zb & 0xF
            zp[:, 1::2] = zb >> 4
            zp = zp[:, :groups]
        else:
            zp = zero_points.reshape(n, groups).astype(np.float64)
        return s_full * (codes - np.repeat(zp, gs, axis=1)[:, :k])
    if method == O.QM_SCALE_BIAS:
        bb = bf16_to_f32(biases).astype(np.float64).reshape(n, groups)
        return s_full * codes + np.repeat(bb, gs, axis=1)[:, :k]
    return s_full * (codes - (1 << (bits - 1)))


def random_quant(rng, n, k, bits, gs, method):
    groups = -(-k // gs)
    packed = rng.integers(0, 256, size=(n, k // 2 if bits == 4 else k), dtype=np.uint8)
    scales = f32_to_bf16(rng.uniform(0.01, 0.3, size=(n, groups)).astype(np.float32))   # tests/matmul/quant.rs:58-111
    zp = biases = None
    if method == O.QM_ZERO_POINT:
        zp = rng.integers(0, 256, size=(n, -(-groups // 2) if bits == 4 else groups), dtype=np.uint8)
    elif method == O.QM_SCALE_BIAS:
        biases = f32_to_bf16(rng.uniform(-0.03, 0.03, size=(n, groups)).astype(np.float32))
    return packed, scales, zp, biases


@pytest.mark.parametrize("bits,gs,method", [(4, 64, O.QM_ZERO_POINT), (4, 32, O.QM_SCALE_BIAS), (8, 64, O.QM_ZERO_POINT),
                                            (4, 128, O.QM_SYMMETRIC), (8, 32, O.QM_SCALE_BIAS), (8, 128, O.QM_SYMMETRIC)])
def test_matmul_matches_float64_dequant(bits, gs, method):
    rng = np.random.default_rng(1)
    m, n, k = 3, 37, 256
    packed, scales, zp, biases = random_quant(rng, n, k, bits, gs, method)
    x = f32_to_bf16(rng.uniform(-0.3, 0.3, size=(m, k)).astype(np.float32))
    d = O.matmul(x, packed, m=m, n=n, k=k, scales=scales, zero_points=zp, biases=biases, method=method, bits=bits,
                 group_size=gs, d_f32=True)
    w = dequant_numpy(packed, scales, zp, biases, n=n, k=k, bits=bits, gs=gs, method=method)
    ref = bf16_to_f32(x).astype(np.float64) @ w.T
    np.testing.assert_allclose(d, ref, rtol=2e-5, atol=2e-5)


def test_matmul_full_precision_closed_form():
    # tests/unit/backends/common/kernel/matmul/gemv_test.rs:38-58 input patterns
    for m, k, n in [(1, 33, 3), (4, 128, 11), (8, 128, 64)]:
        a = f32_to_bf16(((np.arange(m * k) % 13) * 0.1 - 0.6).astype(np.float32)).reshape(m, k)
        b = f32_to_bf16(((np.arange(n * k) % 17) * 0.1 - 0.8).astype(np.float32)).reshape(n, k)
        d = O.matmul(a, b, m=m, n=n, k=k, method=O.QM_NONE, d_f32=True)
        ref = bf16_to_f32(a).astype(np.float64) @ bf16_to_f32(b).astype(np.float64).T
        np.testing.assert_allclose(d, ref, rtol=1e-5, atol=1e-5)


def test_gather_readout_equals_dense():
    # gemv_test.rs:188-260: out[r][j] == dense[r][gather[r][j]], bit for bit on the CPU backend
    rng = np.random.default_rng(2)
    m, n_full, k, n_sel = 2, 64, 128, 9
    packed, scales, zp, _ = random_quant(rng, n_full, k, 4, 64, O.QM_ZERO_POINT)
    x = f32_to_bf16(rng.uniform(-0.3, 0.3, size=(m, k)).astype(np.float32))
    dense = O.matmul(x, packed, m=m, n=n_full, k=k, scales=scales, zero_points=zp, method=O.QM_ZERO_POINT, group_size=64)
    gather = rng.integers(0, n_full, size=(m, n_sel)).astype(np.uint32)
    sparse = O.matmul(x, packed, m=m, n=n_sel, k=k, scales=scales, zero_points=zp, method=O.QM_ZERO_POINT, group_size=64,
                      gather=gather)
    for r in range(m):
        assert (sparse[r] == dense[r][gather[r]]).all()


def test_matmul_epilogue_and_threads():
    rng = np.random.default_rng(3)
    m, n, k = 2, 40, 128
    packed, scales, zp, _ = random_quant(rng, n, k, 4, 64, O.QM_ZERO_POINT)
    x = f32_to_bf16(rng.uniform(-0.3, 0.3, size=(m, k)).astype(np.float32))
    bias = f32_to_bf16(rng.uniform(-1, 1, size=n).astype(np.float32))
    base = O.matmul(x, packed, m=m, n=n, k=k, scales=scales, zero_points=zp, method=O.QM_ZERO_POINT, d_f32=True)
    d0 = np.full((m, n), 0.5, np.float32)
    out = O.matmul(x, packed, m=m, n=n, k=k, scales=scales, zero_points=zp, method=O.QM_ZERO_POINT, d=d0.copy(),
                   ab_scale=2.0, accumulate=True, bias=bias, soft_cap=3.0)
    ref = 3.0 * np.tanh((np.float32(2.0) * base + np.float32(0.5) + bf16_to_f32(bias)[None, :]) / np.float32(3.0))
    np.testing.assert_allclose(out, ref, rtol=1e-6, atol=1e-6)
    # the OpenMP "courtesy" variant computes every element identically
    mt = O.matmul(x, packed, m=m, n=n, k=k, scales=scales, zero_points=zp, method=O.QM_ZERO_POINT, d_f32=True, threads=4)
    assert (mt == base).all()


def test_signed_codes_equals_flipped_bytes():
    # WeightMatrix::make_codes_signed XORs 0x88 / 0x80 into the stored bytes; signed_codes undoes it at read time
    rng = np.random.default_rng(4)
    m, n, k = 1, 16, 128
    for bits, mask in [(4, 0x88), (8, 0x80)]:
        packed, scales, zp, _ = random_quant(rng, n, k, bits, 64, O.QM_ZERO_POINT)
        x = f32_to_bf16(rng.uniform(-0.3, 0.3, size=(m, k)).astype(np.float32))
        a = O.matmul(x, packed, m=m, n=n, k=k, scales=scales, zero_points=zp, method=O.QM_ZERO_POINT, bits=bits)
        b = O.matmul(x, packed ^ np.uint8(mask), m=m, n=n, k=k, scales=scales, zero_points=zp, method=O.QM_ZERO_POINT,
                     bits=bits, signed_codes=True)
        assert (a == b).all()


# ---- attention ---------------------------------------------------------------------------------------------
def attention_inputs(num_heads, num_kv_heads, seq, suffix, D):
    # tests/unit/backends/common/kernel/attention/attention_single_pass_test.rs:33-78 (closed-form sin/cos inputs,
    # K/V laid out [kv_head, seq, D]: head stride = seq*D, seq stride = D)
    q = f32_to_bf16((np.sin(np.arange(num_heads * suffix * D, dtype=np.float32) * np.float32(0.13) + np.float32(0.5)) * np.float32(0.5)))
    k = f32_to_bf16((np.cos(np.arange(num_kv_heads * seq * D, dtype=np.float32) * np.float32(0.07) + np.float32(1.0)) * np.float32(0.5)))
    v = f32_to_bf16((np.sin(np.arange(num_kv_heads * seq * D, dtype=np.float32) * np.float32(0.11) + np.float32(2.0)) * np.float32(0.5)))
    return q, k, v


def softmax_reference(q, k, v, num_heads, num_kv_heads, seq, suffix, D, causal):
    """Independent float64 softmax attention (tests/unit/encodable_block/attention_test.rs:26-124 style)."""
    qf = bf16_to_f32(q).astype(np.float64).reshape(num_heads, suffix, D)
    kf = bf16_to_f32(k).astype(np.float64).reshape(num_kv_heads, seq, D)
    vf = bf16_to_f32(v).astype(np.float64).reshape(num_kv_heads, seq, D)
    g = num_heads // num_kv_heads
    out = np.zeros((suffix, num_heads, D))
    prefix = seq - suffix
    for h in range(num_heads):
        s = qf[h] @ kf[h // g].T / np.sqrt(D)
        if causal:
            for t in range(suffix):
                s[t, prefix + t + 1:] = -np.inf
        p = np.exp(s - s.max(axis=1, keepdims=True))
        p /= p.sum(axis=1, keepdims=True)
        out[:, h, :] = p @ vf[h // g]
    return out


@pytest.mark.parametrize("H,Hkv,seq,suffix,D,causal", [(4, 4, 16, 1, 64, False), (8, 2, 40, 4, 128, True), (4, 1, 64, 8, 64, True)])
def test_attention_single_and_two_pass_vs_softmax(H, Hkv, seq, suffix, D, causal):
    q, k, v = attention_inputs(H, Hkv, seq, suffix, D)
    kw = dict(head_dim=D, gqa_factor=H // Hkv, sequence_length=seq, k_head_stride=seq * D, k_seq_stride=D,
              v_head_stride=seq * D, v_seq_stride=D, scale=float(np.float32(1.0) / np.sqrt(np.float32(D))), num_heads=H,
              suffix_length=suffix, is_causal=causal)
    sp = O.attention_single_pass(q, k, v, **kw)
    tp = O.attention_two_pass(q, k, v, **kw)
    ref = softmax_reference(q, k, v, H, Hkv, seq, suffix, D, causal)
    np.testing.assert_allclose(bf16_to_f32(sp), ref, rtol=1e-2, atol=4e-3)   # bf16 output
    assert_bf16_close(tp, sp, max_ulp=1, min_exact=0.95, what="two-pass vs single-pass")


def test_attention_mask_ring_and_window():
    # mask.rs:3-62: sliding window on a ring: only ring_length filled slots, window relative to query position
    H, D, prefix, suffix = 1, 64, 8, 1
    rng = np.random.default_rng(5)
    q = f32_to_bf16(rng.standard_normal((H, suffix, D)).astype(np.float32))
    k = f32_to_bf16(rng.standard_normal((prefix + suffix, D)).astype(np.float32))
    v = f32_to_bf16(rng.standard_normal((prefix + suffix, D)).astype(np.float32))
    kw = dict(head_dim=D, gqa_factor=1, sequence_length=prefix + suffix, k_head_stride=D, k_seq_stride=D, v_head_stride=D,
              v_seq_stride=D, scale=0.125, num_heads=H, suffix_length=suffix, is_causal=True)
    full = O.attention_single_pass(q, k, v, **kw)
    # ring with offset 3, 5 filled entries, window 4: key positions (8 + i - 3) % 8 for the prefix, suffix position = 5
    ring = O.attention_single_pass(q, k, v, ring=(3, 5), sliding_window=4, **kw)
    pos = [(prefix + i - 3) % prefix for i in range(prefix)]
    used = [i for i in range(prefix) if pos[i] < 5 and pos[i] <= 5 and 5 - pos[i] < 4] + [prefix]
    qf = bf16_to_f32(q)[0, 0].astype(np.float64) * 0.125
    s = bf16_to_f32(k).astype(np.float64)[used] @ qf
    p = np.exp(s - s.max()); p /= p.sum()
    ref = p @ bf16_to_f32(v).astype(np.float64)[used]
    np.testing.assert_allclose(bf16_to_f32(ring)[0, 0], ref, rtol=1e-2, atol=4e-3)
    assert not (ring == full).all()


def test_attention_trie_mask_is_ancestor_or_self_visibility():
    """mask.rs:21-29: suffix key i is visible to query q iff trie_start(i) <= q <= trie_end(i), i.e. q lies in the subtree of node i -- key i is
    an ancestor of the query's node or the node itself; every prefix key is visible. Float64 softmax over exactly that set, for a trie
    with siblings at two depths, single-pass and two-pass cores."""
    from uzu_b200.trie import TrieNode
    root = TrieNode(0, 0)
    a = TrieNode(1, 0); a.add(TrieNode(2, 0)); a.add(TrieNode(3, 0))
    b = TrieNode(4, 0); c = TrieNode(5, 0); c.add(TrieNode(6, 0)); b.add(c)
    root.add(a); root.add(b); root.add(TrieNode(7, 0))
    flat = root.linearize()
    nodes, parents = flat.nodes(), flat.parents()
    H, Hkv, D, prefix, suffix = 4, 2, 64, 21, len(flat)
    seq = prefix + suffix
    q, k, v = attention_inputs(H, Hkv, seq, suffix, D)
    kw = dict(head_dim=D, gqa_factor=H // Hkv, sequence_length=seq, k_head_stride=seq * D, k_seq_stride=D, v_head_stride=seq * D,
              v_seq_stride=D, scale=float(np.float32(1.0) / np.sqrt(np.float32(D))), num_heads=H, suffix_length=suffix, is_causal=True)
    sp = bf16_to_f32(O.attention_single_pass(q, k, v, trie=nodes, **kw))
    tp = O.attention_two_pass(q, k, v, trie=nodes, **kw)
    qf = bf16_to_f32(q).astype(np.float64).reshape(H, suffix, D)
    kf = bf16_to_f32(k).astype(np.float64).reshape(Hkv, seq, D)
    vf = bf16_to_f32(v).astype(np.float64).reshape(Hkv, seq, D)
    for node in range(suffix):
        anc, pnode = [], node
        while pnode >= 0:
            anc.append(pnode); pnode = parents[pnode]
        visible = list(range(prefix)) + [prefix + i for i in sorted(anc)]
        assert sorted(anc) == [i for i in range(suffix) if nodes[i, 0] <= node <= nodes[i, 1]]
        for h in range(H):
            s_ = kf[h // (H // Hkv)][visible] @ qf[h, node] / np.sqrt(D)
            p_ = np.exp(s_ - s_.max()); p_ /= p_.sum()
            np.testing.assert_allclose(sp[node, h], p_ @ vf[h // (H // Hkv)][visible], rtol=1e-2, atol=4e-3)
    assert_bf16_close(tp, f32_to_bf16(sp), max_ulp=1, min_exact=0.9, what="two-pass vs single-pass with a trie")
    # a flat chain is the causal mask
    chain = TrieNode.flat(0, range(suffix), __import__("uzu_b200.trie", fromlist=["PRng"]).PRng(0)).linearize().nodes()
    assert (O.attention_single_pass(q, k, v, trie=chain, **kw) == O.attention_single_pass(q, k, v, **kw)).all()


def test_qkv_norm_and_logit_transform_closed_forms():
    """qkv_norm.rs:36-76: per-head RMS norm over head_dim on the selected heads only, in place; logit_transform: scale then soft cap
    cap * tanh(x / cap)."""
    rng = np.random.default_rng(41)
    rows, heads, D = 3, 6, 64
    qkv = f32_to_bf16(rng.standard_normal((rows, heads * D)).astype(np.float32))
    scales = (1.0 + 0.1 * rng.standard_normal(D)).astype(np.float32)
    got = O.qkv_norm(qkv.copy(), scales, total_heads=heads, head_dim=D, epsilon=1e-6, scale_offset=0.0, head_offset=1, head_count=3, full_layer=False)
    x = bf16_to_f32(qkv).astype(np.float64).reshape(rows, heads, D)
    want = x.copy()
    sel = x[:, 1:4]
    want[:, 1:4] = sel / np.sqrt((sel ** 2).mean(axis=2, keepdims=True) + 1e-6) * scales
    g = bf16_to_f32(got).reshape(rows, heads, D)
    np.testing.assert_allclose(g[:, 1:4], want[:, 1:4], rtol=2e-2, atol=2e-3)
    assert (got.reshape(rows, heads, D)[:, [0, 4, 5]] == qkv.reshape(rows, heads, D)[:, [0, 4, 5]]).all()   # untouched heads: bit-identical
    logits = f32_to_bf16((rng.standard_normal(500) * 40).astype(np.float32)).reshape(1, 500)
    out = O.logit_transform(logits.copy(), 0.5, 30.0)
    ref = 30.0 * np.tanh(bf16_to_f32(logits).astype(np.float64) * 0.5 / 30.0)
    np.testing.assert_allclose(bf16_to_f32(out), ref, rtol=1e-2, atol=1e-2)


# ---- normalization / rope / prepare ------------------------------------------------------------------------
def test_normalization_closed_form():
    # tests/unit/backends/common/kernel/normalization_test.rs:87-92 input pattern 0.5 + i*0.01
    n, rows = 256, 3
    x = f32_to_bf16((0.5 + np.arange(rows * n, dtype=np.float32) * np.float32(0.01)).reshape(rows, n))
    scales = (1.0 + 0.001 * np.arange(n)).astype(np.float32)
    for full_layer in (False, True):
        out = O.normalization(x, scales, epsilon=1e-5, full_layer=full_layer)
        xf = bf16_to_f32(x).astype(np.float64)
        ref = xf / np.sqrt((xf ** 2).mean(axis=1, keepdims=True) + 1e-5) * scales
        np.testing.assert_allclose(bf16_to_f32(out), ref, rtol=1.2e-2, atol=1e-3)
    # residual add: shortcut <- bf16(x + shortcut) and the norm reads the updated shortcut
    sc = f32_to_bf16(np.full((rows, n), 0.25, np.float32))
    sc0 = sc.copy()
    out = O.normalization(x, scales, shortcut=sc, residual_add=True, epsilon=1e-5)
    assert (sc == f32_to_bf16(bf16_to_f32(x) + bf16_to_f32(sc0))).all()
    out2 = O.normalization(sc.copy(), scales, epsilon=1e-5)
    assert (out == out2).all()


def test_rope_tables_and_prepare():
    cfg = {"type": "LlamaRoPEConfig", "base": 500000.0, "head_dim": 128, "scaling_factor": 8.0,
           "original_context_length": 8192, "low_frequency_factor": 1.0, "high_frequency_factor": 4.0}
    pos = np.array([0, 1, 17, 4095], np.uint32)
    cos, sin = O.rope_tables(cfg, pos)
    inv = 1.0 / np.float64(500000.0) ** (np.arange(0, 128, 2) / 128.0)
    wl = 2 * np.pi / inv
    scaled = inv / 8.0
    smooth = (8192 / wl - 1.0) / 3.0
    f = np.where(wl < 8192 / 4.0, inv, np.where(wl > 8192.0, scaled, smooth * inv + (1 - smooth) * scaled))
    ang = pos[:, None].astype(np.float64) * f[None, :]
    np.testing.assert_allclose(cos[:, :64], np.cos(ang), atol=2e-3)
    np.testing.assert_allclose(sin[:, 64:], np.sin(ang), atol=2e-3)
    assert (cos[0] == 1).all() and (sin[0] == 0).all()
    # prepare: Q transposed to [Hq, m, D], K/V appended at kv_token_offset, half-rotation pairing
    rng = np.random.default_rng(6)
    Hq, Hkv, D, m = 4, 2, 128, 3
    qkv = f32_to_bf16(rng.standard_normal((m, (Hq + 2 * Hkv) * D)).astype(np.float32))
    keys = np.zeros((16, Hkv * D), np.uint16); values = np.zeros((16, Hkv * D), np.uint16)
    cos, sin = O.rope_tables(cfg, np.arange(5, 5 + m))
    queries = O.attention_prepare(qkv, keys, values, cos, sin, num_q_heads=Hq, num_kv_heads=Hkv, head_dim=D, rope_dim=128,
                                  kv_token_offset=5)
    x = bf16_to_f32(qkv).reshape(m, Hq + 2 * Hkv, D)
    rot = np.concatenate([-x[..., 64:], x[..., :64]], axis=-1)
    ref = x * cos[:, None, :] + rot * sin[:, None, :]
    np.testing.assert_allclose(bf16_to_f32(queries), ref[:, :Hq].transpose(1, 0, 2), rtol=1e-2, atol=1e-2)
    np.testing.assert_allclose(bf16_to_f32(keys[5:8]).reshape(m, Hkv, D), ref[:, Hq:Hq + Hkv], rtol=1e-2, atol=1e-2)
    assert (values[5:8].reshape(m, Hkv, D) == qkv.reshape(m, Hq + 2 * Hkv, D)[:, Hq + Hkv:]).all()   # V is copied, not rotated
    assert (keys[:5] == 0).all() and (keys[8:] == 0).all()


# ---- gate / embedding / sampling -----------------------------------------------------------------------------
def test_gated_act_mul_two_roundings():
    rng = np.random.default_rng(7)
    F = 96
    up = f32_to_bf16(rng.standard_normal((2, 2 * F)).astype(np.float32) * 3)
    out = O.gated_act_mul(up, F)
    v, g = bf16_to_f32(up[:, :F]), bf16_to_f32(up[:, F:])
    act = bf16_to_f32(f32_to_bf16((g / (1 + np.exp(-g.astype(np.float64)))).astype(np.float32)))
    ref = f32_to_bf16(v * act)
    assert_bf16_close(out, ref, max_ulp=1, min_exact=0.98, what="gated_act_mul")


def test_quant_embedding_lookup_is_a_dequantised_row():
    rng = np.random.default_rng(8)
    V, H = 50, 128
    for bits, method in [(4, O.QM_ZERO_POINT), (8, O.QM_SCALE_BIAS), (4, O.QM_SYMMETRIC)]:
        packed, scales, zp, biases = random_quant(rng, V, H, bits, 64, method)
        toks = np.array([3, 49, 50, 0], np.uint32)   # 50 is out of range -> zeros (quant_embedding.rs:51-54)
        out = O.quant_embedding_lookup(toks, packed, scales, zero_points=zp, biases=biases, vocab_size=V, model_dim=H,
                                       input_scale=2.0, group_size=64, mode=O.MODE_U4 if bits == 4 else O.MODE_U8, method=method)
        w = dequant_numpy(packed, scales, zp, biases, n=V, k=H, bits=bits, gs=64, method=method)
        np.testing.assert_allclose(bf16_to_f32(out[[0, 1, 3]]), 2.0 * w[[3, 49, 0]], rtol=1e-2, atol=1e-3)
        assert (out[2] == 0).all()


def test_sampling_greedy_ties_and_gumbel():
    V = 5000
    rng = np.random.default_rng(9)
    logits = f32_to_bf16(rng.standard_normal((3, V)).astype(np.float32))
    logits[0, 100] = logits[0, 4000] = f32_to_bf16(np.array([9.0], np.float32))[0]   # exact tie -> lowest index
    assert list(O.unified_sampling(logits)) == [100, int(np.argmax(bf16_to_f32(logits[1]))), int(np.argmax(bf16_to_f32(logits[2])))]
    # top_k = 1 keeps only the maximum whatever the noise
    seeds = np.array([1, 2, 3], np.uint64)
    assert list(O.unified_sampling(logits, seeds=seeds, top_k=1)) == list(O.unified_sampling(logits))
    # Gumbel-max == argmax(l/T + g) with g from the per-logit Philox stream
    lib = O.lib()
    T = 0.7
    got = O.unified_sampling(logits, seeds=seeds, temperature=T)
    for r in range(3):
        g = np.array([lib.oracle_gumbel_float(int(seeds[r]), *O.revidx(i, V)) for i in range(V)], np.float32)
        l = bf16_to_f32(logits[r]) * (np.float32(1.0) / np.float32(T))
        assert got[r] == int(np.argmax(l + g))
    # a bitmask removes tokens
    bm = np.full((3, (V + 31) // 32), 0xFFFFFFFF, np.uint32)
    best = int(O.unified_sampling(logits[1:2])[0])
    bm[1, best // 32] &= ~np.uint32(1 << (best % 32))
    assert int(O.unified_sampling(logits, bitmask=bm)[1]) != best


def test_sampling_filters_match_sorted_definition():
    V = 300
    rng = np.random.default_rng(10)
    logits = f32_to_bf16((rng.standard_normal((4, V)) * 2).astype(np.float32))
    seeds = np.arange(4, dtype=np.uint64) + 11
    lib = O.lib()
    for kw in [dict(top_k=5), dict(top_p=0.6), dict(min_p=0.1), dict(top_k=20, top_p=0.9, min_p=0.01, temperature=1.3)]:
        got = O.unified_sampling(logits, seeds=seeds, **kw)
        for r in range(4):
            l = bf16_to_f32(logits[r]).astype(np.float32)
            if "temperature" in kw:
                l = l * (np.float32(1.0) / np.float32(kw["temperature"]))
            order = sorted(range(V), key=lambda i: (-l[i], i))
            p = np.exp(l - l[order[0]]); p = p / p.sum()
            keep, mass = [], 0.0
            for rank, i in enumerate(order):
                if ("top_k" in kw and rank >= kw["top_k"]) or ("top_p" in kw and mass >= kw["top_p"]) or \
                        ("min_p" in kw and l[i] < l[order[0]] + np.log(kw["min_p"])):
                    break
                keep.append(i); mass += p[i]
            assert int(got[r]) in keep
            g = {i: lib.oracle_gumbel_float(int(seeds[r]), *O.revidx(i, V)) for i in keep}
            assert int(got[r]) == max(keep, key=lambda i: (np.float32(l[i]) + np.float32(g[i]), -i))


def test_delta_net_update_vs_float64():
    rng = np.random.default_rng(11)
    Hv, Hk, Dk, Dv = 4, 2, 128, 128
    kd, vd = Hk * Dk, Hv * Dv
    total = 2 * kd + vd + vd + 2 * Hv
    x = f32_to_bf16(rng.standard_normal(total).astype(np.float32))
    a_log = rng.uniform(-1, 1, Hv).astype(np.float32); dt = rng.uniform(-1, 1, Hv).astype(np.float32)
    nw = (1 + 0.1 * rng.standard_normal(Dv)).astype(np.float32)
    state = (rng.standard_normal((Hv, Dv, Dk)) * 0.1).astype(np.float32)
    s0 = state.astype(np.float64)
    out = O.delta_net_update(x, a_log, dt, nw, state, num_v_heads=Hv, num_k_heads=Hk, head_k_dim=Dk, head_v_dim=Dv, key_dim=kd,
                             value_dim=vd, norm_epsilon=1e-6)
    xf = bf16_to_f32(x).astype(np.float64)
    ref = np.zeros(vd)
    for hv in range(Hv):
        hk = hv // (Hv // Hk)
        q = xf[hk * Dk:(hk + 1) * Dk]; k = xf[kd + hk * Dk: kd + (hk + 1) * Dk]
        q = q / np.sqrt((q * q).sum() + 1e-6) / np.sqrt(Dk); k = k / np.sqrt((k * k).sum() + 1e-6)
        beta = 1 / (1 + np.exp(-xf[2 * kd + 2 * vd + hv]))
        sp = np.log1p(np.exp(xf[2 * kd + 2 * vd + Hv + hv] + dt[hv]))
        decay = np.exp(-np.exp(a_log[hv]) * sp)
        v = xf[2 * kd + hv * Dv: 2 * kd + (hv + 1) * Dv]
        S = s0[hv]
        delta = beta * (v - decay * (S @ k))
        o = decay * (S @ q) + delta * (k @ q)
        S_new = decay * S + np.outer(delta, k)
        np.testing.assert_allclose(state[hv], S_new, rtol=1e-4, atol=1e-5)
        z = xf[2 * kd + vd + hv * Dv: 2 * kd + vd + (hv + 1) * Dv]
        ref[hv * Dv:(hv + 1) * Dv] = o / np.sqrt((o * o).mean() + 1e-6) * nw * (z / (1 + np.exp(-z)))
    np.testing.assert_allclose(bf16_to_f32(out), ref, rtol=1.5e-2, atol=2e-3)


# ---- round 2: float64 twins for the remaining thin spots (conv update, two-pass block merge, sigmoid gate, KV-cache update) ----

def test_delta_net_conv_update_vs_float64():
    """gdn/conv_update.rs:8-55: out = silu(bias + sum_t state[t] * w[t] + x * w[last]) rounded to bf16; state shifts left and takes x."""
    rng = np.random.default_rng(21)
    for K, C_, has_bias in [(4, 96, True), (2, 40, False), (7, 17, True)]:
        w = rng.standard_normal((C_, K)).astype(np.float32) * 0.5
        bias = rng.standard_normal(C_).astype(np.float32) * 0.1 if has_bias else None
        x = f32_to_bf16(rng.standard_normal(C_).astype(np.float32))
        st = rng.standard_normal((C_, K - 1)).astype(np.float32)
        st0, xf = st.astype(np.float64), bf16_to_f32(x).astype(np.float64)
        io = x.copy()
        O.delta_net_conv_update(w, bias if has_bias else np.zeros(C_, np.float32), io, st, K, C_) if has_bias else \
            O.delta_net_conv_update(w, None, io, st, K, C_)
        acc = (bias.astype(np.float64) if has_bias else 0.0) + (st0 * w[:, :K - 1]).sum(1) + xf * w[:, K - 1]
        ref = acc / (1 + np.exp(-acc))
        np.testing.assert_allclose(bf16_to_f32(io), ref, rtol=1e-2, atol=2e-3)          # bf16 output
        np.testing.assert_array_equal(st[:, :K - 2], st0[:, 1:].astype(np.float32))       # shifted left, exactly
        np.testing.assert_array_equal(st[:, K - 2], bf16_to_f32(x))                       # newest input, exactly


def test_two_pass_partials_merge_like_a_float64_softmax():
    """attention_two_pass.rs:55-189: pass 1 leaves per-block (max, sum, unnormalised output); pass 2 merges 32 blocks. The merged
    result must be the float64 softmax over all keys, and each block's statistics must be the float64 statistics of ITS keys
    (block of key i = i % 32)."""
    H, Hkv, seq, suffix, D = 4, 2, 200, 1, 64
    q, k, v = attention_inputs(H, Hkv, seq, suffix, D)
    kw = dict(head_dim=D, gqa_factor=H // Hkv, sequence_length=seq, k_head_stride=seq * D, k_seq_stride=D, v_head_stride=seq * D,
              v_seq_stride=D, scale=float(np.float32(1.0) / np.sqrt(np.float32(D))), num_heads=H, suffix_length=suffix, is_causal=True)
    out, partials, sums, maxs = O.attention_two_pass(q, k, v, return_partials=True, **kw)
    qf = bf16_to_f32(q).astype(np.float64).reshape(H, suffix, D)
    kf = bf16_to_f32(k).astype(np.float64).reshape(Hkv, seq, D)
    vf = bf16_to_f32(v).astype(np.float64).reshape(Hkv, seq, D)
    for h in range(H):
        s = (kf[h // (H // Hkv)] @ (qf[h, 0] * kw["scale"]))
        for blk in range(32):
            idx = np.arange(blk, seq, 32)
            m = s[idx].max()
            np.testing.assert_allclose(maxs[0, h, blk], m, rtol=1e-5, atol=1e-5)
            p = np.exp(s[idx] - m)
            np.testing.assert_allclose(sums[0, h, blk], p.sum(), rtol=1e-4)
            np.testing.assert_allclose(partials[0, h, blk], p @ vf[h // (H // Hkv)][idx], rtol=1e-3, atol=1e-4)
        p = np.exp(s - s.max()); p /= p.sum()
        np.testing.assert_allclose(bf16_to_f32(out[0, h]), p @ vf[h // (H // Hkv)], rtol=1e-2, atol=2e-3)


def test_sigmoid_gate_and_kv_cache_update_closed_forms():
    rng = np.random.default_rng(22)
    gate = f32_to_bf16(rng.standard_normal(300).astype(np.float32) * 3)
    val = f32_to_bf16(rng.standard_normal(300).astype(np.float32))
    got = O.sigmoid_gate(gate, val.copy())
    ref = bf16_to_f32(val).astype(np.float64) / (1 + np.exp(-bf16_to_f32(gate).astype(np.float64)))
    np.testing.assert_allclose(bf16_to_f32(got), ref, rtol=1e-2, atol=1e-3)
    # kv_cache_update.rs:7-28: copies are applied in order, a later copy sees the rows written by an earlier one
    E = 16
    keys = f32_to_bf16(rng.standard_normal((10, E)).astype(np.float32)); values = f32_to_bf16(rng.standard_normal((10, E)).astype(np.float32))
    copies = [(7, 2), (2, 5), (9, 7)]
    ek, ev = keys.copy(), values.copy()
    for s_, d_ in copies:
        ek[d_] = ek[s_]; ev[d_] = ev[s_]
    O.kv_cache_update(keys, values, copies, E)
    assert (keys == ek).all() and (values == ev).all()


# ---- Mirai RHT: ActivationTransform + the matmul's output-RHT epilogue ------------------------------------------

def _rht_case(batch, channels):
    # the reference's own test inputs (tests/unit/backends/common/kernel/activation_transform_test.rs:64-76)
    data = (np.sin(np.arange(batch * channels, dtype=np.float64) * 0.1) * 2.0).reshape(batch, channels)
    factors = np.where(np.arange(channels) % 3 == 0, -1, 1).astype(np.int32)
    return data, factors


def _hadamard_f64(n=32):
    h = np.array([[1.0]])
    while h.shape[0] < n:
        h = np.block([[h, h], [h, -h]])
    return h / np.sqrt(n)


@pytest.mark.parametrize("batch,channels", [(1, 32), (1, 64), (1, 128), (4, 32), (4, 256), (2, 2048)])
def test_activation_transform_against_float64_hadamard(batch, channels):
    """activation_transform.rs:73-99 + mod.rs:31-44: the butterfly network IS the Sylvester-ordered Walsh-Hadamard matrix / sqrt(32)
    applied per 32-wide stripe; InputRht = H (s o x), OutputRht = s o (H x). Shapes and inputs of the reference's test."""
    data, factors = _rht_case(batch, channels)
    H = _hadamard_f64()
    blocks = data.reshape(batch, channels // 32, 32)
    fb = factors.reshape(channels // 32, 32).astype(np.float64)
    want_in = np.einsum("ij,bsj->bsi", H, blocks * fb).reshape(batch, channels)
    want_out = (np.einsum("ij,bsj->bsi", H, blocks) * fb).reshape(batch, channels)
    x32 = data.astype(np.float32)
    for op, want in ((O.RHT_INPUT, want_in), (O.RHT_OUTPUT, want_out)):
        got = O.activation_transform(x32, factors, op=op)
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5)             # reference tolerance for f32: 1e-4
        inplace = x32.copy()
        assert O.activation_transform(inplace, factors, op=op, in_place=True) is inplace and (inplace == got).all()
        xb = f32_to_bf16(x32)
        gb = O.activation_transform(xb, factors, op=op)
        # bf16: exact f32 transform of the bf16 inputs, rounded once
        blocks_b = bf16_to_f32(xb).astype(np.float64).reshape(batch, channels // 32, 32)
        wb = (np.einsum("ij,bsj->bsi", H, blocks_b * fb) if op == O.RHT_INPUT else np.einsum("ij,bsj->bsi", H, blocks_b) * fb).reshape(batch, channels)
        np.testing.assert_allclose(bf16_to_f32(gb), wb, rtol=2 ** -8, atol=1e-6)
    # H is orthogonal and symmetric: OutputRht undoes InputRht
    back = O.activation_transform(O.activation_transform(x32, factors, op=O.RHT_INPUT), factors, op=O.RHT_OUTPUT)
    np.testing.assert_allclose(back, x32, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("group,sum_group", [(32, None), (64, 32), (128, 128)])
def test_activation_quantize_closed_form(group, sum_group):
    """activation_transform.rs:11-42, mod.rs:9-29: per activation group, divisor = max|t| / 127 (1 if the group is all zero), code =
    round-half-away(t / divisor) clamped to +-127, group sums = integer sums of the codes."""
    rng = np.random.default_rng(0x5EED)
    rows, cols = 3, 256
    x = rng.uniform(-1, 1, (rows, cols)).astype(np.float32)
    x[1, 64:128] = 0.0
    factors = np.where(np.arange(cols) % 3 == 0, -1, 1).astype(np.int32)
    t = O.activation_transform(x, factors, op=O.RHT_INPUT)
    op = O.RHT_QUANTIZE if sum_group is None else O.RHT_QUANTIZE_WITH_GROUP_SUMS
    q, sc, gs = O.activation_transform(x, factors, op=op, activation_group_size=group, sum_group_size=sum_group or 0)
    tg = t.reshape(rows, cols // group, group)
    mag = np.abs(tg).max(axis=2)
    want_sc = np.where(mag > 0, mag / np.float32(127.0), np.float32(1.0)).astype(np.float32)
    assert (sc == want_sc).all()
    r = tg / want_sc[:, :, None]
    want_q = np.clip(np.sign(r) * np.floor(np.abs(r) + np.float32(0.5)), -127, 127).astype(np.int8).reshape(rows, cols)
    assert (q == want_q).all()
    assert np.abs(q.reshape(rows, cols // group, group)).max(axis=2)[mag > 0].min() == 127
    if sum_group is not None:
        assert (gs == q.astype(np.int32).reshape(rows, cols // sum_group, sum_group).sum(axis=2)).all()
    else:
        assert gs is None


def test_matmul_output_rht_runs_bias_after_the_transform():
    """cpu/kernel/matmul/kernel.rs:64,162,285,297-303: with rht_factors the in-loop bias is skipped, D is output-transformed in place
    (rounded to bf16 first, as stored), and the bias is added afterwards by TensorAddBias."""
    rng = np.random.default_rng(31)
    m, n, k = 3, 64, 128
    w = rng.integers(0, 256, (n, k // 2), dtype=np.uint8)
    scales = f32_to_bf16(rng.uniform(0.01, 0.3, (n, k // 64)).astype(np.float32))
    zp = rng.integers(0, 256, (n, 1), dtype=np.uint8)
    x = f32_to_bf16(rng.uniform(-0.3, 0.3, (m, k)).astype(np.float32))
    bias = f32_to_bf16(rng.uniform(-0.03, 0.03, n).astype(np.float32))
    factors = rng.choice(np.array([-1, 1], np.int32), n)
    kw = dict(m=m, n=n, k=k, scales=scales, zero_points=zp, method=O.QM_ZERO_POINT, bits=4, group_size=64)
    plain = O.matmul(x, w, **kw)
    got = O.matmul(x, w, bias=bias, rht_factors=factors, **kw)
    H = _hadamard_f64()
    t = (np.einsum("ij,bsj->bsi", H, bf16_to_f32(plain).astype(np.float64).reshape(m, n // 32, 32)) * factors.reshape(n // 32, 32)).reshape(m, n)
    t = bf16_to_f32(f32_to_bf16(t.astype(np.float32)))
    want = t.astype(np.float64) + bf16_to_f32(bias)
    np.testing.assert_allclose(bf16_to_f32(got), want, rtol=2 ** -7, atol=2e-3)
    # and it is NOT the bias-before-transform order
    wrong = O.activation_transform(O.matmul(x, w, bias=bias, **kw), factors, op=O.RHT_OUTPUT)
    assert (wrong != got).any()


def test_tensor_glue_fp_embedding_and_bitmask_sampling_closed_forms():
    """The small kernels around the hot path: tensor_{add_scale,add_bias,add_swap,copy}/*.rs (one bf16 rounding of the f32 result),
    full_precision_embedding.rs (row copy x input_scale, out-of-range id -> zeros), unified_sampling.rs:48-55 (a cleared grammar bit makes
    the logit -inf before anything else)."""
    rng = np.random.default_rng(51)
    rows, n = 3, 40
    x = f32_to_bf16(rng.standard_normal((rows, n)).astype(np.float32))
    b = f32_to_bf16(rng.standard_normal(n).astype(np.float32))
    xf, bf = bf16_to_f32(x).astype(np.float64), bf16_to_f32(b).astype(np.float64)
    assert (O.tensor_add_bias(x, b, n) == f32_to_bf16((xf + bf).astype(np.float32))).all()
    assert (O.tensor_add_scale(x, b, n, 0.25) == f32_to_bf16(((xf + bf) * 0.25).astype(np.float32))).all()
    skip, main = x.copy(), f32_to_bf16(rng.standard_normal((rows, n)).astype(np.float32))
    want = f32_to_bf16((bf16_to_f32(skip) + bf16_to_f32(main)).astype(np.float32))
    O.tensor_add_swap(skip, main)
    assert (skip == want).all() and (main == want).all()
    V, H = 9, 16
    w = f32_to_bf16(rng.standard_normal((V, H)).astype(np.float32))
    out = O.fp_embedding_lookup(np.array([8, 0, 9], np.uint32), w, vocab_size=V, model_dim=H, input_scale=1.5)
    np.testing.assert_allclose(bf16_to_f32(out[:2]), 1.5 * bf16_to_f32(w[[8, 0]]), rtol=2 ** -8)
    assert (out[2] == 0).all()
    # bitmask: only allowed ids can win, greedy and stochastic alike
    Vv = 100
    logits = f32_to_bf16(rng.standard_normal((2, Vv)).astype(np.float32))
    allowed = [[3, 64, 65], [99]]
    mask = np.zeros((2, (Vv + 31) // 32), np.uint32)
    for r, ids in enumerate(allowed):
        for i in ids:
            mask[r, i // 32] |= np.uint32(1 << (i % 32))
    g = O.unified_sampling(logits, bitmask=mask)
    f = bf16_to_f32(logits)
    assert int(g[0]) == max(allowed[0], key=lambda i: (f[0, i], -i)) and int(g[1]) == 99
    s = O.unified_sampling(logits, bitmask=mask, seeds=np.array([7, 8], np.uint64), temperature=1.3)
    assert int(s[0]) in allowed[0] and int(s[1]) == 99
