"""Host logic of the persistent decode kernel's work partition (uzu_b200/csrc/decode_mega.cu `mk_my_range` / `mk_range_of`, mirrored by
`MegaBuilder::range_of` in engine.cu): the units of a GEMV phase are cut into min(W, U) contiguous ranges, every (tile, range)
intersection is one "piece", and the pieces of a tile must be numbered 0..count-1 without gaps -- consumers sum exactly `count` slots.
This restates the integer arithmetic in Python and checks the invariants for the BASELINE shapes and for random ones."""
import random


def range_of(u, U, W):
    return ((u + 1) * W - 1) // U


def simulate(tiles_list, C, W):
    unit0, mats = 0, []
    for t in tiles_list:
        mats.append((unit0, t))
        unit0 += t * C
    U = unit0
    weff = min(W, U)
    written = set()
    sizes = []
    for ri in range(weff):
        ub, ue = ri * U // weff, (ri + 1) * U // weff
        assert ue > ub, "every range holds at least one unit"
        sizes.append(ue - ub)
        u = ub
        while u < ue:
            mi = 1 if len(mats) > 1 and u >= mats[1][0] else 0
            m0, mt = mats[mi]
            tile, c = divmod(u - m0, C)
            mend = min(ue, m0 + mt * C)
            while u < mend:
                if c + 1 == C or u + 1 == mend:
                    key = (mi, tile, ri - range_of(m0 + tile * C, U, weff))
                    assert key not in written
                    written.add(key)
                u += 1
                c += 1
                if c == C:
                    c, tile = 0, tile + 1
    total = 0
    for mi, (m0, mt) in enumerate(mats):
        for t in range(mt):
            cnt = range_of(m0 + (t + 1) * C - 1, U, weff) - range_of(m0 + t * C, U, weff) + 1
            assert all((mi, t, p) in written for p in range(cnt)) and (mi, t, cnt) not in written
            total += cnt
    assert total == len(written)
    assert max(sizes) - min(sizes) <= 1, "ranges are balanced to one unit"


def test_baseline_shapes():
    W = 148 * 15
    simulate([384], 8, W)            # Llama-3-8B qkv 6144 x 4096
    simulate([256], 8, W)            # out 4096 x 4096
    simulate([1792], 8, W)           # up 28672 x 4096
    simulate([256], 28, W)           # down 4096 x 14336
    simulate([8016], 8, W)           # readout 128256 x 4096
    simulate([128, 384], 2, W)       # Qwen3.5 gate + qkv sharing one phase, K = 1024
    simulate([15520], 2, W)          # Qwen3.5 readout 248320 x 1024


def test_random_shapes():
    rnd = random.Random(7)
    for W in (148 * 15, 148 * 11, 148 * 7, 132 * 15, 7, 33):
        for _ in range(40):
            C = rnd.choice([1, 2, 4, 7, 8, 28, 32])
            tiles = [rnd.randint(1, 1500)] + ([rnd.randint(1, 400)] if rnd.random() < 0.4 else [])
            simulate(tiles, C, W)
