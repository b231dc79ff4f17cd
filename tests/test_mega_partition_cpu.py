"""Host logic of the persistent decode kernel's work partition (uzu_b200/csrc/decode_mega.cu `mk_my_range` / `mk_range_of` and the
CTA-level combine of partial tiles, mirrored by `MegaBuilder::range_of` in engine.cu). The units of a GEMV phase are cut into
min(grid, U) equal CTA ranges, each CTA range into min(NCW, length) equal warp ranges; a tile summed entirely by one warp goes
straight to piece slot 0, the partial tiles of a CTA meet in shared memory (owner = first warp of the CTA that touches the tile) and
leave ONE piece in slot (cta - first cta of the tile). This restates the integer arithmetic and the ownership rules in Python and checks:
every unit is summed exactly once, every (tile, slot) is written at most once, slots are numbered 0..count-1 without gaps and
count <= the P the host allocates."""
import random


def range_of(u, U, W):
    return ((u + 1) * W - 1) // U


def simulate(tiles_list, C, grid, ncw):
    unit0, mats = 0, []
    for t in tiles_list:
        mats.append((unit0, t))
        unit0 += t * C
    U = unit0
    geff = min(grid, U)
    pieces = {}          # (mi, tile, slot) -> set of units
    per_cta_units = []
    for cta in range(geff):
        cb, ce = cta * U // geff, (cta + 1) * U // geff
        ln = ce - cb
        assert ln >= 1
        per_cta_units.append(ln)
        nw = min(ncw, ln)
        frags = [[] for _ in range(ncw)]     # per warp: list of (key, c_first, units)
        for w in range(nw):
            ub, ue = cb + w * ln // nw, cb + (w + 1) * ln // nw
            assert ue > ub
            u = ub
            while u < ue:
                mi = 1 if len(mats) > 1 and u >= mats[1][0] else 0
                m0, mt = mats[mi]
                tile, c = divmod(u - m0, C)
                c_first = c
                mend = min(ue, m0 + mt * C)
                cur = []
                while u < mend:
                    cur.append(u)
                    if c + 1 == C or u + 1 == mend:
                        if c_first == 0 and c + 1 == C:
                            key = (mi, tile, 0)
                            assert key not in pieces
                            pieces[key] = set(cur)
                        else:
                            frags[w].append(((mi, tile), c_first, set(cur)))
                        cur, c_first = [], 0
                    u += 1
                    c += 1
                    if c == C:
                        c, tile = 0, tile + 1
            assert len(frags[w]) <= 2
        for w in range(ncw):
            for (key, c_first, units) in frags[w]:
                if not (c_first == 0 or w == 0):
                    continue
                acc = set(units)
                w2 = w + 1
                while w2 < ncw and frags[w2] and frags[w2][0][0] == key:
                    assert not (acc & frags[w2][0][2])
                    acc |= frags[w2][0][2]
                    w2 += 1
                mi, tile = key
                m0 = mats[mi][0]
                slot = 0 if c_first == 0 else cta - range_of(m0 + tile * C, U, geff)
                k = (mi, tile, slot)
                assert k not in pieces, k
                pieces[k] = acc
    assert max(per_cta_units) - min(per_cta_units) <= 1, "every SM streams the same number of units (+-1)"
    # every unit exactly once, slots gap-free, count == host formula
    seen = set()
    for mi, (m0, mt) in enumerate(mats):
        for t in range(mt):
            cnt = range_of(m0 + (t + 1) * C - 1, U, geff) - range_of(m0 + t * C, U, geff) + 1
            got = set()
            for p in range(cnt):
                assert (mi, t, p) in pieces, (mi, t, p, cnt)
                assert not (got & pieces[(mi, t, p)])
                got |= pieces[(mi, t, p)]
            assert (mi, t, cnt) not in pieces
            assert got == set(range(m0 + t * C, m0 + (t + 1) * C))
            seen |= got
    assert len(seen) == U
    return max(range_of(m0 + (t + 1) * C - 1, U, geff) - range_of(m0 + t * C, U, geff) + 1 for (m0, mt) in mats for t in range(mt))


def test_baseline_shapes():
    assert simulate([384], 8, 148, 15) <= 2            # Llama-3-8B qkv 6144 x 4096
    assert simulate([256], 8, 148, 15) <= 2            # out 4096 x 4096
    assert simulate([1792], 8, 148, 15) <= 2           # up 28672 x 4096
    assert simulate([256], 28, 148, 15) <= 2           # down 4096 x 14336
    assert simulate([8016], 8, 148, 15) <= 2           # readout 128256 x 4096
    assert simulate([128, 384], 2, 148, 15) <= 2       # Qwen3.5 gate + qkv sharing one phase, K = 1024
    assert simulate([15520], 2, 148, 15) <= 2          # Qwen3.5 readout 248320 x 1024
    assert simulate([64], 7, 148, 15) <= 4             # Qwen3.5 down 1024 x 3584


def test_random_shapes():
    rnd = random.Random(7)
    for grid, ncw in ((148, 15), (148, 11), (148, 7), (132, 15), (7, 3), (33, 2)):
        for _ in range(40):
            C = rnd.choice([1, 2, 4, 7, 8, 28, 32])
            tiles = [rnd.randint(1, 1500)] + ([rnd.randint(1, 400)] if rnd.random() < 0.4 else [])
            simulate(tiles, C, grid, ncw)
