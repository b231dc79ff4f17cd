"""Lane-level model of the rows kernel's data movement (uzu_b200/csrc/matmul.cu, qmv_rows_kernel), checked on the CPU by enumeration:
the 16-byte cp.async layout, the quad exchange that separates the two 64-nibble halves of a chunk by word index, the activation item
each word is multiplied with, the mma.sync.m16n8k16 fragment ownership (A rows g / g+8, B column g = activation row g, D columns 2t /
2t+1), the per-group affine with the zero-point word / shift selection, and the fragment -> row-major transposition of the epilogue.
The model is the kernel's index arithmetic restated in numpy; it must reproduce sum_k x[k] * scale[g(k)] * (code[k] - zp[g(k)]) for every
(weight row, activation row) of a 16 x 512-nibble super-chunk. (The GPU tests compare the kernel itself with the oracle; this test pins
the DESIGN of the layout independently of the hardware.)"""
import numpy as np
import pytest


def lane_words(chunk_bytes, t):
    """16 contiguous bytes [16t, 16t+16) of a 64-byte row chunk as four little-endian u32 words (issue_stage, FEED 1)."""
    return [int.from_bytes(bytes(chunk_bytes[16 * t + 4 * w:16 * t + 4 * w + 4]), "little") for w in range(4)]


def exchange(words_by_t):
    """The quad exchange of the consumer: lanes t < 2 keep words 0-1 and take words 0-1 of lane t + 2 into slots 2-3; lanes t >= 2 keep words
    2-3 and take words 2-3 of lane t - 2 into slots 0-1 (__shfl_xor(..., 2) of the half each lane gives away)."""
    out = {}
    for t in range(4):
        own, partner = words_by_t[t], words_by_t[t ^ 2]
        lo = t < 2
        send_partner = (partner[0], partner[1]) if lo else (partner[2], partner[3])     # what the partner (other side) sends us
        out[t] = [own[0], own[1], send_partner[0], send_partner[1]] if lo else [send_partner[0], send_partner[1], own[2], own[3]]
    return out


def item_index(t, w):
    """activation item (8 consecutive nibble positions) of chunk-local index for lane t, word slot w: xi + (w >> 1) * 8 + (w & 1)"""
    return 4 * (t & 1) + 2 * (t >> 1) + (w >> 1) * 8 + (w & 1)


def nib_pair(word, shift):
    """(code_i, code_{i+4}) with i = shift / 4: the two halves of the bf16x2 register"""
    return (word >> shift) & 15, (word >> (shift + 16)) & 15


@pytest.mark.parametrize("npg", [64, 128])
def test_exchange_puts_every_nibble_next_to_its_activation_and_group(npg):
    rng = np.random.default_rng(npg)
    chunk = rng.integers(0, 256, 64, dtype=np.uint8)
    codes = np.empty(128, np.int64)
    codes[0::2] = chunk & 15
    codes[1::2] = chunk >> 4                       # element c*K+k at bit (idx % 8) * 4 of its u32 word = low nibble first
    words = exchange({t: lane_words(chunk, t) for t in range(4)})
    seen = np.zeros(128, bool)
    for t in range(4):
        for w in range(4):
            li = item_index(t, w)
            for shift in (0, 4, 8, 12):
                i = shift // 4
                c_lo, c_hi = nib_pair(words[t][w], shift)
                for code, e in ((c_lo, i), (c_hi, i + 4)):
                    pos = li * 8 + e               # the activation element this code is multiplied with (xs item li, element e)
                    assert codes[pos] == code, (t, w, shift)
                    assert not seen[pos]
                    seen[pos] = True
                    if npg == 64:                  # words 0-1 accumulate the chunk's first group, words 2-3 its second
                        assert pos // 64 == (w >> 1)
    assert seen.all()


def test_bank_layout_of_the_fragment_loads():
    """xs rows are xstride = slice_items + 1 items of 16 bytes apart (odd): the 8 lanes of a quarter warp (two activation rows x four t)
    must touch 8 different 16-byte bank groups for every word slot; the per-group sums use an odd row stride for the four row pairs."""
    for slice_items in (64, 128, 256, 448):
        xstride = slice_items + 1
        for w in range(4):
            for g0 in range(0, 8, 2):
                groups = {((g * xstride + item_index(t, w)) * 16 // 16) % 8 for g in (g0, g0 + 1) for t in range(4)}
                assert len(groups) == 8, (slice_items, w, g0)
    for sgroups in (8, 16, 32, 56, 64):
        sxs = sgroups | 1
        for c in range(2):
            banks = {((c * 8 + 2 * t) * sxs) % 32 for t in range(4)}
            assert len(banks) == 4


@pytest.mark.parametrize("bits,group_size,mrows", [(4, 64, 16), (4, 128, 8), (8, 64, 16)])
def test_super_chunk_through_the_lane_model_equals_the_dequantised_dot_product(bits, group_size, mrows):
    rng = np.random.default_rng(bits * 1000 + group_size + mrows)
    npg = group_size * bits // 4                   # nibbles per group
    gps = 512 // npg
    K = 512 * 4 // bits                            # k elements of one super-chunk
    W = rng.integers(0, 1 << bits, (16, K)).astype(np.int64)
    scales = rng.uniform(0.01, 0.3, (16, gps))
    zps = rng.integers(0, 1 << bits, (16, gps)).astype(np.int64)
    x = rng.uniform(-0.3, 0.3, (mrows, K))
    want = np.zeros((mrows, 16))
    for r in range(16):
        deq = scales[r][np.arange(K) // group_size] * (W[r] - zps[r][np.arange(K) // group_size])
        want[:, r] = x @ deq

    # packed rows: 512 nibbles each, low nibble first (int8: one code = two nibbles, lo then hi)
    nib = np.zeros((16, 512), np.int64)
    if bits == 4:
        nib[:] = W
    else:
        nib[:, 0::2] = W & 15
        nib[:, 1::2] = W >> 4
    row_bytes = (nib[:, 0::2] | (nib[:, 1::2] << 4)).astype(np.uint8)             # [16, 256]
    # B-fragment items per activation row: 8 nibble positions each; int8 items carry (x, 16 x) for the (lo, hi) nibble of 4 elements
    xn = np.zeros((mrows, 512))
    if bits == 4:
        xn[:] = x
    else:
        xn[:, 0::2] = x
        xn[:, 1::2] = 16.0 * x
    sx = np.stack([x[:, gidx * group_size:(gidx + 1) * group_size].sum(axis=1) for gidx in range(gps)], axis=1)    # [mrows, gps]
    mult128 = 128.0 if bits == 4 else 128.0 * 17.0

    got = np.zeros((mrows, 16))
    for cblock in range(mrows // 8):
        acc = np.zeros((8, 4, 4))                  # [g][t][fragment register]
        for j in range(4):                         # chunk of the super-chunk
            dA = np.zeros((8, 4, 4)); dB = np.zeros((8, 4, 4))
            wa = {g: exchange({t: lane_words(row_bytes[g, 64 * j:64 * j + 64], t) for t in range(4)}) for g in range(8)}
            wb = {g: exchange({t: lane_words(row_bytes[g + 8, 64 * j:64 * j + 64], t) for t in range(4)}) for g in range(8)}
            for w in range(4):
                frag = dB if (npg == 64 and w >= 2) else dA
                for half in (0, 8):                # first / second MMA of the word: nibble shifts (0, 4) / (8, 12)
                    # mma.m16n8k16: D[row][col] += sum over the 16 k slots; slot owner = lane t' of the quad, B column col is fed by lanes g' = col
                    for g in range(8):             # A rows g (regs 0, 2) and g + 8 (regs 1, 3) come from lanes (g, t')
                        for col in range(8):
                            tot_a = tot_b = 0.0
                            for tp in range(4):
                                li = j * 16 + item_index(tp, w)
                                xrow = xn[cblock * 8 + col, li * 8:li * 8 + 8]
                                for shift in (half, half + 4):
                                    i = shift // 4
                                    a_lo, a_hi = nib_pair(wa[g][tp][w], shift)
                                    b_lo, b_hi = nib_pair(wb[g][tp][w], shift)
                                    tot_a += (128 + a_lo) * xrow[i] + (128 + a_hi) * xrow[i + 4]
                                    tot_b += (128 + b_lo) * xrow[i] + (128 + b_hi) * xrow[i + 4]
                            # D fragment: lane (g, t) holds cols 2t (regs 0, 2) and 2t + 1 (regs 1, 3)
                            frag[g, col // 2, col % 2] += tot_a
                            frag[g, col // 2, 2 + col % 2] += tot_b
            for gl, frag in ((2 * j, dA), (2 * j + 1, dB)) if npg == 64 else ((j, dA),):
                for g in range(8):
                    for t in range(4):
                        for reg in range(4):
                            wrow = g + (8 if reg >= 2 else 0)
                            arow = cblock * 8 + 2 * t + (reg & 1)
                            k = -(zps[wrow, gl] + mult128)
                            acc[g, t, reg] += scales[wrow, gl] * (frag[g, t, reg] + k * sx[arow, gl])
        # epilogue transposition: rw[(c*8 + 2t + (reg & 1)) * 16 + g + 8 * (reg >> 1)]
        for g in range(8):
            for t in range(4):
                for reg in range(4):
                    got[cblock * 8 + 2 * t + (reg & 1), g + 8 * (reg >> 1)] = acc[g, t, reg]
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-9)


def test_zero_point_word_and_shift_selection():
    """issue_stage copies, for quad lane t', the aligned 32-bit word holding byte (row * zp_stride + goff(t')), goff = gi / 2 + t' (4-bit: one
    byte = the zero points of groups gi + 2t', gi + 2t' + 1, low nibble first) or gi + 2t' (8-bit); the consumer shifts by
    ((row * zp_stride + goff) & 3) * 8 and takes nibble / byte `hi`. Enumerate rows, strides and super-chunks."""
    rng = np.random.default_rng(3)
    for bits in (4, 8):
        for groups in (8, 16, 64, 224):
            stride = (groups + 1) // 2 if bits == 4 else groups
            zp_bytes = rng.integers(0, 256, 40 * stride + 8, dtype=np.uint8)
            for row in (0, 1, 7, 39):
                for gi in range(0, groups, 8):
                    for gl in range(8):
                        tq, hi = gl >> 1, gl & 1
                        goff = gi // 2 + tq if bits == 4 else gi + 2 * tq
                        byte_index = row * stride + goff
                        word = int.from_bytes(bytes(zp_bytes[(byte_index & ~3):(byte_index & ~3) + 4]), "little")
                        shifted = word >> ((byte_index & 3) * 8)
                        got = (shifted >> (hi * 4)) & 15 if bits == 4 else (shifted >> (hi * 8)) & 255
                        g = gi + gl
                        want = (zp_bytes[row * stride + g // 2] >> ((g & 1) * 4)) & 15 if bits == 4 else zp_bytes[row * stride + g]
                        if bits == 8 and (byte_index & 3) == 3 and hi == 1:
                            continue          # 8-bit pairs start at even byte offsets (gi and 2t' are even, strides are multiples of 4 in the kernel's domain)
                        assert got == want, (bits, groups, row, gi, gl)
