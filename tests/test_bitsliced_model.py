"""Executable model (numpy) of the bit-sliced counters of csm_coarse_bits_kernel
(gloc3d_b200/csrc/csm.cu): per point NP packed 32-bit words (two candidate rows of <= 16
column bits each), 16 points folded by a carry-save adder tree into one weight-16 carry that
ripples into 8 high planes, counts read back bit by bit.  Checks the adder tree against plain
integer counting and the capacity invariant behind kBitChunk = 16 x 255 = 4080 points."""
import re
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def csa(l, x, y):
    """(carry, sum) of three bit-vectors: the kernel's `csa` lambda."""
    u = l ^ x
    return (l & x) | (u & y), u ^ y


def count_bitsliced(words):
    """words: [n_points, NP] uint32.  Returns the 12 counter planes after the kernel's loop."""
    n, NP = words.shape
    pad = (-n) % 16
    w = np.concatenate([words, np.zeros((pad, NP), np.uint32)])
    ones = np.zeros(NP, np.uint32); twos = ones.copy(); fours = ones.copy(); eights = ones.copy()
    hi = np.zeros((8, NP), np.uint32)
    for p in range(0, len(w), 16):
        e = []
        for o in (0, 8):                         # oct_words
            f = []
            for q in (0, 4):                     # two pair_words + csa into twos
                ta, ones = csa(ones, w[p + o + q], w[p + o + q + 1])
                tb, ones = csa(ones, w[p + o + q + 2], w[p + o + q + 3])
                fa, twos = csa(twos, ta, tb)
                f.append(fa)
            ea, fours = csa(fours, f[0], f[1])
            e.append(ea)
        c16, eights = csa(eights, e[0], e[1])
        for i in range(8):                       # ripple the weight-16 carry into the high planes
            t = hi[i] & c16
            hi[i] ^= c16
            c16 = t
        assert not c16.any(), "high planes overflowed"
    return [ones, twos, fours, eights] + [hi[i] for i in range(8)]


def extract(planes, word, bit):
    return sum(int((pl[word] >> np.uint32(bit)) & np.uint32(1)) << i for i, pl in enumerate(planes))


def test_adder_tree_equals_integer_counts():
    rng = np.random.default_rng(0)
    for n, NP, dens in ((1, 1, 0.5), (17, 4, 0.3), (1000, 4, 0.05), (2049, 7, 0.9), (4080, 4, 1.0)):
        bits = rng.random((n, NP, 32)) < dens
        words = (bits.astype(np.uint64) << np.arange(32, dtype=np.uint64)).sum(axis=2).astype(np.uint32)
        planes = count_bitsliced(words)
        ref = bits.sum(axis=0)                   # [NP, 32] integer counts
        for wd in range(NP):
            for b in range(32):
                assert extract(planes, wd, b) == ref[wd, b], (n, NP, wd, b)


def test_capacity_matches_the_kernel_constant():
    # every bit set in every point: 4080 points is the most 4 low planes (<= 15) + 8 high planes
    # (<= 255 carries of weight 16) can count
    src = open(os.path.join(ROOT, "gloc3d_b200", "csrc", "csm.cu")).read()
    chunk = int(re.search(r"constexpr int kBitChunk = (\d+);", src).group(1))
    assert chunk == 16 * 255 and chunk % 16 == 0
    ones = np.full((chunk, 1), 0xFFFFFFFF, np.uint32)
    planes = count_bitsliced(ones)
    assert extract(planes, 0, 0) == chunk and extract(planes, 0, 31) == chunk
    try:
        count_bitsliced(np.full((chunk + 16, 1), 1, np.uint32))
    except AssertionError:
        pass
    else:
        raise AssertionError("one more block of 16 points must overflow the 8 high planes")
