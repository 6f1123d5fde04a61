"""CPU model of k_expand3's index arithmetic (csrc/srm_label.cu): the run of pixel x from a 1 bit/px bitmap of run starts
and an exclusive popcount prefix per 32-pixel word, and the runs of the next three pixels from their own start bits,
against numpy's searchsorted on the run starts."""
import numpy as np


def _expand_bitmap(starts, labels, n):
    words = np.zeros(1024, np.uint64)
    for s in starts:
        words[s >> 5] |= np.uint64(1) << np.uint64(s & 31)
    pc = np.array([bin(int(w)).count("1") for w in words])
    pre = np.concatenate([[0], np.cumsum(pc)[:-1]])
    out = np.empty(n, np.int64)
    for q in range(n >> 2):
        x = q << 2; w = x >> 5; b = x & 31
        word = int(words[w])
        e0 = max(pre[w] + bin(word & ((2 << b) - 1)).count("1") - 1, 0)
        e1 = e0 + ((word >> (b + 1)) & 1); e2 = e1 + ((word >> (b + 2)) & 1); e3 = e2 + ((word >> (b + 3)) & 1)
        out[x:x + 4] = [labels[e0], labels[e1], labels[e2], labels[e3]]
    return out


def test_bitmap_popcount_expansion_matches_searchsorted():
    rng = np.random.default_rng(2)
    for n, k in ((256, 1), (256, 256), (1024, 40), (8192, 316), (32768, 1200), (512, 300)):
        starts = np.unique(np.concatenate([[0], rng.choice(n, size=k, replace=False)]))[:max(k, 1)]
        starts[0] = 0
        labels = rng.integers(0, 1 << 30, size=len(starts))
        got = _expand_bitmap(starts, labels, n)
        exp = labels[np.searchsorted(starts, np.arange(n), side="right") - 1]
        assert np.array_equal(got, exp), (n, k)
