"""CPU model of the warp algorithm of k_centroid_dense2 (csrc/srm_centroid.cu), lane by lane: in-thread prefix sums,
sub-run sums picked by the first / last boundary of a 3-bit mask, hand-off of the leading sub-run to the left lane,
heads, the segmented shuffle tree that stops at the first distance no lane can use.  Checked against direct per-run sums
on rows full of short runs (lengths 1-3 inside a thread, a-b-a patterns, labels repeated in non-adjacent runs, runs
across the 128-pixel warp boundary), where every branch of the kernel is taken.  The GPU test of the kernel itself is
tests/test_gpu_centroid.py; this model pins the logic the CUDA code transcribes."""
import numpy as np


def _warp_step(lab, d, x_base, emit):
    """lab, d: 128 pixels of one row (int64 labels, float64 density); x_base = x of pixel 0.  emit(label, W, X)."""
    L = lab.reshape(32, 4); D = d.reshape(32, 4)
    item = np.zeros(32, np.int64); W = np.zeros(32); X = np.zeros(32); WL = np.zeros(32); XL = np.zeros(32)
    uni = np.zeros(32, bool); first = np.zeros(32, np.int64)
    for t in range(32):
        d0, d1, d2, d3 = D[t]
        x0 = float(x_base + 4 * t)
        p = [0.0, d0, d0 + d1, d0 + d1 + d2, d0 + d1 + d2 + d3]
        q = [0.0, 0.0, d1, d1 + 2.0 * d2, d1 + 2.0 * d2 + 3.0 * d3]
        m = int(L[t, 1] != L[t, 0]) | (int(L[t, 2] != L[t, 1]) << 1) | (int(L[t, 3] != L[t, 2]) << 2)
        f = (m & -m).bit_length()            # __ffs
        l = m.bit_length()                   # 32 - __clz
        pf = p[1] if f == 1 else p[2] if f == 2 else p[3]
        qf = 0.0 if f == 1 else q[2] if f == 2 else q[3]
        pl = 0.0 if l == 0 else p[1] if l == 1 else p[2] if l == 2 else p[3]
        ql = 0.0 if l <= 1 else q[2] if l == 2 else q[3]
        uni[t] = m == 0; first[t] = L[t, 0]; item[t] = L[t, 3]
        W[t] = p[4] - pl; X[t] = x0 * W[t] + (q[4] - ql)
        if l > f:
            if m == 7:
                emit(L[t, 1], d1, x0 * d1 + d1)
                emit(L[t, 2], d2, x0 * d2 + 2.0 * d2)
            else:
                w = pl - pf
                emit(L[t, 1] if f == 1 else L[t, 2], w, x0 * w + (ql - qf))
        WL[t] = pf; XL[t] = x0 * pf + qf
    left = np.concatenate([item[:1], item[:-1]])                      # __shfl_up(item, 1): lane 0 keeps its own
    lane = np.arange(32)
    accepted = ~uni & (lane > 0) & (left == first)
    for t in range(32):
        if not uni[t] and not accepted[t]:
            emit(first[t], WL[t], XL[t])
    gW = np.where(accepted, WL, 0.0); gX = np.where(accepted, XL, 0.0)
    W[:31] += gW[1:]; X[:31] += gX[1:]                               # __shfl_down(.., 1), lane 31 excluded
    head = (lane == 0) | ~uni | (item != left)
    heads = sum(1 << t for t in range(32) if head[t])
    after = [(heads >> 1) >> t for t in range(32)]
    steps = 0
    o = 1
    while o < 32:
        can = np.array([t + o < 32 and (after[t] & ((1 << o) - 1)) == 0 for t in range(32)])
        if not can.any():
            break
        W2 = np.concatenate([W[o:], W[32 - o:]]); X2 = np.concatenate([X[o:], X[32 - o:]])   # out-of-range lanes: own value
        W = np.where(can, W + W2, W); X = np.where(can, X + X2, X)
        steps += 1
        o <<= 1
    for t in range(32):
        if head[t]:
            emit(item[t], W[t], X[t])
    return steps


def _row_labels(rng, n, mean_run, nlabels):
    lab = np.empty(n, np.int64)
    x = 0
    prev = -1
    while x < n:
        ln = 1 + rng.geometric(1.0 / mean_run) - 1 if mean_run > 1 else 1
        ln = max(1, int(ln))
        v = int(rng.integers(0, nlabels))
        if v == prev:
            v = (v + 1) % nlabels
        lab[x:x + ln] = v
        prev = v
        x += ln
    return lab


def test_warp_model_matches_direct_run_sums():
    rng = np.random.default_rng(11)
    n = 1024
    max_steps = 0
    for mean_run, nlabels in ((1, 3), (1.5, 2), (2, 4), (3, 5), (7, 50), (26, 1000), (200, 10), (5000, 3)):
        for rep in range(6):
            lab = _row_labels(rng, n, mean_run, nlabels)
            if rep == 0:
                lab[100:140] = np.tile([7, 8, 7, 7], 10)            # a-b-a inside threads, the same label in non-adjacent runs
                lab[125:131] = 9                                      # a run across the warp boundary at 128
            d = rng.random(n) + 0.25
            got = {}

            def emit(label, W, X):
                a = got.setdefault(int(label), [0.0, 0.0, 0])
                a[0] += W; a[1] += X; a[2] += 1
            for w in range(n // 128):
                max_steps = max(max_steps, _warp_step(lab[128 * w:128 * (w + 1)], d[128 * w:128 * (w + 1)], 128 * w, emit))
            xs = np.arange(n, dtype=np.float64)
            runs = 1 + int((lab[1:] != lab[:-1]).sum())
            crossing = int(sum(lab[128 * w] == lab[128 * w - 1] for w in range(1, n // 128)))
            emits = sum(a[2] for a in got.values())
            assert emits == runs + crossing, (mean_run, emits, runs, crossing)   # one emit per run and warp segment: nothing split further
            for v in np.unique(lab):
                sel = lab == v
                W = d[sel].sum(); X = (d * xs)[sel].sum()
                assert abs(got[int(v)][0] - W) <= 1e-12 * max(W, 1) and abs(got[int(v)][1] - X) <= 1e-12 * max(X, 1), (mean_run, v)
    assert max_steps == 5   # the 5000-pixel runs need the whole tree; short runs stop early (checked below)


def test_tree_stops_early_on_short_runs():
    rng = np.random.default_rng(5)
    lab = _row_labels(rng, 128, 6, 100)
    assert (np.diff(np.flatnonzero(np.concatenate([[True], lab[1:] != lab[:-1]]))).max()) <= 60
    steps = _warp_step(lab, rng.random(128), 0, lambda *a: None)
    assert steps < 5
