"""CPU model of the band-order kernel (csrc/srm_band.cu:k_band_order).

The band kernel's CTAs may take the 8-row bands by decreasing cost of an earlier iteration (option "band_order").  The
order is built on the device by one CTA: key = cost << 12 | (4095 - band), padded to a power of two with -1, sorted in
decreasing order by a bitonic network.  This model restates the key packing and the network exactly (same loop nest,
same compare-exchange rule) and checks what the band kernel relies on: the result is a permutation of the bands, by
decreasing cost, equal costs in row order (so a flat profile is the identity).  The CUDA kernel is compared with this
contract in tests/test_gpu_band_order.py."""
import numpy as np
import pytest

ORDER_MAX = 4096


def band_order_model(cnt, R=8):
    """cnt: runs per row of the last labelling (any ints; clamped like the kernel)."""
    nb = len(cnt) // R
    cost = np.clip(np.asarray(cnt[:nb * R], np.int64), 0, 32767).reshape(nb, R).sum(1)
    P = 1
    while P < nb:
        P <<= 1
    key = np.full(P, -1, np.int64)
    key[:nb] = (np.minimum(cost, (1 << 18) - 1) << 12) | (ORDER_MAX - 1 - np.arange(nb))
    k = 2
    while k <= P:
        j = k >> 1
        while j > 0:
            i = np.arange(P)
            q = i ^ j
            sel = q > i
            i, q = i[sel], q[sel]
            a, b = key[i].copy(), key[q].copy()
            desc = (i & k) == 0
            swap = (a < b) == desc
            key[i] = np.where(swap, b, a)
            key[q] = np.where(swap, a, b)
            j >>= 1
        k <<= 1
    return (ORDER_MAX - 1 - (key[:nb] & (ORDER_MAX - 1))).astype(np.int32), cost


def check_order(perm, cost):
    nb = len(cost)
    assert sorted(perm.tolist()) == list(range(nb))
    c = cost[perm]
    assert (np.diff(c) <= 0).all()
    same = np.diff(c) == 0
    assert (np.diff(perm)[same] > 0).all()   # equal costs keep the row order


@pytest.mark.parametrize("nrows", [64, 192, 1024, 5376, 8192, 32768])
def test_band_order_is_a_permutation_by_decreasing_cost(nrows):
    rng = np.random.default_rng(nrows)
    for cnt in (rng.integers(0, 1800, nrows), rng.integers(0, 4, nrows), rng.integers(-5, 10 ** 9, nrows),
                np.zeros(nrows, np.int64), np.full(nrows, 40000)):
        perm, cost = band_order_model(cnt)
        check_order(perm, cost)


def test_flat_profile_is_the_identity():
    perm, _ = band_order_model(np.zeros(8192, np.int64))
    assert (perm == np.arange(1024)).all()
    perm, _ = band_order_model(np.full(2048, 7), R=16)
    assert (perm == np.arange(128)).all()


def test_list_scheduling_gain_on_the_headline_profile():
    """Why the order helps: runs per band of the C3 workload at 8192^2 / 100k sites after 5 Lloyd iterations
    (tests/golden/c3_8192_band_runs.npy, written by tools/sim_band_order.py).  1.73 waves of CTAs in row order start the
    expensive bands late; longest-first ends with the cheap ones (greedy list scheduling on 592 resident CTAs)."""
    import heapq
    import os
    cost = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c3_8192_band_runs.npy")).astype(np.int64)
    nb, slots = len(cost), 592

    def makespan(order):
        h = [0] * slots
        heapq.heapify(h)
        end = 0
        for b in order:
            t = heapq.heappop(h) + int(cost[b])
            end = max(end, t)
            heapq.heappush(h, t)
        return end
    rows = np.zeros(8 * nb, np.int64); rows[::8] = cost
    perm, c = band_order_model(rows)
    assert (c == cost).all()
    assert makespan(perm) < 0.9 * makespan(range(nb))
