#!/usr/bin/env python
"""Why the band kernel takes its bands longest-first (csrc/srm_band.cu "Band order"): list-scheduling model of the CTA
waves on the run counts of a real labelling.  CPU only (uses the test oracle for the labelling).

Labels the C3 workload after a few Lloyd iterations with the oracle, counts the runs per 8-row band (the cost the band
kernel records), and schedules the bands greedily on `slots` resident CTAs in row order, in decreasing order of the true
cost, and in decreasing order of a site-count proxy.  Writes the profile to tests/golden/c3_<n>_band_runs.npy (used by
tests/test_band_order_model.py).

    python tools/sim_band_order.py [n=8192] [sites=100000] [iterations=5] [slots=592]
Result at the defaults: row order 5490, longest-first 4767 (-13 %), site-count proxy 5051, lower bound 4315.
"""
import heapq
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _inputs as I   # noqa: E402
import _oracle as O   # noqa: E402


def makespan(order, cost, slots):
    h = [0.0] * slots
    heapq.heapify(h)
    end = 0.0
    for b in order:
        t = heapq.heappop(h) + cost[b]
        end = max(end, t)
        heapq.heappush(h, t)
    return end


def main():
    a = sys.argv[1:]
    n = int(a[0]) if len(a) > 0 else 8192
    k = int(a[1]) if len(a) > 1 else 100000
    iters = int(a[2]) if len(a) > 2 else 5
    slots = int(a[3]) if len(a) > 3 else 592
    dens = I.density_c3(n)
    mask = I.mask_c3(dens, every=16)
    seeds, _, _ = O.seed(dens, mask, k)
    lab = O.gcvt(seeds, dens, mask, iters, stop_rule=0)[0] if iters else O.label_exact(seeds)
    lab32 = lab.view(np.int32).reshape(n, n)
    runs = 1 + (lab32[:, 1:] != lab32[:, :-1]).sum(1)
    ys, xs = np.mgrid[0:n, 0:n]
    site_rows = ((lab[..., 0] == xs) & (lab[..., 1] == ys)).sum(1)
    nb = n // 8
    cost = runs.reshape(nb, 8).sum(1).astype(float)
    proxy = np.convolve(site_rows.reshape(nb, 8).sum(1).astype(float), np.ones(9), mode="same")
    print(f"bands {nb}, slots {slots}: cost min / mean / max {cost.min():.0f} / {cost.mean():.0f} / {cost.max():.0f}")
    print(f"row order           {makespan(range(nb), cost, slots):.0f}")
    print(f"longest first       {makespan(np.argsort(-cost, kind='stable'), cost, slots):.0f}")
    print(f"site-count proxy    {makespan(np.argsort(-proxy, kind='stable'), cost, slots):.0f}")
    print(f"lower bound         {max(cost.max(), cost.sum() / slots):.0f}")
    np.save(os.path.join(ROOT, "tests", "golden", f"c3_{n}_band_runs.npy"), cost.astype(np.int32))


if __name__ == "__main__":
    main()
