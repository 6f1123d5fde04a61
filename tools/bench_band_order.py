#!/usr/bin/env python
"""A/B of the band kernel's CTA -> band order (option "band_order", csrc/srm_band.cu "Band order") on the headline
workload, in one process: row order (0) against longest-first by the run counts of an earlier iteration (1), alternated
twice.  Per setting: a fresh site set, 12 warm-up iterations (the order is rebuilt in iterations 1 and 11), 100 iterations
with CUDA events between the stages (srm_iterate_profiled) and 100 plain iterations between two events.  The site lists
after the run must be identical in every setting.

    python tools/bench_band_order.py [--n 8192] [--sites 100000] > gpurun_out/band_order.json
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench                      # noqa: E402
import surface_remesher_b200 as S  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--sites", type=int, default=100000)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--orders", default="0,1,0,1", help="settings to run, in order (e.g. \"0\" for one timing of the row order, "
                    "used with SRM_BAND_RPW=2 in the environment for the 16-row band A/B)")
    a = ap.parse_args()
    import torch
    n = a.n
    dens, mask, vor = bench.make_inputs(n, a.sites, pinned=False)
    out = {"grid": n, "sites": a.sites, "steps": a.steps, "SRM_BAND_RPW": os.environ.get("SRM_BAND_RPW", "1"), "runs": []}
    ref = None
    with S.Context(n) as c:
        c.set_stream(torch.cuda.current_stream().cuda_stream)
        c.set_density(dens); c.set_mask(mask)
        for order in [int(x) for x in a.orders.split(",")]:
            c.set_option("band_order", order)
            c.set_site_map(vor)
            c.iterate(12)
            st = c.iterate_profiled(a.steps)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            c.iterate(a.steps)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            sites = np.sort(c.get_sites())
            perm, cost = c.debug_band_order()
            if ref is None:
                ref = sites
            r = {"band_order": order, "k_band_us": round(st["band_fused"] / a.steps * 1e3, 2),
                 "profiled_step_us": round(st["iteration"] / a.steps * 1e3, 2), "plain_step_us": round(ms / a.steps * 1e3, 2),
                 "it_per_s": round(a.steps / ms * 1e3, 1), "sites_identical": bool(np.array_equal(sites, ref)),
                 "sites_sha1": bench.sites_sha1(sites), "perm_is_identity": bool((perm == np.arange(len(perm))).all()),
                 "cost_min_mean_max": [int(cost.min()), int(cost.mean()), int(cost.max())]}
            out["runs"].append(r)
            print(json.dumps(r), file=sys.stderr, flush=True)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
