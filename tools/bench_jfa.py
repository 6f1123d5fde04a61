#!/usr/bin/env python
"""Jump-flooding family (csrc/srm_jfa.cu) on the headline grid: device time of every launch of the 1+JFA schedule at
8192^2 / 100k sites, fused shared-memory tile form (mode 1) next to one plain kernel per pass (mode 0), with the
algorithmic bandwidth 8 B/px per launch against the measured HBM peak, and the exact labelling of the product path for
scale.  The two modes must give identical labels (checked here on the device output).

    python tools/bench_jfa.py [--n 8192] [--sites 100000] > gpurun_out/jfa.json
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
import test_jfa_tile_model as M    # noqa: E402  (launch plan of srm_launch_jfa)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--sites", type=int, default=100000)
    a = ap.parse_args()
    n = a.n
    dens, mask, vor = bench.make_inputs(n, a.sites, pinned=False)
    peak, src = bench.measured_peaks()
    steps = [1] + [n >> (i + 1) for i in range(int(np.log2(n)))]
    plan = M.group_steps(steps)
    out = {"grid": n, "sites": a.sites, "schedule": steps, "peak_GBs": peak, "peak_source": src,
           "bytes_per_launch": 8 * n * n}
    with S.Context(n) as c:
        import torch
        c.set_stream(torch.cuda.current_stream().cuda_stream)
        c.set_density(dens); c.set_mask(mask); c.set_site_map(vor)
        labs = {}
        for mode in (0, 1):
            c.set_option("jfa_mode", mode)
            c.label_jfa_timed(steps, mode)                       # warm-up
            best = None
            for _ in range(3):
                ms = c.label_jfa_timed(steps, mode)
                best = ms if best is None or sum(ms) < sum(best) else best
            labs[mode] = c.label_jfa(steps)
            names = [f"pass {s}" for s in steps] if mode == 0 else \
                    [("tile " + "+".join(map(str, arg))) if kind == "tile" else f"far {arg}" for kind, arg in plan]
            out[f"mode{mode}"] = {
                "total_ms": sum(best), "launches": len(best),
                "per_launch": [{"launch": nm, "ms": round(t, 4), "algorithmic_GBs": round(8 * n * n / t / 1e6, 1),
                                "frac_of_peak": round(8 * n * n / t / 1e6 / peak, 3)} for nm, t in zip(names, best)]}
        out["modes_identical"] = bool((labs[0] != labs[1]).sum() == 0)
        # the exact labelling of the product path (carry + band + expand), for scale
        c.label(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            c.label()
        e1.record(); torch.cuda.synchronize()
        out["exact_labelling_ms"] = e0.elapsed_time(e1) / 5
        exact = c.get_labels()
        ys, xs = np.mgrid[0:n, 0:n]
        d_j = (labs[1][..., 0].astype(np.int64) - xs) ** 2 + (labs[1][..., 1].astype(np.int64) - ys) ** 2
        d_e = (exact[..., 0].astype(np.int64) - xs) ** 2 + (exact[..., 1].astype(np.int64) - ys) ** 2
        out["jfa_pixels_farther_than_exact"] = float((d_j > d_e).mean())
        out["jfa_pixels_nearer_than_exact"] = int((d_j < d_e).sum())
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
