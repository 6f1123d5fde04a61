#!/usr/bin/env python
"""The two streaming kernels of a gCVT call on the headline grid, both builds of each (srm_set_variant):
k_prefix (4 B/px read, 24 B/px written; 0 = 128/64-bit stores, 1 = 256-bit stores) and k_expand (4 B/px written from
~20 MB of runs; 0 = binary search per 4-pixel group, 1 = two-level lookup), device time per launch from CUDA events
(srm_time_kernel: 20 launches after one untimed), as GB/s against the measured copy and write-only stream rates; and the
stand-alone centroid pass over the dense label map (k_centroid_dense, 8 B/px read).

    python tools/bench_streams.py [--n 8192] [--sites 100000] > gpurun_out/streams.json
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench                      # noqa: E402
import surface_remesher_b200 as S  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--sites", type=int, default=100000)
    a = ap.parse_args()
    n = a.n
    dens, mask, vor = bench.make_inputs(n, a.sites, pinned=False)
    peak, src = bench.measured_peaks()
    out = {"grid": n, "sites": a.sites, "copy_peak_GBs": peak, "peak_source": src}
    sp = os.path.join(ROOT, "profiles", "r2_stream_peaks.json")
    if os.path.exists(sp):
        out["write_only_peak_GBs"] = json.load(open(sp))["write_only_fill_GBs"]
    N = n * n
    with S.Context(n) as c:
        c.set_density(dens); c.set_mask(mask); c.set_site_map(vor)
        c.iterate(20, stop_rule=False)   # Lloyd-relaxed sites: the run structure of a real final labelling
        c.label()
        runs, _ = c.debug_counts()
        out["runs"] = int(runs)
        for v in (0, 1, 2):   # expand: 0 / 1 / 2 (bitmap form, default); prefix: 0 / 1 (default)
            S.api.set_variant("prefix", min(v, 1)); S.api.set_variant("expand", v)
            tp = min(c.time_kernel("prefix", 20) for _ in range(3))
            te = min(c.time_kernel("expand", 20) for _ in range(3))
            out[f"variant{v}"] = {
                "k_prefix_ms": tp, "k_prefix_GBs": 28.0 * N / tp / 1e6, "k_prefix_frac_of_copy_peak": 28.0 * N / tp / 1e6 / peak,
                "k_expand_ms": te, "k_expand_GBs": (4.0 * N + 8.0 * runs) / te / 1e6,
                "k_expand_frac_of_copy_peak": (4.0 * N + 8.0 * runs) / te / 1e6 / peak}
        S.api.set_variant("prefix", -1); S.api.set_variant("expand", -1)
        # the stand-alone centroid pass over the dense label map (srm_centroid.cu): 8 B/px read + one hash bucket and three
        # fp64 REDs per run; next to it the run-based accumulation kernel it is the alternative to (per run: 16 B prefix pair)
        for v in (0, 1):   # 0 = k_centroid_dense (first build), 1 = k_centroid_dense2 (default)
            S.api.set_variant("centroid", v)
            tc = min(c.time_kernel("centroid", 20) for _ in range(3))
            tce = min(c.time_kernel("centroid_energy", 20) for _ in range(3))
            out[f"centroid_dense_v{v}"] = {"ms": tc, "GBs": 8.0 * N / tc / 1e6, "frac_of_copy_peak": 8.0 * N / tc / 1e6 / peak,
                                           "with_energy_ms": tce, "with_energy_frac_of_copy_peak": 8.0 * N / tce / 1e6 / peak,
                                           "algorithmic_bytes": 8.0 * N, "SRM_CEN_WAVES": os.environ.get("SRM_CEN_WAVES", "4")}
        S.api.set_variant("centroid", -1)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
