#!/usr/bin/env python
"""Host side of the drop-in boundary on the headline grid: where the time of a gCVT(host buffers) call goes, and how the
pageable-copy pipeline (csrc/srm_host.cu) responds to its two knobs (worker threads, staging chunk size).

For every (threads, chunk) setting: upload of the 268 MB density from a pageable numpy array (srm_set_density: staging
pipeline + k_prefix), download of the 268 MB label map into a pageable array (srm_get_labels: k_expand + pipeline), and a
whole gCVT call with `--iters` iterations (wall clock, best of 3).  SRM_TRACE=1 prints the stages of every gCVT call to
stderr.

    python tools/bench_host.py [--iters 20] > gpurun_out/host.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench                      # noqa: E402
import surface_remesher_b200 as S  # noqa: E402


def best_of(fn, reps=3):
    b = None
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); dt = time.perf_counter() - t0
        b = dt if b is None else min(b, dt)
    return b * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--sites", type=int, default=100000)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--quick", action="store_true", help="default setting and (8 threads, 2 MB chunks) only")
    a = ap.parse_args()
    n = a.n
    import torch
    dens, mask, vor = bench.make_inputs(n, a.sites, pinned=False)
    out = {"grid": n, "iterations_per_call": a.iters, "host_threads_available": os.cpu_count(), "settings": []}
    buf = np.empty((n, n, 2), np.int16)
    lab = np.empty((n, n, 2), np.int16)
    grid = [(0, 4096), (4, 4096), (8, 2048), (8, 8192), (12, 4096), (16, 4096), (16, 8192), (0, 4096)]
    for threads, chunk_kb in (grid[:1] + grid[2:3] if a.quick else grid):
        S.api.host_config(threads, chunk_kb)
        with S.Context(n) as c:
            c.set_mask(mask); c.set_site_map(vor)
            up = best_of(lambda: (c.set_density(dens), c.synchronize()))
            c.label()
            down = best_of(lambda: c.get_labels(lab))

        def call():
            buf[:] = vor
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            S.gCVT(buf, dens, mask, n, 1, a.iters)
            torch.cuda.synchronize()
            return time.perf_counter() - t0
        call()
        e2e = min(call() for _ in range(3)) * 1e3
        out["settings"].append({"threads": threads, "chunk_kb": chunk_kb, "upload_density_ms": round(up, 2),
                                "upload_GBs": round(dens.nbytes / up / 1e6, 1), "download_labels_ms": round(down, 2),
                                "download_GBs": round(lab.nbytes / down / 1e6, 1), "gcvt_call_ms": round(e2e, 2),
                                "gcvt_it_per_s": round(a.iters / e2e * 1e3, 1)})
        print(json.dumps(out["settings"][-1]), file=sys.stderr, flush=True)
    S.api.host_config(0, 4096)
    S.lib().srm_release_cache()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
