#!/usr/bin/env python
"""Secondary measurements for DESIGN.md / profiles (not the driver's bench): BASELINE.json configs C2, C3, C5
device-resident, the reference CUDA loop on the same GPU where it supports the size, the JFA kernel family
against the HBM roofline, the final-labelling expand kernel and the rasteriser.

    gpurun -- 'python tools/bench_configs.py > gpurun_out/configs.json'
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _inputs as I  # noqa: E402
import _ref as R  # noqa: E402
import surface_remesher_b200 as S  # noqa: E402
import torch  # noqa: E402


def seeded(n, k, kind):
    dens = I.density_uniform(n) if kind == "uniform" else np.concatenate(
        [I.density_c3(n, rows=(r, min(n, r + 512))) for r in range(0, n, 512)])
    mask = None if kind == "uniform" else I.mask_c3(dens)
    vor = np.empty((n, n, 2), np.int16)
    S.api._ck(S.lib().srm_seed(vor.ctypes.data, dens.ctypes.data, None if mask is None else mask.ctypes.data, k, n, None))
    return dens, mask, vor


def time_ours(n, dens, mask, vor, iters, warm=10):
    with S.Context(n) as c:
        c.set_density(dens); c.set_mask(mask); c.set_site_map(vor)
        c.iterate(warm); c.synchronize()
        st = c.iterate_profiled(iters)
        runs, ovf = c.debug_counts()
        return {"it_per_s": iters / (st["iteration"] / 1e3), "ms_per_it": st["iteration"] / iters,
                "stages_us": {k: round(v / iters * 1e3, 1) for k, v in st.items()}, "runs": runs, "robust_rows": ovf,
                "sites": c.state()["num_sites"]}


def main():
    out = {}
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    for name, n, k, kind, iters in [("C2 uniform 4096^2 20k", 4096, 20000, "uniform", 100),
                                    ("C3 aniso 8192^2 100k", 8192, 100000, "c3", 50),
                                    ("C5 one mesh 2048^2 10k", 2048, 10000, "c3", 100),
                                    ("C1-size 1024^2 2k", 1024, 2000, "c3", 100)]:
        dens, mask, vor = seeded(n, k, kind)
        rec = {"ours": time_ours(n, dens, mask, vor, iters)}
        if R.available() and n <= 8192:
            try:
                ms = R.loop_timed(vor, dens, mask, min(iters, 30))
                rec["reference_cuda_loop_it_per_s"] = min(iters, 30) / (ms / 1e3)
            except Exception as e:
                rec["reference_cuda_loop_it_per_s"] = f"failed: {e}"
        out[name] = rec
        print(name, json.dumps(rec), file=sys.stderr, flush=True)
        del dens, mask, vor

    # C5: batch of independent 2048^2 meshes on one GPU ("replicas"): several contexts driven from threads, each on
    # its own stream, so that the small grids (128 bands = 128..256 CTAs each) fill the GPU together
    import threading
    n, k = 2048, 10000
    dens, mask, vor = seeded(n, k, "c3")
    for conc in (1, 2, 4, 8):
        ctxs = []
        for _ in range(conc):
            c = S.Context(n); c.set_density(dens); c.set_mask(mask); c.set_site_map(vor); c.iterate(5); ctxs.append(c)
        for c in ctxs: c.synchronize()
        iters = 200
        t0 = time.perf_counter()
        th = [threading.Thread(target=lambda c=c: (c.iterate(iters), c.synchronize())) for c in ctxs]
        for t in th: t.start()
        for t in th: t.join()
        dt = time.perf_counter() - t0
        out[f"C5 batch 2048^2 x{conc} concurrent"] = {"mesh_iterations_per_s": conc * iters / dt,
                                                      "meshes_of_100_iterations_per_s": conc * iters / dt / 100}
        print(f"C5 x{conc}", out[f"C5 batch 2048^2 x{conc} concurrent"], file=sys.stderr, flush=True)
        for c in ctxs: c.close()
    del dens, mask, vor

    # JFA family + expand at 8192^2: HBM streaming kernels
    n = 8192
    dens, mask, vor = seeded(n, 100000, "c3")
    with S.Context(n) as c:
        c.set_site_map(vor)
        steps = [1] + [n >> (i + 1) for i in range(13)]
        lab = torch.empty((n, n, 2), dtype=torch.int16, device="cuda")
        c.label_jfa(steps, out=lab)  # warm-up
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(3):
            c.label_jfa(steps, out=lab)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
        N = n * n
        out["jfa_1+JFA_8192"] = {"passes": len(steps), "ms_per_labelling": dt * 1e3, "GBs_algorithmic": 8.0 * N * len(steps) / dt / 1e9,
                                 "frac_of_measured_peak": 8.0 * N * len(steps) / dt / 1e9 / peak,
                                 "note": "includes a 4 B/px fill + scatter and a 4 B/px device copy of the result"}
        c.label()
        glab = torch.empty((n, n, 2), dtype=torch.int16, device="cuda")
        c.get_labels(out=glab)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5):
            c.get_labels(out=glab)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
        out["expand_8192"] = {"ms": dt * 1e3, "GBs_algorithmic": 4.0 * N / dt / 1e9, "frac_of_measured_peak": 4.0 * N / dt / 1e9 / peak}
        # exact vs JFA error
        ex = glab.cpu().numpy(); jf = lab.cpu().numpy()
        ys, xs = np.mgrid[0:n, 0:n]
        dj = (jf[..., 0].astype(np.int64) - xs) ** 2 + (jf[..., 1].astype(np.int64) - ys) ** 2
        de = (ex[..., 0].astype(np.int64) - xs) ** 2 + (ex[..., 1].astype(np.int64) - ys) ** 2
        out["jfa_1+JFA_8192"]["wrong_distance_pixels"] = int((dj > de).sum())
        out["jfa_1+JFA_8192"]["label_mismatch_pixels"] = int((jf != ex).any(axis=2).sum())
    # rasteriser
    for side, n in [(56, 1024), (80, 2048)]:
        pts, wt, tri = I.random_mesh(side, 5)
        d = np.empty((n, n), np.float32)
        S.discretization_d(pts, wt, len(wt), tri, len(tri), d, 1.0 / (n - 1), n)
        t0 = time.perf_counter(); S.discretization_d(pts, wt, len(wt), tri, len(tri), d, 1.0 / (n - 1), n); t_our = time.perf_counter() - t0
        rec = {"triangles": len(tri), "ours_ms": t_our * 1e3}
        if R.available():
            R.discretize(pts, wt, tri, 1.0 / (n - 1), n)
            t0 = time.perf_counter(); dr = R.discretize(pts, wt, tri, 1.0 / (n - 1), n); rec["reference_ms"] = (time.perf_counter() - t0) * 1e3
            rec["bit_identical"] = bool(np.array_equal(d.view(np.uint32), dr.view(np.uint32)))
        out[f"raster_{n}_{len(tri)}tri"] = rec
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
