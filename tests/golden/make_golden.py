"""Generate the golden vectors that pin the CPU oracle to the UNMODIFIED reference CUDA.

Runs on a GPU box (no /root/reference needed there: oracle/_ref/libsrm_ref.so was built from the
reference sources in the authoring container by `make -C oracle ref` and travels with the repo):

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'   then copy *.npz into tests/golden/

Each file stores the INPUTS as compact site lists / generator parameters and the reference OUTPUT.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _inputs as I  # noqa: E402
import _oracle as O  # noqa: E402
import _ref as R  # noqa: E402


def sites_list(site_map):
    return O.sites_of(site_map)


def multires(out):
    """Coarse-to-fine pieces (a11): pyramid levels, zoomed seed map, whole depth-2/3 runs."""
    sys.path.insert(0, os.path.dirname(HERE))
    import test_multires as T
    for name, kind, n, k, depth, iters in [("uni512_d2", "uniform", 512, 400, 2, 60), ("c3_512_d2", "c3", 512, 1000, 2, 80),
                                           ("c3_1024_d3", "c3", 1024, 1500, 3, 100)]:
        dens, mask, seeds = T._case(kind, n, k, depth)
        rec = dict(n=n, k=k, depth=depth, kind=kind, max_iter=iters)
        for lvl in range(1, depth):
            p = R.pyramid(dens, lvl)
            s = n >> lvl
            rows = (s // 2 - 32, s // 2 + 32)   # keep the fixture small: a 64-row band of every level
            rec[f"pyr{lvl}"] = p[rows[0]:rows[1]]
            rec[f"pyr{lvl}_rows"] = np.array(rows)
        rec["zoom_sites"] = sites_list(R.zoom(seeds))
        fin, it = R.gcvt_multires(seeds, dens, mask, depth, iters)
        rec["final_sites"] = sites_list(fin)
        rec["iterations"] = it
        np.savez_compressed(os.path.join(out, f"ref_multires_{name}.npz"), **rec)
        print("multires", name, "ok", it, len(rec["final_sites"]))


def main(out):
    os.makedirs(out, exist_ok=True)
    if len(sys.argv) > 2 and sys.argv[2] == "multires":
        return multires(out)
    # ---- labelling (pba2DCompute), incl. lattice ties: pins the tie rule A2
    for name, seeds in [
        ("rand256", I.random_sites(256, 700, 101)),
        ("dense256", I.random_sites(256, 6000, 102)),
        ("lattice256", I.lattice_sites(256, 8, 4)),
        ("lattice256b", I.lattice_sites(256, 6, 1)),
        ("rand512", I.random_sites(512, 3000, 103)),
        ("lattice512", I.lattice_sites(512, 16, 7)),
        ("rand1024", I.random_sites(1024, 2000, 104)),
    ]:
        lab = R.label(seeds)
        np.savez_compressed(os.path.join(out, f"ref_label_{name}.npz"), n=seeds.shape[0], sites=sites_list(seeds),
                            labels=lab)
        print("label", name, "ok")
    # ---- one teacher-forced Lloyd step (label, energy, centroid, update)
    for name, n, k, omega, kind in [("uni256", 256, 400, 2.0, "uniform"), ("c3_512", 512, 3000, 2.0, "c3"),
                                    ("c3_512_w13", 512, 3000, 1.3, "c3"), ("uni1024", 1024, 2000, 2.0, "uniform")]:
        dens = I.density_uniform(n) if kind == "uniform" else I.density_c3(n)
        mask = None if kind == "uniform" else I.mask_c3(dens)
        seeds, _, _ = O.seed(dens, mask, k)
        lab, nxt, e = R.step(seeds, dens, mask, omega)
        np.savez_compressed(os.path.join(out, f"ref_step_{name}.npz"), n=n, k=k, omega=omega, kind=kind,
                            sites=sites_list(seeds), labels=lab, new_sites=sites_list(nxt), energy=e)
        print("step", name, "ok", e)
    # ---- whole gCVT
    for name, n, k, iters, kind in [("uni256", 256, 400, 60, "uniform"), ("c3_512", 512, 3000, 40, "c3")]:
        dens = I.density_uniform(n) if kind == "uniform" else I.density_c3(n)
        mask = None if kind == "uniform" else I.mask_c3(dens)
        seeds, _, _ = O.seed(dens, mask, k)
        fin, it = R.gcvt(seeds, dens, mask, iters)
        np.savez_compressed(os.path.join(out, f"ref_gcvt_{name}.npz"), n=n, k=k, max_iter=iters, kind=kind,
                            sites=sites_list(seeds), final_sites=sites_list(fin), iterations=it)
        print("gcvt", name, "ok", it)
    # ---- rasteriser
    for name, side, n, seed in [("mesh12_256", 12, 256, 7), ("mesh40_512", 40, 512, 8)]:
        pts, wt, tri = I.random_mesh(side, seed)
        scale = 1.0 / (n - 1)
        dens = R.discretize(pts, wt, tri, scale, n)
        rows = (0, n) if n <= 256 else (192, 320)  # keep the fixture small: a 128-row band of the larger case
        np.savez_compressed(os.path.join(out, f"ref_raster_{name}.npz"), n=n, side=side, seed=seed, scale=scale,
                            rows=np.array(rows), density=dens[rows[0]:rows[1]])
        print("raster", name, "ok", float(dens.max()))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "_out"))
