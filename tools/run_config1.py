#!/usr/bin/env python
"""BASELINE.json configs[0], the reference pipeline end to end on a bundled mesh (nefertiti.off through the CGAL-free
front end; fixture tests/golden/c1_nefertiti.npz):

    rasterise (srm_discretize) -> mask (srm_generate_mask) -> seed + gCVT (srm_seed, srm_gcvt) -> CDT input in
    delaunayInput order -> constrained Delaunay by the UNMODIFIED reference gDel2D (oracle/_ref/libgdel2d_ref.so, test
    infrastructure, in a subprocess) -> lift to the surface (srm_recover) -> result.off

and the same with the CPU oracle in place of every libsrm call; both must give identical sites, CDT input and final
vertices.  Prints one JSON object; writes the remeshed surface to `out_off` if given.

    python tools/run_config1.py [n=1024] [sites=2000] [iters=100] [out_off] [mesh=nefertiti|horse]

mesh=horse: the closed horse.off cut along data/horse.selection.txt (fixture tests/golden/c1_horse.npz); the two sides of
the seam are welded in the result, which must be a closed surface again (Euler characteristic 2).
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _inputs as I      # noqa: E402
import _oracle as O      # noqa: E402
import _ref as R         # noqa: E402
import surface_remesher_b200 as S   # noqa: E402
from surface_remesher_b200 import frontend as FE   # noqa: E402


def cdt_input(labels, mask, scale, l, b, border_uv):
    """delaunayInput (delaunay.h:46-79): free sites x-outer / y-inner, then the constraint points; segments = border edges."""
    n = labels.shape[0]
    sites = sorted((x, y) for (x, y) in I.site_set(labels) if not mask[y, x])
    free = np.array(sites, np.float64).reshape(-1, 2) * scale + np.array([l, b])
    pts = np.concatenate([free, border_uv])
    k0, m = len(free), len(border_uv)
    segs = np.array([[k0 + i, k0 + (i + 1) % m] for i in range(m)], np.int32)
    return np.ascontiguousarray(pts), segs, k0


def write_off(path, V, T):
    with open(path, "w") as f:
        f.write(f"OFF\n{len(V)} {len(T)} 0\n")
        for v in V:
            f.write(f"{v[0]:.17g} {v[1]:.17g} {v[2]:.17g}\n")
        for t in T:
            f.write(f"3 {t[0]} {t[1]} {t[2]}\n")


def mesh_stats(V, T):
    e = np.sort(np.concatenate([T[:, [0, 1]], T[:, [1, 2]], T[:, [2, 0]]]), axis=1)
    ue, cnt = np.unique(e, axis=0, return_counts=True)
    used = np.unique(T)
    a = 0.5 * np.linalg.norm(np.cross(V[T[:, 1]] - V[T[:, 0]], V[T[:, 2]] - V[T[:, 0]]), axis=1)
    return {"vertices_used": int(len(used)), "edges": int(len(ue)), "faces": int(len(T)),
            "euler": int(len(used) - len(ue) + len(T)), "border_edges": int((cnt == 1).sum()),
            "nonmanifold_edges": int((cnt > 2).sum()), "area": float(a.sum()),
            "mean_edge": float(np.linalg.norm(V[ue[:, 0]] - V[ue[:, 1]], axis=1).mean())}


def weld(T, num_free, cpoint_orig):
    """Merge the constraint points (CDT points num_free..) that are copies of one surface vertex — the two sides of the
    seam — and return the triangles with merged indices."""
    idx = np.arange(num_free + len(cpoint_orig))
    first = {}
    for i, o in enumerate(np.asarray(cpoint_orig).tolist()):
        idx[num_free + i] = first.setdefault(o, num_free + i)
    Tw = idx[T]
    # The border (one point per seam vertex and side) is sampled more densely than the interior (2000 sites), so the CDT
    # has "ears" (i, i+1, i+2) along it.  Where both sides of the seam have the same ear, welding makes them one
    # triangle twice — a flat pillow: drop both copies (their middle vertex then simply is not used any more).
    key = np.sort(Tw, axis=1)
    _, inv, cnt = np.unique(key, axis=0, return_inverse=True, return_counts=True)
    keep = (cnt[inv.reshape(-1)] == 1) & (key[:, 0] != key[:, 1]) & (key[:, 1] != key[:, 2])
    return Tw[keep]


def run(n=1024, sites=2000, iters=100, out_off=None, mesh="nefertiti"):
    z = np.load(os.path.join(ROOT, "tests", "golden", f"c1_{mesh}.npz"))
    V3, F, uv, loop, wt = z["V"], np.ascontiguousarray(z["F"], np.int32), z["uv"], z["loop"], np.ascontiguousarray(z["weights"])
    pts, scale, l, b = FE.discretization_arrays(uv, n)
    out = {"n": n, "sites": sites, "max_iter": iters, "mesh": {"V": len(V3), "F": len(F), "border": len(loop)}}
    t0 = time.time()
    dens = np.empty((n, n), np.float32)
    S.discretization_d(pts, wt, len(wt), F, len(F), dens, scale, n)
    mask = np.zeros((n, n), np.uint8)
    S.generateMask(uv[loop], mask, n, scale, l, b)
    vor = np.empty((n, n, 2), np.int16)
    st = S.centroidalVoronoi(vor, dens, mask, sites, n, 1, iters)
    out["gcvt"] = st
    cpts, segs, k0 = cdt_input(vor, mask, scale, l, b, uv[loop])
    out["t_hot_path_s"] = time.time() - t0
    # the same with the oracle
    edens = O.rasterise(pts, wt, F, scale, n)
    seeds, _, _ = O.seed(edens, mask, sites)
    exp, it, _, _ = O.gcvt(seeds, edens, mask, iters, stop_rule=1)
    epts, esegs, ek0 = cdt_input(exp, mask, scale, l, b, uv[loop])
    out["density_bit_exact"] = bool(np.array_equal(dens.view(np.uint32), edens.view(np.uint32)))
    out["labels_bit_exact"] = bool((vor != exp).sum() == 0 and st["iterations"] == it)
    out["cdt_input_identical"] = bool(np.array_equal(cpts, epts) and np.array_equal(segs, esegs))
    out["cdt_points"] = int(len(cpts)); out["cdt_segments"] = int(len(segs))
    t0 = time.time()
    tri = R.cdt(cpts, segs)
    out["t_cdt_s"] = time.time() - t0
    out["cdt_triangles"] = int(len(tri))
    t0 = time.time()
    verts, keep = S.recover(uv, V3, F, cpts, loop.astype(np.int32), tri)
    out["t_recover_s"] = time.time() - t0
    everts, ekeep, kept = O.recover(uv, V3, F, cpts, loop.astype(np.int32), tri)
    out["vertices_bit_exact"] = bool(np.array_equal(verts.view(np.uint64), everts.view(np.uint64)))
    out["kept_identical"] = bool(np.array_equal(keep, ekeep))
    T = tri[keep.astype(bool)]
    out["result"] = mesh_stats(verts, T)
    if "orig" in z.files:   # a mesh that was cut along a seam: weld the two sides, the surface must close again
        Tw = weld(T, k0, z["orig"][loop])
        out["result_welded"] = mesh_stats(verts, Tw)
        src = mesh_stats(z["V_src"], z["F_src"])
        out["source"] = {"euler": src["euler"], "border_edges": src["border_edges"]}
        T = Tw
    a3 = 0.5 * np.linalg.norm(np.cross(V3[F[:, 1]] - V3[F[:, 0]], V3[F[:, 2]] - V3[F[:, 0]]), axis=1).sum()
    out["source_area"] = float(a3)
    ext = float((V3.max(0) - V3.min(0)).max())
    out["max_vertex_diff_over_extent"] = float(np.abs(verts - everts).max() / ext)
    if out_off:
        write_off(out_off, verts, T)
    return out, verts, T


if __name__ == "__main__":
    a = sys.argv[1:]
    res, _, _ = run(int(a[0]) if len(a) > 0 else 1024, int(a[1]) if len(a) > 1 else 2000, int(a[2]) if len(a) > 2 else 100,
                    a[3] if len(a) > 3 else None, a[4] if len(a) > 4 else "nefertiti")
    print(json.dumps(res))
