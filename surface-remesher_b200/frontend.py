"""CGAL-free host front end for BASELINE config 1: bring a bundled disk-topology mesh to the hot-path
boundary with the reference's array contract.

The reference reaches `discretization_d` / `gCVT` through CGAL (seam cutting, ARAP parameterisation, main.cpp:141-216).
None of that exists in this environment, and none of it is on the hot path; this module provides the minimum that
lets an OFF mesh with one boundary loop go through the same boundary:

    read_off -> boundary_loop -> tutte_parameterize (uniform-weight harmonic map onto the unit circle; ARAP is
    CGAL-only) -> area_ratio_weights (weighting.h:38-97) -> discretization_arrays (discretization.h:89-118:
    bbox, scale = max extent / (n-1), points shifted by the bbox corner) -> constraint points = border vertices
    (constrains.h:61-76) -> srm_generate_mask / srm_discretize / srm_seed / srm_gcvt.

Closed meshes need a seam cut first (split.h); long-edge splitting of the border (main.cpp:153-155) is not done.
"""
import numpy as np


def read_off(path):
    toks = open(path).read().split()
    assert toks[0] == "OFF"
    nv, nf = int(toks[1]), int(toks[2])
    p = 4
    V = np.array(toks[p:p + 3 * nv], np.float64).reshape(nv, 3)
    p += 3 * nv
    F = []
    for _ in range(nf):
        k = int(toks[p]); assert k == 3, "triangle meshes only"
        F.append([int(toks[p + 1]), int(toks[p + 2]), int(toks[p + 3])])
        p += 4
    return V, np.asarray(F, np.int32)


def boundary_loop(F):
    """Ordered vertex loop of the single boundary (edges used by exactly one triangle)."""
    cnt = {}
    for a, b, c in F.tolist():
        for u, v in ((a, b), (b, c), (c, a)):
            cnt[(u, v)] = cnt.get((u, v), 0) + 1
    nxt = {u: v for (u, v), k in cnt.items() if (v, u) not in cnt}
    if not nxt:
        raise ValueError("closed mesh: cut it along a seam first")
    start = next(iter(nxt))
    loop, cur = [start], nxt[start]
    while cur != start:
        loop.append(cur); cur = nxt[cur]
    if len(loop) != len(nxt):
        raise ValueError("more than one boundary loop")
    return np.asarray(loop, np.int64)


def tutte_parameterize(V, F, loop):
    """Boundary on the unit circle by chord length, interior vertices at the average of their neighbours (Tutte)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    nv = len(V)
    seg = np.linalg.norm(V[np.roll(loop, -1)] - V[loop], axis=1)
    t = 2 * np.pi * np.concatenate([[0], np.cumsum(seg)[:-1]]) / seg.sum()
    UV = np.zeros((nv, 2))
    UV[loop, 0], UV[loop, 1] = np.cos(t), np.sin(t)
    I = np.concatenate([F[:, 0], F[:, 1], F[:, 2], F[:, 1], F[:, 2], F[:, 0]])
    J = np.concatenate([F[:, 1], F[:, 2], F[:, 0], F[:, 0], F[:, 1], F[:, 2]])
    A = sp.coo_matrix((np.ones(len(I)), (I, J)), shape=(nv, nv)).tocsr()
    A.data[:] = 1.0                                   # uniform weights (duplicates collapse to 1)
    Lap = sp.diags(np.asarray(A.sum(1)).ravel()) - A
    interior = np.setdiff1d(np.arange(nv), loop)
    rhs = -Lap[interior][:, loop] @ UV[loop]
    UV[interior] = spl.splu(Lap[interior][:, interior].tocsc()).solve(rhs)
    return UV


def area_ratio_weights(V, F, UV):
    """weighting_vertices_with_area (weighting.h:38-97): per vertex, sum of incident 3-D face areas / sum of incident
    2-D face areas; 'infinite' ratios (> 1e10) are replaced by the maximum finite one."""
    a3 = 0.5 * np.linalg.norm(np.cross(V[F[:, 1]] - V[F[:, 0]], V[F[:, 2]] - V[F[:, 0]]), axis=1)
    e1, e2 = UV[F[:, 1]] - UV[F[:, 0]], UV[F[:, 2]] - UV[F[:, 0]]
    a2 = 0.5 * np.abs(e1[:, 0] * e2[:, 1] - e2[:, 0] * e1[:, 1])
    A3, A2 = np.zeros(len(V)), np.zeros(len(V))
    for k in range(3):
        np.add.at(A3, F[:, k], a3); np.add.at(A2, F[:, k], a2)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = A3 / A2
    bad = ~(w <= 1e10)
    w[bad] = w[~bad].max()
    return w


def discretization_arrays(UV, n):
    """discretization.h:89-118: bbox corner (left, lower), scale, points shifted to the corner."""
    l, b = UV[:, 0].min(), UV[:, 1].min()
    r, u = UV[:, 0].max(), UV[:, 1].max()
    scale = max(u - b, r - l) / (n - 1)
    pts = np.ascontiguousarray(UV - np.array([l, b]))
    return pts, float(scale), float(l), float(b)


def prepare(V, F, n):
    """Everything the boundary needs for one mesh: dict(points, weights, triangles, scale, left, lower, border_uv)."""
    loop = boundary_loop(F)
    UV = tutte_parameterize(V, F, loop)
    wt = area_ratio_weights(V, F, UV)
    pts, scale, l, b = discretization_arrays(UV, n)
    return {"points": pts, "weights": np.ascontiguousarray(wt), "triangles": np.ascontiguousarray(F, np.int32),
            "scale": scale, "left": l, "lower": b, "border_uv": np.ascontiguousarray(UV[loop]), "uv": UV, "loop": loop}
