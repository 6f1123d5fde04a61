"""CGAL-free host front end for BASELINE config 1: bring a bundled disk-topology mesh to the hot-path
boundary with the reference's array contract.

The reference reaches `discretization_d` / `gCVT` through CGAL (seam cutting, ARAP parameterisation, main.cpp:141-216).
None of that exists in this environment, and none of it is on the hot path; this module provides the minimum that
lets an OFF mesh with one boundary loop go through the same boundary:

    read_off -> boundary_loop -> tutte_parameterize (uniform-weight harmonic map onto the unit circle; ARAP is
    CGAL-only) -> area_ratio_weights (weighting.h:38-97) -> discretization_arrays (discretization.h:89-118:
    bbox, scale = max extent / (n-1), points shifted by the bbox corner) -> constraint points = border vertices
    (constrains.h:61-76) -> srm_generate_mask / srm_discretize / srm_seed / srm_gcvt.

Closed meshes (horse.off, pig.off) are first cut to a disk along the seam of their selection file, as main.cpp:157-168
does through CGAL's Seam_mesh: read_seam_pairs (split.h:28-57 addSeams) -> split_long_edges on the seam edges
(main.cpp:165, PMP::split_long_edges with the 0.72 limit of main.cpp:150; the border edges of main.cpp:153-155 likewise)
-> cut_along_seam (what Seam_mesh does virtually: every seam vertex is duplicated once per sector of its face fan
between two seam edges, so the seam opens into a border loop of twice its edges).  `orig` maps the vertices of the cut
mesh back to the surface vertices, which is how the two sides of the seam are welded after the lift.
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


def read_seam_pairs(path):
    """split.h:40-43: pairs of vertex indices, one seam edge per line."""
    toks = open(path).read().split()
    return [(int(toks[i]), int(toks[i + 1])) for i in range(0, len(toks) - 1, 2)]


def _edge_key(a, b):
    return (a, b) if a < b else (b, a)


def seam_edges_of(F, pairs):
    """split.h:44-55: the listed pairs that are edges of the mesh and not border edges, without repetitions."""
    cnt = {}
    for a, b, c in np.asarray(F).tolist():
        for u, v in ((a, b), (b, c), (c, a)):
            cnt[_edge_key(u, v)] = cnt.get(_edge_key(u, v), 0) + 1
    out, seen = [], set()
    for s, t in pairs:
        k = _edge_key(s, t)
        if cnt.get(k, 0) == 2 and k not in seen:     # an edge, and not on the border
            seen.add(k); out.append(k)
    return out


def split_long_edges(V, F, edges, limit):
    """PMP::split_long_edges (main.cpp:155,165) on a set of edges: while one of them is longer than `limit`, split the
    longest at its midpoint (both incident triangles are split through the new vertex); the two halves stay in the set.
    Returns (V, F, edges) of the refined mesh."""
    import heapq
    V = [np.asarray(p, np.float64) for p in np.asarray(V, np.float64)]
    F = [list(map(int, f)) for f in np.asarray(F).tolist()]
    inc = {}
    for fi, (a, b, c) in enumerate(F):
        for u, v in ((a, b), (b, c), (c, a)):
            inc.setdefault(_edge_key(u, v), set()).add(fi)
    live = set(_edge_key(a, b) for a, b in edges)
    heap = [(-float(np.linalg.norm(V[a] - V[b])), a, b) for a, b in live]
    heapq.heapify(heap)
    while heap:
        negl, a, b = heapq.heappop(heap)
        k = _edge_key(a, b)
        if k not in live or -negl <= limit:
            continue
        m = len(V)
        V.append(0.5 * (V[a] + V[b]))
        live.discard(k)
        for fi in sorted(inc.pop(k, ())):
            f = F[fi]
            i = [j for j in range(3) if _edge_key(f[j], f[(j + 1) % 3]) == k][0]
            p, q, r = f[i], f[(i + 1) % 3], f[(i + 2) % 3]          # edge p-q, apex r
            F[fi] = [p, m, r]
            F.append([m, q, r])
            fj = len(F) - 1
            inc[_edge_key(q, r)].discard(fi); inc[_edge_key(q, r)].add(fj)
            inc.setdefault(_edge_key(p, m), set()).add(fi)
            inc.setdefault(_edge_key(m, q), set()).add(fj)
            inc.setdefault(_edge_key(m, r), set()).update((fi, fj))
        for u in (a, b):
            kk = _edge_key(u, m)
            live.add(kk)
            heapq.heappush(heap, (-float(np.linalg.norm(V[u] - V[m])), kk[0], kk[1]))
    return np.asarray(V), np.asarray(F, np.int32), sorted(live)


def cut_along_seam(V, F, seam):
    """Open a mesh along its seam edges (the explicit form of CGAL's Seam_mesh, main.cpp:157-168).  Around a vertex the
    incident faces fall into sectors separated by seam edges; every sector gets its own copy of the vertex (the first
    keeps the index).  Returns (V2, F2, orig) with orig[new vertex] = vertex of the uncut mesh."""
    V = np.asarray(V, np.float64); F = np.asarray(F, np.int32)
    seam = set(_edge_key(a, b) for a, b in seam)
    on_seam = set(v for e in seam for v in e)
    faces_of = {}
    for fi, f in enumerate(F.tolist()):
        for v in f:
            if v in on_seam:
                faces_of.setdefault(v, []).append(fi)
    F2 = F.copy()
    Vout, orig = [p for p in V], list(range(len(V)))
    for v, fs in faces_of.items():
        parent = {fi: fi for fi in fs}

        def find(x):
            while parent[x] != x:
                parent[x] = parent[parent[x]]; x = parent[x]
            return x
        by_edge = {}
        for fi in fs:
            for w in F[fi].tolist():
                if w != v:
                    by_edge.setdefault(w, []).append(fi)
        for w, pair in by_edge.items():
            if len(pair) == 2 and _edge_key(v, w) not in seam:   # the two faces share a non-seam edge at v: same sector
                parent[find(pair[0])] = find(pair[1])
        roots = sorted(set(find(fi) for fi in fs))
        for root in roots[1:]:
            nv = len(Vout)
            Vout.append(V[v]); orig.append(v)
            for fi in fs:
                if find(fi) == root:
                    F2[fi][F2[fi] == v] = nv
    return np.asarray(Vout), F2, np.asarray(orig, np.int64)


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


def prepare(V, F, n, seam_pairs=None, edge_length_limit=0.72):
    """Everything the boundary needs for one mesh: dict(points, weights, triangles, scale, left, lower, border_uv, ...).
    With seam_pairs (a closed mesh and its selection file) the mesh is refined and cut first, main.cpp:150-168; the
    result then carries V / F of the cut mesh and `orig`."""
    V = np.asarray(V, np.float64); F = np.asarray(F, np.int32)
    orig = np.arange(len(V))
    if seam_pairs is not None:
        seam = seam_edges_of(F, seam_pairs)
        V, F, seam = split_long_edges(V, F, seam, edge_length_limit)
        V, F, orig = cut_along_seam(V, F, seam)
    loop = boundary_loop(F)
    UV = tutte_parameterize(V, F, loop)
    wt = area_ratio_weights(V, F, UV)
    pts, scale, l, b = discretization_arrays(UV, n)
    return {"points": pts, "weights": np.ascontiguousarray(wt), "triangles": np.ascontiguousarray(F, np.int32),
            "scale": scale, "left": l, "lower": b, "border_uv": np.ascontiguousarray(UV[loop]), "uv": UV, "loop": loop,
            "V": V, "F": np.ascontiguousarray(F, np.int32), "orig": orig}
