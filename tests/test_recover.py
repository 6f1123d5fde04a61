"""Point location and the lift back to 3-D (recover.h:63-153).

CPU: the oracle's brute-force `locate` / `recover` against plain numpy statements.
GPU: srm_locate / srm_recover through the C ABI against the oracle — face ids, barycentric weights, lifted
vertices and kept-triangle flags bit-identical (the product evaluates only a grid cell's face list with the
same predicate; the oracle tests every face like the reference)."""
import numpy as np
import pytest

import _inputs as I
import _oracle as O


def _surface(pts, seed):
    """A 3-D embedding of the 2-D mesh vertices (what vertex_2d_to_3d + the surface's point map give)."""
    rng = np.random.default_rng(seed)
    a, b, c = rng.random(3) * 3
    x, y = pts[:, 0], pts[:, 1]
    return np.ascontiguousarray(np.stack([x + 0.1 * np.sin(a * y), y * np.cos(b * x), np.sin(c * x) * np.cos(a * y)], 1))


def _cdt_like(n_free, n_cp, pts, seed):
    """CDT-input-like arrays: free points inside the unit square (a few outside), constraint points = mesh vertices,
    triangles over them from scipy's Delaunay (only their centroids matter to recover)."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    free = rng.random((n_free, 2))
    free[0] = [0.31, 0.47]                               # the first site must lie in a face
    if n_free > 10:
        free[5] = [1.7, 0.2]; free[9] = [-0.3, 0.5]      # sites in no face: the stale-f_loc path (recover.h:92-96)
    cpv = rng.choice(len(pts), size=n_cp, replace=False).astype(np.int32)
    extra = np.array([[1.6, 1.6], [-0.5, 1.5]])          # far-away points so that some triangle centroids fall outside
    allp = np.concatenate([free, pts[cpv]])
    tri = Delaunay(np.concatenate([allp, extra])).simplices.astype(np.int32)
    tri = tri[(tri < len(allp)).all(axis=1)]
    return np.ascontiguousarray(allp), cpv, np.ascontiguousarray(tri)


def test_oracle_locate_matches_numpy_statement():
    pts, wt, tri = I.random_mesh(9, 3)
    rng = np.random.default_rng(0)
    q = rng.random((300, 2)) * 1.2 - 0.1
    face, w = O.locate(pts, tri, q)
    for k in range(len(q)):
        exp = -1
        for t, (a, b, c) in enumerate(tri):
            v0, v1, v2 = pts[b] - pts[a], pts[c] - pts[a], q[k] - pts[a]
            d00, d01, d11, d20, d21 = v0 @ v0, v0 @ v1, v1 @ v1, v2 @ v0, v2 @ v1
            den = d00 * d11 - d01 * d01
            if den == 0:
                continue
            w1 = (d11 * d20 - d01 * d21) / den; w2 = (d00 * d21 - d01 * d20) / den; w3 = 1.0 - w1 - w2
            if w1 >= 0 and w2 >= 0 and w3 >= 0:
                exp = t
                assert np.allclose(w[k], [w3, w1, w2], rtol=0, atol=1e-12)
                break
        assert face[k] == exp
    inside = (q >= 0).all(1) & (q <= 1).all(1)
    assert (face[inside] >= 0).all() and (face[~inside] == -1).sum() > 0


def test_oracle_recover_lift_and_filter():
    pts, wt, tri = I.random_mesh(12, 4)
    p3 = _surface(pts, 1)
    allp, cpv, cdt = _cdt_like(200, 20, pts, 2)
    out, keep, kept = O.recover(pts, p3, tri, allp, cpv, cdt)
    face, w = O.locate(pts, tri, allp[:200])
    ok = np.nonzero(face >= 0)[0]
    exp = np.einsum("kj,kjd->kd", w[ok], p3[tri[face[ok]]])
    assert np.allclose(out[ok], exp, rtol=0, atol=1e-12)
    assert np.array_equal(out[5], out[4]) and np.array_equal(out[9], out[8])      # stale f_loc: previous vertex again
    assert np.array_equal(out[200:], p3[cpv])
    assert kept == keep.sum() and 0 < kept < len(cdt)
    cen = allp[cdt].sum(1) / 3.0
    inside = (cen >= 0).all(1) & (cen <= 1).all(1)
    assert np.array_equal(keep.astype(bool), inside)


@pytest.mark.gpu
@pytest.mark.parametrize("side,nq,seed", [(4, 500, 1), (12, 5000, 2), (60, 40000, 3), (150, 100000, 4)])
def test_locate_bit_exact(side, nq, seed):
    import surface_remesher_b200 as S
    pts, wt, tri = I.random_mesh(side, seed)
    rng = np.random.default_rng(seed)
    q = rng.random((nq, 2)) * 1.1 - 0.05
    q[: len(pts)] = pts[: nq][: len(pts)]            # mesh vertices themselves (on edges of several faces: first wins)
    mid = 0.5 * (pts[tri[:, 0]] + pts[tri[:, 1]])    # edge midpoints
    q[len(pts): len(pts) + len(mid)] = mid[: max(0, nq - len(pts))][: len(q[len(pts): len(pts) + len(mid)])]
    face, w = S.locate(pts, tri, q)
    eface, ew = O.locate(pts, tri, q)
    assert np.array_equal(face, eface)
    hit = face >= 0
    assert np.array_equal(w[hit].view(np.uint64), ew[hit].view(np.uint64))
    assert hit.mean() > 0.7


@pytest.mark.gpu
def test_locate_degenerate_faces_and_empty():
    import surface_remesher_b200 as S
    pts = np.array([[0.2, 0.2], [0.8, 0.25], [0.5, 0.9], [0.5, 0.5], [0.5, 0.5], [0.6, 0.6], [0.1, 0.1], [0.3, 0.1],
                    [0.2, 0.1], [0.0, 0.0], [1.0, 0.0], [0.0, 1.0]])
    tri = np.array([[3, 4, 5], [6, 7, 8], [0, 1, 2], [9, 10, 11]], np.int32)   # degenerate, collinear, two overlapping
    q = np.array([[0.5, 0.5], [0.2, 0.1], [0.1, 0.15], [0.9, 0.9], [0.5, 0.4]])
    face, w = S.locate(pts, tri, q)
    eface, ew = O.locate(pts, tri, q)
    assert np.array_equal(face, eface) and face.tolist() == [2, 3, 3, -1, 2]
    f0, _ = S.locate(pts, tri, np.zeros((0, 2)))
    assert len(f0) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("side,nfree,ncp,seed", [(12, 200, 20, 2), (80, 20000, 300, 5)])
def test_recover_bit_exact(side, nfree, ncp, seed):
    import surface_remesher_b200 as S
    pts, wt, tri = I.random_mesh(side, seed)
    p3 = _surface(pts, seed)
    allp, cpv, cdt = _cdt_like(nfree, ncp, pts, seed + 1)
    out, keep = S.recover(pts, p3, tri, allp, cpv, cdt)
    eout, ekeep, kept = O.recover(pts, p3, tri, allp, cpv, cdt)
    assert np.array_equal(out.view(np.uint64), eout.view(np.uint64))
    assert np.array_equal(keep, ekeep) and keep.sum() == kept
