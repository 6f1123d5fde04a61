"""BASELINE config 1: a bundled mesh (nefertiti, via the CGAL-free front end) -> rasterise -> mask -> seed -> gCVT at
1024^2 / 2k sites / 100 iterations -> CDT input points.  CPU part: front-end sanity; GPU part: parity with the oracle."""
import os

import numpy as np
import pytest

import _inputs as I
import _oracle as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _fixture(n):
    from surface_remesher_b200 import frontend as FE
    z = np.load(os.path.join(G, "c1_nefertiti.npz"))
    pts, scale, l, b = FE.discretization_arrays(z["uv"], n)
    return z, pts, scale, l, b


def test_frontend_on_synthetic_disk():
    """Tutte map of a curved disk: bijective (all 2-D triangles keep their orientation), weights finite and positive."""
    from surface_remesher_b200 import frontend as FE
    pts, _, tri = I.random_mesh(9, 3)
    V = np.c_[pts, 0.3 * np.sin(3 * pts[:, 0]) * np.cos(2 * pts[:, 1])]
    loop = FE.boundary_loop(tri)
    assert len(loop) == 4 * 8
    UV = FE.tutte_parameterize(V, tri, loop)
    e1, e2 = UV[tri[:, 1]] - UV[tri[:, 0]], UV[tri[:, 2]] - UV[tri[:, 0]]
    area = e1[:, 0] * e2[:, 1] - e2[:, 0] * e1[:, 1]
    assert (np.sign(area) == np.sign(area[0])).all() and (np.abs(area) > 1e-12).all()
    w = FE.area_ratio_weights(V, tri, UV)
    assert np.isfinite(w).all() and (w > 0).all()
    with pytest.raises(ValueError):
        FE.boundary_loop(np.array([[0, 1, 2], [0, 2, 3], [0, 3, 1], [1, 3, 2]], np.int32))  # closed tetrahedron


def test_config1_fixture_is_a_valid_parameterisation():
    z, pts, scale, l, b = _fixture(1024)
    F, uv = z["F"], z["uv"]
    e1, e2 = uv[F[:, 1]] - uv[F[:, 0]], uv[F[:, 2]] - uv[F[:, 0]]
    area = e1[:, 0] * e2[:, 1] - e2[:, 0] * e1[:, 1]
    assert (np.sign(area) == np.sign(area[0])).all()
    assert abs(np.linalg.norm(uv[z["loop"]], axis=1) - 1).max() < 1e-12
    assert pts.min() == 0 and abs(pts.max() / scale - 1023) < 1e-6


@pytest.mark.gpu
def test_config1_end_to_end_parity():
    import surface_remesher_b200 as S
    n, sites, iters = 1024, 2000, 100
    z, pts, scale, l, b = _fixture(n)
    wt, tri = np.ascontiguousarray(z["weights"]), np.ascontiguousarray(z["F"], np.int32)
    dens = np.empty((n, n), np.float32)
    S.discretization_d(pts, wt, len(wt), tri, len(tri), dens, scale, n)
    assert np.array_equal(dens.view(np.uint32), O.rasterise(pts, wt, tri, scale, n).view(np.uint32))
    mask = np.zeros((n, n), np.uint8)
    S.generateMask(z["uv"][z["loop"]], mask, n, scale, l, b)          # constraint points = border vertices
    assert 0 < mask.sum() <= len(z["loop"])
    vor = np.empty((n, n, 2), np.int16)
    st = S.centroidalVoronoi(vor, dens, mask, sites, n, 1, iters)      # putConstrains + randomPoints + gCVT
    seeds, _, _ = O.seed(dens, mask, sites)
    exp, it, _, _ = O.gcvt(seeds, dens, mask, iters, stop_rule=1)
    assert st["iterations"] == it
    assert (vor != exp).sum() == 0
    # CDT input points in delaunayInput order, from a resident context with the same state
    with S.Context(n) as c:
        c.set_density(dens); c.set_mask(mask); c.set_site_map(seeds)
        c.run(iters, stop_rule=True)
        p = c.extract_sites(mask, scale, l, b)
    free = [(x, y) for (x, y) in sorted(I.site_set(exp)) if not mask[y, x]]
    assert len(p) == len(free) and np.allclose(p, np.array(free) * scale + np.array([l, b]))


@pytest.mark.gpu
def test_config1_full_pipeline_to_remeshed_surface():
    """The whole reference pipeline behind the CGAL-free front end: rasterise -> gCVT -> delaunayInput -> constrained
    Delaunay by the UNMODIFIED reference gDel2D (test infrastructure, oracle/_ref/libgdel2d_ref.so, in a subprocess) ->
    recover.  Product (libsrm) and oracle must agree on every intermediate and on the final vertices bit for bit
    (north_star: final-mesh vertices within 1e-5 of the extent); the result must be a disk like the source."""
    import sys
    import _ref as R
    if not R.cdt_available():
        pytest.skip("oracle/_ref/libgdel2d_ref.so not built")
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import run_config1 as RC
    try:
        out, verts, T = RC.run(1024, 2000, 100)
    except RuntimeError as e:   # the decade-old reference CDT is not the product: report, do not fail
        pytest.skip(str(e))
    assert out["density_bit_exact"] and out["labels_bit_exact"] and out["cdt_input_identical"]
    assert out["vertices_bit_exact"] and out["kept_identical"] and out["max_vertex_diff_over_extent"] <= 1e-5
    r = out["result"]
    assert r["faces"] > 2000 and r["nonmanifold_edges"] == 0
    assert r["euler"] == 1                                   # a disk, like nefertiti.off
    assert abs(r["area"] - out["source_area"]) / out["source_area"] < 0.05
