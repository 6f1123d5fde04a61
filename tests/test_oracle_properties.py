"""Randomised properties of the CPU oracle (hypothesis; CPU only).  The oracle is the checker of every GPU parity
test, so its own statements are cross-checked against each other on inputs that stress the tie rules."""
import numpy as np
from hypothesis import given, settings, strategies as st

import _inputs as I
import _oracle as O


def _map(n, xs, ys):
    v = np.full((n, n, 2), I.MARK, np.int16)
    v[ys, xs, 0] = xs
    v[ys, xs, 1] = ys
    return v


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 2**31 - 1), st.sampled_from([2, 3, 4, 6, 8]), st.floats(0.05, 1.0))
def test_separable_labelling_equals_brute_force_on_lattices(seed, pitch, keep):
    """Lattice subsets produce exact distance ties in both directions (column rule incl. the 64-row band, row rule)."""
    n = 128
    rng = np.random.default_rng(seed)
    gx, gy = np.meshgrid(np.arange(rng.integers(0, pitch), n, pitch), np.arange(rng.integers(0, pitch), n, pitch))
    sel = rng.random(gx.size) < keep
    if not sel.any():
        sel[rng.integers(0, gx.size)] = True
    seeds = _map(n, gx.ravel()[sel], gy.ravel()[sel])
    assert (O.label_brute(seeds) != O.label_exact(seeds)).sum() == 0


@settings(max_examples=20, deadline=None)
@given(st.integers(0, 2**31 - 1), st.integers(1, 400))
def test_labels_are_nearest_sites_and_sites_label_themselves(seed, k):
    n = 96
    seeds = I.random_sites(n, k, seed)
    lab = O.label_exact(seeds).astype(np.int64)
    sites = np.array(sorted(I.site_set(seeds)), np.int64)
    yy, xx = np.mgrid[0:n, 0:n]
    d_lab = (lab[..., 0] - xx) ** 2 + (lab[..., 1] - yy) ** 2
    d_min = ((sites[:, 0][None, None, :] - xx[..., None]) ** 2 + (sites[:, 1][None, None, :] - yy[..., None]) ** 2).min(-1)
    assert np.array_equal(d_lab, d_min)                                   # exact Euclidean Voronoi
    assert np.array_equal(lab[sites[:, 1], sites[:, 0]], sites)            # a site owns its pixel
    assert I.site_set(lab) == I.site_set(seeds)                            # self-labelled pixels are exactly the sites


@settings(max_examples=20, deadline=None)
@given(st.integers(0, 2**31 - 1), st.integers(3, 14))
def test_locate_reconstructs_interior_points(seed, side):
    pts, _, tri = I.random_mesh(side, seed % 1000)
    rng = np.random.default_rng(seed)
    f = rng.integers(0, len(tri), 200)
    w = rng.dirichlet([2, 2, 2], 200)                                      # strictly interior
    q = np.einsum("kj,kjd->kd", w, pts[tri[f]])
    face, ww = O.locate(pts, tri, q)
    assert (face >= 0).all()
    rec = np.einsum("kj,kjd->kd", ww, pts[tri[face]])
    assert np.allclose(rec, q, rtol=0, atol=1e-12)
    assert (ww >= 0).all() and np.allclose(ww.sum(1), 1.0, atol=1e-12)
    assert np.array_equal(face, f)                                         # interior points: the face is unique


@settings(max_examples=10, deadline=None)
@given(st.integers(0, 2**31 - 1))
def test_pyramid_preserves_mass_and_zoom_roundtrips(seed):
    rng = np.random.default_rng(seed)
    d = rng.random((512, 512)).astype(np.float32)
    d1 = O.density_scale(d)
    assert abs(float(d1.astype(np.float64).sum()) * 4 - float(d.astype(np.float64).sum())) < 1e-3 * d.size * 1e-3
    seeds = I.random_sites(256, int(rng.integers(1, 500)), int(rng.integers(0, 10**6)))
    z = O.zoom_in(seeds)
    back = {(x // 2, y // 2) for (x, y) in I.site_set(z)}
    assert back == I.site_set(seeds) and all(x % 2 == 0 and y % 2 == 0 for (x, y) in I.site_set(z))


@settings(max_examples=8, deadline=None)
@given(st.integers(0, 2**31 - 1))
def test_lloyd_step_conserves_mass_and_respects_constraints(seed):
    """Centroid sums add up to the total density; constrained sites do not move; free sites land on positive density."""
    n = 128
    rng = np.random.default_rng(seed)
    dens = (rng.random((n, n)) * (rng.random((n, n)) > 0.2)).astype(np.float32)
    mask = (rng.random((n, n)) > 0.995).astype(np.uint8)
    seeds, _, _ = O.seed(dens, mask, 60)
    lab, nxt, e = O.lloyd_step(seeds, dens, mask, 2.0)
    W, X, Y = O.centroid(lab, dens)
    assert abs(W.sum() - dens.astype(np.float64).sum()) <= 1e-9 * max(1.0, float(dens.sum()))
    fixed = {(x, y) for (x, y) in I.site_set(seeds) if mask[y, x]}
    assert fixed <= I.site_set(nxt)
    for (x, y) in I.site_set(nxt) - fixed - I.site_set(seeds):
        assert dens[y, x] > 0                                              # a moved free site never enters zero density
    assert e >= 0
