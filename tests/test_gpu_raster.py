"""Rasteriser parity: CUDA (tile-binned) vs the CPU oracle and vs the reference CUDA golden — bit-exact floats."""
import glob
import os

import numpy as np
import pytest

import _inputs as I
import _oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _run(pts, wt, tri, scale, n):
    import surface_remesher_b200 as S
    out = np.empty((n, n), np.float32)
    S.discretization_d(np.ascontiguousarray(pts, np.float64), np.ascontiguousarray(wt, np.float64), len(wt),
                       np.ascontiguousarray(tri, np.int32).reshape(-1, 3), len(tri), out, scale, n)
    return out


@pytest.mark.parametrize("side,n,seed", [(6, 256, 1), (12, 256, 2), (40, 512, 3), (90, 1024, 4)])
def test_raster_random_mesh(side, n, seed):
    pts, wt, tri = I.random_mesh(side, seed)
    scale = 1.0 / (n - 1)
    got = _run(pts, wt, tri, scale, n)
    exp = O.rasterise(pts, wt, tri, scale, n)
    assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))
    assert (got > 0).mean() > 0.95


def test_raster_partial_coverage_degenerate_and_overlap():
    n = 256
    pts = np.array([[0.2, 0.2], [0.8, 0.25], [0.5, 0.9], [0.5, 0.5], [0.5, 0.5], [0.6, 0.6], [-0.5, -0.5], [3.0, -0.5],
                    [-0.5, 3.0], [0.1, 0.1], [0.3, 0.1], [0.2, 0.1]])
    wt = np.arange(1, len(pts) + 1, dtype=np.float64)
    tri = np.array([[3, 4, 5],      # degenerate (two identical vertices): never hits
                    [9, 10, 11],    # collinear: never hits
                    [0, 1, 2],      # ordinary
                    [6, 7, 8],      # huge triangle covering the grid, later in index order
                    [0, 2, 1]], np.int32)
    scale = 1.0 / (n - 1)
    got = _run(pts, wt, tri, scale, n)
    exp = O.rasterise(pts, wt, tri, scale, n)
    assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))
    assert (got != 0).all()      # the huge triangle covers everything


def test_raster_no_triangles_is_zero():
    n = 256
    pts = np.zeros((3, 2)); wt = np.ones(3)
    got = _run(pts, wt, np.zeros((0, 3), np.int32), 1.0 / (n - 1), n)
    assert (got == 0).all()


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(G, "ref_raster_*.npz"))), ids=os.path.basename)
def test_raster_vs_reference_golden(path):
    z = np.load(path)
    n = int(z["n"])
    pts, wt, tri = I.random_mesh(int(z["side"]), int(z["seed"]))
    got = _run(pts, wt, tri, float(z["scale"]), n)
    r0, r1 = [int(v) for v in z["rows"]]
    assert np.array_equal(got[r0:r1].view(np.uint32), z["density"].view(np.uint32))
