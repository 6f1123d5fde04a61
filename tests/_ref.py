"""ctypes binding of oracle/_ref/libsrm_ref.so — the UNMODIFIED reference CUDA path (gcvt.cu +
discretization.cu compiled from /root/reference by oracle/Makefile) behind oracle/ref_driver.cu.
Test / baseline infrastructure only; needs a GPU."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "oracle", "_ref", "libsrm_ref.so")
_lib = None


def available():
    return os.path.exists(PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(PATH)
        p, i, f = C.c_void_p, C.c_int, C.c_float
        L.ref_gcvt.argtypes = [p, p, p, i, i, i]
        L.ref_gcvt_timed.argtypes = [p, p, p, i, i, i, p]
        L.ref_label.argtypes = [p, i]
        L.ref_step.argtypes = [p, p, p, i, f, p, p, p]
        L.ref_loop_timed.argtypes = [p, p, p, i, i, p]
        L.ref_pyramid.argtypes = [p, p, i, i, p]
        L.ref_zoom.argtypes = [p, i, p]
        L.ref_discretize.argtypes = [p, p, i, p, i, p, C.c_double, i]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _mask(mask, n):
    return np.zeros((n, n), np.uint8) if mask is None else np.ascontiguousarray(mask, np.uint8)


def label(seeds):
    n = seeds.shape[0]
    v = np.array(seeds, np.int16, copy=True, order="C")
    rc = lib().ref_label(_p(v), n)
    assert rc == 0, rc
    return v


def step(seeds, density, mask, omega):
    n = seeds.shape[0]
    s = np.ascontiguousarray(seeds, np.int16)
    d = np.ascontiguousarray(density, np.float32)
    m = _mask(mask, n)
    lab = np.empty_like(s); out = np.empty_like(s); e = np.zeros(1, np.float32)
    rc = lib().ref_step(_p(s), _p(d), _p(m), n, float(omega), _p(lab), _p(out), _p(e))
    assert rc == 0, rc
    return lab, out, float(e[0])


def gcvt(seeds, density, mask, max_iter, timed=False):
    n = seeds.shape[0]
    v = np.array(seeds, np.int16, copy=True, order="C")
    d = np.ascontiguousarray(density, np.float32)
    m = _mask(mask, n)
    if timed:
        ms = np.zeros(1, np.float32)
        it = lib().ref_gcvt_timed(_p(v), _p(d), _p(m), n, 1, int(max_iter), _p(ms))
        return v, it, float(ms[0])
    it = lib().ref_gcvt(_p(v), _p(d), _p(m), n, 1, int(max_iter))
    return v, it


def gcvt_multires(coarse_seeds, density, mask, depth, max_iter):
    """Reference gCVT with depth > 1 (gcvt.cu:1087-1156): the seed map has side n >> (depth-1) and is read from the
    first entries of the n^2 buffer.  Returns (final n^2 label map, iterations)."""
    n = density.shape[0]
    s = coarse_seeds.shape[0]
    assert s == n >> (depth - 1)
    v = np.full((n, n, 2), -32768, np.int16)
    v.reshape(-1)[: 2 * s * s] = np.ascontiguousarray(coarse_seeds, np.int16).reshape(-1)
    d = np.ascontiguousarray(density, np.float32)
    m = _mask(mask, n)
    it = lib().ref_gcvt(_p(v), _p(d), _p(m), n, int(depth), int(max_iter))
    return v, it


def pyramid(density, level):
    n = density.shape[0]
    d = np.ascontiguousarray(density, np.float32)
    m = _mask(None, n)
    s = n >> level
    out = np.empty((s, s), np.float32)
    rc = lib().ref_pyramid(_p(d), _p(m), n, int(level), _p(out))
    assert rc == 0, rc
    return out


def zoom(seeds):
    s = seeds.shape[0]
    a = np.ascontiguousarray(seeds, np.int16)
    out = np.empty((2 * s, 2 * s, 2), np.int16)
    rc = lib().ref_zoom(_p(a), s, _p(out))
    assert rc == 0, rc
    return out


def loop_timed(seeds, density, mask, iters):
    n = seeds.shape[0]
    s = np.ascontiguousarray(seeds, np.int16)
    d = np.ascontiguousarray(density, np.float32)
    m = _mask(mask, n)
    ms = np.zeros(1, np.float32)
    rc = lib().ref_loop_timed(_p(s), _p(d), _p(m), n, int(iters), _p(ms))
    assert rc == 0, rc
    return float(ms[0])


def discretize(points, weight, triangles, scale, n):
    pts = np.ascontiguousarray(points, np.float64); w = np.ascontiguousarray(weight, np.float64)
    tri = np.ascontiguousarray(triangles, np.int32)
    out = np.empty((n, n), np.float32)
    rc = lib().ref_discretize(_p(pts), _p(w), len(w), _p(tri), len(tri), _p(out), float(scale), n)
    assert rc == 0, rc
    return out


GDEL_PATH = os.path.join(ROOT, "oracle", "_ref", "libgdel2d_ref.so")


def cdt_available():
    return os.path.exists(GDEL_PATH)


def cdt(points, segs, timeout=180):
    """Constrained Delaunay triangulation by the unmodified reference gDel2D, in a subprocess (tests/_cdt_child.py).
    Returns (T,3) int32 triangles; raises RuntimeError with the child's stderr if it fails or times out."""
    import subprocess
    import sys
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        fi, fo = os.path.join(d, "in.npz"), os.path.join(d, "out.npz")
        np.savez(fi, points=np.ascontiguousarray(points, np.float64), segs=np.ascontiguousarray(segs, np.int32))
        try:
            p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_cdt_child.py"), fi, fo], capture_output=True,
                               text=True, timeout=timeout)
        except subprocess.TimeoutExpired:
            raise RuntimeError("reference gDel2D timed out")
        if p.returncode != 0 or not os.path.exists(fo):
            raise RuntimeError("reference gDel2D failed: " + (p.stderr or p.stdout)[-400:])
        return np.load(fo)["tri"]
