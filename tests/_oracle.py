"""ctypes bindings of the CPU oracle (oracle/liboracle.so) — test infrastructure only.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this module.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
MARK = -32768

_lib = None


def build():
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    src = os.path.join(ORACLE_DIR, "srm_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        L = _lib
        p = C.c_void_p
        L.orc_num_threads.restype = C.c_int
        L.orc_put_constraints.argtypes = [p, p, C.c_int]
        L.orc_random_points.argtypes = [p, p, C.c_int, C.c_int, p, C.c_longlong]
        L.orc_random_points.restype = C.c_longlong
        L.orc_label_brute.argtypes = [p, p, C.c_int]
        L.orc_label_exact.argtypes = [p, p, C.c_int]
        L.orc_label_band.argtypes = [p, p, C.c_int, C.c_int, C.c_int, C.c_int, p]
        L.orc_label_jfa.argtypes = [p, p, C.c_int, p, C.c_int]
        L.orc_centroid.argtypes = [p, p, C.c_int, p, p, p]
        L.orc_energy.argtypes = [p, p, C.c_int]
        L.orc_energy.restype = C.c_double
        L.orc_update_sites.argtypes = [p, p, p, p, p, p, C.c_int, C.c_float, p]
        L.orc_lloyd_step.argtypes = [p, p, p, C.c_int, C.c_float, p, p, p]
        L.orc_gcvt.argtypes = [p, p, p, C.c_int, C.c_int, C.c_int, p, p]
        L.orc_gcvt.restype = C.c_int
        L.orc_density_scale.argtypes = [p, p, C.c_int]
        L.orc_zoom_in.argtypes = [p, p, C.c_int]
        L.orc_gcvt_multires.argtypes = [p, p, p, C.c_int, C.c_int, C.c_int, C.c_int, p, p, p]
        L.orc_gcvt_multires.restype = C.c_int
        L.orc_locate.argtypes = [p, p, C.c_int, p, C.c_int, p, p]
        L.orc_recover.argtypes = [p, p, p, C.c_int, p, C.c_int, p, C.c_int, p, C.c_int, p, p]
        L.orc_recover.restype = C.c_int
        L.orc_rasterise.argtypes = [p, p, C.c_int, p, C.c_int, p, C.c_double, C.c_int]
        L.orc_fast_scratch_bytes.argtypes = [C.c_int, C.c_int]
        L.orc_fast_scratch_bytes.restype = C.c_size_t
        L.orc_fast_step.argtypes = [p, p, C.c_int, p, p, C.c_int, C.c_float, p, p]
        L.orc_fast_step.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _mask(mask, n):
    if mask is None:
        return None
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    assert m.shape == (n, n)
    return m


def seed(density, mask, num, state=0, max_attempts=0):
    """putConstrains + randomPoints (gcvt.h:76-139).  Returns (site map int16[n,n,2], attempts, state)."""
    n = density.shape[0]
    d = np.ascontiguousarray(density, dtype=np.float32)
    m = _mask(mask, n)
    vor = np.empty((n, n, 2), np.int16)
    lib().orc_put_constraints(_p(vor), _p(m), n)
    st = np.array([state], np.uint64)
    att = lib().orc_random_points(_p(vor), _p(d), int(num), n, _p(st), int(max_attempts))
    return vor, att, int(st[0])


def label_brute(seeds):
    n = seeds.shape[0]
    s = np.ascontiguousarray(seeds, np.int16)
    out = np.empty_like(s)
    lib().orc_label_brute(_p(s), _p(out), n)
    return out


def label_exact(seeds):
    n = seeds.shape[0]
    s = np.ascontiguousarray(seeds, np.int16)
    out = np.empty_like(s)
    lib().orc_label_exact(_p(s), _p(out), n)
    return out


def label_band(sites_xy, n, r0, r1):
    """Rule A2 for rows [r0, r1) from a site list (K,2) int16 (x,y): labels int16[r1-r0, n, 2]."""
    xy = np.asarray(sites_xy, np.int16)
    sx = np.ascontiguousarray(xy[:, 0]); sy = np.ascontiguousarray(xy[:, 1])
    out = np.empty((r1 - r0, n, 2), np.int16)
    lib().orc_label_band(_p(sx), _p(sy), len(sx), int(n), int(r0), int(r1), _p(out))
    return out


def label_jfa(seeds, steps):
    n = seeds.shape[0]
    s = np.ascontiguousarray(seeds, np.int16)
    st = np.ascontiguousarray(steps, np.int32)
    out = np.empty_like(s)
    lib().orc_label_jfa(_p(s), _p(out), n, _p(st), len(st))
    return out


def centroid(labels, density):
    n = labels.shape[0]
    l = np.ascontiguousarray(labels, np.int16)
    d = np.ascontiguousarray(density, np.float32)
    W = np.empty((n, n)); X = np.empty((n, n)); Y = np.empty((n, n))
    lib().orc_centroid(_p(l), _p(d), n, _p(W), _p(X), _p(Y))
    return W, X, Y


def energy(labels, density):
    n = labels.shape[0]
    return lib().orc_energy(_p(np.ascontiguousarray(labels, np.int16)), _p(np.ascontiguousarray(density, np.float32)), n)


def lloyd_step(seeds, density, mask, omega, want_energy=True):
    """Returns (labels, new site map, energy or None)."""
    n = seeds.shape[0]
    s = np.ascontiguousarray(seeds, np.int16)
    d = np.ascontiguousarray(density, np.float32)
    m = _mask(mask, n)
    lab = np.empty_like(s)
    out = np.empty_like(s)
    e = np.zeros(1)
    lib().orc_lloyd_step(_p(s), _p(d), _p(m), n, float(omega), _p(lab), _p(out), _p(e) if want_energy else None)
    return lab, out, (float(e[0]) if want_energy else None)


def gcvt(seeds, density, mask, max_iter, stop_rule=1):
    """Returns (final label map, iterations, energies[list], final omega)."""
    n = seeds.shape[0]
    vor = np.array(seeds, np.int16, copy=True, order="C")
    d = np.ascontiguousarray(density, np.float32)
    m = _mask(mask, n)
    en = np.zeros(max_iter // 10 + 2)
    om = np.zeros(1, np.float32)
    it = lib().orc_gcvt(_p(vor), _p(d), _p(m), n, int(max_iter), int(stop_rule), _p(en), _p(om))
    return vor, it, en[: (it + 9) // 10].tolist(), float(om[0])


def density_scale(density):
    """kernelDensityScaling: one 2x2 box-filter level."""
    n = density.shape[0]
    d = np.ascontiguousarray(density, np.float32)
    out = np.empty((n // 2, n // 2), np.float32)
    lib().orc_density_scale(_p(d), _p(out), n // 2)
    return out


def zoom_in(seeds):
    s = seeds.shape[0]
    a = np.ascontiguousarray(seeds, np.int16)
    out = np.empty((2 * s, 2 * s, 2), np.int16)
    lib().orc_zoom_in(_p(a), _p(out), s)
    return out


def gcvt_multires(coarse_seeds, density, mask, depth, max_iter, stop_rule=1):
    """Returns (final n^2 label map, iterations, iteration count at the end of each level, omega, Energy)."""
    n = density.shape[0]
    s = coarse_seeds.shape[0]
    vor = np.full((n, n, 2), MARK, np.int16)
    vor.reshape(-1)[: 2 * s * s] = np.ascontiguousarray(coarse_seeds, np.int16).reshape(-1)
    d = np.ascontiguousarray(density, np.float32)
    m = _mask(mask, n)
    li = np.zeros(16, np.int32); om = np.zeros(1, np.float32); en = np.zeros(1, np.float32)
    it = lib().orc_gcvt_multires(_p(vor), _p(d), _p(m), n, int(depth), int(max_iter), int(stop_rule), _p(li), _p(om), _p(en))
    nl = 0
    dd = max(1, depth)
    while nl < dd and (n >> nl) >= 256:
        nl += 1
    return vor, it, li[:nl].tolist(), float(om[0]), float(en[0])


def locate(points, triangles, queries):
    """recover.h locate, brute force: (face ids int32[Q], weights float64[Q,3])."""
    pts = np.ascontiguousarray(points, np.float64); tri = np.ascontiguousarray(triangles, np.int32)
    q = np.ascontiguousarray(queries, np.float64)
    face = np.empty(len(q), np.int32); w = np.zeros((len(q), 3))
    lib().orc_locate(_p(pts), _p(tri), len(tri), _p(q), len(q), _p(face), _p(w))
    return face, w


def recover(points, points3d, triangles, pxy, cpoint_vertex, cdt_tri):
    """recover.h recover on arrays: (vertices float64[P,3], keep uint8[M], kept)."""
    pts = np.ascontiguousarray(points, np.float64); p3 = np.ascontiguousarray(points3d, np.float64)
    tri = np.ascontiguousarray(triangles, np.int32); q = np.ascontiguousarray(pxy, np.float64)
    cpv = np.ascontiguousarray(cpoint_vertex, np.int32); cdt = np.ascontiguousarray(cdt_tri, np.int32)
    out = np.zeros((len(q), 3)); keep = np.zeros(len(cdt), np.uint8)
    kept = lib().orc_recover(_p(pts), _p(p3), _p(tri), len(tri), _p(q), len(q), _p(cpv), len(cpv), _p(cdt), len(cdt),
                             _p(out), _p(keep))
    return out, keep, kept


def rasterise(points, weight, triangles, scale, n):
    pts = np.ascontiguousarray(points, np.float64)
    w = np.ascontiguousarray(weight, np.float64)
    tri = np.ascontiguousarray(triangles, np.int32)
    out = np.empty((n, n), np.float32)
    lib().orc_rasterise(_p(pts), _p(w), len(w), _p(tri), len(tri), _p(out), float(scale), n)
    return out


def sites_of(site_map):
    """Site pixels of a site/label map in row-major scan order -> (K,2) int16 (x,y)."""
    n = site_map.shape[0]
    ys, xs = np.nonzero((site_map[..., 0] == np.arange(n)[None, :]) & (site_map[..., 1] == np.arange(n)[:, None]))
    return np.stack([xs, ys], 1).astype(np.int16)


class FastLloyd:
    """OpenMP CPU baseline loop (orc_fast_step) over a site list."""

    def __init__(self, seeds, density, mask):
        self.n = seeds.shape[0]
        xy = sites_of(seeds)
        self.sx = np.ascontiguousarray(xy[:, 0]); self.sy = np.ascontiguousarray(xy[:, 1])
        self.K = len(xy)
        self.d = np.ascontiguousarray(density, np.float32)
        self.m = _mask(mask, self.n)
        self.scratch = np.empty(lib().orc_fast_scratch_bytes(self.n, self.K), np.uint8)

    def step(self, omega=2.0):
        e = np.zeros(1)
        self.K = lib().orc_fast_step(_p(self.sx), _p(self.sy), self.K, _p(self.d), _p(self.m), self.n,
                                     float(omega), _p(self.scratch), _p(e))
        return float(e[0])
