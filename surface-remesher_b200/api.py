"""ctypes binding of libsrm.so and the reference-named host functions.

Names, argument order and meaning follow the reference (source/gcvt.h:29,76-159,
source/discretization.h:66-67); arrays are numpy, modified in place where the reference
modifies its buffers in place.
"""
import ctypes as C
import os

import numpy as np

MARKER = -32768
_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


class SrmError(RuntimeError):
    pass


class Stats(C.Structure):
    _fields_ = [("iterations", C.c_int), ("num_sites", C.c_int), ("stopped", C.c_int), ("omega", C.c_float),
                ("energy", C.c_float), ("ms_device", C.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def lib_path():
    # SRM_LIB: alternative build of the same library (kernel-variant A/B measurements only)
    return os.environ.get("SRM_LIB") or os.path.join(_HERE, "libsrm.so")


def lib():
    """Load libsrm.so (built in-tree by __graft_entry__.build() / csrc/Makefile).  Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise SrmError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    L = C.CDLL(path)
    p, i, d = C.c_void_p, C.c_int, C.c_double
    L.srm_last_error.restype = C.c_char_p
    L.srm_version.restype = i
    L.srm_launch_count.restype = C.c_longlong
    L.srm_launch_count.argtypes = []
    L.srm_gcvt.argtypes = [p, p, p, i, i, i, p]
    L.srm_discretize.argtypes = [p, p, i, p, i, p, d, i]
    L.srm_release_cache.argtypes = []
    L.srm_seed.argtypes = [p, p, p, i, i, p]
    L.srm_generate_mask.argtypes = [p, p, i, i, d, d, d]
    L.srm_locate.argtypes = [p, i, p, i, p, i, p, p]
    L.srm_recover.argtypes = [p, p, i, p, i, p, i, p, i, p, i, p, p, C.POINTER(i)]
    L.srm_create.argtypes = [C.POINTER(p), i, i, i, i]
    L.srm_destroy.argtypes = [p]
    L.srm_set_stream.argtypes = [p, p]
    L.srm_nccl_unique_id.argtypes = [p]
    L.srm_nccl_init.argtypes = [p, p, i, i]
    L.srm_p2p_info.argtypes = [p, p]
    L.srm_p2p_connect.argtypes = [p, p, i, i]
    L.srm_p2p_disconnect.argtypes = [p]
    L.srm_synchronize.argtypes = [p]
    L.srm_set_density.argtypes = [p, p, i]
    L.srm_set_mask.argtypes = [p, p, i]
    L.srm_set_density_band.argtypes = [p, p, i]
    L.srm_set_mask_pixels.argtypes = [p, p, i]
    L.srm_scan_site_map_host.argtypes = [p, C.c_size_t, p, i, C.POINTER(i)]
    L.srm_scan_mask_host.argtypes = [p, i, i, i, p, i, C.POINTER(i)]
    L.srm_shared_bits.argtypes = [p, i, C.POINTER(p), C.POINTER(C.c_size_t)]
    L.srm_set_site_map.argtypes = [p, p, i]
    L.srm_set_sites.argtypes = [p, p, i, i]
    L.srm_get_sites.argtypes = [p, p, i, C.POINTER(i)]
    L.srm_set_omega.argtypes = [p, C.c_float]
    L.srm_extract_sites.argtypes = [p, p, d, d, d, p, i, C.POINTER(i)]
    L.srm_set_option.argtypes = [p, C.c_char_p, i]
    L.srm_label.argtypes = [p]
    L.srm_accumulate.argtypes = [p, i]
    L.srm_accumulate_dense.argtypes = [p, p, i]
    L.srm_label_accumulate.argtypes = [p, i]
    L.srm_update.argtypes = [p]
    L.srm_acc_buffer.argtypes = [p, C.POINTER(p), C.POINTER(C.c_size_t)]
    L.srm_iterate.argtypes = [p, i, i]
    L.srm_iterate_profiled.argtypes = [p, i, i, p]
    L.srm_run.argtypes = [p, i, i, p]
    L.srm_get_state.argtypes = [p, p]
    L.srm_debug_counts.argtypes = [p, C.POINTER(C.c_longlong), C.POINTER(i)]
    L.srm_debug_get.argtypes = [p, i, C.POINTER(C.c_longlong)]
    L.srm_debug_band_order.argtypes = [p, p, p, i, C.POINTER(i)]
    L.srm_get_labels.argtypes = [p, p, i]
    L.srm_label_jfa.argtypes = [p, p, i, p, i]
    L.srm_set_variant.argtypes = [C.c_char_p, i]
    L.srm_host_config.argtypes = [i, i]
    L.srm_host_config.restype = i
    L.srm_time_kernel.argtypes = [p, C.c_char_p, i, C.POINTER(C.c_float)]
    L.srm_time_kernel.restype = i
    L.srm_set_variant.restype = i
    L.srm_label_jfa_timed.argtypes = [p, p, i, i, p, i, C.POINTER(i)]
    for name in ("srm_gcvt", "srm_release_cache", "srm_discretize", "srm_seed", "srm_generate_mask", "srm_locate", "srm_recover", "srm_create", "srm_destroy",
                 "srm_set_density_band", "srm_set_mask_pixels", "srm_scan_site_map_host", "srm_scan_mask_host", "srm_shared_bits", "srm_set_stream", "srm_nccl_unique_id", "srm_nccl_init", "srm_p2p_info", "srm_p2p_connect", "srm_p2p_disconnect", "srm_synchronize", "srm_set_density", "srm_set_mask", "srm_set_site_map",
                 "srm_set_sites", "srm_get_sites", "srm_extract_sites", "srm_set_omega", "srm_set_option", "srm_label", "srm_accumulate", "srm_accumulate_dense", "srm_debug_band_order", "srm_label_accumulate", "srm_update",
                 "srm_acc_buffer", "srm_iterate", "srm_iterate_profiled", "srm_run", "srm_get_state", "srm_debug_counts", "srm_debug_get", "srm_get_labels", "srm_label_jfa", "srm_label_jfa_timed"):
        getattr(L, name).restype = i
    _lib = L
    return L


def _ck(rc):
    if rc != 0:
        raise SrmError(f"libsrm error {rc}: {lib().srm_last_error().decode()}")


def _np(a, dtype, shape=None):
    if not isinstance(a, np.ndarray) or a.dtype != np.dtype(dtype) or not a.flags["C_CONTIGUOUS"]:
        raise TypeError(f"expected a C-contiguous numpy array of {np.dtype(dtype)}")
    if shape is not None and a.size != int(np.prod(shape)):
        raise ValueError(f"expected {int(np.prod(shape))} elements, got {a.size}")
    return a.ctypes.data_as(C.c_void_p)


def _mask_ptr(mask, n):
    if mask is None:
        return None, None
    m = mask if (isinstance(mask, np.ndarray) and mask.dtype in (np.uint8, np.bool_) and mask.flags["C_CONTIGUOUS"]) \
        else np.ascontiguousarray(mask, dtype=np.uint8)
    if m.size != n * n:
        raise ValueError("mask must have n*n elements")
    return m, m.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------- reference-named functions

def gCVT(Voronoi, density_d, mask, size, depth, maxIter):
    """gcvt.h:29 / gcvt.cu:1087.  Voronoi: int16[size*size*2] in = seed map, out = final label map."""
    st = Stats()
    keep, mp = _mask_ptr(mask, size)
    _ck(lib().srm_gcvt(_np(Voronoi, np.int16, (size, size, 2)), _np(density_d, np.float32, (size, size)), mp,
                       int(size), int(depth), int(maxIter), C.byref(st)))
    return st.as_dict()


def discretization_d(points, weight, num_point, triangle, num_tri, density, scale, n):
    """discretization.h:66 / discretization.cu:87.  density: float32[n*n], written."""
    _ck(lib().srm_discretize(_np(points, np.float64, (num_point, 2)), _np(weight, np.float64, (num_point,)),
                             int(num_point), _np(triangle, np.int32, (num_tri, 3)), int(num_tri),
                             _np(density, np.float32, (n, n)), float(scale), int(n)))


def putConstrains(Voronoi, mask, n):
    """gcvt.h:106-122."""
    keep, mp = _mask_ptr(mask, n)
    d = np.ones((1,), np.float32)  # unused with num=0
    dens = np.ones((n, n), np.float32)
    _ck(lib().srm_seed(_np(Voronoi, np.int16, (n, n, 2)), _np(dens, np.float32), mp, 0, int(n), None))
    del d


def randomPoints(Voronoi, density, num, size, state=0):
    """gcvt.h:76-104 on an already constrained map; returns the RNG state (reference: starts at 0)."""
    n = size
    st = C.c_ulonglong(state)
    # srm_seed = putConstrains + randomPoints; keep the sites already present by passing them as the mask
    v = Voronoi.reshape(n, n, 2)
    present = np.ascontiguousarray((v[..., 0] != MARKER).astype(np.uint8))
    _ck(lib().srm_seed(_np(Voronoi, np.int16, (n, n, 2)), _np(density, np.float32, (n, n)),
                       present.ctypes.data_as(C.c_void_p), int(num), int(n), C.byref(st)))
    return st.value


def centroidalVoronoi(Voronoi, density, constrainMask, vertices, imageSize, depth, maxIter):
    """gcvt.h:124-139: putConstrains, randomPoints, gCVT."""
    keep, mp = _mask_ptr(constrainMask, imageSize)
    _ck(lib().srm_seed(_np(Voronoi, np.int16, (imageSize, imageSize, 2)), _np(density, np.float32), mp,
                       int(vertices), int(imageSize), None))
    return gCVT(Voronoi, density, constrainMask, imageSize, depth, maxIter)


def generateMask(points_xy, mask, imageSize, scale, l, b):
    """gcvt.h:143-159 with the constraint points given as an (M,2) float64 array."""
    pts = np.ascontiguousarray(points_xy, np.float64)
    _ck(lib().srm_generate_mask(_np(mask, np.uint8, (imageSize, imageSize)), _np(pts, np.float64), len(pts),
                                int(imageSize), float(scale), float(l), float(b)))


def locate(points_2d, faces, queries):
    """recover.h:63-83 for an array of points: (face ids int32[Q] with -1 = in no face, weights float64[Q,3])."""
    pts = np.ascontiguousarray(points_2d, np.float64); tri = np.ascontiguousarray(faces, np.int32)
    q = np.ascontiguousarray(queries, np.float64)
    face = np.empty(len(q), np.int32); w = np.zeros((len(q), 3))
    _ck(lib().srm_locate(_np(pts, np.float64), len(pts), _np(tri, np.int32), len(tri), _np(q, np.float64), len(q),
                         _np(face, np.int32), _np(w, np.float64)))
    return face, w


def recover(points_2d, points_3d, faces, points_xy, cpoint_vertex, cdt_triangles):
    """recover.h:85-153 on arrays: points_3d[v] = surface position of 2-D mesh vertex v; points_xy = CDT input points
    (free sites, then the constraint points, whose mesh vertices are cpoint_vertex); cdt_triangles index points_xy.
    Returns (vertices float64[P,3], keep uint8[M]): the triangles with keep == 1 form the remeshed surface."""
    pts = np.ascontiguousarray(points_2d, np.float64); p3 = np.ascontiguousarray(points_3d, np.float64)
    tri = np.ascontiguousarray(faces, np.int32); q = np.ascontiguousarray(points_xy, np.float64)
    cpv = np.ascontiguousarray(cpoint_vertex, np.int32); cdt = np.ascontiguousarray(cdt_triangles, np.int32)
    out = np.zeros((len(q), 3)); keep = np.zeros(max(len(cdt), 1), np.uint8)
    kept = C.c_int(0)
    _ck(lib().srm_recover(_np(pts, np.float64), _np(p3, np.float64), len(pts), _np(tri, np.int32), len(tri),
                          _np(q, np.float64), len(q), cpv.ctypes.data_as(C.c_void_p) if len(cpv) else None, len(cpv),
                          cdt.ctypes.data_as(C.c_void_p) if len(cdt) else None, len(cdt), _np(out, np.float64),
                          _np(keep, np.uint8), C.byref(kept)))
    return out, keep[: len(cdt)]


def scan_site_map(site_map_rows):
    """Sites of (rows of) a dense seed map as a packed int32 list in row-major order (multi-threaded host scan)."""
    a = np.ascontiguousarray(site_map_rows, np.int16)
    pixels = a.size // 2
    cnt = C.c_int()
    cap = max(1024, pixels // 64)
    while True:
        out = np.empty(cap, np.int32)
        _ck(lib().srm_scan_site_map_host(a.ctypes.data_as(C.c_void_p), pixels, out.ctypes.data_as(C.c_void_p), cap, C.byref(cnt)))
        if cnt.value <= cap:
            return out[: cnt.value]
        cap = cnt.value


def scan_mask(mask, n, row0, row1):
    """Non-zero pixels of rows [row0, row1) of a 1 B/px mask as packed int32 (x | y << 16)."""
    m = mask if (isinstance(mask, np.ndarray) and mask.dtype in (np.uint8, np.bool_) and mask.flags["C_CONTIGUOUS"]) \
        else np.ascontiguousarray(mask, dtype=np.uint8)
    cnt = C.c_int()
    cap = 1 << 16
    while True:
        out = np.empty(cap, np.int32)
        _ck(lib().srm_scan_mask_host(m.ctypes.data_as(C.c_void_p), int(n), int(row0), int(row1), out.ctypes.data_as(C.c_void_p), cap,
                                     C.byref(cnt)))
        if cnt.value <= cap:
            return out[: cnt.value]
        cap = cnt.value


# ---------------------------------------------------------------- handle API

def row_bands(n, world):
    """Row-band partition of an n-row grid over `world` ranks: equal bands, multiples of 64 rows."""
    if n % (64 * world):
        raise ValueError(f"n={n} is not divisible into {world} bands of a multiple of 64 rows")
    h = n // world
    return [(r * h, (r + 1) * h) for r in range(world)]


def row_bands_balanced(n, world, site_rows, unit=None, fixed=2.3):
    """Row bands of roughly equal WORK instead of equal height.  The band kernel's time per row grows with the number
    of sites near the row (band-list and envelope sizes) on top of a fixed part (Phase A scans every column, and rows
    far from any site prune badly), so on densities with empty regions equal-height bands leave the ranks of the dense
    rows as stragglers (C3 on 8 ranks: 1.8x the mean number of sites in the middle bands, 0.1x at the top and bottom).
    Weight of a `unit`-row block = sites in the block + `fixed` x the mean.  `fixed` = 2.3 is calibrated on C3 at
    32768^2 on 8 B200s (0.116 us per sparse row against 0.195 us per dense row); with fixed = 0.3 the sparse bands
    became the stragglers and every configuration measured slower than equal heights, which is
    why bench.py defaults to equal heights. boundaries are multiples of `unit` rows (default 256 = whole carry
    segments).
    Deterministic in its inputs: every rank computes the same partition from the replicated site map.
    site_rows: y coordinates of the sites."""
    if world == 1:
        return [(0, n)]
    if unit is None:
        unit = 256   # whole carry segments: other heights take the generic column sweep (measured 4 % slower at N = 2)
    if n % unit or n // unit < world:
        return row_bands(n, world)
    nb = n // unit
    hist = np.bincount(np.asarray(site_rows, np.int64) // unit, minlength=nb)[:nb].astype(np.float64)
    w = hist + fixed * max(hist.mean(), 1e-9)
    cum = np.concatenate([[0.0], np.cumsum(w)])
    cuts = [0]
    for k in range(1, world):
        target = cum[-1] * k / world
        i = int(np.searchsorted(cum, target))
        if i > 0 and abs(cum[i - 1] - target) <= abs(cum[min(i, nb)] - target):
            i -= 1
        i = max(i, cuts[-1] + 1)            # at least one block per band ...
        i = min(i, nb - (world - k))        # ... and one left for every later band
        cuts.append(i)
    cuts.append(nb)
    return [(cuts[r] * unit, cuts[r + 1] * unit) for r in range(world)]


def rebalance_bands(bands, times, unit=256):
    """Row bands of equal MEASURED work: given the current partition and the band-kernel time of every rank on it, assume
    the cost per row is constant inside each old band and cut the rows so that every new band integrates to the same
    cost.  Cuts are multiples of `unit` rows (whole carry segments), every band keeps at least one unit.  One round from
    equal bands removes most of the imbalance of a density with sparse and dense regions (C4 on 8 GPUs: the slowest band
    0.50 ms against a mean of 0.41 ms).  Deterministic: every rank computes the same partition from the gathered times."""
    world = len(bands)
    n = bands[-1][1]
    if world == 1 or n % unit or n // unit < world:
        return list(bands)
    t = np.maximum(np.asarray(times, np.float64), 1e-9)
    # cumulative cost at every unit boundary
    nb = n // unit
    cost = np.zeros(nb)
    for (a, b), tt in zip(bands, t):
        cost[a // unit: b // unit] = tt / max((b - a) // unit, 1)
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    cuts = [0]
    for k in range(1, world):
        target = cum[-1] * k / world
        i = int(np.searchsorted(cum, target))
        if i > 0 and abs(cum[i - 1] - target) <= abs(cum[min(i, nb)] - target):
            i -= 1
        i = max(i, cuts[-1] + 1)
        i = min(i, nb - (world - k))
        cuts.append(i)
    cuts.append(nb)
    return [(cuts[r] * unit, cuts[r + 1] * unit) for r in range(world)]


def _ptr(a):
    """Device or host pointer of a numpy array / torch tensor -> (void*, on_device)."""
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise TypeError("array must be C-contiguous")
        return a.ctypes.data_as(C.c_void_p), 0
    if hasattr(a, "data_ptr"):  # torch tensor
        if not a.is_contiguous():
            raise TypeError("tensor must be contiguous")
        return C.c_void_p(a.data_ptr()), 1 if a.is_cuda else 0
    raise TypeError(type(a))


class Context:
    """Device-resident Lloyd state for rows [row0,row1) of an n x n grid (include/srm.h handle API)."""

    def __init__(self, n, row0=0, row1=None, device=0):
        self.n, self.row0, self.row1 = int(n), int(row0), int(n if row1 is None else row1)
        self._h = C.c_void_p()
        _ck(lib().srm_create(C.byref(self._h), self.n, self.row0, self.row1, int(device)))

    def close(self):
        if self._h:
            lib().srm_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, cuda_stream_ptr):
        _ck(lib().srm_set_stream(self._h, C.c_void_p(int(cuda_stream_ptr))))

    @staticmethod
    def nccl_unique_id():
        buf = C.create_string_buffer(128)
        _ck(lib().srm_nccl_unique_id(buf))
        return buf.raw

    def nccl_init(self, id128, rank, world):
        _ck(lib().srm_nccl_init(self._h, C.create_string_buffer(bytes(id128), 128), int(rank), int(world)))

    def p2p_info(self):
        buf = C.create_string_buffer(160)
        _ck(lib().srm_p2p_info(self._h, buf))
        return buf.raw

    def p2p_connect(self, blobs, rank, world):
        raw = b"".join(blobs)
        assert len(raw) == 160 * world
        _ck(lib().srm_p2p_connect(self._h, C.create_string_buffer(raw, len(raw)), int(rank), int(world)))

    def p2p_disconnect(self):
        _ck(lib().srm_p2p_disconnect(self._h))

    def synchronize(self):
        _ck(lib().srm_synchronize(self._h))

    def set_density(self, density):
        p, dev = _ptr(density)
        _ck(lib().srm_set_density(self._h, p, dev))

    def set_density_band(self, band_rows):
        """Row bands: only the context's own rows, (row1-row0, n) float32.  The caller then exchanges the slices of
        shared_bits(0) between the ranks."""
        p, dev = _ptr(band_rows)
        _ck(lib().srm_set_density_band(self._h, p, dev))

    def set_mask_pixels(self, packed_xy):
        """Constraint pixels as a packed list (x | y << 16, int32)."""
        a = np.ascontiguousarray(packed_xy, np.int32)
        _ck(lib().srm_set_mask_pixels(self._h, a.ctypes.data_as(C.c_void_p), int(a.shape[0])))

    def shared_bits(self, which):
        """(device pointer, number of 32-bit words) of a full-grid bitmap: 0 = density != 0, 1 = constraint pixels."""
        p, cnt = C.c_void_p(), C.c_size_t()
        _ck(lib().srm_shared_bits(self._h, int(which), C.byref(p), C.byref(cnt)))
        return p.value, cnt.value

    def set_mask(self, mask):
        if mask is None:
            _ck(lib().srm_set_mask(self._h, None, 0))
            return
        if isinstance(mask, np.ndarray) and mask.dtype == np.bool_:
            mask = mask.view(np.uint8)
        p, dev = _ptr(mask)
        _ck(lib().srm_set_mask(self._h, p, dev))

    def set_site_map(self, site_map):
        p, dev = _ptr(site_map)
        _ck(lib().srm_set_site_map(self._h, p, dev))

    def set_sites(self, packed_xy):
        p, dev = _ptr(packed_xy)
        _ck(lib().srm_set_sites(self._h, p, int(packed_xy.shape[0]), dev))

    def get_sites(self):
        k = C.c_int()
        _ck(lib().srm_get_sites(self._h, None, 0, C.byref(k)))
        out = np.empty(k.value, np.int32)
        _ck(lib().srm_get_sites(self._h, out.ctypes.data_as(C.c_void_p), k.value, C.byref(k)))
        return out[: k.value]

    def extract_sites(self, mask=None, scale=1.0, l=0.0, b=0.0):
        """Free sites as (M,2) float64 points in delaunayInput's scan order (delaunay.h:46-57)."""
        k = C.c_int()
        mp = None
        if mask is not None:
            m = np.ascontiguousarray(mask, np.uint8)
            mp = m.ctypes.data_as(C.c_void_p)
        _ck(lib().srm_extract_sites(self._h, mp, float(scale), float(l), float(b), None, 0, C.byref(k)))
        out = np.empty((k.value, 2), np.float64)
        _ck(lib().srm_extract_sites(self._h, mp, float(scale), float(l), float(b), out.ctypes.data_as(C.c_void_p),
                                    k.value, C.byref(k)))
        return out

    def set_option(self, name, value):
        _ck(lib().srm_set_option(self._h, name.encode(), int(value)))

    def set_omega(self, omega):
        _ck(lib().srm_set_omega(self._h, float(omega)))

    def label(self):
        _ck(lib().srm_label(self._h))

    def accumulate(self, want_energy):
        _ck(lib().srm_accumulate(self._h, int(bool(want_energy))))

    def accumulate_dense(self, labels=None, want_energy=False):
        """Stand-alone centroid pass over a dense label map on the device (torch CUDA tensor of this band's rows, int16
        [rows, n, 2]); None = the labels of the last label(), expanded into the context's own dense buffer first."""
        ptr = None
        if labels is not None:
            ptr, on_dev = _ptr(labels)
            if not on_dev:
                raise TypeError("accumulate_dense: labels must be a device tensor")
        _ck(lib().srm_accumulate_dense(self._h, ptr, int(bool(want_energy))))

    def label_accumulate(self, want_energy):
        _ck(lib().srm_label_accumulate(self._h, int(bool(want_energy))))

    def update(self):
        _ck(lib().srm_update(self._h))

    def acc_buffer(self):
        """(device pointer, number of doubles) of the per-site accumulators, for an external all-reduce."""
        p, cnt = C.c_void_p(), C.c_size_t()
        _ck(lib().srm_acc_buffer(self._h, C.byref(p), C.byref(cnt)))
        return p.value, cnt.value

    def iterate(self, iters, stop_rule=False):
        _ck(lib().srm_iterate(self._h, int(iters), int(bool(stop_rule))))

    STAGES = ("bitmap+carry", "band_fused", "robust_rows", "allreduce", "update", "iteration")

    def iterate_profiled(self, iters, stop_rule=False):
        """dict stage -> summed device ms over `iters` iterations (CUDA events between the stages)."""
        ms = (C.c_float * 6)()
        _ck(lib().srm_iterate_profiled(self._h, int(iters), int(bool(stop_rule)), ms))
        return dict(zip(self.STAGES, [float(v) for v in ms]))

    def run(self, max_iter, stop_rule=True):
        st = Stats()
        _ck(lib().srm_run(self._h, int(max_iter), int(bool(stop_rule)), C.byref(st)))
        return st.as_dict()

    def state(self):
        st = Stats()
        _ck(lib().srm_get_state(self._h, C.byref(st)))
        return st.as_dict()

    def debug_counts(self):
        """(total runs of the last labelling, rows that took the robust path)."""
        runs, ovf = C.c_longlong(), C.c_int()
        _ck(lib().srm_debug_counts(self._h, C.byref(runs), C.byref(ovf)))
        return runs.value, ovf.value

    def debug_get(self, which):
        """Statistics counter `which` of the band kernel (option dbg_stats = 1): 0/1/2 band-list max/sum/bands,
        6 = warps that took Phase A's staging-overflow fallback."""
        v = C.c_longlong()
        _ck(lib().srm_debug_get(self._h, int(which), C.byref(v)))
        return v.value

    def debug_band_order(self):
        """(perm, cost): the band kernel's CTA -> band order (option "band_order") and the runs per 8-row band of the
        last labelling, from which the order is rebuilt every 10th iteration."""
        nb = C.c_int()
        _ck(lib().srm_debug_band_order(self._h, None, None, 0, C.byref(nb)))
        perm = np.empty(nb.value, np.int32); cost = np.empty(nb.value, np.int32)
        _ck(lib().srm_debug_band_order(self._h, _ptr(perm)[0], _ptr(cost)[0], nb.value, C.byref(nb)))
        return perm, cost

    def get_labels(self, out=None):
        rows = self.row1 - self.row0
        if out is None:
            out = np.empty((rows, self.n, 2), np.int16)
        p, dev = _ptr(out)
        _ck(lib().srm_get_labels(self._h, p, dev))
        return out

    def label_jfa(self, steps, out=None):
        st = np.ascontiguousarray(steps, np.int32)
        if out is None:
            out = np.empty((self.n, self.n, 2), np.int16)
        p, dev = _ptr(out)
        _ck(lib().srm_label_jfa(self._h, st.ctypes.data_as(C.c_void_p), len(st), p, dev))
        return out

    def time_kernel(self, which, reps=10):
        """Measurement: device ms per launch of a streaming kernel ("prefix" / "expand" / "centroid" / "centroid_energy")
        on the resident data."""
        ms = C.c_float()
        _ck(lib().srm_time_kernel(self._h, which.encode(), int(reps), C.byref(ms)))
        return ms.value

    def label_jfa_timed(self, steps, mode=1):
        """Measurement: device milliseconds of every launch of the schedule (mode 1: runs of small steps are one fused
        shared-memory tile launch; mode 0: one plain kernel per pass)."""
        st = np.ascontiguousarray(steps, np.int32)
        ms = np.zeros(len(st), np.float32)
        nl = C.c_int()
        _ck(lib().srm_label_jfa_timed(self._h, st.ctypes.data_as(C.c_void_p), len(st), int(mode),
                                      ms.ctypes.data_as(C.c_void_p), len(ms), C.byref(nl)))
        return ms[:nl.value].astype(float).tolist()


def host_config(threads=0, chunk_kb=0):
    """Worker threads / staging chunk size of the pageable-copy pipeline and the host scans (0 = default / unchanged)."""
    _ck(lib().srm_host_config(int(threads), int(chunk_kb)))


def set_variant(which, value):
    """Measurement / A-B tests: process-wide choice between the builds of a streaming kernel ("expand": 0 / 1 / 2,
    "prefix": 0 / 1, "centroid": 0 / 1); -1 = back to the compiled default."""
    _ck(lib().srm_set_variant(which.encode(), int(value)))


def unpack_sites(packed):
    """int32 packed sites -> (K,2) int16 (x,y)."""
    p = np.asarray(packed, np.int32)
    return np.stack([(p & 0xFFFF).astype(np.uint16).view(np.int16), (p >> 16).astype(np.int16)], 1)
