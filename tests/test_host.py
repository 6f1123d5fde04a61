"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/srm.h declares,
host-only entry points (seeding, mask) match the oracle, argument errors are reported, and compute entry
points fail loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import _inputs as I
import _oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import surface_remesher_b200 as S
    hdr = open(os.path.join(ROOT, "include", "srm.h")).read()
    names = sorted(set(re.findall(r"\b(srm_[a-z_]+)\s*\(", hdr)))
    assert len(names) >= 20
    L = C.CDLL(S.lib_path())
    for nm in names:
        assert hasattr(L, nm), nm


def test_header_is_plain_c_and_example_links(tmp_path):
    """include/srm.h must be usable from C (the boundary is a C ABI): compile and link the C example against libsrm.so;
    without a GPU it must fail loudly with SRM_ERR_CUDA (exit code 3), not fall back to anything."""
    import subprocess
    import surface_remesher_b200 as S
    exe = str(tmp_path / "gcvt_c_abi")
    libdir = os.path.dirname(S.lib_path())
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "gcvt_c_abi.c"), "-o", exe, "-L", libdir, "-lsrm",
                           f"-Wl,-rpath,{libdir}"])
    import torch
    p = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    if torch.cuda.is_available():
        assert p.returncode == 0 and "iterations" in p.stdout, p.stderr
    else:
        assert p.returncode == 3 and "no CUDA device" in p.stderr, (p.returncode, p.stderr)


def test_dropin_caller_links_and_fails_like_gpuErrchk(tmp_path):
    """A C++ caller that only declares gCVT / discretization_d (as gcvt.h:29 and discretization.h:66 do) links against
    libsrm_dropin.so; without a GPU the shim reproduces the reference's error behaviour (gcvt.cu:41-47): a
    `GPUassert: ...` line on stderr and exit(code)."""
    import subprocess
    import surface_remesher_b200 as S
    exe = str(tmp_path / "dropin_caller")
    libdir = os.path.dirname(S.lib_path())
    subprocess.check_call(["/usr/bin/g++", "-std=c++14", "-Wall", os.path.join(ROOT, "examples", "dropin_caller.cpp"), "-o", exe,
                           "-L", libdir, "-lsrm_dropin", "-lsrm", f"-Wl,-rpath,{libdir}"])
    import torch
    p = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    if torch.cuda.is_available():
        assert p.returncode == 0 and "gcvtIterations" in p.stdout, p.stderr
    else:
        assert p.returncode == 2 and p.stderr.startswith("GPUassert: "), (p.returncode, p.stderr)


def test_dropin_library_exports_reference_signatures():
    so = os.path.join(ROOT, "surface-remesher_b200", "libsrm_dropin.so")
    L = C.CDLL(so)
    # mangled names of gCVT(short*,float*,bool*,int,int,int) and discretization_d(double*,double*,int,int*,int,float*,double,int)
    assert hasattr(L, "_Z4gCVTPsPfPbiii")
    assert hasattr(L, "_Z16discretization_dPdS_iPiiPfdi")


@pytest.mark.parametrize("n,k,kind", [(256, 300, "c3"), (256, 1000, "uniform"), (512, 5000, "c3")])
def test_seed_matches_oracle(n, k, kind):
    import surface_remesher_b200 as S
    dens = I.density_uniform(n) if kind == "uniform" else I.density_c3(n)
    mask = None if kind == "uniform" else I.mask_c3(dens)
    exp, att, st = O.seed(dens, mask, k)
    vor = np.empty((n, n, 2), np.int16)
    state = C.c_ulonglong(0)
    rc = S.lib().srm_seed(vor.ctypes.data, dens.ctypes.data, None if mask is None else mask.ctypes.data, k, n,
                          C.byref(state))
    assert rc == 0
    assert np.array_equal(vor, exp) and state.value == st


def test_seed_reports_unsatisfiable_request():
    import surface_remesher_b200 as S
    n = 16
    dens = np.zeros((n, n), np.float32); dens[3, 3] = 1
    vor = np.empty((n, n, 2), np.int16)
    rc = S.lib().srm_seed(vor.ctypes.data, dens.ctypes.data, None, 2, n, None)
    assert rc == 4 and b"placed 1 of 2" in S.lib().srm_last_error()


def test_generate_mask():
    import surface_remesher_b200 as S
    n = 64
    m = np.ones((n, n), np.uint8)
    pts = np.array([[0.0, 0.0], [0.5, 0.25], [0.999, 0.999]])
    S.generateMask(pts, m, n, 1.0 / (n - 1), 0.0, 0.0)
    assert m.sum() == 3 and m[0, 0] and m[int(0.25 * (n - 1)), int(0.5 * (n - 1))] and m[62, 62]
    with pytest.raises(S.SrmError):
        S.generateMask(np.array([[2.0, 0.0]]), m, n, 1.0 / (n - 1), 0.0, 0.0)


def test_argument_errors():
    import surface_remesher_b200 as S
    L = S.lib()
    h = C.c_void_p()
    assert L.srm_create(C.byref(h), 100, 0, 100, 0) == 1          # n not a multiple of 256
    assert L.srm_create(C.byref(h), 256, 0, 100, 0) == 1          # band not a multiple of 64
    assert L.srm_gcvt(None, None, None, 256, 1, 10, None) == 1
    d = np.zeros((16, 16), np.float32)
    p = np.zeros((3, 2)); w = np.zeros(3); t = np.array([[0, 1, 7]], np.int32)
    assert L.srm_discretize(p.ctypes.data, w.ctypes.data, 3, t.ctypes.data, 1, d.ctypes.data, 1.0, 16) == 1  # index out of range
    assert L.srm_discretize(p.ctypes.data, w.ctypes.data, 3, t.ctypes.data, 0, d.ctypes.data, 0.0, 16) == 1  # scale


def test_argument_errors_locate_recover():
    import surface_remesher_b200 as S
    L = S.lib()
    pts = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]]); p3 = np.zeros((3, 3))
    tri = np.array([[0, 1, 2]], np.int32); bad = np.array([[0, 1, 5]], np.int32)
    q = np.array([[0.2, 0.2]]); face = np.zeros(1, np.int32); w = np.zeros(3)
    assert L.srm_locate(None, 3, tri.ctypes.data, 1, q.ctypes.data, 1, face.ctypes.data, w.ctypes.data) == 1
    assert L.srm_locate(pts.ctypes.data, 3, bad.ctypes.data, 1, q.ctypes.data, 1, face.ctypes.data, w.ctypes.data) == 1
    assert b"out of range" in L.srm_last_error()
    out = np.zeros((1, 3)); keep = np.zeros(1, np.uint8); cdt = np.array([[0, 0, 3]], np.int32); cpv = np.array([9], np.int32)
    args = lambda c, v, ncp: (pts.ctypes.data, p3.ctypes.data, 3, tri.ctypes.data, 1, q.ctypes.data, 1, v, ncp, c, 1,
                              out.ctypes.data, keep.ctypes.data, None)
    assert L.srm_recover(*args(cdt.ctypes.data, None, 0)) == 1            # CDT vertex index out of range
    assert L.srm_recover(*args(None, cpv.ctypes.data, 1)) == 1            # constraint vertex out of range / null triangles
    assert L.srm_recover(*args(None, None, 2)) == 1                       # more constraint points than points


def test_tuning_entry_points_validate_their_arguments():
    """srm_host_config / srm_set_variant need no device: range and name checks, and back to the defaults."""
    import surface_remesher_b200 as S
    S.api.host_config(4, 1024)
    S.api.host_config(16, 65536)
    for bad in [(17, 0), (-1, 0), (4, 100), (4, 1 << 20)]:
        with pytest.raises(S.SrmError, match="out of range"):
            S.api.host_config(*bad)
    S.api.host_config(0, 4096)
    S.api.set_variant("expand", 1); S.api.set_variant("prefix", 1)
    S.api.set_variant("expand", -1); S.api.set_variant("prefix", -1)
    with pytest.raises(S.SrmError, match="unknown kernel"):
        S.api.set_variant("band", 1)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import surface_remesher_b200 as S
    with pytest.raises(S.SrmError, match="no CUDA device"):
        S.Context(256)
    v = np.full((256, 256, 2), -32768, np.int16); v[5, 5] = (5, 5)
    with pytest.raises(S.SrmError):
        S.gCVT(v, np.ones((256, 256), np.float32), None, 256, 1, 3)
    with pytest.raises(S.SrmError, match="no CUDA device"):
        S.locate(np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]]), np.array([[0, 1, 2]], np.int32), np.array([[0.2, 0.2]]))


def test_product_never_references_the_oracle():
    pkg = os.path.join(ROOT, "surface-remesher_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                txt = open(os.path.join(dp, f)).read()
                assert "liboracle" not in txt and "_oracle" not in txt and "orc_" not in txt, f


@pytest.mark.gpu
def test_examples_run_on_the_gpu(tmp_path):
    """The same two programs on a GPU box: the C-ABI example and the drop-in caller complete a Lloyd run."""
    test_header_is_plain_c_and_example_links(tmp_path)
    test_dropin_caller_links_and_fails_like_gpuErrchk(tmp_path)


@pytest.mark.parametrize("threads", [1, 3, 8])
def test_host_scans_match_numpy(threads):
    """srm_host.cu: the block-wise scans of the seed map (128-bit OR reduction, pixel loop only in blocks that hold a site)
    and of the constraint mask against a numpy restatement: sizes that leave ragged tails per thread, a 2-byte aligned
    base pointer, sites at block borders and in the x = 0 / y = 0 corner (packed value 0), dense stretches."""
    import surface_remesher_b200 as S
    L = S.api.lib()
    rng = np.random.default_rng(7 + threads)
    assert L.srm_host_config(threads, 0) == 0
    try:
        for pixels in (1, 31, 64, 65, 1000, 4099, 256 * 256 + 17):
            raw = np.full(2 * pixels + 1, I.MARK, np.int16)
            a = raw[1:]                                   # base pointer aligned to 2 bytes only
            k = min(pixels, max(1, pixels // 50))
            idx = np.unique(np.concatenate([rng.choice(pixels, size=k, replace=False), [0, pixels - 1, min(63, pixels - 1), min(64, pixels - 1)]]))
            a[2 * idx] = rng.integers(0, 32767, size=len(idx)); a[2 * idx + 1] = rng.integers(0, 32767, size=len(idx))
            a[0] = 0; a[1] = 0                            # site (0, 0): packed value 0 must be found
            if pixels > 300:
                a[2 * 100:2 * 228:2] = 5                  # 128 consecutive sites
            exp = np.array([int(np.uint16(a[2 * i])) | (int(np.uint16(a[2 * i + 1])) << 16) for i in range(pixels) if a[2 * i] != I.MARK], np.int64)
            out = np.empty(pixels + 1, np.int32); cnt = C.c_int()
            assert L.srm_scan_site_map_host(a.ctypes.data_as(C.c_void_p), C.c_size_t(pixels), out.ctypes.data_as(C.c_void_p), pixels + 1, C.byref(cnt)) == 0
            assert cnt.value == len(exp) and (out[: cnt.value].astype(np.int64) & 0xffffffff == exp).all()
            # capacity smaller than the count: the count is still reported
            assert L.srm_scan_site_map_host(a.ctypes.data_as(C.c_void_p), C.c_size_t(pixels), out.ctypes.data_as(C.c_void_p), 0, C.byref(cnt)) == 0
            assert cnt.value == len(exp)
        for n, r0, r1 in ((256, 0, 256), (512, 64, 448), (264, 3, 200)):
            m = np.zeros((n, n), np.uint8)
            ys = rng.integers(0, n, 300); xs = rng.integers(0, n, 300)
            m[ys, xs] = rng.integers(1, 255, 300)
            m[r0, :] = 1; m[r1 - 1, n - 1] = 255; m[r0, 0] = 7
            got = np.sort(S.api.scan_mask(m, n, r0, r1))
            yy, xx = np.nonzero(m[r0:r1])
            assert (got == np.sort((xx | ((yy + r0) << 16)).astype(np.int32))).all()
    finally:
        L.srm_host_config(0, 0)
