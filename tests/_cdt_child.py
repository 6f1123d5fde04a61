"""Runs the UNMODIFIED reference gDel2D (oracle/_ref/libgdel2d_ref.so, built by oracle/Makefile from
/root/reference/source/gDel2D with flag-only shims) on one input, in its own process: it is decade-old GPU code that
is not part of the product, so a crash or hang there must not take the test process with it.
    python tests/_cdt_child.py in.npz out.npz      in: points (P,2) float64, segs (S,2) int32    out: tri (T,3) int32"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "oracle", "_ref", "libgdel2d_ref.so")

if __name__ == "__main__":
    z = np.load(sys.argv[1])
    pts = np.ascontiguousarray(z["points"], np.float64)
    segs = np.ascontiguousarray(z["segs"], np.int32).reshape(-1, 2)
    L = C.CDLL(PATH)
    L.ref_cdt.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    cap = 4 * len(pts) + 64
    tri = np.zeros((cap, 3), np.int32)
    nt = L.ref_cdt(pts.ctypes.data, len(pts), segs.ctypes.data if len(segs) else None, len(segs), tri.ctypes.data, cap)
    if nt < 0 or nt > cap:
        sys.exit(f"ref_cdt returned {nt}")
    np.savez(sys.argv[2], tri=tri[:nt])
