// oracle/gdel_driver.cu — C entry point around the UNMODIFIED reference gDel2D (GpuDel::compute,
// source/gDel2D/gDel2D/GpuDelaunay.h:175), the constrained Delaunay triangulation that consumes the hot path's output
// (main.cpp:229-236).  TEST INFRASTRUCTURE ONLY: gDel2D is out of scope for the product (SURVEY §2.1, §8(f4) "build
// shims only"); this lets the config-1 test run the whole pipeline sites -> CDT -> lift on the GPU box.
#include "gDel2D/GpuDelaunay.h"
#include <cstring>

extern "C" {

// points: np (x,y) pairs; segs: ns pairs of point indices.  tri_out: capacity `cap` triangles (3 ints each).
// Returns the number of triangles (may exceed cap: nothing beyond cap is written), or -1 on bad arguments.
int ref_cdt(const double *points, int np, const int *segs, int ns, int *tri_out, int cap) {
    if (!points || np < 3 || ns < 0 || (ns > 0 && !segs) || !tri_out) return -1;
    GDel2DInput in;
    GDel2DOutput out;
    in.pointVec.resize(np);
    for (int i = 0; i < np; ++i) { in.pointVec[i]._p[0] = (RealType)points[2 * i]; in.pointVec[i]._p[1] = (RealType)points[2 * i + 1]; }
    in.constraintVec.resize(ns);
    for (int i = 0; i < ns; ++i) { in.constraintVec[i]._v[0] = segs[2 * i]; in.constraintVec[i]._v[1] = segs[2 * i + 1]; }
    GpuDel gpuDel;
    gpuDel.compute(in, &out);
    const int nt = (int)out.triVec.size();
    for (int i = 0; i < nt && i < cap; ++i)
        for (int k = 0; k < 3; ++k) tri_out[3 * i + k] = out.triVec[i]._v[k];
    return nt;
}

}  // extern "C"
