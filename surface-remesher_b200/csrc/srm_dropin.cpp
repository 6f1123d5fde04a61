// srm_dropin.cpp — the two C++ symbols Surface-Remesher's headers declare `extern` and main.cpp
// reaches through centroidalVoronoi() / discretization():
//     void gCVT(short*, float*, bool*, int, int, int)                       gcvt.h:29   (def gcvt.cu:1087)
//     void discretization_d(double*, double*, int, int*, int, float*, double, int)
//                                                                           discretization.h:66 (def discretization.cu:87)
// Same mangled names, same argument meaning, same error behaviour as the reference's gpuErrchk
// (gcvt.cu:41-47): print "GPUassert: ..." to stderr and exit.  Linking an unmodified main.cpp
// against libsrm_dropin.so + libsrm.so instead of gcvt.cu/discretization.cu swaps the hot path.
#include "../../include/srm.h"

#include <cstdio>
#include <cstdlib>

int gcvtIterations = 0;  // the reference exposes its iteration count as a global (gcvt.cu:1086)

static void die(int code, const char *file, int line) {
    std::fprintf(stderr, "GPUassert: %s %s %d\n", srm_last_error(), file, line);
    std::exit(code);
}

void gCVT(short *Voronoi, float *density_d, bool *mask, int size, int depth, int maxIter) {
    static_assert(sizeof(bool) == 1, "bool mask is one byte per pixel (gcvt.cu:866)");
    srm_stats st;
    int rc = srm_gcvt(Voronoi, density_d, reinterpret_cast<const unsigned char *>(mask), size, depth, maxIter, &st);
    if (rc != SRM_OK) die(rc, __FILE__, __LINE__);
    gcvtIterations = st.iterations;
}

void discretization_d(double *points, double *weight, int num_point, int *triangle, int num_tri, float *density,
                      double scale, int n) {
    int rc = srm_discretize(points, weight, num_point, triangle, num_tri, density, scale, n);
    if (rc != SRM_OK) die(rc, __FILE__, __LINE__);
}
