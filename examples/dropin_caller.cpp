// examples/dropin_caller.cpp — a caller written the way the reference's headers reach the hot path: the two
// functions are only DECLARED (gcvt.h:29, discretization.h:66-67) and resolved at link time — here against
// libsrm_dropin.so instead of the reference's gcvt.cu / discretization.cu.
//   g++ examples/dropin_caller.cpp -Lsurface-remesher_b200 -lsrm_dropin -lsrm -Wl,-rpath,$PWD/surface-remesher_b200
#include <cstdio>
#include <cstdlib>
#include <vector>

extern void gCVT(short *Voronoi, float *density_d, bool *mask, int size, int depth, int maxIter);
extern void discretization_d(double *points, double *weight, int num_point, int *triangle, int num_tri,
                             float *density, double scale, int n);
extern int gcvtIterations;

int main() {
    const int n = 256;
    // one triangle covering the lower-left half of the unit square, rasterised like discretization.h:120 does
    std::vector<double> pts = {0.0, 0.0, 1.0, 0.0, 0.0, 1.0}, wt = {1.0, 2.0, 3.0};
    std::vector<int> tri = {0, 1, 2};
    float *density = (float *)malloc(sizeof(float) * n * n);
    discretization_d(pts.data(), wt.data(), 3, tri.data(), 1, density, 1.0 / (n - 1), n);
    // seeds the way gcvt.h does it: MARKER everywhere, a few sites on positive density
    short *vor = (short *)malloc(sizeof(short) * 2 * n * n);
    bool *mask = (bool *)calloc((size_t)n * n, sizeof(bool));
    for (int i = 0; i < 2 * n * n; ++i) vor[i] = -32768;
    for (int k = 0; k < 40; ++k) {
        const int x = 5 + (k * 37) % 100, y = 5 + (k * 53) % 100;
        if (density[y * n + x] > 0) { vor[2 * (y * n + x)] = (short)x; vor[2 * (y * n + x) + 1] = (short)y; }
    }
    gCVT(vor, density, mask, n, 1, 30);
    std::printf("gcvtIterations %d, label of pixel (0,0): (%d,%d)\n", gcvtIterations, vor[0], vor[1]);
    free(density); free(vor); free(mask);
    return 0;
}
