// oracle/ref_driver.cu — thin C entry points around the UNMODIFIED reference CUDA path.
//
// TEST INFRASTRUCTURE ONLY.  Compiled by oracle/Makefile together with
// /root/reference/source/gcvt.cu and discretization.cu (from where they lie) into
// oracle/_ref/libsrm_ref.so.  It calls the reference's own externally visible host
// functions and globals (gcvt.cu:50-62, 840-1156; discretization.cu:87-120); it
// contains no algorithm of its own.  Used by: tests/ (-m gpu) to pin the oracle and
// the product against the real reference on the B200, tests/golden/make_golden.py,
// and bench.py --impl reference (north_star baseline (a): "the reference's own CUDA
// gCVT on the same B200").
#include <cuda_runtime.h>
#include <unordered_map>
#include <cstdio>
#include <cstring>

// ---- declarations of reference symbols (definitions live in gcvt.cu / discretization.cu)
extern short2 **pbaTextures, *pbaMargin;
extern short2 *pbaVoronoi, *pbaTemp;
extern float **pbaDensity;
extern float pbaOmega;
extern int pbaScale, pbaBuffer, pbaMemSize, pbaTexSize;
extern int gcvtIterations;
extern std::unordered_map<int, int> m_1, m_2, m_3;
void gcvtInitialization(int textureSize);
void pbaCVDDeinitialization();
void pba2DInitializeInput(float *density, bool *mask);
void pba2DCompute(int m1, int m2, int m3);
void pbaCVDDensityScaling(int k);
void pbaCVDComputeWeightedPrefix(int k);
void pbaCVDComputeCentroid();
void pbaCVDUpdateSites();
void pbaCVDZoomIn();
float pbaCVDCalcEnergy();
void gCVT(short *Voronoi, float *density_d, bool *mask, int size, int depth, int maxIter);
void discretization_d(double *points, double *weight, int num_point, int *triangle, int num_tri,
                      float *density, double scale, int n);

static bool size_ok(int n) { return n == 256 || n == 512 || n == 1024 || n == 2048 || n == 4096 || n == 8192; }

extern "C" {

// Whole reference gCVT (gcvt.cu:1087).  Returns iterations run, <0 on bad size.
int ref_gcvt(short *vor, float *density, unsigned char *mask, int n, int depth, int max_iter) {
    if (!size_ok(n)) return -1;
    gCVT(vor, density, (bool *)mask, n, depth, max_iter);
    return gcvtIterations;
}

// Same, timed with CUDA events around the call (alloc + H2D + loop + D2H + free, i.e.
// what a caller of the reference pays).  ms_out receives milliseconds.
int ref_gcvt_timed(short *vor, float *density, unsigned char *mask, int n, int depth, int max_iter, float *ms_out) {
    if (!size_ok(n)) return -1;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    gCVT(vor, density, (bool *)mask, n, depth, max_iter);
    cudaEventRecord(b); cudaEventSynchronize(b);
    cudaEventElapsedTime(ms_out, a, b);
    cudaEventDestroy(a); cudaEventDestroy(b);
    return gcvtIterations;
}

// Labelling only: the reference's pba2DCompute on a seed map (gcvt.cu:967-978).
int ref_label(short *vor_inout, int n) {
    if (!size_ok(n)) return -1;
    gcvtInitialization(n);
    pbaVoronoi = pbaTextures[0]; pbaTemp = pbaTextures[1]; pbaBuffer = 0;
    cudaMemcpy(pbaVoronoi, vor_inout, (size_t)n * n * sizeof(short2), cudaMemcpyHostToDevice);
    pba2DCompute(m_1[n], m_2[n], m_3[n]);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(vor_inout, pbaVoronoi, (size_t)n * n * sizeof(short2), cudaMemcpyDeviceToHost);
    pbaCVDDeinitialization();
    return e == cudaSuccess ? 0 : -2;
}

// One teacher-forced Lloyd iteration with the reference kernels, in the order of the
// loop body at gcvt.cu:1112-1123: label, (energy), centroid, update.  Any output may be NULL.
int ref_step(const short *seeds, float *density, unsigned char *mask, int n, float omega,
             short *labels_out, short *seeds_out, float *energy_out) {
    if (!size_ok(n)) return -1;
    gcvtInitialization(n);
    pba2DInitializeInput(density, (bool *)mask);
    pbaCVDDensityScaling(1);
    pbaCVDComputeWeightedPrefix(1);
    pbaScale = 0;
    cudaMemcpy(pbaVoronoi, seeds, (size_t)n * n * sizeof(short2), cudaMemcpyHostToDevice);
    pbaOmega = omega;
    pba2DCompute(m_1[n], m_2[n], m_3[n]);
    if (labels_out) cudaMemcpy(labels_out, pbaVoronoi, (size_t)n * n * sizeof(short2), cudaMemcpyDeviceToHost);
    if (energy_out) *energy_out = pbaCVDCalcEnergy();
    pbaCVDComputeCentroid();
    pbaCVDUpdateSites();
    cudaError_t e = cudaDeviceSynchronize();
    if (seeds_out) cudaMemcpy(seeds_out, pbaVoronoi, (size_t)n * n * sizeof(short2), cudaMemcpyDeviceToHost);
    pbaCVDDeinitialization();
    return e == cudaSuccess ? 0 : -2;
}

// Device-resident timing of the reference loop body: `iters` iterations (label, centroid,
// update; energy every 10th like gcvt.cu:1116) between CUDA events, inputs already on the GPU.
int ref_loop_timed(const short *seeds, float *density, unsigned char *mask, int n, int iters, float *ms_out) {
    if (!size_ok(n)) return -1;
    gcvtInitialization(n);
    pba2DInitializeInput(density, (bool *)mask);
    pbaCVDDensityScaling(1);
    pbaCVDComputeWeightedPrefix(1);
    pbaScale = 0;
    cudaMemcpy(pbaVoronoi, seeds, (size_t)n * n * sizeof(short2), cudaMemcpyHostToDevice);
    pbaOmega = 2.0f;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int it = 0; it < iters; ++it) {
        pba2DCompute(m_1[n], m_2[n], m_3[n]);
        if (it % 10 == 0) (void)pbaCVDCalcEnergy();
        pbaCVDComputeCentroid();
        pbaCVDUpdateSites();
    }
    cudaEventRecord(b); cudaEventSynchronize(b);
    cudaEventElapsedTime(ms_out, a, b);
    cudaEventDestroy(a); cudaEventDestroy(b);
    cudaError_t e = cudaDeviceSynchronize();
    pbaCVDDeinitialization();
    return e == cudaSuccess ? 0 : -2;
}

// Multires pieces (a11).  Level `level` (>= 1) of the reference's density pyramid
// (pbaCVDDensityScaling -> kernelDensityScaling, gcvt.cu:985-993, 497-511): out = (n >> level)^2 floats.
int ref_pyramid(float *density, unsigned char *mask, int n, int level, float *out) {
    if (!size_ok(n) || level < 1 || (n >> level) < 256) return -1;
    gcvtInitialization(n);
    pba2DInitializeInput(density, (bool *)mask);
    pbaCVDDensityScaling(level + 1);
    cudaError_t e = cudaDeviceSynchronize();
    const size_t s = (size_t)(n >> level);
    cudaMemcpy(out, pbaDensity[level], s * s * sizeof(float), cudaMemcpyDeviceToHost);
    pbaCVDDeinitialization();
    return e == cudaSuccess ? 0 : -2;
}

// pbaCVDZoomIn (gcvt.cu:1036-1051): seed map of side s -> seed map of side 2s.
int ref_zoom(const short *seeds_s, int s, short *seeds_2s) {
    if (!size_ok(2 * s) || s < 256) return -1;
    gcvtInitialization(2 * s);
    pbaVoronoi = pbaTextures[0]; pbaTemp = pbaTextures[1]; pbaBuffer = 0;
    pbaTexSize = s;
    cudaMemcpy(pbaVoronoi, seeds_s, (size_t)s * s * sizeof(short2), cudaMemcpyHostToDevice);
    pbaCVDZoomIn();
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(seeds_2s, pbaVoronoi, (size_t)4 * s * s * sizeof(short2), cudaMemcpyDeviceToHost);
    pbaTexSize = 2 * s;
    pbaCVDDeinitialization();
    return e == cudaSuccess ? 0 : -2;
}

int ref_discretize(double *points, double *weight, int num_point, int *triangle, int num_tri,
                   float *density, double scale, int n) {
    discretization_d(points, weight, num_point, triangle, num_tri, density, scale, n);
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : -2;
}

}  // extern "C"
