// srm_recover.cu — point location in the parameter-domain mesh and the lift back to 3-D, sm_100a.
// Replaces the host loops of recover.h:63-153: `locate` tests a point against EVERY
// face until one passes (O(K * F) fp64 barycentric evaluations on one host thread: 20 k sites x 6 k
// faces), once per site and once per CDT triangle centroid.  Here the faces are binned to a uniform
// grid over the mesh's bounding box (conservative: one cell of slack on every side), and one thread per
// query evaluates only its cell's list with the IDENTICAL predicate, keeping the lowest face index that
// passes — recover.h's "first face in index order wins" (:67-81).
//
// Arithmetic: recover.h is host code; its `barycentric` (:30-53) is restated with explicit
// round-to-nearest multiplies / adds (no fused multiply-add, what an x86-64 host build without -mfma
// computes), so face ids AND weights are bit-identical to the CPU oracle.
#include "srm_common.cuh"

#define LOC_NT 128

struct LocGrid {
    double x0, y0, inv;   // cell = floor((p - origin) * inv), clamped to [0, G-1]
    int G;
};

// recover.h:30-53.  w[0] = weight of p1, w[1] = of p2, w[2] = of p3 (the out-parameter rotation of the
// declaration (&w3,&w1,&w2) and of the call (_w[0],_w[1],_w[2]) cancels).
__device__ __forceinline__ bool loc_bary(double x1, double y1, double x2, double y2, double x3, double y3, double x0,
                                         double y0, double *w) {
    const double v0x = __dsub_rn(x2, x1), v0y = __dsub_rn(y2, y1);
    const double v1x = __dsub_rn(x3, x1), v1y = __dsub_rn(y3, y1);
    const double v2x = __dsub_rn(x0, x1), v2y = __dsub_rn(y0, y1);
    const double d00 = __dadd_rn(__dmul_rn(v0x, v0x), __dmul_rn(v0y, v0y));
    const double d01 = __dadd_rn(__dmul_rn(v0x, v1x), __dmul_rn(v0y, v1y));
    const double d11 = __dadd_rn(__dmul_rn(v1x, v1x), __dmul_rn(v1y, v1y));
    const double d20 = __dadd_rn(__dmul_rn(v2x, v0x), __dmul_rn(v2y, v0y));
    const double d21 = __dadd_rn(__dmul_rn(v2x, v1x), __dmul_rn(v2y, v1y));
    const double denom = __dsub_rn(__dmul_rn(d00, d11), __dmul_rn(d01, d01));
    if (denom == 0.0) return false;   // w1 = w2 = w3 = -1: never accepted (:44-47)
    const double w1 = __ddiv_rn(__dsub_rn(__dmul_rn(d11, d20), __dmul_rn(d01, d21)), denom);
    const double w2 = __ddiv_rn(__dsub_rn(__dmul_rn(d00, d21), __dmul_rn(d01, d20)), denom);
    const double w3 = __dsub_rn(__dsub_rn(1.0, w1), w2);
    w[0] = w3; w[1] = w1; w[2] = w2;
    return !(w3 < 0 || w1 < 0 || w2 < 0);
}

__device__ __forceinline__ int loc_cell(double v, double o, double inv, int G) {
    const double c = floor((v - o) * inv);
    if (!(c == c)) return 0;   // NaN
    return (int)fmax(0.0, fmin((double)(G - 1), c));
}

__global__ void k_loc_bin(const double *__restrict__ pts, const int *__restrict__ tri, int T, LocGrid g, int *cnt,
                          const int *__restrict__ off, int *__restrict__ list) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int a = tri[3 * t], b = tri[3 * t + 1], c = tri[3 * t + 2];
    const double xa = pts[2 * a], ya = pts[2 * a + 1], xb = pts[2 * b], yb = pts[2 * b + 1], xc = pts[2 * c],
                 yc = pts[2 * c + 1];
    int cx0 = 0, cx1 = g.G - 1, cy0 = 0, cy1 = g.G - 1;
    const double lox = fmin(xa, fmin(xb, xc)), hix = fmax(xa, fmax(xb, xc));
    const double loy = fmin(ya, fmin(yb, yc)), hiy = fmax(ya, fmax(yb, yc));
    if (lox == lox && hix == hix && loy == loy && hiy == hiy) {   // else NaN: cover everything, the predicate decides
        cx0 = max(0, loc_cell(lox, g.x0, g.inv, g.G) - 1); cx1 = min(g.G - 1, loc_cell(hix, g.x0, g.inv, g.G) + 1);
        cy0 = max(0, loc_cell(loy, g.y0, g.inv, g.G) - 1); cy1 = min(g.G - 1, loc_cell(hiy, g.y0, g.inv, g.G) + 1);
    }
    for (int cy = cy0; cy <= cy1; ++cy)
        for (int cx = cx0; cx <= cx1; ++cx) {
            const int slot = atomicAdd(&cnt[cy * g.G + cx], 1);
            if (list) list[off[cy * g.G + cx] + slot] = t;
        }
}

// One thread per query point.  centroid_of != nullptr: query q is the centroid of CDT triangle q
// ((a + b + c) / 3 per coordinate, CGAL::centroid as used at recover.h:134) of the points in `qxy`.
__global__ void __launch_bounds__(LOC_NT) k_locate(const double *__restrict__ pts, const int *__restrict__ tri,
                                                   LocGrid g, const int *__restrict__ off, const int *__restrict__ cnt,
                                                   const int *__restrict__ list, const double *__restrict__ qxy,
                                                   const int *__restrict__ centroid_of, int Q, int *__restrict__ face,
                                                   double *__restrict__ wout) {
    const int q = blockIdx.x * LOC_NT + threadIdx.x;
    if (q >= Q) return;
    double x, y;
    if (centroid_of) {
        const int a = centroid_of[3 * q], b = centroid_of[3 * q + 1], c = centroid_of[3 * q + 2];
        x = __ddiv_rn(__dadd_rn(__dadd_rn(qxy[2 * a], qxy[2 * b]), qxy[2 * c]), 3.0);
        y = __ddiv_rn(__dadd_rn(__dadd_rn(qxy[2 * a + 1], qxy[2 * b + 1]), qxy[2 * c + 1]), 3.0);
    } else {
        x = qxy[2 * q]; y = qxy[2 * q + 1];
    }
    const int cell = loc_cell(y, g.y0, g.inv, g.G) * g.G + loc_cell(x, g.x0, g.inv, g.G);
    const int m = cnt[cell];
    const int *lst = list + off[cell];
    int best = INT_MAX;
    double w[3], bw[3] = {0, 0, 0};
    for (int k = 0; k < m; ++k) {
        const int t = lst[k];
        if (t >= best) continue;
        const int a = tri[3 * t], b = tri[3 * t + 1], c = tri[3 * t + 2];
        if (loc_bary(pts[2 * a], pts[2 * a + 1], pts[2 * b], pts[2 * b + 1], pts[2 * c], pts[2 * c + 1], x, y, w)) {
            best = t; bw[0] = w[0]; bw[1] = w[1]; bw[2] = w[2];
        }
    }
    face[q] = best == INT_MAX ? -1 : best;
    if (wout) { wout[3 * q] = bw[0]; wout[3 * q + 1] = bw[1]; wout[3 * q + 2] = bw[2]; }
}

// Lift (recover.h:100-106): X = 0.0 + v0.x*w0 + v1.x*w1 + v2.x*w2, in that order, per coordinate.
__global__ void k_lift(const int *__restrict__ tri, const double *__restrict__ pts3d, const int *__restrict__ face,
                       const double *__restrict__ w, int Q, double *__restrict__ out) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    const int f = face[q];
    if (f < 0) { out[3 * q] = out[3 * q + 1] = out[3 * q + 2] = 0.0; return; }   // fixed up on the host (stale f_loc)
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < 3; ++j) s = __dadd_rn(s, __dmul_rn(pts3d[3 * tri[3 * f + j] + d], w[3 * q + j]));
        out[3 * q + d] = s;
    }
}

// ---- host driver.  All pointers are device pointers; bbox/grid from the host copy of the points.
struct SrmLocator {
    LocGrid g{};
    int *cnt = nullptr, *off = nullptr, *list = nullptr;
};

void srm_locator_free(SrmLocator *L) {
    if (!L) return;
    cudaFree(L->cnt); cudaFree(L->off); cudaFree(L->list);
    delete L;
}

cudaError_t srm_locator_build(cudaStream_t st, const double *pts_host, int P, const double *pts_dev, const int *tri_dev,
                              int T, SrmLocator **out) {
    double x0 = 1e300, y0 = 1e300, x1 = -1e300, y1 = -1e300;
    for (int i = 0; i < P; ++i) {
        const double x = pts_host[2 * i], y = pts_host[2 * i + 1];
        if (x == x && y == y) { x0 = x < x0 ? x : x0; x1 = x > x1 ? x : x1; y0 = y < y0 ? y : y0; y1 = y > y1 ? y : y1; }
    }
    if (!(x1 >= x0)) { x0 = y0 = 0; x1 = y1 = 1; }
    double ext = (x1 - x0) > (y1 - y0) ? (x1 - x0) : (y1 - y0);
    if (!(ext > 0)) ext = 1;
    int G = 1;
    while (G < 1024 && G * G < T / 2) G <<= 1;   // ~2 faces per cell before the slack ring
    SrmLocator *L = new SrmLocator();
    L->g.G = G; L->g.x0 = x0; L->g.y0 = y0; L->g.inv = (double)G / ext;
    const int nc = G * G;
    int *total_d = nullptr;
    int total = 0;
    cudaError_t e;
    if ((e = cudaMalloc(&L->cnt, sizeof(int) * nc)) != cudaSuccess) goto fail;
    if ((e = cudaMalloc(&L->off, sizeof(int) * nc)) != cudaSuccess) goto fail;
    if ((e = cudaMalloc(&total_d, sizeof(int))) != cudaSuccess) goto fail;
    cudaMemsetAsync(L->cnt, 0, sizeof(int) * nc, st);
    cudaMemsetAsync(L->off, 0, sizeof(int) * nc, st);
    if (T > 0) {
        SRM_COUNT(), k_loc_bin<<<(T + 127) / 128, 128, 0, st>>>(pts_dev, tri_dev, T, L->g, L->cnt, nullptr, nullptr);
        srm_launch_scan_counts(st, L->cnt, L->off, nc, total_d);
        cudaMemcpyAsync(&total, total_d, sizeof(int), cudaMemcpyDeviceToHost, st);
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) goto fail;
    }
    if ((e = cudaMalloc(&L->list, sizeof(int) * (size_t)(total > 0 ? total : 1))) != cudaSuccess) goto fail;
    if (T > 0) {
        cudaMemsetAsync(L->cnt, 0, sizeof(int) * nc, st);
        SRM_COUNT(), k_loc_bin<<<(T + 127) / 128, 128, 0, st>>>(pts_dev, tri_dev, T, L->g, L->cnt, L->off, L->list);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) goto fail;
    cudaFree(total_d);
    *out = L;
    return cudaSuccess;
fail:
    cudaFree(total_d);
    srm_locator_free(L);
    return e;
}

cudaError_t srm_locator_query(cudaStream_t st, const SrmLocator *L, const double *pts_dev, const int *tri_dev,
                              const double *qxy_dev, const int *centroid_of_dev, int Q, int *face_dev, double *w_dev) {
    if (Q > 0)
        SRM_COUNT(), k_locate<<<(Q + LOC_NT - 1) / LOC_NT, LOC_NT, 0, st>>>(pts_dev, tri_dev, L->g, L->off, L->cnt, L->list, qxy_dev,
                                                               centroid_of_dev, Q, face_dev, w_dev);
    return cudaGetLastError();
}

cudaError_t srm_launch_lift(cudaStream_t st, const int *tri_dev, const double *pts3d_dev, const int *face_dev,
                            const double *w_dev, int Q, double *out_dev) {
    if (Q > 0) SRM_COUNT(), k_lift<<<(Q + 127) / 128, 128, 0, st>>>(tri_dev, pts3d_dev, face_dev, w_dev, Q, out_dev);
    return cudaGetLastError();
}
