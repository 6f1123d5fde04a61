// srm_raster.cu — density rasteriser, sm_100a.  Replaces discretization_d + kernelDiscretization
// (discretization.cu:57-120), which tests every pixel against every triangle (N*T fp64 barycentric
// evaluations).  Here triangles are binned to 16x16-pixel tiles by a conservative bounding box
// (one pixel of slack on every side), and each pixel evaluates only its tile's list with the
// IDENTICAL fp64 predicate, keeping the lowest triangle index that passes — which is the
// reference's "first triangle in index order wins" (discretization.cu:72-82).
//
// The barycentric arithmetic is spelled with explicit round-to-nearest intrinsics in the exact
// contraction pattern nvcc emitted for the reference (DMUL/DFMA order read from the SASS of
// oracle/_ref/libsrm_ref.so), so results are bit-identical, not just close.
#include "srm_common.cuh"

#define RT_TILE 16

__device__ __forceinline__ bool bary_hit(double x1, double y1, double x2, double y2, double x3, double y3, int tx,
                                         int ty, double scale, double &wA, double &wB, double &wC) {
    const double v0x = __dsub_rn(x2, x1), v0y = __dsub_rn(y2, y1);
    const double v1x = __dsub_rn(x3, x1), v1y = __dsub_rn(y3, y1);
    const double v2x = __fma_rn((double)tx, scale, -x1), v2y = __fma_rn((double)ty, scale, -y1);
    const double d00 = __fma_rn(v0x, v0x, __dmul_rn(v0y, v0y));
    const double d01 = __fma_rn(v0x, v1x, __dmul_rn(v0y, v1y));
    const double d11 = __fma_rn(v1x, v1x, __dmul_rn(v1y, v1y));
    const double denom = __fma_rn(d00, d11, -__dmul_rn(d01, d01));
    if (denom == 0.0) return false;  // degenerate triangle never hits (discretization.cu:48-51)
    const double d20 = __fma_rn(v0x, v2x, __dmul_rn(v0y, v2y));
    const double d21 = __fma_rn(v1x, v2x, __dmul_rn(v1y, v2y));
    wB = __ddiv_rn(__fma_rn(d11, d20, -__dmul_rn(d01, d21)), denom);  // weight of p2
    wC = __ddiv_rn(__fma_rn(d00, d21, -__dmul_rn(d01, d20)), denom);  // weight of p3
    wA = __dsub_rn(__dsub_rn(1.0, wB), wC);                            // weight of p1
    return !(wA < 0 || wB < 0 || wC < 0);
}

__device__ __forceinline__ void tri_tiles(const double *pts, const int *tri, int t, double scale, int n, int &tx0,
                                          int &tx1, int &ty0, int &ty1) {
    const int a = tri[3 * t], b = tri[3 * t + 1], c = tri[3 * t + 2];
    const double xa = pts[2 * a], ya = pts[2 * a + 1], xb = pts[2 * b], yb = pts[2 * b + 1], xc = pts[2 * c],
                 yc = pts[2 * c + 1];
    const double lox = fmin(xa, fmin(xb, xc)) / scale, hix = fmax(xa, fmax(xb, xc)) / scale;
    const double loy = fmin(ya, fmin(yb, yc)) / scale, hiy = fmax(ya, fmax(yb, yc)) / scale;
    // NaN / inf coordinates: cover everything (the predicate decides)
    int px0 = 0, px1 = n - 1, py0 = 0, py1 = n - 1;
    if (lox == lox && hix == hix && loy == loy && hiy == hiy) {
        px0 = (int)fmax(0.0, fmin((double)n, floor(lox) - 1.0));
        px1 = (int)fmax(-1.0, fmin((double)(n - 1), ceil(hix) + 1.0));
        py0 = (int)fmax(0.0, fmin((double)n, floor(loy) - 1.0));
        py1 = (int)fmax(-1.0, fmin((double)(n - 1), ceil(hiy) + 1.0));
    }
    tx0 = px0 / RT_TILE; tx1 = px1 < 0 ? -1 : px1 / RT_TILE;
    ty0 = py0 / RT_TILE; ty1 = py1 < 0 ? -1 : py1 / RT_TILE;
}

__global__ void k_tri_bin(const double *__restrict__ pts, const int *__restrict__ tri, int num_tri, double scale, int n,
                          int *cnt, const int *__restrict__ off, int *__restrict__ list) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_tri) return;
    int tx0, tx1, ty0, ty1;
    tri_tiles(pts, tri, t, scale, n, tx0, tx1, ty0, ty1);
    const int nt = n / RT_TILE;
    for (int ty = ty0; ty <= ty1; ++ty)
        for (int tx = tx0; tx <= tx1; ++tx) {
            int slot = atomicAdd(&cnt[ty * nt + tx], 1);
            if (list) list[off[ty * nt + tx] + slot] = t;
        }
}

#define RT_CHUNK 64
__global__ void __launch_bounds__(RT_TILE *RT_TILE) k_raster(const double *__restrict__ pts,
                                                              const double *__restrict__ wt,
                                                              const int *__restrict__ tri, const int *__restrict__ off,
                                                              const int *__restrict__ cnt, const int *__restrict__ list,
                                                              double scale, int n, float *__restrict__ density) {
    __shared__ double sx[RT_CHUNK][6];
    __shared__ int sid[RT_CHUNK];
    const int nt = n / RT_TILE;
    const int tile = blockIdx.y * nt + blockIdx.x;
    const int tx = blockIdx.x * RT_TILE + threadIdx.x, ty = blockIdx.y * RT_TILE + threadIdx.y;
    const int tid = threadIdx.y * RT_TILE + threadIdx.x;
    const int m = cnt[tile];
    const int *lst = list + off[tile];
    int best = INT_MAX;
    for (int base = 0; base < m; base += RT_CHUNK) {
        const int c = min(RT_CHUNK, m - base);
        __syncthreads();
        if (tid < c) {
            const int t = lst[base + tid];
            sid[tid] = t;
            const int a = tri[3 * t], b = tri[3 * t + 1], cc = tri[3 * t + 2];
            sx[tid][0] = pts[2 * a]; sx[tid][1] = pts[2 * a + 1];
            sx[tid][2] = pts[2 * b]; sx[tid][3] = pts[2 * b + 1];
            sx[tid][4] = pts[2 * cc]; sx[tid][5] = pts[2 * cc + 1];
        }
        __syncthreads();
        for (int k = 0; k < c; ++k) {
            const int t = sid[k];
            if (t >= best) continue;
            double wA, wB, wC;
            if (bary_hit(sx[k][0], sx[k][1], sx[k][2], sx[k][3], sx[k][4], sx[k][5], tx, ty, scale, wA, wB, wC)) best = t;
        }
    }
    float res = 0.0f;
    if (best != INT_MAX) {
        const int a = tri[3 * best], b = tri[3 * best + 1], c = tri[3 * best + 2];
        double wA, wB, wC;
        bary_hit(pts[2 * a], pts[2 * a + 1], pts[2 * b], pts[2 * b + 1], pts[2 * c], pts[2 * c + 1], tx, ty, scale, wA,
                 wB, wC);
        res = (float)__fma_rn(wC, wt[c], __fma_rn(wB, wt[b], __dmul_rn(wA, wt[a])));
    }
    density[(size_t)ty * n + tx] = res;
}

// Host driver: all buffers are device pointers except num_*.  Synchronises once (list size).
cudaError_t srm_raster(cudaStream_t st, const double *pts, const double *wt, int num_point, const int *tri, int num_tri,
                       float *density, double scale, int n) {
    (void)num_point;
    const int nt = n / RT_TILE, ntiles = nt * nt;
    int *cnt = nullptr, *off = nullptr, *list = nullptr, *total_d = nullptr;
    cudaError_t e;
    if ((e = cudaMalloc(&cnt, sizeof(int) * ntiles)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&off, sizeof(int) * ntiles)) != cudaSuccess) { cudaFree(cnt); return e; }
    if ((e = cudaMalloc(&total_d, sizeof(int))) != cudaSuccess) { cudaFree(cnt); cudaFree(off); return e; }
    cudaMemsetAsync(cnt, 0, sizeof(int) * ntiles, st);
    cudaMemsetAsync(off, 0, sizeof(int) * ntiles, st);
    int total = 0;
    if (num_tri > 0) {
        SRM_COUNT(), k_tri_bin<<<(num_tri + 127) / 128, 128, 0, st>>>(pts, tri, num_tri, scale, n, cnt, nullptr, nullptr);
        srm_launch_scan_counts(st, cnt, off, ntiles, total_d);
        cudaMemcpyAsync(&total, total_d, sizeof(int), cudaMemcpyDeviceToHost, st);
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) goto done;
        if ((e = cudaMalloc(&list, sizeof(int) * (size_t)(total > 0 ? total : 1))) != cudaSuccess) goto done;
        cudaMemsetAsync(cnt, 0, sizeof(int) * ntiles, st);
        SRM_COUNT(), k_tri_bin<<<(num_tri + 127) / 128, 128, 0, st>>>(pts, tri, num_tri, scale, n, cnt, off, list);
    }
    {
        dim3 grid(nt, nt), block(RT_TILE, RT_TILE);
        SRM_COUNT(), k_raster<<<grid, block, 0, st>>>(pts, wt, tri, off, cnt, list, scale, n, density);
    }
    e = cudaStreamSynchronize(st);
done:
    cudaFree(cnt); cudaFree(off); cudaFree(total_d);
    if (list) cudaFree(list);
    return e;
}
