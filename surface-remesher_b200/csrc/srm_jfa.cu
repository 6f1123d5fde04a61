// srm_jfa.cu — jump flooding in north_star's form: int-packed labels, vectorised coalesced loads for the far passes, a
// shared-memory tile with a halo that FUSES the consecutive small-step passes of a schedule into one launch.
//
// Not the reference's algorithm (the reference labels exactly, gcvt.cu:73-479) and not the product path
// (DESIGN.md section 2.1); pinned bit-exact to the CPU JFA of the same schedule and key (dist^2, x, y)
// (tests/test_gpu_label.py, tests/test_jfa_tile_model.py).
//
//   k_jfa_far   one pass with any step: 4 pixels per thread, the nine taps of a 4-pixel group are nine 128-bit loads
//               when step % 4 == 0 (every power of two >= 4); algorithmic 8 B/px.
//   k_jfa_tile  a run of consecutive passes whose steps sum to H <= 15 (e.g. 8, 4, 2, 1): the CTA stages its 64 x 64
//               tile plus a halo of H pixels in shared memory (94 x 94 ints), runs the passes between two
//               shared-memory buffers — pass p on the tile widened by the steps still to come, so that every tap of
//               every later pass is a value this CTA computed itself — and writes the last pass straight to global
//               memory: 8 B/px + halo (served by L2) for the whole run instead of 8 B/px per pass.  Values computed
//               redundantly in the halo are the same function of the same inputs as in the neighbouring CTA, so the
//               result is bit-identical to the unfused passes.
//
// Both kernels use one branch-free tap (a rare tie takes the only branch); the round-1 kernel (srm_label.cu:k_jfa) stays
// selectable (option "jfa_mode" 0) as the A/B baseline.
#include "srm_common.cuh"

#define JT_W 64            // tile side
#define JT_HALO 15         // largest sum of fused steps
#define JT_P (JT_W + 2 * JT_HALO)   // pitch and height of a staged tile
#define JT_NT 256
#define JT_MAXFUSE 4

// Candidates are compared by the key (dist^2, x, y) as ONE 64-bit unsigned number: high word dist^2, low word the label
// in "ordered" form (y in the low half, x in the high half; the halves of the packed format swapped — one PRMT).  The
// empty label 0x80008000 is its own ordered form; its key is larger than every real key: for n <= 16384 by arithmetic
// (halves read as unsigned: |32768 - X| > 16384 > any real |dx|, and 2 * 32768^2 = 2^31 still fits), for larger
// grids (BIG) by an explicit select.  No branches: a tie is just a comparison of the low words.
__device__ __forceinline__ unsigned jfa_swap(unsigned v) { return __byte_perm(v, 0u, 0x1032); }
template <bool BIG>
__device__ __forceinline__ unsigned long long jfa_key(unsigned ord, int X, int Y) {
    const int dx = (int)(ord >> 16) - X, dy = (int)(ord & 0xffffu) - Y;
    unsigned d = (unsigned)(dx * dx + dy * dy);
    if (BIG) d = (ord == (unsigned)SRM_SENT) ? 0xffffffffu : d;
    return ((unsigned long long)d << 32) | ord;
}
#define JFA_KEY_EMPTY 0xffffffff80008000ull

template <bool BIG>
__global__ void __launch_bounds__(128) k_jfa_far(const int *__restrict__ in, int *__restrict__ out, int n, int k) {
    const int X0 = (blockIdx.x * blockDim.x + threadIdx.x) << 2, Y = blockIdx.y;
    if (X0 >= n) return;
    unsigned long long best[4] = {JFA_KEY_EMPTY, JFA_KEY_EMPTY, JFA_KEY_EMPTY, JFA_KEY_EMPTY};
    const bool vec = (k & 3) == 0;
#pragma unroll
    for (int j = -1; j <= 1; ++j) {
        const int qy = Y + j * k;
        if (qy < 0 || qy >= n) continue;
        const int *rowp = in + (size_t)qy * n;
#pragma unroll
        for (int i = -1; i <= 1; ++i) {
            const int qx0 = X0 + i * k;
            if ((vec || i == 0) && qx0 >= 0 && qx0 + 3 < n) {   // aligned group inside the grid
                const int4 v = *reinterpret_cast<const int4 *>(rowp + qx0);
                best[0] = min(best[0], jfa_key<BIG>(jfa_swap((unsigned)v.x), X0, Y));
                best[1] = min(best[1], jfa_key<BIG>(jfa_swap((unsigned)v.y), X0 + 1, Y));
                best[2] = min(best[2], jfa_key<BIG>(jfa_swap((unsigned)v.z), X0 + 2, Y));
                best[3] = min(best[3], jfa_key<BIG>(jfa_swap((unsigned)v.w), X0 + 3, Y));
            } else {
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const int qx = qx0 + p;
                    if (qx >= 0 && qx < n)
                        best[p] = min(best[p], jfa_key<BIG>(jfa_swap((unsigned)__ldg(rowp + qx)), X0 + p, Y));
                }
            }
        }
    }
    *reinterpret_cast<int4 *>(out + (size_t)Y * n + X0) =
        make_int4((int)jfa_swap((unsigned)best[0]), (int)jfa_swap((unsigned)best[1]), (int)jfa_swap((unsigned)best[2]),
                  (int)jfa_swap((unsigned)best[3]));
}

struct JfaRun { unsigned steps; int count; };   // step p in bits 4p .. 4p+3 (each <= 15)
__host__ __device__ __forceinline__ int jfa_run_step(const JfaRun &r, int p) { return (int)((r.steps >> (4 * p)) & 15u); }

template <bool BIG>
__global__ void __launch_bounds__(JT_NT) k_jfa_tile(const int *__restrict__ in, int *__restrict__ out, int n, JfaRun run) {
    extern __shared__ unsigned jt_smem[];   // two staged tiles, labels in ordered form
    const int t = threadIdx.x;
    const int gx0 = blockIdx.x * JT_W - JT_HALO, gy0 = blockIdx.y * JT_W - JT_HALO;   // grid position of local (0, 0)
    int H = 0;
    for (int p = 0; p < run.count; ++p) H += jfa_run_step(run, p);

    // stage the tile widened by H; pixels outside the grid hold "empty" and stay empty in every pass
    {
        const int lo = JT_HALO - H, side = JT_W + 2 * H;
        const float inv = 1.0f / (float)side;
        for (int idx = t; idx < side * side; idx += JT_NT) {
            const int ry = (int)(((float)idx + 0.5f) * inv), rx = idx - ry * side;   // exact for idx < 2^14
            const int lx = lo + rx, ly = lo + ry, gx = gx0 + lx, gy = gy0 + ly;
            unsigned v = (unsigned)SRM_SENT;
            if (gx >= 0 && gx < n && gy >= 0 && gy < n) v = jfa_swap((unsigned)__ldg(in + (size_t)gy * n + gx));
            jt_smem[ly * JT_P + lx] = v;
        }
    }
    __syncthreads();

    int m = H;
    for (int p = 0; p < run.count; ++p) {
        const int s = jfa_run_step(run, p);
        m -= s;   // this pass is computed on the tile widened by the steps still to come
        const unsigned *src = jt_smem + (p & 1) * (JT_P * JT_P);
        unsigned *dst = jt_smem + ((p & 1) ^ 1) * (JT_P * JT_P);
        const bool last = p == run.count - 1;
        const int lo = JT_HALO - m, side = JT_W + 2 * m;
        const float inv = 1.0f / (float)side;
        for (int idx = t; idx < side * side; idx += JT_NT) {
            const int ry = (int)(((float)idx + 0.5f) * inv), rx = idx - ry * side;
            const int lx = lo + rx, ly = lo + ry, gx = gx0 + lx, gy = gy0 + ly;
            unsigned long long best = JFA_KEY_EMPTY;
            const bool inside = gx >= 0 && gx < n && gy >= 0 && gy < n;
            if (inside) {
                const unsigned *c = src + ly * JT_P + lx;
#pragma unroll
                for (int j = -1; j <= 1; ++j)
#pragma unroll
                    for (int i = -1; i <= 1; ++i) best = min(best, jfa_key<BIG>(c[j * s * JT_P + i * s], gx, gy));
            }
            if (!last) dst[ly * JT_P + lx] = (unsigned)best;
            else if (inside) out[(size_t)gy * n + gx] = (int)jfa_swap((unsigned)best);
        }
        __syncthreads();
    }
}

// Runs the schedule; fused & 1: consecutive steps with a sum <= 15 go through k_jfa_tile, the others through
// k_jfa_far; fused = 0: one round-1 kernel per pass; fused & 2 (tests): the large-grid instantiations at any size.  Returns the buffer that holds the result (a or b).
// ev != nullptr: an event is recorded before the first launch and after every launch (nlaunch + 1 events).
int *srm_launch_jfa(cudaStream_t st, int *a, int *b, int n, const int *steps, int nsteps, int fused, cudaEvent_t *ev,
                    int evcap, int *nlaunch, cudaError_t *err) {
    *err = cudaSuccess;
    int nl = 0;
    auto mark = [&]() { if (ev && nl < evcap) cudaEventRecord(ev[nl], st); };
    mark();
    if (fused) {   // per function and device; cheap, so not cached
        const int smem = 2 * JT_P * JT_P * (int)sizeof(int);
        *err = cudaFuncSetAttribute(k_jfa_tile<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (*err == cudaSuccess) *err = cudaFuncSetAttribute(k_jfa_tile<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (*err != cudaSuccess) return a;
    }
    const bool big = n > 16384 || (fused & 2);   // the empty label needs an explicit test (jfa_key)
    int s = 0;
    while (s < nsteps) {
        if (!fused) {
            srm_launch_jfa_pass(st, a, b, n, steps[s]);
            ++s;
        } else {
            JfaRun run{};
            int sum = 0;
            while (s + run.count < nsteps && run.count < JT_MAXFUSE && sum + steps[s + run.count] <= JT_HALO) {
                run.steps |= (unsigned)steps[s + run.count] << (4 * run.count);
                sum += steps[s + run.count];
                ++run.count;
            }
            if (run.count > 0) {
                const unsigned tiles = (unsigned)((n + JT_W - 1) / JT_W);
                if (big) SRM_COUNT(), k_jfa_tile<true><<<dim3(tiles, tiles), JT_NT, 2 * JT_P * JT_P * sizeof(int), st>>>(a, b, n, run);
                else SRM_COUNT(), k_jfa_tile<false><<<dim3(tiles, tiles), JT_NT, 2 * JT_P * JT_P * sizeof(int), st>>>(a, b, n, run);
                s += run.count;
            } else {
                const dim3 grid((unsigned)((n / 4 + 127) / 128), (unsigned)n);
                if (big) SRM_COUNT(), k_jfa_far<true><<<grid, 128, 0, st>>>(a, b, n, steps[s]);
                else SRM_COUNT(), k_jfa_far<false><<<grid, 128, 0, st>>>(a, b, n, steps[s]);
                ++s;
            }
        }
        int *tmp = a; a = b; b = tmp;
        ++nl;
        mark();
    }
    if (nlaunch) *nlaunch = nl;
    *err = cudaGetLastError();
    return a;
}
