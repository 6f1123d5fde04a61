// srm_centroid.cu — the centroid pass in north_star's stand-alone form: ONE streaming pass over a dense label map and the
// density (8 B/px read, nothing per-pixel written), warp-level segmented reduction, one set of REDs per run segment.
//
// Reference counterpart: pbaCVDComputeCentroid (gcvt.cu:1008-1023 -> kernelComputeWeightedPrefix*, kernelTotal_X,
// kernelScan_Y, gcvt.cu:514-732: fp32 prefix images + two reductions over the label map) and, with want_energy,
// kernelCalcEnergy + kernelReduce (gcvt.cu:788-832, 1059-1083).
//
// NOT the product path.  The Lloyd loop never materialises labels: the band kernel (srm_band.cu) adds the fp64 prefix
// differences of every run to its site while it still holds the run in shared memory, which
// costs 22 us of REDs + 11 us of prefix fetches per iteration at 8192^2.  This kernel exists for callers that already
// hold a dense label map (the jump-flooding family srm_label_jfa, external labellings), as the parity counterpart of
// the reference's own data flow (labels -> sums), and as the measured answer to "what would a label-map centroid pass
// cost at its roofline": 8 B/px * N streamed once.
//
// One warp = 128 consecutive pixels of one row (n is a multiple of 256, so a warp never straddles rows); a thread owns 4
// pixels (one 128-bit streaming load of labels, one of the density; the loads of the next step are issued before the
// current one is reduced).  Every thread ends up with ONE item (label, W, X) for the warp's segmented reduction:
//   * a thread whose 4 labels agree ("uniform", ~85 % at 26-pixel runs): its label and sums;
//   * a thread that straddles run boundaries: its TRAILING sub-run (the run that continues into the next lane); its
//     LEADING sub-run is handed to the lane on its left (one shuffle step) when that lane's item carries the same label
//     — i.e. it is the tail of the run the left lanes are summing — and complete runs inside the thread (<= 3 pixels,
//     rare) go out directly.
// Heads are the lanes that start a run: lane 0, every straddling lane (its item starts inside the lane), and uniform
// lanes whose label differs from the item on their left.  A lane adds its partner at distance o only if no head lies
// in between (read off the ballot of the heads), so non-adjacent runs of one label (possible in approximate labellings)
// are never merged, and five shuffle steps leave every run's (W, X) in its head lane.  Heads look the site id up in the
// pixel -> id hash and issue three fp64 REDs (W, X, Y * W): one set per run and row (+ one per run that crosses a
// 128-pixel warp boundary), what the band kernel issues from its run lists.  North_star's "shared-memory per-site
// accumulators": a shared-memory fp64 atomicAdd is a CAS loop (ATOMS.CAST.SPIN, ~64 cycles per warp instruction),
// dearer than the global RED it would save (1.3 cycles per lane; profiles/r2_kernel_log.md), so the per-warp run sums go
// to global memory directly.
//
// Sums are fp64 in tree order; the float the update law rounds them to is what the parity tests compare
// (tests/test_gpu_centroid.py: per-site sums against the run-based kernel and direct numpy sums, the updated sites
// bit-exact against the oracle's step).
#include "srm_common.cuh"
#include <stdlib.h>

#define CEN_NT 256
#define CEN_MINCTA 6

__device__ __forceinline__ void cen_emit(int label, double W, double X, int Y, const SrmHash &hash, double *__restrict__ acc,
                                         int Kcap, int touch) {
    if (label == SRM_SENT) return;   // pixel without a site (empty site set)
    const int id = srm_hash_find(hash, (unsigned)label);
    if (id < 0) return;              // not a live site of this iteration (foreign label map): ignored
    double *a = acc + 4 * (size_t)id;
    atomicAdd(a, W);
    atomicAdd(a + 1, X);
    atomicAdd(a + 2, (double)Y * W);
    if (touch) reinterpret_cast<unsigned char *>(acc + 4 * (size_t)Kcap + 4)[id] = 1;
}

__global__ void __launch_bounds__(CEN_NT, CEN_MINCTA)
k_centroid_dense(const int4 *__restrict__ labels4, const float4 *__restrict__ dens4, SrmHash hash, int n, int row0, int nrows,
                 double *__restrict__ acc, int Kcap, int want_energy, int touch) {
    const int lane = threadIdx.x & 31;
    const unsigned gpr = (unsigned)n >> 2;                               // 4-pixel groups per row (a multiple of 64)
    const unsigned groups = (unsigned)nrows * gpr;                       // <= 32768 * 8192 = 2^28
    const unsigned stride = gridDim.x * CEN_NT;
    double e_loc = 0;
    unsigned g = blockIdx.x * CEN_NT + threadIdx.x;                      // warp-uniform trip count: groups % 32 == 0
    int4 L = make_int4(0, 0, 0, 0);
    float4 D = make_float4(0, 0, 0, 0);
    if (g < groups) { L = __ldcs(labels4 + g); D = __ldcs(dens4 + g); }
    while (g < groups) {
        const unsigned gn = g + stride;
        int4 Ln = L;
        float4 Dn = D;
        if (gn < groups) { Ln = __ldcs(labels4 + gn); Dn = __ldcs(dens4 + gn); }   // in flight while this step is reduced
        const unsigned r = g / gpr;
        const int x0 = (int)(g - r * gpr) << 2, Y = row0 + (int)r;
        const int lab[4] = {L.x, L.y, L.z, L.w};
        const double d[4] = {(double)D.x, (double)D.y, (double)D.z, (double)D.w};
        if (want_energy) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int dx = (int)(short)(lab[k] & 0xffff) - (x0 + k), dy = (lab[k] >> 16) - Y;
                if (lab[k] != SRM_SENT) e_loc += d[k] * (double)(dx * dx + dy * dy);
            }
        }
        const bool uni = lab[0] == lab[1] && lab[1] == lab[2] && lab[2] == lab[3];
        // item: the sub-run that reaches the thread's last pixel; WL / XL: the leading sub-run of a straddling thread
        int item = lab[3];
        double W, X, WL = 0, XL = 0;
        if (uni) {
            W = (d[0] + d[1]) + (d[2] + d[3]);
            X = ((double)x0 * d[0] + (double)(x0 + 1) * d[1]) + ((double)(x0 + 2) * d[2] + (double)(x0 + 3) * d[3]);
        } else {
            int cur = lab[0];
            double w = d[0], x = (double)x0 * d[0];
            bool leading = true;
#pragma unroll
            for (int k = 1; k < 4; ++k) {
                if (lab[k] == cur) { w += d[k]; x += (double)(x0 + k) * d[k]; }
                else {
                    if (leading) { WL = w; XL = x; leading = false; }
                    else cen_emit(cur, w, x, Y, hash, acc, Kcap, touch);   // a complete run inside the thread
                    cur = lab[k]; w = d[k]; x = (double)(x0 + k) * d[k];
                }
            }
            W = w; X = x;
        }
        // hand the leading sub-run to the left lane if that lane's item is the same run (same label, adjacent pixels)
        const int left = __shfl_up_sync(0xffffffffu, item, 1);
        const bool accepted = !uni && lane > 0 && left == lab[0];
        if (!uni && !accepted) cen_emit(lab[0], WL, XL, Y, hash, acc, Kcap, touch);   // run ends here and starts at or before the thread's first pixel
        {
            const double gW = __shfl_down_sync(0xffffffffu, accepted ? WL : 0.0, 1), gX = __shfl_down_sync(0xffffffffu, accepted ? XL : 0.0, 1);
            if (lane < 31) { W += gW; X += gX; }
        }
        const bool head = lane == 0 || !uni || item != left;
        const unsigned heads = __ballot_sync(0xffffffffu, head);
        const unsigned after = (heads >> 1) >> lane;   // bit k: lane + 1 + k starts a new run
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {             // lane + o lies in this lane's run iff no head in (lane, lane + o]
            const double W2 = __shfl_down_sync(0xffffffffu, W, o), X2 = __shfl_down_sync(0xffffffffu, X, o);
            if (lane + o < 32 && (after & ((1u << o) - 1u)) == 0) { W += W2; X += X2; }
        }
        if (head) cen_emit(item, W, X, Y, hash, acc, Kcap, touch);
        L = Ln; D = Dn; g = gn;
    }
    if (want_energy) {
        e_loc = warp_sum(e_loc);
        if (lane == 0 && e_loc != 0.0) atomicAdd(acc + 4 * (size_t)Kcap, e_loc);
    }
}

// ---- Build 2 (default): the same reduction with fewer instructions.  The session-4 capture of the kernel above
// (profiles/r2_ncu_streams_s14.md) shows it bound by instruction issue, not by HBM: 138 M warp instructions per launch
// (264 per 128 pixels), issue slots 69 % busy, DRAM traffic = the algorithmic 537 MB.  Where they went and what replaces them:
//   * g / gpr per step (an emulated integer division, 6.5 %)          -> row / column carried along, one division per thread;
//   * the sub-run loop of straddling threads, executed by nearly every warp with a few active lanes (13 %) and the four
//     int -> double conversions of the x coordinates (5 %)             -> straight-line: in-thread prefix sums p[k], q[k] of
//     d and k*d, sub-run sums as differences picked by the first / last boundary (ffs / clz of a 3-bit mask), X = x0 * W + q;
//   * five steps of the segmented tree regardless of the run lengths (30 % with its shuffles) -> the tree stops at the first
//     distance no lane can use (warp vote): runs of ~26 pixels span 7 lanes, so distances 16 and mostly 8 are skipped;
//   * one launch per wave of resident CTAs was slower than four (203 vs 163 us): several short CTAs per slot.
template <bool ENERGY>
__global__ void __launch_bounds__(CEN_NT, CEN_MINCTA)
k_centroid_dense2(const int4 *__restrict__ labels4, const float4 *__restrict__ dens4, SrmHash hash, int n, int row0, int nrows,
                  double *__restrict__ acc, int Kcap, int touch) {
    const int lane = threadIdx.x & 31;
    const unsigned gpr = (unsigned)n >> 2;
    const unsigned groups = (unsigned)nrows * gpr;
    const unsigned stride = gridDim.x * CEN_NT;
    const unsigned sr = stride / gpr, sc = stride - sr * gpr;            // the step in (rows, groups of the row)
    unsigned g = blockIdx.x * CEN_NT + threadIdx.x;                      // warp-uniform trip count: groups % 32 == 0
    unsigned r = g / gpr, c = g - r * gpr;
    double e_loc = 0;
    int4 L = make_int4(0, 0, 0, 0);
    float4 D = make_float4(0, 0, 0, 0);
    if (g < groups) { L = __ldcs(labels4 + g); D = __ldcs(dens4 + g); }
    while (g < groups) {
        const unsigned gn = g + stride;
        int4 Ln = L;
        float4 Dn = D;
        if (gn < groups) { Ln = __ldcs(labels4 + gn); Dn = __ldcs(dens4 + gn); }
        const int x0 = (int)c << 2, Y = row0 + (int)r;
        const double d0 = (double)D.x, d1 = (double)D.y, d2 = (double)D.z, d3 = (double)D.w;
        if (ENERGY) {
            const int lab[4] = {L.x, L.y, L.z, L.w};
            const double d[4] = {d0, d1, d2, d3};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int dx = (int)(short)(lab[k] & 0xffff) - (x0 + k), dy = (lab[k] >> 16) - Y;
                if (lab[k] != SRM_SENT) e_loc += d[k] * (double)(dx * dx + dy * dy);
            }
        }
        // in-thread prefix sums: p_k = d_0 + .. + d_{k-1}, q_k = sum of j * d_j (x relative to x0); p_0 = q_0 = q_1 = 0
        const double p1 = d0, p2 = p1 + d1, p3 = p2 + d2, p4 = p3 + d3;
        const double q2 = d1, q3 = fma(2.0, d2, q2), q4 = fma(3.0, d3, q3);
        // boundaries inside the thread: bit k-1 set iff pixel k starts a new run
        const unsigned m = (unsigned)(L.y != L.x) | ((unsigned)(L.z != L.y) << 1) | ((unsigned)(L.w != L.z) << 2);
        const bool uni = m == 0;
        const int f = __ffs(m), l = 32 - __clz(m);                        // first / last boundary (1..3); uniform: 0 / 0
        const double pf = f == 1 ? p1 : f == 2 ? p2 : p3, qf = f == 1 ? 0.0 : f == 2 ? q2 : q3;      // leading sub-run [0, f)
        const double pl = l == 0 ? 0.0 : l == 1 ? p1 : l == 2 ? p2 : p3, ql = l <= 1 ? 0.0 : l == 2 ? q2 : q3;
        const double xd = (double)x0;
        const int item = L.w;                                              // the sub-run [l, 4) that reaches the last pixel
        double W = p4 - pl, X = fma(xd, W, q4 - ql);
        if (l > f) {   // complete runs inside the thread (rare: runs of 1-2 pixels); m is 011, 101, 110 or 111 here
            if (m == 7u) {                                                 // boundaries at 1, 2, 3: runs [1,2) and [2,3)
                cen_emit(L.y, d1, fma(xd, d1, d1), Y, hash, acc, Kcap, touch);
                cen_emit(L.z, d2, fma(xd, d2, 2.0 * d2), Y, hash, acc, Kcap, touch);
            } else {                                                       // two boundaries: one run [f, l)
                const double w = pl - pf;
                cen_emit(f == 1 ? L.y : L.z, w, fma(xd, w, ql - qf), Y, hash, acc, Kcap, touch);
            }
        }
        const int left = __shfl_up_sync(0xffffffffu, item, 1);
        const bool accepted = !uni && lane > 0 && left == L.x;
        const double WL = pf, XL = fma(xd, pf, qf);
        if (!uni && !accepted) cen_emit(L.x, WL, XL, Y, hash, acc, Kcap, touch);
        {
            const double gW = __shfl_down_sync(0xffffffffu, accepted ? WL : 0.0, 1), gX = __shfl_down_sync(0xffffffffu, accepted ? XL : 0.0, 1);
            if (lane < 31) { W += gW; X += gX; }
        }
        const bool head = lane == 0 || !uni || item != left;
        const unsigned heads = __ballot_sync(0xffffffffu, head);
        const unsigned after = (heads >> 1) >> lane;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const bool can = lane + o < 32 && (after & ((1u << o) - 1u)) == 0;
            if (!__any_sync(0xffffffffu, can)) break;   // no run spans o lanes; none spans more (warp-uniform)
            const double W2 = __shfl_down_sync(0xffffffffu, W, o), X2 = __shfl_down_sync(0xffffffffu, X, o);
            if (can) { W += W2; X += X2; }
        }
        if (head) cen_emit(item, W, X, Y, hash, acc, Kcap, touch);
        L = Ln; D = Dn; g = gn;
        c += sc; r += sr;
        if (c >= gpr) { c -= gpr; ++r; }
    }
    if (ENERGY) {
        e_loc = warp_sum(e_loc);
        if (lane == 0 && e_loc != 0.0) atomicAdd(acc + 4 * (size_t)Kcap, e_loc);
    }
}

#ifndef SRM_CENTROID_DEFAULT
#define SRM_CENTROID_DEFAULT 1
#endif
// 1 = k_centroid_dense2 (default), 0 = k_centroid_dense (the first build, validated on the B200 in session 4; A/B baseline,
// tests/test_gpu_centroid.py runs both).  SRM_CENTROID_V in the environment (read once) or srm_set_variant("centroid", v).
int g_srm_centroid_v = -1;
static int centroid_variant() {
    if (g_srm_centroid_v < 0) { const char *e = getenv("SRM_CENTROID_V"); g_srm_centroid_v = e ? atoi(e) != 0 : SRM_CENTROID_DEFAULT; }
    return g_srm_centroid_v;
}

// labels: dense packed labels of the context's rows (device, nrows * n int32, 16-byte aligned); density: the same rows.
cudaError_t srm_launch_centroid_dense(cudaStream_t st, const int *labels, const float *density, SrmHash hash, SrmGrid g,
                                      double *acc, int Kcap, int want_energy, int touch) {
    const size_t groups = (size_t)g.nrows() * (size_t)(g.n >> 2);
    size_t blocks = (groups + CEN_NT - 1) / CEN_NT;
    // a few CTAs per resident slot, every thread streams several groups (measured: 4 waves 163 us, 1 wave 203 us at 8192^2;
    // SRM_CEN_WAVES, read once: measurement knob)
    static const int waves = []() { const char *e = getenv("SRM_CEN_WAVES"); const int v = e ? atoi(e) : 4; return v >= 1 && v <= 64 ? v : 4; }();
    const size_t resident = (size_t)148 * CEN_MINCTA * (size_t)waves;
    if (blocks > resident) blocks = resident;
    const int4 *l4 = reinterpret_cast<const int4 *>(labels);
    const float4 *d4 = reinterpret_cast<const float4 *>(density);
    if (!centroid_variant())
        SRM_COUNT(), k_centroid_dense<<<(unsigned)blocks, CEN_NT, 0, st>>>(l4, d4, hash, g.n, g.row0, g.nrows(), acc, Kcap, want_energy, touch);
    else if (want_energy)
        SRM_COUNT(), k_centroid_dense2<true><<<(unsigned)blocks, CEN_NT, 0, st>>>(l4, d4, hash, g.n, g.row0, g.nrows(), acc, Kcap, touch);
    else
        SRM_COUNT(), k_centroid_dense2<false><<<(unsigned)blocks, CEN_NT, 0, st>>>(l4, d4, hash, g.n, g.row0, g.nrows(), acc, Kcap, touch);
    return cudaGetLastError();
}
