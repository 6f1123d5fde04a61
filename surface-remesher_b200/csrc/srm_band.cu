// srm_band.cu — fused band kernel: exact labelling of R = 8 (or 16) consecutive rows straight from the
// column bitmap, run-length output and (optionally) the per-site centroid/energy accumulation, in
// one launch.  This is the hot kernel of the Lloyd loop; it replaces the reference's pba2DCompute +
// pbaCVDComputeCentroid (+ pbaCVDCalcEnergy) chain (gcvt.cu:921-978, 1008-1023, 1059-1083: ~15
// launches, ~75 B/px) and never materialises a per-pixel array.
//
// One CTA = 8 warps = one band of R = 8*RPW rows (RPW = 1 by default: 8-row bands).
//   Phase A (whole CTA, once per band): per column, from the bitmap word and the up/dn carries, the
//     nearest site row above the band (U), below it (D) and the in-band bits.  Columns that are
//     dominated for EVERY row of the band by both neighbouring 8-column blocks are dropped
//     (bounds: gmin = min over the band of |dy|, gmax <= gmin + R - 1); survivors are compacted
//     into the band list (x, U, D, inband bits) in shared memory, ordered by x.
//   Phase B (one warp per row): the row's lower envelope by DOMINANCE ROUNDS.  Element e of the current
//     list wins on the integer interval (B(e-1,e), B(e,e+1)]; if that interval is empty (or beyond the
//     grid) e is dropped.  All elements of a round are tested in parallel against the round's input
//     list (sound with stale neighbours), survivors are compacted in place in a 4 KB per-warp buffer,
//     and rounds repeat until nothing is dropped.  Round 0 reads the band list (row candidate c(x,Y) by
//     the tie rule of A2).  On Voronoi-like data the list shrinks ~3x per round (measured 1654 -> 554 ->
//     187 -> 95 -> 76 -> 70 -> 69; total work 1.9x the band list), every instruction runs on 32 lanes,
//     and there are no stacks, no merges, no capacities other than "envelope <= 993 elements per row".
//     The output pass writes the runs and, in accumulate mode, adds the fp64 prefix differences of each
//     run to its site's accumulators (centroid + energy).
// A band whose list exceeds CL entries, or a row whose envelope exceeds the buffer, is handed to the
// robust path (k_row in srm_label.cu).
//
// Compile-time switches (defaults = the measured best; every alternative was A/B-measured on a B200 unless noted,
// profiles/r1_kernel_log.md):
//   BAND_REACH    3   neighbouring 8-column blocks per side used by the band-level pruning (1: band list 1.7x longer)
//   BAND_GS0/GS1  0   in-chunk Gauss-Seidel sweeps in round 0 / later rounds (fewer passes, same time)
//   BAND_NCH      2   chunks in flight in the round loop (3, 4: no effect)
//   BAND_LUT      0   reciprocal table instead of the float division in the breakpoint (512: exact, no effect)
//   BAND_MINCTA   4   resident CTAs per SM the register allocation aims at; BAND_C8K / BAND_CL8K = buffer and band-list
//                     capacities for n <= 8192 (5 or 6 CTAs per SM with smaller capacities: no gain / slower)
//   BAND_PERSIST  0   persistent CTAs taking bands from a ticket (same results, 3 % slower)
//   BAND_SITETAB  0   per-band shared-memory site table for the accumulation (written, NOT yet run on a GPU)
//   SRM_PFX_TILE  1   (srm_common.cuh) rows interleaved in the fp64 prefix arrays (8: kernel -1 %, k_prefix slower)
// Variants are built into build/variants/ and compared in one process with tools/ab_inproc.py.
#include "srm_common.cuh"
#include "srm_envelope.cuh"
#include <stdlib.h>

#define BAND_NT 256
#define BAND_NW 8
#ifndef BAND_MINCTA
#define BAND_MINCTA 4      // resident CTAs per SM the register allocation aims at (5 needs <= 48 registers)
#endif
#ifndef BAND_C8K
#define BAND_C8K 1024      // per-warp element buffer (entries) for n <= 8192
#endif
#ifndef BAND_CL8K
#define BAND_CL8K 2816     // band-list capacity (entries) for n <= 8192
#endif

struct Col8 {          // 8 consecutive columns of one band
    int U[8], D[8];    // nearest site row above / below the band (SRM_MARK if none)
    uint32_t inb[8];   // site bits inside the band, bit k = row Y0 + k
    int gmin[8];       // lower bound of |dy| over the band's rows (SRM_BIG: no site in the column)
    int M;             // min over the 8 columns of the upper bound gmax
};

struct Raw8 { uint4 w0, w1, u4, d4; };   // 8 columns of one word row: bitmap words + up / dn carries (64 B)

__device__ __forceinline__ Raw8 load_raw8(const uint32_t *__restrict__ bits, const short *__restrict__ up,
                                          const short *__restrict__ dn, size_t o) {
    Raw8 r;
    r.w0 = *reinterpret_cast<const uint4 *>(bits + o); r.w1 = *reinterpret_cast<const uint4 *>(bits + o + 4);
    r.u4 = *reinterpret_cast<const uint4 *>(up + o); r.d4 = *reinterpret_cast<const uint4 *>(dn + o);
    return r;
}

template <int R>
__device__ __forceinline__ void unpack_col8(const Raw8 &r, int j, int k0, int Y0, Col8 &c) {
    const uint32_t w[8] = {r.w0.x, r.w0.y, r.w0.z, r.w0.w, r.w1.x, r.w1.y, r.w1.z, r.w1.w};
    const uint32_t uu[4] = {r.u4.x, r.u4.y, r.u4.z, r.u4.w}, dd[4] = {r.d4.x, r.d4.y, r.d4.z, r.d4.w};
    const uint32_t rmask = (R == 32) ? 0xffffffffu : ((1u << R) - 1u);
    c.M = SRM_BIG;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int u0 = (short)((uu[k >> 1] >> ((k & 1) * 16)) & 0xffff), d0 = (short)((dd[k >> 1] >> ((k & 1) * 16)) & 0xffff);
        const uint32_t lo = k0 ? (w[k] & ((1u << k0) - 1u)) : 0u;
        const uint32_t hi = (k0 + R < 32) ? (w[k] >> (k0 + R)) : 0u;
        const uint32_t in = (w[k] >> k0) & rmask;
        const int U = lo ? 32 * j + 31 - __clz(lo) : u0;
        const int D = hi ? Y0 + R + __ffs(hi) - 1 : d0;
        int gmin, gmax;
        if (in) { gmin = 0; gmax = R - 1; }
        else {
            const int gu = (U == SRM_MARK) ? SRM_BIG : Y0 - U, gd = (D == SRM_MARK) ? SRM_BIG : D - (Y0 + R - 1);
            gmin = min(gu, gd);
            gmax = (gmin == SRM_BIG) ? SRM_BIG : gmin + R - 1;
        }
        c.U[k] = U; c.D[k] = D; c.inb[k] = in; c.gmin[k] = gmin;
        c.M = min(c.M, gmax);
    }
}

template <int R>
__device__ __forceinline__ void load_col8(const uint32_t *__restrict__ bits, const short *__restrict__ up,
                                          const short *__restrict__ dn, size_t o, int j, int k0, int Y0, Col8 &c) {
    unpack_col8<R>(load_raw8(bits, up, dn, o), j, k0, Y0, c);
}

template <int R>
__device__ __forceinline__ int block_gmax(const uint32_t *__restrict__ bits, const short *__restrict__ up,
                                          const short *__restrict__ dn, size_t o, int j, int k0, int Y0) {
    Col8 c;
    load_col8<R>(bits, up, dn, o, j, k0, Y0, c);
    return c.M;
}

// A column with lower bound g = gmin is dead for the whole band when a candidate on its left beats it at its own
// column x (then it loses for every X <= x) and a candidate on its right does too (every X >= x; strictly, the
// smaller x wins ties).  TL / TR bound the squared distance of such candidates from above: the best of
// M_k^2 + (8k+7)^2 over the k-th neighbouring 8-column blocks (M_k = smallest upper bound gmax in the block,
// 8k+7 = its farthest column), k = 1..BAND_REACH on each side.
template <int R>
__device__ __forceinline__ unsigned live_mask(const Col8 &c, int TL, int TR) {
    unsigned live = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int g2 = c.gmin[k] * c.gmin[k];
        const bool dead = (c.gmin[k] == SRM_BIG) || (g2 >= TL && g2 > TR);
        live |= dead ? 0u : (1u << k);
    }
    return live;
}

#ifndef BAND_REACH
#define BAND_REACH 3   // neighbouring 8-column blocks per side used by the band-level pruning (1 = adjacent only)
#endif                 // simulated on Lloyd-relaxed C3 sites: 16.7 % of the columns survive with 1, 12.8 % with 2, 11.9 % with 3

// Row candidate of a band-list entry for row Y = Y0 + k (Appendix A2 column rule), branch-free: U = nearest
// site row <= Y, D = nearest > Y (in-band bits override the band-level U, D), nearer wins, tie -> D iff it lies
// in Y's 64-row band (semantics of kernelFloodDown/Up + kernelPropagateInterband + kernelUpdateVertical).
__device__ __forceinline__ int row_candidate(int U, int D, uint32_t inb, int Y0, int k, int Y) {
    const uint32_t lo = inb & (0xffffffffu >> (31 - k));
    const uint32_t hi = inb & ~(0xffffffffu >> (31 - k));
    U = lo ? Y0 + 31 - __clz(lo) : U;
    D = hi ? Y0 + __ffs(hi) - 1 : D;
    const int du = (U == SRM_MARK) ? SRM_BIG : Y - U, dd = (D == SRM_MARK) ? SRM_BIG : D - Y;
    const bool pickD = dd < du || (dd == du && (D >> SRM_TIE_BAND_SHIFT) == (Y >> SRM_TIE_BAND_SHIFT));
    return pickD ? D : U;
}

// Integer breakpoint between neighbours p < q of a row: p wins (ties included, smallest x first, reference
// kernelColor gcvt.cu:449-466) exactly for X <= B = floor((H_q - H_p) / (2 (x_q - x_p))), clamped to [-1, n-1].
// Branch-free (so that independent evaluations interleave): float quotient estimate, |error| <= 1 whenever the
// true quotient is below n, exact +-1 fix-up, clamps.  num < 0 gives estimate 0, remainder < 0, hence -1.
__device__ __forceinline__ int breakpoint(int num, int den, int n) {
    int q = __float2int_rz(__fdividef(__int2float_rz(max(num, 0)), __int2float_rn(den)));
    q = min(q, n);                       // q * den <= 32768 * 65534 < 2^31
    const int r = num - q * den;
    q += (int)(r >= den) - (int)(r < 0);
    return min(q, n - 1);
}

#ifndef BAND_PERSIST
#define BAND_PERSIST 0 // 1: persistent CTAs that take bands from a ticket (grid = resident CTAs) instead of one CTA per band
#endif
#ifndef BAND_NCH
#define BAND_NCH 2     // 31-element chunks in flight per iteration of the round loop over the element buffer
#endif
#ifndef BAND_LUT
#define BAND_LUT 0     // > 0: neighbours less than BAND_LUT columns apart take their quotient estimate from a shared-memory
#endif                 // table of reciprocals (one LDS + IMAD.HI) instead of I2F, I2F, MUFU.RCP, F2I (four XU-pipe operations)

// Same result as breakpoint(num, 2 * gap, n).  lut[g] = floor(2^32 / (2g)) + 1, so umulhi(num, lut[g]) is the true
// quotient or one more for 0 <= num < 2^31 (excess num * eps / 2^32 < 1/2); the +-1 fix-up below makes it exact.
__device__ __forceinline__ int breakpoint_gap(int num, int gap, int n, const unsigned *__restrict__ lut) {
#if BAND_LUT > 0
    if ((unsigned)gap < (unsigned)BAND_LUT) {
        const int den = 2 * gap;
        int q = (int)__umulhi((unsigned)max(num, 0), lut[gap]);
        q = min(q, n);
        const int r = num - q * den;
        q += (int)(r >= den) - (int)(r < 0);
        return min(q, n - 1);
    }
#endif
    return breakpoint(num, 2 * gap, n);
}

// Candidate of band-list entry i for row Y = Y0 + k: packed x | c << 16 and H = x^2 + (c - Y)^2.
__device__ __forceinline__ void load_cand(const uint2 *__restrict__ L, int i, int Y0, int k, int Y, unsigned &v, int &x,
                                          int &H) {
    const uint2 e = L[i];
    x = (int)(e.x & 0xffffu);
    const int c = row_candidate((int)(short)(e.y & 0xffffu), (int)e.y >> 16, e.x >> 16, Y0, k, Y);
    const int g = c - Y;
    v = (unsigned)x | ((unsigned)c << 16);
    H = x * x + g * g;
}

// One step of a dominance round over 31 consecutive elements (lane 31 is a read-only lookahead).
// Element e wins on the integer interval (Bc, B]: B = breakpoint with its successor, Bc = breakpoint with its
// predecessor (carry across steps).  It is dropped when that interval is empty or lies beyond the grid; dropping
// is sound with stale neighbours (a pair that beats e everywhere exists either way), so all elements of a round
// are tested against the round's input list, in parallel.
struct RoundStep {
    int B, Bc;
    bool owned, keep;
};
__device__ __forceinline__ RoundStep round_step(bool valid, bool validn, unsigned v, int x, int H, int Y, int lane, int n,
                                                int &carryB, const unsigned *__restrict__ lut) {
    RoundStep r;
    const unsigned vn = __shfl_down_sync(0xffffffffu, v, 1);
    const int xn = (int)(vn & 0xffffu), gn = (int)(vn >> 16) - Y, Hn = xn * xn + gn * gn;
    r.owned = valid && lane < 31;
    r.B = validn ? breakpoint_gap(Hn - H, xn - x, n, lut) : n - 1;
    r.Bc = __shfl_up_sync(0xffffffffu, r.B, 1);
    if (lane == 0) r.Bc = carryB;
    carryB = __shfl_sync(0xffffffffu, r.B, 30);
    r.keep = r.owned && r.B > r.Bc && r.Bc < n - 1;
    return r;
}

// A dominance round over one chunk, with in-chunk Gauss-Seidel: after the parallel (Jacobi) test, lanes that
// survive relink to their nearest surviving neighbours inside the chunk and are re-tested until the chunk is stable,
// so a cascade of dominated elements inside 31 neighbours collapses in one pass instead of one pass per link.
// Every test uses real candidates as dominators, so every drop is sound.  The carry handed to the next chunk is
// the breakpoint of (last owned element, lookahead element) from the Jacobi step: a valid bound either way.
template <int MAX_INNER>
__device__ __forceinline__ bool chunk_round(bool valid, bool validn, unsigned v, int x, int H, int Y, int lane, int n,
                                            int &carryB, const unsigned *__restrict__ lut) {
    const unsigned vn = __shfl_down_sync(0xffffffffu, v, 1);
    const int xn = (int)(vn & 0xffffu), gn = (int)(vn >> 16) - Y, Hn = xn * xn + gn * gn;
    const bool owned = valid && lane < 31;
    int B = validn ? breakpoint_gap(Hn - H, xn - x, n, lut) : n - 1;
    int Bc = __shfl_up_sync(0xffffffffu, B, 1);
    if (lane == 0) Bc = carryB;
    carryB = __shfl_sync(0xffffffffu, B, 30);
    bool keep = owned && B > Bc && Bc < n - 1;
    unsigned bal = __ballot_sync(0xffffffffu, keep);
    unsigned prevbal = __ballot_sync(0xffffffffu, owned);
    const bool look = __shfl_sync(0xffffffffu, (int)valid, 31) != 0;  // the lookahead element exists
#pragma unroll 1
    for (int inner = 0; inner < MAX_INNER && bal != prevbal; ++inner) {  // something was dropped: relink, re-test (uniform)
        prevbal = bal;
        const unsigned link = bal | (look ? 0x80000000u : 0u);
        const unsigned right = (lane < 31) ? (link >> (lane + 1)) : 0u;
        const unsigned left = bal & ((1u << lane) - 1u);
        const int nl = right ? lane + __ffs(right) : 0;     // nearest kept lane to the right (or the lookahead)
        const int pl = left ? 31 - __clz(left) : 0;         // nearest kept lane to the left
        const unsigned vr = __shfl_sync(0xffffffffu, v, nl);
        const int xr = (int)(vr & 0xffffu), gr = (int)(vr >> 16) - Y, Hr = xr * xr + gr * gr;
        const int Bn = right ? breakpoint_gap(Hr - H, xr - x, n, lut) : n - 1;  // against my nearest kept right neighbour
        const int Bl = __shfl_sync(0xffffffffu, Bn, pl);  // B(nearest kept left neighbour, me): just computed against me
        B = min(B, Bn);                 // both are bounds by real candidates to my right
        if (left) Bc = max(Bc, Bl);     // likewise on the left (without a kept left lane the Jacobi bound stays)
        keep = keep && B > Bc && Bc < n - 1;
        bal = __ballot_sync(0xffffffffu, keep);
    }
    return keep;
}

#ifndef BAND_SITETAB
#define BAND_SITETAB 0   // > 0 (slots, e.g. 768): per-band shared-memory table that sums a site's runs over the band's rows,
#endif                   // so that the site-id lookup and the three global fp64 REDs happen once per site per band instead
                         // of once per run (~6x fewer).  PREPARED IN ROUND 1 WITHOUT GPU TIME LEFT: compiles, never run.

#if BAND_SITETAB > 0
#define SITETAB_EMPTY 0xffffffffu   // no packed site has the top bit set (row < 32768)
struct SiteTab {
    unsigned *key;     // packed site (x | c << 16) or SITETAB_EMPTY
    double *sum;       // 3 doubles per slot: W, X, Y*W
};
__device__ __forceinline__ unsigned sitetab_slot(unsigned v) {
    return (unsigned)(((unsigned long long)(v * 2654435761u) * (unsigned)BAND_SITETAB) >> 32);   // multiplicative hash -> [0, slots)
}
// Adds a run's sums to its site's slot; false if no slot was found within a few probes (the caller then updates the
// global accumulators directly, as the table-less kernel does).
__device__ __forceinline__ bool sitetab_add(const SiteTab &T, unsigned v, double W, double X, double YW) {
    unsigned h = sitetab_slot(v);
#pragma unroll 1
    for (int probe = 0; probe < 8; ++probe) {
        const unsigned k = atomicCAS(&T.key[h], SITETAB_EMPTY, v);
        if (k == SITETAB_EMPTY || k == v) {
            atomicAdd(&T.sum[3 * h], W);
            atomicAdd(&T.sum[3 * h + 1], X);
            atomicAdd(&T.sum[3 * h + 2], YW);
            return true;
        }
        h = (h + 1 == (unsigned)BAND_SITETAB) ? 0u : h + 1;
    }
    return false;
}
#endif

// Measurement code is compiled in only with -DSRM_MEASURE (tools/prof_band.py, tools/ablate_band.py build that variant):
// per-phase clock64 counters (dbg & 1) and the ablation switches of the accumulation (dbg & 2: no atomics, dbg & 4:
// synthetic site ids, dbg & 8: no prefix loads).  The default build keeps only the band-list statistics below.
#ifdef SRM_MEASURE
#define PROF_T0() long long t0__ = (dbg & 1) ? clock64() : 0
#define PROF_ADD(slot) do { if (dbg & 1) { long long t1__ = clock64(); if (lane == 0) atomicAdd(&ctl->prof[slot], (unsigned long long)(t1__ - t0__)); t0__ = t1__; } } while (0)
#define PROF_CNT(slot, v) do { if ((dbg & 1) && lane == 0) atomicAdd(&ctl->prof[slot], (unsigned long long)(v)); } while (0)
#define ABL(bit) (dbg & (bit))
#else
#define PROF_T0() do { } while (0)
#define PROF_ADD(slot) do { } while (0)
#define PROF_CNT(slot, v) do { } while (0)
#define ABL(bit) 0
#endif
// statistics counters in SrmCtl::dbg (option "dbg_stats"): [0] max / [1] sum of the band-list length, [2] bands,
// [6] warps that took the staging-overflow fallback of Phase A
#define SRM_STAT_ADD(slot, v) do { if ((dbg & 1) && lane == 0) atomicAdd(&ctl->dbg[slot], (int)(v)); } while (0)

template <int RPW, int C, int GS0, int GS1>
__global__ void __launch_bounds__(BAND_NT, BAND_MINCTA) k_band(const uint32_t *__restrict__ bits, const short *__restrict__ up,
                                                  const short *__restrict__ dn, int n, int row0, int CL,
                                                  int2 *__restrict__ rle, int *__restrict__ rle_cnt, int *ovf_rows,
                                                  const double2 *__restrict__ P2, const double *__restrict__ PXX,
                                                  const int *__restrict__ idmap, double *__restrict__ acc, int Kcap,
                                                  SrmCtl *ctl, int accumulate, int want_energy, int respect_stop,
                                                  int dbg, int nbands) {
    constexpr int R = BAND_NW * RPW;
    static_assert(R <= 16, "in-band bits are packed in 16 bits");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int wcnt[BAND_NW];
    if (respect_stop && ctl->stop) return;

    uint2 *L = reinterpret_cast<uint2 *>(smem_raw);                       // band list: {x | inband << 16, U | D << 16}
    unsigned char *masks = reinterpret_cast<unsigned char *>(L + CL);      // live mask per 8-column block
    unsigned *buf0 = reinterpret_cast<unsigned *>(masks + ((n / 8 + 15) & ~15));  // per-warp element buffers
    const unsigned *lut = nullptr;
#if BAND_LUT > 0
    {   // reciprocal table behind the element buffers; entry 0 is never used (neighbours are at least one column apart)
        unsigned *lw = buf0 + (size_t)BAND_NW * C;
        for (int g = threadIdx.x; g < BAND_LUT; g += BAND_NT) lw[g] = g ? (unsigned)(0x100000000ull / (unsigned long long)(2 * g)) + 1u : 0u;
        lut = lw;   // published by the __syncthreads() that ends Phase A
    }
#endif

    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
#if BAND_SITETAB > 0
    SiteTab tab;
    {   // behind the element buffers and the reciprocal table; 8-byte aligned (all sizes before it are multiples of 16)
        unsigned char *tb = reinterpret_cast<unsigned char *>(buf0 + (size_t)BAND_NW * C + BAND_LUT);
        tab.sum = reinterpret_cast<double *>(tb);
        tab.key = reinterpret_cast<unsigned *>(tb + (size_t)BAND_SITETAB * 24);
    }
#endif
    double e_loc = 0;
#if BAND_PERSIST
    // Persistent form: the grid holds as many CTAs as are resident at once and every CTA takes bands from a ticket
    // until none is left, so an SM keeps its full complement of CTAs until the very end (no partial second wave).
    __shared__ int s_band;
    for (;;) {
        __syncthreads();   // every warp is done with the previous band's shared memory
        if (t == 0) s_band = atomicAdd(&ctl->band_ticket, 1);
        __syncthreads();
        const int bi = s_band;
        if (bi >= nbands) break;
#else
    {
        const int bi = blockIdx.x;
        (void)nbands;
#endif
    const int rb = bi * R, Y0 = row0 + rb, j = Y0 >> 5, k0 = Y0 & 31;
#if BAND_SITETAB > 0
    if (accumulate)   // cleared here, first used after the two barriers of Phase A
        for (int q = t; q < BAND_SITETAB; q += BAND_NT) { tab.key[q] = SITETAB_EMPTY; tab.sum[3 * q] = 0; tab.sum[3 * q + 1] = 0; tab.sum[3 * q + 2] = 0; }
#endif
    const size_t wrow = (size_t)j * n;
    const int nb = n >> 3;
    const int bw0 = (w * nb) / BAND_NW, bw1 = ((w + 1) * nb) / BAND_NW;

    PROF_T0();
    // ---- Phase A, pass 1: live mask of every 8-column block (30 owned blocks per step + one halo block each side).
    // Live columns are staged, in order, in the warp's own (still unused) element buffer so that the list can be
    // assembled by a plain copy once the per-warp offsets are known.
    uint2 *stage = reinterpret_cast<uint2 *>(buf0 + (size_t)w * C);
    constexpr int STAGE_CAP = C / 2;
    int mycount = 0;
    bool staged = true;
    constexpr int OWN = 32 - 2 * BAND_REACH;   // owned blocks per step; BAND_REACH halo blocks on each side
    for (int b0 = bw0; b0 < bw1; b0 += OWN) {
        const int b = b0 - BAND_REACH + lane;
        Col8 col;
        col.M = SRM_BIG;
        if (b >= 0 && b < nb) load_col8<R>(bits, up, dn, wrow + (size_t)b * 8, j, k0, Y0, col);
        int TL = INT_MAX, TR = INT_MAX;
#pragma unroll
        for (int k = 1; k <= BAND_REACH; ++k) {
            const int ML = __shfl_up_sync(0xffffffffu, col.M, k), MR = __shfl_down_sync(0xffffffffu, col.M, k);
            const int d2 = (8 * k + 7) * (8 * k + 7);
            TL = min(TL, ML * ML + d2);   // SRM_BIG^2 + 31^2 < 2^31
            TR = min(TR, MR * MR + d2);
        }
        unsigned live = 0;
        if (lane >= BAND_REACH && lane < 32 - BAND_REACH && b < bw1) {
            live = live_mask<R>(col, TL, TR);
            masks[b] = (unsigned char)live;
        }
        const int cnt = __popc(live);
        const int incl = warp_incl_scan(cnt, lane);
        const int tot = __shfl_sync(0xffffffffu, incl, 31);
        if (staged && mycount + tot <= STAGE_CAP) {
            int o = mycount + incl - cnt;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (live & (1u << k)) {
                    stage[o] = make_uint2((unsigned)(b * 8 + k) | (col.inb[k] << 16),
                                          ((unsigned)col.U[k] & 0xffffu) | ((unsigned)col.D[k] << 16));
                    ++o;
                }
        } else staged = false;
        mycount += tot;
    }
    if (lane == 0) wcnt[w] = mycount;
    __syncthreads();
    int mb = 0, wbase = 0;
#pragma unroll
    for (int k = 0; k < BAND_NW; ++k) { if (k < w) wbase += wcnt[k]; mb += wcnt[k]; }
    if ((dbg & 1) && t == 0) { atomicMax(&ctl->dbg[0], mb); atomicAdd(&ctl->dbg[1], mb); atomicAdd(&ctl->dbg[2], 1); }
    if (mb > CL) {  // band list does not fit: every row of the band goes to the robust path
        if (t < R) ovf_rows[atomicAdd(&ctl->ovf, 1)] = rb + t;
#if BAND_PERSIST
        continue;
#else
        return;
#endif
    }
    // Every warp assembles its own section [wbase, wbase + mycount) of the band list, so the choice between the two
    // forms below is warp-local (`staged` is warp-uniform): no flag is shared between warps.
    if (staged) {
        // ---- copy the staged entries to their place
        for (int i = lane; i < mycount; i += 32) L[wbase + i] = stage[i];
    } else {
        // ---- fallback (this warp's staging area overflowed): recompute its live columns from the masks it wrote in
        // pass 1 and write the list in order
        SRM_STAT_ADD(6, 1);
        int base = wbase;
        for (int b0 = bw0; b0 < bw1; b0 += 32) {
            const int b = b0 + lane;
            const unsigned live = (b < bw1) ? masks[b] : 0u;
            const int cnt = __popc(live);
            const int incl = warp_incl_scan(cnt, lane);
            if (live) {
                Col8 col;
                load_col8<R>(bits, up, dn, wrow + (size_t)b * 8, j, k0, Y0, col);
                int o = base + incl - cnt;
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (live & (1u << k)) {
                        L[o] = make_uint2((unsigned)(b * 8 + k) | (col.inb[k] << 16),
                                          ((unsigned)col.U[k] & 0xffffu) | ((unsigned)col.D[k] << 16));
                        ++o;
                    }
            }
            base += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    __syncthreads();
    PROF_ADD(0);   // phase A

    // ---- Phase B: warp w computes the envelopes of rows rb + w*RPW .. +RPW-1 by dominance rounds.
    // Every loop below handles TWO 31-element chunks per iteration: the chunks are independent up to one carry
    // shuffle, so their shared-memory / global loads, breakpoint divisions and shuffles overlap (the kernel is
    // latency bound: ncu short/long scoreboard stalls, profiles/r1_ncu_full_summary.md).
    unsigned *buf = buf0 + (size_t)w * C;
    const unsigned lt = (1u << lane) - 1u;
    for (int rr = 0; rr < RPW; ++rr) {
        const int k = w * RPW + rr, r = rb + k, Y = Y0 + k;
        int m = 0, pos = 0, carry0 = -1;
        bool overflow = false;
        for (;;) {
            // round 0: candidates of the band list, tested against their list neighbours, appended to buf
            while (pos < mb && m + 62 <= C) {
                const int ea = pos + lane, eb = pos + 31 + lane;
                const bool va = ea < mb, vb = eb < mb;
                unsigned v0 = 0, v1 = 0;
                int x0 = 0, H0 = 0, x1 = 0, H1 = 0;
                load_cand(L, min(ea, mb - 1), Y0, k, Y, v0, x0, H0);  // clamped index: no branch, result unused if !va
                load_cand(L, min(eb, mb - 1), Y0, k, Y, v1, x1, H1);
                const bool ka = chunk_round<GS0>(va, ea + 1 < mb, v0, x0, H0, Y, lane, n, carry0, lut);
                const bool kb = chunk_round<GS0>(vb, eb + 1 < mb, v1, x1, H1, Y, lane, n, carry0, lut);
                const unsigned ba = __ballot_sync(0xffffffffu, ka), bb = __ballot_sync(0xffffffffu, kb);
                if (ka) buf[m + __popc(ba & lt)] = v0;
                m += __popc(ba);
                if (kb) buf[m + __popc(bb & lt)] = v1;
                m += __popc(bb);
                pos += 62;
            }
            __syncwarp();
            PROF_ADD(2); PROF_CNT(10, m); PROF_CNT(11, 1);
            // rounds over buf until nothing is dropped
            for (;;) {
                int wp = 0, carryB = -1;
                for (int base = 0; base < m; base += 31 * BAND_NCH) {
                    // BAND_NCH independent 31-element chunks per iteration (only the carry shuffle links them)
                    unsigned vv[BAND_NCH], bal[BAND_NCH];
                    bool kk[BAND_NCH];
                    int xs[BAND_NCH], Hs[BAND_NCH];
#pragma unroll
                    for (int q = 0; q < BAND_NCH; ++q) {
                        const int e = base + 31 * q + lane;
                        vv[q] = (e < m) ? buf[e] : 0u;
                        xs[q] = (int)(vv[q] & 0xffffu);
                        const int g = (int)(vv[q] >> 16) - Y;
                        Hs[q] = xs[q] * xs[q] + g * g;
                    }
#pragma unroll
                    for (int q = 0; q < BAND_NCH; ++q) {
                        const int e = base + 31 * q + lane;
                        kk[q] = chunk_round<GS1>(e < m, e + 1 < m, vv[q], xs[q], Hs[q], Y, lane, n, carryB, lut);
                    }
#pragma unroll
                    for (int q = 0; q < BAND_NCH; ++q) bal[q] = __ballot_sync(0xffffffffu, kk[q]);
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < BAND_NCH; ++q) {
                        if (kk[q]) buf[wp + __popc(bal[q] & lt)] = vv[q];
                        wp += __popc(bal[q]);
                    }
                }
                __syncwarp();
                const bool removed = wp != m;
                PROF_CNT(12, (m + 30) / 31);   // round steps
                PROF_CNT(13, 1);               // passes
                m = wp;
                if (!removed) break;
            }
            PROF_ADD(3);   // rounds
            if (pos >= mb) break;
            if (m + 62 > C) { overflow = true; break; }  // the envelope itself does not fit
        }
        if (overflow) {
            if (lane == 0) ovf_rows[atomicAdd(&ctl->ovf, 1)] = r;
            continue;
        }
        // output pass: runs -> global run-length row, and (accumulate mode) fp64 prefix differences -> site sums
        int2 *out = rle + (size_t)r * n;
        const double2 *p2 = P2 + srm_pfx_row(r, n);   // tiled layout: element x at [x * SRM_PFX_TILE]
        const double *pxx = PXX + srm_pfx_row(r, n);
        int carryB = -1;
        double2 carryP = make_double2(0, 0);
        double carryXX = 0;
        for (int base = 0; base < m; base += 62) {
            int ee[2] = {base + lane, base + 31 + lane};
            unsigned vv[2];
            int xx[2], cc[2], HH[2];
            RoundStep st[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const bool valid = ee[q] < m;
                vv[q] = valid ? buf[ee[q]] : 0u;
                xx[q] = (int)(vv[q] & 0xffffu); cc[q] = (int)(vv[q] >> 16);
                const int g = cc[q] - Y;
                HH[q] = xx[q] * xx[q] + g * g;
                st[q] = round_step(valid, ee[q] + 1 < m, vv[q], xx[q], HH[q], Y, lane, n, carryB, lut);
                if (st[q].owned) out[ee[q]] = make_int2((int)vv[q], st[q].Bc + 1);
            }
            if (accumulate) {
                double2 pb[2];
                double xb[2];
                int id[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {  // all loads of both chunks first
                    pb[q] = (st[q].owned && !ABL(8)) ? p2[(size_t)st[q].B * SRM_PFX_TILE] : make_double2(1, 1);
                    xb[q] = (st[q].owned && want_energy) ? pxx[(size_t)st[q].B * SRM_PFX_TILE] : 0;
#if BAND_SITETAB > 0
                    id[q] = -1;   // looked up only by the runs that find no table slot
#else
                    id[q] = st[q].owned ? (ABL(4) ? (ee[q] + 37 * r) % Kcap : idmap[(size_t)cc[q] * n + xx[q]]) : 0;
#endif
                }
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    double2 pa;
                    pa.x = __shfl_up_sync(0xffffffffu, pb[q].x, 1);
                    pa.y = __shfl_up_sync(0xffffffffu, pb[q].y, 1);
                    double xa = __shfl_up_sync(0xffffffffu, xb[q], 1);
                    if (lane == 0) { pa = carryP; xa = carryXX; }
                    carryP.x = __shfl_sync(0xffffffffu, pb[q].x, 30);
                    carryP.y = __shfl_sync(0xffffffffu, pb[q].y, 30);
                    carryXX = __shfl_sync(0xffffffffu, xb[q], 30);
                    if (st[q].owned) {
                        const double W = pb[q].x - pa.x, X = pb[q].y - pa.y;
#if BAND_SITETAB > 0
                        if (!sitetab_add(tab, vv[q], W, X, (double)Y * W)) id[q] = idmap[(size_t)cc[q] * n + xx[q]];
                        double *a = acc + 4 * (size_t)max(id[q], 0);
                        if (id[q] >= 0) {
#else
                        double *a = acc + 4 * (size_t)id[q];
                        if (!ABL(2)) {
#endif
                            atomicAdd(a, W);
                            atomicAdd(a + 1, X);
                            atomicAdd(a + 2, (double)Y * W);
                            reinterpret_cast<unsigned char *>(acc + 4 * (size_t)Kcap + 4)[id[q]] = 1;
                        }
#ifdef SRM_MEASURE
                        else if (W == -1.5) a[3] = X;   // keeps the loads alive when the atomics are ablated
#endif
                        if (want_energy) e_loc += (xb[q] - xa) - 2.0 * (double)xx[q] * X + (double)(HH[q]) * W;
                    }
                }
            }
        }
        if (lane == 0) rle_cnt[r] = m;
        __syncwarp();
        PROF_ADD(5);   // output + accumulate
        PROF_CNT(14, m);
    }
#if BAND_SITETAB > 0
    if (accumulate) {   // every warp's rows are in the table: one id lookup and one set of global REDs per site of the band
        __syncthreads();
        for (int q = t; q < BAND_SITETAB; q += BAND_NT) {
            const unsigned v = tab.key[q];
            if (v == SITETAB_EMPTY) continue;
            const int id = idmap[(size_t)(v >> 16) * n + (v & 0xffffu)];
            double *a = acc + 4 * (size_t)id;
            atomicAdd(a, tab.sum[3 * q]);
            atomicAdd(a + 1, tab.sum[3 * q + 1]);
            atomicAdd(a + 2, tab.sum[3 * q + 2]);
            reinterpret_cast<unsigned char *>(acc + 4 * (size_t)Kcap + 4)[id] = 1;
        }
    }
#endif
    }   // band (loop in the persistent form)
    if (accumulate && want_energy) {
        e_loc = warp_sum(e_loc);
        if (lane == 0) atomicAdd(acc + 4 * (size_t)Kcap, e_loc);
    }
}

#ifndef BAND_GS0
#define BAND_GS0 0   // in-chunk Gauss-Seidel sweeps allowed in round 0 (band list)
#endif
#ifndef BAND_GS1
#define BAND_GS1 0   // ... and in the later rounds.  Measured on a B200 (8192^2, C3): sweeps cut the passes per row
                      // from 8.3 to 2.6-3.2 but add serial latency; (0,0) 253 us, (0,8) 257, (1,4) 253, (2,4) 278, (3,8) 295.
#endif

// Per-warp element buffer (entries of 4 B): must hold a row's envelope + 62.  Rows of an n-wide grid with the
// BASELINE site densities have ~n/26 runs (316 at 8192^2/100k, 520 at 16384^2/250k, 950 at 32768^2/1M).
static int band_bufcap(int n) { return n <= 8192 ? BAND_C8K : n <= 16384 ? 1280 : 1792; }

static int band_cap(int n) {
    // Band-list capacity.  Measured on C3-like inputs with the 3-block pruning: band list mean 0.11 n, max
    // 0.19 n entries.  Capacities are chosen for CTAs per SM (the kernel is latency bound, resident warps are what count):
    //   n <=  8192: 2816 entries -> 55 KB  -> 4 CTAs/SM (the register file allows no more)
    //   n <= 16384: 3584 entries -> 70 KB  -> 3 CTAs/SM
    //   n <= 32768: 6144 entries -> 108 KB -> 2 CTAs/SM
    // A band or row that exceeds them goes to the robust path (k_row), which has worst-case capacity.
    int cl = (3 * n) / 8;
    const int cap = n <= 8192 ? BAND_CL8K : n <= 16384 ? 3584 : 6144;
    if (cl > cap) cl = cap;
    if (cl < 512) cl = 512;
    return cl;
}

static size_t band_smem(int n, int CL) {
    return (size_t)CL * 8 + (size_t)((n / 8 + 15) & ~15) + (size_t)BAND_NW * band_bufcap(n) * 4 + (size_t)BAND_LUT * 4 +
           (size_t)BAND_SITETAB * 28;   // site table: 3 doubles + key per slot
}

template <int RPW, int C>
static cudaError_t band_setup_one(int smem) {
    return cudaFuncSetAttribute(k_band<RPW, C, BAND_GS0, BAND_GS1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
}

cudaError_t srm_band_setup(int n) {
    // The attribute is per function (and device), not per context: contexts of different sizes coexist (multires
    // levels, batches), so every instantiation is opted in for the largest grid it serves.
    (void)n;
    const int s8 = (int)band_smem(8192, band_cap(8192)), s16 = (int)band_smem(16384, band_cap(16384)),
              s32 = (int)band_smem(32768, band_cap(32768));
    cudaError_t e = band_setup_one<1, BAND_C8K>(s8);
    if (e == cudaSuccess) e = band_setup_one<2, BAND_C8K>(s8);
    if (e == cudaSuccess) e = band_setup_one<1, 1280>(s16);
    if (e == cudaSuccess) e = band_setup_one<2, 1280>(s16);
    if (e == cudaSuccess) e = band_setup_one<1, 1792>(s32);
    if (e == cudaSuccess) e = band_setup_one<2, 1792>(s32);
    return e;
}

// Band height: 8 rows (one row per warp) or 16 rows (two rows per warp, Phase A amortised over twice the rows).
// Measured on a B200 (C3): 8192^2: 231 us vs 253 us, 4096^2: 61 vs 81 us, 8192^2 on 2 GPUs: 16-row bands leave 256
// CTAs for 592 CTA slots.  8-row bands prune harder (bounds gmax <= gmin + 7) and give twice the CTAs, which matters
// more than Phase A for a latency-bound kernel; 16 stays selectable for experiments (SRM_BAND_RPW=2).
static int band_rpw(int nrows) {
    (void)nrows;
    static const int rpw = []() {   // read once per process, not per launch
        const char *env = getenv("SRM_BAND_RPW");
        return (env && (env[0] == '1' || env[0] == '2')) ? env[0] - '0' : 1;
    }();
    return rpw;
}

template <int RPW, int C>
static void band_launch_one(cudaStream_t st, size_t smem, const uint32_t *bits, const short *up, const short *dn, SrmGrid g,
                            int CL, int2 *rle, int *rle_cnt, int *ovf_rows, const double2 *P2, const double *PXX,
                            const int *idmap, double *acc, int Kcap, SrmCtl *ctl, int accumulate, int want_energy,
                            int respect_stop, int dbg) {
    const int nbands = g.nrows() / (BAND_NW * RPW);
    int grid = nbands;
#if BAND_PERSIST
    {
        static int per_sm[2][3] = {{0, 0, 0}, {0, 0, 0}}, sms = 0;   // resident CTAs per SM of this instantiation
        const int ci = C == BAND_C8K ? 0 : C == 1280 ? 1 : 2;
        if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
        if (!per_sm[RPW - 1][ci])
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[RPW - 1][ci], k_band<RPW, C, BAND_GS0, BAND_GS1>, BAND_NT, smem);
        grid = min(nbands, max(1, per_sm[RPW - 1][ci]) * max(1, sms));
    }
#endif
    SRM_COUNT(), k_band<RPW, C, BAND_GS0, BAND_GS1><<<grid, BAND_NT, smem, st>>>(
        bits, up, dn, g.n, g.row0, CL, rle, rle_cnt, ovf_rows, P2, PXX, idmap, acc, Kcap, ctl, accumulate, want_energy,
        respect_stop, dbg, nbands);
}

cudaError_t srm_launch_band(cudaStream_t st, const uint32_t *bits, const short *up, const short *dn, SrmGrid g, int2 *rle,
                            int *rle_cnt, int *ovf_rows, const double2 *P2, const double *PXX, const int *idmap,
                            double *acc, int Kcap, SrmCtl *ctl, int accumulate, int want_energy, int respect_stop,
                            int dbg) {
    const int CL = band_cap(g.n);
    const size_t smem = band_smem(g.n, CL);
    const int rpw = band_rpw(g.nrows()), C = band_bufcap(g.n);
#define BAND_ARGS st, smem, bits, up, dn, g, CL, rle, rle_cnt, ovf_rows, P2, PXX, idmap, acc, Kcap, ctl, accumulate, want_energy, respect_stop, dbg
    if (C == BAND_C8K) { if (rpw == 1) band_launch_one<1, BAND_C8K>(BAND_ARGS); else band_launch_one<2, BAND_C8K>(BAND_ARGS); }
    else if (C == 1280) { if (rpw == 1) band_launch_one<1, 1280>(BAND_ARGS); else band_launch_one<2, 1280>(BAND_ARGS); }
    else { if (rpw == 1) band_launch_one<1, 1792>(BAND_ARGS); else band_launch_one<2, 1792>(BAND_ARGS); }
#undef BAND_ARGS
    return cudaGetLastError();
}
