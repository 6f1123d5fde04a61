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
// Output + accumulation (this round): the fp64 prefix pair at every run end and the run's site id are fetched with
// cp.async (LDGSTS) into a ring of 32-run stages in the free part of the warp's own buffer, up to 4 stages (128 runs)
// in flight per warp, and consumed from shared memory (neighbour = previous slot: no shuffles).  Round 1 issued 62
// loads per warp and then stalled on them, five times per row.  Inside the loop the run-length rows are not written
// (nothing reads them) and the per-site "touched" bytes only when a peer pulls the sums.
//
// Compile-time switches (defaults = the measured best; profiles/r1_kernel_log.md, profiles/r2_kernel_log.md):
//   BAND_REACH    3   neighbouring 8-column blocks per side used by the band-level pruning (1: band list 1.7x longer)
//   BAND_MINCTA   4   resident CTAs per SM the register allocation aims at; BAND_C8K / BAND_CL8K = buffer and band-list
//                     capacities for n <= 8192 (5 or 6 CTAs per SM with smaller capacities: no gain / slower)
//   BAND_NST      4   stages of the accumulation ring (1 = issue, wait, consume: the round-1 behaviour)
//   SRM_PFX_TILE  1   (srm_common.cuh) rows interleaved in the fp64 prefix arrays (8: kernel -1 %, k_prefix slower)
// Removed after round 1 (measured no gain, history in git): in-chunk Gauss-Seidel sweeps, reciprocal table for the
// breakpoint division, persistent CTAs on a band ticket, the shared-memory site table with CAS-loop fp64 atomics
// (an ATOMS.CAS costs more LSU time than the global RED it replaces, profiles/r2_kernel_log.md).
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
#ifndef BAND_NST
#define BAND_NST 4         // stages (of 32 runs) of the accumulation ring
#endif
// one stage of the ring: 32 x (16 B prefix pair + 16 B hash bucket of the run's site [+ 8 B x^2 prefix on energy steps])
#define BAND_STAGE_BYTES(energy) ((energy) ? 1280 : 1024)

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

__device__ __forceinline__ unsigned load_cand_v(const uint2 *__restrict__ L, int i, int Y0, int k, int Y) {
    const uint2 e = L[i];
    const int c = row_candidate((int)(short)(e.y & 0xffffu), (int)e.y >> 16, e.x >> 16, Y0, k, Y);
    return (e.x & 0xffffu) | ((unsigned)c << 16);
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
                                                int &carryB) {
    RoundStep r;
    const unsigned vn = __shfl_down_sync(0xffffffffu, v, 1);
    const int xn = (int)(vn & 0xffffu), gn = (int)(vn >> 16) - Y, Hn = xn * xn + gn * gn;
    r.owned = valid && lane < 31;
    r.B = validn ? breakpoint(Hn - H, 2 * (xn - x), n) : n - 1;
    r.Bc = __shfl_up_sync(0xffffffffu, r.B, 1);
    if (lane == 0) r.Bc = carryB;
    carryB = __shfl_sync(0xffffffffu, r.B, 30);
    r.keep = r.owned && r.B > r.Bc && r.Bc < n - 1;
    return r;
}

// The same dominance test WITHOUT the integer division, for the rounds: element e (neighbours p < e < q in the round's
// input list) can hold a pixel only if the REAL interval (B(p,e), B(e,q)] is non-empty and meets [0, n-1], i.e.
//     N1/D1 > N0/D0,   N1 >= 0,   N0/D0 < n-1       with N1/D1 = B(e,q), N0/D0 = B(p,e) as exact fractions,
// compared by cross-multiplication in 64 bits (|N| < 2^31, D < 2^16).  This is necessary for the exact test of
// round_step (which also asks for an INTEGER in the interval), so every drop is sound, and its fixpoint is the real
// lower envelope: consecutive breakpoints strictly increasing.  Elements of that envelope whose interval holds no
// integer ("ghosts", rare: sites are several pixels apart) survive the rounds; they own no pixel, so
//   * the accumulation adds exact zeros for them (their run is empty: floor(B(g,q)) == floor(B(p,g)), and the floors of
//     all breakpoints across a chain of ghosts coincide, so the neighbouring runs keep their exact extents), and
//   * the run-length output removes them with ONE exact pass (their envelope neighbours are their list neighbours).
// Replaces ~20 instructions incl. four XU-pipe operations per element and round (28 % of the kernel's instructions
// in the round-2 ncu capture, profiles/r2_ncu_band_summary.md) by ~10 ALU instructions.
__device__ __forceinline__ bool dom_step(bool valid, bool validn, bool first, unsigned v, int Y, int lane, int n,
                                         unsigned &carryV) {
    const unsigned vn = __shfl_down_sync(0xffffffffu, v, 1);
    unsigned vp = __shfl_up_sync(0xffffffffu, v, 1);
    if (lane == 0) vp = carryV;
    carryV = __shfl_sync(0xffffffffu, v, 30);
    const int x = (int)(v & 0xffffu), g = (int)(v >> 16) - Y;
    const int xn = (int)(vn & 0xffffu), gn = (int)(vn >> 16) - Y;
    const int xp = (int)(vp & 0xffffu), gp = (int)(vp >> 16) - Y;
    const int D1 = xn - x, D0 = x - xp;                         // half denominators
    const int N1 = D1 * (xn + x) + (gn - g) * (gn + g);         // H_q - H_e
    const int N0 = D0 * (x + xp) + (g - gp) * (g + gp);         // H_e - H_p
    const bool hasp = !(first && lane == 0);
    const bool c1 = !validn || N1 >= 0;
    const bool c2 = !hasp || N0 < 2 * (n - 1) * D0;
    const bool c3 = !(validn && hasp) || (long long)N1 * (long long)D0 > (long long)N0 * (long long)D1;
    return valid && lane < 31 && c1 && c2 && c3;
}

// ---- cp.async (LDGSTS): global -> shared without a register round trip, completion by commit groups
__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async_wait_pending(int pending) {   // warp-uniform
    switch (pending) {
        case 0: cp_async_wait<0>(); break;
        case 1: cp_async_wait<1>(); break;
        case 2: cp_async_wait<2>(); break;
        default: cp_async_wait<3>(); break;
    }
}

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

template <int RPW, int C>
__global__ void __launch_bounds__(BAND_NT, BAND_MINCTA) k_band(const uint32_t *__restrict__ bits, const short *__restrict__ up,
                                                  const short *__restrict__ dn, int n, int row0, int CL,
                                                  SrmRle rle, int *ovf_rows,
                                                  const double2 *__restrict__ P2, const double *__restrict__ PXX,
                                                  SrmHash hash, double *__restrict__ acc, int Kcap,
                                                  SrmCtl *ctl, int flags, int dbg, const int *__restrict__ perm) {
    constexpr int R = BAND_NW * RPW;
    static_assert(R <= 16, "in-band bits are packed in 16 bits");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int wcnt[BAND_NW];
    srm_pdl_enter();
    if ((flags & SRM_BF_STOP) && ctl->stop) return;
    const bool accumulate = flags & SRM_BF_ACC, want_energy = flags & SRM_BF_ENERGY;

    uint2 *L = reinterpret_cast<uint2 *>(smem_raw);                       // band list: {x | inband << 16, U | D << 16}
    unsigned char *masks = reinterpret_cast<unsigned char *>(L + CL);      // live mask per 8-column block
    unsigned *buf0 = reinterpret_cast<unsigned *>(masks + ((n / 8 + 15) & ~15));  // per-warp element buffers

    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    double e_loc = 0;
    // CTA -> band: in launch order, or (band order, below) the bands that were the most expensive in an earlier
    // iteration first, so that the last wave of CTAs is made of the cheap ones
    const int bi = perm ? perm[blockIdx.x] : (int)blockIdx.x;
    const int rb = bi * R, Y0 = row0 + rb, j = Y0 >> 5, k0 = Y0 & 31;
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
        return;
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
        int m = 0, pos = 0;
        unsigned carry0 = 0;
        bool overflow = false;
        for (;;) {
            // round 0: candidates of the band list, tested against their list neighbours, appended to buf
            while (pos < mb && m + 62 <= C) {
                const int ea = pos + lane, eb = pos + 31 + lane;
                const unsigned v0 = load_cand_v(L, min(ea, mb - 1), Y0, k, Y);   // clamped index: no branch, unused if invalid
                const unsigned v1 = load_cand_v(L, min(eb, mb - 1), Y0, k, Y);
                const bool ka = dom_step(ea < mb, ea + 1 < mb, pos == 0, v0, Y, lane, n, carry0);
                const bool kb = dom_step(eb < mb, eb + 1 < mb, false, v1, Y, lane, n, carry0);
                const unsigned ba = __ballot_sync(0xffffffffu, ka), bb = __ballot_sync(0xffffffffu, kb);
                if (ka) buf[m + __popc(ba & lt)] = v0;
                m += __popc(ba);
                if (kb) buf[m + __popc(bb & lt)] = v1;
                m += __popc(bb);
                pos += 62;
            }
            __syncwarp();
            PROF_ADD(2); PROF_CNT(10, m); PROF_CNT(11, 1);
            // rounds over buf until nothing is dropped.  Two independent 31-element chunks per iteration (only the carry
            // shuffle links them): a warp's step is a chain of dependent shared-memory / shuffle / ballot latencies, and
            // the second chunk fills it.  Measured alternatives (profiles/r2_kernel_log.md): one chunk per step with
            // "dirty" flags that skip steps whose elements kept both neighbours (fewer instructions, 165 us against 148:
            // the chain per step got longer), 3 or 4 chunks (no effect).
            for (;;) {
                int wp = 0;
                unsigned carryV = 0;
                for (int base = 0; base < m; base += 62) {
                    unsigned vv[2], bal[2];
                    bool kk[2];
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int e = base + 31 * q + lane;
                        vv[q] = (e < m) ? buf[e] : 0u;
                    }
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int e = base + 31 * q + lane;
                        kk[q] = dom_step(e < m, e + 1 < m, base == 0 && q == 0, vv[q], Y, lane, n, carryV);
                    }
#pragma unroll
                    for (int q = 0; q < 2; ++q) bal[q] = __ballot_sync(0xffffffffu, kk[q]);
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
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
        // the accumulation ring lives in the free space behind the runs (rows too close to the capacity: plain loads)
        const int m4 = (m + 3) & ~3;
        const int stage_bytes = BAND_STAGE_BYTES(want_energy);
        if (overflow) {
            if (lane == 0) ovf_rows[atomicAdd(&ctl->ovf, 1)] = r;
            continue;
        }
        if (flags & SRM_BF_RLE) {
            // exact passes (integer breakpoints): drop the elements whose interval holds no pixel.  One pass removes
            // them all (see dom_step); the loop runs until a pass changes nothing, i.e. one more to confirm.
            for (;;) {
                int wp = 0, carryB = -1;
                for (int base = 0; base < m; base += 31) {
                    const int e = base + lane;
                    const unsigned v = (e < m) ? (buf[e] & 0x7fffffffu) : 0u;
                    const int x = (int)(v & 0xffffu), g = (int)(v >> 16) - Y;
                    const bool kp = round_step(e < m, e + 1 < m, v, x, x * x + g * g, Y, lane, n, carryB).keep;
                    const unsigned bal = __ballot_sync(0xffffffffu, kp);
                    __syncwarp();
                    if (kp) buf[wp + __popc(bal & lt)] = v;
                    wp += __popc(bal);
                }
                __syncwarp();
                const bool removed = wp != m;
                m = wp;
                if (!removed) break;
            }
        }
        if (!(flags & SRM_BF_RLE) && lane == 0) rle.cnt[r] = m;   // statistics only (srm_debug_counts)
        if (flags & SRM_BF_RLE) {
            // runs -> run-length pool (final labelling and the stepwise API; nothing in the loop reads it)
            int po = 0;
            if (lane == 0) po = srm_rle_alloc(rle, ctl, r, m);
            po = __shfl_sync(0xffffffffu, po, 0);
            int2 *out = rle.pool + max(po, 0);
            int carryB = -1;
            for (int base = 0; base < m; base += 62) {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int e = base + 31 * q + lane;
                    const bool valid = e < m;
                    const unsigned v = valid ? (buf[e] & 0x7fffffffu) : 0u;
                    const int x = (int)(v & 0xffffu), g = (int)(v >> 16) - Y;
                    const RoundStep st = round_step(valid, e + 1 < m, v, x, x * x + g * g, Y, lane, n, carryB);
                    if (st.owned && po >= 0) out[e] = make_int2((int)v, st.Bc + 1);
                }
            }
        }
        if (accumulate) {
            // Per-site sums of the row's runs: run e = (B(e-1), B(e)] contributes the fp64 prefix differences of d and
            // x*d (and x^2*d for the energy) to its site.  The prefix entries at the run ends and the home buckets of the
            // sites in the pixel -> id hash are fetched by cp.async into a ring of 32-run stages behind the runs in this
            // warp's buffer.
            unsigned char *ring = reinterpret_cast<unsigned char *>(buf + m4);
            const int nst = min(BAND_NST, ((C - m4) * 4) / stage_bytes);   // 0: no room for a stage
            const int nbat = (m + 31) >> 5;
            const double2 *p2 = P2 + srm_pfx_row(r, n);   // tiled layout: element x at [x * SRM_PFX_TILE]
            const double *pxx = PXX + srm_pfx_row(r, n);
            unsigned char *touched = reinterpret_cast<unsigned char *>(acc + 4 * (size_t)Kcap + 4);
            // run e with prefix entries pb (at its end) and pa (at the end of the previous run) -> its site's sums
            auto apply = [&](unsigned v, int id, double2 pb, double2 pa, double xb, double xa) {
                const int x = (int)(v & 0xffffu), g = (int)(v >> 16) - Y;
                const double W = pb.x - pa.x, X = pb.y - pa.y;
                double *a = acc + 4 * (size_t)id;
                if (!ABL(2)) {
                    atomicAdd(a, W);
                    atomicAdd(a + 1, X);
                    atomicAdd(a + 2, (double)Y * W);
                    if (flags & SRM_BF_TOUCH) touched[id] = 1;
                }
#ifdef SRM_MEASURE
                else if (W == -1.5) a[3] = X;   // keeps the loads alive when the atomics are ablated
#endif
                if (want_energy) e_loc += (xb - xa) - 2.0 * (double)x * X + (double)(x * x + g * g) * W;
            };
            auto run_end = [&](int e, unsigned v) {   // last pixel of run e (exact integer breakpoint with its successor)
                if (e + 1 >= m) return n - 1;
                const unsigned vn = buf[e + 1] & 0x7fffffffu;
                const int x = (int)(v & 0xffffu), g = (int)(v >> 16) - Y;
                const int xn = (int)(vn & 0xffffu), gn = (int)(vn >> 16) - Y;
                return breakpoint(xn * xn + gn * gn - (x * x + g * g), 2 * (xn - x), n);
            };
            double2 carryP = make_double2(0, 0);
            double carryXX = 0;
            if (nst == 0) {
                // a row whose runs leave no room for a stage: plain loads, neighbour by shuffle (the round-1 form)
                for (int b = 0; b < nbat; ++b) {
                    const int e = 32 * b + lane;
                    const bool valid = e < m;
                    const unsigned v = valid ? (buf[e] & 0x7fffffffu) : 0u;
                    const int B = valid ? run_end(e, v) : 0;
                    const double2 pb = valid ? p2[(size_t)B * SRM_PFX_TILE] : make_double2(0, 0);
                    const double xb = (valid && want_energy) ? pxx[(size_t)B * SRM_PFX_TILE] : 0;
                    const int id = valid ? max(srm_hash_find(hash, v), 0) : 0;
                    double2 pa;
                    pa.x = __shfl_up_sync(0xffffffffu, pb.x, 1); pa.y = __shfl_up_sync(0xffffffffu, pb.y, 1);
                    double xa = __shfl_up_sync(0xffffffffu, xb, 1);
                    if (lane == 0) { pa = carryP; xa = carryXX; }
                    carryP.x = __shfl_sync(0xffffffffu, pb.x, 31); carryP.y = __shfl_sync(0xffffffffu, pb.y, 31);
                    carryXX = __shfl_sync(0xffffffffu, xb, 31);
                    if (valid) apply(v, id, pb, pa, xb, xa);
                }
            } else {
            auto issue = [&](int b) {
                unsigned char *st = ring + (b % nst) * stage_bytes;
                const int e = 32 * b + lane;
                if (e < m) {
                    const unsigned v = buf[e] & 0x7fffffffu;
                    const int B = run_end(e, v);
                    if (!ABL(8)) cp_async16(st + 16 * lane, p2 + (size_t)B * SRM_PFX_TILE);
                    if (!ABL(4)) cp_async16(st + 512 + 16 * lane, hash.b + srm_hash_bucket(hash, v));
                    if (want_energy) cp_async8(st + 1024 + 8 * lane, pxx + (size_t)B * SRM_PFX_TILE);
                }
                cp_async_commit();   // one group per batch, empty ones included: uniform group accounting
            };
            for (int b = 0; b < nst; ++b) issue(b);
            for (int b = 0; b < nbat; ++b) {
                cp_async_wait_pending(nst - 1);
                __syncwarp();   // the stage holds every lane's entries
                const unsigned char *st = ring + (b % nst) * stage_bytes;
                const int e = 32 * b + lane;
                const bool valid = e < m;
                const double2 *sp = reinterpret_cast<const double2 *>(st);
                const double *sx = reinterpret_cast<const double *>(st + 1024);
                const double2 pb = (valid && !ABL(8)) ? sp[lane] : make_double2(1, 1);
                const double2 pa = lane ? ((valid && !ABL(8)) ? sp[lane - 1] : make_double2(1, 1)) : carryP;
                const double xb = (valid && want_energy) ? sx[lane] : 0, xa = lane ? ((valid && want_energy) ? sx[lane - 1] : 0) : carryXX;
                carryP = sp[31]; carryXX = want_energy ? sx[31] : 0;   // broadcast reads; only full batches have a successor
                if (valid) {
                    const unsigned v = buf[e] & 0x7fffffffu;
                    // every label is a live site, so the lookup succeeds; it leaves the home bucket for < 1 % of the sites
                    const int id = ABL(4) ? (e + 37 * r) % Kcap
                                          : max(srm_hash_find_from(hash, v, srm_hash_bucket(hash, v),
                                                                   reinterpret_cast<const uint4 *>(st + 512)[lane]), 0);
                    apply(v, id, pb, pa, xb, xa);
                }
                __syncwarp();   // every lane has read the stage before it is refilled
                if (b + nst < nbat) issue(b + nst); else cp_async_commit();
            }
            cp_async_wait<0>();
            }
        }
        __syncwarp();
        PROF_ADD(5);   // output + accumulate
        PROF_CNT(14, m);
    }
    if (accumulate && want_energy) {
        e_loc = warp_sum(e_loc);
        if (lane == 0) atomicAdd(acc + 4 * (size_t)Kcap, e_loc);
    }
}

// ---- Band order.  The grid of k_band is 1.7 waves of CTAs at the headline sizes (1024 bands for 592 resident CTAs at
// 8192^2; 512 for 296 on each of 8 GPUs at 32768^2) and a band's cost varies 2.7x with the local site density (C3: 1430
// to 3850 runs per band), so in launch order the expensive middle bands start in the second wave and the kernel ends
// with a long tail of half-empty SMs.  Longest-processing-time-first: the bands are issued by decreasing cost of an
// EARLIER iteration (runs per band, which k_band records per row anyway; sites move a few pixels per iteration, so the
// cost profile is stable), and the tail is made of the cheap bands.  List-scheduling model on the measured run counts
// (tools/sim_band_order.py): makespan -13 %.  The order is a hint: any permutation gives the same results.
//
// k_band_order: one CTA; key = cost << 12 | (4095 - band) sorted in decreasing order by a bitonic network in shared
// memory (equal costs keep the launch order; an all-equal profile, e.g. before the first iteration, is the identity).
#define BAND_ORDER_MAX 4096   // bands of one context: rows / 8 <= 32768 / 8
#define BAND_ORDER_NT 1024
__global__ void __launch_bounds__(BAND_ORDER_NT) k_band_order(const int *__restrict__ cnt, int nb, int R, int *__restrict__ perm) {
    __shared__ int key[BAND_ORDER_MAX];
    srm_pdl_enter();
    const int t = threadIdx.x;
    int P = 1;
    while (P < nb) P <<= 1;
    for (int i = t; i < P; i += BAND_ORDER_NT) {
        int k = -1;   // padding sorts last
        if (i < nb) {
            int s = 0;
            for (int r = 0; r < R; ++r) s += min(max(cnt[i * R + r], 0), 32767);
            k = (min(s, (1 << 18) - 1) << 12) | (BAND_ORDER_MAX - 1 - i);
        }
        key[i] = k;
    }
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = t; i < P; i += BAND_ORDER_NT) {
                const int q = i ^ j;
                if (q > i) {
                    const int a = key[i], b = key[q];
                    const bool desc = (i & k) == 0;
                    if ((a < b) == desc) { key[i] = b; key[q] = a; }
                }
            }
            __syncthreads();
        }
    for (int i = t; i < nb; i += BAND_ORDER_NT) perm[i] = BAND_ORDER_MAX - 1 - (key[i] & (BAND_ORDER_MAX - 1));
}

// Per-warp element buffer (entries of 4 B): must hold a row's envelope + 62 (and, when accumulating, one 1024 / 1280-byte
// stage of the prefix ring).  Rows of an n-wide grid with the BASELINE site densities have ~n/26 runs (316 at
// 8192^2/100k, 520 at 16384^2/250k, 950 at 32768^2/1M).
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
    return (size_t)CL * 8 + (size_t)((n / 8 + 15) & ~15) + (size_t)BAND_NW * band_bufcap(n) * 4;
}

template <int RPW, int C>
static cudaError_t band_setup_one(int smem) {
    return cudaFuncSetAttribute(k_band<RPW, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
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
                            int CL, SrmRle rle, int *ovf_rows, const double2 *P2, const double *PXX,
                            SrmHash hash, double *acc, int Kcap, SrmCtl *ctl, int flags, int dbg, const int *perm) {
    const int nbands = g.nrows() / (BAND_NW * RPW);
    srm_launch_pdl(st, dim3(nbands), dim3(BAND_NT), smem, k_band<RPW, C>, bits, up, dn, g.n, g.row0, CL, rle, ovf_rows, P2,
                   PXX, hash, acc, Kcap, ctl, flags, dbg, perm);
}

int srm_band_bufcap(int n) { return band_bufcap(n); }

cudaError_t srm_launch_band(cudaStream_t st, const uint32_t *bits, const short *up, const short *dn, SrmGrid g, SrmRle rle,
                            int *ovf_rows, const double2 *P2, const double *PXX, SrmHash hash,
                            double *acc, int Kcap, SrmCtl *ctl, int flags, int dbg, const int *perm, int refresh_order) {
    const int CL = band_cap(g.n);
    const size_t smem = band_smem(g.n, CL);
    const int rpw = band_rpw(g.nrows()), C = band_bufcap(g.n);
    if (perm && refresh_order) {
        const int R = BAND_NW * rpw;
        srm_launch_pdl(st, dim3(1), dim3(BAND_ORDER_NT), 0, k_band_order, (const int *)rle.cnt, g.nrows() / R, R, const_cast<int *>(perm));
    }
#define BAND_ARGS st, smem, bits, up, dn, g, CL, rle, ovf_rows, P2, PXX, hash, acc, Kcap, ctl, flags, dbg, perm
    if (C == BAND_C8K) { if (rpw == 1) band_launch_one<1, BAND_C8K>(BAND_ARGS); else band_launch_one<2, BAND_C8K>(BAND_ARGS); }
    else if (C == 1280) { if (rpw == 1) band_launch_one<1, 1280>(BAND_ARGS); else band_launch_one<2, 1280>(BAND_ARGS); }
    else { if (rpw == 1) band_launch_one<1, 1792>(BAND_ARGS); else band_launch_one<2, 1792>(BAND_ARGS); }
#undef BAND_ARGS
    return cudaGetLastError();
}
