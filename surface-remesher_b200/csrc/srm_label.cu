// srm_label.cu — exact Voronoi labelling of a site set on an n x n grid, sm_100a.
//
// Replaces the reference's pba2DCompute (gcvt.cu:921-978: 9 kernels, ~48 B/px, pointer-chasing
// stacks in global memory) with a from-scratch formulation that produces bit-identical labels
// (rule: SURVEY Appendix A2) but never materialises a dense label map inside the Lloyd loop:
//
//   sites (K packed int32) --k_bits--> column bitmap  bits[(y>>5)*n + x] bit (y&31)     N/8 B
//   k_carry: per column, nearest site row above / below every 32-row word                N/8 B
//   k_band (srm_band.cu): per 16-row band, candidates straight from the bitmap, per row the lower
//            envelope of the parabolas (X-x)^2+(c(x,Y)-Y)^2 with INTEGER breakpoints
//            -> run-length labels  rle[row] = {(site, first X)}  (+ fused accumulation)   ~8 B/run
//   k_row:   the same result for rows that overflow k_band's shared-memory budget (worst-case capacity)
//   k_expand (final labelling only): runs -> dense short2 labels                         4 B/px
//
// All arithmetic is integer; ties follow the reference (column tie: gcvt.cu:97-119 + :172-216;
// row tie -> smallest x: gcvt.cu:449-466).
#include "srm_common.cuh"
#include <stdlib.h>
#include "srm_envelope.cuh"
#include <algorithm>

// ------------------------------------------------------------------ sites -> bitmap

// Row-band contexts fill only their own word rows of the bitmap; what lies outside the band enters through two
// per-column edge values (the nearest site row above / below the band), which seed the band's carry scan.  This is
// what kernelPropagateInterband (gcvt.cu:121-170) hands from band to band, here straight from the replicated site list.
#define EDGE_NONE_TOP ((int)0x80808080)   // "no site above" (< 0)
#define EDGE_NONE_BOT ((int)0x7f7f7f7f)   // "no site below" (> 32767)

// One site into the per-iteration structures: bitmap bit (own rows) or band edge.
__device__ __forceinline__ void srm_place_site(int p, int n, uint32_t *bits, int *edge, int row0, int row1) {
    const int x = srm_x(p), y = srm_y(p);
    if (y >= row0 && y < row1) atomicOr(&bits[(size_t)(y >> 5) * n + x], 1u << (y & 31));
    else if (y < row0) atomicMax(&edge[x], y);
    else atomicMin(&edge[n + x], y);
}

// Initial build from a site list (srm_set_sites / srm_set_site_map); inside the loop the update kernel places the
// surviving sites of the next iteration itself (srm_lloyd.cu).  The buffers were cleared by the launcher.
__global__ void k_init_sites(const int *__restrict__ sites, const SrmCtl *__restrict__ ctl, int n, uint32_t *bits, int *edge,
                             SrmHash hash, int row0, int row1) {
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= ctl->K) return;
    const int p = sites[id];
    if (p == SRM_SENT) return;
    srm_place_site(p, n, bits, edge, row0, row1);
    srm_hash_claim(hash, (unsigned)p, id);
}

__global__ void k_fill_edge(int *edge, int n) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x < n) { edge[x] = EDGE_NONE_TOP; edge[n + x] = EDGE_NONE_BOT; }
}

void srm_launch_init_sites(cudaStream_t st, const int *sites, SrmCtl *ctl, int Kcap, int n, uint32_t *bits, size_t bits_words,
                           int *edge, SrmHash hash, int row0, int row1) {
    // bits: base pointer for absolute word-row indexing; the context's own words start at (row0 >> 5) * n
    cudaMemsetAsync(bits + (size_t)(row0 >> 5) * n, 0, bits_words * sizeof(uint32_t), st);
    cudaMemsetAsync(hash.b, 0xff, ((size_t)hash.mask + 1) * sizeof(uint4), st);
    if (edge) SRM_COUNT(), k_fill_edge<<<(n + 255) / 256, 256, 0, st>>>(edge, n);
    SRM_COUNT(), k_init_sites<<<(max(Kcap, 1) + 255) / 256, 256, 0, st>>>(sites, ctl, n, bits, edge, hash, row0, row1);
}

__device__ __forceinline__ int edge_top(const int *edge, int x) {
    if (!edge) return SRM_MARK;
    const int v = edge[x];
    return v < 0 ? SRM_MARK : v;
}
__device__ __forceinline__ int edge_bot(const int *edge, int n, int x) {
    if (!edge) return SRM_MARK;
    const int v = edge[n + x];
    return v > 32767 ? SRM_MARK : v;
}

// Every carry kernel also prepares the NEXT iteration's buffers (they are filled by the update kernels later in this
// iteration): it zeroes the bitmap words at the indices it scans, and clears the hash table, the band edges and the
// robust-path row counter grid-wide.  This replaces the per-iteration memsets and the k_bits kernel of round 1.
__device__ __forceinline__ void carry_clear_next(const SrmStep &s, int n, SrmCtl *ctl, size_t tid, size_t nthreads) {
    if (tid == 0) { ctl->ovf = 0; ctl->rle_used = 0; }
    const uint4 e = make_uint4(SRM_HEMPTY, 0xffffffffu, SRM_HEMPTY, 0xffffffffu);
    for (size_t i = tid; i <= (size_t)s.hash_next.mask; i += nthreads) s.hash_next.b[i] = e;
    if (s.edge_next)
        for (size_t x = tid; x < (size_t)n; x += nthreads) { s.edge_next[x] = EDGE_NONE_TOP; s.edge_next[n + x] = EDGE_NONE_BOT; }
}

// Generic form (any n): one thread per column, one sweep per direction.
__global__ void k_carry_any(SrmStep s, int n, short *__restrict__ up, short *__restrict__ dn, SrmCtl *ctl, int respect_stop,
                            int jbeg, int jend) {
    srm_pdl_enter();
    if (respect_stop && ctl->stop) return;
    const uint32_t *__restrict__ bits = s.bits;
    const size_t tid = ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    carry_clear_next(s, n, ctl, tid, (size_t)gridDim.x * gridDim.y * blockDim.x);
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n) return;
    if (blockIdx.y == 0) {
        int last = edge_top(s.edge, x);
#pragma unroll 8
        for (int j = jbeg; j < jend; ++j) {
            size_t o = (size_t)j * n + x;
            uint32_t w = bits[o];
            up[o] = (short)last;
            s.bits_next[o] = 0u;
            if (w) last = 32 * j + 31 - __clz(w);
        }
    } else {
        int next = edge_bot(s.edge, n, x);
#pragma unroll 8
        for (int j = jend - 1; j >= jbeg; --j) {
            size_t o = (size_t)j * n + x;
            uint32_t w = bits[o];
            dn[o] = (short)next;
            if (w) next = 32 * j + __ffs(w) - 1;
        }
    }
}

// Per column: up[j][x] = largest site row < 32j, dn[j][x] = smallest site row >= 32(j+1) (MARK if none).
// This is the in-GPU analogue of kernelPropagateInterband (gcvt.cu:121-170) and, for row-band
// sharding, what makes halo exchange unnecessary: every band scans the replicated bitmap.
// A CTA owns 32 columns; its 8 warps own 8 segments of the column's word-rows.  Each thread scans its
// segment once (words kept in registers), segment summaries meet in shared memory, then the thread
// writes both carries of its words.
// WPS words per segment are held in registers (<= 32: more spills), CARRY_SEG = 8, 16 or 32 segments per column.
template <int WPS, int CARRY_SEG>
__global__ void __launch_bounds__(32 * CARRY_SEG) k_carry(SrmStep s, int n, short *__restrict__ up, short *__restrict__ dn,
                                                          SrmCtl *ctl, int respect_stop, int jbeg) {
    __shared__ short s_last[CARRY_SEG][32], s_first[CARRY_SEG][32];
    srm_pdl_enter();
    if (respect_stop && ctl->stop) return;
    const uint32_t *__restrict__ bits = s.bits;
    carry_clear_next(s, n, ctl, (size_t)blockIdx.x * (32 * CARRY_SEG) + threadIdx.y * 32 + threadIdx.x,
                     (size_t)gridDim.x * (32 * CARRY_SEG));
    const int x = blockIdx.x * 32 + threadIdx.x, seg = threadIdx.y;
    const int j0 = jbeg + seg * WPS;
    uint32_t w[WPS];
    int last = SRM_MARK, first = SRM_MARK;
#pragma unroll
    for (int k = 0; k < WPS; ++k) {
        w[k] = bits[(size_t)(j0 + k) * n + x];
        s.bits_next[(size_t)(j0 + k) * n + x] = 0u;
        if (w[k]) {
            last = 32 * (j0 + k) + 31 - __clz(w[k]);
            if (first == SRM_MARK) first = 32 * (j0 + k) + __ffs(w[k]) - 1;
        }
    }
    s_last[seg][threadIdx.x] = (short)last;
    s_first[seg][threadIdx.x] = (short)first;
    __syncthreads();
    int cu = edge_top(s.edge, x), cd = edge_bot(s.edge, n, x);  // nearest site row above / below this segment
    for (int q = 0; q < seg; ++q) { const int v = s_last[q][threadIdx.x]; if (v != SRM_MARK) cu = v; }
    for (int q = CARRY_SEG - 1; q > seg; --q) { const int v = s_first[q][threadIdx.x]; if (v != SRM_MARK) cd = v; }
#pragma unroll
    for (int k = 0; k < WPS; ++k) {
        up[(size_t)(j0 + k) * n + x] = (short)cu;
        if (w[k]) cu = 32 * (j0 + k) + 31 - __clz(w[k]);
    }
#pragma unroll
    for (int k = WPS - 1; k >= 0; --k) {
        dn[(size_t)(j0 + k) * n + x] = (short)cd;
        if (w[k]) cd = 32 * (j0 + k) + __ffs(w[k]) - 1;
    }
}

// Same result without a register-resident segment: every thread walks its segment three times (summary, up
// carries, down carries; the re-reads hit L2).  For long columns (n > 8192 on one GPU), where 32+ words per thread
// would spill: 32 segments of nw/32 words.
template <int CARRY_SEG>
__global__ void __launch_bounds__(32 * CARRY_SEG) k_carry_loop(SrmStep s, int n, short *__restrict__ up, short *__restrict__ dn,
                                                               SrmCtl *ctl, int respect_stop, int jbeg, int wps) {
    __shared__ short s_last[CARRY_SEG][32], s_first[CARRY_SEG][32];
    srm_pdl_enter();
    if (respect_stop && ctl->stop) return;
    const uint32_t *__restrict__ bits = s.bits;
    carry_clear_next(s, n, ctl, (size_t)blockIdx.x * (32 * CARRY_SEG) + threadIdx.y * 32 + threadIdx.x,
                     (size_t)gridDim.x * (32 * CARRY_SEG));
    const int x = blockIdx.x * 32 + threadIdx.x, seg = threadIdx.y;
    const int j0 = jbeg + seg * wps;
    int last = SRM_MARK, first = SRM_MARK;
#pragma unroll 4
    for (int k = 0; k < wps; ++k) {
        const uint32_t w = bits[(size_t)(j0 + k) * n + x];
        s.bits_next[(size_t)(j0 + k) * n + x] = 0u;
        if (w) {
            last = 32 * (j0 + k) + 31 - __clz(w);
            if (first == SRM_MARK) first = 32 * (j0 + k) + __ffs(w) - 1;
        }
    }
    s_last[seg][threadIdx.x] = (short)last;
    s_first[seg][threadIdx.x] = (short)first;
    __syncthreads();
    int cu = edge_top(s.edge, x), cd = edge_bot(s.edge, n, x);
    for (int q = 0; q < seg; ++q) { const int v = s_last[q][threadIdx.x]; if (v != SRM_MARK) cu = v; }
    for (int q = CARRY_SEG - 1; q > seg; --q) { const int v = s_first[q][threadIdx.x]; if (v != SRM_MARK) cd = v; }
#pragma unroll 4
    for (int k = 0; k < wps; ++k) {
        const size_t o = (size_t)(j0 + k) * n + x;
        const uint32_t w = bits[o];
        up[o] = (short)cu;
        if (w) cu = 32 * (j0 + k) + 31 - __clz(w);
    }
#pragma unroll 4
    for (int k = wps - 1; k >= 0; --k) {
        const size_t o = (size_t)(j0 + k) * n + x;
        const uint32_t w = bits[o];
        dn[o] = (short)cd;
        if (w) cd = 32 * (j0 + k) + __ffs(w) - 1;
    }
}

void srm_launch_carry(cudaStream_t st, const SrmStep &s, int n, short *up, short *dn, SrmCtl *ctl, int respect_stop,
                      int row0, int row1) {
    // Carries of the band's own word rows only; the rest of the column is summarised by the edge values that the update
    // collected (whole-grid contexts: no edges).  8 segments per column, words in registers.
    const int jbeg = row0 >> 5, jend = row1 >> 5, nw = jend - jbeg;
    const int wps = (nw % 8) ? 0 : nw / 8;
    dim3 grid(n / 32), block(32, 8);
    if (nw % 32 == 0 && nw / 32 >= 4) {
        // long columns: 32 segments per column (1024 threads per CTA), words in registers up to 16 per thread (a
        // 32-word segment needs 206 registers: one CTA per SM and 12 % of the warps, round-2 ncu capture), looped beyond
        const int w32 = nw / 32;
#define CARRY32(W) if (w32 == W) { srm_launch_pdl(st, grid, dim3(32, 32), 0, k_carry<W, 32>, s, n, up, dn, ctl, respect_stop, jbeg); return; }
        CARRY32(4) CARRY32(8) CARRY32(16)
#undef CARRY32
        srm_launch_pdl(st, grid, dim3(32, 32), 0, k_carry_loop<32>, s, n, up, dn, ctl, respect_stop, jbeg, w32);
        return;
    }
#define CARRY_CASE(W) if (wps == W) { srm_launch_pdl(st, grid, block, 0, k_carry<W, 8>, s, n, up, dn, ctl, respect_stop, jbeg); return; }
    CARRY_CASE(1) CARRY_CASE(2) CARRY_CASE(3) CARRY_CASE(4) CARRY_CASE(8)
    if (wps > 0) {   // other band heights (work-balanced row bands): 8 segments, looped
        srm_launch_pdl(st, grid, block, 0, k_carry_loop<8>, s, n, up, dn, ctl, respect_stop, jbeg, wps);
        return;
    }
    srm_launch_pdl(st, dim3((n + 63) / 64, 2), dim3(64), 0, k_carry_any, s, n, up, dn, ctl, respect_stop, jbeg, jend);
#undef CARRY_CASE
}

// ------------------------------------------------------------------ robust row path (fallback)
//
// k_row handles ANY row with worst-case capacity (every column live): column candidates of the row from
// the bitmap -> prune -> 128 thread stacks -> 7 bridging levels -> runs (-> accumulate).  The fused band
// kernel (srm_band.cu) is the fast path; rows that overflow its shared-memory budget are appended to a
// row list and processed here.  rows == nullptr means "all rows of the band" (tests pin this path so).

#define ROW_NT 128
#define ROW_NW (ROW_NT / 32)

// Column candidates c(x,Y) of 8 consecutive columns for row Y (word-row j, bit k), straight from the bitmap and
// the carries (semantics of gcvt.cu:77-216); returns the minimum |c - Y| (SRM_BIG for a column without sites).
__device__ __forceinline__ int cand8_dist(const uint32_t *__restrict__ bits, const short *__restrict__ up,
                                          const short *__restrict__ dn, size_t o, int j, int k, int Y, int *g, short *cs) {
    const uint4 w0 = *reinterpret_cast<const uint4 *>(bits + o), w1 = *reinterpret_cast<const uint4 *>(bits + o + 4);
    const uint4 u4 = *reinterpret_cast<const uint4 *>(up + o), d4 = *reinterpret_cast<const uint4 *>(dn + o);
    const uint32_t w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    const uint32_t uu[4] = {u4.x, u4.y, u4.z, u4.w}, dd[4] = {d4.x, d4.y, d4.z, d4.w};
    int M = SRM_BIG;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int u0 = (short)((uu[q >> 1] >> ((q & 1) * 16)) & 0xffff), d0 = (short)((dd[q >> 1] >> ((q & 1) * 16)) & 0xffff);
        const uint32_t mlo = w[q] & (0xffffffffu >> (31 - k));
        const uint32_t mhi = (k == 31) ? 0u : (w[q] & (0xffffffffu << (k + 1)));
        const int U = mlo ? 32 * j + 31 - __clz(mlo) : u0;
        const int D = mhi ? 32 * j + __ffs(mhi) - 1 : d0;
        const int c = srm_choose_col(U, D, Y);
        const int d = (c == SRM_MARK) ? SRM_BIG : abs(c - Y);
        if (cs) { cs[q] = (short)c; g[q] = d; }
        M = min(M, d);
    }
    return M;
}

// One CTA per row.  P1: prune columns that are dominated from both sides by a neighbouring 8-column
// block (sound: DESIGN.md §row pass), compact the survivors.  P2: per-thread stacks over short
// segments + log2(128) bridging levels.  P3: compact the envelope to global.
__global__ void __launch_bounds__(ROW_NT) k_row(const uint32_t *__restrict__ bits, const short *__restrict__ up,
                                                const short *__restrict__ dn, int n, int row0, int nrows, SrmRle R,
                                                const int *__restrict__ rows,
                                                const int *__restrict__ count, const double2 *__restrict__ P2,
                                                const double *__restrict__ PXX, SrmHash hash,
                                                double *__restrict__ acc, int Kcap, SrmCtl *ctl,
                                                int accumulate, int want_energy, int respect_stop, int write_rle,
                                                SrmPeers peers) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned short sb_[ROW_NT], se_[ROW_NT];
    __shared__ int wtot[ROW_NW];
    __shared__ int row_total;   // runs of the row (its own word: wtot[] is still being read by slower warps)
    __shared__ int pool_off;
    srm_pdl_enter();
    if (respect_stop && ctl->stop) return;
    EnvSmem s;
    s.x = (unsigned short *)smem_raw;
    s.c = (short *)(s.x + n);
    s.S = s.c + n;
    s.sb = sb_;
    s.se = se_;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int total = rows ? *count : nrows;
    for (int q = blockIdx.x; q < total; q += gridDim.x) {
        const int r = rows ? rows[q] : q, Y = row0 + r;
        const int j = Y >> 5, kbit = Y & 31;
        const size_t wrow = (size_t)j * n;
        const int nchunks = n >> 8;
        const int c0 = (w * nchunks) / ROW_NW, c1 = ((w + 1) * nchunks) / ROW_NW;
        const int rb = c0 << 8;
        int base = rb, carryM = SRM_BIG;

        for (int c = c0; c < c1; ++c) {
            const int x0 = (c << 8) + lane * 8;
            int g[8];
            short cs[8];
            int M = cand8_dist(bits, up, dn, wrow + x0, j, kbit, Y, g, cs);
            int ML = __shfl_up_sync(0xffffffffu, M, 1), MR = __shfl_down_sync(0xffffffffu, M, 1);
            if (lane == 0) {
                if (c == c0) {
                    ML = SRM_BIG;
                    if (x0 > 0) ML = cand8_dist(bits, up, dn, wrow + x0 - 8, j, kbit, Y, nullptr, nullptr);
                } else ML = carryM;
            }
            if (lane == 31) {
                MR = SRM_BIG;
                if (x0 + 8 < n) MR = cand8_dist(bits, up, dn, wrow + x0 + 8, j, kbit, Y, nullptr, nullptr);
            }
            carryM = __shfl_sync(0xffffffffu, M, 31);
            const int TL = ML * ML + 225, TR = MR * MR + 225;  // 15^2: farthest column of an adjacent block
            unsigned live = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                int g2 = g[k] * g[k];
                bool dead = (g[k] == SRM_BIG) || (g2 >= TL && g2 > TR);
                live |= dead ? 0u : (1u << k);
            }
            int cnt = __popc(live);
            int incl = warp_incl_scan(cnt, lane);
            int o = base + incl - cnt;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (live & (1u << k)) { s.x[o] = (unsigned short)(x0 + k); s.c[o] = cs[k]; ++o; }
            base += __shfl_sync(0xffffffffu, incl, 31);
        }
        __syncwarp();

        // P2a: per-lane stack, in place over its slice of the warp's survivors
        {
            const int m = base - rb, qq = (m + 31) >> 5;
            const int beg = rb + min(lane * qq, m), end = rb + min((lane + 1) * qq, m);
            s.sb[t] = (unsigned short)beg;
            s.se[t] = (unsigned short)env_lane_stack(s, beg, end, Y, n);
        }
        __syncthreads();

        // P2b: bridge neighbouring groups, doubling the group size each level
        for (int span = 1; span < ROW_NT; span <<= 1) {
            if ((t & (2 * span - 1)) == span) env_merge(s, t - span, t, t + span, Y, n);
            __syncthreads();
        }

        // P3: compact to run-length form in this CTA's scratch row (the accumulation reads it from there), then — if the
        // labels are wanted — into the pool
        int2 *mine = R.scratch + (size_t)blockIdx.x * n;
        {
            const int b = s.sb[t], e = s.se[t], cnt = e - b;
            int incl = warp_incl_scan(cnt, lane);
            if (lane == 31) wtot[w] = incl;
            __syncthreads();
            int off = incl - cnt;
            for (int k = 0; k < w; ++k) off += wtot[k];
            int2 *out = mine + off;
            for (int i = b; i < e; ++i) out[i - b] = make_int2(srm_pack(s.x[i], s.c[i]), (int)s.S[i] + 1);
            if (t == ROW_NT - 1) {
                row_total = off + cnt;
                pool_off = write_rle ? srm_rle_alloc(R, ctl, r, off + cnt) : -1;
                if (!write_rle) R.cnt[r] = off + cnt;
            }
        }
        __syncthreads();
        if (pool_off >= 0)
            for (int i = t; i < row_total; i += ROW_NT) R.pool[pool_off + i] = mine[i];
        if (accumulate && w == 0) {
            double e_loc = acc_row(mine, row_total, P2 + srm_pfx_row(r, n), PXX + srm_pfx_row(r, n), hash, n, Y, acc,
                                   Kcap, want_energy, lane);
            if (want_energy) {
                e_loc = warp_sum(e_loc);
                if (lane == 0) atomicAdd(acc + 4 * (size_t)Kcap, e_loc);
            }
        }
        __syncthreads();
    }
    // Fused all-reduce over peer memory: this is the last kernel that adds to the accumulators of the iteration, so the
    // last CTA to get here tells every peer "my sums of iteration it are complete" (what the one-warp k_signal launch
    // did in round 1).  k_update_pos on the peers waits for these flags and then pulls the sums over NVLink.
    if (peers.world > 1) {
        __shared__ int last;
        __syncthreads();
        if (t == 0) {
            __threadfence();
            last = atomicAdd(&ctl->row_ticket, 1) == (int)gridDim.x - 1;
        }
        __syncthreads();
        if (last && t < peers.world) {
            if (t == 0) ctl->row_ticket = 0;
            const int target = (ctl->epoch << 20) | (ctl->it + 1);
            __threadfence_system();
            *(volatile int *)(peers.flags[t] + peers.rank) = target;
        }
    }
}

static size_t row_smem_bytes(int n) { return (size_t)n * 6; }  // 192 KB at n = 32768

__global__ void k_expand(SrmRle R, int n, int *__restrict__ labels);

cudaError_t srm_label_setup(int n) {
    // per function, not per context: opt in for the largest grid (contexts of different sizes coexist)
    cudaError_t e = cudaFuncSetAttribute(k_row, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_smem_bytes(32768));
    if (e != cudaSuccess) return e;
    return srm_band_setup(n);
}

int srm_row_scratch_ctas(int nrows) { (void)nrows; return 148; }

cudaError_t srm_launch_row(cudaStream_t st, const uint32_t *bits, const short *up, const short *dn, SrmGrid g, SrmRle rle,
                           const int *rows, const int *count, const double2 *P2, const double *PXX,
                           SrmHash hash, double *acc, int Kcap, SrmCtl *ctl, int accumulate, int want_energy,
                           int respect_stop, int write_rle, SrmPeers signal) {
    // rows == nullptr: every row of the band (CTAs loop over the rows); else the listed rows
    const int grid = rows ? 148 : std::min(g.nrows(), rle.scratch_ctas);
    srm_launch_pdl(st, dim3(grid), dim3(ROW_NT), row_smem_bytes(g.n), k_row, bits, up, dn, g.n, g.row0, g.nrows(), rle, rows,
                   count, P2, PXX, hash, acc, Kcap, ctl, accumulate, want_energy, respect_stop, write_rle, signal);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ runs -> dense labels

#define EXP_NT 256
#define EXP_CAP 4096   // runs of a row staged in shared memory (32 KB); longer rows are searched in global memory

// ---- TMA 1-D bulk copy global -> shared with an mbarrier (sm_90+: SASS UBLKCP + SYNCS)
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem), b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(src_gmem),
                 "r"(bytes), "r"(b)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(b),
        "r"(parity)
        : "memory");
}

// Runs -> dense labels, one CTA per row.  The row's run list ({site, first X}, sorted by X; rows allocate an even number
// of entries, so it is 16-byte aligned) is fetched with ONE bulk copy (TMA) into shared memory; every thread then fills
// groups of 4 pixels: binary search of the group's first pixel among the run starts (neighbouring threads probe the
// same entries: broadcasts), at most three advances inside the group, one 128-bit store.  4 B/px written, nothing else
// per pixel — the round-1 kernel built the row in shared memory (fill, scatter, scan: three passes over 4 B/px of
// shared memory and a block-wide scan) and reached 41 % of the HBM write rate.
__global__ void __launch_bounds__(EXP_NT) k_expand(SrmRle R, int n, int *__restrict__ labels) {
    extern __shared__ __align__(16) int2 s_runs[];
    __shared__ __align__(8) unsigned long long bar;
    const int r = blockIdx.x, t = threadIdx.x;
    if (R.off[r] < 0) return;   // the pool was exhausted: the host repeats the labelling with a larger one
    const int cnt = R.cnt[r];
    const int2 *runs = R.pool + R.off[r];
    if (cnt == 0) {   // no site at all
        for (int i = t; i < n; i += EXP_NT) labels[(size_t)r * n + i] = SRM_SENT;
        return;
    }
    if (cnt <= EXP_CAP) {
        if (t == 0) mbar_init(&bar, 1);
        __syncthreads();
        if (t == 0) bulk_load(s_runs, runs, (unsigned)(((cnt + 1) & ~1) * sizeof(int2)), &bar);
        mbar_wait(&bar, 0);
        runs = s_runs;
    }
    int4 *out = reinterpret_cast<int4 *>(labels + (size_t)r * n);
    for (int q = t; q < (n >> 2); q += EXP_NT) {
        const int x = q << 2;
        int lo = 0, hi = cnt;   // largest e with start(e) <= x; start(0) == 0
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (runs[mid].y <= x) lo = mid; else hi = mid;
        }
        int e = lo;
        int4 v;
        v.x = runs[e].x;
        if (e + 1 < cnt && runs[e + 1].y <= x + 1) ++e;
        v.y = runs[e].x;
        if (e + 1 < cnt && runs[e + 1].y <= x + 2) ++e;
        v.z = runs[e].x;
        if (e + 1 < cnt && runs[e + 1].y <= x + 3) ++e;
        v.w = runs[e].x;
        out[q] = v;
    }
}

// Two-level form: the binary search is done once per block of 32 pixels (n / 32 searches per row instead of n / 4), the
// result kept in shared memory; a group of 4 pixels starts from its block's run and walks forward (runs are ~26 pixels
// long: one or two steps).  The round-2 kernel above spends ~100 instructions and a chain of ~12 dependent shared-memory
// loads per group, which is what bounds it (2.6 TB/s against a 7.1 TB/s write-only stream, profiles/r2_stream_peaks.json).
#define EXP2_CAP 2048   // runs of a row staged in shared memory (16 KB)
__global__ void __launch_bounds__(EXP_NT) k_expand2(SrmRle R, int n, int *__restrict__ labels) {
    __shared__ __align__(16) int2 s_runs2[EXP2_CAP];
    __shared__ int s_lo[1024];   // n / 32 <= 1024
    __shared__ __align__(8) unsigned long long bar;
    const int r = blockIdx.x, t = threadIdx.x;
    if (R.off[r] < 0) return;   // the pool was exhausted: the host repeats the labelling with a larger one
    const int cnt = R.cnt[r];
    const int2 *runs = R.pool + R.off[r];
    int4 *out = reinterpret_cast<int4 *>(labels + (size_t)r * n);
    if (cnt == 0) {   // no site at all
        for (int q = t; q < (n >> 2); q += EXP_NT) out[q] = make_int4(SRM_SENT, SRM_SENT, SRM_SENT, SRM_SENT);
        return;
    }
    if (cnt <= EXP2_CAP) {
        if (t == 0) mbar_init(&bar, 1);
        __syncthreads();
        if (t == 0) bulk_load(s_runs2, runs, (unsigned)(((cnt + 1) & ~1) * sizeof(int2)), &bar);
        mbar_wait(&bar, 0);
        runs = s_runs2;
    }
    for (int b = t; b < (n >> 5); b += EXP_NT) {
        const int x = b << 5;
        int lo = 0, hi = cnt;   // largest e with start(e) <= x; start(0) == 0
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (runs[mid].y <= x) lo = mid; else hi = mid;
        }
        s_lo[b] = lo;
    }
    __syncthreads();
    auto start = [&](int e) { return e < cnt ? runs[e].y : INT_MAX; };
    for (int q = t; q < (n >> 2); q += EXP_NT) {
        const int x = q << 2;
        int e = s_lo[x >> 5];
        int nxt = start(e + 1);
        while (nxt <= x) { ++e; nxt = start(e + 1); }
        int lab = runs[e].x;
        int4 v;
        v.x = lab;
        while (nxt <= x + 1) { ++e; lab = runs[e].x; nxt = start(e + 1); }
        v.y = lab;
        while (nxt <= x + 2) { ++e; lab = runs[e].x; nxt = start(e + 1); }
        v.z = lab;
        while (nxt <= x + 3) { ++e; lab = runs[e].x; nxt = start(e + 1); }
        v.w = lab;
        out[q] = v;
    }
}

// Bitmap form: the run starts of the row become a 1 bit/px bitmap in shared memory with an exclusive popcount prefix per
// 32-pixel word, so the run of pixel x is a prefix popcount — two shared-memory words, one POPC — and the runs of the three
// following pixels are that plus their own start bits: no search and no data-dependent loop.  k_expand2 is bound by
// instruction issue (ncu, profiles/r2_ncu_streams_s14.md: 65.7 M warp instructions = 125 per 4-pixel group of a warp,
// issue slots 82 % busy, 3.6 TB/s of stores); this form needs ~30 per group.  Run lists longer than the staging area
// are read from global memory by the same code.
__global__ void __launch_bounds__(EXP_NT) k_expand3(SrmRle R, int n, int *__restrict__ labels) {
    __shared__ __align__(16) int2 s_runs3[EXP2_CAP];
    __shared__ unsigned s_bits[1024];   // n / 32 <= 1024 words: bit x set iff a run starts at pixel x
    __shared__ int s_pre[1024];         // run starts in the words before this one
    __shared__ int s_wsum[EXP_NT / 32];
    __shared__ __align__(8) unsigned long long bar;
    const int r = blockIdx.x, t = threadIdx.x, lane = t & 31, wid = t >> 5;
    if (R.off[r] < 0) return;   // the pool was exhausted: the host repeats the labelling with a larger one
    const int cnt = R.cnt[r];
    const int2 *runs = R.pool + R.off[r];
    int4 *out = reinterpret_cast<int4 *>(labels + (size_t)r * n);
    if (cnt == 0) {   // no site at all
        for (int q = t; q < (n >> 2); q += EXP_NT) out[q] = make_int4(SRM_SENT, SRM_SENT, SRM_SENT, SRM_SENT);
        return;
    }
    const bool staged = cnt <= EXP2_CAP;
    if (staged && t == 0) mbar_init(&bar, 1);
    for (int w = t; w < 1024; w += EXP_NT) s_bits[w] = 0u;
    __syncthreads();
    if (staged) {
        if (t == 0) bulk_load(s_runs3, runs, (unsigned)(((cnt + 1) & ~1) * sizeof(int2)), &bar);
        mbar_wait(&bar, 0);
        runs = s_runs3;
    }
    for (int e = t; e < cnt; e += EXP_NT) {   // starts are distinct and increasing; run 0 starts at pixel 0
        const int x = runs[e].y;
        atomicOr(&s_bits[x >> 5], 1u << (x & 31));
    }
    __syncthreads();
    {   // exclusive prefix of the popcounts over the 1024 words: 4 consecutive words per thread, warp scan, 8 warp totals
        const int c0 = __popc(s_bits[4 * t]), c1 = __popc(s_bits[4 * t + 1]), c2 = __popc(s_bits[4 * t + 2]), c3 = __popc(s_bits[4 * t + 3]);
        const int tot = c0 + c1 + c2 + c3;
        const int incl = warp_incl_scan(tot, lane);
        if (lane == 31) s_wsum[wid] = incl;
        __syncthreads();
        int base = incl - tot;
#pragma unroll
        for (int k = 0; k < EXP_NT / 32; ++k) base += (k < wid) ? s_wsum[k] : 0;
        s_pre[4 * t] = base; s_pre[4 * t + 1] = base + c0; s_pre[4 * t + 2] = base + c0 + c1; s_pre[4 * t + 3] = base + c0 + c1 + c2;
    }
    __syncthreads();
    for (int q = t; q < (n >> 2); q += EXP_NT) {
        const int x = q << 2, w = x >> 5, b = x & 31;   // b is a multiple of 4: b + 3 <= 31
        const unsigned word = s_bits[w];
        const int e0 = max(s_pre[w] + __popc(word & ((2u << b) - 1u)) - 1, 0);   // run starts at pixels <= x, minus one
        const int e1 = e0 + (int)((word >> (b + 1)) & 1u), e2 = e1 + (int)((word >> (b + 2)) & 1u), e3 = e2 + (int)((word >> (b + 3)) & 1u);
        int4 v;
        v.x = runs[e0].x;
        v.y = runs[e1].x;
        v.z = runs[e2].x;
        v.w = runs[e3].x;
        out[q] = v;
    }
}

#ifndef SRM_EXPAND_DEFAULT
#define SRM_EXPAND_DEFAULT 2   // measured on the B200 (profiles/r2_stream_kernels*.json): 1: 73.9 us against 0: 100.0 us at 8192^2
#endif
// 2 = bitmap form k_expand3 (default), 1 = two-level k_expand2, 0 = k_expand (all three run in tests/test_gpu_variants.py).  SRM_EXPAND_V in the environment (read once) or
// srm_set_variant("expand", v) (measurement tools) override the compiled default.
int g_srm_expand_v = -1;
static int expand_variant() {
    if (g_srm_expand_v < 0) { const char *e = getenv("SRM_EXPAND_V"); const int v = e ? atoi(e) : SRM_EXPAND_DEFAULT; g_srm_expand_v = v < 0 || v > 2 ? SRM_EXPAND_DEFAULT : v; }
    return g_srm_expand_v;
}

cudaError_t srm_launch_expand(cudaStream_t st, SrmRle rle, SrmGrid g, int *labels) {
    const int ev = expand_variant();
    if (ev == 2) SRM_COUNT(), k_expand3<<<g.nrows(), EXP_NT, 0, st>>>(rle, g.n, labels);
    else if (ev == 1) SRM_COUNT(), k_expand2<<<g.nrows(), EXP_NT, 0, st>>>(rle, g.n, labels);
    else SRM_COUNT(), k_expand<<<g.nrows(), EXP_NT, EXP_CAP * sizeof(int2), st>>>(rle, g.n, labels);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ jump flooding (north_star family)

// One JFA pass with step k: 3x3 stencil at +-k, key (dist^2, x, y).  Not the reference's algorithm
// (SURVEY F1); pinned by the tests to a CPU JFA with the same schedule and key.  4 pixels per thread, 128-bit stores;
// the centre row load is 128-bit, the +-k column loads are 128-bit when k % 4 == 0.
__device__ __forceinline__ void jfa_take(int cand, int X, int Y, int &best, unsigned &bd) {
    if (cand == SRM_SENT) return;
    int sx = srm_x(cand), sy = srm_y(cand);
    int dx = sx - X, dy = sy - Y;
    unsigned d = (unsigned)(dx * dx + dy * dy);
    if (best == SRM_SENT || d < bd) { best = cand; bd = d; return; }
    if (d == bd) {
        int bx = srm_x(best), by = srm_y(best);
        if (sx < bx || (sx == bx && sy < by)) best = cand;
    }
}

__global__ void __launch_bounds__(256) k_jfa(const int *__restrict__ in, int *__restrict__ out, int n, int k) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;  // group of 4 pixels
    const int X0 = (q << 2) % n, Y = (q << 2) / n;
    if (Y >= n) return;
    int best[4] = {SRM_SENT, SRM_SENT, SRM_SENT, SRM_SENT};
    unsigned bd[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = -1; j <= 1; ++j) {
        const int qy = Y + j * k;
        if (qy < 0 || qy >= n) continue;
        const int *rowp = in + (size_t)qy * n;
#pragma unroll
        for (int i = -1; i <= 1; ++i) {
            const int qx0 = X0 + i * k;
            if ((k & 3) == 0 || i == 0) {
                if (qx0 < 0 || qx0 + 3 >= n) {
                    if (qx0 + 3 < 0 || qx0 >= n) continue;
#pragma unroll
                    for (int p = 0; p < 4; ++p)
                        if (qx0 + p >= 0 && qx0 + p < n) jfa_take(rowp[qx0 + p], X0 + p, Y, best[p], bd[p]);
                } else {
                    int4 v = *reinterpret_cast<const int4 *>(rowp + qx0);
                    jfa_take(v.x, X0, Y, best[0], bd[0]);
                    jfa_take(v.y, X0 + 1, Y, best[1], bd[1]);
                    jfa_take(v.z, X0 + 2, Y, best[2], bd[2]);
                    jfa_take(v.w, X0 + 3, Y, best[3], bd[3]);
                }
            } else {
#pragma unroll
                for (int p = 0; p < 4; ++p)
                    if (qx0 + p >= 0 && qx0 + p < n) jfa_take(__ldg(rowp + qx0 + p), X0 + p, Y, best[p], bd[p]);
            }
        }
    }
    *reinterpret_cast<int4 *>(out + (size_t)Y * n + X0) = make_int4(best[0], best[1], best[2], best[3]);
}

void srm_launch_jfa_pass(cudaStream_t st, const int *in, int *out, int n, int step) {
    size_t groups = (size_t)n * n / 4;
    SRM_COUNT(), k_jfa<<<(unsigned)((groups + 255) / 256), 256, 0, st>>>(in, out, n, step);
}

__global__ void k_scatter_sites(const int *__restrict__ sites, const SrmCtl *__restrict__ ctl, int n, int *map) {
    int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= ctl->K) return;
    int p = sites[id];
    if (p != SRM_SENT) map[(size_t)srm_y(p) * n + srm_x(p)] = p;
}

void srm_launch_scatter_sites(cudaStream_t st, const int *sites, const SrmCtl *ctl, int Kcap, int n, int *map) {
    if (Kcap > 0) SRM_COUNT(), k_scatter_sites<<<(Kcap + 255) / 256, 256, 0, st>>>(sites, ctl, n, map);
}

__global__ void k_fill_int(int4 *p, size_t count4, int value) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    int4 v = make_int4(value, value, value, value);
    for (; i < count4; i += stride) p[i] = v;
}

void srm_launch_fill_int(cudaStream_t st, int *p, size_t count, int value) {
    size_t c4 = count / 4;
    unsigned blocks = (unsigned)((c4 + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    SRM_COUNT(), k_fill_int<<<blocks, 256, 0, st>>>(reinterpret_cast<int4 *>(p), c4, value);
}
