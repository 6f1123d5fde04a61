// srm_lloyd.cu — centroid reduction, site update, energy/convergence control, sm_100a.
//
// Replaces pbaCVDComputeWeightedPrefix / pbaCVDComputeCentroid / pbaCVDUpdateSites /
// pbaCVDCalcEnergy and the control block of gCVT (gcvt.cu:995-1034, 1059-1083, 1105-1142).
//
// Design: the density never changes during a gCVT call, so (like the reference, gcvt.cu:514-560,
// but in fp64 instead of fp32) the row-inclusive prefix sums of d, x*d and x^2*d are built once.
// A Lloyd iteration then needs NO per-pixel work for the centroid: every run of the run-length
// labels contributes prefix[end] - prefix[start-1] to its site (sum y*d = Y * sum d within a row),
// and the CVT energy follows from the same three sums.  Accumulators are fp64 per site
// (W, X, Y, pad) and are all-reduced across row bands by the caller when sharded.
#include "srm_common.cuh"
#include "srm_envelope.cuh"
#include <algorithm>
#include <stdlib.h>

#ifndef SRM_PREFIX_DEFAULT
#define SRM_PREFIX_DEFAULT 1   // measured on the B200 (profiles/r2_stream_kernels.json): 0.319 ms against 0.572 ms at 8192^2
#endif

// ------------------------------------------------------------------ prefix sums (once per call)

#define PFX_NT 256
// 256-bit global stores (sm_100+: STG.E.ENL2.256): a thread's four consecutive prefix entries leave as full 32-byte
// sectors (P2: two stores of two double2 each, PXX: one store) instead of four 16-byte and four 8-byte partial-sector
// stores that L2 has to merge.
__device__ __forceinline__ void st_global_256(double *p, double a, double b, double c, double d) {
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

template <bool WIDE>
__global__ void __launch_bounds__(PFX_NT) k_prefix(const float *__restrict__ density, int n, double2 *__restrict__ P2,
                                                   double *__restrict__ PXX) {
    __shared__ double sw[3][PFX_NT / 32];
    const int r = blockIdx.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
    const float *d = density + (size_t)r * n;
    double cW = 0, cX = 0, cXX = 0;
    for (int base = 0; base < n; base += 4 * PFX_NT) {
        const int x = base + 4 * t;
        const bool in = x < n;
        float4 v = in ? *reinterpret_cast<const float4 *>(d + x) : make_float4(0, 0, 0, 0);
        double a[4] = {v.x, v.y, v.z, v.w};
        double W[4], X[4], XX[4];
        double rw = 0, rx = 0, rxx = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double px = (double)(x + k);
            rw += a[k]; rx += px * a[k]; rxx += px * px * a[k];
            W[k] = rw; X[k] = rx; XX[k] = rxx;
        }
        double iw = rw, ix = rx, ixx = rxx;  // warp inclusive scan of the per-thread totals
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double yw = __shfl_up_sync(0xffffffffu, iw, o), yx = __shfl_up_sync(0xffffffffu, ix, o),
                   yxx = __shfl_up_sync(0xffffffffu, ixx, o);
            if (lane >= o) { iw += yw; ix += yx; ixx += yxx; }
        }
        if (lane == 31) { sw[0][w] = iw; sw[1][w] = ix; sw[2][w] = ixx; }
        __syncthreads();
        double ew = cW + (iw - rw), ex = cX + (ix - rx), exx = cXX + (ixx - rxx);
        double tw = 0, tx = 0, txx = 0;
#pragma unroll
        for (int k = 0; k < PFX_NT / 32; ++k) {
            if (k < w) { ew += sw[0][k]; ex += sw[1][k]; exx += sw[2][k]; }
            tw += sw[0][k]; tx += sw[1][k]; txx += sw[2][k];
        }
        if (in) {
            const size_t o = srm_pfx_row(r, n) + (size_t)x * SRM_PFX_TILE;   // tiled layout, srm_common.cuh
            if (WIDE && SRM_PFX_TILE == 1) {
                double *p2 = reinterpret_cast<double *>(P2 + o);   // 64-byte aligned: x is a multiple of 4
                st_global_256(p2, ew + W[0], ex + X[0], ew + W[1], ex + X[1]);
                st_global_256(p2 + 4, ew + W[2], ex + X[2], ew + W[3], ex + X[3]);
                st_global_256(PXX + o, exx + XX[0], exx + XX[1], exx + XX[2], exx + XX[3]);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    P2[o + (size_t)k * SRM_PFX_TILE] = make_double2(ew + W[k], ex + X[k]);
                    PXX[o + (size_t)k * SRM_PFX_TILE] = exx + XX[k];
                }
            }
        }
        cW += tw; cX += tx; cXX += txx;
        __syncthreads();
    }
}

// 1 = 256-bit stores (default), 0 = 128/64-bit stores (the round-1 form; A/B baseline, kept for tests/test_gpu_variants.py).  SRM_PREFIX_V in the environment (read
// once) or srm_set_variant("prefix", v) (measurement tools) override the compiled default.
int g_srm_prefix_v = -1;
static int prefix_variant() {
    if (g_srm_prefix_v < 0) { const char *e = getenv("SRM_PREFIX_V"); g_srm_prefix_v = e ? atoi(e) != 0 : SRM_PREFIX_DEFAULT; }
    return g_srm_prefix_v;
}

void srm_launch_prefix(cudaStream_t st, const float *density_band, SrmGrid g, double2 *P2, double *PXX) {
    if (prefix_variant()) SRM_COUNT(), k_prefix<true><<<g.nrows(), PFX_NT, 0, st>>>(density_band, g.n, P2, PXX);
    else SRM_COUNT(), k_prefix<false><<<g.nrows(), PFX_NT, 0, st>>>(density_band, g.n, P2, PXX);
}

// ------------------------------------------------------------------ per-run accumulation

#define ACC_NT 128
// Robust-path accumulation: one warp per listed row (rows == nullptr: all rows of the band).
__global__ void __launch_bounds__(ACC_NT) k_acc(SrmRle R,
                                                const double2 *__restrict__ P2, const double *__restrict__ PXX,
                                                SrmHash hash, int n, int row0, int nrows,
                                                double *__restrict__ acc, int Kcap, const int *__restrict__ rows,
                                                const int *__restrict__ count, const SrmCtl *__restrict__ ctl,
                                                int want_energy, int respect_stop) {
    if (respect_stop && ctl->stop) return;
    const int lane = threadIdx.x & 31;
    const int total = rows ? *count : nrows;
    const int nwarps = gridDim.x * (ACC_NT / 32);
    double e_loc = 0;
    for (int q = blockIdx.x * (ACC_NT / 32) + (threadIdx.x >> 5); q < total; q += nwarps) {
        const int r = rows ? rows[q] : q;
        if (R.off[r] < 0) continue;   // pool exhausted for this row (flagged in SrmCtl::rle_fail)
        e_loc += acc_row(R.pool + R.off[r], R.cnt[r], P2 + srm_pfx_row(r, n), PXX + srm_pfx_row(r, n), hash, n, row0 + r, acc,
                         Kcap, want_energy, lane);
    }
    if (want_energy) {
        e_loc = warp_sum(e_loc);
        if (lane == 0 && e_loc != 0.0) atomicAdd(acc + 4 * (size_t)Kcap, e_loc);
    }
}

void srm_launch_acc(cudaStream_t st, SrmRle rle, const double2 *P2, const double *PXX,
                    SrmHash hash, SrmGrid g, double *acc, int Kcap, const int *rows, const int *count,
                    const SrmCtl *ctl, int want_energy, int respect_stop) {
    const int rows_per_block = ACC_NT / 32;
    const int grid = rows ? 148 : (g.nrows() + rows_per_block - 1) / rows_per_block;
    SRM_COUNT(), k_acc<<<grid, ACC_NT, 0, st>>>(rle, P2, PXX, hash, g.n, g.row0, g.nrows(), acc, Kcap, rows, count, ctl,
                                   want_energy, respect_stop);
}

// ------------------------------------------------------------------ block-count scan helper

// Exclusive scan of nb block counts by one CTA; total -> *total_out.
__global__ void __launch_bounds__(1024) k_scan_counts(const int *__restrict__ cnt, int *__restrict__ off, int nb,
                                                      int *total_out) {
    __shared__ int ws[32];
    __shared__ int carry_s;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        int i = base + t;
        int v = i < nb ? cnt[i] : 0;
        int incl = warp_incl_scan(v, lane);
        if (lane == 31) ws[w] = incl;
        __syncthreads();
        int wo = 0, tot = 0;
        for (int k = 0; k < 32; ++k) { if (k < w) wo += ws[k]; tot += ws[k]; }
        int c = carry_s;
        if (i < nb) off[i] = c + wo + incl - v;
        __syncthreads();
        if (t == 0) carry_s = c + tot;
        __syncthreads();
    }
    if (t == 0) *total_out = carry_s;
}

void srm_launch_scan_counts(cudaStream_t st, const int *cnt, int *off, int nb, int *total_out) {
    SRM_COUNT(), k_scan_counts<<<1, 1024, 0, st>>>(cnt, off, nb, total_out);
}

// ------------------------------------------------------------------ site update

// kernelUpdateSites (gcvt.cu:753-781): centroid, over-relaxation, round, clamp, reject.  The float
// expression is written with the roundings nvcc produced for the reference (FADD, FFMA, FADD,
// F2I.TRUNC; checked in the SASS of oracle/_ref).  Collisions: every site claims its target pixel in
// the NEXT iteration's pixel -> id hash with atomicMin(id); the smallest id survives and the others
// become holes (SRM_SENT) in the list — the reference merges sites the same way, by overwriting one
// pixel (gcvt.cu:779-780).  The list is never compacted, so ids (accumulator slots) are stable and
// identical on every rank.  The survivors are then placed into the next iteration's site bitmap / band
// edges, so no separate "sites -> bitmap" kernel runs inside the loop.
__device__ __forceinline__ int ld_volatile_int(const int *p) { return *(const volatile int *)p; }

// Fused all-reduce (row bands on several GPUs): every rank signals "my accumulators of iteration it are complete"
// into every peer's flag array (k_signal, after the band kernel); k_update_pos waits for all arrivals and then pulls
// each site's partial sums straight from the peers' accumulators over NVLink (cache-bypassing loads, fixed rank
// order, so every rank computes bit-identical totals and the replicated update stays consistent).  Accumulators are
// double-buffered by iteration parity; a buffer is cleared one iteration after its use, when every peer has
// provably finished reading it (they have signalled the next iteration).
__global__ void k_signal(SrmCtl *ctl, SrmPeers p, int respect_stop) {
    srm_pdl_enter();
    if (respect_stop && ctl->stop) return;
    const int target = (ctl->epoch << 20) | (ctl->it + 1);
    __threadfence_system();
    if ((int)threadIdx.x < p.world) *(volatile int *)(p.flags[threadIdx.x] + p.rank) = target;
}

void srm_launch_signal(cudaStream_t st, SrmCtl *ctl, SrmPeers peers, int respect_stop) {
    if (peers.world > 1) srm_launch_pdl(st, dim3(1), dim3(32), 0, k_signal, ctl, peers, respect_stop);
}

// nzbits / maskbits: one bit per pixel, word (y * n + x) >> 5, bit x & 31: "density != 0" (gcvt.cu:777) and "constraint
// pixel" (gcvt.cu:776).  Full-grid bitmaps of N/8 bytes each replace the replicated 4 B/px density and 1 B/px mask of
// round 1: the update is replicated on every rank and may look at any pixel.
__device__ __forceinline__ bool srm_bit(const uint32_t *__restrict__ b, size_t i) { return (b[i >> 5] >> (i & 31)) & 1u; }

// Owner-computes form of the fused all-reduce (r2): with W ranks, rank r computes the new position only of the sites with
// id in [K r / W, K (r+1) / W) — it pulls their partial sums from the ranks that contributed (remote loads over NVLink,
// 1/W of what every rank pulled in round 1) — and stores the result into the newpos array of EVERY rank (remote stores,
// coalesced).  The last block to finish tells every peer "my slice of newpos is complete"; k_update_claim waits for all
// slices and then every rank runs the (cheap, replicated) claim + resolve on the complete list, so the site lists stay
// identical without a broadcast of the list itself.
__device__ __forceinline__ bool wait_flags(const int *flags, int world, int target, SrmCtl *ctl) {
    for (int q = 0; q < world; ++q) {
        long long spins = 0;
        while (ld_volatile_int(flags + q) < target)
            if (++spins > (1ll << 28)) {   // fail-safe: never hang the GPU.  Stop the loop (every later kernel returns
                ctl->p2p_timeout = 1;      // at once) and let the host report it (fetch_ctl -> SRM_ERR_CUDA)
                ctl->stop = 1;
                return false;
            }
    }
    return true;
}

__global__ void k_update_pos(const int *__restrict__ sites, const double *acc, const uint32_t *__restrict__ nzbits,
                             const uint32_t *__restrict__ maskbits, int n, SrmCtl *ctl, int *__restrict__ newpos,
                             SrmHash claim, int respect_stop, SrmPeers peers) {
    __shared__ int s_last;
    srm_pdl_enter();
    if (respect_stop && ctl->stop) return;
    const bool pull = peers.world > 1;                  // the sums live in the peers' accumulators
    const bool shared_update = pull && peers.owner;     // ... and this rank updates only its slice of the ids
    if (pull) {  // wait for every rank's accumulators of this iteration
        if (threadIdx.x == 0) wait_flags(peers.flags_local, peers.world, (ctl->epoch << 20) | (ctl->it + 1), ctl);
        __syncthreads();
        if (ctl->p2p_timeout) return;   // a peer never arrived: do not update from partial sums
    }
    const int K = ctl->K;
    // this rank's slice of the ids (everything on a single GPU / with the NCCL all-reduce)
    const int lo = shared_update ? (int)((long long)K * peers.rank / peers.world) : 0;
    const int hi = shared_update ? (int)((long long)K * (peers.rank + 1) / peers.world) : K;
    const int id = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (id < hi) {
        const int p = sites[id];
        int np = SRM_SENT;
        if (p != SRM_SENT) {
            const int tx = srm_x(p), ty = srm_y(p);
            int rx = tx, ry = ty;
            if (!(maskbits && srm_bit(maskbits, (size_t)ty * n + tx))) {
                double sW = 0, sX = 0, sY = 0;
                if (pull) {
                    // every rank marks the sites it contributed to (one byte per site, behind its accumulators): a site's
                    // cell spans one or two bands, so only those ranks' sums are pulled over NVLink.  Remote loads cost
                    // ~2 us each, so they are issued in independent batches of 8 ranks (all "touched" bytes, then all sums)
                    // and added in rank order (bit-identical totals whoever computes them).
                    for (int q0 = 0; q0 < peers.world; q0 += 8) {
                        const double *base[8];
                        unsigned char tch[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int q = min(q0 + k, peers.world - 1);
                            base[k] = peers.acc[q] + (size_t)peers.parity * peers.stride;
                            tch[k] = (q0 + k < peers.world)
                                         ? __ldcv(reinterpret_cast<const unsigned char *>(base[k] + 4 * (size_t)peers.kcap + 4) + id)
                                         : (unsigned char)0;
                        }
                        double2 wx[8];
                        double yy[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            wx[k] = make_double2(0, 0); yy[k] = 0;
                            if (tch[k]) {
                                const double *a = base[k] + 4 * (size_t)id;
                                wx[k] = __ldcv(reinterpret_cast<const double2 *>(a));
                                yy[k] = __ldcv(a + 2);
                            }
                        }
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            if (tch[k]) { sW += wx[k].x; sX += wx[k].y; sY += yy[k]; }
                    }
                } else {
                    const double *a = acc + 4 * (size_t)id;
                    sW = a[0]; sX = a[1]; sY = a[2];
                }
                const float pW = (float)sW, pX = (float)sX, pY = (float)sY;
                const float omega = ctl->omega;
                const float _x = __fdiv_rn(pX, pW), _y = __fdiv_rn(pY, pW);
                const float fx = __fadd_rn(__fmaf_rn(__fsub_rn(_x, (float)tx), omega, (float)tx), 0.5f);
                const float fy = __fadd_rn(__fmaf_rn(__fsub_rn(_y, (float)ty), omega, (float)ty), 0.5f);
                int cx = __float2int_rz(fx), cy = __float2int_rz(fy);  // NaN -> 0, like F2I.TRUNC
                cx = max(min(cx, n - 1), 0);
                cy = max(min(cy, n - 1), 0);
                if (srm_bit(nzbits, (size_t)cy * n + cx)) { rx = cx; ry = cy; }
            }
            np = srm_pack(rx, ry);
        }
        if (shared_update) {
            // newpos lives behind the accumulator pair of every rank (one IPC mapping covers both)
            for (int q = 0; q < peers.world; ++q)
                reinterpret_cast<int *>(const_cast<double *>(peers.acc[q]) + 2 * peers.stride)[id] = np;
        } else {
            newpos[id] = np;
            if (np != SRM_SENT) srm_hash_claim(claim, (unsigned)np, id);
        }
    }
    if (shared_update) {   // the last block to finish publishes this rank's slice
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            s_last = atomicAdd(&ctl->pos_ticket, 1) == (int)gridDim.x - 1;
        }
        __syncthreads();
        if (s_last && (int)threadIdx.x < peers.world) {
            if (threadIdx.x == 0) ctl->pos_ticket = 0;
            const int target = (ctl->epoch << 20) | (ctl->it + 1);
            __threadfence_system();
            *(volatile int *)(peers.flags[threadIdx.x] + 64 + peers.rank) = target;
        }
    }
}

// Row bands with the fused all-reduce: claims of ALL sites, once every rank's slice of newpos has arrived.
__global__ void k_update_claim(const int *newpos, SrmCtl *ctl, SrmHash claim, int respect_stop, SrmPeers peers) {
    srm_pdl_enter();
    if (respect_stop && ctl->stop) return;
    if (threadIdx.x == 0) wait_flags(peers.flags_local + 64, peers.world, (ctl->epoch << 20) | (ctl->it + 1), ctl);
    __syncthreads();
    if (ctl->p2p_timeout) return;
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= ctl->K) return;
    const int np = ld_volatile_int(newpos + id);   // written by peers: not through the read-only path
    if (np != SRM_SENT) srm_hash_claim(claim, (unsigned)np, id);
}

#define UPD_NT 256
// Resolve the claims, clear the accumulators, and — in the last block to finish — run the loop control of
// gCVT (gcvt.cu:1116-1140) on device: latch the energy, count the iteration, every 10th iteration update omega
// and apply the stopping rule.
__global__ void __launch_bounds__(UPD_NT) k_update_resolve(const int *__restrict__ newpos, SrmHash claim, int n, int row0,
                                                           int row1, uint32_t *bits_next, int *edge_next, SrmCtl *ctl,
                                                           int *__restrict__ sites_out, double *acc, int Kcap,
                                                           int want_energy, int stop_rule, int respect_stop,
                                                           SrmPeers peers) {
    __shared__ int is_last;
    srm_pdl_enter();
    if (respect_stop && ctl->stop) return;  // set only by a previous launch's last block
    if (ctl->p2p_timeout) return;           // k_update_pos gave up waiting for a peer: leave the sites as they are
    const int id = blockIdx.x * UPD_NT + threadIdx.x;
    int alive = 0;
    if (id < ctl->K) {
        const int p = newpos[id];
        if (p != SRM_SENT) alive = srm_hash_find(claim, (unsigned)p) == id;
        sites_out[id] = alive ? p : SRM_SENT;
        if (alive) {   // into the next iteration's bitmap (own rows) or band edges
            const int x = srm_x(p), y = srm_y(p);
            if (y >= row0 && y < row1) atomicOr(&bits_next[(size_t)(y >> 5) * n + x], 1u << (y & 31));
            else if (y < row0) atomicMax(&edge_next[x], y);
            else atomicMin(&edge_next[n + x], y);
        }
        // clear for a later iteration: the buffer just used (single GPU), or the other one of the pair (peers may
        // still be reading the current one; the other one was last read an iteration ago)
        double *abase = acc + (peers.world > 1 ? (size_t)(peers.parity ^ 1) * peers.stride : 0);
        double *a = abase + 4 * (size_t)id;
        reinterpret_cast<unsigned char *>(abase + 4 * (size_t)Kcap + 4)[id] = 0;
        a[0] = 0; a[1] = 0; a[2] = 0; a[3] = 0;
    }
    const int c = __syncthreads_count(alive);
    if (threadIdx.x == 0) {
        if (c) atomicAdd(&ctl->live_acc, c);
        __threadfence();
        is_last = atomicAdd(&ctl->ticket, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last || threadIdx.x >= 32) return;   // the first warp of the last block runs the control
    __threadfence();
    const int lane = threadIdx.x;
    if (peers.world > 1) {
        if (want_energy) {
            // one remote load per lane, then a sum in rank order (identical on every rank)
            double e = 0;
            for (int q0 = 0; q0 < peers.world; q0 += 32) {
                const int q = q0 + lane;
                const double eq = (q < peers.world) ? __ldcv(peers.acc[q] + (size_t)peers.parity * peers.stride + 4 * (size_t)Kcap) : 0.0;
                const int cnt = min(32, peers.world - q0);
                for (int k = 0; k < cnt; ++k) e += __shfl_sync(0xffffffffu, eq, k);
            }
            if (lane == 0) ctl->E = (float)(e / ((double)n * (double)n)) * ctl->escale;
        }
        if (lane == 0) acc[(size_t)(peers.parity ^ 1) * peers.stride + 4 * (size_t)Kcap] = 0;
    } else if (lane == 0) {
        double *acc_energy = acc + 4 * (size_t)Kcap;
        if (want_energy) {
            ctl->E = (float)(acc_energy[0] / ((double)n * (double)n)) * ctl->escale;  // * powf(2, 2 level), gcvt.cu:1082
            acc_energy[0] = 0;
        }
    }
    if (lane != 0) return;
    ctl->nlive = atomicExch(&ctl->live_acc, 0);
    ctl->ticket = 0;
    const int it = ctl->it + 1;
    ctl->it = it;
    if (it % 10 == 0) {
        const float diffEnergy = ctl->lastE - ctl->E;
        const float gradientEnergy = (float)((double)diffEnergy / 10.0);
        const double om = 1.0 + (double)diffEnergy;
        ctl->omega = (float)(om < 2.0 ? om : 2.0);
        if (stop_rule && (double)gradientEnergy < ctl->thresh) ctl->stop = 1;
        else ctl->lastE = ctl->E;
    }
}

void srm_launch_update(cudaStream_t st, const int *sites_in, int *sites_out, double *acc, const uint32_t *nzbits,
                       const uint32_t *maskbits, SrmGrid g, SrmCtl *ctl, int Kcap, int *newpos, const SrmStep &s,
                       int want_energy, int stop_rule, int respect_stop, SrmPeers peers) {
    // acc: single GPU: the accumulator buffer; peers: the BASE of this rank's buffer pair
    const int k1 = Kcap > 0 ? Kcap : 1;
    if (peers.world > 1 && peers.owner) {   // owner computes its slice of the ids, then every rank claims all of them
        const int slice = k1 / peers.world + 1;
        srm_launch_pdl(st, dim3((slice + 255) / 256), dim3(256), 0, k_update_pos, sites_in, (const double *)acc, nzbits, maskbits, g.n,
                       ctl, newpos, s.hash_next, respect_stop, peers);
        srm_launch_pdl(st, dim3((k1 + 255) / 256), dim3(256), 0, k_update_claim, (const int *)newpos, ctl, s.hash_next, respect_stop, peers);
    } else
        srm_launch_pdl(st, dim3((k1 + 255) / 256), dim3(256), 0, k_update_pos, sites_in, (const double *)acc, nzbits, maskbits, g.n, ctl,
                       newpos, s.hash_next, respect_stop, peers);
    srm_launch_pdl(st, dim3((k1 + UPD_NT - 1) / UPD_NT), dim3(UPD_NT), 0, k_update_resolve, (const int *)newpos, s.hash_next, g.n, g.row0,
                   g.row1, s.bits_next, s.edge_next, ctl, sites_out, acc, Kcap, want_energy, stop_rule, respect_stop, peers);
}

void srm_preload_kernels() {
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, k_update_pos);
    cudaFuncGetAttributes(&a, k_update_claim);
    cudaFuncGetAttributes(&a, k_update_resolve);
    cudaFuncGetAttributes(&a, k_signal);
    cudaFuncGetAttributes(&a, k_acc);
    cudaGetLastError();
}

// ------------------------------------------------------------------ multires (coarse-to-fine, gcvt.cu:485-511)

// kernelDensityScaling: float adds in the reference's loop order (x outer, y inner), then an exact /4.
__global__ void __launch_bounds__(256) k_density_scale(const float *__restrict__ in, float *__restrict__ out, int s) {
    const int tx = blockIdx.x * 32 + (threadIdx.x & 31), ty = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (tx >= s || ty >= s) return;
    const size_t n2 = 2 * (size_t)s;
    const float2 a = *reinterpret_cast<const float2 *>(in + (size_t)(2 * ty) * n2 + 2 * tx);
    const float2 b = *reinterpret_cast<const float2 *>(in + (size_t)(2 * ty + 1) * n2 + 2 * tx);
    float d = __fadd_rn(0.0f, a.x);   // (2tx, 2ty)
    d = __fadd_rn(d, b.x);            // (2tx, 2ty+1)
    d = __fadd_rn(d, a.y);            // (2tx+1, 2ty)
    d = __fadd_rn(d, b.y);            // (2tx+1, 2ty+1)
    out[(size_t)ty * s + tx] = d * 0.25f;
}

void srm_launch_density_scale(cudaStream_t st, const float *in, float *out, int s) {
    SRM_COUNT(), k_density_scale<<<dim3((s + 31) / 32, (s + 7) / 8), 256, 0, st>>>(in, out, s);
}

// kernelZoomIn on the site list: (x, y) -> (2x, 2y); holes stay holes.
__global__ void k_zoom_sites(const int *__restrict__ in, int *__restrict__ out, int K) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K) return;
    const int p = in[i];
    out[i] = (p == SRM_SENT) ? SRM_SENT : srm_pack(srm_x(p) << 1, srm_y(p) << 1);
}

void srm_launch_zoom_sites(cudaStream_t st, const int *in, int *out, int K) {
    if (K > 0) SRM_COUNT(), k_zoom_sites<<<(K + 255) / 256, 256, 0, st>>>(in, out, K);
}

// constraint pixels found by the host scan (srm_host.cu) -> mask bitmap (cleared by the caller)
__global__ void k_scatter_mask(const int *__restrict__ pixels, int count, int n, uint32_t *__restrict__ maskbits) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int p = pixels[i];
    const size_t px = (size_t)srm_y(p) * n + srm_x(p);
    atomicOr(&maskbits[px >> 5], 1u << (px & 31));
}

void srm_launch_scatter_mask(cudaStream_t st, const int *pixels, int count, int n, uint32_t *maskbits) {
    if (count > 0) SRM_COUNT(), k_scatter_mask<<<(count + 255) / 256, 256, 0, st>>>(pixels, count, n, maskbits);
}

// 32 consecutive values -> one bitmap word ("value != 0"); one warp per 32 words: lane l reads float / byte l of every
// group (coalesced) and the ballot is the word.
template <typename T>
__global__ void __launch_bounds__(256) k_nonzero_bits(const T *__restrict__ v, size_t words, uint32_t *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * 256 + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * 256) >> 5;
    for (size_t w0 = warp * 32; w0 < words; w0 += nwarps * 32) {
        uint32_t mine = 0;
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
            const size_t w = w0 + k;
            const bool nz = w < words && v[w * 32 + lane] != (T)0;
            const uint32_t b = __ballot_sync(0xffffffffu, nz);
            if (lane == k) mine = b;
        }
        if (w0 + lane < words) out[w0 + lane] = mine;
    }
}

void srm_launch_nonzero_bits_f32(cudaStream_t st, const float *v, size_t count, uint32_t *out) {
    const size_t words = count / 32;
    const unsigned blocks = (unsigned)std::min<size_t>((words + 255) / 256, 148 * 8);
    if (words) SRM_COUNT(), k_nonzero_bits<float><<<blocks, 256, 0, st>>>(v, words, out);
}
void srm_launch_nonzero_bits_u8(cudaStream_t st, const unsigned char *v, size_t count, uint32_t *out) {
    const size_t words = count / 32;
    const unsigned blocks = (unsigned)std::min<size_t>((words + 255) / 256, 148 * 8);
    if (words) SRM_COUNT(), k_nonzero_bits<unsigned char><<<blocks, 256, 0, st>>>(v, words, out);
}

// ------------------------------------------------------------------ dense seed map -> site list

#define SFM_NT 256
#define SFM_TILE (SFM_NT * 4)
__device__ __forceinline__ int is_site4(int4 v, int *f) {
    f[0] = (short)(v.x & 0xffff) != SRM_MARK;
    f[1] = (short)(v.y & 0xffff) != SRM_MARK;
    f[2] = (short)(v.z & 0xffff) != SRM_MARK;
    f[3] = (short)(v.w & 0xffff) != SRM_MARK;
    return f[0] + f[1] + f[2] + f[3];
}

// Sites of a seed map = pixels whose x half is not MARKER (gcvt.cu:90, :240), in scan order.
__global__ void __launch_bounds__(SFM_NT) k_sites_count(const int4 *__restrict__ map4, size_t n4, int *blockcnt) {
    __shared__ int ws[SFM_NT / 32];
    size_t q = (size_t)blockIdx.x * SFM_NT + threadIdx.x;
    int f[4], c = 0;
    if (q < n4) c = is_site4(map4[q], f);
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int k = 0; k < SFM_NT / 32; ++k) s += ws[k];
        blockcnt[blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(SFM_NT) k_sites_write(const int4 *__restrict__ map4, size_t n4,
                                                        const int *__restrict__ blockoff, int *__restrict__ sites) {
    __shared__ int ws[SFM_NT / 32];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    size_t q = (size_t)blockIdx.x * SFM_NT + t;
    int f[4] = {0, 0, 0, 0}, c = 0;
    int4 v = make_int4(0, 0, 0, 0);
    if (q < n4) { v = map4[q]; c = is_site4(v, f); }
    int incl = warp_incl_scan(c, lane);
    if (lane == 31) ws[w] = incl;
    __syncthreads();
    int off = blockoff[blockIdx.x] + incl - c;
    for (int k = 0; k < w; ++k) off += ws[k];
    if (f[0]) sites[off++] = v.x;
    if (f[1]) sites[off++] = v.y;
    if (f[2]) sites[off++] = v.z;
    if (f[3]) sites[off++] = v.w;
}

// count_only != 0: fill blockcnt/blockoff and *total_out (the caller sizes the site arrays from it);
// count_only == 0: blockoff must hold the scan; writes the list.
void srm_launch_sites_from_map(cudaStream_t st, const int *site_map, size_t N, int *sites_out, int *blockcnt,
                               int *blockoff, int *total_out, int count_only) {
    const size_t n4 = N / 4;
    const int nb = (int)((n4 + SFM_NT - 1) / SFM_NT);
    if (count_only) {
        SRM_COUNT(), k_sites_count<<<nb, SFM_NT, 0, st>>>(reinterpret_cast<const int4 *>(site_map), n4, blockcnt);
        SRM_COUNT(), k_scan_counts<<<1, 1024, 0, st>>>(blockcnt, blockoff, nb, total_out);
    } else {
        SRM_COUNT(), k_sites_write<<<nb, SFM_NT, 0, st>>>(reinterpret_cast<const int4 *>(site_map), n4, blockoff, sites_out);
    }
}
