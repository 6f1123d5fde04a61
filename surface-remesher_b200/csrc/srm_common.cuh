// srm_common.cuh — shared definitions of the sm_100a discrete-CVT engine (libsrm.so).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <limits.h>

#define SRM_MARK (-32768)
#ifndef SRM_BAND_ORDER_DEFAULT
#define SRM_BAND_ORDER_DEFAULT 0   // default of the option "band_order" (srm_band.cu "Band order"; environment SRM_BAND_ORDER)
#endif
#define SRM_TIE_BAND_SHIFT 6   // 64-row bands of the reference's phase 1 (n/m1 == 64, gcvt.cu:842-847)
#define SRM_BIG 40000          // "no site in this column" distance; > any real |dy| (<= 32767), BIG^2+225 < 2^31
#define SRM_SENT ((int)0x80008000)  // (MARK,MARK) packed

// Device-side loop state (gcvt.cu:1105-1142 keeps these on the host; here the host never syncs).
struct SrmCtl {
    int K;          // length of the site list (fixed; merged sites become SRM_SENT holes, gcvt.cu:779-780)
    int nlive;      // live sites after the last update
    int live_acc;   // live-site counter of the running update
    int ticket;     // last-block-done counter of the running update
    int stop;       // reference stopping rule fired
    int it;         // iterations done (gcvtIterations)
    float omega;    // pbaOmega
    float lastE;    // lastEnergy
    float E;        // Energy (float like the reference)
    int ovf;        // rows handed to the robust path by the band kernel (this labelling)
    int band_ticket; // next band of the persistent band kernel (reset with ovf by k_bits)
    int p2p_timeout; // set if a peer never arrived (fail-safe of the spin wait)
    int epoch;       // bumped whenever the sites are (re)set: arrival flags carry epoch << 20 | (it + 1), never reset
    int pos_ticket;  // owner-computes update: blocks of k_update_pos done (the last one publishes the rank's newpos slice)
    int row_ticket;  // robust row kernel: CTAs done (the last one signals the peers in the fused all-reduce)
    int rle_used;    // entries of the run-length pool handed out by the current labelling (reset by the carry kernel)
    int rle_fail;    // a row found the pool exhausted (its offset is -1): the host grows the pool and labels again
    float escale;    // multires: energy factor 4^level (gcvt.cu:1082); 1 on the finest level
    double thresh;   // stopping threshold on the energy gradient: 1e-5 finest, 3e-1 coarse levels (gcvt.cu:1132-1137)
    int dbg[8];     // optional statistics of the band kernel: max/sum of band-list and row-survivor sizes
    unsigned long long prof[16];  // optional per-phase clock / element counters of the band kernel (dbg & 1)
};

// Layout of the fp64 prefix arrays (P2, PXX): SRM_PFX_TILE consecutive rows are interleaved, element (r, x) at
// srm_pfx_row(r, n) + x * SRM_PFX_TILE.  With 8 rows a 128-byte line of P2 holds one column of an 8-row band: the run
// ends of neighbouring rows lie at nearly the same x (cell boundaries are continuous), so the scattered end-of-run
// lookups of a band's 8 warps could share lines in L2 instead of pulling one DRAM sector each.  Measured on a B200
// (C3, 8192^2): k_band 212.7 us tiled against 214.8 us row-major, but k_prefix's strided stores cost more per call
// than that saves (e2e 2302 against 2338 it/s), so the default is 1 = plain row-major.
#ifndef SRM_PFX_TILE
#define SRM_PFX_TILE 1
#endif
__host__ __device__ __forceinline__ size_t srm_pfx_row(int r, int n) {
    return (size_t)(r / SRM_PFX_TILE) * (size_t)n * SRM_PFX_TILE + (size_t)(r % SRM_PFX_TILE);
}

__host__ __device__ __forceinline__ int srm_pack(int x, int y) { return (x & 0xffff) | (y << 16); }
__host__ __device__ __forceinline__ int srm_x(int p) { return (int)(short)(p & 0xffff); }
__host__ __device__ __forceinline__ int srm_y(int p) { return p >> 16; }

// Column candidate (semantics of kernelFloodDown/Up + kernelPropagateInterband +
// kernelUpdateVertical, gcvt.cu:77-216): U = nearest site row <= Y, D = nearest site row > Y.
__device__ __forceinline__ int srm_choose_col(int U, int D, int Y) {
    if (U == SRM_MARK) return D;
    if (D == SRM_MARK) return U;
    int du = Y - U, dd = D - Y;
    if (du < dd) return U;
    if (dd < du) return D;
    return ((D >> SRM_TIE_BAND_SHIFT) == (Y >> SRM_TIE_BAND_SHIFT)) ? D : U;
}

// ---- pixel -> site-id hash (replaces the two image-sized int maps of round 1: site ids and dedupe claims).
// Buckets of two {key, id} pairs (16 bytes, one load); key = packed pixel (x | y << 16), empty = 0xffffffff with id
// 0xffffffff (all-ones bytes: a plain memset clears a table; ids are compared as unsigned).  Open addressing over buckets, no deletions inside a step: an empty slot ends a probe sequence.  Two tables
// alternate by iteration parity: the update of iteration t claims the new pixels in table (t+1)&1 with atomicMin(id)
// (the smallest id keeps a contested pixel, the others become holes), which is then the id lookup of iteration t+1.
#define SRM_HEMPTY 0xffffffffu
struct SrmHash {
    uint4 *b = nullptr;    // buckets
    unsigned mask = 0;     // buckets - 1 (power of two)
    int shift = 32;        // 32 - log2(buckets)
};
__device__ __forceinline__ unsigned srm_hash_bucket(const SrmHash &H, unsigned key) { return (key * 2654435761u) >> H.shift; }

// id stored for `key` given its (already loaded) home bucket e; probes on only if both slots hold other keys.
__device__ __forceinline__ int srm_hash_find_from(const SrmHash &H, unsigned key, unsigned b, uint4 e) {
    for (unsigned probe = 0; probe <= H.mask; ++probe) {
        if (e.x == key) return (int)e.y;
        if (e.z == key) return (int)e.w;
        if (e.x == SRM_HEMPTY || e.z == SRM_HEMPTY) return -1;
        b = (b + 1) & H.mask;
        e = H.b[b];
    }
    return -1;
}
__device__ __forceinline__ int srm_hash_find(const SrmHash &H, unsigned key) {
    const unsigned b = srm_hash_bucket(H, key);
    return srm_hash_find_from(H, key, b, H.b[b]);
}
// insert-or-min: after all claims of a step, the entry of `key` holds the smallest claiming id
__device__ __forceinline__ void srm_hash_claim(const SrmHash &H, unsigned key, int id) {
    unsigned b = srm_hash_bucket(H, key);
    for (unsigned probe = 0; probe <= H.mask; ++probe) {
        unsigned *w = reinterpret_cast<unsigned *>(H.b + b);
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const unsigned old = atomicCAS(w + 2 * s, SRM_HEMPTY, key);
            if (old == SRM_HEMPTY || old == key) { atomicMin(w + 2 * s + 1, (unsigned)id); return; }
        }
        b = (b + 1) & H.mask;
    }
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += y;
    }
    return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Peer view of the row-band accumulators (fused all-reduce over NVLink peer memory, srm_lloyd.cu).
struct SrmPeers {
    const double *const *acc = nullptr;  // device array [world]: base of every rank's accumulator pair
    int *const *flags = nullptr;         // device array [world]: every rank's arrival-flag array
    int *flags_local = nullptr;          // this rank's flag array: slot q = last iteration rank q finished accumulating,
                                         // slot 64 + q = last iteration whose newpos slice rank q has delivered
    int world = 1, rank = 0;
    size_t stride = 0;                   // doubles per accumulator buffer (two buffers, used by iteration parity)
    int kcap = 0;                        // sites per buffer: 4*kcap+4 doubles of sums, then kcap "touched" bytes
    int parity = 0;
    int owner = 0;                       // 1: owner-computes update (every rank updates 1/world of the ids and delivers the
                                         // result to all peers); 0: every rank pulls the sums of all sites (small site sets)
};

// Kernel launches issued by this library since it was loaded (srm_launch_count(); bench.py reports the difference
// over its timed region as gpu_launches).  Every `<<<>>>` site counts itself: SRM_COUNT(), k_x<<<...>>>(...).
extern long long g_srm_launches;
#define SRM_COUNT() ((void)__sync_fetch_and_add(&g_srm_launches, 1ll))
#define SRM_COUNT_N(k) ((void)__sync_fetch_and_add(&g_srm_launches, (long long)(k)))

// ---- Programmatic dependent launch (PDL): the kernels of the Lloyd loop are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so that a kernel's CTAs are scheduled while its predecessor
// drains instead of after the usual kernel-boundary gap.  Contract: every such kernel executes srm_pdl_enter() as its
// FIRST statement on every path (it must not return, nor touch global memory, before it): griddepcontrol.wait blocks
// until the preceding grid has completed and its writes are visible, so the chain stays transitively ordered;
// griddepcontrol.launch_dependents then lets the successor be scheduled as soon as all CTAs of this grid are resident.
// Without the launch attribute both instructions are no-ops.  SRM_PDL=0 in the environment disables the attribute.
__device__ __forceinline__ void srm_pdl_enter() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
bool srm_pdl_enabled();
template <typename... KArgs, typename... Args>
static inline cudaError_t srm_launch_pdl(cudaStream_t st, dim3 grid, dim3 block, size_t smem, void (*kern)(KArgs...), Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = srm_pdl_enabled() ? 1 : 0;
    SRM_COUNT();
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// Run-length labels of a labelling: row r holds cnt[r] runs {packed site, first X} at pool[off[r]..].  Rows allocate
// exactly what they need from one pool (atomicAdd on SrmCtl::rle_used); the pool has rows * min(n, buffer capacity)
// entries instead of rows * n (537 MB at 8192^2 in round 1), and a labelling that exhausts it (adversarial inputs only:
// more runs per row than the band kernel's buffer on average) is repeated by the host with a full-size pool.
struct SrmRle {
    int2 *pool = nullptr;
    int *off = nullptr, *cnt = nullptr;
    int cap = 0;
    int2 *scratch = nullptr;   // robust row kernel: n entries per CTA (it accumulates from here, pool or not)
    int scratch_ctas = 0;
};
__device__ __forceinline__ int srm_rle_alloc(const SrmRle &R, SrmCtl *ctl, int r, int count) {   // one thread per row
    const int padded = (count + 1) & ~1;   // rows start 16-byte aligned (bulk copies of a row's list, k_expand)
    const int o = atomicAdd(&ctl->rle_used, padded);
    const bool ok = o + padded <= R.cap;
    R.off[r] = ok ? o : -1;
    R.cnt[r] = count;
    if (!ok) atomicExch(&ctl->rle_fail, 1);
    return ok ? o : -1;
}

// ---- launchers (host), one per pipeline stage; all asynchronous on `st`.
struct SrmGrid {           // geometry of one context
    int n, row0, row1;
    int nrows() const { return row1 - row0; }
};

// Per-iteration device state that alternates by iteration parity (cur = it & 1 is read by the labelling, next is
// cleared by the carry kernel and filled by the update): site bitmap, band edges, pixel -> id hash.
struct SrmStep {
    const uint32_t *bits;   // site bitmap of this iteration, indexed [(y >> 5) * n + x] (band contexts: own word rows)
    const int *edge;        // row bands: per column nearest site row above / below the band (2n ints), else nullptr
    SrmHash hash;           // pixel -> id of this iteration's sites
    uint32_t *bits_next;    // next iteration's buffers
    int *edge_next;
    SrmHash hash_next;
};

// Build bitmap / edges / hash of buffer set `cur` from a site list (after srm_set_sites / srm_set_site_map).
void srm_launch_init_sites(cudaStream_t st, const int *sites, SrmCtl *ctl, int Kcap, int n, uint32_t *bits, size_t bits_words,
                           int *edge, SrmHash hash, int row0, int row1);
// Carries of the current bitmap; also clears the next iteration's bitmap / edges / hash and the robust-path row list.
void srm_launch_carry(cudaStream_t st, const SrmStep &s, int n, short *up, short *dn, SrmCtl *ctl, int respect_stop,
                      int row0, int row1);
// fused fast path (srm_band.cu)
cudaError_t srm_band_setup(int n);
int srm_band_bufcap(int n);   // entries of a warp's element buffer = most runs a band-kernel row can have
// flags of the band kernel
enum {
    SRM_BF_ACC = 1,      // accumulate the per-site sums (fused centroid pass)
    SRM_BF_ENERGY = 2,   // ... and the CVT energy
    SRM_BF_STOP = 4,     // return at once if the device-side stop flag is set
    SRM_BF_RLE = 8,      // write the run-length rows (final labelling, stepwise API; the loop does not need them)
    SRM_BF_TOUCH = 16    // mark the sites this rank contributed to (read by the peer-memory all-reduce)
};
cudaError_t srm_launch_band(cudaStream_t st, const uint32_t *bits, const short *up, const short *dn, SrmGrid g, SrmRle rle,
                            int *ovf_rows, const double2 *P2, const double *PXX, SrmHash hash,
                            double *acc, int Kcap, SrmCtl *ctl, int flags, int dbg = 0, const int *perm = nullptr,
                            int refresh_order = 0);   // perm: CTA -> band order (srm_band.cu "Band order"), rebuilt first if asked
// robust path, driven by a row list (rows == nullptr: every row of the band)
cudaError_t srm_launch_row(cudaStream_t st, const uint32_t *bits, const short *up, const short *dn, SrmGrid g, SrmRle rle,
                           const int *rows, const int *count, const double2 *P2, const double *PXX,
                           SrmHash hash, double *acc, int Kcap, SrmCtl *ctl, int accumulate, int want_energy,
                           int respect_stop, int write_rle, SrmPeers signal = SrmPeers());
int srm_row_scratch_ctas(int nrows);   // CTAs of the robust row kernel (= n-entry slices of SrmRle::scratch)
cudaError_t srm_launch_expand(cudaStream_t st, SrmRle rle, SrmGrid g, int *labels);
cudaError_t srm_label_setup(int n);  // opt-in shared memory sizes

void srm_launch_prefix(cudaStream_t st, const float *density_band, SrmGrid g, double2 *P2, double *PXX);
void srm_launch_acc(cudaStream_t st, SrmRle rle, const double2 *P2, const double *PXX,
                    SrmHash hash, SrmGrid g, double *acc, int Kcap, const int *rows, const int *count,
                    const SrmCtl *ctl, int want_energy, int respect_stop);
// Site update (two kernels): new positions + claims in s.hash_next, then winners -> sites_out / s.bits_next / s.edge_next.
void srm_launch_update(cudaStream_t st, const int *sites_in, int *sites_out, double *acc, const uint32_t *nzbits,
                       const uint32_t *maskbits, SrmGrid g, SrmCtl *ctl, int Kcap, int *newpos, const SrmStep &s,
                       int want_energy, int stop_rule, int respect_stop, SrmPeers peers = SrmPeers());
// one bit per value: "value != 0"; count must be a multiple of 32
void srm_launch_nonzero_bits_f32(cudaStream_t st, const float *v, size_t count, uint32_t *out);
void srm_launch_nonzero_bits_u8(cudaStream_t st, const unsigned char *v, size_t count, uint32_t *out);
void srm_launch_signal(cudaStream_t st, SrmCtl *ctl, SrmPeers peers, int respect_stop);
// stand-alone centroid pass over a dense label map (srm_centroid.cu): labels / density = the context's rows
cudaError_t srm_launch_centroid_dense(cudaStream_t st, const int *labels, const float *density, SrmHash hash, SrmGrid g,
                                      double *acc, int Kcap, int want_energy, int touch);
void srm_preload_kernels();   // resolves every kernel of the loop up front (lazy module loading may otherwise synchronise
                              // the device in the middle of a launch sequence, while a kernel spins on a peer's flag)
void srm_launch_sites_from_map(cudaStream_t st, const int *site_map, size_t N, int *sites_out, int *blockcnt,
                               int *blockoff, int *total_out, int count_only);
void srm_launch_scan_counts(cudaStream_t st, const int *cnt, int *off, int nb, int *total_out);
void srm_launch_jfa_pass(cudaStream_t st, const int *in, int *out, int n, int step);
// whole JFA schedule between two dense maps (srm_jfa.cu); returns the buffer holding the result
int *srm_launch_jfa(cudaStream_t st, int *a, int *b, int n, const int *steps, int nsteps, int fused, cudaEvent_t *ev,
                    int evcap, int *nlaunch, cudaError_t *err);
void srm_launch_scatter_sites(cudaStream_t st, const int *sites, const SrmCtl *ctl, int Kcap, int n, int *map);
void srm_launch_fill_int(cudaStream_t st, int *p, size_t count, int value);
// multires (gcvt.cu:485-511): 2x2 box filter of the density (s = output side), site zoom x2
void srm_launch_density_scale(cudaStream_t st, const float *in, float *out, int s);
void srm_launch_zoom_sites(cudaStream_t st, const int *in, int *out, int K);

// point location + lift (srm_recover.cu; recover.h:63-153)
struct SrmLocator;
cudaError_t srm_locator_build(cudaStream_t st, const double *pts_host, int P, const double *pts_dev, const int *tri_dev,
                              int T, SrmLocator **out);
void srm_locator_free(SrmLocator *L);
cudaError_t srm_locator_query(cudaStream_t st, const SrmLocator *L, const double *pts_dev, const int *tri_dev,
                              const double *qxy_dev, const int *centroid_of_dev, int Q, int *face_dev, double *w_dev);
cudaError_t srm_launch_lift(cudaStream_t st, const int *tri_dev, const double *pts3d_dev, const int *face_dev,
                            const double *w_dev, int Q, double *out_dev);

cudaError_t srm_raster(cudaStream_t st, const double *pts, const double *wt, int num_point, const int *tri, int num_tri,
                       float *density, double scale, int n);

// host side of the boundary (srm_host.cu): pageable <-> device copies at PCIe speed, host scans of the sparse inputs
#ifdef __cplusplus
#include <vector>
cudaError_t srm_h2d_pageable(void *dst_dev, const void *src_host, size_t bytes, cudaStream_t after);
cudaError_t srm_d2h_pageable(void *dst_host, const void *src_dev, size_t bytes, cudaStream_t after);
void srm_scan_site_map(const int *site_map, size_t N, std::vector<int> &sites);
void srm_scan_mask(const unsigned char *mask, int n, std::vector<int> &pixels, int row0 = 0, int row1 = -1);
void srm_host_pool_release();
int srm_host_set_config(int threads, int chunk_kb);
void srm_launch_scatter_mask(cudaStream_t st, const int *pixels, int count, int n, uint32_t *maskbits);
#endif
