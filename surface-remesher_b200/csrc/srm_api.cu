// srm_api.cu — context management and the C ABI of libsrm.so (include/srm.h).
//
// Host side of the hot path: what gcvtInitialization / pba2DInitializeInput / gCVT /
// pbaCVDDeinitialization (gcvt.cu:840-914, 1087-1156) and discretization_d
// (discretization.cu:87-120) do around the kernels, re-designed so that all loop state stays on
// the device (no host sync per iteration; the reference blocks on a 4-byte D2H every 10th
// iteration, gcvt.cu:1080) and a context can own a row band of the grid for multi-GPU sharding.
#include "../../include/srm.h"
#include "srm_common.cuh"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <algorithm>
#include <dlfcn.h>
#include <mutex>
#include <thread>
#include <vector>

static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return fail(SRM_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                        __LINE__);                                                                       \
    } while (0)

long long g_srm_launches = 0;
static thread_local bool g_pdl_suppress = false;   // set while a CUDA graph is being captured (plain edges there)
bool srm_pdl_enabled() {
#ifdef SRM_NO_PDL
    return false;
#endif
    if (g_pdl_suppress) return false;
    static const bool on = []() { const char *e = getenv("SRM_PDL"); return !(e && e[0] == '0'); }();
    return on;
}
extern "C" long long srm_launch_count(void) { return __sync_fetch_and_add(&g_srm_launches, 0ll); }
extern "C" const char *srm_last_error(void) { return g_err; }
extern "C" int srm_version(void) { return 100; }

// ---- NCCL, bound at run time (dlopen) so that libsrm.so has no link-time dependency on it.  Only the row-band
// all-reduce uses it; single-GPU use never loads it.
typedef struct { char internal[128]; } SrmNcclId;
struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(SrmNcclId *) = nullptr;
    int (*CommInitRank)(void **, int, SrmNcclId, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static std::mutex g_nccl_mu;

static int load_nccl() {
    std::lock_guard<std::mutex> lock(g_nccl_mu);
    if (g_nccl.handle) return SRM_OK;
    const char *names[] = {getenv("SRM_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *nm : names)
        if (nm && (h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!h) return fail(SRM_ERR_CUDA, "NCCL library not found (libnccl.so.2; set SRM_NCCL_LIB): %s", dlerror());
    g_nccl.GetUniqueId = (int (*)(SrmNcclId *))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void **, int, SrmNcclId, int))dlsym(h, "ncclCommInitRank");
    g_nccl.AllReduce = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t))dlsym(h, "ncclAllReduce");
    g_nccl.CommDestroy = (int (*)(void *))dlsym(h, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy)
        return fail(SRM_ERR_CUDA, "NCCL library lacks a required symbol");
    g_nccl.handle = h;
    return SRM_OK;
}

// Default of the option "band_order" (SRM_BAND_ORDER=0/1 overrides it; read once per process).
static int default_band_order() {
    static const int v = []() {
        const char *e = getenv("SRM_BAND_ORDER");
        return (e && (e[0] == '0' || e[0] == '1')) ? e[0] - '0' : SRM_BAND_ORDER_DEFAULT;
    }();
    return v;
}

struct srm_ctx {
    SrmGrid g{};
    int device = 0;
    size_t N = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // inputs: the density of the context's own rows (the prefix sums are built from it), and two full-grid bitmaps of
    // N/8 bytes each that the replicated site update reads at arbitrary pixels: "density != 0" and "constraint pixel"
    float *density = nullptr;
    uint32_t *nzbits = nullptr, *maskbits = nullptr;
    bool has_mask = false, has_density = false, has_sites = false, labelled = false;
    double2 *P2 = nullptr;
    double *PXX = nullptr;
    // sites
    int *sites[2] = {nullptr, nullptr};
    int cur = 0, Kcap = 0;
    double *acc = nullptr;
    int *newpos = nullptr, *blockcnt = nullptr, *blockoff = nullptr;
    size_t blockcap = 0;
    // labelling state.  bits / up / dn hold the context's own word rows only; the base pointers are shifted so that the
    // kernels index them with absolute rows: [(y >> 5) * n + x].  bits, edge and hash alternate by iteration parity.
    uint32_t *bits_alloc[2] = {nullptr, nullptr}, *bits[2] = {nullptr, nullptr};
    short *up_alloc = nullptr, *dn_alloc = nullptr, *up = nullptr, *dn = nullptr;
    size_t band_words = 0;
    int2 *rle = nullptr, *row_scratch = nullptr;   // run-length pool (SrmRle), scratch rows of the robust row kernel
    int rle_cap = 0;
    int *rle_cnt = nullptr, *rle_off = nullptr, *labels = nullptr, *scratch_map = nullptr;
    int *ovf_rows = nullptr;   // rows the band kernel hands to the robust path
    // option "band_order": the band kernel's CTAs take the bands by decreasing cost (runs per band of an earlier
    // iteration, srm_band.cu "Band order") instead of in row order; the permutation is rebuilt every 10th iteration
    int *band_perm = nullptr;
    int band_order = default_band_order();
    int *edge[2] = {nullptr, nullptr};   // row bands: per column, nearest site row above / below the band (2n ints)
    SrmHash hash[2];           // pixel -> site id (and the dedupe claims of the update)
    int dbg_stats = 0;
    bool robust_only = false;  // option: label every row with the robust path (tests pin it this way)
    int jfa_mode = 1;          // option: 1 = srm_jfa.cu (fused small-step tile kernel + vectorised far passes), 0 = one round-1
                               // kernel per pass (A/B baseline), 3 = like 1 with the large-grid (n > 16384) instantiations
    // option "graph": srm_iterate replays a captured CUDA graph of 10 iterations (the period of the loop: buffer
    // parity 2, energy every 10th) instead of launching 50 kernels from the host — for launch-bound sizes / batches
    bool use_graph = false;
    cudaGraphExec_t graph[2] = {nullptr, nullptr};   // [stop_rule]
    int graph_key = 0;         // configuration the graphs were captured for (invalidated when it changes)
    int graph_kernels[2] = {0, 0};   // kernel nodes per graph (for srm_launch_count)
    SrmCtl *ctl = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int it_host = 0;   // iterations executed since the sites were set (host mirror of SrmCtl::it)
    bool stopped = false;
    bool unsynced = false;   // srm_iterate with the stop rule returned without reading the device's iteration count back
    void *comm = nullptr;  // NCCL communicator of the row bands (world > 1)
    int world = 1;
    // fused all-reduce over peer memory
    bool p2p = false;
    int rank = 0, epoch = 0;
    size_t acc_stride = 0;            // doubles per accumulator buffer (two buffers)
    int *flags = nullptr;             // arrival flags, one slot per rank
    const double **d_peer_acc = nullptr;
    int **d_peer_flags = nullptr;
    std::vector<void *> ipc_opened;
};

// CTA -> band permutation of the band kernel, or nullptr for row order
static const int *perm_of(const srm_ctx *c) { return c->band_order ? c->band_perm : nullptr; }

static void drop_graphs(srm_ctx *c) {
    for (int i = 0; i < 2; ++i) if (c->graph[i]) { cudaGraphExecDestroy(c->graph[i]); c->graph[i] = nullptr; }
}

static int valid_n(int n) { return n >= 256 && n <= 32768 && (n % 256) == 0; }

// Site-indexed buffers.  They are kept when a later site set fits (Kcap is then the capacity, ctl->K the list
// length), so that peer mappings of the accumulators (srm_p2p_connect) survive repeated calls.
static int alloc_sites(srm_ctx *c, int K) {
    if (c->sites[0] && K <= c->Kcap) {
        CK(cudaMemsetAsync(c->acc, 0, 2 * c->acc_stride * sizeof(double), c->stream));
        c->cur = 0;
        return SRM_OK;
    }
    if (c->world > 1)   // peer mappings (srm_p2p_connect) and the NCCL element count refer to the current buffers
        return fail(SRM_ERR_STATE, "site set of %d entries exceeds the capacity %d the row-band collective was set up "
                                   "with: destroy the band contexts and reconnect", K, c->Kcap);
    drop_graphs(c);   // captured kernel arguments point into the buffers freed below
    for (int i = 0; i < 2; ++i) if (c->sites[i]) { cudaFree(c->sites[i]); c->sites[i] = nullptr; }
    for (int i = 0; i < 2; ++i) if (c->hash[i].b) { cudaFree(c->hash[i].b); c->hash[i].b = nullptr; }
    if (c->acc) { cudaFree(c->acc); c->acc = nullptr; }
    c->newpos = nullptr;   // lives behind the accumulator pair (one allocation, one IPC mapping for the peers)
    c->Kcap = K;
    size_t k1 = (size_t)(K > 0 ? K : 1);
    {   // hash tables: two-slot buckets, buckets = power of two >= 2 K (slot load <= 0.25, 16 bytes per bucket)
        int lg = 10;
        while ((1ull << lg) < 2 * k1 && lg < 30) ++lg;
        for (int i = 0; i < 2; ++i) {
            CK(cudaMalloc(&c->hash[i].b, sizeof(uint4) << lg));
            c->hash[i].mask = (1u << lg) - 1u;
            c->hash[i].shift = 32 - lg;
        }
    }
    CK(cudaMalloc(&c->sites[0], k1 * sizeof(int)));
    CK(cudaMalloc(&c->sites[1], k1 * sizeof(int)));
    // two buffers (the peer-memory all-reduce alternates them by iteration parity), each: 4K+4 doubles of sums
    // followed by K "this rank contributed to the site" bytes
    c->acc_stride = (4 * (size_t)K + 4 + ((size_t)K + 7) / 8 + 3) & ~(size_t)3;  // 32-byte aligned buffers
    CK(cudaMalloc(&c->acc, 2 * c->acc_stride * sizeof(double) + k1 * sizeof(int)));
    c->newpos = reinterpret_cast<int *>(c->acc + 2 * c->acc_stride);   // peers store their slice of the new positions here
    CK(cudaMemsetAsync(c->acc, 0, 2 * c->acc_stride * sizeof(double), c->stream));
    c->p2p = false;  // peer mappings refer to the old buffers
    c->cur = 0;
    return SRM_OK;
}

static int reset_ctl(srm_ctx *c, int K) {
    SrmCtl h;
    memset(&h, 0, sizeof(h));
    h.K = K; h.nlive = K; h.omega = 2.0f; h.lastE = 1e18f; h.E = 0.0f;  // gcvt.cu:1105-1108
    h.escale = 1.0f; h.thresh = 1e-5;                                     // finest level (gcvt.cu:1082,1136)
    h.epoch = ++c->epoch;
    CK(cudaMemcpyAsync(c->ctl, &h, sizeof(h), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));  // h is a stack object
    c->it_host = 0;
    c->stopped = false;
    c->unsynced = false;
    c->labelled = false;
    return SRM_OK;
}

static bool is_band(const srm_ctx *c) { return c->g.row0 > 0 || c->g.row1 < c->g.n; }

// Buffer sets of iteration `it`: cur = it & 1 is read by the labelling, next is cleared by the carry kernel and
// filled by the update.
static SrmStep step_of(srm_ctx *c, int it) {
    const int cur = it & 1, nxt = cur ^ 1;
    SrmStep s;
    s.bits = c->bits[cur]; s.edge = is_band(c) ? c->edge[cur] : nullptr; s.hash = c->hash[cur];
    s.bits_next = c->bits[nxt]; s.edge_next = is_band(c) ? c->edge[nxt] : nullptr; s.hash_next = c->hash[nxt];
    return s;
}

// Bitmap, band edges and hash of the buffer set `parity` from the site list sites[parity] (ctl->K entries).
static int init_step_buffers(srm_ctx *c, int parity) {
    srm_launch_init_sites(c->stream, c->sites[parity], c->ctl, c->Kcap, c->g.n, c->bits[parity], c->band_words,
                          is_band(c) ? c->edge[parity] : nullptr, c->hash[parity], c->g.row0, c->g.row1);
    CK(cudaGetLastError());
    return SRM_OK;
}

extern "C" int srm_create(srm_ctx **out, int n, int row0, int row1, int device) {
    if (!out) return fail(SRM_ERR_ARG, "srm_create: null out");
    *out = nullptr;
    if (!valid_n(n)) return fail(SRM_ERR_ARG, "srm_create: n=%d unsupported (multiple of 256 in [256,32768])", n);
    if (row0 < 0 || row1 > n || row0 >= row1 || (row0 % 64) || (row1 % 64))
        return fail(SRM_ERR_ARG, "srm_create: rows [%d,%d) must be multiples of 64 within [0,%d]", row0, row1, n);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(SRM_ERR_CUDA, "srm_create: no CUDA device (libsrm has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(SRM_ERR_ARG, "srm_create: device %d of %d", device, ndev);
    CK(cudaSetDevice(device));
    srm_ctx *c = new srm_ctx();
    c->g.n = n; c->g.row0 = row0; c->g.row1 = row1;
    c->device = device;
    c->N = (size_t)n * n;
    const size_t NB = (size_t)c->g.nrows() * n, NW = (size_t)(n >> 5) * n;
#define CKD(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess) {                                                                        \
            srm_destroy(c);                                                                              \
            return fail(SRM_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                        __LINE__);                                                                       \
        }                                                                                                \
    } while (0)
    CKD(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
    CKD(cudaEventCreate(&c->ev0));
    CKD(cudaEventCreate(&c->ev1));
    CKD(cudaMalloc(&c->density, NB * sizeof(float)));
    CKD(cudaMalloc(&c->nzbits, c->N / 8));
    CKD(cudaMalloc(&c->maskbits, c->N / 8));
    CKD(cudaMalloc(&c->P2, NB * sizeof(double2)));
    CKD(cudaMalloc(&c->PXX, NB * sizeof(double)));
    (void)NW;
    c->band_words = (size_t)(c->g.nrows() >> 5) * n;
    const size_t woff = (size_t)(row0 >> 5) * n;   // kernels index with absolute word rows
    for (int i = 0; i < 2; ++i) {
        CKD(cudaMalloc(&c->bits_alloc[i], c->band_words * sizeof(uint32_t)));
        c->bits[i] = c->bits_alloc[i] - woff;
        CKD(cudaMalloc(&c->edge[i], 2 * (size_t)n * sizeof(int)));
    }
    CKD(cudaMalloc(&c->up_alloc, c->band_words * sizeof(short)));
    CKD(cudaMalloc(&c->dn_alloc, c->band_words * sizeof(short)));
    c->up = c->up_alloc - woff; c->dn = c->dn_alloc - woff;
    c->rle_cap = c->g.nrows() * std::min(n, srm_band_bufcap(n));   // enough for every row at the band kernel's capacity
    CKD(cudaMalloc(&c->rle, (size_t)c->rle_cap * sizeof(int2)));
    CKD(cudaMalloc(&c->row_scratch, (size_t)srm_row_scratch_ctas(c->g.nrows()) * n * sizeof(int2)));
    CKD(cudaMalloc(&c->rle_cnt, (size_t)c->g.nrows() * sizeof(int)));
    CKD(cudaMemsetAsync(c->rle_cnt, 0, (size_t)c->g.nrows() * sizeof(int), c->stream));
    {   // band order: the identity until the first rebuild
        std::vector<int> ident((size_t)c->g.nrows() / 8);
        for (size_t i = 0; i < ident.size(); ++i) ident[i] = (int)i;
        CKD(cudaMalloc(&c->band_perm, ident.size() * sizeof(int)));
        CKD(cudaMemcpy(c->band_perm, ident.data(), ident.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    CKD(cudaMalloc(&c->rle_off, (size_t)c->g.nrows() * sizeof(int)));
    CKD(cudaMalloc(&c->ovf_rows, (size_t)c->g.nrows() * sizeof(int)));
    CKD(cudaMalloc(&c->ctl, sizeof(SrmCtl)));
    CKD(cudaMalloc(&c->flags, 128 * sizeof(int)));   // [0,64): sums complete, [64,128): newpos slice delivered, per rank
    CKD(cudaMemsetAsync(c->flags, 0, 128 * sizeof(int), c->stream));
    c->blockcap = c->N / 256 + 2;  // covers both the seed-map compaction (N/1024 tiles) and K <= N sites
    CKD(cudaMalloc(&c->blockcnt, c->blockcap * sizeof(int)));
    CKD(cudaMalloc(&c->blockoff, c->blockcap * sizeof(int)));
    CKD(cudaMemsetAsync(c->ctl, 0, sizeof(SrmCtl), c->stream));
    CKD(srm_label_setup(n));
    srm_preload_kernels();
    CKD(cudaStreamSynchronize(c->stream));
#undef CKD
    *out = c;
    return SRM_OK;
}

extern "C" int srm_destroy(srm_ctx *c) {
    if (!c) return SRM_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    drop_graphs(c);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    for (void *p : c->ipc_opened) cudaIpcCloseMemHandle(p);
    if (c->flags) cudaFree(c->flags);
    if (c->d_peer_acc) cudaFree((void *)c->d_peer_acc);
    if (c->d_peer_flags) cudaFree((void *)c->d_peer_flags);
    void *ptrs[] = {c->density, c->nzbits, c->maskbits, c->P2, c->PXX, c->sites[0], c->sites[1], c->acc, c->blockcnt,
                    c->blockoff, c->bits_alloc[0], c->bits_alloc[1], c->up_alloc, c->dn_alloc, c->rle, c->row_scratch, c->rle_cnt, c->rle_off, c->ovf_rows, c->band_perm,
                    c->edge[0], c->edge[1], c->hash[0].b, c->hash[1].b, c->labels, c->scratch_map, c->ctl};
    for (void *p : ptrs) if (p) cudaFree(p);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return SRM_OK;
}

// Row-band collective: one all-reduce (sum, fp64) of the per-site accumulators per iteration.
// rank 0 creates the id, the caller distributes the 128 bytes (torch.distributed / MPI / files: plumbing),
// every rank calls srm_nccl_init on its band context.
extern "C" int srm_nccl_unique_id(char *id128) {
    if (!id128) return fail(SRM_ERR_ARG, "srm_nccl_unique_id: null argument");
    int rc = load_nccl();
    if (rc) return rc;
    SrmNcclId id;
    int e = g_nccl.GetUniqueId(&id);
    if (e) return fail(SRM_ERR_CUDA, "ncclGetUniqueId: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(e) : "error");
    memcpy(id128, id.internal, 128);
    return SRM_OK;
}

extern "C" int srm_nccl_init(srm_ctx *c, const char *id128, int rank, int world) {
    if (!c || !id128 || world < 1 || rank < 0 || rank >= world) return fail(SRM_ERR_ARG, "srm_nccl_init: bad argument");
    int rc = load_nccl();
    if (rc) return rc;
    CK(cudaSetDevice(c->device));
    SrmNcclId id;
    memcpy(id.internal, id128, 128);
    int e = g_nccl.CommInitRank(&c->comm, world, id, rank);
    if (e) { c->comm = nullptr; return fail(SRM_ERR_CUDA, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(e) : "error"); }
    c->world = world;
    return SRM_OK;
}

static SrmPeers peers_of(srm_ctx *c, int it) {
    SrmPeers p;
    if (c->p2p) {
        p.acc = c->d_peer_acc; p.flags = c->d_peer_flags; p.flags_local = c->flags;
        p.world = c->world; p.rank = c->rank; p.stride = c->acc_stride; p.parity = it & 1; p.kcap = c->Kcap;
        // owner-computes pays one more kernel and flag exchange per iteration and saves (world-1)/world of the remote
        // loads: measured break-even far above 10^5 sites (8192^2 / 100k: 33 -> 46 us per update on 2 GPUs; 32768^2 /
        // 10^6: 420 -> 125 us)
        p.owner = c->Kcap >= 300000;
    }
    return p;
}
static double *cur_acc(srm_ctx *c, int it) { return c->acc + (c->p2p ? (size_t)(it & 1) * c->acc_stride : 0); }

// Peer-memory all-reduce set-up.  Every rank fills a blob describing its accumulator pair and flag array (CUDA IPC
// handles + raw pointers + pid), the caller gathers the blobs of all ranks (plumbing), every rank connects.
// Call after the sites are set (that allocates the accumulators) and before iterating; collective.
struct SrmP2PBlob {
    cudaIpcMemHandle_t acc, flags;
    unsigned long long acc_ptr, flags_ptr;
    long long pid;
    int device, pad;
};
static_assert(sizeof(SrmP2PBlob) == 160, "blob layout is part of the C ABI");

extern "C" int srm_p2p_info(srm_ctx *c, void *blob160) {
    if (!c || !blob160) return fail(SRM_ERR_ARG, "srm_p2p_info: null argument");
    if (!c->has_sites) return fail(SRM_ERR_STATE, "srm_p2p_info: set the sites first");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    SrmP2PBlob b;
    memset(&b, 0, sizeof(b));
    CK(cudaIpcGetMemHandle(&b.acc, c->acc));
    CK(cudaIpcGetMemHandle(&b.flags, c->flags));
    b.acc_ptr = (unsigned long long)c->acc; b.flags_ptr = (unsigned long long)c->flags;
    b.pid = (long long)getpid(); b.device = c->device;
    memcpy(blob160, &b, sizeof(b));
    return SRM_OK;
}

extern "C" int srm_p2p_connect(srm_ctx *c, const void *blobs, int rank, int world) {
    if (!c || !blobs || world < 1 || world > 64 || rank < 0 || rank >= world)
        return fail(SRM_ERR_ARG, "srm_p2p_connect: bad argument");
    CK(cudaSetDevice(c->device));
    const SrmP2PBlob *b = (const SrmP2PBlob *)blobs;
    std::vector<const double *> pa((size_t)world);
    std::vector<int *> pf((size_t)world);
    for (int q = 0; q < world; ++q) {
        if (q == rank) { pa[q] = c->acc; pf[q] = c->flags; continue; }
        if (b[q].pid == (long long)getpid()) {  // same process (tests): raw pointers, peer access if another device
            if (b[q].device != c->device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(b[q].device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    return fail(SRM_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", b[q].device, cudaGetErrorString(e));
                cudaGetLastError();
            }
            pa[q] = (const double *)b[q].acc_ptr; pf[q] = (int *)b[q].flags_ptr;
        } else {
            void *p1 = nullptr, *p2 = nullptr;
            CK(cudaIpcOpenMemHandle(&p1, b[q].acc, cudaIpcMemLazyEnablePeerAccess));
            c->ipc_opened.push_back(p1);
            CK(cudaIpcOpenMemHandle(&p2, b[q].flags, cudaIpcMemLazyEnablePeerAccess));
            c->ipc_opened.push_back(p2);
            pa[q] = (const double *)p1; pf[q] = (int *)p2;
        }
    }
    if (!c->d_peer_acc) CK(cudaMalloc((void **)&c->d_peer_acc, 64 * sizeof(void *)));
    if (!c->d_peer_flags) CK(cudaMalloc((void **)&c->d_peer_flags, 64 * sizeof(void *)));
    CK(cudaMemcpy((void *)c->d_peer_acc, pa.data(), world * sizeof(void *), cudaMemcpyHostToDevice));
    CK(cudaMemcpy((void *)c->d_peer_flags, pf.data(), world * sizeof(void *), cudaMemcpyHostToDevice));
    c->world = world; c->rank = rank; c->p2p = world > 1;
    srm_preload_kernels();
    return SRM_OK;
}

// Undo srm_p2p_connect: closes the mappings of the peers' buffers.  Collective in the sense that every rank must have
// disconnected (caller: barrier) before any rank destroys its context — a peer's mapping must not outlive the buffer.
extern "C" int srm_p2p_disconnect(srm_ctx *c) {
    if (!c) return fail(SRM_ERR_ARG, "srm_p2p_disconnect: null ctx");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    for (void *p : c->ipc_opened) cudaIpcCloseMemHandle(p);
    c->ipc_opened.clear();
    c->p2p = false; c->world = c->comm ? c->world : 1; c->rank = 0;
    drop_graphs(c);
    return SRM_OK;
}

// A row band (not the whole grid) updates from partial sums unless a collective is bound: refuse instead of diverging.
static int require_collective(srm_ctx *c, const char *who) {
    if (c->p2p || c->comm) return SRM_OK;
    if (c->world > 1)
        return fail(SRM_ERR_STATE, "%s: world = %d but neither the peer-memory nor the NCCL all-reduce is connected", who, c->world);
    if (c->g.row0 > 0 || c->g.row1 < c->g.n)
        return fail(SRM_ERR_STATE, "%s: rows [%d,%d) of %d are a row band; connect srm_p2p_connect / srm_nccl_init first, or "
                                   "step it with srm_label_accumulate / srm_acc_buffer / srm_update", who, c->g.row0, c->g.row1, c->g.n);
    return SRM_OK;
}

static int allreduce_acc(srm_ctx *c) {
    if (c->p2p) return SRM_OK;   // peer-memory mode: the update kernel pulls the partial sums itself
    if (!c->comm) return require_collective(c, "all-reduce");
    int e = g_nccl.AllReduce(c->acc, c->acc, 4 * (size_t)c->Kcap + 4, /*ncclFloat64*/ 8, /*ncclSum*/ 0, c->comm, c->stream);
    if (e) return fail(SRM_ERR_CUDA, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(e) : "error");
    return SRM_OK;
}

extern "C" int srm_set_stream(srm_ctx *c, void *cuda_stream) {
    if (!c) return fail(SRM_ERR_ARG, "null ctx");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    if (c->own_stream) cudaStreamDestroy(c->stream);
    c->stream = (cudaStream_t)cuda_stream;
    c->own_stream = false;
    return SRM_OK;
}

extern "C" int srm_synchronize(srm_ctx *c) {
    if (!c) return fail(SRM_ERR_ARG, "null ctx");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    return SRM_OK;
}

// pbaCVDComputeWeightedPrefix (gcvt.cu:995-1006), fp64, rows of this band only; c->density is already filled.
static int density_ready(srm_ctx *c) {
    srm_launch_prefix(c->stream, c->density, c->g, c->P2, c->PXX);
    if (!is_band(c)) srm_launch_nonzero_bits_f32(c->stream, c->density, c->N, c->nzbits);   // (bands: see srm_set_density*)
    CK(cudaGetLastError());
    c->has_density = true;
    return SRM_OK;
}

// Pinned (cudaHostAlloc / cudaHostRegister) memory is DMA-able as it is; pageable memory goes through the staging
// pipeline of srm_host.cu.
static bool host_ptr_is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

static int upload(srm_ctx *c, void *dst, const void *src, size_t bytes, int on_device) {
    if (on_device) { CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, c->stream)); return SRM_OK; }
    if (host_ptr_is_pinned(src) || bytes < (1u << 20)) {
        CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));   // the caller may reuse its buffer when we return
        return SRM_OK;
    }
    CK(srm_h2d_pageable(dst, src, bytes, c->stream));
    return SRM_OK;
}

extern "C" int srm_set_density(srm_ctx *c, const float *density, int on_device) {
    if (!c || !density) return fail(SRM_ERR_ARG, "srm_set_density: null argument");
    CK(cudaSetDevice(c->device));
    const int n = c->g.n, r0 = c->g.row0, r1 = c->g.row1;
    const size_t NB = (size_t)(r1 - r0) * n;
    int rc = upload(c, c->density, density + (size_t)r0 * n, NB * sizeof(float), on_device);
    if (rc) return rc;
    if (!is_band(c)) {
        // whole grid: density_ready() derives the bitmap from c->density
    } else if (on_device) {
        srm_launch_nonzero_bits_f32(c->stream, density, c->N, c->nzbits);   // straight from the caller's device array
    } else {
        // row band, full-grid host array: the band's rows stay on the device; the other rows only pass through a scratch
        // buffer for their "non-zero" bits.  (srm_set_density_band + an exchange of the bitmap slices avoids this.)
        srm_launch_nonzero_bits_f32(c->stream, c->density, NB, c->nzbits + (size_t)r0 * n / 32);
        float *tmp = nullptr;
        CK(cudaMalloc(&tmp, (size_t)64 * n * sizeof(float)));
        for (int b = 0; b < n && rc == SRM_OK; b += 64) {   // band boundaries are multiples of 64 rows
            if (b >= r0 && b < r1) continue;
            rc = upload(c, tmp, density + (size_t)b * n, (size_t)64 * n * sizeof(float), 0);
            if (rc == SRM_OK) srm_launch_nonzero_bits_f32(c->stream, tmp, (size_t)64 * n, c->nzbits + (size_t)b * n / 32);
            if (rc == SRM_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = fail(SRM_ERR_CUDA, "srm_set_density: sync failed");
        }
        cudaFree(tmp);
        if (rc) return rc;
    }
    CK(cudaGetLastError());
    return density_ready(c);
}

// Row bands: only the band's own rows ((row1 - row0) * n floats).  The "density != 0" bits of the OTHER rows must then be
// supplied by the caller: every rank fills its slice, the slices are exchanged through srm_shared_bits (all-gather).
extern "C" int srm_set_density_band(srm_ctx *c, const float *band_rows, int on_device) {
    if (!c || !band_rows) return fail(SRM_ERR_ARG, "srm_set_density_band: null argument");
    CK(cudaSetDevice(c->device));
    const size_t NB = (size_t)c->g.nrows() * c->g.n;
    int rc = upload(c, c->density, band_rows, NB * sizeof(float), on_device);
    if (rc) return rc;
    srm_launch_nonzero_bits_f32(c->stream, c->density, NB, c->nzbits + (size_t)c->g.row0 * c->g.n / 32);
    CK(cudaGetLastError());
    return density_ready(c);
}

// The host scans of the sparse inputs (srm_host.cu), for callers that shard them: sites of `pixels` seed-map pixels
// (packed x | y << 16 = the map's own values, row-major order) and the non-zero bytes of rows [row0, row1) of a mask.
// Return the number found in *count (which may exceed `capacity`; only `capacity` entries are written).
extern "C" int srm_scan_site_map_host(const short *site_map, size_t pixels, int *packed_out, int capacity, int *count) {
    if (!site_map || !count || (!packed_out && capacity > 0)) return fail(SRM_ERR_ARG, "srm_scan_site_map_host: bad argument");
    std::vector<int> v;
    srm_scan_site_map((const int *)site_map, pixels, v);
    *count = (int)v.size();
    if (packed_out) memcpy(packed_out, v.data(), sizeof(int) * std::min((size_t)capacity, v.size()));
    return SRM_OK;
}

extern "C" int srm_scan_mask_host(const unsigned char *mask, int n, int row0, int row1, int *packed_out, int capacity, int *count) {
    if (!mask || !count || n <= 0 || (n % 8) || row0 < 0 || row1 > n || row0 > row1 || (!packed_out && capacity > 0))
        return fail(SRM_ERR_ARG, "srm_scan_mask_host: bad argument");
    std::vector<int> v;
    srm_scan_mask(mask, n, v, row0, row1);
    *count = (int)v.size();
    if (packed_out) memcpy(packed_out, v.data(), sizeof(int) * std::min((size_t)capacity, v.size()));
    return SRM_OK;
}

// Device pointers of the two full-grid bitmaps (which = 0: density != 0, 1: constraint pixels): n*n/32 words, row y
// starts at word y*n/32.  For the exchange of the band slices between ranks.
extern "C" int srm_shared_bits(srm_ctx *c, int which, void **device_ptr, size_t *num_words) {
    if (!c || !device_ptr || !num_words || which < 0 || which > 1) return fail(SRM_ERR_ARG, "srm_shared_bits: bad argument");
    *device_ptr = which ? c->maskbits : c->nzbits;
    *num_words = c->N / 32;
    return SRM_OK;
}

// dense device mask from the list of constraint pixels (packed x | y << 16)
static int set_mask_pixels(srm_ctx *c, const std::vector<int> &px) {
    CK(cudaMemsetAsync(c->maskbits, 0, c->N / 8, c->stream));
    if (!px.empty()) {
        int *d = nullptr;
        CK(cudaMalloc(&d, px.size() * sizeof(int)));
        cudaError_t e = cudaMemcpyAsync(d, px.data(), px.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) { srm_launch_scatter_mask(c->stream, d, (int)px.size(), c->g.n, c->maskbits); e = cudaGetLastError(); }
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);   // px may be a local of the caller
        cudaFree(d);
        if (e != cudaSuccess) return fail(SRM_ERR_CUDA, "srm_set_mask: %s", cudaGetErrorString(e));
    }
    c->has_mask = true;
    return SRM_OK;
}

// Constraint pixels as a list (packed x | y << 16) instead of the 1 B/px mask: what srm_set_mask extracts from a host
// mask.  Row bands: every rank scans its own rows of the mask and the lists are concatenated by the caller.
extern "C" int srm_set_mask_pixels(srm_ctx *c, const int *packed_xy, int count) {
    if (!c || (!packed_xy && count > 0) || count < 0) return fail(SRM_ERR_ARG, "srm_set_mask_pixels: bad argument");
    CK(cudaSetDevice(c->device));
    for (int i = 0; i < count; ++i) {
        const int x = srm_x(packed_xy[i]), y = srm_y(packed_xy[i]);
        if (x < 0 || y < 0 || x >= c->g.n || y >= c->g.n) return fail(SRM_ERR_ARG, "srm_set_mask_pixels: pixel %d outside the grid", i);
    }
    std::vector<int> px(packed_xy, packed_xy + count);
    return set_mask_pixels(c, px);
}

extern "C" int srm_set_mask(srm_ctx *c, const unsigned char *mask, int on_device) {
    if (!c) return fail(SRM_ERR_ARG, "null ctx");
    CK(cudaSetDevice(c->device));
    if (!mask) { c->has_mask = false; return SRM_OK; }
    if (on_device) {
        srm_launch_nonzero_bits_u8(c->stream, mask, c->N, c->maskbits);
        CK(cudaGetLastError());
        c->has_mask = true;
        return SRM_OK;
    }
    // the mask marks a few constraint pixels (generateMask, gcvt.h:143-159): scan it on the host, upload the list
    std::vector<int> px;
    srm_scan_mask(mask, c->g.n, px);
    return set_mask_pixels(c, px);
}

extern "C" int srm_set_site_map(srm_ctx *c, const short *site_map, int on_device) {
    if (!c || !site_map) return fail(SRM_ERR_ARG, "srm_set_site_map: null argument");
    CK(cudaSetDevice(c->device));
    const int *dmap = (const int *)site_map;
    if (!on_device) {
        // a seed map holds a handful of sites in 4 B/px: scan it on the host (row-major order = the order of the device
        // compaction below, so the site ids are the same on both paths) and upload the list
        std::vector<int> sites;
        srm_scan_site_map((const int *)site_map, c->N, sites);
        return srm_set_sites(c, sites.data(), (int)sites.size(), 0);
    }
    int rc;
    srm_launch_sites_from_map(c->stream, dmap, c->N, nullptr, c->blockcnt, c->blockoff, &c->ctl->nlive, 1);
    CK(cudaGetLastError());
    SrmCtl h;
    CK(cudaMemcpyAsync(&h, c->ctl, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    const int K = h.nlive;
    rc = alloc_sites(c, K);
    if (rc) return rc;
    srm_launch_sites_from_map(c->stream, dmap, c->N, c->sites[0], c->blockcnt, c->blockoff, nullptr, 0);
    CK(cudaGetLastError());
    rc = reset_ctl(c, K);
    if (rc) return rc;
    rc = init_step_buffers(c, 0);
    if (rc) return rc;
    c->has_sites = true;
    return SRM_OK;
}

extern "C" int srm_set_sites(srm_ctx *c, const int *packed_xy, int num, int on_device) {
    if (!c || (!packed_xy && num > 0) || num < 0) return fail(SRM_ERR_ARG, "srm_set_sites: bad argument");
    CK(cudaSetDevice(c->device));
    if (!on_device)   // the kernels index the bitmap and the site-id map with these coordinates
        for (int i = 0; i < num; ++i) {
            const int v = packed_xy[i], x = srm_x(v), y = srm_y(v);
            if (v != SRM_SENT && (x < 0 || y < 0 || x >= c->g.n || y >= c->g.n))
                return fail(SRM_ERR_ARG, "srm_set_sites: site %d = (%d,%d) outside the %d x %d grid", i, x, y, c->g.n, c->g.n);
        }
    int rc = alloc_sites(c, num);
    if (rc) return rc;
    if (num > 0)
        CK(cudaMemcpyAsync(c->sites[0], packed_xy, (size_t)num * sizeof(int),
                           on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
    rc = reset_ctl(c, num);
    if (rc) return rc;
    rc = init_step_buffers(c, 0);
    if (rc) return rc;
    c->has_sites = true;
    return SRM_OK;
}

static int fetch_ctl(srm_ctx *c, SrmCtl *h) {
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(h, c->ctl, sizeof(*h), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->it_host = h->it;  // iterations enqueued after a device-side stop were no-ops
    c->stopped = h->stop != 0;
    c->unsynced = false;
    if (h->p2p_timeout)
        return fail(SRM_ERR_CUDA, "row-band all-reduce: a peer never signalled iteration %d (rank %d of %d gave up waiting); "
                                  "the site lists are no longer consistent", h->it + 1, c->rank, c->world);
    return SRM_OK;
}

// srm_iterate with the stopping rule does not wait for the device (batches of contexts are enqueued back to back): the
// host mirror of the iteration count — the buffer parity of everything that follows — is read back by the next call
// that needs it.
static int resync(srm_ctx *c) {
    if (!c->unsynced) return SRM_OK;
    SrmCtl h;
    return fetch_ctl(c, &h);
}

// The live site list is in sites[it & 1]: every executed iteration flips the buffer.
static int current_buffer(srm_ctx *c) { return c->it_host & 1; }

extern "C" int srm_get_sites(srm_ctx *c, int *packed_xy_host, int capacity, int *num_out) {
    if (!c || !num_out) return fail(SRM_ERR_ARG, "srm_get_sites: null argument");
    if (!c->has_sites) return fail(SRM_ERR_STATE, "srm_get_sites: no sites set");
    SrmCtl h;
    int rc = fetch_ctl(c, &h);
    if (rc) return rc;
    *num_out = h.nlive;
    if (packed_xy_host && h.K > 0) {  // the device list keeps holes for merged sites: filter them here
        std::vector<int> all((size_t)h.K);
        CK(cudaMemcpy(all.data(), c->sites[current_buffer(c)], all.size() * sizeof(int), cudaMemcpyDeviceToHost));
        int k = 0;
        for (int v : all)
            if (v != SRM_SENT && k < capacity) packed_xy_host[k++] = v;
    }
    return SRM_OK;
}

// Measurement helper: total number of runs of the last labelling and rows that took the robust path.
extern "C" int srm_debug_counts(srm_ctx *c, long long *total_runs, int *overflow_rows) {
    if (!c || !total_runs || !overflow_rows) return fail(SRM_ERR_ARG, "srm_debug_counts: null argument");
    CK(cudaSetDevice(c->device));
    std::vector<int> h((size_t)c->g.nrows());
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(h.data(), c->rle_cnt, h.size() * sizeof(int), cudaMemcpyDeviceToHost));
    long long tot = 0;
    int mx = 0;
    for (int v : h) { tot += v; if (v > mx) mx = v; }
    *total_runs = tot;
    if (c->dbg_stats) fprintf(stderr, "[srm dbg] runs per row: mean %.1f max %d\n", (double)tot / (double)h.size(), mx);
    SrmCtl hc;
    CK(cudaMemcpy(&hc, c->ctl, sizeof(hc), cudaMemcpyDeviceToHost));
    *overflow_rows = c->robust_only ? c->g.nrows() : hc.ovf;
    if (c->dbg_stats & 1) {
        const unsigned long long *p = hc.prof;
        const double tot = (double)(p[0] + p[1] + p[2] + p[3] + p[4] + p[5]);
        fprintf(stderr, "[srm prof] warp-cycles: phaseA %.1f%% prefilter0 %.1f%% plain0 %.1f%% rounds %.1f%% move %.1f%% output %.1f%% | "
                        "prefilter survivors/row %.0f (%llu rows) plain survivors/row %.0f (%llu rows) | round steps/row %.1f passes/row %.2f "
                        "envelope/row %.0f\n",
                100 * p[0] / tot, 100 * p[1] / tot, 100 * p[2] / tot, 100 * p[3] / tot, 100 * p[4] / tot, 100 * p[5] / tot,
                p[9] ? (double)p[8] / p[9] : 0.0, p[9], p[11] ? (double)p[10] / p[11] : 0.0, p[11],
                (double)p[12] / (p[9] + p[11] + 1e-9), (double)p[13] / (p[9] + p[11] + 1e-9), (double)p[14] / (p[9] + p[11] + 1e-9));
    }
    if (c->dbg_stats)
        fprintf(stderr, "[srm dbg] band list: max %d mean %.1f (%d bands) | row survivors: max %d mean %.1f (%d rows)\n",
                hc.dbg[0], hc.dbg[2] ? (double)hc.dbg[1] / hc.dbg[2] : 0.0, hc.dbg[2], hc.dbg[3],
                hc.dbg[5] ? (double)hc.dbg[4] / hc.dbg[5] : 0.0, hc.dbg[5]);
    return SRM_OK;
}

// Statistics counter SrmCtl::dbg[which] (collected while the option "dbg_stats" is on): 0 max / 1 sum of the band-list
// length, 2 bands, 6 warps that took the staging-overflow fallback of the band kernel's Phase A.
extern "C" int srm_debug_get(srm_ctx *c, int which, long long *value) {
    if (!c || !value || which < 0 || which >= 8) return fail(SRM_ERR_ARG, "srm_debug_get: bad argument");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    int v = 0;
    CK(cudaMemcpy(&v, &c->ctl->dbg[which], sizeof(int), cudaMemcpyDeviceToHost));
    *value = v;
    return SRM_OK;
}

// Measurement / tests: the band kernel's CTA -> band order and the cost (runs per band of the last labelling) it is rebuilt from.
extern "C" int srm_debug_band_order(srm_ctx *c, int *perm_out, int *cost_out, int capacity, int *num_bands) {
    if (!c || !num_bands) return fail(SRM_ERR_ARG, "srm_debug_band_order: null argument");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    const int nb = c->g.nrows() / 8;
    *num_bands = nb;
    if (capacity < nb) return SRM_OK;
    if (perm_out) CK(cudaMemcpy(perm_out, c->band_perm, (size_t)nb * sizeof(int), cudaMemcpyDeviceToHost));
    if (cost_out) {
        std::vector<int> h((size_t)c->g.nrows());
        CK(cudaMemcpy(h.data(), c->rle_cnt, h.size() * sizeof(int), cudaMemcpyDeviceToHost));
        for (int b = 0; b < nb; ++b) {
            int sum = 0;
            for (int r = 0; r < 8; ++r) sum += std::min(std::max(h[(size_t)b * 8 + r], 0), 32767);
            cost_out[b] = sum;
        }
    }
    return SRM_OK;
}

// Options: "robust_only" (0/1): label with the worst-case-capacity row path only (no fused band kernel).
extern "C" int srm_set_option(srm_ctx *c, const char *name, int value) {
    if (!c || !name) return fail(SRM_ERR_ARG, "srm_set_option: null argument");
    if (!strcmp(name, "robust_only")) { c->robust_only = value != 0; return SRM_OK; }
    if (!strcmp(name, "graph")) { c->use_graph = value != 0; return SRM_OK; }
    if (!strcmp(name, "jfa_mode")) { c->jfa_mode = value & 3; return SRM_OK; }
    if (!strcmp(name, "band_order")) { c->band_order = value != 0; drop_graphs(c); return SRM_OK; }
    if (!strcmp(name, "dbg_stats")) {
        c->dbg_stats = value;
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaMemset(c->ctl->dbg, 0, sizeof(c->ctl->dbg)));
        CK(cudaMemset(c->ctl->prof, 0, sizeof(c->ctl->prof)));
        return SRM_OK;
    }
    return fail(SRM_ERR_ARG, "srm_set_option: unknown option %s", name);
}

// Tuning of the host side of the boundary (srm_host.cu): worker threads of the pageable-copy pipeline and of the host
// scans (1..16, 0 = default), staging chunk size in KB (256..65536, 0 = unchanged).
extern "C" int srm_host_config(int threads, int chunk_kb) {
    if (srm_host_set_config(threads, chunk_kb)) return fail(SRM_ERR_ARG, "srm_host_config: threads %d / chunk %d KB out of range", threads, chunk_kb);
    return SRM_OK;
}

// Process-wide choice between two builds of a streaming kernel (measurement tools and A/B tests): which = "expand"
// (runs -> dense labels: 0 one binary search per 4-pixel group, 1 two-level lookup) or "prefix" (fp64 prefix sums:
// 0 128/64-bit stores, 1 256-bit stores); value < 0 = back to the environment / compiled default.
extern int g_srm_expand_v, g_srm_prefix_v, g_srm_centroid_v;
extern "C" int srm_set_variant(const char *which, int value) {
    if (!which) return fail(SRM_ERR_ARG, "srm_set_variant: null argument");
    if (!strcmp(which, "expand")) { g_srm_expand_v = value < 0 ? -1 : (value > 2 ? 2 : value); return SRM_OK; }
    if (!strcmp(which, "prefix")) { g_srm_prefix_v = value < 0 ? -1 : (value != 0); return SRM_OK; }
    if (!strcmp(which, "centroid")) { g_srm_centroid_v = value < 0 ? -1 : (value != 0); return SRM_OK; }
    return fail(SRM_ERR_ARG, "srm_set_variant: unknown kernel %s", which);
}

// Site extraction in the order delaunayInput scans the label map (delaunay.h:46-57: x outer, y inner; sites are
// pixels with label == self that are not constraint pixels; point = (x*scale + l, y*scale + b)).  The reference
// downloads the 2N-short label map and scans it on the host; here the K-entry site list is read back instead.
extern "C" int srm_extract_sites(srm_ctx *c, const unsigned char *mask_host, double scale, double l, double b,
                                 double *points_xy, int capacity, int *num_out) {
    if (!c || !num_out) return fail(SRM_ERR_ARG, "srm_extract_sites: null argument");
    if (!c->has_sites) return fail(SRM_ERR_STATE, "srm_extract_sites: no sites set");
    SrmCtl h;
    int rc = fetch_ctl(c, &h);
    if (rc) return rc;
    std::vector<int> all((size_t)(h.K > 0 ? h.K : 0));
    if (h.K > 0) CK(cudaMemcpy(all.data(), c->sites[current_buffer(c)], all.size() * sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<unsigned> keys;  // x in the high half: sorting by key = x outer, y inner
    keys.reserve(all.size());
    const int n = c->g.n;
    for (int v : all) {
        if (v == SRM_SENT) continue;
        const int x = srm_x(v), y = srm_y(v);
        if (mask_host && mask_host[(size_t)y * n + x]) continue;
        keys.push_back(((unsigned)x << 16) | (unsigned)y);
    }
    std::sort(keys.begin(), keys.end());
    *num_out = (int)keys.size();
    if (points_xy) {
        const int k = std::min(capacity, (int)keys.size());
        for (int i = 0; i < k; ++i) {
            points_xy[2 * i] = (double)(keys[i] >> 16) * scale + l;
            points_xy[2 * i + 1] = (double)(keys[i] & 0xffffu) * scale + b;
        }
    }
    return SRM_OK;
}

extern "C" int srm_set_omega(srm_ctx *c, float omega) {
    if (!c) return fail(SRM_ERR_ARG, "null ctx");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(&c->ctl->omega, &omega, sizeof(float), cudaMemcpyHostToDevice));
    return SRM_OK;
}

// it_host counts enqueued iterations; the buffer parity on the device is (executed iterations) & 1.
// Inside the loop both agree until a stop; after a stop every kernel is a no-op, so using the
// host parity for the (skipped) launches is harmless.  For calls outside the loop (final labelling)
// the parity is read back from the device.
static SrmRle rle_of(srm_ctx *c) {
    SrmRle R;
    R.pool = c->rle; R.off = c->rle_off; R.cnt = c->rle_cnt; R.cap = c->rle_cap;
    R.scratch = c->row_scratch; R.scratch_ctas = srm_row_scratch_ctas(c->g.nrows());
    return R;
}

static int band_flags(srm_ctx *c, int respect_stop, int accumulate, int want_energy, int write_rle) {
    return (accumulate ? SRM_BF_ACC : 0) | (want_energy ? SRM_BF_ENERGY : 0) | (respect_stop ? SRM_BF_STOP : 0) |
           (write_rle ? SRM_BF_RLE : 0) | (c->p2p ? SRM_BF_TOUCH : 0);
}

// write_rle: the run-length rows are needed by srm_get_labels / srm_accumulate after this labelling (not inside the loop)
// in_loop: an iteration of the Lloyd loop (the band order is rebuilt from the run counts of iteration it - 1 when
// it % 10 == 1: the period of the captured graphs)
static int label_with(srm_ctx *c, int it, int respect_stop, int accumulate, int want_energy, int write_rle = 1,
                      int in_loop = 0) {
    double *acc = cur_acc(c, it);
    const SrmStep s = step_of(c, it);
    srm_launch_carry(c->stream, s, c->g.n, c->up, c->dn, c->ctl, respect_stop, c->g.row0, c->g.row1);
    const int *rows = nullptr, *count = nullptr;
    if (!c->robust_only) {
        CK(srm_launch_band(c->stream, s.bits, c->up, c->dn, c->g, rle_of(c), c->ovf_rows, c->P2, c->PXX,
                           s.hash, acc, c->Kcap, c->ctl, band_flags(c, respect_stop, accumulate, want_energy, write_rle),
                           c->dbg_stats, perm_of(c), in_loop && it % 10 == 1));
        rows = c->ovf_rows;
        count = &c->ctl->ovf;
    }
    // (fused all-reduce: the last CTA of the row kernel signals the peers that this rank's sums are complete)
    CK(srm_launch_row(c->stream, s.bits, c->up, c->dn, c->g, rle_of(c), rows, count, c->P2, c->PXX, s.hash,
                      acc, c->Kcap, c->ctl, accumulate, want_energy, respect_stop, write_rle,
                      accumulate ? peers_of(c, it) : SrmPeers()));
    return SRM_OK;
}

static int require_ready(srm_ctx *c, const char *who, bool need_density) {
    if (!c) return fail(SRM_ERR_ARG, "%s: null ctx", who);
    if (!c->has_sites) return fail(SRM_ERR_STATE, "%s: sites not set", who);
    if (need_density && !c->has_density) return fail(SRM_ERR_STATE, "%s: density not set", who);
    return SRM_OK;
}

// After a labelling that wrote run-length rows: if a row found the pool exhausted (more runs per row on average than
// the band kernel's buffer holds: adversarial site sets only), grow the pool to the worst case (one run per pixel) and
// label again.  Synchronises the stream.
static int ensure_rle(srm_ctx *c) {
    CK(cudaStreamSynchronize(c->stream));
    int failed = 0;
    CK(cudaMemcpy(&failed, &c->ctl->rle_fail, sizeof(int), cudaMemcpyDeviceToHost));
    if (!failed) return SRM_OK;
    const size_t full = (size_t)c->g.nrows() * c->g.n;
    if ((size_t)c->rle_cap >= full) return fail(SRM_ERR_STATE, "run-length pool exhausted at its full size");
    cudaFree(c->rle);
    c->rle = nullptr;
    CK(cudaMalloc(&c->rle, full * sizeof(int2)));
    c->rle_cap = (int)full;
    failed = 0;
    CK(cudaMemcpy(&c->ctl->rle_fail, &failed, sizeof(int), cudaMemcpyHostToDevice));
    int rc = label_with(c, c->it_host, 0, 0, 0);   // labels only: the sums of the first pass are complete
    if (rc) return rc;
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(&failed, &c->ctl->rle_fail, sizeof(int), cudaMemcpyDeviceToHost));
    if (failed) return fail(SRM_ERR_STATE, "run-length pool exhausted at its full size");
    return SRM_OK;
}

extern "C" int srm_label(srm_ctx *c) {
    int rc = require_ready(c, "srm_label", false);
    if (rc) return rc;
    rc = resync(c);
    if (rc) return rc;
    CK(cudaSetDevice(c->device));
    rc = label_with(c, c->it_host, 0, 0, 0);
    if (rc) return rc;
    rc = ensure_rle(c);
    if (rc) return rc;
    c->labelled = true;
    return SRM_OK;
}

// Fused stepwise variant: labelling and per-site accumulation in one pass of the band kernel (what
// srm_iterate does per iteration), for callers that all-reduce the accumulators before srm_update.
extern "C" int srm_label_accumulate(srm_ctx *c, int want_energy) {
    int rc = require_ready(c, "srm_label_accumulate", true);
    if (rc) return rc;
    rc = resync(c);
    if (rc) return rc;
    CK(cudaSetDevice(c->device));
    rc = label_with(c, c->it_host, 0, 1, want_energy);
    if (rc) return rc;
    rc = ensure_rle(c);
    if (rc) return rc;
    c->labelled = true;
    return SRM_OK;
}

extern "C" int srm_accumulate(srm_ctx *c, int want_energy) {
    int rc = require_ready(c, "srm_accumulate", true);
    if (rc) return rc;
    rc = resync(c);
    if (rc) return rc;
    if (!c->labelled) return fail(SRM_ERR_STATE, "srm_accumulate: call srm_label first");
    CK(cudaSetDevice(c->device));
    srm_launch_acc(c->stream, rle_of(c), c->P2, c->PXX, c->hash[c->it_host & 1], c->g, cur_acc(c, c->it_host), c->Kcap,
                   nullptr, nullptr, c->ctl, want_energy, 0);
    srm_launch_signal(c->stream, c->ctl, peers_of(c, c->it_host), 0);
    CK(cudaGetLastError());
    return SRM_OK;
}

// The centroid pass in north_star's stand-alone form (srm_centroid.cu): per-site sums of this band from a DENSE label map
// on the device and the density, one streaming pass.  labels_dev == nullptr: the run-length labels of the last srm_label
// are expanded into the context's own dense buffer first (the reference's data flow: label map -> sums).  Adds to the
// same accumulators as srm_accumulate; srm_update follows as usual.
static int expand_own_labels(srm_ctx *c, const char *who) {
    if (!c->labelled) return fail(SRM_ERR_STATE, "%s: call srm_label first (or pass a label map)", who);
    if (!c->labels) CK(cudaMalloc(&c->labels, (size_t)c->g.nrows() * c->g.n * sizeof(int)));
    CK(srm_launch_expand(c->stream, rle_of(c), c->g, c->labels));
    return SRM_OK;
}

extern "C" int srm_accumulate_dense(srm_ctx *c, const short *labels_dev, int want_energy) {
    int rc = require_ready(c, "srm_accumulate_dense", true);
    if (rc) return rc;
    rc = resync(c);
    if (rc) return rc;
    CK(cudaSetDevice(c->device));
    const int *lab = reinterpret_cast<const int *>(labels_dev);
    if (!lab) {
        rc = expand_own_labels(c, "srm_accumulate_dense");
        if (rc) return rc;
        lab = c->labels;
    }
    if ((reinterpret_cast<uintptr_t>(lab) & 15) != 0) return fail(SRM_ERR_ARG, "srm_accumulate_dense: the label map must be 16-byte aligned");
    CK(srm_launch_centroid_dense(c->stream, lab, c->density, c->hash[c->it_host & 1], c->g, cur_acc(c, c->it_host), c->Kcap,
                                 want_energy, c->p2p ? 1 : 0));
    srm_launch_signal(c->stream, c->ctl, peers_of(c, c->it_host), 0);
    CK(cudaGetLastError());
    return SRM_OK;
}

extern "C" int srm_acc_buffer(srm_ctx *c, void **device_ptr, size_t *num_doubles) {
    if (!c || !device_ptr || !num_doubles) return fail(SRM_ERR_ARG, "srm_acc_buffer: null argument");
    if (!c->has_sites) return fail(SRM_ERR_STATE, "srm_acc_buffer: sites not set");
    { int rc = resync(c); if (rc) return rc; }
    *device_ptr = cur_acc(c, c->it_host);
    *num_doubles = 4 * (size_t)c->Kcap + 4;
    return SRM_OK;
}

// Stepwise update (manual loop: label / accumulate / [all-reduce] / update).  No host sync: the
// iteration count is mirrored on the host; the stopping rule is not applied in stepwise mode (the
// caller reads srm_get_state when it wants to stop), omega still follows gcvt.cu:1131.
extern "C" int srm_update(srm_ctx *c) {
    int rc = require_ready(c, "srm_update", true);
    if (rc) return rc;
    rc = resync(c);
    if (rc) return rc;
    CK(cudaSetDevice(c->device));
    // (stepwise mode with world > 1 and no bound collective is the caller-side all-reduce of ShardedLloyd.step:
    //  srm_acc_buffer hands out the sums, so nothing to check here)
    const int buf = current_buffer(c);
    srm_launch_update(c->stream, c->sites[buf], c->sites[buf ^ 1], c->acc, c->nzbits, c->has_mask ? c->maskbits : nullptr,
                      c->g, c->ctl, c->Kcap, c->newpos, step_of(c, c->it_host), (c->it_host % 10) == 0, 0, 0,
                      peers_of(c, c->it_host));
    CK(cudaGetLastError());
    c->labelled = false;
    c->it_host += 1;
    return SRM_OK;
}

// One Lloyd iteration on the context's stream (the body of the loop at gcvt.cu:1112-1123).
static int one_iteration(srm_ctx *c, int it, int stop_rule) {
    const int buf = it & 1, want_energy = (it % 10) == 0;
    int rc = label_with(c, it, 1, 1, want_energy, /*write_rle=*/0, /*in_loop=*/1);
    if (rc) return rc;
    rc = allreduce_acc(c);  // no-op for a single band and in peer-memory mode
    if (rc) return rc;
    srm_launch_update(c->stream, c->sites[buf], c->sites[buf ^ 1], c->acc, c->nzbits,
                      c->has_mask ? c->maskbits : nullptr, c->g, c->ctl, c->Kcap, c->newpos, step_of(c, it), want_energy,
                      stop_rule, 1, peers_of(c, it));
    return SRM_OK;
}

// Graph of iterations 10 q .. 10 q + 9: the kernel arguments depend on the iteration only through its parity and
// through "every 10th computes the energy", and the stop flag / iteration counter live on the device, so one graph
// serves every block of ten.  Captured without the PDL attribute (plain kernel-to-kernel edges).
static int capture_graph(srm_ctx *c, int stop_rule) {
    const int key = (c->Kcap << 3) ^ (c->has_mask ? 1 : 0) ^ (c->robust_only ? 2 : 0) ^ (c->dbg_stats ? 4 : 0) ^ (c->p2p ? 0x40000000 : 0) ^
                    (c->band_order ? 0x20000000 : 0);
    if (key != c->graph_key) { drop_graphs(c); c->graph_key = key; }
    if (c->graph[stop_rule]) return SRM_OK;
    cudaGraph_t g = nullptr;
    g_pdl_suppress = true;
    const long long count0 = srm_launch_count();
    cudaError_t e = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal);
    int rc = SRM_OK;
    if (e == cudaSuccess) {
        for (int it = 0; it < 10 && rc == SRM_OK; ++it) rc = one_iteration(c, it, stop_rule);
        e = cudaStreamEndCapture(c->stream, &g);
    }
    g_pdl_suppress = false;
    c->graph_kernels[stop_rule] = (int)(srm_launch_count() - count0);
    SRM_COUNT_N(-c->graph_kernels[stop_rule]);   // captured, not launched
    if (rc) { if (g) cudaGraphDestroy(g); return rc; }
    if (e != cudaSuccess) { cudaGetLastError(); return fail(SRM_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e)); }
    e = cudaGraphInstantiate(&c->graph[stop_rule], g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) { c->graph[stop_rule] = nullptr; return fail(SRM_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e)); }
    return SRM_OK;
}

extern "C" int srm_iterate(srm_ctx *c, int iters, int stop_rule) {
    int rc = require_ready(c, "srm_iterate", true);
    if (rc) return rc;
    rc = require_collective(c, "srm_iterate");
    if (rc) return rc;
    rc = resync(c);
    if (rc) return rc;
    CK(cudaSetDevice(c->device));
    if (c->stopped) return SRM_OK;
    int it = c->it_host;
    const int end = it + iters;
    const bool graphs = c->use_graph && !c->comm;   // (an NCCL all-reduce inside the loop is not captured)
    while (it < end) {
        if (graphs && it % 10 == 0 && end - it >= 10) {
            rc = capture_graph(c, stop_rule ? 1 : 0);
            if (rc) return rc;
            CK(cudaGraphLaunch(c->graph[stop_rule ? 1 : 0], c->stream));
            SRM_COUNT_N(c->graph_kernels[stop_rule ? 1 : 0]);
            it += 10;
            continue;
        }
        rc = one_iteration(c, it, stop_rule);
        if (rc) return rc;
        ++it;
    }
    CK(cudaGetLastError());
    c->it_host = it;
    c->labelled = false;
    c->unsynced = stop_rule != 0;   // the device may stop early: the next call that needs the count reads it back
    return SRM_OK;
}

// Same loop as srm_iterate with CUDA events between the stages; stage_ms[6] receives the summed device
// time of {site bitmap + carries, fused band kernel, robust row path, accumulator all-reduce (row bands),
// update + control, whole iteration}.
extern "C" int srm_iterate_profiled(srm_ctx *c, int iters, int stop_rule, float *stage_ms) {
    int rc = require_ready(c, "srm_iterate_profiled", true);
    if (rc) return rc;
    if (!stage_ms || iters <= 0) return fail(SRM_ERR_ARG, "srm_iterate_profiled: bad argument");
    rc = require_collective(c, "srm_iterate_profiled");
    if (rc) return rc;
    rc = resync(c);
    if (rc) return rc;
    CK(cudaSetDevice(c->device));
    for (int k = 0; k < 6; ++k) stage_ms[k] = 0;
    if (c->stopped) return SRM_OK;
    std::vector<cudaEvent_t> ev((size_t)iters * 6);
    for (auto &e : ev) CK(cudaEventCreate(&e));
    int it = c->it_host;
    for (int i = 0; i < iters; ++i, ++it) {
        const int buf = it & 1, want_energy = (it % 10) == 0;
        cudaEvent_t *e = &ev[(size_t)i * 6];
        CK(cudaEventRecord(e[0], c->stream));
        double *acc = cur_acc(c, it);
        const SrmStep s = step_of(c, it);
        srm_launch_carry(c->stream, s, c->g.n, c->up, c->dn, c->ctl, 1, c->g.row0, c->g.row1);
        CK(cudaEventRecord(e[1], c->stream));
        const int *rows = nullptr, *count = nullptr;
        if (!c->robust_only) {
            CK(srm_launch_band(c->stream, s.bits, c->up, c->dn, c->g, rle_of(c), c->ovf_rows, c->P2, c->PXX,
                               s.hash, acc, c->Kcap, c->ctl, band_flags(c, 1, 1, want_energy, 0), c->dbg_stats, perm_of(c),
                               it % 10 == 1));
            rows = c->ovf_rows;
            count = &c->ctl->ovf;
        }
        CK(cudaEventRecord(e[2], c->stream));
        CK(srm_launch_row(c->stream, s.bits, c->up, c->dn, c->g, rle_of(c), rows, count, c->P2, c->PXX, s.hash,
                          acc, c->Kcap, c->ctl, 1, want_energy, 1, 0, peers_of(c, it)));
        CK(cudaEventRecord(e[3], c->stream));
        rc = allreduce_acc(c);
        if (rc) return rc;
        CK(cudaEventRecord(e[4], c->stream));
        srm_launch_update(c->stream, c->sites[buf], c->sites[buf ^ 1], c->acc, c->nzbits,
                          c->has_mask ? c->maskbits : nullptr, c->g, c->ctl, c->Kcap, c->newpos, s, want_energy,
                          stop_rule, 1, peers_of(c, it));
        CK(cudaEventRecord(e[5], c->stream));
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < iters; ++i) {
        cudaEvent_t *e = &ev[(size_t)i * 6];
        for (int k = 0; k < 5; ++k) {
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e[k], e[k + 1]));
            stage_ms[k] += ms;
        }
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e[0], e[5]));
        stage_ms[5] += ms;
    }
    for (auto &e : ev) cudaEventDestroy(e);
    c->it_host = it;
    c->labelled = false;
    SrmCtl h;
    return fetch_ctl(c, &h);
}

static void fill_stats(const SrmCtl &h, srm_stats *s, float ms) {
    if (!s) return;
    s->iterations = h.it; s->num_sites = h.nlive; s->stopped = h.stop; s->omega = h.omega; s->energy = h.E; s->ms_device = ms;
}

extern "C" int srm_get_state(srm_ctx *c, srm_stats *stats) {
    if (!c || !stats) return fail(SRM_ERR_ARG, "srm_get_state: null argument");
    SrmCtl h;
    int rc = fetch_ctl(c, &h);
    if (rc) return rc;
    fill_stats(h, stats, 0.0f);
    return SRM_OK;
}

extern "C" int srm_run(srm_ctx *c, int max_iter, int stop_rule, srm_stats *stats) {
    int rc = require_ready(c, "srm_run", true);
    if (rc) return rc;
    CK(cudaSetDevice(c->device));
    CK(cudaEventRecord(c->ev0, c->stream));
    // the loop always runs at least one iteration (do/while, gcvt.cu:1112-1142)
    rc = srm_iterate(c, std::max(1, max_iter), stop_rule);
    if (rc) return rc;
    rc = srm_label(c);  // final pba2DCompute (gcvt.cu:1149)
    if (rc) return rc;
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaEventSynchronize(c->ev1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    SrmCtl h;
    rc = fetch_ctl(c, &h);
    if (rc) return rc;
    fill_stats(h, stats, ms);
    return SRM_OK;
}

extern "C" int srm_get_labels(srm_ctx *c, short *out, int on_device) {
    if (!c || !out) return fail(SRM_ERR_ARG, "srm_get_labels: null argument");
    if (!c->labelled) return fail(SRM_ERR_STATE, "srm_get_labels: call srm_label first");
    CK(cudaSetDevice(c->device));
    const size_t NB = (size_t)c->g.nrows() * c->g.n;
    int *dst = on_device ? (int *)out : nullptr;
    if (!on_device) {
        if (!c->labels) CK(cudaMalloc(&c->labels, NB * sizeof(int)));
        dst = c->labels;
    }
    CK(srm_launch_expand(c->stream, rle_of(c), c->g, dst));
    if (!on_device) {
        if (host_ptr_is_pinned(out)) CK(cudaMemcpyAsync(out, dst, NB * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        else CK(srm_d2h_pageable(out, dst, NB * sizeof(int), c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));
    return SRM_OK;
}

// Measurement: device milliseconds per launch of one of the two streaming kernels on this context's resident data, `reps`
// launches between two events.  which = "prefix" (needs the density) or "expand" (needs a labelling with run-length rows).
extern "C" int srm_time_kernel(srm_ctx *c, const char *which, int reps, float *ms_per_launch) {
    if (!c || !which || reps <= 0 || !ms_per_launch) return fail(SRM_ERR_ARG, "srm_time_kernel: bad argument");
    CK(cudaSetDevice(c->device));
    const bool prefix = !strcmp(which, "prefix"), centroid = !strcmp(which, "centroid") || !strcmp(which, "centroid_energy");
    if (centroid) {   // the stand-alone centroid pass over this context's dense labels; the accumulators are cleared afterwards
        if (!c->has_density || !c->has_sites) return fail(SRM_ERR_STATE, "srm_time_kernel: density / sites not set");
        int rc = resync(c);
        if (rc) return rc;
        rc = expand_own_labels(c, "srm_time_kernel");
        if (rc) return rc;
        double *acc = cur_acc(c, c->it_host);
        for (int i = -1; i < reps; ++i) {
            if (i == 0) CK(cudaEventRecord(c->ev0, c->stream));
            CK(srm_launch_centroid_dense(c->stream, c->labels, c->density, c->hash[c->it_host & 1], c->g, acc, c->Kcap,
                                         !strcmp(which, "centroid_energy"), 0));
        }
        CK(cudaEventRecord(c->ev1, c->stream));
        CK(cudaMemsetAsync(acc, 0, c->acc_stride * sizeof(double), c->stream));
        CK(cudaEventSynchronize(c->ev1));
        CK(cudaStreamSynchronize(c->stream));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        *ms_per_launch = ms / (float)reps;
        return SRM_OK;
    }
    if (!prefix && strcmp(which, "expand")) return fail(SRM_ERR_ARG, "srm_time_kernel: unknown kernel %s", which);
    if (prefix && !c->has_density) return fail(SRM_ERR_STATE, "srm_time_kernel: density not set");
    if (!prefix && !c->labelled) return fail(SRM_ERR_STATE, "srm_time_kernel: call srm_label first");
    if (!prefix && !c->labels) CK(cudaMalloc(&c->labels, (size_t)c->g.nrows() * c->g.n * sizeof(int)));
    for (int i = -1; i < reps; ++i) {   // one untimed launch first
        if (i == 0) CK(cudaEventRecord(c->ev0, c->stream));
        if (prefix) srm_launch_prefix(c->stream, c->density, c->g, c->P2, c->PXX);
        else CK(srm_launch_expand(c->stream, rle_of(c), c->g, c->labels));
    }
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaEventSynchronize(c->ev1));
    CK(cudaGetLastError());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    *ms_per_launch = ms / (float)reps;
    return SRM_OK;
}

// Seeds the dense map `a` from the site list and runs the JFA schedule between a and b (srm_jfa.cu); *res = the buffer
// that holds the labels.
static int jfa_run(srm_ctx *c, const char *who, const int *steps, int nsteps, int mode, int **res, cudaEvent_t *ev, int evcap,
                   int *nlaunch) {
    int rc = require_ready(c, who, false);
    if (rc) return rc;
    rc = resync(c);
    if (rc) return rc;
    if (!steps || nsteps <= 0) return fail(SRM_ERR_ARG, "%s: bad argument", who);
    if (c->g.row0 != 0 || c->g.row1 != c->g.n) return fail(SRM_ERR_ARG, "%s: whole-grid contexts only", who);
    for (int s = 0; s < nsteps; ++s)
        if (steps[s] <= 0 || steps[s] >= c->g.n) return fail(SRM_ERR_ARG, "%s: step %d", who, steps[s]);
    CK(cudaSetDevice(c->device));
    if (!c->labels) CK(cudaMalloc(&c->labels, c->N * sizeof(int)));
    if (!c->scratch_map) CK(cudaMalloc(&c->scratch_map, c->N * sizeof(int)));
    srm_launch_fill_int(c->stream, c->scratch_map, c->N, SRM_SENT);
    srm_launch_scatter_sites(c->stream, c->sites[current_buffer(c)], c->ctl, c->Kcap, c->g.n, c->scratch_map);
    cudaError_t e = cudaSuccess;
    *res = srm_launch_jfa(c->stream, c->scratch_map, c->labels, c->g.n, steps, nsteps, mode, ev, evcap, nlaunch, &e);
    CK(e);
    return SRM_OK;
}

extern "C" int srm_label_jfa(srm_ctx *c, const int *steps, int nsteps, short *out, int on_device) {
    if (!out) return fail(SRM_ERR_ARG, "srm_label_jfa: bad argument");
    int *a = nullptr;
    int rc = jfa_run(c, "srm_label_jfa", steps, nsteps, c ? c->jfa_mode : 0, &a, nullptr, 0, nullptr);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, a, c->N * sizeof(int), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                       c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return SRM_OK;
}

// Measurement: the same schedule with an event between the launches; ms[i] = device time of launch i (a fused run of
// small steps is one launch).  The labels stay on the device (srm_label_jfa returns them).
extern "C" int srm_label_jfa_timed(srm_ctx *c, const int *steps, int nsteps, int mode, float *ms, int cap, int *nlaunch) {
    if (!ms || !nlaunch || cap <= 0) return fail(SRM_ERR_ARG, "srm_label_jfa_timed: bad argument");
    std::vector<cudaEvent_t> ev((size_t)nsteps + 1, nullptr);
    for (auto &x : ev) CK(cudaEventCreate(&x));
    int *a = nullptr, nl = 0;
    int rc = jfa_run(c, "srm_label_jfa_timed", steps, nsteps, mode, &a, ev.data(), (int)ev.size(), &nl);
    if (rc == SRM_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = fail(SRM_ERR_CUDA, "srm_label_jfa_timed: %s", cudaGetErrorString(cudaGetLastError()));
    if (rc == SRM_OK) {
        *nlaunch = nl;
        for (int i = 0; i < nl && i < cap; ++i) cudaEventElapsedTime(&ms[i], ev[(size_t)i], ev[(size_t)i + 1]);
    }
    for (auto &x : ev) cudaEventDestroy(x);
    return rc;
}

// ------------------------------------------------------------------ one-shot drop-ins

static std::mutex g_cache_mu;
static srm_ctx *g_cache = nullptr;
static int g_cache_n = 0, g_cache_dev = -1;

extern "C" int srm_release_cache(void) {
    std::lock_guard<std::mutex> lock(g_cache_mu);
    if (g_cache) { srm_destroy(g_cache); g_cache = nullptr; }
    srm_host_pool_release();
    return SRM_OK;
}

// Loop state handed from one multires level to the next (gcvt.cu:1105-1147 keeps one set of host variables
// across the levels: the iteration counter, omega, Energy and lastEnergy are NOT reset at a level switch).
static int carry_ctl(srm_ctx *c, const SrmCtl &prev, int level) {
    CK(cudaStreamSynchronize(c->stream));
    SrmCtl h;
    CK(cudaMemcpy(&h, c->ctl, sizeof(h), cudaMemcpyDeviceToHost));
    h.it = prev.it; h.omega = prev.omega; h.lastE = prev.lastE; h.E = prev.E; h.stop = 0;
    h.escale = (float)(1 << (2 * level));         // powf(2, 2 pbaScale), gcvt.cu:1082
    h.thresh = level ? 3e-1 : 1e-5;               // gcvt.cu:1132-1137
    CK(cudaMemcpy(c->ctl, &h, sizeof(h), cudaMemcpyHostToDevice));
    c->it_host = prev.it;
    c->stopped = false;
    return SRM_OK;
}

// Coarse-to-fine gCVT (depth > 1; gcvt.cu:985-993, 1036-1051, 1087-1156).  Level L runs in its own context of side
// n >> L: density = L-times box-filtered input (on device), constraint mask = the first (n >> L)^2 bytes of the
// caller's mask read as an image of that side (the reference indexes the full-resolution mask with the level's
// size, gcvt.cu:1032-1033 / 764-777), sites = the previous level's sites with doubled coordinates.
// voronoi in: seed map of side n >> (depth-1) in the first entries of the buffer (gcvt.cu:1101-1103).
static int gcvt_multires(srm_ctx *c0, short *voronoi, const float *density, const unsigned char *mask, int depth,
                         int max_iter, srm_stats *stats) {
    const int n = c0->g.n, dev = c0->device;
    std::vector<srm_ctx *> lev((size_t)depth, nullptr);
    lev[0] = c0;
    int rc = SRM_OK;
    float ms_total = 0;
    SrmCtl h;
    memset(&h, 0, sizeof(h));
    auto cleanup = [&]() { for (int L = 1; L < depth; ++L) if (lev[L]) srm_destroy(lev[L]); };
#define MR(call) do { rc = (call); if (rc) { cleanup(); return rc; } } while (0)
#define MRC(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup(); return fail(SRM_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__)); } } while (0)
    MR(srm_set_density(c0, density, 0));
    for (int L = 1; L < depth; ++L) {
        MR(srm_create(&lev[L], n >> L, 0, n >> L, dev));
        MRC(cudaStreamSynchronize(lev[L - 1]->stream));   // the finer density is complete (different streams)
        srm_launch_density_scale(lev[L]->stream, lev[L - 1]->density, lev[L]->density, n >> L);
        MRC(cudaGetLastError());
        MR(density_ready(lev[L]));
    }
    for (int L = 0; L < depth; ++L) MR(srm_set_mask(lev[L], mask, 0));  // copies the first (n >> L)^2 bytes
    MR(srm_set_site_map(lev[depth - 1], voronoi, 0));
    MRC(cudaMemcpy(&h, lev[depth - 1]->ctl, sizeof(h), cudaMemcpyDeviceToHost));
    for (int L = depth - 1; L >= 0; --L) {
        srm_ctx *c = lev[L];
        MR(carry_ctl(c, h, L));
        MRC(cudaEventRecord(c->ev0, c->stream));
        // do { ... } while (gcvtIterations < maxIter): every level runs at least one iteration
        MR(srm_iterate(c, std::max(1, max_iter - h.it), 1));
        MRC(cudaEventRecord(c->ev1, c->stream));
        MR(fetch_ctl(c, &h));
        float ms = 0;
        MRC(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        ms_total += ms;
        if (L) {  // pbaCVDZoomIn: same ids, doubled coordinates, into the buffer the next level reads first
            srm_ctx *f = lev[L - 1];
            MR(alloc_sites(f, h.K));
            MR(reset_ctl(f, h.K));
            srm_launch_zoom_sites(f->stream, c->sites[current_buffer(c)], f->sites[h.it & 1], h.K);
            MRC(cudaGetLastError());
            MR(init_step_buffers(f, h.it & 1));   // the finer level continues the iteration count: parity h.it & 1
            f->has_sites = true;
        }
    }
    MRC(cudaEventRecord(c0->ev0, c0->stream));
    MR(srm_label(c0));  // final pba2DCompute (gcvt.cu:1149)
    MRC(cudaEventRecord(c0->ev1, c0->stream));
    MRC(cudaEventSynchronize(c0->ev1));
    float ms = 0;
    MRC(cudaEventElapsedTime(&ms, c0->ev0, c0->ev1));
    fill_stats(h, stats, ms_total + ms);
    MR(srm_get_labels(c0, voronoi, 0));
    cleanup();
#undef MR
#undef MRC
    return SRM_OK;
}

extern "C" int srm_gcvt(short *voronoi, const float *density, const unsigned char *mask, int n, int depth, int max_iter,
                        srm_stats *stats) {
    if (!voronoi || !density) return fail(SRM_ERR_ARG, "srm_gcvt: null argument");
    if (!valid_n(n)) return fail(SRM_ERR_ARG, "srm_gcvt: n=%d unsupported", n);
    if (depth < 1) depth = 1;
    for (int i = 0; i < depth; ++i) if ((n >> i) < 256) { depth = i; break; }  // gcvt.cu:1091
    int dev = 0;
    cudaGetDevice(&dev);
    const bool trace = getenv("SRM_TRACE") != nullptr;
    auto now = []() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    double t0 = now(), t1;
#define TR(what) do { if (trace) { t1 = now(); fprintf(stderr, "[srm_gcvt] %-12s %8.2f ms\n", what, t1 - t0); t0 = t1; } } while (0)
    // The reference allocates and frees ~46 N bytes of device memory on every call (gcvt.cu:840-903); here the
    // context of the last call is kept (per process, one at a time) and reused when n and the device match,
    // so repeated calls pay only for the transfers and the loop.  srm_release_cache() frees it.
    std::lock_guard<std::mutex> lock(g_cache_mu);
    if (g_cache && (g_cache_n != n || g_cache_dev != dev)) { srm_destroy(g_cache); g_cache = nullptr; }
    int rc = SRM_OK;
    if (!g_cache) {
        rc = srm_create(&g_cache, n, 0, n, dev);
        if (rc) { g_cache = nullptr; return rc; }
        g_cache_n = n; g_cache_dev = dev;
    }
    srm_ctx *c = g_cache;
    TR("create");
    if (depth > 1) {
        rc = gcvt_multires(c, voronoi, density, mask, depth, max_iter, stats);
        TR("multires");
        if (rc) { srm_destroy(g_cache); g_cache = nullptr; }
        return rc;
    }
    {   // the two sparse inputs are scanned on host threads while the density goes up through the staging pipeline
        std::vector<int> mask_px, sites;
        std::thread scan([&]() {
            if (mask) srm_scan_mask(mask, n, mask_px);
            srm_scan_site_map((const int *)voronoi, (size_t)n * n, sites);
        });
        rc = srm_set_density(c, density, 0);
        scan.join();
        TR("density+scans");
        if (!rc) { if (mask) rc = set_mask_pixels(c, mask_px); else c->has_mask = false; }
        TR("mask");
        if (!rc) rc = srm_set_sites(c, sites.data(), (int)sites.size(), 0);
        TR("sites");
    }
    if (!rc) rc = srm_run(c, max_iter, 1, stats);
    TR("run");
    if (!rc) rc = srm_get_labels(c, voronoi, 0);
    TR("get_labels");
    if (rc) { srm_destroy(g_cache); g_cache = nullptr; }  // do not keep a context in an unknown state
#undef TR
    return rc;
}

extern "C" int srm_discretize(const double *points, const double *weight, int num_point, const int *triangle,
                              int num_tri, float *density, double scale, int n) {
    if (!points || !weight || (!triangle && num_tri > 0) || !density || num_point <= 0 || num_tri < 0)
        return fail(SRM_ERR_ARG, "srm_discretize: bad argument");
    if (n < 16 || (n % 16)) return fail(SRM_ERR_ARG, "srm_discretize: n=%d must be a multiple of 16", n);
    if (!(scale > 0)) return fail(SRM_ERR_ARG, "srm_discretize: scale must be > 0");
    for (int i = 0; i < 3 * num_tri; ++i)
        if (triangle[i] < 0 || triangle[i] >= num_point) return fail(SRM_ERR_ARG, "srm_discretize: vertex index out of range");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(SRM_ERR_CUDA, "srm_discretize: no CUDA device (libsrm has no CPU fallback)");
    double *dp = nullptr, *dw = nullptr;
    int *dt = nullptr;
    float *dd = nullptr;
    int rc = SRM_OK;
    cudaError_t e;
#define CKR(call) do { e = (call); if (e != cudaSuccess) { rc = fail(SRM_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e)); goto out; } } while (0)
    CKR(cudaMalloc(&dp, sizeof(double) * 2 * num_point));
    CKR(cudaMalloc(&dw, sizeof(double) * num_point));
    CKR(cudaMalloc(&dt, sizeof(int) * 3 * (size_t)(num_tri > 0 ? num_tri : 1)));
    CKR(cudaMalloc(&dd, sizeof(float) * (size_t)n * n));
    CKR(cudaMemcpy(dp, points, sizeof(double) * 2 * num_point, cudaMemcpyHostToDevice));
    CKR(cudaMemcpy(dw, weight, sizeof(double) * num_point, cudaMemcpyHostToDevice));
    if (num_tri > 0) CKR(cudaMemcpy(dt, triangle, sizeof(int) * 3 * (size_t)num_tri, cudaMemcpyHostToDevice));
    CKR(srm_raster(nullptr, dp, dw, num_point, dt, num_tri, dd, scale, n));
    CKR(cudaMemcpy(density, dd, sizeof(float) * (size_t)n * n, cudaMemcpyDeviceToHost));
#undef CKR
out:
    cudaFree(dp); cudaFree(dw); cudaFree(dt); cudaFree(dd);
    return rc;
}

// ------------------------------------------------------------------ locate + lift (recover.h:63-153)

namespace {
struct DevBufs {   // frees what it allocated (and the locator built on it) when it goes out of scope
    std::vector<void *> p;
    SrmLocator *loc = nullptr;
    ~DevBufs() { srm_locator_free(loc); for (void *q : p) cudaFree(q); }
    template <typename T> cudaError_t up(T **d, const T *h, size_t count) {
        cudaError_t e = cudaMalloc((void **)d, sizeof(T) * (count ? count : 1));
        if (e != cudaSuccess) return e;
        p.push_back(*d);
        return (h && count) ? cudaMemcpy(*d, h, sizeof(T) * count, cudaMemcpyHostToDevice) : cudaSuccess;
    }
};
}  // namespace

static int check_faces(const char *who, const int *faces, int num_faces, int num_vertices) {
    for (int i = 0; i < 3 * num_faces; ++i)
        if (faces[i] < 0 || faces[i] >= num_vertices) return fail(SRM_ERR_ARG, "%s: vertex index out of range", who);
    return SRM_OK;
}

extern "C" int srm_locate(const double *mesh_xy, int num_vertices, const int *faces, int num_faces, const double *query_xy,
                          int num_query, int *face_out, double *w_out) {
    if (!mesh_xy || num_vertices <= 0 || (!faces && num_faces > 0) || num_faces < 0 || (!query_xy && num_query > 0) ||
        num_query < 0 || !face_out)
        return fail(SRM_ERR_ARG, "srm_locate: bad argument");
    int rc = check_faces("srm_locate", faces, num_faces, num_vertices);
    if (rc) return rc;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(SRM_ERR_CUDA, "srm_locate: no CUDA device (libsrm has no CPU fallback)");
    DevBufs B;
    double *dp = nullptr, *dq = nullptr, *dw = nullptr;
    int *dt = nullptr, *df = nullptr;
    SrmLocator *L = nullptr;
    CK(B.up(&dp, mesh_xy, 2 * (size_t)num_vertices));
    CK(B.up(&dt, faces, 3 * (size_t)num_faces));
    CK(B.up(&dq, query_xy, 2 * (size_t)num_query));
    CK(B.up(&df, (const int *)nullptr, (size_t)num_query));
    CK(B.up(&dw, (const double *)nullptr, 3 * (size_t)num_query));
    CK(srm_locator_build(nullptr, mesh_xy, num_vertices, dp, dt, num_faces, &L));
    B.loc = L;
    cudaError_t e = srm_locator_query(nullptr, L, dp, dt, dq, nullptr, num_query, df, dw);
    if (e == cudaSuccess && num_query) e = cudaMemcpy(face_out, df, sizeof(int) * num_query, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && num_query && w_out) e = cudaMemcpy(w_out, dw, sizeof(double) * 3 * num_query, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return fail(SRM_ERR_CUDA, "srm_locate: %s", cudaGetErrorString(e));
    return SRM_OK;
}

extern "C" int srm_recover(const double *mesh_xy, const double *mesh_xyz, int num_vertices, const int *faces, int num_faces,
                           const double *points_xy, int num_points, const int *cpoint_vertex, int num_cpoints,
                           const int *cdt_tri, int num_cdt_tri, double *vertices_xyz, unsigned char *tri_keep,
                           int *num_kept) {
    if (!mesh_xy || !mesh_xyz || num_vertices <= 0 || (!faces && num_faces > 0) || num_faces < 0 || !points_xy ||
        num_points < 0 || num_cpoints < 0 || num_cpoints > num_points || (!cpoint_vertex && num_cpoints > 0) ||
        (!cdt_tri && num_cdt_tri > 0) || num_cdt_tri < 0 || !vertices_xyz || (!tri_keep && num_cdt_tri > 0))
        return fail(SRM_ERR_ARG, "srm_recover: bad argument");
    int rc = check_faces("srm_recover", faces, num_faces, num_vertices);
    if (rc) return rc;
    for (int i = 0; i < 3 * num_cdt_tri; ++i)
        if (cdt_tri[i] < 0 || cdt_tri[i] >= num_points) return fail(SRM_ERR_ARG, "srm_recover: CDT vertex index out of range");
    for (int i = 0; i < num_cpoints; ++i)
        if (cpoint_vertex[i] < 0 || cpoint_vertex[i] >= num_vertices) return fail(SRM_ERR_ARG, "srm_recover: constraint vertex out of range");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(SRM_ERR_CUDA, "srm_recover: no CUDA device (libsrm has no CPU fallback)");
    const int nfree = num_points - num_cpoints;
    DevBufs B;
    double *dp = nullptr, *d3 = nullptr, *dq = nullptr, *dw = nullptr, *dout = nullptr;
    int *dt = nullptr, *df = nullptr, *dc = nullptr, *dfc = nullptr;
    SrmLocator *L = nullptr;
    CK(B.up(&dp, mesh_xy, 2 * (size_t)num_vertices));
    CK(B.up(&d3, mesh_xyz, 3 * (size_t)num_vertices));
    CK(B.up(&dt, faces, 3 * (size_t)num_faces));
    CK(B.up(&dq, points_xy, 2 * (size_t)num_points));
    CK(B.up(&dc, cdt_tri, 3 * (size_t)num_cdt_tri));
    CK(B.up(&df, (const int *)nullptr, (size_t)nfree));
    CK(B.up(&dw, (const double *)nullptr, 3 * (size_t)nfree));
    CK(B.up(&dout, (const double *)nullptr, 3 * (size_t)nfree));
    CK(B.up(&dfc, (const int *)nullptr, (size_t)num_cdt_tri));
    CK(srm_locator_build(nullptr, mesh_xy, num_vertices, dp, dt, num_faces, &L));
    B.loc = L;
    // sites: locate + barycentric lift (recover.h:94-109); CDT triangles: keep iff the centroid lies in a face (:115-142)
    cudaError_t e = srm_locator_query(nullptr, L, dp, dt, dq, nullptr, nfree, df, dw);
    if (e == cudaSuccess) e = srm_launch_lift(nullptr, dt, d3, df, dw, nfree, dout);
    if (e == cudaSuccess) e = srm_locator_query(nullptr, L, dp, dt, dq, dc, num_cdt_tri, dfc, nullptr);
    std::vector<int> hf((size_t)nfree), hfc((size_t)num_cdt_tri);
    if (e == cudaSuccess && nfree) e = cudaMemcpy(vertices_xyz, dout, sizeof(double) * 3 * nfree, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && nfree) e = cudaMemcpy(hf.data(), df, sizeof(int) * nfree, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && num_cdt_tri) e = cudaMemcpy(hfc.data(), dfc, sizeof(int) * num_cdt_tri, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return fail(SRM_ERR_CUDA, "srm_recover: %s", cudaGetErrorString(e));
    // A site that lies in no face: the reference ignores locate's return value and lifts with the PREVIOUS site's
    // face and weights (f_loc is not reset, recover.h:92-96), i.e. repeats the previous vertex.
    for (int i = 0; i < nfree; ++i)
        if (hf[i] < 0) {
            if (i == 0) return fail(SRM_ERR_ARG, "srm_recover: the first site lies in no face of the mesh");
            for (int d = 0; d < 3; ++d) vertices_xyz[3 * i + d] = vertices_xyz[3 * (i - 1) + d];
        }
    for (int i = 0; i < num_cpoints; ++i)   // constraint points keep their mesh vertex (recover.h:111-113)
        for (int d = 0; d < 3; ++d) vertices_xyz[3 * (size_t)(nfree + i) + d] = mesh_xyz[3 * (size_t)cpoint_vertex[i] + d];
    int kept = 0;
    for (int i = 0; i < num_cdt_tri; ++i) { tri_keep[i] = hfc[i] >= 0; kept += tri_keep[i]; }
    if (num_kept) *num_kept = kept;
    return SRM_OK;
}

// putConstrains + randomPoints (gcvt.h:58-122).  Host code in the reference too; the RNG is the
// never-seeded KISS generator, which degenerates to the LCG j <- 69069 j + 1234567 on 64-bit
// unsigned long.  Attempts are capped so that an unsatisfiable request fails
// instead of spinning forever like the reference would.
extern "C" int srm_seed(short *voronoi, const float *density, const unsigned char *mask, int num, int n,
                        unsigned long long *rng_state) {
    if (!voronoi || !density || n <= 0 || num < 0) return fail(SRM_ERR_ARG, "srm_seed: bad argument");
    const size_t N = (size_t)n * n;
    for (int y = 0; y < n; ++y)
        for (int x = 0; x < n; ++x) {
            const size_t i = (size_t)y * n + x;
            const bool m = mask && mask[i];
            voronoi[2 * i] = m ? (short)x : (short)SRM_MARKER;
            voronoi[2 * i + 1] = m ? (short)y : (short)SRM_MARKER;
        }
    double mx = 0, avg = 0, cnt = 0;
    for (size_t i = 0; i < N; ++i) {
        if (density[i] > mx) mx = density[i];
        if (density[i] != 0) { cnt += 1; avg += density[i]; }
    }
    mx = std::min(mx, avg / cnt * 100.);
    unsigned long long j = rng_state ? *rng_state : 0ull;
    auto next = [&j]() { j = 69069ull * j + 1234567ull; return (double)j / 18446744073709551616.0; };
    const long long cap = 1000ll * (long long)num + 100000000ll;
    long long attempts = 0;
    for (int k = 0; k < num; ++k) {
        for (;;) {
            if (attempts++ >= cap) return fail(SRM_ERR_SEED, "srm_seed: placed %d of %d sites in %lld attempts", k, num, cap);
            const int x = (int)(next() * n), y = (int)(next() * n);
            const double z = next() * mx;
            if (x >= n || y >= n) continue;
            const size_t i = (size_t)y * n + x;
            if (voronoi[2 * i] == SRM_MARKER && (double)density[i] > z) {
                voronoi[2 * i] = (short)x;
                voronoi[2 * i + 1] = (short)y;
                break;
            }
        }
    }
    if (rng_state) *rng_state = j;
    return SRM_OK;
}

extern "C" int srm_generate_mask(unsigned char *mask, const double *points_xy, int num_points, int n, double scale,
                                 double left, double lower) {
    if (!mask || (!points_xy && num_points > 0) || n <= 0) return fail(SRM_ERR_ARG, "srm_generate_mask: bad argument");
    memset(mask, 0, (size_t)n * n);  // the reference clears 8 bytes and relies on fresh pages (gcvt.h:152)
    for (int k = 0; k < num_points; ++k) {
        const int x = (int)((points_xy[2 * k] - left) / scale), y = (int)((points_xy[2 * k + 1] - lower) / scale);
        if (x < 0 || y < 0 || x >= n || y >= n) return fail(SRM_ERR_ARG, "srm_generate_mask: point %d outside the grid", k);
        mask[(size_t)y * n + x] = 1;
    }
    return SRM_OK;
}
