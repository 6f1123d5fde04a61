/*
 * srm.h — C ABI of libsrm.so: the B200-native (sm_100a) discrete-CVT Lloyd engine that
 * replaces Surface-Remesher's GPU hot path.  Plain pointers and sizes only.
 *
 * Reference interfaces replaced (paths relative to the reference's source/):
 *   srm_gcvt        <- void gCVT(short*,float*,bool*,int,int,int)          gcvt.h:29, gcvt.cu:1087
 *   srm_discretize  <- void discretization_d(double*,double*,int,int*,int,
 *                                            float*,double,int)            discretization.h:66, discretization.cu:87
 *   srm_seed        <- putConstrains + randomPoints                        gcvt.h:76-122
 *   srm_generate_mask <- generateMask                                      gcvt.h:143-159
 *   srm_locate      <- locate(p, mesh, f_loc)                              recover.h:63-83
 *   srm_recover     <- recover(mesh_2d, mesh_3d, vertex_2d_to_3d, point,
 *                              triangle, c_points, resMesh)                recover.h:85-153
 *   srm_extract_sites <- the site scan of delaunayInput                    delaunay.h:46-57
 * The C++ shims with the reference's exact (mangled) signatures live in
 * surface-remesher_b200/csrc/srm_dropin.cpp (libsrm_dropin.so); INTEGRATION.md shows
 * how an unmodified main.cpp links against them.
 *
 * Conventions: index = y*n + x (TOID, gcvt.cu:38); a label / site is a little-endian
 * short2 (x,y) == one int32 `(uint16)x | (uint16)y << 16`; empty = MARKER = -32768 in
 * both halves (gcvt.cu:37).  Every function returns SRM_OK or an error code; the text
 * of the last error on the calling thread is srm_last_error().  There is no CPU
 * fallback: without a CUDA device every compute entry point fails with SRM_ERR_CUDA.
 */
#ifndef SRM_H
#define SRM_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRM_MARKER (-32768)

enum {
    SRM_OK = 0,
    SRM_ERR_ARG = 1,      /* bad size / null pointer / unsupported n */
    SRM_ERR_CUDA = 2,     /* CUDA runtime error (message in srm_last_error) */
    SRM_ERR_STATE = 3,    /* call order (e.g. iterate before density/sites are set) */
    SRM_ERR_SEED = 4      /* seeding could not place the requested number of sites */
};

typedef struct srm_ctx srm_ctx;

typedef struct srm_stats {
    int iterations;      /* Lloyd iterations executed (gcvtIterations, gcvt.cu:1086) */
    int num_sites;       /* sites alive at the end (sites merge, gcvt.cu:779-780) */
    int stopped;         /* 1 if the reference stopping rule fired (gcvt.cu:1136) */
    float omega;         /* over-relaxation factor at exit (gcvt.cu:1131) */
    float energy;        /* last energy computed (gcvt.cu:1117), as float like the reference */
    float ms_device;     /* device time of the loop + final labelling (CUDA events) */
} srm_stats;

const char *srm_last_error(void);
int srm_version(void);

/* ------------------------------------------------------------ one-shot drop-ins (host buffers) */

/* gCVT (gcvt.cu:1087-1156): voronoi in = seed map, out = label map of the final sites.
 * depth > 1 (clamped like gcvt.cu:1091 so that the coarsest level is >= 256) runs the reference's coarse-to-fine
 * loop (gcvt.cu:985-993, 1036-1051, 1110-1147): voronoi in = seed map of side n >> (depth-1) stored in the FIRST
 * (n >> (depth-1))^2 entries of the n^2 buffer (gcvt.cu:1101-1103), out = n^2 labels.  stats may be NULL. */
int srm_gcvt(short *voronoi, const float *density, const unsigned char *mask, int n, int depth, int max_iter,
             srm_stats *stats);

/* srm_gcvt keeps the device context of its last call for reuse (same n, same device); this frees it. */
int srm_release_cache(void);

/* discretization_d (discretization.cu:87-120): first triangle (index order) containing the
 * sample (x*scale, y*scale) gives the barycentric interpolation of the vertex weights; 0 if none. */
int srm_discretize(const double *points, const double *weight, int num_point, const int *triangle, int num_tri,
                   float *density, double scale, int n);

/* putConstrains + randomPoints (gcvt.h:76-122), LP64 semantics of the never-seeded RNG.
 * *rng_state in/out (reference: 0; may be NULL).  mask may be NULL. */
int srm_seed(short *voronoi, const float *density, const unsigned char *mask, int num, int n,
             unsigned long long *rng_state);

/* generateMask (gcvt.h:143-159): mask[int((p-l)/scale)] = 1 for every constraint point (x,y pairs). */
int srm_generate_mask(unsigned char *mask, const double *points_xy, int num_points, int n, double scale,
                      double left, double lower);

/* locate (recover.h:63-83) for num_query points: face_out[q] = the FIRST face (index order) of the 2-D mesh in which
 * all three barycentric weights of the point are >= 0, or -1; w_out (3 per query, may be NULL) = weights of the
 * face's vertices 0,1,2.  mesh_xy: 2 doubles per vertex; faces: 3 vertex indices per face. */
int srm_locate(const double *mesh_xy, int num_vertices, const int *faces, int num_faces, const double *query_xy,
               int num_query, int *face_out, double *w_out);

/* recover (recover.h:85-153): lift the CDT back to the surface.  mesh_xyz[v] = 3-D position of the surface vertex
 * that 2-D mesh vertex v maps to (vertex_2d_to_3d).  points_xy = the CDT input points: first the free sites, then
 * num_cpoints constraint points whose mesh vertices are cpoint_vertex[].  Output: one 3-D vertex per point (free
 * sites: barycentric lift in their face; constraint points: their mesh vertex) and tri_keep[t] = 1 iff the centroid
 * of CDT triangle t lies in a face of the 2-D mesh (the others are dropped, recover.h:134-136). */
int srm_recover(const double *mesh_xy, const double *mesh_xyz, int num_vertices, const int *faces, int num_faces,
                const double *points_xy, int num_points, const int *cpoint_vertex, int num_cpoints,
                const int *cdt_tri, int num_cdt_tri, double *vertices_xyz, unsigned char *tri_keep, int *num_kept);

/* ------------------------------------------------------------ handle API (device-resident state) */

/* A context owns rows [row0,row1) of an n x n grid on `device` (row0 = 0, row1 = n for a whole
 * grid; row bands for multi-GPU sharding).  n: multiple of 256, 256..32768;
 * row0,row1: multiples of 64.  Site list, site bitmap, density and mask are replicated per band. */
int srm_create(srm_ctx **out, int n, int row0, int row1, int device);
int srm_destroy(srm_ctx *ctx);
/* Run all work of this context on an existing CUDA stream (cudaStream_t passed as void*). */
int srm_set_stream(srm_ctx *ctx, void *cuda_stream);
int srm_synchronize(srm_ctx *ctx);

/* Inputs: full-grid arrays (n*n).  on_device != 0 means the pointer is device memory on ctx's device. */
int srm_set_density(srm_ctx *ctx, const float *density, int on_device);
int srm_set_mask(srm_ctx *ctx, const unsigned char *mask, int on_device); /* NULL = no constraints */
/* Host scans of the two sparse inputs (multi-threaded; no device work), for callers that shard them: the sites of
 * `pixels` seed-map pixels in row-major order (packed x | y << 16), and the non-zero pixels of rows [row0, row1) of a
 * 1 B/px mask.  *count = number found (may exceed capacity; at most capacity entries are written). */
int srm_scan_site_map_host(const short *site_map, size_t pixels, int *packed_out, int capacity, int *count);
int srm_scan_mask_host(const unsigned char *mask, int n, int row0, int row1, int *packed_out, int capacity, int *count);
/* The constraint pixels as a list (packed x | y << 16) instead of a 1 B/px mask. */
int srm_set_mask_pixels(srm_ctx *ctx, const int *packed_xy, int count);
/* Row bands without replicating the inputs: only the band's own rows of the density ((row1-row0)*n floats) ... */
int srm_set_density_band(srm_ctx *ctx, const float *band_rows, int on_device);
/* ... and the two full-grid bitmaps the replicated site update reads (which = 0: density != 0, 1: constraint pixels;
 * n*n/32 words each, row y starts at word y*n/32).  After srm_set_density_band every rank holds the density bits of its
 * own rows only: the caller exchanges the slices between the ranks (an all-gather; plumbing) before iterating. */
int srm_shared_bits(srm_ctx *ctx, int which, void **device_ptr, size_t *num_words);
/* Sites from a dense seed map (short2 per pixel) or from a packed list. */
int srm_set_site_map(srm_ctx *ctx, const short *site_map, int on_device);
int srm_set_sites(srm_ctx *ctx, const int *packed_xy, int num, int on_device);
int srm_get_sites(srm_ctx *ctx, int *packed_xy_host, int capacity, int *num_out);
int srm_set_omega(srm_ctx *ctx, float omega);
/* delaunayInput's site scan (delaunay.h:46-57): free sites (label == self, not a constraint pixel) in x-outer /
 * y-inner order as points (x*scale + l, y*scale + b).  mask_host may be NULL.  Replaces <- delaunay.h:30-57. */
int srm_extract_sites(srm_ctx *ctx, const unsigned char *mask_host, double scale, double l, double b,
                      double *points_xy, int capacity, int *num_out);
/* Options: "robust_only" = 1 labels every row with the worst-case-capacity row path instead of the fused band
 * kernel (same results; used by the tests to pin that path). */
int srm_set_option(srm_ctx *ctx, const char *name, int value);

/* Steps of one Lloyd iteration (gcvt.cu:1112-1123), all asynchronous on the context's stream. */
int srm_label(srm_ctx *ctx);                       /* pba2DCompute: exact labels of the current sites (run-length form) */
int srm_accumulate(srm_ctx *ctx, int want_energy); /* pbaCVDComputeCentroid (+pbaCVDCalcEnergy): per-site sums over this band */
int srm_update(srm_ctx *ctx);                      /* pbaCVDUpdateSites + the control block of gcvt.cu:1123-1140 */
int srm_label_accumulate(srm_ctx *ctx, int want_energy); /* srm_label + srm_accumulate fused in one kernel pass */
/* The centroid pass as a stand-alone streaming kernel over a DENSE label map (north_star's form; alternative to the fused
 * accumulation of srm_label_accumulate, NOT used by the loop): per-site sums of this band from labels_dev — device
 * memory, (row1-row0)*n short2, 16-byte aligned, e.g. the output of srm_get_labels / srm_label_jfa with on_device = 1;
 * NULL = the labels of the last srm_label, expanded into the context's own dense buffer first — and the density:
 * 8 B/px read once, warp-segmented reduction, one set of fp64 REDs per run.  Labels that are not live sites of the
 * context are ignored.  Adds to the accumulators srm_accumulate fills.
 * Replaces <- pbaCVDComputeCentroid gcvt.cu:1008-1023 (+ pbaCVDCalcEnergy gcvt.cu:1059-1083 with want_energy). */
int srm_accumulate_dense(srm_ctx *ctx, const short *labels_dev, int want_energy);
/* Device accumulators for an external all-reduce between srm_accumulate and srm_update:
 * 4*capacity+4 doubles: (W, X, Y, 0) per site, then (energy_sum,0,0,0). */
int srm_acc_buffer(srm_ctx *ctx, void **device_ptr, size_t *num_doubles);

/* Row bands on several GPUs: rank 0 obtains a 128-byte id, the caller distributes it, every rank binds its band
 * context; srm_iterate / srm_run then all-reduce (NCCL, sum, fp64) the accumulators between the band kernel and the
 * update, on the context's stream.  NCCL is loaded at run time (libnccl.so.2, or $SRM_NCCL_LIB). */
int srm_nccl_unique_id(char *id128);
int srm_nccl_init(srm_ctx *ctx, const char *id128, int rank, int world);

/* Alternative to the NCCL all-reduce: a FUSED all-reduce over peer memory.  Every rank describes its accumulator
 * pair and arrival flags in a 160-byte blob (CUDA IPC handles), the caller gathers the blobs of all ranks, every rank
 * connects (collective; after the sites are set).  From then on the update kernel itself waits for the peers'
 * arrival flags and pulls the per-site partial sums from their memory over NVLink; no collective kernel runs. */
int srm_p2p_info(srm_ctx *ctx, void *blob160);
int srm_p2p_connect(srm_ctx *ctx, const void *blobs /* world x 160 bytes */, int rank, int world);
/* Closes the mappings of the peers' buffers.  Every rank disconnects, the caller synchronises the ranks (barrier), and
 * only then are the contexts destroyed: a mapping must not outlive the peer's buffer. */
int srm_p2p_disconnect(srm_ctx *ctx);

/* iters x (label, accumulate, update) with the reference's energy/omega schedule; stop_rule != 0
 * honours the reference stopping rule (checked on device; remaining iterations become no-ops). */
int srm_iterate(srm_ctx *ctx, int iters, int stop_rule);
/* Same loop with CUDA events between the stages (measurement only): stage_ms[6] receives the summed device
 * milliseconds of {site bitmap + carries, fused band kernel, robust row path, accumulator all-reduce (row bands),
 * update + control, whole iteration}. */
int srm_iterate_profiled(srm_ctx *ctx, int iters, int stop_rule, float *stage_ms);
/* Whole gCVT on resident inputs: loop + final labelling. */
int srm_run(srm_ctx *ctx, int max_iter, int stop_rule, srm_stats *stats);
int srm_get_state(srm_ctx *ctx, srm_stats *stats);
/* Measurement helper: runs produced by the last labelling and rows that took the robust path. */
int srm_debug_counts(srm_ctx *ctx, long long *total_runs, int *overflow_rows);
/* Kernels launched by this library since it was loaded (process-wide; memsets and copies are not counted). */
long long srm_launch_count(void);
/* Statistics counter of the band kernel, collected while the option "dbg_stats" is 1: which = 0 max / 1 sum of the
 * band-list length, 2 bands, 6 warps that took the staging-overflow fallback of Phase A. */
int srm_debug_get(srm_ctx *ctx, int which, long long *value);

/* Option "band_order" (srm_set_option; default: SRM_BAND_ORDER in the environment, else the compiled default): the CTAs
 * of the band kernel take the 8-row bands by decreasing cost — runs per band of an earlier iteration, rebuilt on the
 * device every 10th iteration — instead of in row order, so that the last wave of CTAs is made of the cheap bands.
 * Results do not depend on it.  This returns the order (perm_out[i] = band of CTA i) and the cost per band of the last
 * labelling; *num_bands = rows / 8; nothing is written when capacity < *num_bands. */
int srm_debug_band_order(srm_ctx *ctx, int *perm_out, int *cost_out, int capacity, int *num_bands);

/* Host side of the boundary: worker threads (1..16, 0 = default: SRM_HOST_THREADS or half the hardware threads, at most
 * 8) of the pageable <-> device copy pipeline and of the host scans of the seed map and the mask; staging chunk size in
 * KB (256..65536, 0 = unchanged, default 4096). */
int srm_host_config(int threads, int chunk_kb);

/* Measurement / A-B tests: process-wide choice between two builds of a streaming kernel.  which = "expand" (runs ->
 * dense labels; 0 / 1 / 2), "prefix" (fp64 prefix sums; 0 / 1) or "centroid" (srm_accumulate_dense; 0 / 1); < 0 =
 * environment (SRM_EXPAND_V, SRM_PREFIX_V, SRM_CENTROID_V) / compiled default (the highest number of each). */
int srm_set_variant(const char *which, int value);

/* Measurement: device milliseconds per launch of a streaming kernel on the context's resident data (`reps` launches
 * between two CUDA events after one untimed launch).  which = "prefix" (density set), "expand" (after srm_label),
 * "centroid" / "centroid_energy" (srm_accumulate_dense on the context's own dense labels, after srm_label; the
 * accumulators are cleared afterwards). */
int srm_time_kernel(srm_ctx *ctx, const char *which, int reps, float *ms_per_launch);

/* Dense labels of this band (rows row0..row1): expands the run-length labels of the last srm_label.
 * out: (row1-row0)*n short2. */
int srm_get_labels(srm_ctx *ctx, short *out, int on_device);
/* Alternative labelling (north_star kernel family, NOT the reference's algorithm): jump flooding with
 * the given step schedule on the current sites, whole-grid contexts only; result as dense labels. */
int srm_label_jfa(srm_ctx *ctx, const int *steps, int nsteps, short *out, int on_device);
/* Measurement: runs the same schedule with CUDA events between the launches.  mode 1 = fused shared-memory tile kernel
 * for runs of small steps (sum <= 15) + vectorised far passes (the default of srm_label_jfa, option "jfa_mode"),
 * mode 0 = one plain kernel per pass.  ms[i] = device milliseconds of launch i, *nlaunch = launches issued. */
int srm_label_jfa_timed(srm_ctx *ctx, const int *steps, int nsteps, int mode, float *ms, int cap, int *nlaunch);

#ifdef __cplusplus
}
#endif
#endif /* SRM_H */
