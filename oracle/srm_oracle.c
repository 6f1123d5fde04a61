/*
 * srm_oracle.c — CPU restatement of Surface-Remesher's discrete-CVT Lloyd path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is product code: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker / CPU baseline.  The product path
 * (surface-remesher_b200/csrc) never links, loads or calls it.
 *
 * Every function restates (does not copy) the semantics of the reference file:line
 * it cites; paths are relative to /root/reference/source/.  The labelling rule is
 * the one SURVEY.md §8(a)-L / Appendix A2 derives from gcvt.cu:77-216,421-479.
 *
 * Parity pin: tests/golden/ref_*.npz hold label maps, site maps and densities
 * produced by the UNMODIFIED reference CUDA (oracle/_ref/libsrm_ref.so, built from
 * /root/reference/source/{gcvt,discretization}.cu by oracle/Makefile) on a B200;
 * tests/test_oracle_golden.py checks this file against them.
 *
 * Build: make -C oracle   (gcc -O3 -fopenmp -ffp-contract=off)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MARK (-32768)
#define IDX(x, y, n) ((size_t)(y) * (size_t)(n) + (size_t)(x))
#define TIE_BAND 64 /* n/m1 == 64 for every row of the schedule table, gcvt.cu:842-847 */

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------ seeding */

/* gcvt.h:106-122 putConstrains: mask pixel -> its own coordinates, else MARK. */
void orc_put_constraints(short *vor, const uint8_t *mask, int n) {
    for (int y = 0; y < n; ++y)
        for (int x = 0; x < n; ++x) {
            size_t i = IDX(x, y, n);
            if (mask && mask[i]) { vor[2 * i] = (short)x; vor[2 * i + 1] = (short)y; }
            else { vor[2 * i] = MARK; vor[2 * i + 1] = MARK; }
        }
}

/* gcvt.h:58-75.  randinit() is never called, so z=w=jsr=0 and rand_int()==CONG();
 * `unsigned long` is 64-bit here (LP64): j <- 69069*j + 1234567 (mod 2^64),
 * random() = (double)j / 2^64. */
static inline double lcg_next(uint64_t *j) {
    *j = 69069ull * (*j) + 1234567ull;
    return (double)(*j) / 18446744073709551616.0;
}

/* gcvt.h:76-104 randomPoints.  Returns the number of attempts made, or -1 if
 * max_attempts (>0) was exhausted before `num` sites were placed (the reference
 * would loop forever).  *state carries the LCG state in and out (reference: 0). */
long long orc_random_points(short *vor, const float *density, int num, int n,
                            uint64_t *state, long long max_attempts) {
    double mx = 0, avg = 0, cnt = 0;
    size_t N = (size_t)n * n;
    for (size_t i = 0; i < N; ++i) {
        if (density[i] > mx) mx = density[i];
        if (density[i] != 0) { cnt += 1; avg += density[i]; }
    }
    double lim = avg / cnt * 100.;
    if (lim < mx) mx = lim; /* std::min(mx, avg/cnt*100.) */
    long long attempts = 0;
    for (int i = 0; i < num; ++i) {
        int x, y;
        double z;
        for (;;) {
            if (max_attempts > 0 && attempts >= max_attempts) return -1;
            x = (int)(lcg_next(state) * n);
            y = (int)(lcg_next(state) * n);
            z = lcg_next(state) * mx;
            ++attempts;
            if (x >= n || y >= n) continue; /* (double)j rounding to 2^64: p = 2^-54 */
            size_t id = IDX(x, y, n);
            if (vor[2 * id] == MARK && (double)density[id] > z) break;
        }
        size_t id = IDX(x, y, n);
        vor[2 * id] = (short)x;
        vor[2 * id + 1] = (short)y;
    }
    return attempts;
}

/* ---------------------------------------------------------------- labelling */

/* Column candidate of Appendix A2: U = nearest site row <= Y, D = nearest > Y
 * (MARK when absent); nearer by |dy|, tie -> D iff same 64-row band as Y. */
static inline int choose_col(int U, int D, int Y) {
    if (U == MARK) return D;
    if (D == MARK) return U;
    int du = Y - U, dd = D - Y;
    if (du < dd) return U;
    if (dd < du) return D;
    return (D / TIE_BAND == Y / TIE_BAND) ? D : U;
}

/* Literal A2, O(n^3): for n <= 512.  Follows the semantics of kernelFloodDown/Up,
 * kernelPropagateInterband, kernelUpdateVertical (gcvt.cu:77-216) for c[x][Y] and
 * kernelProximatePoints..kernelColor (gcvt.cu:227-479) for the argmin over x. */
void orc_label_brute(const short *seeds, short *labels, int n) {
    short *c = (short *)malloc(sizeof(short) * (size_t)n * n);
#pragma omp parallel for schedule(static)
    for (int x = 0; x < n; ++x)
        for (int Y = 0; Y < n; ++Y) {
            int U = MARK, D = MARK;
            for (int y = Y; y >= 0; --y)
                if (seeds[2 * IDX(x, y, n)] != MARK) { U = y; break; }
            for (int y = Y + 1; y < n; ++y)
                if (seeds[2 * IDX(x, y, n)] != MARK) { D = y; break; }
            c[IDX(x, Y, n)] = (short)choose_col(U, D, Y);
        }
#pragma omp parallel for schedule(static)
    for (int Y = 0; Y < n; ++Y)
        for (int X = 0; X < n; ++X) {
            long long best = -1;
            int bx = MARK, by = MARK;
            for (int x = 0; x < n; ++x) {
                int cy = c[IDX(x, Y, n)];
                if (cy == MARK) continue;
                long long dx = x - X, dy = cy - Y, d = dx * dx + dy * dy;
                if (best < 0 || d < best) { best = d; bx = x; by = cy; } /* ties -> smallest x */
            }
            labels[2 * IDX(X, Y, n)] = (short)bx;
            labels[2 * IDX(X, Y, n) + 1] = (short)by;
        }
    free(c);
}

static inline long long floordiv(long long a, long long b) { /* b > 0 */
    long long q = a / b, r = a % b;
    return (r != 0 && r < 0) ? q - 1 : q;
}

/* Same result as orc_label_brute in O(N): column sweep, then per row the lower
 * envelope of the parabolas (X-x)^2 + (c_x-Y)^2 with integer breakpoints; for
 * x1 < x2, x1 wins (ties included) for X <= floor((H2-H1)/(2(x2-x1))), H = x^2+g^2. */
void orc_label_exact(const short *seeds, short *labels, int n) {
    short *c = (short *)malloc(sizeof(short) * (size_t)n * n);
#pragma omp parallel for schedule(static)
    for (int x = 0; x < n; ++x) {
        int U = MARK;
        /* down sweep stores U; up sweep resolves with D */
        for (int Y = 0; Y < n; ++Y) {
            if (seeds[2 * IDX(x, Y, n)] != MARK) U = Y;
            c[IDX(x, Y, n)] = (short)U;
        }
        int D = MARK;
        for (int Y = n - 1; Y >= 0; --Y) {
            int u = c[IDX(x, Y, n)];
            c[IDX(x, Y, n)] = (short)choose_col(u, D, Y);
            if (seeds[2 * IDX(x, Y, n)] != MARK) D = Y;
        }
    }
#pragma omp parallel
    {
        int *sx = (int *)malloc(sizeof(int) * n);
        long long *sH = (long long *)malloc(sizeof(long long) * n);
        long long *sS = (long long *)malloc(sizeof(long long) * n); /* element wins for X > sS */
#pragma omp for schedule(static)
        for (int Y = 0; Y < n; ++Y) {
            int top = 0;
            for (int x = 0; x < n; ++x) {
                int cy = c[IDX(x, Y, n)];
                if (cy == MARK) continue;
                long long g = cy - Y, H = (long long)x * x + g * g, B = 0;
                while (top > 0) {
                    B = floordiv(H - sH[top - 1], 2ll * (x - sx[top - 1]));
                    if (top > 1 && B <= sS[top - 1]) --top; else break;
                }
                if (top > 0 && B >= n - 1) continue; /* never wins inside the grid */
                sx[top] = x; sH[top] = H; sS[top] = (top > 0) ? B : -(1ll << 40);
                ++top;
            }
            if (top == 0) {
                for (int X = 0; X < n; ++X) { labels[2 * IDX(X, Y, n)] = MARK; labels[2 * IDX(X, Y, n) + 1] = MARK; }
                continue;
            }
            int e = 0;
            for (int X = 0; X < n; ++X) {
                while (e + 1 < top && sS[e + 1] < X) ++e;
                labels[2 * IDX(X, Y, n)] = (short)sx[e];
                labels[2 * IDX(X, Y, n) + 1] = c[IDX(sx[e], Y, n)];
            }
        }
        free(sx); free(sH); free(sS);
    }
    free(c);
}

/* Rule A2 for the rows [r0, r1) of an n x n grid, from a site LIST instead of a dense seed map (sizes where
 * the dense maps do not fit a test: 16384^2, 32768^2).  Column candidates come from the sorted site rows of
 * every column; the row pass is the stack of orc_label_exact.  labels: (r1-r0) * n short2. */
void orc_label_band(const short *sx, const short *sy, int K, int n, int r0, int r1, short *labels) {
    int *start = (int *)calloc((size_t)n + 1, sizeof(int));
    short *rows = (short *)malloc(sizeof(short) * (size_t)(K > 0 ? K : 1));
    for (int k = 0; k < K; ++k) start[sx[k] + 1]++;
    for (int x = 0; x < n; ++x) start[x + 1] += start[x];
    {   /* counting sort by column, then by row inside the column */
        int *fill = (int *)malloc(sizeof(int) * (size_t)n);
        memcpy(fill, start, sizeof(int) * (size_t)n);
        for (int k = 0; k < K; ++k) rows[fill[sx[k]]++] = sy[k];
        free(fill);
#pragma omp parallel for schedule(dynamic, 64)
        for (int x = 0; x < n; ++x) {
            short *a = rows + start[x];
            int m = start[x + 1] - start[x];
            for (int i = 1; i < m; ++i) { short v = a[i]; int j = i - 1; while (j >= 0 && a[j] > v) { a[j + 1] = a[j]; --j; } a[j + 1] = v; }
        }
    }
    const int nr = r1 - r0;
    short *c = (short *)malloc(sizeof(short) * (size_t)nr * n);
#pragma omp parallel for schedule(static)
    for (int x = 0; x < n; ++x) {
        const short *a = rows + start[x];
        const int m = start[x + 1] - start[x];
        int p = 0; /* a[p-1] = largest row <= Y, a[p] = smallest row > Y */
        while (p < m && a[p] <= r0) ++p;
        for (int Y = r0; Y < r1; ++Y) {
            while (p < m && a[p] <= Y) ++p;
            const int U = p > 0 ? a[p - 1] : MARK, D = p < m ? a[p] : MARK;
            c[(size_t)(Y - r0) * n + x] = (short)choose_col(U, D, Y);
        }
    }
#pragma omp parallel
    {
        int *ex = (int *)malloc(sizeof(int) * n);
        long long *eH = (long long *)malloc(sizeof(long long) * n);
        long long *eS = (long long *)malloc(sizeof(long long) * n);
#pragma omp for schedule(static)
        for (int Y = r0; Y < r1; ++Y) {
            const short *cr = c + (size_t)(Y - r0) * n;
            short *out = labels + 2 * (size_t)(Y - r0) * n;
            int top = 0;
            for (int x = 0; x < n; ++x) {
                int cy = cr[x];
                if (cy == MARK) continue;
                long long g = cy - Y, H = (long long)x * x + g * g, B = 0;
                while (top > 0) {
                    B = floordiv(H - eH[top - 1], 2ll * (x - ex[top - 1]));
                    if (top > 1 && B <= eS[top - 1]) --top; else break;
                }
                if (top > 0 && B >= n - 1) continue;
                ex[top] = x; eH[top] = H; eS[top] = (top > 0) ? B : -(1ll << 40);
                ++top;
            }
            int e = 0;
            for (int X = 0; X < n; ++X) {
                if (top == 0) { out[2 * X] = MARK; out[2 * X + 1] = MARK; continue; }
                while (e + 1 < top && eS[e + 1] < X) ++e;
                out[2 * X] = (short)ex[e];
                out[2 * X + 1] = cr[ex[e]];
            }
        }
        free(ex); free(eH); free(eS);
    }
    free(c); free(rows); free(start);
}

/* Jump flooding (north_star's kernel family; NOT in the reference).
 * Schedule: nsteps step sizes, each pass reads the 3x3 stencil at +-k from the
 * previous pass's buffer (ping-pong).  Candidate order is fixed (dy=-k,0,+k outer,
 * dx=-k,0,+k inner); the key (dist^2, x, y) makes the result order-independent. */
void orc_label_jfa(const short *seeds, short *labels, int n, const int *steps, int nsteps) {
    size_t N = (size_t)n * n;
    short *a = (short *)malloc(sizeof(short) * 2 * N), *b = (short *)malloc(sizeof(short) * 2 * N);
    memcpy(a, seeds, sizeof(short) * 2 * N);
    for (int s = 0; s < nsteps; ++s) {
        int k = steps[s];
#pragma omp parallel for schedule(static)
        for (int Y = 0; Y < n; ++Y)
            for (int X = 0; X < n; ++X) {
                long long best = -1;
                int bx = MARK, by = MARK;
                for (int j = -1; j <= 1; ++j)
                    for (int i = -1; i <= 1; ++i) {
                        int qx = X + i * k, qy = Y + j * k;
                        if (qx < 0 || qy < 0 || qx >= n || qy >= n) continue;
                        int sx = a[2 * IDX(qx, qy, n)], sy = a[2 * IDX(qx, qy, n) + 1];
                        if (sx == MARK) continue;
                        long long dx = sx - X, dy = sy - Y, d = dx * dx + dy * dy;
                        if (best < 0 || d < best || (d == best && (sx < bx || (sx == bx && sy < by)))) {
                            best = d; bx = sx; by = sy;
                        }
                    }
                b[2 * IDX(X, Y, n)] = (short)bx;
                b[2 * IDX(X, Y, n) + 1] = (short)by;
            }
        short *t = a; a = b; b = t;
    }
    memcpy(labels, a, sizeof(short) * 2 * N);
    free(a); free(b);
}

/* ------------------------------------------------- centroid / update / energy */

/* What kernelVoronoi1D + kernelTotal_X + kernelScan_Y (gcvt.cu:563-732) compute,
 * as direct fp64 sums: W,X,Y accumulated at the owning site's pixel index.
 * Image-sized outputs, zeroed here. */
void orc_centroid(const short *labels, const float *density, int n, double *W, double *X, double *Yv) {
    size_t N = (size_t)n * n;
    memset(W, 0, sizeof(double) * N); memset(X, 0, sizeof(double) * N); memset(Yv, 0, sizeof(double) * N);
    for (int y = 0; y < n; ++y)
        for (int x = 0; x < n; ++x) {
            size_t i = IDX(x, y, n);
            int sx = labels[2 * i], sy = labels[2 * i + 1];
            if (sx == MARK) continue;
            size_t s = IDX(sx, sy, n);
            double d = density[i];
            W[s] += d; X[s] += (double)x * d; Yv[s] += (double)y * d;
        }
}

/* kernelCalcEnergy + kernelReduce (gcvt.cu:788-832,1059-1083), exact fp64:
 * E = sum d*((sx-x)^2+(sy-y)^2) / n^2.  (The reference's fp32 reduction over-reads,
 * SURVEY §8(a) quirk 3; not reproduced.) */
double orc_energy(const short *labels, const float *density, int n) {
    double E = 0;
#pragma omp parallel for schedule(static) reduction(+ : E)
    for (int y = 0; y < n; ++y) {
        double e = 0;
        for (int x = 0; x < n; ++x) {
            size_t i = IDX(x, y, n);
            int sx = labels[2 * i], sy = labels[2 * i + 1];
            if (sx == MARK) continue;
            long long dx = sx - x, dy = sy - y;
            e += (double)density[i] * (double)(dx * dx + dy * dy);
        }
        E += e;
    }
    return E / ((double)n * (double)n);
}

static inline short f2s_rz(float v) { /* cvt.rzi.s16.f32: truncate, saturate, NaN -> 0 */
    if (v != v) return 0;
    if (v >= 32767.f) return 32767;
    if (v <= -32768.f) return -32768;
    return (short)v;
}

/* kernelFillShort + kernelUpdateSites (gcvt.cu:65-71,753-781).  nvcc contracts
 * tx + (_x-tx)*omega into one fma (checked in the SASS of oracle/_ref); fmaf mirrors it. */
void orc_update_sites(const short *labels, const double *W, const double *X, const double *Yv,
                      const float *density, const uint8_t *mask, int n, float omega, short *out) {
    size_t N = (size_t)n * n;
    for (size_t i = 0; i < 2 * N; ++i) out[i] = MARK;
    for (int ty = 0; ty < n; ++ty)
        for (int tx = 0; tx < n; ++tx) {
            size_t id = IDX(tx, ty, n);
            if (labels[2 * id] != tx || labels[2 * id + 1] != ty) continue;
            int rx = tx, ry = ty;
            if (!(mask && mask[id])) {
                float pX = (float)X[id], pY = (float)Yv[id], pW = (float)W[id];
                float _x = pX / pW, _y = pY / pW;
                int cx = f2s_rz(fmaf(_x - (float)tx, omega, (float)tx) + 0.5f);
                int cy = f2s_rz(fmaf(_y - (float)ty, omega, (float)ty) + 0.5f);
                cx = cx > n - 1 ? n - 1 : cx; cx = cx < 0 ? 0 : cx;
                cy = cy > n - 1 ? n - 1 : cy; cy = cy < 0 ? 0 : cy;
                if (density[IDX(cx, cy, n)] != 0) { rx = cx; ry = cy; }
            }
            size_t o = IDX(rx, ry, n);
            out[2 * o] = (short)rx; out[2 * o + 1] = (short)ry;
        }
}

/* One Lloyd iteration on a site map: label -> (energy) -> centroid -> update.
 * labels_out / energy_out may be NULL.  Scratch is allocated per call. */
void orc_lloyd_step(const short *seeds, const float *density, const uint8_t *mask, int n, float omega,
                    short *labels_out, short *seeds_out, double *energy_out) {
    size_t N = (size_t)n * n;
    short *lab = labels_out ? labels_out : (short *)malloc(sizeof(short) * 2 * N);
    double *W = (double *)malloc(sizeof(double) * 3 * N), *X = W + N, *Yv = X + N;
    orc_label_exact(seeds, lab, n);
    if (energy_out) *energy_out = orc_energy(lab, density, n);
    orc_centroid(lab, density, n, W, X, Yv);
    orc_update_sites(lab, W, X, Yv, density, mask, n, omega, seeds_out);
    free(W);
    if (!labels_out) free(lab);
}

/* gCVT driver, single level (gcvt.cu:1087-1156 with depth==1; Appendix A5).
 * vor: in = seed map, out = label map of the final sites.  Returns iterations run.
 * energies (optional, capacity max_iter/10+2) receives E at it = 0,10,20,...
 * stop_rule: 1 = reference stopping rule, 0 = fixed max_iter iterations (omega still
 * follows the reference schedule). */
int orc_gcvt(short *vor, const float *density, const uint8_t *mask, int n, int max_iter, int stop_rule,
             double *energies, float *omega_out) {
    size_t N = (size_t)n * n;
    short *cur = (short *)malloc(sizeof(short) * 2 * N), *nxt = (short *)malloc(sizeof(short) * 2 * N);
    memcpy(cur, vor, sizeof(short) * 2 * N);
    float Energy = 0, lastEnergy = 1e18f, diffEnergy, gradientEnergy, omega = 2.0f;
    int it = 0, ne = 0;
    do {
        double E;
        orc_lloyd_step(cur, density, mask, n, omega, NULL, nxt, (it % 10 == 0) ? &E : NULL);
        if (it % 10 == 0) { Energy = (float)E; if (energies) energies[ne++] = E; }
        short *t = cur; cur = nxt; nxt = t;
        ++it;
        if (it % 10 == 0) {
            diffEnergy = lastEnergy - Energy;
            gradientEnergy = (float)(diffEnergy / 10.0);
            double om = 1.0 + (double)diffEnergy;
            omega = (float)(om < 2.0 ? om : 2.0);
            if (stop_rule && gradientEnergy < 1e-5) break;
            lastEnergy = Energy;
        }
    } while (it < max_iter);
    orc_label_exact(cur, vor, n);
    if (omega_out) *omega_out = omega;
    free(cur); free(nxt);
    return it;
}

/* --------------------------------------------------------------- multires (a11) */

/* kernelDensityScaling (gcvt.cu:497-511): 2x2 box filter, float adds in the loop order
 * x outer / y inner, then /4.0 (a double division of a float, exact).  s = OUTPUT side. */
void orc_density_scale(const float *in, float *out, int s) {
    for (int ty = 0; ty < s; ++ty)
        for (int tx = 0; tx < s; ++tx) {
            float d = 0;
            for (int x = 2 * tx; x < 2 * tx + 2; ++x)
                for (int y = 2 * ty; y < 2 * ty + 2; ++y) d += in[IDX(x, y, 2 * s)];
            out[IDX(tx, ty, s)] = (float)((double)d / 4.0);
        }
}

/* pbaCVDZoomIn = kernelFillShort + kernelZoomIn (gcvt.cu:485-495,1036-1051): a site (x,y) of the
 * s-sided map becomes the site (2x,2y) of the 2s-sided map.  s = INPUT side. */
void orc_zoom_in(const short *in, short *out, int s) {
    size_t N2 = (size_t)4 * s * s;
    for (size_t i = 0; i < 2 * N2; ++i) out[i] = MARK;
    for (int y = 0; y < s; ++y)
        for (int x = 0; x < s; ++x) {
            size_t i = IDX(x, y, s);
            if (in[2 * i] == MARK) continue;
            size_t o = IDX(2 * x, 2 * y, 2 * s);
            out[2 * o] = (short)(in[2 * i] << 1); out[2 * o + 1] = (short)(in[2 * i + 1] << 1);
        }
}

/* gCVT driver with the coarse-to-fine loop (gcvt.cu:1087-1156).  depth is clamped so that the coarsest
 * level is >= 256 (:1091).  vor: in = seed map of side n >> (depth-1) in the FIRST (n >> (depth-1))^2
 * entries (:1101-1103), out = n^2 labels of the final sites.  Level L (0 = finest) runs on the L-times
 * box-filtered density; the constraint mask is indexed with the LEVEL's side on the full-resolution
 * buffer (kernelUpdateSites is handed constrainMask_d and size = pbaTexSize, :1032-1033), i.e. a coarse
 * level sees the first s^2 bytes of the mask as an s x s image.  The iteration counter, omega, Energy
 * and lastEnergy carry over between levels; the energy is scaled by 4^L (:1082); coarse levels stop at
 * gradient < 3e-1 (:1133), every level runs at least one iteration (do/while).
 * level_iters (optional, capacity depth) receives the iteration count at the end of each level, coarsest
 * first (the reference's switch_iter, :1144).  Returns iterations run, or -1 on a bad size. */
int orc_gcvt_multires(short *vor, const float *density, const uint8_t *mask, int n, int depth, int max_iter,
                      int stop_rule, int *level_iters, float *omega_out, float *energy_out) {
    if (depth < 1) depth = 1;
    for (int i = 0; i < depth; ++i) if ((n >> i) < 256) { depth = i; break; }
    if (depth < 1) return -1;
    size_t N = (size_t)n * n;
    float **dens = (float **)malloc(sizeof(float *) * depth);
    dens[0] = (float *)density;
    for (int i = 1; i < depth; ++i) {
        int s = n >> i;
        dens[i] = (float *)malloc(sizeof(float) * (size_t)s * s);
        orc_density_scale(dens[i - 1], dens[i], s);
    }
    short *cur = (short *)malloc(sizeof(short) * 2 * N), *nxt = (short *)malloc(sizeof(short) * 2 * N);
    int s = n >> (depth - 1);
    memcpy(cur, vor, sizeof(short) * 2 * (size_t)s * s);
    float Energy = 0, lastEnergy = 1e18f, diffEnergy, gradientEnergy, omega = 2.0f;
    int it = 0, nl = 0;
    for (int L = depth - 1; L >= 0; --L) {
        s = n >> L;
        do {
            double E;
            orc_lloyd_step(cur, dens[L], mask, s, omega, NULL, nxt, (it % 10 == 0) ? &E : NULL);
            if (it % 10 == 0) Energy = (float)E * powf(2.0f, (float)L * 2.0f);
            short *t = cur; cur = nxt; nxt = t;
            ++it;
            if (it % 10 == 0) {
                diffEnergy = lastEnergy - Energy;
                gradientEnergy = (float)(diffEnergy / 10.0);
                double om = 1.0 + (double)diffEnergy;
                omega = (float)(om < 2.0 ? om : 2.0);
                if (stop_rule && (double)gradientEnergy < (L ? 3e-1 : 1e-5)) break;
                lastEnergy = Energy;
            }
        } while (it < max_iter);
        if (level_iters) level_iters[nl++] = it;
        if (L) { orc_zoom_in(cur, nxt, s); short *t = cur; cur = nxt; nxt = t; }
    }
    orc_label_exact(cur, vor, n);
    if (omega_out) *omega_out = omega;
    if (energy_out) *energy_out = Energy;
    for (int i = 1; i < depth; ++i) free(dens[i]);
    free(dens); free(cur); free(nxt);
    return it;
}

/* --------------------------------------------------------------- rasteriser */

/* discretization.cu:32-85 (Appendix A6).  fp64; nvcc's default -fmad=true contracts
 * the reference's expressions, mirrored here with explicit fma() in the pattern seen
 * in the SASS of oracle/_ref (see DESIGN.md "rasteriser arithmetic"). */
void orc_rasterise(const double *pts, const double *wt, int num_point, const int *tri, int num_tri,
                   float *density, double scale, int n) {
    (void)num_point;
#pragma omp parallel for schedule(dynamic, 4)
    for (int ty = 0; ty < n; ++ty)
        for (int tx = 0; tx < n; ++tx) {
            float res = 0;
            for (int t = 0; t < num_tri; ++t) {
                int p1 = tri[3 * t], p2 = tri[3 * t + 1], p3 = tri[3 * t + 2];
                double x1 = pts[2 * p1], y1 = pts[2 * p1 + 1];
                double x2 = pts[2 * p2], y2 = pts[2 * p2 + 1];
                double x3 = pts[2 * p3], y3 = pts[2 * p3 + 1];
                double v0x = x2 - x1, v0y = y2 - y1, v1x = x3 - x1, v1y = y3 - y1;
                /* x = tx*scale is contracted into the subtraction: DFMA (tx, scale, -x1) */
                double v2x = fma((double)tx, scale, -x1), v2y = fma((double)ty, scale, -y1);
                double d00 = fma(v0x, v0x, v0y * v0y);
                double d01 = fma(v0x, v1x, v0y * v1y);
                double d11 = fma(v1x, v1x, v1y * v1y);
                double d20 = fma(v2x, v0x, v2y * v0y);
                double d21 = fma(v2x, v1x, v2y * v1y);
                double denom = fma(d00, d11, -(d01 * d01));
                if (denom == 0) continue;
                double w2 = fma(d11, d20, -(d01 * d21)) / denom; /* weight of p2 */
                double w3 = fma(d00, d21, -(d01 * d20)) / denom; /* weight of p3 */
                double w1 = 1.0 - w2 - w3;                        /* weight of p1 */
                if (w1 < 0 || w2 < 0 || w3 < 0) continue;
                res = (float)fma(w3, wt[p3], fma(w2, wt[p2], w1 * wt[p1]));
                break;
            }
            density[IDX(tx, ty, n)] = res;
        }
}

/* ------------------------------------------------------- locate + lift (f3) */

/* recover.h:30-53 `barycentric`, host arithmetic (no contraction: this file is built with -ffp-contract=off).
 * w[0],w[1],w[2] = weights of p1,p2,p3 (the out-parameter rotation cancels, Appendix A6).  Returns 0 for a
 * degenerate face (the reference sets all weights to -1, which the caller rejects). */
static inline int rec_bary(double x1, double y1, double x2, double y2, double x3, double y3, double x0, double y0,
                           double *w) {
    double v0x = x2 - x1, v0y = y2 - y1, v1x = x3 - x1, v1y = y3 - y1, v2x = x0 - x1, v2y = y0 - y1;
    double d00 = v0x * v0x + v0y * v0y, d01 = v0x * v1x + v0y * v1y, d11 = v1x * v1x + v1y * v1y;
    double d20 = v2x * v0x + v2y * v0y, d21 = v2x * v1x + v2y * v1y;
    double denom = d00 * d11 - d01 * d01;
    if (denom == 0) return 0;
    double w1 = (d11 * d20 - d01 * d21) / denom, w2 = (d00 * d21 - d01 * d20) / denom, w3 = 1.0 - w1 - w2;
    w[0] = w3; w[1] = w1; w[2] = w2;
    return 1;
}

/* recover.h:63-83 `locate`: the first face (index order) with all weights >= 0; -1 if none.  Brute force over
 * every face, like the reference.  w (3 doubles) is written only on success. */
int orc_locate_one(const double *pts, const int *tri, int T, double x, double y, double *w) {
    for (int t = 0; t < T; ++t) {
        int a = tri[3 * t], b = tri[3 * t + 1], c = tri[3 * t + 2];
        double ww[3];
        if (!rec_bary(pts[2 * a], pts[2 * a + 1], pts[2 * b], pts[2 * b + 1], pts[2 * c], pts[2 * c + 1], x, y, ww)) continue;
        if (ww[0] < 0 || ww[1] < 0 || ww[2] < 0) continue;
        if (w) { w[0] = ww[0]; w[1] = ww[1]; w[2] = ww[2]; }
        return t;
    }
    return -1;
}

void orc_locate(const double *pts, const int *tri, int T, const double *qxy, int Q, int *face, double *w) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int q = 0; q < Q; ++q) {
        double ww[3] = {0, 0, 0};
        face[q] = orc_locate_one(pts, tri, T, qxy[2 * q], qxy[2 * q + 1], ww);
        if (w) { w[3 * q] = ww[0]; w[3 * q + 1] = ww[1]; w[3 * q + 2] = ww[2]; }
    }
}

/* recover.h:85-153 `recover` on plain arrays.  pts3d[v] = 3-D position that 2-D mesh vertex v maps to.
 * points: nfree free sites followed by ncp constraint points (mesh vertices cpv[]).  A site in no face reuses
 * the previous site's face and weights (f_loc is not reset, :92-96).  Returns the number of kept triangles,
 * -1 if the first site lies in no face (the reference reads an uninitialised f_loc there). */
int orc_recover(const double *pts, const double *pts3d, const int *tri, int T, const double *pxy, int npoints,
                const int *cpv, int ncp, const int *cdt, int M, double *out, unsigned char *keep) {
    int nfree = npoints - ncp, f = -1;
    double w[3] = {0, 0, 0};
    for (int i = 0; i < nfree; ++i) {
        double ww[3];
        int g = orc_locate_one(pts, tri, T, pxy[2 * i], pxy[2 * i + 1], ww);
        if (g >= 0) { f = g; w[0] = ww[0]; w[1] = ww[1]; w[2] = ww[2]; }
        if (f < 0) return -1;
        for (int d = 0; d < 3; ++d) {
            double s = 0.0;
            for (int j = 0; j < 3; ++j) s += pts3d[3 * tri[3 * f + j] + d] * w[j];
            out[3 * i + d] = s;
        }
    }
    for (int i = 0; i < ncp; ++i)
        for (int d = 0; d < 3; ++d) out[3 * (size_t)(nfree + i) + d] = pts3d[3 * (size_t)cpv[i] + d];
    int kept = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : kept)
    for (int t = 0; t < M; ++t) {
        int a = cdt[3 * t], b = cdt[3 * t + 1], c = cdt[3 * t + 2];
        double cx = (pxy[2 * a] + pxy[2 * b] + pxy[2 * c]) / 3.0, cy = (pxy[2 * a + 1] + pxy[2 * b + 1] + pxy[2 * c + 1]) / 3.0;
        keep[t] = orc_locate_one(pts, tri, T, cx, cy, NULL) >= 0;
        kept += keep[t];
    }
    return kept;
}

/* ----------------------------------------------- CPU baseline (bench.py only) */

/* OpenMP Lloyd iteration used as BASELINE.md §4(b): separable exact labelling,
 * per-site fp64 sums keyed by a dense id map, reference update expression.
 * Same results as orc_lloyd_step; organised for throughput (row-parallel partial
 * sums) rather than clarity.  sites_x/sites_y: K sites in, updated in place;
 * returns the new K (sites merge).  scratch: caller-provided, see orc_fast_scratch_bytes. */
size_t orc_fast_scratch_bytes(int n, int K) {
    size_t N = (size_t)n * n;
    return sizeof(int) * N /*idmap*/ + sizeof(short) * 4 * N /*seed map + labels*/ +
           sizeof(double) * 3 * (size_t)K * (size_t)orc_num_threads() + 64;
}

int orc_fast_step(short *sx, short *sy, int K, const float *density, const uint8_t *mask, int n, float omega,
                  void *scratch, double *energy_out) {
    size_t N = (size_t)n * n;
    int *idmap = (int *)scratch;
    short *seeds = (short *)(idmap + N);
    short *lab = seeds + 2 * N;
    double *part = (double *)(((uintptr_t)(lab + 2 * N) + 63) & ~(uintptr_t)63);
    int T = orc_num_threads();
    /* scatter sites into a dense seed map */
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < N; ++i) { seeds[2 * i] = MARK; seeds[2 * i + 1] = MARK; }
    for (int k = 0; k < K; ++k) {
        size_t i = IDX(sx[k], sy[k], n);
        seeds[2 * i] = sx[k]; seeds[2 * i + 1] = sy[k]; idmap[i] = k;
    }
    orc_label_exact(seeds, lab, n);
    memset(part, 0, sizeof(double) * 3 * (size_t)K * T);
    double E = 0;
#pragma omp parallel reduction(+ : E)
    {
#ifdef _OPENMP
        int t = omp_get_thread_num();
#else
        int t = 0;
#endif
        double *P = part + 3 * (size_t)K * t;
#pragma omp for schedule(static)
        for (int y = 0; y < n; ++y)
            for (int x = 0; x < n; ++x) {
                size_t i = IDX(x, y, n);
                int lx = lab[2 * i], ly = lab[2 * i + 1];
                int k = idmap[IDX(lx, ly, n)];
                double d = density[i];
                P[3 * k] += d; P[3 * k + 1] += x * d; P[3 * k + 2] += y * d;
                long long dx = lx - x, dy = ly - y;
                E += d * (double)(dx * dx + dy * dy);
            }
    }
    if (energy_out) *energy_out = E / ((double)n * n);
    /* reduce partials, update, dedupe (first id to claim a pixel keeps it) */
    int newK = 0;
    for (int k = 0; k < K; ++k) {
        double W = 0, X = 0, Y = 0;
        for (int t = 0; t < T; ++t) { const double *P = part + 3 * (size_t)K * t + 3 * k; W += P[0]; X += P[1]; Y += P[2]; }
        int tx = sx[k], ty = sy[k], rx = tx, ry = ty;
        size_t id = IDX(tx, ty, n);
        if (!(mask && mask[id])) {
            float _x = (float)X / (float)W, _y = (float)Y / (float)W;
            int cx = f2s_rz(fmaf(_x - (float)tx, omega, (float)tx) + 0.5f);
            int cy = f2s_rz(fmaf(_y - (float)ty, omega, (float)ty) + 0.5f);
            cx = cx > n - 1 ? n - 1 : cx; cx = cx < 0 ? 0 : cx;
            cy = cy > n - 1 ? n - 1 : cy; cy = cy < 0 ? 0 : cy;
            if (density[IDX(cx, cy, n)] != 0) { rx = cx; ry = cy; }
        }
        sx[k] = (short)rx; sy[k] = (short)ry;
    }
    /* dedupe with idmap as the claim map (stale entries are never read as claims:
     * mark with -1-k) */
    for (int k = 0; k < K; ++k) idmap[IDX(sx[k], sy[k], n)] = -1;
    for (int k = 0; k < K; ++k) {
        size_t i = IDX(sx[k], sy[k], n);
        if (idmap[i] == -1) { idmap[i] = newK; sx[newK] = sx[k]; sy[newK] = sy[k]; ++newK; }
    }
    return newK;
}
