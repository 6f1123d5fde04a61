// srm_envelope.cuh — device building blocks shared by the fused band kernel (srm_band.cu) and the
// robust row path (srm_label.cu): exact lower envelope of the row parabolas with integer breakpoints.
//
// For candidates p < q (columns) with H = x^2 + (c - Y)^2, p beats q (ties included: smallest x wins,
// reference kernelColor gcvt.cu:449-466) exactly for X <= B(p,q) = floor((H_q - H_p) / (2 (q - p))).
// A stack element stores S = B(previous, itself): it wins for X in (S, S_next].  Every comparison
// "B <= S" / "B >= S" is done by cross-multiplication; one integer division per pushed element.
// All intermediate values fit int32 for n <= 32768 (2 * 32767^2 < 2^31).
#pragma once
#include "srm_common.cuh"

struct EnvSmem {
    unsigned short *x;        // candidate column
    short *c;                 // its site row c(x,Y)
    short *S;                 // element wins for X > S (within its merged group); -1 at the bottom
    unsigned short *sb, *se;  // per-segment live range [sb,se) inside x/c/S
};

// Sequential stack over candidates [beg,end) (sorted by x), in place; returns the new end.
__device__ __forceinline__ int env_lane_stack(const EnvSmem &s, int beg, int end, int Y, int n) {
    int top = beg;
    for (int i = beg; i < end; ++i) {
        const int xr = s.x[i];
        const short cr = s.c[i];
        const int gr = cr - Y, Hr = xr * xr + gr * gr;
        int num = 0, den = 1;
        while (top > beg) {
            const int xl = s.x[top - 1], gl = s.c[top - 1] - Y;
            num = Hr - (xl * xl + gl * gl);
            den = 2 * (xr - xl);
            if (num < ((int)s.S[top - 1] + 1) * den) --top; else break;  // B <= S_top: top wins nowhere
        }
        short S = -1;
        if (top > beg) {
            if (num >= (n - 1) * den) continue;  // B >= n-1: never wins inside the grid
            S = (short)(num / den);
        }
        s.x[top] = (unsigned short)xr; s.c[top] = cr; s.S[top] = S;
        ++top;
    }
    return top;
}

// Bridge two envelopes: segments [g0,gm) hold the left group, [gm,g1) the right group.  Pops dominated
// elements from the top of L and the bottom of R (the role of kernelMergeBands, gcvt.cu:293-410, over
// integer breakpoints and contiguous shared-memory segments).
__device__ __forceinline__ void env_merge(const EnvSmem &s, int g0, int gm, int g1, int Y, int n) {
    int sl = gm - 1;
    while (sl >= g0 && s.sb[sl] == s.se[sl]) --sl;
    int sr = gm;
    while (sr < g1 && s.sb[sr] == s.se[sr]) ++sr;
    if (sl < g0 || sr >= g1) return;
    int l = s.se[sl] - 1, r = s.sb[sr];
    for (;;) {
        const int xl = s.x[l], xr = s.x[r];
        const int gl = s.c[l] - Y, gr = s.c[r] - Y;
        const int num = (xr * xr + gr * gr) - (xl * xl + gl * gl);
        const int den = 2 * (xr - xl);
        if (num < ((int)s.S[l] + 1) * den) {  // B(l,r) <= S_l : l wins nowhere
            s.se[sl] = (unsigned short)l;
            if (l == s.sb[sl]) {
                do { --sl; } while (sl >= g0 && s.sb[sl] == s.se[sl]);
                if (sl < g0) { s.S[r] = -1; return; }
            }
            l = s.se[sl] - 1;
            continue;
        }
        bool rdead = num >= (n - 1) * den;  // r beats l only beyond the grid
        if (!rdead) {
            int r2 = -1;
            if (r + 1 < s.se[sr]) r2 = r + 1;
            else {
                int s2 = sr + 1;
                while (s2 < g1 && s.sb[s2] == s.se[s2]) ++s2;
                if (s2 < g1) r2 = s.sb[s2];
            }
            if (r2 >= 0 && num >= (int)s.S[r2] * den) rdead = true;  // B(l,r) >= S_r2 : r wins nowhere
        }
        if (rdead) {
            s.sb[sr] = (unsigned short)(r + 1);
            if (s.sb[sr] == s.se[sr]) {
                do { ++sr; } while (sr < g1 && s.sb[sr] == s.se[sr]);
                if (sr >= g1) return;
            }
            r = s.sb[sr];
            continue;
        }
        s.S[r] = (short)(num / den);  // num >= 0 here
        return;
    }
}

// Per-run accumulation of one row (one warp): run e = [start_e, start_{e+1}-1] contributes the fp64
// prefix differences of d, x*d (and x^2*d for the energy) to its site (role of kernelTotal_X +
// kernelScan_Y, gcvt.cu:591-732, and of kernelCalcEnergy, :788-802).  Returns the lane's energy part.
__device__ __forceinline__ double acc_row(const int2 *rr, int cnt, const double2 *__restrict__ p2,
                                          const double *__restrict__ pxx, const SrmHash &hash, int n, int Y,
                                          double *__restrict__ acc, int Kcap, int want_energy, int lane) {
    unsigned char *touched = reinterpret_cast<unsigned char *>(acc + 4 * (size_t)Kcap + 4);  // per-site "this rank contributed"
    double e_loc = 0;
    double2 carry = make_double2(0, 0);  // prefix at the end of the previous run
    double carryxx = 0;
    for (int e0 = 0; e0 < cnt; e0 += 32) {
        const int e = e0 + lane;
        const bool act = e < cnt;
        int2 v = act ? rr[e] : make_int2(0, 0);
        int b = n - 1;
        if (act && e + 1 < cnt) b = rr[e + 1].y - 1;
        double2 pb = act ? p2[(size_t)b * SRM_PFX_TILE] : make_double2(0, 0);   // p2 / pxx: row bases (srm_pfx_row)
        double xb = (act && want_energy) ? pxx[(size_t)b * SRM_PFX_TILE] : 0;
        double2 pa;
        pa.x = __shfl_up_sync(0xffffffffu, pb.x, 1);
        pa.y = __shfl_up_sync(0xffffffffu, pb.y, 1);
        double xa = __shfl_up_sync(0xffffffffu, xb, 1);
        if (lane == 0) { pa = carry; xa = carryxx; }
        const int lastl = min(31, cnt - e0 - 1);  // last active lane's prefix carries into the next batch
        carry.x = __shfl_sync(0xffffffffu, pb.x, lastl);
        carry.y = __shfl_sync(0xffffffffu, pb.y, lastl);
        carryxx = __shfl_sync(0xffffffffu, xb, lastl);
        if (act) {
            const double W = pb.x - pa.x, X = pb.y - pa.y;
            const int sx = srm_x(v.x), sy = srm_y(v.x);
            const int id = srm_hash_find(hash, (unsigned)v.x);   // every label is a live site: always found
            double *a = acc + 4 * (size_t)max(id, 0);
            if (id >= 0) {
                atomicAdd(a, W);
                atomicAdd(a + 1, X);
                atomicAdd(a + 2, (double)Y * W);
                touched[id] = 1;
            }
            if (want_energy) {
                const int dy = sy - Y;
                e_loc += (xb - xa) - 2.0 * (double)sx * X + (double)(sx * sx + dy * dy) * W;
            }
        }
    }
    return e_loc;
}
