/* examples/gcvt_c_abi.c — the C ABI used from plain C: seed, run the Lloyd loop, read the
 * result.  Build:  gcc -Iinclude examples/gcvt_c_abi.c -Lsurface-remesher_b200 -lsrm -Wl,-rpath,$PWD/surface-remesher_b200
 * Needs a CUDA device at run time (libsrm has no CPU fallback: srm_gcvt then returns SRM_ERR_CUDA). */
#include <stdio.h>
#include <stdlib.h>
#include "srm.h"

int main(void) {
    const int n = 256, sites = 300;
    const size_t N = (size_t)n * n;
    short *vor = (short *)malloc(sizeof(short) * 2 * N);          /* seed map in, label map out (gcvt.h:133-138) */
    float *dens = (float *)malloc(sizeof(float) * N);
    unsigned char *mask = (unsigned char *)calloc(N, 1);
    if (!vor || !dens || !mask) return 2;
    for (size_t i = 0; i < N; ++i) dens[i] = 1.0f;
    if (srm_seed(vor, dens, mask, sites, n, NULL) != SRM_OK) { fprintf(stderr, "%s\n", srm_last_error()); return 1; }
    srm_stats st;
    int rc = srm_gcvt(vor, dens, mask, n, /*depth*/ 1, /*maxIter*/ 50, &st);
    if (rc != SRM_OK) { fprintf(stderr, "srm_gcvt: %d %s\n", rc, srm_last_error()); free(vor); free(dens); free(mask); return rc == SRM_ERR_CUDA ? 3 : 1; }
    printf("iterations %d sites %d omega %g energy %g device ms %g\n", st.iterations, st.num_sites, st.omega, st.energy, st.ms_device);
    free(vor); free(dens); free(mask);
    return 0;
}
