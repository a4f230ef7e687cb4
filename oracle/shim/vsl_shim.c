/* oracle/shim/vsl_shim.c — TEST INFRASTRUCTURE ONLY.  See mkl_vsl.h. */
#include "../../include/rsvd_b200_rng.h"
#include "mkl_vsl.h"
#include <stddef.h>

static unsigned long long g_seed = 777ULL; /* the reference's unused `#define SEED 777` */

void oracle_set_seed(unsigned long long seed) { g_seed = seed; }
unsigned long long oracle_get_seed(void) { return g_seed; }

int vslNewStream(VSLStreamStatePtr *stream, int brng, unsigned int seed) {
    (void)brng; (void)seed; *stream = NULL; return 0;
}
int vslDeleteStream(VSLStreamStatePtr *stream) { (void)stream; return 0; }

int vsRngGaussian(int method, VSLStreamStatePtr stream, int n, float *r, float a, float sigma) {
    (void)method; (void)stream;
    long long N = n;
    #pragma omp parallel for schedule(static)
    for (long long b = 0; b < (N + 3) / 4; ++b) {
        float z[4];
        rsvd_normal4(g_seed, (uint64_t)b, z);
        for (int j = 0; j < 4; ++j) {
            long long i = 4 * b + j;
            if (i < N) r[i] = fmaf(sigma, z[j], a);
        }
    }
    return 0;
}

/* standalone entry for tests/oracle twins: fill out[i] = normal(seed, first + i) as double */
void oracle_fill_normal(unsigned long long seed, unsigned long long first, long long count, double *out) {
    #pragma omp parallel for schedule(static)
    for (long long i = 0; i < count; ++i) out[i] = (double)rsvd_normal_at(seed, first + (uint64_t)i);
}
void oracle_philox4x32_10(unsigned int ctr[4], unsigned int k0, unsigned int k1) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    rsvd_philox4x32_10(c, k0, k1);
    for (int i = 0; i < 4; ++i) ctr[i] = c[i];
}
