/* oracle/shim/mkl_vsl.h — TEST INFRASTRUCTURE ONLY.
 * Stand-in for MKL VSL: the reference draws Omega with vslNewStream(BRNG, time(NULL)) +
 * vsRngGaussian (matrix_vector_functions_intel_mkl.c:470-473), which is irreproducible by design.
 * The shim ignores the time seed and fills r[i] = normal(seed, i) from include/rsvd_b200_rng.h,
 * the same (seed, linear index) -> value map the device kernels use ("Omega imported" mode). */
#ifndef ORACLE_SHIM_MKL_VSL_H
#define ORACLE_SHIM_MKL_VSL_H
typedef void *VSLStreamStatePtr;
#define VSL_BRNG_MCG31 1
#define VSL_RNG_METHOD_GAUSSIAN_ICDF 1
int vslNewStream(VSLStreamStatePtr *stream, int brng, unsigned int seed);
int vsRngGaussian(int method, VSLStreamStatePtr stream, int n, float *r, float a, float sigma);
int vslDeleteStream(VSLStreamStatePtr *stream);
/* test hooks */
void oracle_set_seed(unsigned long long seed);
unsigned long long oracle_get_seed(void);
#endif
