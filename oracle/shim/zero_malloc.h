/* oracle/shim/zero_malloc.h — TEST INFRASTRUCTURE ONLY.
 * Force-included when building the reference: the reference hands an uninitialised malloc'd jpvt
 * to dgeqp3 (rank_revealing_algorithms_intel_mkl.c:934,947) and the 64-bit matrix_new does not
 * zero.  Intended semantics are "all zero"; make malloc zero-initialise. */
#include <stdlib.h>
#define malloc(n) calloc((n), 1)
