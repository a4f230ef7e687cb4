/* Drop-in replacement for multi_core_mkl_code/matrix_vector_functions_intel_mkl.h (int indices).
 * Existing drivers `#include "rank_revealing_algorithms_intel_mkl.h"` and relink against
 * librsvd_b200_api32.so unchanged.  See rsvd_b200_matvec_decl.h. */
#ifndef RSVD_INT
#define RSVD_INT int
#endif
#include "rsvd_b200_matvec_decl.h"
