/* Drop-in replacement for multi_core_mkl_code_64bit/matrix_vector_functions_intel_mkl.h (int64_t indices,
 * 16-byte file header).  Link against librsvd_b200_api64.so. */
#include <stdint.h>
#ifndef RSVD_INT
#define RSVD_INT int64_t
#endif
#ifndef RSVD_INDEX_64
#define RSVD_INDEX_64 1
#endif
#include "../rsvd_b200_matvec_decl.h"
