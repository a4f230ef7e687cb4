/* Drop-in replacement for multi_core_mkl_code/rank_revealing_algorithms_intel_mkl.h (int indices). */
#include "matrix_vector_functions_intel_mkl.h"
#include "rsvd_b200_rra_decl.h"
