/* rsvd_b200_matvec_decl.h — declarations behind the drop-in header matrix_vector_functions_intel_mkl.h.
 * Index type is RSVD_INT: `int` for the multi_core_mkl_code ABI, `int64_t` for multi_core_mkl_code_64bit
 * (reference: multi_core_mkl_code/matrix_vector_functions_intel_mkl.h:18-360 and the _64bit twin).
 * Struct layout, names, argument order and ownership rules are the reference's; the bodies are new
 * (lowrankmatrixdecompositioncodes_b200/csrc/host/matrix_vector_functions.c) and route every BLAS/LAPACK-class
 * operation to the sm_100a device layer (rsvd_b200.h).  No MKL headers are needed or included. */
#ifndef RSVD_B200_MATVEC_DECL_H
#define RSVD_B200_MATVEC_DECL_H

#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
#include <time.h>
#include <sys/time.h>

#ifndef RSVD_INT
#define RSVD_INT int
#endif

#ifndef SEED
#define SEED 777
#endif
#ifndef min
#define min(x,y) (((x) < (y)) ? (x) : (y))
#endif
#ifndef max
#define max(x,y) (((x) > (y)) ? (x) : (y))
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* dense column-major matrix: element (i,j) is d[j*nrows + i]  (reference MVH:18-21, MVF:47-55) */
typedef struct {
    RSVD_INT nrows, ncols;
    double *d;
} mat;

/* vector; also carries permutations as 0-based indices stored in doubles (reference MVH:24-27, RRA:967-970) */
typedef struct {
    RSVD_INT nrows;
    double *d;
} vec;

/* -- lifetime (outputs of the algorithms are allocated by the callee with these, freed by the caller) -- */
mat *matrix_new(RSVD_INT nrows, RSVD_INT ncols);          /* zero-initialised */
vec *vector_new(RSVD_INT nrows);
void matrix_delete(mat *M);
void vector_delete(vec *v);

/* -- element access -- */
void matrix_set_element(mat *M, RSVD_INT row_num, RSVD_INT col_num, double val);
double matrix_get_element(mat *M, RSVD_INT row_num, RSVD_INT col_num);
void vector_set_element(vec *v, RSVD_INT row_num, double val);
double vector_get_element(vec *v, RSVD_INT row_num);

/* -- binary file format: header (RSVD_INT m, RSVD_INT n) then m*n doubles in ROW-major order -- */
mat *matrix_load_from_binary_file(char *fname);
void matrix_write_to_binary_file(mat *M, char *fname);

void matrix_print(mat *M);
void vector_print(vec *v);

/* -- elementwise helpers (host loops) -- */
void vector_set_data(vec *v, double *data);
void vector_scale(vec *v, double scalar);
void matrix_scale(mat *M, double scalar);
double vector_get2norm(vec *v);
void vector_get_min_element(vec *v, RSVD_INT *minindex, double *minval);
void vector_get_max_element(vec *v, RSVD_INT *maxindex, double *maxval);
void vector_copy(vec *d, vec *s);
void matrix_copy(mat *D, mat *S);
void matrix_hard_threshold(mat *M, double TOL);
void matrix_build_transpose(mat *Mt, mat *M);
void vector_sub(vec *a, vec *b);
void matrix_sub(mat *A, mat *B);
void matrix_sub_column_times_row_vector(mat *A, vec *u, vec *v);
double get_matrix_frobenius_norm(mat *M);
double get_matrix_max_abs_element(mat *M);
double vector_dot_product(vec *u, vec *v);
double get_matrix_column_norm_squared(mat *M, RSVD_INT colnum);
double matrix_getmaxcolnorm(mat *M);
void compute_matrix_column_norms(mat *M, vec *column_norms);
double get_percent_error_between_two_mats(mat *A, mat *B);

/* -- Gaussian test matrix: float32-valued N(0,1) entries in linear column-major order.  The reference seeds MKL's
 *    MCG31 stream with time(NULL); here entry i is Philox(seed, i) with seed = option "seed" (default SEED) -- */
void initialize_random_matrix(mat *M);

/* -- products (device GEMM) -- */
void matrix_matrix_mult(mat *A, mat *B, mat *C);            /* C = A B   */
void matrix_transpose_matrix_mult(mat *A, mat *B, mat *C);  /* C = A^T B */
void matrix_matrix_transpose_mult(mat *A, mat *B, mat *C);  /* C = A B^T */
void matrix_vector_mult(mat *M, vec *x, vec *y);            /* y = M x   */
void matrix_transpose_vector_mult(mat *M, vec *x, vec *y);  /* y = M^T x */

/* -- rows / columns -- */
void matrix_get_col(mat *M, RSVD_INT j, vec *column_vec);
void matrix_set_col(mat *M, RSVD_INT j, vec *column_vec);
void matrix_get_row(mat *M, RSVD_INT i, vec *row_vec);
void matrix_set_row(mat *M, RSVD_INT i, vec *row_vec);
void matrix_get_selected_columns(mat *M, RSVD_INT *inds, mat *Mc);
void matrix_set_selected_columns(mat *M, RSVD_INT *inds, mat *Mc);
void matrix_get_selected_rows(mat *M, RSVD_INT *inds, mat *Mr);
void matrix_set_selected_rows(mat *M, RSVD_INT *inds, mat *Mr);
void matrix_copy_symmetric(mat *S, mat *M);
void matrix_keep_only_upper_triangular(mat *M);
void initialize_diagonal_matrix(mat *D, vec *data);
void initialize_identity_matrix(mat *D);
void invert_diagonal_matrix(mat *Dinv, mat *D);
void invert_upper_triangular_matrix(mat *Minv);

/* -- slicing -- */
void fill_vector_from_row_list(vec *input, vec *inds, vec *output);
void matrix_copy_first_rows(mat *M_out, mat *M);
void matrix_copy_first_columns(mat *M_out, mat *M);
void matrix_copy_first_columns_with_param(mat *D, mat *S, RSVD_INT num_columns);
void matrix_copy_first_k_rows_and_columns(mat *M_out, mat *M);
void matrix_copy_all_rows_and_last_columns_from_indexk(mat *M_out, mat *M, RSVD_INT k);
void fill_matrix_from_first_rows(mat *M, RSVD_INT k, mat *M_k);
void fill_matrix_from_last_rows(mat *M, RSVD_INT k, mat *M_k);
void fill_matrix_from_first_columns(mat *M, RSVD_INT k, mat *M_k);
void fill_matrix_from_last_columns(mat *M, RSVD_INT k, mat *M_k);
void fill_matrix_from_last_columns_from_specified_one(mat *M, RSVD_INT k, mat *M_k);
void fill_matrix_from_lower_right_corner(mat *M, RSVD_INT k, mat *M_out);
void fill_matrix_from_first_columns_from_list(mat *M, vec *I, RSVD_INT k, mat *M_k);  /* M_k = M(:, I(1:k)) */
void fill_matrix_from_first_rows_from_list(mat *M, vec *I, RSVD_INT k, mat *M_k);     /* M_k = M(I(1:k), :) */
void fill_matrix_from_last_columns_from_list(mat *M, vec *I, RSVD_INT k, mat *M_k);
void resize_matrix_by_columns(mat **M, RSVD_INT k);
void resize_matrix_by_columns_from_end(mat **M, RSVD_INT k);
void resize_matrix_by_rows(mat **M, RSVD_INT k);
void resize_matrix_by_rows_from_end(mat **M, RSVD_INT k);
void append_matrices_horizontally(mat *A, mat *B, mat *C);
void append_matrices_vertically(mat *A, mat *B, mat *C);
void vector_build_rewrapped(vec *Iinv, vec *I);               /* Iinv(I) = 0..len-1 */

/* -- factorizations (device kernels) -- */
void compute_evals_and_evecs_of_symm_matrix(mat *S, vec *evals);   /* ascending; vectors overwrite S */
void compact_QR_factorization(mat *M, mat *Q, mat *R);
void QR_factorization_getQ(mat *M, mat *Q);
void singular_value_decomposition(mat *M, mat *U, mat *S, mat *Vt); /* S is a full diagonal matrix, descending */
void form_svd_product_matrix(mat *U, mat *S, mat *V, mat *P);       /* P = U S V^T */
void form_cur_product_matrix(mat *C, mat *U, mat *R, mat *P);       /* P = C U R   */
void upper_triangular_system_solve(mat *A, mat *B, mat *X, int solve_type);
void square_matrix_system_solve(mat *A, mat *X, mat *B);

/* legacy range-finder helpers (MVH:226, 241, 342, 345) */
void project_vector(vec *v, vec *u, vec *p);
void build_orthonormal_basis_from_mat(mat *A, mat *Q);
void estimate_rank_and_buildQ(mat *M, double frac_of_max_rank, double TOL, mat **Q, RSVD_INT *good_rank);
void estimate_rank_and_buildQ2(mat *M, RSVD_INT kblock, double TOL, mat **Y, mat **Q, RSVD_INT *good_rank);
double get_seconds_frac(struct timeval start_timeval, struct timeval end_timeval);

#ifdef __cplusplus
}
#endif
#endif
