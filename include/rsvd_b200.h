/* rsvd_b200.h — the thin C-ABI device layer of the B200-native RSVDPACK hot path.
 *
 * Everything below is `extern "C"`, plain pointers and 64-bit sizes, no C++/torch types.  Matrices are
 * dense column-major FP64 (d[col*ld + row], the reference's layout — matrix_vector_functions_intel_mkl.c:47-55),
 * and every `double *` parameter is a DEVICE pointer unless its name starts with `h_`.
 *
 * Who calls this: the C host code that exports the reference's own API
 * (include/rank_revealing_algorithms_intel_mkl.h, include/matrix_vector_functions_intel_mkl.h and their
 * 64bit/ twins).  Each entry point names the reference wrapper (file:line under
 * /root/reference/multi_core_mkl_code/, RRA = rank_revealing_algorithms_intel_mkl.c,
 * MVF = matrix_vector_functions_intel_mkl.c) whose vendor call it replaces.
 *
 * All functions return 0 on success, non-zero on error; the message is kept for rsvd_b200_last_error().
 * There is no CPU fallback: without a CUDA device every compute entry point fails with an error.
 */
#ifndef RSVD_B200_H
#define RSVD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef long long rsvd_i64;

/* ---- context, errors, options -------------------------------------------------------------------- */
int rsvd_b200_init(int device);                 /* implicit on first use (device 0 or $RSVD_B200_DEVICE / LOCAL_RANK) */
int rsvd_b200_device_count(void);
int rsvd_b200_status(void);                     /* 0 = no error since last clear */
const char *rsvd_b200_last_error(void);
void rsvd_b200_clear_error(void);
void *rsvd_b200_stream(void);                   /* the cudaStream_t all kernels are launched on */
void rsvd_b200_sync(void);
unsigned long long rsvd_b200_launch_count(void); /* kernels launched by this library so far */
/* options: "seed" (Omega seed, default 777 = the reference's unused `#define SEED 777`, MVH:10),
 * "verbose" (1 messages, 2 synchronising phase timer, 3 CUDA-event phase times), "force_generic_gemm",
 * "force_qr_fallback" (1: TSQR-preconditioned path for every panel, 2: shifted CholeskyQR3), "force_unblocked_qr", "single_device" (host-level calls
 * ignore the worker pool); A/B switches of individual kernels, all default 0: "no_sketch_cluster", "no_chol_dataflow" (l x l
 * Cholesky + inverse as a per-block launch sequence), "no_live_replay" (Jacobi's V rebuilt after the rotations instead of next
 * to them), "no_block_cache" (freed work buffers go straight back to the CUDA memory pool), "jacobi_transpose";
 * "qr_blocked_rows" (inputs with at most this many rows use the blocked pivoted QR; default 2048, at most 4096);
 * in a process-per-GPU job "row0" / "m_global" place this rank's row block in the global matrix.
 * "last_qr_path": 1 CholeskyQR2, 4 shifted CholeskyQR3, 2 TSQR-preconditioned, 3 Householder with explicit Q (singular panel). */
void rsvd_b200_set_option(const char *name, rsvd_i64 value);
rsvd_i64 rsvd_b200_get_option(const char *name); /* also "last_gemm_path", "last_qr_path", "qr_fallbacks", "sms", "rank", "world", "devices" (workers of the single-process pool) */

/* ---- memory --------------------------------------------------------------------------------------- */
double *rsvd_b200_dev_alloc(rsvd_i64 n_doubles);
void rsvd_b200_dev_free(double *d);
int rsvd_b200_h2d(double *d_dst, const double *h_src, rsvd_i64 n_doubles);  /* staged through pinned buffers if h_src is pageable */
int rsvd_b200_d2h(double *h_dst, const double *d_src, rsvd_i64 n_doubles);
void *rsvd_b200_host_alloc(size_t bytes);       /* pinned, zero-initialised (matrix_new for large mats, MVF:8-16) */
void rsvd_b200_host_free(void *p);

/* ---- primitives ----------------------------------------------------------------------------------- */
/* cblas_dgemm NN/TN/NT (MVF:538-561): C = alpha*op(A)*op(B) + beta*C. */
int rsvd_b200_gemm(char ta, char tb, rsvd_i64 m, rsvd_i64 n, rsvd_i64 k, double alpha, const double *A, rsvd_i64 lda,
                   const double *B, rsvd_i64 ldb, double beta, double *C, rsvd_i64 ldc);
/* initialize_random_matrix + dgemm fused (MVF:458-486 + MVF:538-552; RRA:90-95, RRA:1871-1877):
 * C(m x n) = op(A)(m x k) * Omega(k x n) with Omega(kk, j) = normal(seed, off + kk*sk + j*sc) generated on the fly. */
int rsvd_b200_sketch(char ta, rsvd_i64 m, rsvd_i64 n, rsvd_i64 k, const double *A, rsvd_i64 lda, uint64_t seed,
                     rsvd_i64 sk, rsvd_i64 sc, rsvd_i64 off, double *C, rsvd_i64 ldc);
/* initialize_random_matrix alone (MVF:458-486): d[i] = normal(seed, first + i). */
int rsvd_b200_fill_normal(double *d, rsvd_i64 n, uint64_t seed, rsvd_i64 first);
/* QR_factorization_getQ / compact_QR_factorization (MVF:1251-1263, 1214-1245; dgeqrf+dorgqr):
 * Y (m x l) <- Q in place; R (l x l upper, may be NULL).  CholeskyQR2, TSQR-preconditioned fallback. */
int rsvd_b200_orthonormalize(double *Y, rsvd_i64 ldy, rsvd_i64 m, rsvd_i64 l, double *R, rsvd_i64 ldr);
/* The l x l step of one Cholesky-QR pass (the triangular factor the reference gets from dgeqrf, MVF:1251-1263): G (n x n, upper
 * triangle read) <- R with G = R^T R (exact zeros below the diagonal), Rinv <- R^{-1}; dminmax (HOST, 2 doubles, may be NULL) =
 * min / max of diag(R).  Returns 0, the failing column + 1 when G is not positive definite, or -1 on an error. */
int rsvd_b200_chol_inv(double *G, rsvd_i64 ldg, rsvd_i64 n, double *Rinv, rsvd_i64 ldi, double *dminmax);
/* pivotedQR_mkl (RRA:924-976; dgeqp3): in place, R in the upper triangle, jpvt 0-based stored as doubles. */
int rsvd_b200_geqp3(double *A, rsvd_i64 lda, rsvd_i64 m, rsvd_i64 n, double *jpvt);
/* the same followed by dorgqr (RRA:957-964): Q (m x min(m,n)) is formed from the Householder reflectors, orthonormal for any input */
int rsvd_b200_geqp3_q(double *A, rsvd_i64 lda, rsvd_i64 m, rsvd_i64 n, double *jpvt, double *Q, rsvd_i64 ldq);
/* singular_value_decomposition (MVF:1270-1284; dgesvd 'S','S') for square n x n: A = U diag(s) Vt, s descending. */
int rsvd_b200_svd_small(double *A, rsvd_i64 lda, rsvd_i64 n, double *U, rsvd_i64 ldu, double *s, double *Vt, rsvd_i64 ldvt);
/* compute_evals_and_evecs_of_symm_matrix (MVF:1206-1209; dsyev 'V','U'): ascending w, vectors overwrite A. */
int rsvd_b200_eig_small(double *A, rsvd_i64 lda, rsvd_i64 n, double *w);
/* upper_triangular_system_solve type 1 (MVF:1477-1493; dtrsm L,U,N,N): B <- R^{-1} B. */
int rsvd_b200_trsm_left_upper(const double *R, rsvd_i64 ldr, rsvd_i64 k, double *B, rsvd_i64 ldb, rsvd_i64 ncols);
/* square_matrix_system_solve (MVF:1525-1531; dgesv): B <- A^{-1} B, A overwritten by LU. */
int rsvd_b200_lu_solve(double *A, rsvd_i64 lda, rsvd_i64 n, double *B, rsvd_i64 ldb, rsvd_i64 nrhs);
/* get_matrix_frobenius_norm (MVF:360-372). */
double rsvd_b200_frob_norm(const double *A, rsvd_i64 lda, rsvd_i64 m, rsvd_i64 n);
/* matrix_build_transpose (MVF:246-253): B(n x m) = A(m x n)^T. */
int rsvd_b200_transpose(const double *A, rsvd_i64 lda, double *B, rsvd_i64 ldb, rsvd_i64 m, rsvd_i64 n);

/* ---- device-resident algorithms (inputs and outputs stay in HBM) -------------------------------------- */
/* low_rank_svd_rand_decomp_fixed_rank (RRA:73-234).  A m x n (not modified).  omega: NULL = fused Philox
 * (seed), else an imported n x (k+p) Omega.  Outputs U m x k, S k (singular values; vnum 1 descending, vnum 2
 * ascending like the reference), V n x k. */
int rsvd_b200_svd_rand_dev(const double *A, rsvd_i64 m, rsvd_i64 n, rsvd_i64 lda, rsvd_i64 k, rsvd_i64 p, int vnum,
                           int q, int s, uint64_t seed, const double *omega, double *U, rsvd_i64 ldu, double *S,
                           double *V, rsvd_i64 ldv);
/* Same as rsvd_b200_svd_rand_dev for a HOST matrix h_A (m x n, ld m; RRA:73-95): the upload into dA (m x n device buffer
 * supplied by the caller) is pipelined in column blocks with the sketch pass when h_A is pinned memory. */
int rsvd_b200_svd_rand_host(const double *h_A, double *dA, rsvd_i64 m, rsvd_i64 n, rsvd_i64 k, rsvd_i64 p, int vnum, int q, int s,
                            uint64_t seed, double *U, rsvd_i64 ldu, double *S, double *V, rsvd_i64 ldv);
/* randQB_pb_new (RRA:1576-1801).  Awork m x n is OVERWRITTEN by the residual A - QB (the reference's private
 * copy, RRA:1630).  Q m x (kstep*nstep_max), B (kstep*nstep_max) x n; *frank = columns actually produced.
 * nstep <= 0: tolerance mode (absolute Frobenius norm < tol, RRA:1773-1775), evaluated on the device. */
int rsvd_b200_randqb_dev(double *Awork, rsvd_i64 m, rsvd_i64 n, rsvd_i64 lda, rsvd_i64 kstep, rsvd_i64 nstep, double tol,
                         int q, int s, uint64_t seed, double *Q, rsvd_i64 ldq, double *B, rsvd_i64 ldb, rsvd_i64 *frank);
/* tail shared by RRA:133-225 and RRA:289-380: SVD factors from A and an orthonormal Q (m x l). */
int rsvd_b200_svd_from_q_dev(const double *A, rsvd_i64 m, rsvd_i64 n, rsvd_i64 lda, double *Q, rsvd_i64 ldq, rsvd_i64 l,
                             rsvd_i64 k, int vnum, double *U, rsvd_i64 ldu, double *S, double *V, rsvd_i64 ldv);
/* id_rand_decomp_fixed_rank (RRA:1863-1965): I (n doubles, 0-based permutation), T k x (n-k). */
int rsvd_b200_id_rand_dev(const double *A, rsvd_i64 m, rsvd_i64 n, rsvd_i64 lda, rsvd_i64 k, rsvd_i64 p, int q, int s,
                          uint64_t seed, const double *omega, double *I, double *T, rsvd_i64 ldt);
/* id_decomp_fixed_rank_or_prec, k == min(m,n) branch (RRA:1830-1850): full pivoted QR of M (k x n). */
int rsvd_b200_id_full_dev(const double *M, rsvd_i64 k, rsvd_i64 n, rsvd_i64 ldm, double *I, double *T, rsvd_i64 ldt);
/* pivoted QR of M (r x n) + T = R11(k x k)^{-1} R12: id_rand_decomp_fromQB (oneapi_code/rank_revealing_algorithms_one_api.c:421-444)
 * and the B-factor step of id_blockrand_decomp_fixed_rank_or_prec (RRA:1996-2020). */
int rsvd_b200_id_qr_dev(const double *M, rsvd_i64 r, rsvd_i64 n, rsvd_i64 ldm, rsvd_i64 k, double *I, double *T, rsvd_i64 ldt);
/* row ID of M(:, Icol(1:k)) — second half of every two-sided ID (RRA:2071-2078, 2098-2107). */
int rsvd_b200_id_rows_dev(const double *A, rsvd_i64 m, rsvd_i64 n, rsvd_i64 lda, const double *Icol, rsvd_i64 k, double *Irow, double *S,
                          rsvd_i64 lds);
/* CUR factors from a two-sided ID (RRA:2200-2252, 2274-2326). */
int rsvd_b200_cur_from_id_dev(const double *A, rsvd_i64 m, rsvd_i64 n, rsvd_i64 lda, const double *Icol, const double *Irow, const double *T,
                              rsvd_i64 ldt, rsvd_i64 k, double *C, rsvd_i64 ldc, double *U, rsvd_i64 ldu, double *R, rsvd_i64 ldr);
/* low_rank_svd_rand_decomp_fromQB (oneapi_code/rank_revealing_algorithms_one_api.c:244-304, FP64): SVD factors from Q (m x l), B (l x n). */
int rsvd_b200_svd_from_qb_dev(const double *Q, rsvd_i64 m, rsvd_i64 ldq, const double *B, rsvd_i64 l, rsvd_i64 n, rsvd_i64 ldb, double *U,
                              rsvd_i64 ldu, double *S, double *V, rsvd_i64 ldv);
/* id_two_sided_rand_decomp_fixed_rank (RRA:2060-2082). Icol n, Irow m, T k x (n-k), S k x (m-k). */
int rsvd_b200_id_two_sided_rand_dev(const double *A, rsvd_i64 m, rsvd_i64 n, rsvd_i64 lda, rsvd_i64 k, rsvd_i64 p, int q,
                                    int s, uint64_t seed, double *Icol, double *Irow, double *T, rsvd_i64 ldt,
                                    double *S, rsvd_i64 lds);
/* cur_rand_decomp_fixed_rank (RRA:2191-2258). C m x k, U k x k, R k x n. */
int rsvd_b200_cur_rand_dev(const double *A, rsvd_i64 m, rsvd_i64 n, rsvd_i64 lda, rsvd_i64 k, rsvd_i64 p, int q, int s,
                           uint64_t seed, double *C, rsvd_i64 ldc, double *U, rsvd_i64 ldu, double *R, rsvd_i64 ldr);
/* Introspection (host only, no GPU needed): the column-pair schedule of the one-sided Jacobi kernel for an n x n problem with
 * blocks of bw columns (1, 2 or 4).  Returns N = n rounded up to a multiple of 2*bw; with pairs != NULL fills
 * pairs[(step*(N/2) + slot)*2 + {0,1}] for the N-1 steps of a sweep.  Every pair of columns must meet exactly once. */
int rsvd_b200_jacobi_schedule(int n, int bw, int *pairs);

/* ---- binary matrix files <-> device memory ----
 * The reference's format (MVF:77-133; 64-bit MVF64:78-135): two int32 (index_bits = 32) or int64 (64) m, n, then ROW-major
 * doubles.  Row blocks are DMA'd as they lie in the file and transposed on the device; the host never holds the matrix.
 * load: *dA is allocated by the library (free with rsvd_b200_dev_free), column-major m x n with ld m. */
int rsvd_b200_load_binary_dev(const char *path, int index_bits, double **dA, rsvd_i64 *m, rsvd_i64 *n);
int rsvd_b200_store_binary_dev(const char *path, int index_bits, const double *dA, rsvd_i64 lda, rsvd_i64 m, rsvd_i64 n);

/* ---- deterministic baselines and legacy entry points ---- */
/* randQB_pb (RRA:1425-1572): as randqb_dev in rank mode but re-orthogonalising against all previous blocks on EVERY step
 * (RRA:1503-1528); p = power steps (loop j <= p), no tolerance.  Q m x kstep*nstep, B kstep*nstep x n. */
int rsvd_b200_randqb_legacy_dev(double *Awork, rsvd_i64 m, rsvd_i64 n, rsvd_i64 lda, rsvd_i64 kstep, rsvd_i64 nstep, int p, int s,
                                uint64_t seed, double *Q, rsvd_i64 ldq, double *B, rsvd_i64 ldb);
/* randQB_p (RRA:1343-1421): single-vector randQB, k columns one at a time with p power steps each (BLAS-2, HBM-bound). */
int rsvd_b200_randqb_single_dev(double *Awork, rsvd_i64 m, rsvd_i64 n, rsvd_i64 lda, rsvd_i64 k, rsvd_i64 p, uint64_t seed, double *Q,
                                rsvd_i64 ldq, double *B, rsvd_i64 ldb);
/* SVD factors from (Q, B) in ASCENDING singular-value order: tail of randomized_low_rank_svd4 (RRA:660-688, dsyev order). */
int rsvd_b200_svd_from_qb_asc_dev(const double *Q, rsvd_i64 m, rsvd_i64 ldq, const double *B, rsvd_i64 l, rsvd_i64 n, rsvd_i64 ldb, double *U,
                                  rsvd_i64 ldu, double *S, double *V, rsvd_i64 ldv);
/* full thin SVD A = U diag(S) V^T, r = min(m,n) (dgesvd 'S','S' of low_rank_svd_decomp_fixed_rank_or_prec, RRA:7-69); r <= 4096 (Jacobi kernel limit). */
int rsvd_b200_svd_full_dev(const double *A, rsvd_i64 m, rsvd_i64 n, rsvd_i64 lda, double *U, rsvd_i64 ldu, double *S, double *V,
                           rsvd_i64 ldv);
/* estimate_rank_and_buildQ (MVF:1339-1400): sketch of width maxdim, sequential Gram-Schmidt with the reference's stop rule;
 * Q is an m x maxdim buffer whose first *rank columns are the orthonormal basis. */
int rsvd_b200_estimate_rank1_dev(const double *A, rsvd_i64 m, rsvd_i64 n, rsvd_i64 lda, rsvd_i64 maxdim, double tol, uint64_t seed, double *Q,
                                 rsvd_i64 ldq, rsvd_i64 *rank);
/* estimate_rank_and_buildQ2 (MVF:1404-1467): sketch grown by kblock columns until ||QQ^T A - A||_F/||QQ^T A||_F <= tol or max_cols.
 * Y, Q: m x max_cols buffers; *rank = columns used.  Successive blocks continue the Philox stream. */
int rsvd_b200_estimate_rank2_dev(const double *A, rsvd_i64 m, rsvd_i64 n, rsvd_i64 lda, rsvd_i64 kblock, double tol, uint64_t seed, double *Y,
                                 rsvd_i64 ldy, double *Q, rsvd_i64 ldq, rsvd_i64 max_cols, rsvd_i64 *rank);
/* power iterations (loop j < q, RRA:858-886) + SVD tail from an existing sketch Y = A*Omega (m x l): U m x l, S l, V n x l. */
int rsvd_b200_svd_rand_from_sketch_dev(const double *A, rsvd_i64 m, rsvd_i64 n, rsvd_i64 lda, double *Y, rsvd_i64 ldy, rsvd_i64 l, int q, int s,
                                       double *U, rsvd_i64 ldu, double *S, double *V, rsvd_i64 ldv);
/* the reference's own partial pivoted Householder QR (pivoted_QR_of_specified_rank, RRA:1012-1155, with zero_exact = 1;
 * pivoted_QR_of_specified_rank_or_prec, RRA:1159-1334, with zero_exact = 0).  k > 0: at most k steps; k <= 0: tolerance mode.
 * Awork (m x n) is destroyed.  *frank = steps taken; I n doubles (0-based); Q m x frank; R frank x n (any may be NULL). */
int rsvd_b200_pqr_partial_dev(double *Awork, rsvd_i64 lda, rsvd_i64 m, rsvd_i64 n, rsvd_i64 k, double tol, int zero_exact, double *I,
                              double *Q, rsvd_i64 ldq, double *R, rsvd_i64 ldr, rsvd_i64 *frank);
/* streamed 100*||A - U diag(S) V^T||_F/||A||_F (get_percent_error_between_two_mats after form_svd_product_matrix,
 * MVF:391-405,1304-1318) without forming the dense m x n product at once. */
double rsvd_b200_svd_percent_error_dev(const double *A, rsvd_i64 m, rsvd_i64 n, rsvd_i64 lda, const double *U, rsvd_i64 ldu,
                                       const double *S, const double *V, rsvd_i64 ldv, rsvd_i64 k);

/* ---- host-level entry points: HOST matrix in, HOST factors out (hostapi.cu) -----------------------------------------------
 * What the C host code calls for the reference API functions of the hot path.  h_A is column-major m x n with leading
 * dimension ldh (pinned memory uploads at PCIe speed; pageable memory works).  The upload is pipelined in column chunks with
 * the first pass of the algorithm.  With several active devices (RSVD_B200_DEVICES / rsvd_b200_set_devices) the matrix is
 * row-partitioned over them inside the call; otherwise the call runs on the calling thread's device (and, in a
 * process-per-GPU job set up with rsvd_b200_comm_init, h_A is this rank's row block). */
int rsvd_b200_svd_rand_h(const double *h_A, rsvd_i64 m, rsvd_i64 n, rsvd_i64 ldh, rsvd_i64 k, rsvd_i64 p, int vnum, int q, int s,
                         uint64_t seed, double *h_U, rsvd_i64 ldu, double *h_S, double *h_V, rsvd_i64 ldv);   /* RRA:73-234 */
int rsvd_b200_id_rand_h(const double *h_A, rsvd_i64 m, rsvd_i64 n, rsvd_i64 ldh, rsvd_i64 k, rsvd_i64 p, int q, int s, uint64_t seed,
                        double *h_I, double *h_T, rsvd_i64 ldt);                                              /* RRA:1863-1965 */
int rsvd_b200_id_two_sided_rand_h(const double *h_A, rsvd_i64 m, rsvd_i64 n, rsvd_i64 ldh, rsvd_i64 k, rsvd_i64 p, int q, int s, uint64_t seed,
                                  double *h_Icol, double *h_Irow, double *h_T, rsvd_i64 ldt, double *h_S, rsvd_i64 lds);   /* RRA:2060-2082 */
int rsvd_b200_cur_rand_h(const double *h_A, rsvd_i64 m, rsvd_i64 n, rsvd_i64 ldh, rsvd_i64 k, rsvd_i64 p, int q, int s, uint64_t seed,
                         double *h_C, rsvd_i64 ldc, double *h_U, rsvd_i64 ldu, double *h_R, rsvd_i64 ldr);   /* RRA:2191-2258 */
/* randQB_pb_new (RRA:1576-1801): Q (m x cap), B (cap x n) and the residual M - QB stay in HBM behind the returned handle
 * (NULL on failure); *frank = columns produced (nstep <= 0: tolerance mode, evaluated on the device). */
void *rsvd_b200_randqb_h(const double *h_A, rsvd_i64 m, rsvd_i64 n, rsvd_i64 ldh, rsvd_i64 kstep, rsvd_i64 nstep, rsvd_i64 cap, double tol,
                         int q, int s, uint64_t seed, rsvd_i64 *frank);
rsvd_i64 rsvd_b200_qb_parts(void *qb);                     /* devices the result is partitioned over */
int rsvd_b200_qb_download(void *qb, rsvd_i64 cols, double *h_Q, rsvd_i64 ldq, double *h_B, rsvd_i64 ldb);
/* SVD tail of low_rank_svd_blockrand_decomp_fixed_rank_or_prec (RRA:289-380) from the QB result alone: M^T Q is rebuilt as
 * Ares^T Q + B^T (Q^T Q), so the original M is not uploaded again.  l = rows of B used, kk = rank kept. */
int rsvd_b200_qb_svd(void *qb, rsvd_i64 l, rsvd_i64 kk, int vnum, double *h_U, rsvd_i64 ldu, double *h_S, double *h_V, rsvd_i64 ldv);
int rsvd_b200_qb_dev_ptrs(void *qb, double **dAres, double **dQ, double **dB);   /* single-device results only */
void rsvd_b200_qb_release_handle(void *qb);   /* after rsvd_b200_qb_dev_ptrs: the three buffers now belong to the caller (rsvd_b200_dev_free) */
void rsvd_b200_qb_free(void *qb);
/* residency across calls: after rsvd_b200_pin_matrix(h_A, m, n) the first host-level call that uploads h_A keeps the device
 * copy, later calls on the same pointer skip the upload (the caller promises not to modify h_A meanwhile);
 * rsvd_b200_unpin_matrix releases it.  Reference drivers call several routines on one M (driver_multi_core_mkl3.c:51,82,92,112). */
int rsvd_b200_pin_matrix(const double *h_A, rsvd_i64 m, rsvd_i64 n);
void rsvd_b200_unpin_matrix(const double *h_A);
int rsvd_b200_is_resident(const double *h_A);

/* ---- single-process multi-GPU (multi.cu): one worker thread and one NCCL rank per device inside this process ---------------
 * The reference's caller is one single-threaded C process (multi_core_mkl_code_64bit/driver1.c:40-50); with
 * RSVD_B200_DEVICES=0-7 (or "all", "0,2,5") in the environment — or this call — the host-level entry points above use all
 * listed GPUs for one API call.  n <= 1 returns to single-device operation. */
int rsvd_b200_set_devices(int n, const int *ids);
int rsvd_b200_active_devices(void);

/* ---- row-partitioned multi-GPU (one process per GPU; A_g = rows of this rank) ------------------------- */
/* NCCL is used only for the sums of n x l products and l x l Gram matrices. */
int rsvd_b200_comm_unique_id(char id_out[128]);
int rsvd_b200_comm_init(int rank, int world, const char id[128]);
void rsvd_b200_comm_destroy(void);
int rsvd_b200_allreduce_sum(double *d, rsvd_i64 count);
/* even row split used by the drivers: rows [*row0, *row0 + *rows) of an m-row matrix for `rank` of `world`. */
void rsvd_b200_row_partition(rsvd_i64 m, int world, int rank, rsvd_i64 *row0, rsvd_i64 *rows);

/* ---- measurement helpers ------------------------------------------------------------------------------ */
/* register-resident DMMA (use_dfma = 0) or DFMA (1) loop: measured FP64 peak of this GPU in TFLOP/s. */
double rsvd_b200_fp64_peak_tflops(int iters, int use_dfma);

#ifdef __cplusplus
}
#endif
#endif /* RSVD_B200_H */
