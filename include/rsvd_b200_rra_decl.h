/* rsvd_b200_rra_decl.h — declarations behind the drop-in header rank_revealing_algorithms_intel_mkl.h
 * (reference: multi_core_mkl_code/rank_revealing_algorithms_intel_mkl.h:3-110 and the _64bit twin).
 * The randomized range-finder / QB hot path is implemented on the B200 (sm_100a); see SURVEY.md §8a.
 *
 * LIMITS of this build (violations are reported through rsvd_b200_api_status() / rsvd_b200_api_last_error(); outputs are still
 * allocated, zero-filled, so existing matrix_delete calls stay safe):
 *   - sketch width k + p <= 4096, and min(m, n) <= 4096 for the full-SVD baseline low_rank_svd_decomp_fixed_rank_or_prec
 *     (the l x l factor goes through the one-sided Jacobi kernel);
 *   - k + p <= min(m, n), s > 0 (unchecked undefined behaviour in the reference);
 *   - several GPUs (RSVD_B200_DEVICES=0-7, or one process per GPU): the hot-path entry points
 *     low_rank_svd_rand_decomp_fixed_rank, randQB_pb_new, low_rank_svd_blockrand_decomp_fixed_rank_or_prec,
 *     id_rand_decomp_fixed_rank, id_two_sided_rand_decomp_fixed_rank, cur_rand_decomp_fixed_rank are row-partitioned;
 *     everything else (deterministic baselines, legacy randomized_low_rank_svd*, randQB_p[b], the block-randomized ID/CUR
 *     tails, the matrix_vector_functions helpers) runs on ONE device, the first one listed;
 *   - randQB_pb_new in tolerance mode (nstep <= 0) allocates Q (m x kstep*floor(min(m,n)/kstep)) and B at their maximum size,
 *     as the reference does (RRA:1607-1609), on the device(s) as well as on the host. */
#ifndef RSVD_B200_RRA_DECL_H
#define RSVD_B200_RRA_DECL_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- randomized SVD (RRH:5-8) ---- */
void low_rank_svd_rand_decomp_fixed_rank(mat *M, RSVD_INT k, RSVD_INT p, RSVD_INT vnum, RSVD_INT q, RSVD_INT s,
                                         RSVD_INT *frank, mat **U, mat **S, mat **V);
void low_rank_svd_blockrand_decomp_fixed_rank_or_prec(mat *M, RSVD_INT k, RSVD_INT p, double TOL, RSVD_INT vnum,
                                                      RSVD_INT kstep, RSVD_INT q, RSVD_INT s, RSVD_INT *frank,
                                                      mat **U, mat **S, mat **V);

/* ---- blocked randQB with power scheme (RRH:60) ---- */
void randQB_pb_new(mat *M, RSVD_INT kstep, RSVD_INT nstep, double TOL, RSVD_INT q, RSVD_INT s, RSVD_INT *frank,
                   mat **Q, mat **B);

/* ---- pivoted QR through the dgeqp3-compatible device kernel (RRH:40) ---- */
void pivotedQR_mkl(mat *M, mat **Q, mat **R, vec **I);

/* ---- interpolative decompositions (RRH:65-78) ---- */
void id_decomp_fixed_rank_or_prec(mat *M, RSVD_INT k, double TOL, RSVD_INT *frank, vec **I, mat **T);
void id_rand_decomp_fixed_rank(mat *M, RSVD_INT k, RSVD_INT p, RSVD_INT q, RSVD_INT s, vec **I, mat **T);
void id_two_sided_rand_decomp_fixed_rank(mat *M, RSVD_INT k, RSVD_INT p, RSVD_INT q, RSVD_INT s, vec **Icol,
                                         vec **Irow, mat **T, mat **S);

/* ---- CUR (RRH:88) ---- */
void cur_rand_decomp_fixed_rank(mat *M, RSVD_INT k, RSVD_INT p, RSVD_INT q, RSVD_INT s, mat **C, mat **U, mat **R);

/* ---- block-randomized ID / two-sided ID / CUR on top of the device QB (RRH:71, 81, 91) ---- */
void id_blockrand_decomp_fixed_rank_or_prec(mat *M, RSVD_INT k, RSVD_INT p, double TOL, RSVD_INT kstep, RSVD_INT q, RSVD_INT s,
                                            RSVD_INT *frank, vec **I, mat **T);
void id_two_sided_blockrand_decomp_fixed_rank_or_prec(mat *M, RSVD_INT k, RSVD_INT p, double TOL, RSVD_INT kstep, RSVD_INT q,
                                                      RSVD_INT s, RSVD_INT *frank, vec **Icol, vec **Irow, mat **T, mat **S);
void cur_blockrand_decomp_fixed_rank_or_prec(mat *M, RSVD_INT k, RSVD_INT p, double TOL, RSVD_INT kstep, RSVD_INT q, RSVD_INT s,
                                             RSVD_INT *frank, mat **C, mat **U, mat **R);

/* ---- SVD / ID from an existing QB (oneapi_code/rank_revealing_algorithms_one_api.h:6,10), FP64 ---- */
void low_rank_svd_rand_decomp_fromQB(mat *Q, mat *B, mat **U, mat **S, mat **V);
void id_rand_decomp_fromQB(mat *Q, mat *B, vec **I, mat **T);

/* ---- evaluation helpers used by every driver (RRH:95-110): print 100*||M - approx||_F/||M||_F ---- */
void use_low_rank_svd_for_approximation(mat *M, mat *U, mat *S, mat *V);
void use_QB_decomp_for_approximation(mat *M, mat *Q, mat *B);
void use_id_decomp_for_approximation(mat *M, mat *T, vec *I, RSVD_INT k);
void use_id_two_sided_decomp_for_approximation(mat *M, mat *T, mat *S, vec *Icol, vec *Irow, RSVD_INT k);
void use_cur_decomp_for_approximation(mat *M, mat *C, mat *U, mat *R);

/* ---- deterministic baselines (RRH:11, 44-51, 65, 75, 85, 98) ---- */
void low_rank_svd_decomp_fixed_rank_or_prec(mat *M, RSVD_INT k, double TOL, RSVD_INT *frank, mat **U, mat **S, mat **V);
void get_householder_matrix(vec *x, RSVD_INT ind1, RSVD_INT ind2, mat *H);
void pivoted_QR_of_specified_rank(mat *M, RSVD_INT k, RSVD_INT *frank, mat **Qk, mat **Rk, vec **I);
void pivoted_QR_of_specified_rank_or_prec(mat *M, RSVD_INT k, double TOL, RSVD_INT *frank, mat **Qk, mat **Rk, vec **I);
void id_two_sided_decomp_fixed_rank_or_prec(mat *M, RSVD_INT k, double TOL, RSVD_INT *frank, vec **Icol, vec **Irow, mat **T, mat **S);
void cur_decomp_fixed_rank_or_prec(mat *M, RSVD_INT k, double TOL, RSVD_INT *frank, mat **C, mat **U, mat **R);
void use_pivoted_QR_decomp_for_approximation(mat *M, mat *Qk, mat *Rk, vec *I);

/* ---- legacy entry points (RRH:14-37, 54-57) ---- */
void randomized_low_rank_svd1(mat *M, RSVD_INT k, mat **U, mat **S, mat **V);
void randomized_low_rank_svd2(mat *M, RSVD_INT k, mat **U, mat **S, mat **V);
void randomized_low_rank_svd3(mat *M, RSVD_INT k, RSVD_INT q, RSVD_INT s, mat **U, mat **S, mat **V);
void randomized_low_rank_svd4(mat *M, RSVD_INT kstep, RSVD_INT nstep, RSVD_INT p, mat **U, mat **S, mat **V);
void randomized_low_rank_svd2_autorank1(mat *M, double frac_of_max_rank, double TOL, mat **U, mat **S, mat **V);
void randomized_low_rank_svd2_autorank2(mat *M, RSVD_INT kblocksize, double TOL, mat **U, mat **S, mat **V);
void randomized_low_rank_svd3_autorank2(mat *M, RSVD_INT kblocksize, double TOL, RSVD_INT q, RSVD_INT s, mat **U, mat **S, mat **V);
void randQB_p(mat *M, RSVD_INT k, RSVD_INT p, mat **Q, mat **B);
void randQB_pb(mat *M, RSVD_INT kstep, RSVD_INT nstep, RSVD_INT p, RSVD_INT s, mat **Q, mat **B);

/* ---- out-of-band status (the reference API is void and unchecked) ---- */
int rsvd_b200_api_status(void);                 /* 0 = last call succeeded */
const char *rsvd_b200_api_last_error(void);
void rsvd_b200_api_clear_error(void);           /* reset the status (every rank_revealing_algorithms entry point does so on entry) */
double rsvd_b200_api_last_percent_error(void);  /* value printed by the last use_*_for_approximation call */

#ifdef __cplusplus
}
#endif
#endif
