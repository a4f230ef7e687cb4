// pipeline.cuh — declarations of the device-resident algorithms (pipeline.cu) shared by the C-ABI layers (capi.cu, hostapi.cu).
#pragma once
#include "common.cuh"

namespace rsvd {
int svd_rand_impl(const double *A, i64 m, i64 n, i64 lda, i64 k, i64 p, int vnum, int q, int s, uint64_t seed, const double *omega, DBuf *Y0,
                  const Upload *up, double *U, i64 ldu, double *S, double *V, i64 ldv);
int svd_from_q(const double *A, i64 m, i64 n, i64 lda, double *Q, i64 ldq, i64 l, i64 k, int vnum, double *U, i64 ldu,
               double *S, double *V, i64 ldv);
int svd_from_q_residual(const double *Ares, i64 m, i64 n, i64 lda, double *Q, i64 ldq, i64 l, const double *B, i64 ldb, i64 k, int vnum,
                        double *U, i64 ldu, double *S, double *V, i64 ldv);
int svd_rand(const double *A, i64 m, i64 n, i64 lda, i64 k, i64 p, int vnum, int q, int s, uint64_t seed,
             const double *omega, double *U, i64 ldu, double *S, double *V, i64 ldv);
int svd_rand_host(const double *hA, double *dA, i64 m, i64 n, i64 k, i64 p, int vnum, int q, int s, uint64_t seed, double *U, i64 ldu,
                  double *S, double *V, i64 ldv);
int randqb(double *A, i64 m, i64 n, i64 lda, i64 kstep, i64 nstep, double tol, int q, int s, uint64_t seed, double *Q,
           i64 ldq, double *B, i64 ldb, i64 max_rank, i64 *frank_out, int legacy_reorth, const Upload *up = nullptr);
int randqb_single(double *A, i64 m, i64 n, i64 lda, i64 k, i64 p, uint64_t seed, double *Q, i64 ldq, double *B, i64 ldb);
int svd_full(const double *A, i64 m, i64 n, i64 lda, double *U, i64 ldu, double *S, double *V, i64 ldv);
int jacobi_schedule(int n, int bw, int *pairs);
int estimate_rank1(const double *A, i64 m, i64 n, i64 lda, i64 maxdim, double tol, uint64_t seed, double *Q, i64 ldq, i64 *rank_out);
int estimate_rank2(const double *A, i64 m, i64 n, i64 lda, i64 kblock, double tol, uint64_t seed, double *Y, i64 ldy, double *Q, i64 ldq,
                   i64 max_cols, i64 *rank_out);
int svd_rand_from_sketch(const double *A, i64 m, i64 n, i64 lda, double *Y, i64 ldy, i64 l, int q, int s, double *U, i64 ldu,
                         double *S, double *V, i64 ldv);
int id_rand(const double *A, i64 m, i64 n, i64 lda, i64 k, i64 p, int q, int s, uint64_t seed, const double *omega,
            double *I, double *T, i64 ldt, const Upload *up = nullptr);
int id_full(const double *M, i64 k, i64 n, i64 ldm, double *I, double *T, i64 ldt);
int id_qr(const double *M, i64 r, i64 n, i64 ldm, i64 k, double *I, double *T, i64 ldt);
int id_rows(const double *A, i64 m, i64 n, i64 lda, const double *Icol, i64 k, double *Irow, double *S, i64 lds, i64 m_global);
int cur_from_id(const double *A, i64 m, i64 n, i64 lda, const double *Icol, const double *Irow, const double *T, i64 ldt, i64 k,
                double *Cm, i64 ldc, double *U, i64 ldu, double *R, i64 ldr);
int svd_from_qb(const double *Q, i64 m, i64 ldq, const double *B, i64 l, i64 n, i64 ldb, double *U, i64 ldu, double *S, double *V, i64 ldv,
                int ascending);
int id_two_sided_rand(const double *A, i64 m, i64 n, i64 lda, i64 k, i64 p, int q, int s, uint64_t seed, double *Icol,
                      double *Irow, double *T, i64 ldt, double *S, i64 lds, i64 m_global, const Upload *up = nullptr);
int cur_rand(const double *A, i64 m, i64 n, i64 lda, i64 k, i64 p, int q, int s, uint64_t seed, double *Cm, i64 ldc,
             double *U, i64 ldu, double *R, i64 ldr, i64 m_global, const Upload *up = nullptr);
double svd_percent_error(const double *A, i64 m, i64 n, i64 lda, const double *U, i64 ldu, const double *S,
                         const double *V, i64 ldv, i64 k);
}  // namespace rsvd
