// capi.cu — extern "C" surface of the device layer (declared in include/rsvd_b200.h).
#include "common.cuh"

#include "pipeline.cuh"

using namespace rsvd;

#define READY()            \
    do {                   \
        ensure_init();     \
        if (!ctx().inited) return 1; \
    } while (0)

// Algorithm-level entry points return with their results complete (like the reference's synchronous API).  Besides the
// simpler contract this keeps row-partitioned ranks in step: a rank whose host thread ran ahead into the next call while
// its peer was still finishing the previous one produced sporadic 100-600 ms stalls at 2 GPUs.
static int finish(int rc) {
    RSVD_CUDA(cudaStreamSynchronize(ctx().stream));
    return rc ? rc : g_status;
}

static i64 global_rows(i64 m_local) {
    // in a row partition the option "m_global" carries the total row count; default = local
    return (ctx().world > 1 && ctx().m_global > 0) ? (i64)ctx().m_global : m_local;
}

extern "C" {

int rsvd_b200_gemm(char ta, char tb, rsvd_i64 m, rsvd_i64 n, rsvd_i64 k, double alpha, const double *A, rsvd_i64 lda,
                   const double *B, rsvd_i64 ldb, double beta, double *C, rsvd_i64 ldc) {
    READY();
    Gemm g;
    g.ta = ta; g.tb = tb; g.m = m; g.n = n; g.k = k; g.alpha = alpha; g.beta = beta;
    g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc;
    gemm(g);
    return g_status;
}

int rsvd_b200_sketch(char ta, rsvd_i64 m, rsvd_i64 n, rsvd_i64 k, const double *A, rsvd_i64 lda, uint64_t seed,
                     rsvd_i64 sk, rsvd_i64 sc, rsvd_i64 off, double *C, rsvd_i64 ldc) {
    READY();
    Gemm g;
    g.ta = ta; g.tb = 'N'; g.m = m; g.n = n; g.k = k; g.A = A; g.lda = lda; g.C = C; g.ldc = ldc;
    g.philox = true; g.seed = seed; g.ph_sk = sk; g.ph_sc = sc; g.ph_off = off;
    gemm(g);
    return g_status;
}

int rsvd_b200_fill_normal(double *d, rsvd_i64 n, uint64_t seed, rsvd_i64 first) {
    READY();
    fill_normal(d, n, seed, first);
    return g_status;
}

int rsvd_b200_orthonormalize(double *Y, rsvd_i64 ldy, rsvd_i64 m, rsvd_i64 l, double *R, rsvd_i64 ldr) {
    READY();
    if (l > m && ctx().world == 1) { set_error("rsvd_b200_orthonormalize: need m >= l (got %lld x %lld)", (long long)m, (long long)l); return 1; }
    orthonormalize(Y, ldy, m, l, R, ldr, true);
    return g_status;
}

int rsvd_b200_chol_inv(double *G, rsvd_i64 ldg, rsvd_i64 n, double *Rinv, rsvd_i64 ldi, double *dminmax) {
    READY();
    int info = ctx().no_chol_dataflow ? -1 : chol_inv_upper(G, ldg, n, Rinv, ldi, dminmax);
    if (info < 0 && !g_status) {      // per-block launch sequence (option no_chol_dataflow, or no cooperative launch)
        info = potrf_upper(G, ldg, n);
        if (info == 0) trtri_upper(G, ldg, n, Rinv, ldi);
        if (dminmax) dminmax[0] = dminmax[1] = 0.0;
    }
    return g_status ? -1 : info;
}

int rsvd_b200_geqp3(double *A, rsvd_i64 lda, rsvd_i64 m, rsvd_i64 n, double *jpvt) {
    READY();
    geqp3(A, lda, m, n, jpvt);
    return g_status;
}

int rsvd_b200_geqp3_q(double *A, rsvd_i64 lda, rsvd_i64 m, rsvd_i64 n, double *jpvt, double *Q, rsvd_i64 ldq) {
    READY();
    geqp3_q(A, lda, m, n, jpvt, Q, ldq);
    return finish(g_status);
}

int rsvd_b200_svd_small(double *A, rsvd_i64 lda, rsvd_i64 n, double *U, rsvd_i64 ldu, double *s, double *Vt, rsvd_i64 ldvt) {
    READY();
    jacobi_svd(A, lda, n, U, ldu, s, Vt, ldvt);
    return g_status;
}

int rsvd_b200_eig_small(double *A, rsvd_i64 lda, rsvd_i64 n, double *w) {
    READY();
    jacobi_eig(A, lda, n, w);
    return g_status;
}

int rsvd_b200_trsm_left_upper(const double *R, rsvd_i64 ldr, rsvd_i64 k, double *B, rsvd_i64 ldb, rsvd_i64 ncols) {
    READY();
    trsm_left_upper(R, ldr, k, B, ldb, ncols);
    return g_status;
}

int rsvd_b200_lu_solve(double *A, rsvd_i64 lda, rsvd_i64 n, double *B, rsvd_i64 ldb, rsvd_i64 nrhs) {
    READY();
    int info = lu_solve(A, lda, n, B, ldb, nrhs);
    if (info) set_error("rsvd_b200_lu_solve: zero pivot at column %d", info);
    return g_status;
}

double rsvd_b200_frob_norm(const double *A, rsvd_i64 lda, rsvd_i64 m, rsvd_i64 n) {
    ensure_init();
    if (!ctx().inited) return -1.0;
    return frob_norm(A, lda, m, n);
}

int rsvd_b200_transpose(const double *A, rsvd_i64 lda, double *B, rsvd_i64 ldb, rsvd_i64 m, rsvd_i64 n) {
    READY();
    transpose(A, lda, B, ldb, m, n);
    return g_status;
}

int rsvd_b200_svd_rand_dev(const double *A, rsvd_i64 m, rsvd_i64 n, rsvd_i64 lda, rsvd_i64 k, rsvd_i64 p, int vnum, int q,
                           int s, uint64_t seed, const double *omega, double *U, rsvd_i64 ldu, double *S, double *V,
                           rsvd_i64 ldv) {
    READY();
    return finish(svd_rand(A, m, n, lda, k, p, vnum, q, s, seed, omega, U, ldu, S, V, ldv));
}

int rsvd_b200_svd_rand_host(const double *h_A, double *dA, rsvd_i64 m, rsvd_i64 n, rsvd_i64 k, rsvd_i64 p, int vnum, int q, int s,
                            uint64_t seed, double *U, rsvd_i64 ldu, double *S, double *V, rsvd_i64 ldv) {
    READY();
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, h_A) != cudaSuccess || a.type != cudaMemoryTypeHost) {
        (void)cudaGetLastError();
        // pageable source: staged upload, then the ordinary device-resident algorithm
        if (rsvd_b200_h2d(dA, h_A, m * n)) return g_status;
        return finish(svd_rand(dA, m, n, m, k, p, vnum, q, s, seed, nullptr, U, ldu, S, V, ldv));
    }
    return finish(svd_rand_host(h_A, dA, m, n, k, p, vnum, q, s, seed, U, ldu, S, V, ldv));
}

int rsvd_b200_randqb_dev(double *Awork, rsvd_i64 m, rsvd_i64 n, rsvd_i64 lda, rsvd_i64 kstep, rsvd_i64 nstep, double tol,
                         int q, int s, uint64_t seed, double *Q, rsvd_i64 ldq, double *B, rsvd_i64 ldb, rsvd_i64 *frank) {
    READY();
    // capacity of Q/B in columns/rows: ldb rows of B were allocated by the caller
    return finish(randqb(Awork, m, n, lda, kstep, nstep, tol, q, s, seed, Q, ldq, B, ldb, ldb, frank, 0));
}

int rsvd_b200_randqb_legacy_dev(double *Awork, rsvd_i64 m, rsvd_i64 n, rsvd_i64 lda, rsvd_i64 kstep, rsvd_i64 nstep, int p, int s,
                                uint64_t seed, double *Q, rsvd_i64 ldq, double *B, rsvd_i64 ldb) {
    READY();
    if (nstep <= 0) { set_error("rsvd_b200: randQB_pb needs nstep > 0"); return 1; }
    i64 frank = 0;
    return finish(randqb(Awork, m, n, lda, kstep, nstep, 0.0, p, s, seed, Q, ldq, B, ldb, ldb, &frank, 1));
}

int rsvd_b200_randqb_single_dev(double *Awork, rsvd_i64 m, rsvd_i64 n, rsvd_i64 lda, rsvd_i64 k, rsvd_i64 p, uint64_t seed, double *Q,
                                rsvd_i64 ldq, double *B, rsvd_i64 ldb) {
    READY();
    return finish(randqb_single(Awork, m, n, lda, k, p, seed, Q, ldq, B, ldb));
}

int rsvd_b200_svd_full_dev(const double *A, rsvd_i64 m, rsvd_i64 n, rsvd_i64 lda, double *U, rsvd_i64 ldu, double *S, double *V,
                           rsvd_i64 ldv) {
    READY();
    return finish(svd_full(A, m, n, lda, U, ldu, S, V, ldv));
}

int rsvd_b200_jacobi_schedule(int n, int bw, int *pairs) { return jacobi_schedule(n, bw, pairs); }

int rsvd_b200_estimate_rank1_dev(const double *A, rsvd_i64 m, rsvd_i64 n, rsvd_i64 lda, rsvd_i64 maxdim, double tol, uint64_t seed, double *Q,
                                 rsvd_i64 ldq, rsvd_i64 *rank) {
    READY();
    return finish(estimate_rank1(A, m, n, lda, maxdim, tol, seed, Q, ldq, rank));
}

int rsvd_b200_estimate_rank2_dev(const double *A, rsvd_i64 m, rsvd_i64 n, rsvd_i64 lda, rsvd_i64 kblock, double tol, uint64_t seed, double *Y,
                                 rsvd_i64 ldy, double *Q, rsvd_i64 ldq, rsvd_i64 max_cols, rsvd_i64 *rank) {
    READY();
    return finish(estimate_rank2(A, m, n, lda, kblock, tol, seed, Y, ldy, Q, ldq, max_cols, rank));
}

int rsvd_b200_svd_rand_from_sketch_dev(const double *A, rsvd_i64 m, rsvd_i64 n, rsvd_i64 lda, double *Y, rsvd_i64 ldy, rsvd_i64 l, int q, int s,
                                       double *U, rsvd_i64 ldu, double *S, double *V, rsvd_i64 ldv) {
    READY();
    return finish(svd_rand_from_sketch(A, m, n, lda, Y, ldy, l, q, s, U, ldu, S, V, ldv));
}

int rsvd_b200_pqr_partial_dev(double *Awork, rsvd_i64 lda, rsvd_i64 m, rsvd_i64 n, rsvd_i64 k, double tol, int zero_exact, double *I,
                              double *Q, rsvd_i64 ldq, double *R, rsvd_i64 ldr, rsvd_i64 *frank) {
    READY();
    *frank = pqr_partial(Awork, lda, m, n, k, tol, zero_exact, I, Q, ldq, R, ldr);
    return finish(g_status);
}

int rsvd_b200_svd_from_q_dev(const double *A, rsvd_i64 m, rsvd_i64 n, rsvd_i64 lda, double *Q, rsvd_i64 ldq, rsvd_i64 l,
                             rsvd_i64 k, int vnum, double *U, rsvd_i64 ldu, double *S, double *V, rsvd_i64 ldv) {
    READY();
    return finish(svd_from_q(A, m, n, lda, Q, ldq, l, k, vnum, U, ldu, S, V, ldv));
}

int rsvd_b200_id_rand_dev(const double *A, rsvd_i64 m, rsvd_i64 n, rsvd_i64 lda, rsvd_i64 k, rsvd_i64 p, int q, int s,
                          uint64_t seed, const double *omega, double *I, double *T, rsvd_i64 ldt) {
    READY();
    return finish(id_rand(A, m, n, lda, k, p, q, s, seed, omega, I, T, ldt));
}

int rsvd_b200_id_full_dev(const double *M, rsvd_i64 k, rsvd_i64 n, rsvd_i64 ldm, double *I, double *T, rsvd_i64 ldt) {
    READY();
    return finish(id_full(M, k, n, ldm, I, T, ldt));
}

int rsvd_b200_id_qr_dev(const double *M, rsvd_i64 r, rsvd_i64 n, rsvd_i64 ldm, rsvd_i64 k, double *I, double *T, rsvd_i64 ldt) {
    READY();
    return finish(id_qr(M, r, n, ldm, k, I, T, ldt));
}

int rsvd_b200_id_rows_dev(const double *A, rsvd_i64 m, rsvd_i64 n, rsvd_i64 lda, const double *Icol, rsvd_i64 k, double *Irow, double *S,
                          rsvd_i64 lds) {
    READY();
    return finish(id_rows(A, m, n, lda, Icol, k, Irow, S, lds, global_rows(m)));
}

int rsvd_b200_cur_from_id_dev(const double *A, rsvd_i64 m, rsvd_i64 n, rsvd_i64 lda, const double *Icol, const double *Irow, const double *T,
                              rsvd_i64 ldt, rsvd_i64 k, double *C, rsvd_i64 ldc, double *U, rsvd_i64 ldu, double *R, rsvd_i64 ldr) {
    READY();
    return finish(cur_from_id(A, m, n, lda, Icol, Irow, T, ldt, k, C, ldc, U, ldu, R, ldr));
}

int rsvd_b200_svd_from_qb_dev(const double *Q, rsvd_i64 m, rsvd_i64 ldq, const double *B, rsvd_i64 l, rsvd_i64 n, rsvd_i64 ldb, double *U,
                              rsvd_i64 ldu, double *S, double *V, rsvd_i64 ldv) {
    READY();
    return finish(svd_from_qb(Q, m, ldq, B, l, n, ldb, U, ldu, S, V, ldv, 0));
}

int rsvd_b200_svd_from_qb_asc_dev(const double *Q, rsvd_i64 m, rsvd_i64 ldq, const double *B, rsvd_i64 l, rsvd_i64 n, rsvd_i64 ldb, double *U,
                                  rsvd_i64 ldu, double *S, double *V, rsvd_i64 ldv) {
    READY();
    return finish(svd_from_qb(Q, m, ldq, B, l, n, ldb, U, ldu, S, V, ldv, 1));
}

int rsvd_b200_id_two_sided_rand_dev(const double *A, rsvd_i64 m, rsvd_i64 n, rsvd_i64 lda, rsvd_i64 k, rsvd_i64 p, int q,
                                    int s, uint64_t seed, double *Icol, double *Irow, double *T, rsvd_i64 ldt, double *S,
                                    rsvd_i64 lds) {
    READY();
    return finish(id_two_sided_rand(A, m, n, lda, k, p, q, s, seed, Icol, Irow, T, ldt, S, lds, global_rows(m)));
}

int rsvd_b200_cur_rand_dev(const double *A, rsvd_i64 m, rsvd_i64 n, rsvd_i64 lda, rsvd_i64 k, rsvd_i64 p, int q, int s,
                           uint64_t seed, double *C, rsvd_i64 ldc, double *U, rsvd_i64 ldu, double *R, rsvd_i64 ldr) {
    READY();
    return finish(cur_rand(A, m, n, lda, k, p, q, s, seed, C, ldc, U, ldu, R, ldr, global_rows(m)));
}

double rsvd_b200_svd_percent_error_dev(const double *A, rsvd_i64 m, rsvd_i64 n, rsvd_i64 lda, const double *U, rsvd_i64 ldu,
                                       const double *S, const double *V, rsvd_i64 ldv, rsvd_i64 k) {
    return svd_percent_error(A, m, n, lda, U, ldu, S, V, ldv, k);
}

double rsvd_b200_fp64_peak_tflops(int iters, int use_dfma) {
    ensure_init();
    if (!ctx().inited) return -1.0;
    return dmma_peak_tflops(iters, use_dfma);
}

}  // extern "C"
