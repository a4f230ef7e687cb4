/* rank_revealing_algorithms.c — C host side of the algorithms layer (reference: rank_revealing_algorithms_intel_mkl.c
 * = RRA).  Each exported function keeps the reference's name, argument order, output allocation rules
 * (callee allocates with matrix_new/vector_new, caller frees) and parameter quirks, uploads M once, runs the whole
 * algorithm on the B200 through the C-ABI of rsvd_b200.h, and downloads the factors.  The host never touches the
 * numbers in between: A stays resident in HBM for all passes. */
#include "rank_revealing_algorithms_intel_mkl.h"
#include "rsvd_b200.h"
#include "rsvd_b200_host_util.h"
#include <stdarg.h>

typedef RSVD_INT idx_t;

/* ---- status ------------------------------------------------------------------------------------------------------ */
static int g_api_status = 0;
static char g_api_err[1024] = "";
static double g_last_percent_error = -1.0;
static char g_api_warn[512];

void rsvd_api_error(const char *fmt, ...) {
    if (g_api_status) return;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_api_err, sizeof(g_api_err), fmt, ap);
    va_end(ap);
    g_api_status = 1;
    if (getenv("RSVD_B200_VERBOSE")) fprintf(stderr, "[rsvd_b200 api] %s\n", g_api_err);
}
void rsvd_api_sync_error(void) {
    if (rsvd_b200_status()) rsvd_api_error("%s", rsvd_b200_last_error());
}
static int g_api_depth = 0;                   /* > 0: inside a composite call (one exported function calling others) */
void rsvd_api_begin(void) {
    if (g_api_depth > 0) return;
    g_api_status = 0; g_api_err[0] = 0; g_api_warn[0] = 0;
    rsvd_b200_clear_error();
}
void rsvd_api_enter(void) { rsvd_api_begin(); ++g_api_depth; }
void rsvd_api_leave(void) { if (g_api_depth > 0) --g_api_depth; }
int rsvd_b200_api_status(void) { rsvd_api_sync_error(); return g_api_status; }
void rsvd_b200_api_clear_error(void) { g_api_depth = 0; rsvd_api_begin(); }   /* the helpers of matrix_vector_functions do not reset the status themselves */
const char *rsvd_b200_api_last_error(void) { return g_api_err; }
/* warnings (an argument was adjusted, the call went ahead): recorded, never turned into an error status */
static void rsvd_api_warning(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_api_warn, sizeof(g_api_warn), fmt, ap);
    va_end(ap);
    if (getenv("RSVD_B200_VERBOSE")) fprintf(stderr, "[rsvd_b200 api] warning: %s\n", g_api_warn);
}
const char *rsvd_b200_api_last_warning(void) { return g_api_warn; }
double rsvd_b200_api_last_percent_error(void) { return g_last_percent_error; }

static int verbose(void) {
    static int v = -1;
    if (v < 0) { const char *s = getenv("RSVD_B200_VERBOSE"); v = s ? atoi(s) : 0; }
    return v;
}
static uint64_t omega_seed(void) { return (uint64_t)rsvd_b200_get_option("seed"); }

mat *rsvd_matrix_new_uninit(idx_t nrows, idx_t ncols);
static mat *download_mat(const double *d, idx_t r, idx_t c) {
    mat *M = d ? rsvd_matrix_new_uninit(r, c) : matrix_new(r, c);   /* fully overwritten by the download (zeros if there is nothing to download) */
    rsvd_download(M->d, d, (size_t)r * (size_t)c);
    if (d && (g_api_status || rsvd_b200_status())) memset(M->d, 0, (size_t)r * (size_t)c * sizeof(double));   /* failed call: zeros, as documented */
    return M;
}
static vec *download_vec(const double *d, idx_t n) {
    vec *v = vector_new(n);
    rsvd_download(v->d, d, (size_t)n);
    return v;
}
static mat *diag_from_device(const double *ds, idx_t k) {
    mat *S = matrix_new(k, k);
    double *s = (double *)malloc((size_t)(k ? k : 1) * sizeof(double));
    rsvd_download(s, ds, (size_t)k);
    for (idx_t i = 0; i < k; ++i) S->d[(size_t)i * (size_t)k + (size_t)i] = s[i];
    free(s);
    return S;
}

/* sets S (k x k, zeroed) to diag(s) */
static void fill_diag(mat *S, const double *s, idx_t k) {
    for (idx_t i = 0; i < k; ++i) S->d[(size_t)i * (size_t)k + (size_t)i] = s[i];
}
/* a failed call returns zeros, as documented */
static void zero_mat(mat *M) { if (M && M->d) memset(M->d, 0, (size_t)M->nrows * (size_t)M->ncols * sizeof(double)); }
static void zero_vec(vec *v) { if (v && v->d) memset(v->d, 0, (size_t)v->nrows * sizeof(double)); }

/* ---- low_rank_svd_rand_decomp_fixed_rank (RRA:73-234) -------------------------------------------------------------
 * *frank is never written by the reference; neither here.  One host-level call: the upload of M is
 * pipelined with the sketch pass, all 2q passes run on the resident copy (row-partitioned over the active devices when
 * RSVD_B200_DEVICES lists several), the factors are downloaded straight into the callee-allocated outputs. */
void low_rank_svd_rand_decomp_fixed_rank(mat *M, idx_t k, idx_t p, idx_t vnum, idx_t q, idx_t s, idx_t *frank,
                                         mat **U, mat **S, mat **V) {
    (void)frank;
    rsvd_api_begin();
    idx_t m = M->nrows, n = M->ncols;
    *U = NULL; *S = NULL; *V = NULL;
    if (k <= 0 || p < 0 || s <= 0 || k + p > min(m, n)) {
        rsvd_api_error("low_rank_svd_rand_decomp_fixed_rank: need 0 < k+p <= min(m,n) and s > 0 (k=%lld p=%lld s=%lld, M %lld x %lld)",
                       (long long)k, (long long)p, (long long)s, (long long)m, (long long)n);
        *U = matrix_new(m, max(k, 0)); *S = matrix_new(max(k, 0), max(k, 0)); *V = matrix_new(n, max(k, 0));
        return;
    }
    struct timeval t0, t1, t2;
    gettimeofday(&t0, NULL);
    *U = rsvd_matrix_new_uninit(m, k); *S = matrix_new(k, k); *V = rsvd_matrix_new_uninit(n, k);   /* U, V are fully overwritten */
    double *sv = (double *)calloc((size_t)k, sizeof(double));
    gettimeofday(&t1, NULL);
    rsvd_b200_svd_rand_h(M->d, m, n, m, k, p, (int)vnum, (int)q, (int)s, omega_seed(), (*U)->d, m, sv, (*V)->d, n);
    gettimeofday(&t2, NULL);
    rsvd_api_sync_error();
    if (g_api_status) { zero_mat(*U); zero_mat(*V); }
    else fill_diag(*S, sv, k);
    free(sv);
    if (verbose())
        fprintf(stderr, "[rsvd_b200 api] output allocation %.3f s, upload + device + download %.3f s\n", get_seconds_frac(t0, t1), get_seconds_frac(t1, t2));
}

/* ---- randQB_pb_new (RRA:1576-1801) -------------------------------------------------------------------------------- */
/* Runs the blocked QB on the device(s); Q, B and the residual stay in HBM behind *qb_out. */
static int randqb_device(mat *M, idx_t kstep, idx_t nstep, double TOL, idx_t q, idx_t s, idx_t *frank, void **qb_out, idx_t *cap_out) {
    idx_t m = M->nrows, n = M->ncols, r = min(m, n);
    *qb_out = NULL; *cap_out = 0; *frank = 0;
    if (kstep > (idx_t)(r / 2)) {                       /* RRA:1589-1592 */
        kstep = (idx_t)r / 10;
        if (verbose()) printf("kstep resized to %lld\n", (long long)kstep);
    }
    if (kstep <= 0 || s <= 0) { rsvd_api_error("randQB_pb_new: invalid kstep/s"); return 1; }
    idx_t cap;
    if (nstep <= 0) cap = (idx_t)(r / kstep) * kstep;   /* tolerance mode (RRA:1595-1602) */
    else cap = kstep * nstep;
    if (cap > r) cap = (r / kstep) * kstep;
    rsvd_i64 fr = 0;
    void *qb = rsvd_b200_randqb_h(M->d, m, n, m, kstep, nstep <= 0 ? 0 : cap / kstep, cap, TOL, (int)q, (int)s, omega_seed(), &fr);
    *frank = (idx_t)fr;
    *qb_out = qb; *cap_out = cap;
    rsvd_api_sync_error();
    return g_api_status;
}

void randQB_pb_new(mat *M, idx_t kstep, idx_t nstep, double TOL, idx_t q, idx_t s, idx_t *frank, mat **Q, mat **B) {
    rsvd_api_begin();
    idx_t m = M->nrows, n = M->ncols, cap = 0;
    void *qb = NULL;
    *Q = NULL; *B = NULL;
    if (randqb_device(M, kstep, nstep, TOL, q, s, frank, &qb, &cap)) {
        rsvd_b200_qb_free(qb);
        *Q = matrix_new(m, 0); *B = matrix_new(0, n);
        return;
    }
    /* rank mode returns all kstep*nstep columns; tolerance mode is cut to frank (RRA:1782-1789) */
    idx_t cols = (nstep <= 0) ? *frank : cap;
    *Q = rsvd_matrix_new_uninit(m, cols); *B = rsvd_matrix_new_uninit(cols, n);
    rsvd_b200_qb_download(qb, cols, (*Q)->d, m, (*B)->d, cols);
    rsvd_b200_qb_free(qb);
    rsvd_api_sync_error();
    if (g_api_status) { zero_mat(*Q); zero_mat(*B); }
}

/* ---- low_rank_svd_blockrand_decomp_fixed_rank_or_prec (RRA:239-381) -------------------------------------------------
 * Parameter handling follows the reference exactly, including quirk Q1: nstep = (k+p)/kstep by INTEGER
 * division is computed after the k <= 0 branch and overwrites its nstep = 0, so for ordinary arguments the QB runs in rank
 * mode; only when that division itself yields 0 (k = -1 with p <= kstep, or k = p = 0 with kstep >= min(m,n)/2, ...) does
 * randQB_pb_new see nstep = 0 and run tolerance-driven, exactly as in the reference (RRA:251-266).
 * The SVD tail (RRA:289-380: Bt = M^T Q, ...) works from the QB residual in HBM; M is not uploaded a second time. */
void low_rank_svd_blockrand_decomp_fixed_rank_or_prec(mat *M, idx_t k, idx_t p, double TOL, idx_t vnum, idx_t kstep,
                                                      idx_t q, idx_t s, idx_t *frank, mat **U, mat **S, mat **V) {
    rsvd_api_begin();
    idx_t m = M->nrows, n = M->ncols;
    int rankMode = k > 0;
    if (p < kstep && (p + kstep) < min(m, n)) p = kstep;          /* RRA:260-262 */
    idx_t nstep = (kstep > 0) ? (k + p) / kstep : 0;              /* RRA:263; <= 0: tolerance mode inside randQB_pb_new */
    *U = NULL; *S = NULL; *V = NULL;
    idx_t cap = 0, fr = 0;
    void *qb = NULL;
    if (randqb_device(M, kstep, nstep, TOL, q, s, &fr, &qb, &cap) || fr <= 0) {
        if (!g_api_status) rsvd_api_error("low_rank_svd_blockrand: the QB produced no columns");
        rsvd_b200_qb_free(qb);
        *U = matrix_new(m, 0); *S = matrix_new(0, 0); *V = matrix_new(n, 0);
        return;
    }
    idx_t l = (nstep <= 0) ? fr : cap;                             /* l = B->nrows (RRA:271): tolerance mode cuts B to frank rows */
    if (rankMode) *frank = k;                                      /* RRA:274-275 */
    else *frank = (idx_t)round(((double)fr / ((double)fr + (double)p + 1e-6)) * (double)fr);   /* RRA:278 */
    idx_t kk = *frank;
    if (kk > l) kk = l;
    if (kk <= 0) { rsvd_api_error("low_rank_svd_blockrand: rank %lld", (long long)kk); rsvd_b200_qb_free(qb); *U = matrix_new(m, 0); *S = matrix_new(0, 0); *V = matrix_new(n, 0); return; }
    *U = rsvd_matrix_new_uninit(m, kk); *S = matrix_new(kk, kk); *V = rsvd_matrix_new_uninit(n, kk);
    double *sv = (double *)calloc((size_t)kk, sizeof(double));
    rsvd_b200_qb_svd(qb, l, kk, (int)vnum, (*U)->d, m, sv, (*V)->d, n);
    rsvd_b200_qb_free(qb);
    rsvd_api_sync_error();
    if (g_api_status) { zero_mat(*U); zero_mat(*V); }
    else fill_diag(*S, sv, kk);
    free(sv);
}

/* ---- pivotedQR_mkl (RRA:924-976) ------------------------------------------------------------------------------------
 * dgeqp3 + dorgqr on the device: R and I from the dgeqp3-compatible kernel, the explicit Q from its Householder reflectors
 * (orthonormal whatever the rank or conditioning of M, like the reference's). */
static void pivotedQR_mkl_impl(mat *M, mat **Q, mat **R, vec **I) {
    idx_t m = M->nrows, n = M->ncols, k = min(m, n);
    idx_t Rcols = (m <= n) ? n : k;
    double *dW = rsvd_upload(M->d, (size_t)m * (size_t)n);
    double *dI = rsvd_b200_dev_alloc(n + 1), *dQ = rsvd_b200_dev_alloc((rsvd_i64)m * k + 1);
    *Q = NULL; *R = NULL; *I = NULL;
    if (dW && dI && dQ) rsvd_b200_geqp3_q(dW, m, m, n, dI, dQ, m);
    mat *W = download_mat(dW, m, n);
    *I = download_vec(dI, n);
    *Q = download_mat(dQ, m, k);
    rsvd_b200_dev_free(dW); rsvd_b200_dev_free(dI); rsvd_b200_dev_free(dQ);
    *R = matrix_new(k, Rcols);
    for (idx_t j = 0; j < Rcols; ++j)
        for (idx_t i = 0; i <= j && i < k; ++i) (*R)->d[(size_t)j * k + i] = W->d[(size_t)j * m + i];
    matrix_delete(W);
    rsvd_api_sync_error();
}
void pivotedQR_mkl(mat *M, mat **Q, mat **R, vec **I) {   /* composite: nested calls keep one status */
    rsvd_api_enter();
    pivotedQR_mkl_impl(M, Q, R, I);
    rsvd_api_leave();
}

/* ---- ID family ---------------------------------------------------------------------------------------------------- */
void id_rand_decomp_fixed_rank(mat *M, idx_t k, idx_t p, idx_t q, idx_t s, vec **I, mat **T) {
    rsvd_api_begin();
    idx_t m = M->nrows, n = M->ncols;
    *I = NULL; *T = NULL;
    if (k <= 0 || p < 0 || s <= 0 || k + p > min(m, n)) {
        rsvd_api_error("id_rand_decomp_fixed_rank: need 0 < k+p <= min(m,n) and s > 0");
        *I = vector_new(n); *T = matrix_new(max(k, 0), n - max(k, 0));
        return;
    }
    *I = vector_new(n); *T = matrix_new(k, n - k);
    rsvd_b200_id_rand_h(M->d, m, n, m, k, p, (int)q, (int)s, omega_seed(), (*I)->d, (*T)->d, k);
    rsvd_api_sync_error();
    if (g_api_status) { zero_vec(*I); zero_mat(*T); }
}

/* Runs the reference's own partial pivoted QR on the device and leaves I, R (kmax x n, ld kmax) there.  Returns frank. */
static idx_t pqr_device(mat *M, idx_t k, double TOL, int zero_exact, double **dI_out, double **dQ_out, double **dR_out, idx_t *kmax_out) {
    idx_t m = M->nrows, n = M->ncols, r = min(m, n);
    idx_t kmax = (k <= 0) ? r : min(k, r);
    double *dW = rsvd_upload(M->d, (size_t)m * (size_t)n);          /* the private copy R = M (RRA:1192) */
    double *dI = rsvd_b200_dev_alloc(n + 1);
    double *dQ = dQ_out ? rsvd_b200_dev_alloc((rsvd_i64)m * kmax + 1) : NULL;
    double *dR = rsvd_b200_dev_alloc((rsvd_i64)kmax * n + 1);
    rsvd_i64 fr = 0;
    if (dW && dI && dR && (!dQ_out || dQ))
        rsvd_b200_pqr_partial_dev(dW, m, m, n, k, TOL, zero_exact, dI, dQ, m, dR, kmax, &fr);
    rsvd_b200_dev_free(dW);
    *dI_out = dI; if (dQ_out) *dQ_out = dQ; *dR_out = dR; *kmax_out = kmax;
    rsvd_api_sync_error();
    return (idx_t)fr;
}

static void pqr_host(mat *M, idx_t k, double TOL, int zero_exact, const char *who, idx_t *frank, mat **Qk, mat **Rk, vec **I) {
    rsvd_api_begin();
    idx_t m = M->nrows, n = M->ncols, kmax = 0;
    if (k > min(m, n)) rsvd_api_warning("%s: k = %lld exceeds min(m,n) = %lld; clamped", who, (long long)k, (long long)min(m, n));
    double *dI = NULL, *dQ = NULL, *dR = NULL;
    *frank = pqr_device(M, k, TOL, zero_exact, &dI, &dQ, &dR, &kmax);
    *I = download_vec(dI, n);
    *Qk = matrix_new(m, *frank);
    if (dQ) rsvd_download((*Qk)->d, dQ, (size_t)m * (size_t)*frank);
    *Rk = download_mat(dR, kmax, n);
    if (*frank != kmax) resize_matrix_by_rows(Rk, *frank);
    rsvd_b200_dev_free(dI); rsvd_b200_dev_free(dQ); rsvd_b200_dev_free(dR);
    rsvd_api_sync_error();
}

/* RRA:1012-1155 (stops on an exactly zero pivot norm) and RRA:1159-1334 (|norm^2| < 1e-10, or R22norm < TOL when k <= 0) */
void pivoted_QR_of_specified_rank(mat *M, idx_t k, idx_t *frank, mat **Qk, mat **Rk, vec **I) {
    if (k <= 0) { rsvd_api_begin(); rsvd_api_error("pivoted_QR_of_specified_rank: need k > 0"); *frank = 0; *Qk = matrix_new(M->nrows, 0); *Rk = matrix_new(0, M->ncols); *I = vector_new(M->ncols); return; }
    pqr_host(M, k, 0.0, 1, "pivoted_QR_of_specified_rank", frank, Qk, Rk, I);
}
void pivoted_QR_of_specified_rank_or_prec(mat *M, idx_t k, double TOL, idx_t *frank, mat **Qk, mat **Rk, vec **I) {
    pqr_host(M, k, TOL, 0, "pivoted_QR_of_specified_rank_or_prec", frank, Qk, Rk, I);
}

/* RRA:982-1008: H(ind1:ind2) = x(ind1:ind2), H(ind1) -= ||x(ind1:ind2)||, scaled so that I - H H^T is the reflector. */
void get_householder_matrix(vec *x, idx_t ind1, idx_t ind2, mat *H) {
    double nrm = 0.0;
    for (idx_t i = ind1; i < ind2; ++i) { H->d[i] = x->d[i]; nrm += x->d[i] * x->d[i]; }
    H->d[ind1] -= sqrt(nrm);
    double val = get_matrix_frobenius_norm(H);
    if (val > 0) matrix_scale(H, sqrt(2.0 / (val * val)));
}

/* RRA:1807-1854.  k == min(m,n): full dgeqp3-rule pivoted QR (the branch the randomized hot path reaches, RRA:1830-1834);
 * k < min(m,n) or k <= 0: the reference's own partial Householder QR in rank or tolerance mode (RRA:1828-1829). */
void id_decomp_fixed_rank_or_prec(mat *M, idx_t k, double TOL, idx_t *frank, vec **I, mat **T) {
    rsvd_api_begin();
    idx_t m = M->nrows, n = M->ncols;
    *I = NULL; *T = NULL;
    if (k < min(m, n)) {
        idx_t kmax = 0;
        double *dI = NULL, *dR = NULL;
        idx_t f = pqr_device(M, k, TOL, 0, &dI, NULL, &dR, &kmax);
        *frank = f;
        *I = download_vec(dI, n);
        if (f > 0 && n > f && dR) rsvd_b200_trsm_left_upper(dR, kmax, f, dR + (size_t)f * kmax, kmax, n - f);   /* T = triu(Rk1)^{-1} Rk2 */
        mat *Tfull = matrix_new(kmax, n - f);
        if (dR && n > f) rsvd_download(Tfull->d, dR + (size_t)f * kmax, (size_t)kmax * (size_t)(n - f));
        if (f != kmax) resize_matrix_by_rows(&Tfull, f);
        *T = Tfull;
        rsvd_b200_dev_free(dI); rsvd_b200_dev_free(dR);
        rsvd_api_sync_error();
        return;
    }
    if (m > n) { rsvd_api_error("id_decomp_fixed_rank_or_prec: need nrows <= ncols"); *I = vector_new(n); *T = matrix_new(k, 0); return; }
    k = min(m, n);
    *frank = k;
    double *dM = rsvd_upload(M->d, (size_t)m * (size_t)n);
    double *dI = rsvd_b200_dev_alloc(n + 1), *dT = rsvd_b200_dev_alloc((rsvd_i64)k * (n - k) + 1);
    if (dM && dI && dT) rsvd_b200_id_full_dev(dM, k, n, m, dI, dT, k);
    rsvd_b200_dev_free(dM);
    *I = download_vec(dI, n);
    *T = download_mat(dT, k, n - k);
    rsvd_b200_dev_free(dI); rsvd_b200_dev_free(dT);
    rsvd_api_sync_error();
}

/* RRA:2034-2056: column ID of M, then the (full) row ID of M(:, Icol(1:frank))^T */
static void id_two_sided_decomp_fixed_rank_or_prec_impl(mat *M, idx_t k, double TOL, idx_t *frank, vec **Icol, vec **Irow, mat **T, mat **S) {
    idx_t m = M->nrows;
    id_decomp_fixed_rank_or_prec(M, k, TOL, frank, Icol, T);
    if (g_api_status) { *Irow = vector_new(m); *S = matrix_new(*frank, m - *frank); return; }
    mat *MI = matrix_new(m, *frank), *MIt = matrix_new(*frank, m);
    fill_matrix_from_first_columns_from_list(M, *Icol, *frank, MI);
    matrix_build_transpose(MIt, MI);
    id_decomp_fixed_rank_or_prec(MIt, *frank, 0, frank, Irow, S);
    matrix_delete(MI); matrix_delete(MIt);
}
void id_two_sided_decomp_fixed_rank_or_prec(mat *M, idx_t k, double TOL, idx_t *frank, vec **Icol, vec **Irow, mat **T, mat **S) {   /* composite: nested calls keep one status */
    rsvd_api_enter();
    id_two_sided_decomp_fixed_rank_or_prec_impl(M, k, TOL, frank, Icol, Irow, T, S);
    rsvd_api_leave();
}

/* RRA:2115-2187: two-sided ID, then the same CUR tail as cur_rand_decomp_fixed_rank (device: rsvd_b200_cur_from_id_dev) */
static void cur_decomp_fixed_rank_or_prec_impl(mat *M, idx_t k, double TOL, idx_t *frank, mat **C, mat **U, mat **R) {
    idx_t m = M->nrows, n = M->ncols;
    vec *Icol = NULL, *Irow = NULL;
    mat *T = NULL, *S = NULL;
    *C = NULL; *U = NULL; *R = NULL;
    id_two_sided_decomp_fixed_rank_or_prec(M, k, TOL, frank, &Icol, &Irow, &T, &S);
    k = *frank;                                                      /* RRA:2125-2127 (k <= 0) — and the only consistent size otherwise */
    if (g_api_status || k <= 0) {
        *C = matrix_new(m, max(k, 0)); *U = matrix_new(max(k, 0), max(k, 0)); *R = matrix_new(max(k, 0), n);
    } else {
        double *dA = rsvd_upload(M->d, (size_t)m * (size_t)n);
        double *dIc = rsvd_upload(Icol->d, (size_t)n), *dIr = rsvd_upload(Irow->d, (size_t)m), *dT = rsvd_upload(T->d, (size_t)k * (size_t)(n - k));
        double *dC = rsvd_b200_dev_alloc((rsvd_i64)m * k + 1), *dU = rsvd_b200_dev_alloc((rsvd_i64)k * k + 1), *dR = rsvd_b200_dev_alloc((rsvd_i64)k * n + 1);
        if (dA && dIc && dIr && dT && dC && dU && dR) rsvd_b200_cur_from_id_dev(dA, m, n, m, dIc, dIr, dT, k, k, dC, m, dU, k, dR, k);
        rsvd_b200_dev_free(dA); rsvd_b200_dev_free(dIc); rsvd_b200_dev_free(dIr); rsvd_b200_dev_free(dT);
        *C = download_mat(dC, m, k); *U = download_mat(dU, k, k); *R = download_mat(dR, k, n);
        rsvd_b200_dev_free(dC); rsvd_b200_dev_free(dU); rsvd_b200_dev_free(dR);
        rsvd_api_sync_error();
    }
    vector_delete(Icol); vector_delete(Irow); matrix_delete(T); matrix_delete(S);
}
void cur_decomp_fixed_rank_or_prec(mat *M, idx_t k, double TOL, idx_t *frank, mat **C, mat **U, mat **R) {   /* composite: nested calls keep one status */
    rsvd_api_enter();
    cur_decomp_fixed_rank_or_prec_impl(M, k, TOL, frank, C, U, R);
    rsvd_api_leave();
}

void id_two_sided_rand_decomp_fixed_rank(mat *M, idx_t k, idx_t p, idx_t q, idx_t s, vec **Icol, vec **Irow, mat **T, mat **S) {
    rsvd_api_begin();
    idx_t m = M->nrows, n = M->ncols;
    *Icol = NULL; *Irow = NULL; *T = NULL; *S = NULL;
    if (k <= 0 || p < 0 || s <= 0 || k + p > min(m, n)) {
        rsvd_api_error("id_two_sided_rand_decomp_fixed_rank: need 0 < k+p <= min(m,n) and s > 0");
        *Icol = vector_new(n); *Irow = vector_new(m); *T = matrix_new(max(k, 0), n - max(k, 0)); *S = matrix_new(max(k, 0), m - max(k, 0));
        return;
    }
    *Icol = vector_new(n); *Irow = vector_new(m); *T = matrix_new(k, n - k); *S = matrix_new(k, m - k);
    rsvd_b200_id_two_sided_rand_h(M->d, m, n, m, k, p, (int)q, (int)s, omega_seed(), (*Icol)->d, (*Irow)->d, (*T)->d, k, (*S)->d, k);
    rsvd_api_sync_error();
    if (g_api_status) { zero_vec(*Icol); zero_vec(*Irow); zero_mat(*T); zero_mat(*S); }
}

void cur_rand_decomp_fixed_rank(mat *M, idx_t k, idx_t p, idx_t q, idx_t s, mat **C, mat **U, mat **R) {
    rsvd_api_begin();
    idx_t m = M->nrows, n = M->ncols;
    *C = NULL; *U = NULL; *R = NULL;
    if (k <= 0 || p < 0 || s <= 0 || k + p > min(m, n)) {
        rsvd_api_error("cur_rand_decomp_fixed_rank: need 0 < k+p <= min(m,n) and s > 0");
        *C = matrix_new(m, max(k, 0)); *U = matrix_new(max(k, 0), max(k, 0)); *R = matrix_new(max(k, 0), n);
        return;
    }
    *C = rsvd_matrix_new_uninit(m, k); *U = matrix_new(k, k); *R = rsvd_matrix_new_uninit(k, n);
    rsvd_b200_cur_rand_h(M->d, m, n, m, k, p, (int)q, (int)s, omega_seed(), (*C)->d, m, (*U)->d, k, (*R)->d, k);
    rsvd_api_sync_error();
    if (g_api_status) { zero_mat(*C); zero_mat(*U); zero_mat(*R); }
}

/* ---- block-randomized ID / two-sided ID / CUR (RRA:1969-2027, 2086-2111, 2262-2332): consumers of the device QB ------------
 * Column ID of the small factor B (pivoted QR on the device), then — for the two-sided and CUR variants — the same row-ID and
 * CUR tails as the non-blocked routines, on the ORIGINAL M (re-uploaded over the QB residual). */
static int blockrand_column_id(mat *M, idx_t k, idx_t p, double TOL, idx_t kstep, idx_t q, idx_t s, idx_t *frank,
                               double **dA_out, double **dI_out, double **dT_out) {
    idx_t n = M->ncols;
    idx_t nstep = (kstep > 0) ? (k + p) / kstep : 0;          /* RRA:1974 (integer division) */
    int rankMode = k > 0;
    if (!rankMode) nstep = 0;                                 /* RRA:1980-1983 */
    idx_t cap = 0, fr = 0;
    double *dA = NULL, *dQ = NULL, *dB = NULL;
    void *qb = NULL;
    *dA_out = NULL; *dI_out = NULL; *dT_out = NULL;
    /* these tails gather columns / rows of the ORIGINAL M on one device: keep the QB on the calling thread's GPU */
    rsvd_b200_set_option("single_device", 1);
    int bad = randqb_device(M, kstep, nstep, TOL, q, s, &fr, &qb, &cap);
    rsvd_b200_set_option("single_device", 0);
    if (bad || rsvd_b200_qb_dev_ptrs(qb, &dA, &dQ, &dB)) {
        rsvd_b200_qb_free(qb);
        rsvd_api_sync_error();
        return 1;
    }
    rsvd_b200_qb_release_handle(qb);                       /* the three buffers are ours now */
    rsvd_b200_dev_free(dQ);
    idx_t rows = (nstep <= 0) ? fr : cap;                     /* B is cut to frank rows only in tolerance mode (RRA:1782-1789) */
    if (rankMode) *frank = k;                                 /* RRA:2001-2002 */
    else *frank = (idx_t)round(((double)fr / ((double)fr + (double)p + 1e-6)) * (double)fr);   /* RRA:2005 */
    idx_t kk = *frank;
    if (kk <= 0 || kk > rows || kk > n) {
        rsvd_api_error("id_blockrand: rank %lld exceeds the %lld rows of B (the reference reads out of bounds here, SURVEY Q7)", (long long)kk, (long long)rows);
        rsvd_b200_dev_free(dA); rsvd_b200_dev_free(dB);
        return 1;
    }
    double *dI = rsvd_b200_dev_alloc(n + 1), *dT = rsvd_b200_dev_alloc((rsvd_i64)kk * (n - kk) + 1);
    if (dI && dT) rsvd_b200_id_qr_dev(dB, rows, n, cap, kk, dI, dT, kk);     /* pivotedQR_mkl(B) + T = Rk1^{-1} Rk2 (RRA:1996-2020) */
    rsvd_b200_dev_free(dB);
    *dA_out = dA; *dI_out = dI; *dT_out = dT;
    rsvd_api_sync_error();
    return g_api_status;
}

void id_blockrand_decomp_fixed_rank_or_prec(mat *M, idx_t k, idx_t p, double TOL, idx_t kstep, idx_t q, idx_t s, idx_t *frank, vec **I, mat **T) {
    rsvd_api_begin();
    idx_t n = M->ncols;
    double *dA = NULL, *dI = NULL, *dT = NULL;
    *I = NULL; *T = NULL;
    if (blockrand_column_id(M, k, p, TOL, kstep, q, s, frank, &dA, &dI, &dT)) { *I = vector_new(n); *T = matrix_new(0, n); rsvd_b200_dev_free(dA); return; }
    rsvd_b200_dev_free(dA);
    *I = download_vec(dI, n);
    *T = download_mat(dT, *frank, n - *frank);
    rsvd_b200_dev_free(dI); rsvd_b200_dev_free(dT);
    rsvd_api_sync_error();
}

void id_two_sided_blockrand_decomp_fixed_rank_or_prec(mat *M, idx_t k, idx_t p, double TOL, idx_t kstep, idx_t q, idx_t s, idx_t *frank,
                                                      vec **Icol, vec **Irow, mat **T, mat **S) {
    rsvd_api_begin();
    idx_t m = M->nrows, n = M->ncols;
    double *dA = NULL, *dI = NULL, *dT = NULL;
    *Icol = NULL; *Irow = NULL; *T = NULL; *S = NULL;
    if (blockrand_column_id(M, k, p, TOL, kstep, q, s, frank, &dA, &dI, &dT)) {
        *Icol = vector_new(n); *Irow = vector_new(m); *T = matrix_new(0, n); *S = matrix_new(0, m); rsvd_b200_dev_free(dA); return;
    }
    idx_t kk = *frank;
    if (rsvd_b200_h2d(dA, M->d, (rsvd_i64)m * n)) rsvd_api_sync_error();     /* the original M replaces the QB residual */
    double *dIr = rsvd_b200_dev_alloc(m + 1), *dS = rsvd_b200_dev_alloc((rsvd_i64)kk * (m - kk) + 1);
    if (dIr && dS) rsvd_b200_id_rows_dev(dA, m, n, m, dI, kk, dIr, dS, kk);  /* RRA:2098-2107 */
    rsvd_b200_dev_free(dA);
    *Icol = download_vec(dI, n); *Irow = download_vec(dIr, m);
    *T = download_mat(dT, kk, n - kk); *S = download_mat(dS, kk, m - kk);
    rsvd_b200_dev_free(dI); rsvd_b200_dev_free(dIr); rsvd_b200_dev_free(dT); rsvd_b200_dev_free(dS);
    rsvd_api_sync_error();
}

void cur_blockrand_decomp_fixed_rank_or_prec(mat *M, idx_t k, idx_t p, double TOL, idx_t kstep, idx_t q, idx_t s, idx_t *frank,
                                             mat **C, mat **U, mat **R) {
    rsvd_api_begin();
    idx_t m = M->nrows, n = M->ncols;
    double *dA = NULL, *dI = NULL, *dT = NULL;
    *C = NULL; *U = NULL; *R = NULL;
    if (blockrand_column_id(M, k, p, TOL, kstep, q, s, frank, &dA, &dI, &dT)) {
        *C = matrix_new(m, 0); *U = matrix_new(0, 0); *R = matrix_new(0, n); rsvd_b200_dev_free(dA); return;
    }
    idx_t kk = *frank;                                                       /* RRA:2272 */
    if (rsvd_b200_h2d(dA, M->d, (rsvd_i64)m * n)) rsvd_api_sync_error();
    double *dIr = rsvd_b200_dev_alloc(m + 1), *dS = rsvd_b200_dev_alloc((rsvd_i64)kk * (m - kk) + 1);
    if (dIr && dS) rsvd_b200_id_rows_dev(dA, m, n, m, dI, kk, dIr, dS, kk);
    rsvd_b200_dev_free(dS);
    double *dC = rsvd_b200_dev_alloc((rsvd_i64)m * kk + 1), *dU = rsvd_b200_dev_alloc((rsvd_i64)kk * kk + 1), *dR = rsvd_b200_dev_alloc((rsvd_i64)kk * n + 1);
    if (dIr && dC && dU && dR) rsvd_b200_cur_from_id_dev(dA, m, n, m, dI, dIr, dT, kk, kk, dC, m, dU, kk, dR, kk);   /* RRA:2274-2326 */
    rsvd_b200_dev_free(dA); rsvd_b200_dev_free(dI); rsvd_b200_dev_free(dIr); rsvd_b200_dev_free(dT);
    *C = download_mat(dC, m, kk); *U = download_mat(dU, kk, kk); *R = download_mat(dR, kk, n);
    rsvd_b200_dev_free(dC); rsvd_b200_dev_free(dU); rsvd_b200_dev_free(dR);
    rsvd_api_sync_error();
}

/* ---- SVD / ID from an existing QB (oneapi_code/rank_revealing_algorithms_one_api.c:244-304, 421-444, in FP64) ---------------- */
void low_rank_svd_rand_decomp_fromQB(mat *Q, mat *B, mat **U, mat **S, mat **V) {
    rsvd_api_begin();
    idx_t m = Q->nrows, l = Q->ncols, n = B->ncols;
    *U = NULL; *S = NULL; *V = NULL;
    if (B->nrows != l) { rsvd_api_error("low_rank_svd_rand_decomp_fromQB: Q is %lld x %lld but B has %lld rows", (long long)m, (long long)l, (long long)B->nrows);
                         *U = matrix_new(m, l); *S = matrix_new(l, l); *V = matrix_new(n, l); return; }
    double *dQ = rsvd_upload(Q->d, (size_t)m * l), *dB = rsvd_upload(B->d, (size_t)l * n);
    double *dU = rsvd_b200_dev_alloc((rsvd_i64)m * l + 1), *dS = rsvd_b200_dev_alloc(l + 1), *dV = rsvd_b200_dev_alloc((rsvd_i64)n * l + 1);
    if (dQ && dB && dU && dS && dV) rsvd_b200_svd_from_qb_dev(dQ, m, m, dB, l, n, l, dU, m, dS, dV, n);
    rsvd_b200_dev_free(dQ); rsvd_b200_dev_free(dB);
    *U = download_mat(dU, m, l); *S = diag_from_device(dS, l); *V = download_mat(dV, n, l);
    rsvd_b200_dev_free(dU); rsvd_b200_dev_free(dS); rsvd_b200_dev_free(dV);
    rsvd_api_sync_error();
}

void id_rand_decomp_fromQB(mat *Q, mat *B, vec **I, mat **T) {
    rsvd_api_begin();
    idx_t l = B->nrows, n = B->ncols, k = Q->ncols;              /* k = Q->ncols (:429) */
    *I = NULL; *T = NULL;
    if (k <= 0 || k > l || k > n) { rsvd_api_error("id_rand_decomp_fromQB: need 0 < cols(Q) <= rows(B) <= cols(B)"); *I = vector_new(n); *T = matrix_new(0, n); return; }
    double *dB = rsvd_upload(B->d, (size_t)l * n);
    double *dI = rsvd_b200_dev_alloc(n + 1), *dT = rsvd_b200_dev_alloc((rsvd_i64)k * (n - k) + 1);
    if (dB && dI && dT) rsvd_b200_id_qr_dev(dB, l, n, l, k, dI, dT, k);
    rsvd_b200_dev_free(dB);
    *I = download_vec(dI, n); *T = download_mat(dT, k, n - k);
    rsvd_b200_dev_free(dI); rsvd_b200_dev_free(dT);
    rsvd_api_sync_error();
}

/* ---- evaluation helpers (RRA:2337-2573): 100*||M - approx||_F/||M||_F, printed like the reference ------------------- */
void use_low_rank_svd_for_approximation(mat *M, mat *U, mat *S, mat *V) {
    rsvd_api_begin();
    idx_t m = M->nrows, n = M->ncols, k = S->nrows;
    double *s = (double *)malloc((size_t)(k ? k : 1) * sizeof(double));
    for (idx_t i = 0; i < k; ++i) s[i] = S->d[(size_t)i * k + i];
    double *dA = rsvd_upload(M->d, (size_t)m * n), *dU = rsvd_upload(U->d, (size_t)m * k), *dV = rsvd_upload(V->d, (size_t)n * k),
           *dS = rsvd_upload(s, (size_t)k);
    free(s);
    double pe = -1.0;
    if (dA && dU && dV && dS) pe = rsvd_b200_svd_percent_error_dev(dA, m, n, m, dU, m, dS, dV, n, k);   /* streamed, no dense m x n product */
    rsvd_b200_dev_free(dA); rsvd_b200_dev_free(dU); rsvd_b200_dev_free(dV); rsvd_b200_dev_free(dS);
    g_last_percent_error = pe;
    printf("percent_error between M and U S V^T = %f\n", pe);
    rsvd_api_sync_error();
}

static void report(mat *M, mat *P, const char *what) {
    g_last_percent_error = get_percent_error_between_two_mats(M, P);
    printf("percent_error between M and %s = %f\n", what, g_last_percent_error);
}

static void use_QB_decomp_for_approximation_impl(mat *M, mat *Q, mat *B) {
    rsvd_api_begin();
    mat *P = matrix_new(M->nrows, M->ncols);
    matrix_matrix_mult(Q, B, P);
    report(M, P, "QB");
    matrix_delete(P);
}
void use_QB_decomp_for_approximation(mat *M, mat *Q, mat *B) {   /* composite: nested calls keep one status */
    rsvd_api_enter();
    use_QB_decomp_for_approximation_impl(M, Q, B);
    rsvd_api_leave();
}

/* M ~ M(:,I(1:k)) [I_k T] P^T  (RRA:2402-2476) */
static void use_id_decomp_for_approximation_impl(mat *M, mat *T, vec *I, idx_t k) {
    rsvd_api_begin();
    idx_t m = M->nrows, n = M->ncols;
    mat *C = matrix_new(m, k), *CT = matrix_new(m, n - k), *P = matrix_new(m, n);
    fill_matrix_from_first_columns_from_list(M, I, k, C);
    matrix_matrix_mult(C, T, CT);
    for (idx_t j = 0; j < n; ++j) {
        idx_t dst = (idx_t)I->d[j];
        const double *src = (j < k) ? &C->d[(size_t)j * m] : &CT->d[(size_t)(j - k) * m];
        memcpy(&P->d[(size_t)dst * m], src, (size_t)m * sizeof(double));
    }
    report(M, P, "ID approximation");
    matrix_delete(C); matrix_delete(CT); matrix_delete(P);
}
void use_id_decomp_for_approximation(mat *M, mat *T, vec *I, idx_t k) {   /* composite: nested calls keep one status */
    rsvd_api_enter();
    use_id_decomp_for_approximation_impl(M, T, I, k);
    rsvd_api_leave();
}

/* M ~ [I_k S]^T(Irow^{-1}) M(Irow(1:k), Icol(1:k)) [I_k T](Icol^{-1})  (RRA:2481-2558) */
static void use_id_two_sided_decomp_for_approximation_impl(mat *M, mat *T, mat *S, vec *Icol, vec *Irow, idx_t k) {
    rsvd_api_begin();
    idx_t m = M->nrows, n = M->ncols;
    mat *Ms = matrix_new(k, k);
    for (idx_t j = 0; j < k; ++j)
        for (idx_t i = 0; i < k; ++i)
            Ms->d[(size_t)j * k + i] = M->d[(size_t)((idx_t)Icol->d[j]) * m + (idx_t)Irow->d[i]];
    /* W = Ms [I T] with columns scattered by Icol (k x n) */
    mat *MsT = matrix_new(k, n - k), *W = matrix_new(k, n);
    matrix_matrix_mult(Ms, T, MsT);
    for (idx_t j = 0; j < n; ++j) {
        idx_t dst = (idx_t)Icol->d[j];
        const double *src = (j < k) ? &Ms->d[(size_t)j * k] : &MsT->d[(size_t)(j - k) * k];
        memcpy(&W->d[(size_t)dst * k], src, (size_t)k * sizeof(double));
    }
    /* P = [I S]^T W with rows scattered by Irow (m x n) */
    mat *StW = matrix_new(m - k, n), *P = matrix_new(m, n);
    matrix_transpose_matrix_mult(S, W, StW);
    for (idx_t j = 0; j < n; ++j)
        for (idx_t i = 0; i < m; ++i) {
            idx_t dst = (idx_t)Irow->d[i];
            P->d[(size_t)j * m + dst] = (i < k) ? W->d[(size_t)j * k + i] : StW->d[(size_t)j * (m - k) + (i - k)];
        }
    report(M, P, "two sided ID approximation");
    matrix_delete(Ms); matrix_delete(MsT); matrix_delete(W); matrix_delete(StW); matrix_delete(P);
}
void use_id_two_sided_decomp_for_approximation(mat *M, mat *T, mat *S, vec *Icol, vec *Irow, idx_t k) {   /* composite: nested calls keep one status */
    rsvd_api_enter();
    use_id_two_sided_decomp_for_approximation_impl(M, T, S, Icol, Irow, k);
    rsvd_api_leave();
}

static void use_cur_decomp_for_approximation_impl(mat *M, mat *C, mat *U, mat *R) {
    rsvd_api_begin();
    mat *P = matrix_new(M->nrows, M->ncols);
    form_cur_product_matrix(C, U, R, P);
    report(M, P, "C U R");
    matrix_delete(P);
}
void use_cur_decomp_for_approximation(mat *M, mat *C, mat *U, mat *R) {   /* composite: nested calls keep one status */
    rsvd_api_enter();
    use_cur_decomp_for_approximation_impl(M, C, U, R);
    rsvd_api_leave();
}

/* M(:, I) ~ Qk Rk  (RRA:2351-2384).  Deliberate deviation: the reference also frees the CALLER's Qk and Rk here (RRA:2378) and
 * its own driver 4 frees them again at exit (driver_multi_core_mkl4.c:139, a double free); the inputs are left alone. */
static void use_pivoted_QR_decomp_for_approximation_impl(mat *M, mat *Qk, mat *Rk, vec *I) {
    rsvd_api_begin();
    idx_t m = M->nrows, n = M->ncols;
    mat *QR = matrix_new(m, n), *P = matrix_new(m, n);
    matrix_matrix_mult(Qk, Rk, QR);
    for (idx_t j = 0; j < n; ++j) memcpy(&P->d[(size_t)((idx_t)I->d[j]) * m], &QR->d[(size_t)j * m], (size_t)m * sizeof(double));
    g_last_percent_error = get_percent_error_between_two_mats(M, P);
    printf("percent_error between M and QkRkPt = %f\n", g_last_percent_error);
    matrix_delete(QR); matrix_delete(P);
}
void use_pivoted_QR_decomp_for_approximation(mat *M, mat *Qk, mat *Rk, vec *I) {   /* composite: nested calls keep one status */
    rsvd_api_enter();
    use_pivoted_QR_decomp_for_approximation_impl(M, Qk, Rk, I);
    rsvd_api_leave();
}

/* ---- deterministic SVD baseline (RRA:7-69) ---------------------------------------------------------------------------
 * Tolerance mode keeps the reference's integer abs(): `abs(sval) < TOL` truncates sval to int first (RRA:49), so the rank
 * is the first i with |trunc(sigma_i)| < TOL. */
void low_rank_svd_decomp_fixed_rank_or_prec(mat *M, idx_t k, double TOL, idx_t *frank, mat **U, mat **S, mat **V) {
    rsvd_api_begin();
    idx_t m = M->nrows, n = M->ncols, r = min(m, n);
    int tolMode = (k <= 0);
    if (tolMode) k = r;
    if (k > r) { rsvd_api_warning("low_rank_svd_decomp_fixed_rank_or_prec: k = %lld exceeds min(m,n) = %lld; clamped", (long long)k, (long long)r); k = r; }
    double *dA = rsvd_upload(M->d, (size_t)m * (size_t)n);
    double *dU = rsvd_b200_dev_alloc((rsvd_i64)m * r + 1), *dS = rsvd_b200_dev_alloc(r + 1), *dV = rsvd_b200_dev_alloc((rsvd_i64)n * r + 1);
    if (dA && dU && dS && dV) rsvd_b200_svd_full_dev(dA, m, n, m, dU, m, dS, dV, n);
    rsvd_b200_dev_free(dA);
    *frank = k;
    if (tolMode && dS) {
        double *sv = (double *)malloc((size_t)(r ? r : 1) * sizeof(double));
        rsvd_download(sv, dS, (size_t)r);
        for (idx_t i = 0; i < r; ++i)
            if ((double)labs((long)sv[i]) < TOL && i < r - 1) { *frank = i + 1; break; }
        free(sv);
    }
    *U = download_mat(dU, m, *frank);
    *S = diag_from_device(dS, *frank);
    *V = download_mat(dV, n, *frank);
    rsvd_b200_dev_free(dU); rsvd_b200_dev_free(dS); rsvd_b200_dev_free(dV);
    rsvd_api_sync_error();
}

/* ---- legacy randomized SVD entry points (RRA:385-918): re-sequencings of the kernels above ------------------------- */
/* RRA:385-461 = eig(B B^T) variant without oversampling or power iterations */
void randomized_low_rank_svd1(mat *M, idx_t k, mat **U, mat **S, mat **V) {
    idx_t fr = 0;
    low_rank_svd_rand_decomp_fixed_rank(M, k, 0, 2, 1, 1, &fr, U, S, V);
}
/* RRA:465-532 = QR/SVD variant without oversampling or power iterations */
void randomized_low_rank_svd2(mat *M, idx_t k, mat **U, mat **S, mat **V) {
    idx_t fr = 0;
    low_rank_svd_rand_decomp_fixed_rank(M, k, 0, 1, 1, 1, &fr, U, S, V);
}
/* RRA:536-634 = QR/SVD variant with the (M M^T)^(q-1) M power scheme, loop j < q */
void randomized_low_rank_svd3(mat *M, idx_t k, idx_t q, idx_t s, mat **U, mat **S, mat **V) {
    idx_t fr = 0;
    low_rank_svd_rand_decomp_fixed_rank(M, k, 0, 1, q, s, &fr, U, S, V);
}

/* RRA:1425-1572 */
static int randqb_pb_device(mat *M, idx_t kstep, idx_t nstep, idx_t p, idx_t s, double **dQ_out, double **dB_out) {
    idx_t m = M->nrows, n = M->ncols, l = kstep * nstep;
    *dQ_out = NULL; *dB_out = NULL;
    if (kstep <= 0 || nstep <= 0 || s <= 0 || p < 0 || l > min(m, n)) { rsvd_api_error("randQB_pb: need kstep, nstep, s > 0, p >= 0 and kstep*nstep <= min(m,n)"); return 1; }
    double *dA = rsvd_upload(M->d, (size_t)m * (size_t)n);
    double *dQ = rsvd_b200_dev_alloc((rsvd_i64)m * l + 1), *dB = rsvd_b200_dev_alloc((rsvd_i64)l * n + 1);
    if (dA && dQ && dB) rsvd_b200_randqb_legacy_dev(dA, m, n, m, kstep, nstep, (int)p, (int)s, omega_seed(), dQ, m, dB, l);
    rsvd_b200_dev_free(dA);
    *dQ_out = dQ; *dB_out = dB;
    rsvd_api_sync_error();
    return g_api_status;
}
void randQB_pb(mat *M, idx_t kstep, idx_t nstep, idx_t p, idx_t s, mat **Q, mat **B) {
    rsvd_api_begin();
    idx_t m = M->nrows, n = M->ncols, l = max(kstep, 0) * max(nstep, 0);
    double *dQ = NULL, *dB = NULL;
    if (randqb_pb_device(M, kstep, nstep, p, s, &dQ, &dB)) { *Q = matrix_new(m, l); *B = matrix_new(l, n); }
    else { *Q = download_mat(dQ, m, l); *B = download_mat(dB, l, n); }
    rsvd_b200_dev_free(dQ); rsvd_b200_dev_free(dB);
    rsvd_api_sync_error();
}

/* RRA:1343-1421 */
void randQB_p(mat *M, idx_t k, idx_t p, mat **Q, mat **B) {
    rsvd_api_begin();
    idx_t m = M->nrows, n = M->ncols;
    if (k <= 0 || p < 0 || k > min(m, n)) { rsvd_api_error("randQB_p: need 0 < k <= min(m,n) and p >= 0"); *Q = matrix_new(m, max(k, 0)); *B = matrix_new(max(k, 0), n); return; }
    double *dA = rsvd_upload(M->d, (size_t)m * (size_t)n);
    double *dQ = rsvd_b200_dev_alloc((rsvd_i64)m * k + 1), *dB = rsvd_b200_dev_alloc((rsvd_i64)k * n + 1);
    if (dA && dQ && dB) rsvd_b200_randqb_single_dev(dA, m, n, m, k, p, omega_seed(), dQ, m, dB, k);
    rsvd_b200_dev_free(dA);
    *Q = download_mat(dQ, m, k); *B = download_mat(dB, k, n);
    rsvd_b200_dev_free(dQ); rsvd_b200_dev_free(dB);
    rsvd_api_sync_error();
}

/* RRA:638-697: randQB_pb(M, kstep, nstep, p, 1) then the eig(B B^T) tail, ascending singular values */
void randomized_low_rank_svd4(mat *M, idx_t kstep, idx_t nstep, idx_t p, mat **U, mat **S, mat **V) {
    rsvd_api_begin();
    idx_t m = M->nrows, n = M->ncols, l = max(kstep, 0) * max(nstep, 0);
    double *dQ = NULL, *dB = NULL;
    *U = NULL; *S = NULL; *V = NULL;
    if (randqb_pb_device(M, kstep, nstep, p, 1, &dQ, &dB)) { *U = matrix_new(m, l); *S = matrix_new(l, l); *V = matrix_new(n, l); }
    else {
        double *dU = rsvd_b200_dev_alloc((rsvd_i64)m * l + 1), *dS = rsvd_b200_dev_alloc(l + 1), *dV = rsvd_b200_dev_alloc((rsvd_i64)n * l + 1);
        if (dU && dS && dV) rsvd_b200_svd_from_qb_asc_dev(dQ, m, m, dB, l, n, l, dU, m, dS, dV, n);
        *U = download_mat(dU, m, l); *S = diag_from_device(dS, l); *V = download_mat(dV, n, l);
        rsvd_b200_dev_free(dU); rsvd_b200_dev_free(dS); rsvd_b200_dev_free(dV);
    }
    rsvd_b200_dev_free(dQ); rsvd_b200_dev_free(dB);
    rsvd_api_sync_error();
}

/* shared tail of the autorank variants: SVD factors from M and an orthonormal Q (RRA:717-747) or, with q > 1, from the
 * sketch Y through the power scheme first (RRA:858-907) */
static void autorank_tail(mat *M, mat *Y, mat *Q, idx_t k, idx_t q, idx_t s, mat **U, mat **S, mat **V) {
    idx_t m = M->nrows, n = M->ncols;
    if (g_api_status || k <= 0) { *U = matrix_new(m, max(k, 0)); *S = matrix_new(max(k, 0), max(k, 0)); *V = matrix_new(n, max(k, 0)); return; }
    double *dA = rsvd_upload(M->d, (size_t)m * (size_t)n);
    double *dP = rsvd_upload(Y ? Y->d : Q->d, (size_t)m * (size_t)k);
    double *dU = rsvd_b200_dev_alloc((rsvd_i64)m * k + 1), *dS = rsvd_b200_dev_alloc(k + 1), *dV = rsvd_b200_dev_alloc((rsvd_i64)n * k + 1);
    if (dA && dP && dU && dS && dV) {
        if (Y) rsvd_b200_svd_rand_from_sketch_dev(dA, m, n, m, dP, m, k, (int)q, (int)s, dU, m, dS, dV, n);
        else rsvd_b200_svd_from_q_dev(dA, m, n, m, dP, m, k, k, 1, dU, m, dS, dV, n);
    }
    rsvd_b200_dev_free(dA); rsvd_b200_dev_free(dP);
    *U = download_mat(dU, m, k); *S = diag_from_device(dS, k); *V = download_mat(dV, n, k);
    rsvd_b200_dev_free(dU); rsvd_b200_dev_free(dS); rsvd_b200_dev_free(dV);
    rsvd_api_sync_error();
}
/* RRA:701-759 */
static void randomized_low_rank_svd2_autorank1_impl(mat *M, double frac_of_max_rank, double TOL, mat **U, mat **S, mat **V) {
    rsvd_api_begin();
    mat *Q = NULL; idx_t k = 0;
    estimate_rank_and_buildQ(M, frac_of_max_rank, TOL, &Q, &k);
    autorank_tail(M, NULL, Q, k, 1, 1, U, S, V);
    matrix_delete(Q);
}
void randomized_low_rank_svd2_autorank1(mat *M, double frac_of_max_rank, double TOL, mat **U, mat **S, mat **V) {   /* composite: nested calls keep one status */
    rsvd_api_enter();
    randomized_low_rank_svd2_autorank1_impl(M, frac_of_max_rank, TOL, U, S, V);
    rsvd_api_leave();
}
/* RRA:763-822 */
static void randomized_low_rank_svd2_autorank2_impl(mat *M, idx_t kblocksize, double TOL, mat **U, mat **S, mat **V) {
    rsvd_api_begin();
    mat *Y = NULL, *Q = NULL; idx_t k = 0;
    estimate_rank_and_buildQ2(M, kblocksize, TOL, &Y, &Q, &k);
    autorank_tail(M, NULL, Q, k, 1, 1, U, S, V);
    matrix_delete(Y); matrix_delete(Q);
}
void randomized_low_rank_svd2_autorank2(mat *M, idx_t kblocksize, double TOL, mat **U, mat **S, mat **V) {   /* composite: nested calls keep one status */
    rsvd_api_enter();
    randomized_low_rank_svd2_autorank2_impl(M, kblocksize, TOL, U, S, V);
    rsvd_api_leave();
}
/* RRA:826-918 */
static void randomized_low_rank_svd3_autorank2_impl(mat *M, idx_t kblocksize, double TOL, idx_t q, idx_t s, mat **U, mat **S, mat **V) {
    rsvd_api_begin();
    mat *Y = NULL, *Q = NULL; idx_t k = 0;
    if (s <= 0) { rsvd_api_error("randomized_low_rank_svd3_autorank2: need s > 0"); }
    estimate_rank_and_buildQ2(M, kblocksize, TOL, &Y, &Q, &k);
    autorank_tail(M, Y, Q, k, q, s, U, S, V);
    matrix_delete(Y); matrix_delete(Q);
}
void randomized_low_rank_svd3_autorank2(mat *M, idx_t kblocksize, double TOL, idx_t q, idx_t s, mat **U, mat **S, mat **V) {   /* composite: nested calls keep one status */
    rsvd_api_enter();
    randomized_low_rank_svd3_autorank2_impl(M, kblocksize, TOL, q, s, U, S, V);
    rsvd_api_leave();
}
