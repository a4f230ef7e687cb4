/* matrix_vector_functions.c — C host side of the primitives layer (reference: matrix_vector_functions_intel_mkl.c).
 * Compiled twice: int indices (librsvd_b200_api32.so) and, with -DRSVD_INDEX_64, int64_t indices (api64).
 *
 * mat/vec storage, the binary file formats and the element/slicing helpers are plain host C; everything that was
 * an MKL call in the reference (dgemm, dgeqrf+dorgqr, dgesvd, dsyev, dtrsm, dgesv, the VSL generator) is executed
 * by the sm_100a device layer through the C-ABI of rsvd_b200.h: upload, run, download.  There is no CPU fallback:
 * without the device layer these entry points record an error (rsvd_b200_api_status()) and leave outputs zero. */
#include "matrix_vector_functions_intel_mkl.h"
#include "rsvd_b200.h"
#include "rsvd_b200_host_util.h"

typedef RSVD_INT idx_t;

/* ---- allocation: large matrices live in pinned memory so uploads/downloads run at PCIe speed.  Pinning costs
 * ~0.5 ms/MB, so released pinned blocks are kept in a small cache (<= 1 GiB) and reused by later matrix_new calls
 * (repeated API calls then allocate their U/V outputs for free). ---------------------------------------------------- */
#include <pthread.h>
/* reference drivers call vector_new / vector_delete inside omp parallel loops: the tables below are shared state */
static pthread_mutex_t g_tab_mu = PTHREAD_MUTEX_INITIALIZER;
#define PINNED_MAX 256
static void *g_pinned[PINNED_MAX];
static size_t g_pinned_bytes[PINNED_MAX];
static int g_npinned = 0;
#define CACHE_MAX 16
static void *g_cache[CACHE_MAX];
static size_t g_cache_bytes[CACHE_MAX];
static int g_ncache = 0;
static size_t g_cache_total = 0;
static size_t pinned_threshold(void) {
    static long thr = -2;
    if (thr == -2) {
        const char *s = getenv("RSVD_B200_PINNED_MB");   /* matrices of at least this many MB are pinned; <0 disables */
        thr = s ? atol(s) : 64;
    }
    return thr < 0 ? (size_t)-1 : (size_t)thr << 20;
}
/* zero = 0: the caller overwrites the whole block (downloads of results): a recycled pinned block is handed out as it is —
 * clearing 280 MB of outputs cost 19 of the 24 ms "download" phase of an end-to-end call at BASELINE configs[1]. */
static double *host_alloc_impl(size_t n, int zero) {
    size_t bytes = n * sizeof(double);
    if (bytes >= pinned_threshold() && rsvd_b200_device_count() > 0) {   /* small allocations never touch the tables */
        void *p = NULL;
        pthread_mutex_lock(&g_tab_mu);
        if (g_npinned >= PINNED_MAX) { pthread_mutex_unlock(&g_tab_mu); return (double *)calloc(n ? n : 1, sizeof(double)); }
        for (int i = 0; i < g_ncache; ++i)
            if (g_cache_bytes[i] >= bytes && g_cache_bytes[i] <= bytes + (bytes >> 3)) {   /* reuse a cached block of ~the same size */
                p = g_cache[i];
                size_t cb = g_cache_bytes[i];
                g_cache_total -= cb;
                g_cache[i] = g_cache[g_ncache - 1]; g_cache_bytes[i] = g_cache_bytes[g_ncache - 1]; --g_ncache;
                g_pinned[g_npinned] = p; g_pinned_bytes[g_npinned++] = cb;
                pthread_mutex_unlock(&g_tab_mu);
                if (zero) memset(p, 0, bytes);
                return (double *)p;
            }
        pthread_mutex_unlock(&g_tab_mu);
        p = rsvd_b200_host_alloc(bytes);
        if (p) {
            pthread_mutex_lock(&g_tab_mu);
            if (g_npinned < PINNED_MAX) { g_pinned[g_npinned] = p; g_pinned_bytes[g_npinned++] = bytes; pthread_mutex_unlock(&g_tab_mu); return (double *)p; }
            pthread_mutex_unlock(&g_tab_mu);
            rsvd_b200_host_free(p);
        }
    }
    return (double *)calloc(n ? n : 1, sizeof(double));
}
double *rsvd_host_calloc(size_t n) { return host_alloc_impl(n, 1); }
double *rsvd_host_alloc_uninit(size_t n) { return host_alloc_impl(n, 0); }
void rsvd_host_free(double *p) {
    if (!p) return;
    if (__atomic_load_n(&g_npinned, __ATOMIC_RELAXED) > 0) {     /* a pinned block is only ever freed by the thread that holds it */
        pthread_mutex_lock(&g_tab_mu);
        for (int i = 0; i < g_npinned; ++i)
            if (g_pinned[i] == (void *)p) {
                size_t bytes = g_pinned_bytes[i];
                int keep = 0;
                g_pinned[i] = g_pinned[g_npinned - 1]; g_pinned_bytes[i] = g_pinned_bytes[g_npinned - 1]; --g_npinned;
                if (g_ncache < CACHE_MAX && g_cache_total + bytes <= ((size_t)1 << 30)) {
                    g_cache[g_ncache] = p; g_cache_bytes[g_ncache++] = bytes; g_cache_total += bytes;
                    keep = 1;
                }
                pthread_mutex_unlock(&g_tab_mu);
                if (!keep) rsvd_b200_host_free(p);
                return;
            }
        pthread_mutex_unlock(&g_tab_mu);
    }
    free(p);
}

mat *rsvd_matrix_new_uninit(idx_t nrows, idx_t ncols) {   /* internal: contents unspecified, for buffers that are fully overwritten */
    mat *M = (mat *)malloc(sizeof(mat));
    M->nrows = nrows; M->ncols = ncols;
    M->d = rsvd_host_alloc_uninit((size_t)nrows * (size_t)ncols);
    return M;
}
mat *matrix_new(idx_t nrows, idx_t ncols) {
    mat *M = (mat *)malloc(sizeof(mat));
    M->nrows = nrows; M->ncols = ncols;
    M->d = rsvd_host_calloc((size_t)nrows * (size_t)ncols);
    return M;
}
vec *vector_new(idx_t nrows) {
    vec *v = (vec *)malloc(sizeof(vec));
    v->nrows = nrows;
    v->d = rsvd_host_calloc((size_t)nrows);
    return v;
}
void matrix_delete(mat *M) { if (M) { rsvd_host_free(M->d); free(M); } }
void vector_delete(vec *v) { if (v) { rsvd_host_free(v->d); free(v); } }

#define EL(M, i, j) ((M)->d[(size_t)(j) * (size_t)(M)->nrows + (size_t)(i)])

void matrix_set_element(mat *M, idx_t i, idx_t j, double val) { EL(M, i, j) = val; }
double matrix_get_element(mat *M, idx_t i, idx_t j) { return EL(M, i, j); }
void vector_set_element(vec *v, idx_t i, double val) { v->d[i] = val; }
double vector_get_element(vec *v, idx_t i) { return v->d[i]; }

/* ---- binary I/O: header (idx_t m, idx_t n) + ROW-major doubles (MVF:77-133; 64-bit MVF64:78-135).
 * The reference issues one fread/fwrite per element; here rows are moved in blocks and transposed in cache tiles. */
mat *matrix_load_from_binary_file(char *fname) {
    FILE *fp = fopen(fname, "rb");
    if (!fp) { rsvd_api_error("matrix_load_from_binary_file: cannot open %s", fname); return NULL; }
    idx_t m = 0, n = 0;
    if (fread(&m, sizeof(idx_t), 1, fp) != 1 || fread(&n, sizeof(idx_t), 1, fp) != 1 || m < 0 || n < 0) {
        rsvd_api_error("matrix_load_from_binary_file: bad header in %s", fname);
        fclose(fp);
        return NULL;
    }
    mat *M = matrix_new(m, n);
    /* Blocks of up to 32 MB of whole rows, double-buffered: inside one parallel region a single thread reads block i+1 from
     * the file while the others (and then that thread too: dynamic schedule) scatter block i into the column-major matrix, one
     * cache line of the row-major block (8 columns of one row) at a time.  Page-cache reads (~4.6 GB/s) and the scatter
     * (~4 GB/s on 8 cores) overlap instead of adding up. */
    size_t RB = n ? (((size_t)32 << 20) / ((size_t)n * sizeof(double))) : 1;
    if (RB < 8) RB = 8;
    if (RB > (size_t)(m ? m : 1)) RB = (size_t)(m ? m : 1);
    const size_t bufn = RB * (size_t)(n ? n : 1);
    double *bufs[2] = {(double *)malloc(bufn * sizeof(double)), (double *)malloc(bufn * sizeof(double))};
    int truncated = 0;
    size_t rb0 = (size_t)m < RB ? (size_t)m : RB;
    if (m > 0 && n > 0 && fread(bufs[0], sizeof(double), rb0 * (size_t)n, fp) != rb0 * (size_t)n) truncated = 1;
    int cur = 0;
    for (size_t i0 = 0; i0 < (size_t)m && !truncated; i0 += RB, cur ^= 1) {
        const size_t rb = (size_t)m - i0 < RB ? (size_t)m - i0 : RB;
        const size_t i1 = i0 + RB;
        const size_t rb_next = i1 < (size_t)m ? ((size_t)m - i1 < RB ? (size_t)m - i1 : RB) : 0;
        const double *buf = bufs[cur];
        double *next = bufs[cur ^ 1];
        int bad = 0;
        #pragma omp parallel
        {
            #pragma omp single nowait
            {
                if (rb_next && fread(next, sizeof(double), rb_next * (size_t)n, fp) != rb_next * (size_t)n) bad = 1;
            }
            #pragma omp for schedule(dynamic, 64)
            for (long long j0 = 0; j0 < (long long)n; j0 += 8) {
                size_t j1 = (size_t)j0 + 8 < (size_t)n ? (size_t)j0 + 8 : (size_t)n;
                for (size_t r = 0; r < rb; ++r)
                    for (size_t j = (size_t)j0; j < j1; ++j) M->d[j * (size_t)m + i0 + r] = buf[r * (size_t)n + j];
            }
        }
        if (bad) truncated = 1;     /* the block just scattered was complete; the next one is not */
    }
    if (truncated) rsvd_api_error("matrix_load_from_binary_file: %s is truncated", fname);
    free(bufs[0]); free(bufs[1]);
    fclose(fp);
    return M;
}

void matrix_write_to_binary_file(mat *M, char *fname) {
    FILE *fp = fopen(fname, "wb");
    if (!fp) { rsvd_api_error("matrix_write_to_binary_file: cannot open %s", fname); return; }
    idx_t m = M->nrows, n = M->ncols;
    fwrite(&m, sizeof(idx_t), 1, fp);
    fwrite(&n, sizeof(idx_t), 1, fp);
    size_t RB = n ? (((size_t)32 << 20) / ((size_t)n * sizeof(double))) : 1;
    if (RB < 8) RB = 8;
    if (RB > (size_t)(m ? m : 1)) RB = (size_t)(m ? m : 1);
    /* double-buffered like the loader: one thread writes block i-1 while the others gather block i */
    const size_t bufn = RB * (size_t)(n ? n : 1);
    double *bufs[2] = {(double *)malloc(bufn * sizeof(double)), (double *)malloc(bufn * sizeof(double))};
    int cur = 0, short_write = 0;
    size_t pending = 0;                                      /* doubles of the previous block still to be written (in bufs[cur ^ 1]) */
    for (size_t i0 = 0; i0 < (size_t)m || pending; i0 += RB, cur ^= 1) {
        const size_t rb = i0 < (size_t)m ? ((size_t)m - i0 < RB ? (size_t)m - i0 : RB) : 0;
        double *buf = bufs[cur];
        const double *prev = bufs[cur ^ 1];
        const size_t npend = pending;
        int bad = 0;
        #pragma omp parallel
        {
            #pragma omp single nowait
            {
                if (npend && fwrite(prev, sizeof(double), npend, fp) != npend) bad = 1;
            }
            #pragma omp for schedule(dynamic, 64)
            for (long long j0 = 0; j0 < (long long)n; j0 += 8) {
                size_t j1 = (size_t)j0 + 8 < (size_t)n ? (size_t)j0 + 8 : (size_t)n;
                for (size_t r = 0; r < rb; ++r)
                    for (size_t j = (size_t)j0; j < j1; ++j) buf[r * (size_t)n + j] = M->d[j * (size_t)m + i0 + r];
            }
        }
        if (bad) { short_write = 1; break; }
        pending = rb * (size_t)n;
    }
    if (short_write) rsvd_api_error("matrix_write_to_binary_file: short write to %s", fname);
    free(bufs[0]); free(bufs[1]);
    fclose(fp);
}

void matrix_print(mat *M) {
    for (idx_t i = 0; i < M->nrows; ++i) {
        for (idx_t j = 0; j < M->ncols; ++j) printf("%f  ", EL(M, i, j));
        printf("\n");
    }
}
void vector_print(vec *v) {
    for (idx_t i = 0; i < v->nrows; ++i) printf("%f\n", v->d[i]);
}

/* ---- elementwise helpers ----------------------------------------------------------------------------------------- */
void vector_set_data(vec *v, double *data) { memcpy(v->d, data, (size_t)v->nrows * sizeof(double)); }
void vector_scale(vec *v, double a) { for (idx_t i = 0; i < v->nrows; ++i) v->d[i] *= a; }
void matrix_scale(mat *M, double a) {
    size_t N = (size_t)M->nrows * (size_t)M->ncols;
    for (size_t i = 0; i < N; ++i) M->d[i] *= a;
}
double vector_get2norm(vec *v) {
    double s = 0;
    for (idx_t i = 0; i < v->nrows; ++i) s += v->d[i] * v->d[i];
    return sqrt(s);
}
/* MVF:314-340 stage every entry through an `int` (values are truncated before they are compared and returned); here the true
 * minimum / maximum of the doubles is returned — identical for the integer-valued index vectors these helpers are used on. */
void vector_get_min_element(vec *v, idx_t *minindex, double *minval) {
    idx_t bi = 0; double b = v->d[0];
    for (idx_t i = 1; i < v->nrows; ++i) if (v->d[i] < b) { b = v->d[i]; bi = i; }
    *minindex = bi; *minval = b;
}
void vector_get_max_element(vec *v, idx_t *maxindex, double *maxval) {
    idx_t bi = 0; double b = v->d[0];
    for (idx_t i = 1; i < v->nrows; ++i) if (v->d[i] > b) { b = v->d[i]; bi = i; }
    *maxindex = bi; *maxval = b;
}
void vector_copy(vec *d, vec *s) { memcpy(d->d, s->d, (size_t)s->nrows * sizeof(double)); }
void matrix_copy(mat *D, mat *S) { memcpy(D->d, S->d, (size_t)S->nrows * (size_t)S->ncols * sizeof(double)); }
void matrix_hard_threshold(mat *M, double TOL) {
    size_t N = (size_t)M->nrows * (size_t)M->ncols;
    for (size_t i = 0; i < N; ++i) if (fabs(M->d[i]) < TOL) M->d[i] = 0.0;
}
void matrix_build_transpose(mat *Mt, mat *M) {
    const size_t m = (size_t)M->nrows, n = (size_t)M->ncols, TB = 32;
    for (size_t j0 = 0; j0 < n; j0 += TB)
        for (size_t i0 = 0; i0 < m; i0 += TB)
            for (size_t j = j0; j < n && j < j0 + TB; ++j)
                for (size_t i = i0; i < m && i < i0 + TB; ++i) Mt->d[i * n + j] = M->d[j * m + i];
}
void vector_sub(vec *a, vec *b) { for (idx_t i = 0; i < a->nrows; ++i) a->d[i] -= b->d[i]; }
void matrix_sub(mat *A, mat *B) {
    size_t N = (size_t)A->nrows * (size_t)A->ncols;
    for (size_t i = 0; i < N; ++i) A->d[i] -= B->d[i];
}
void matrix_sub_column_times_row_vector(mat *A, vec *u, vec *v) {
    for (idx_t j = 0; j < A->ncols; ++j)
        for (idx_t i = 0; i < A->nrows; ++i) EL(A, i, j) -= u->d[i] * v->d[j];
}
double get_matrix_frobenius_norm(mat *M) {
    size_t N = (size_t)M->nrows * (size_t)M->ncols;
    double s = 0;
    for (size_t i = 0; i < N; ++i) s += M->d[i] * M->d[i];
    return sqrt(s);
}
double get_matrix_max_abs_element(mat *M) {
    size_t N = (size_t)M->nrows * (size_t)M->ncols;
    double b = 0;
    for (size_t i = 0; i < N; ++i) if (fabs(M->d[i]) > b) b = fabs(M->d[i]);
    return b;
}
double vector_dot_product(vec *u, vec *v) {
    double s = 0;
    for (idx_t i = 0; i < u->nrows; ++i) s += u->d[i] * v->d[i];
    return s;
}
double get_matrix_column_norm_squared(mat *M, idx_t c) {
    double s = 0;
    for (idx_t i = 0; i < M->nrows; ++i) s += EL(M, i, c) * EL(M, i, c);
    return s;
}
double matrix_getmaxcolnorm(mat *M) {
    double b = 0;
    for (idx_t j = 0; j < M->ncols; ++j) { double s = sqrt(get_matrix_column_norm_squared(M, j)); if (s > b) b = s; }
    return b;
}
/* MVF:445-454: despite the name these are the SQUARED column norms (the reference's own pivoted QR downdates them as such) */
void compute_matrix_column_norms(mat *M, vec *norms) {
    for (idx_t j = 0; j < M->ncols; ++j) norms->d[j] = get_matrix_column_norm_squared(M, j);
}
double get_percent_error_between_two_mats(mat *A, mat *B) {
    size_t N = (size_t)A->nrows * (size_t)A->ncols;
    double sa = 0, sd = 0;
    for (size_t i = 0; i < N; ++i) { double d = A->d[i] - B->d[i]; sa += A->d[i] * A->d[i]; sd += d * d; }
    return 100.0 * sqrt(sd) / sqrt(sa);
}

/* ---- device round trips ------------------------------------------------------------------------------------------ */
double *rsvd_upload(const double *h, size_t n) {
    double *d = rsvd_b200_dev_alloc((rsvd_i64)(n ? n : 1));
    if (!d) { rsvd_api_sync_error(); return NULL; }
    if (n && rsvd_b200_h2d(d, h, (rsvd_i64)n)) rsvd_api_sync_error();
    return d;
}
void rsvd_download(double *h, const double *d, size_t n) {
    if (n && rsvd_b200_d2h(h, d, (rsvd_i64)n)) rsvd_api_sync_error();
}

void initialize_random_matrix(mat *M) {
    rsvd_api_begin();   /* a failure of an EARLIER top-level call must not turn this one into a silent no-op */
    size_t N = (size_t)M->nrows * (size_t)M->ncols;
    double *d = rsvd_b200_dev_alloc((rsvd_i64)(N ? N : 1));
    if (!d) { rsvd_api_sync_error(); return; }
    rsvd_b200_fill_normal(d, (rsvd_i64)N, (uint64_t)rsvd_b200_get_option("seed"), 0);
    rsvd_download(M->d, d, N);
    rsvd_b200_dev_free(d);
    rsvd_api_sync_error();
}

static void host_gemm(char ta, char tb, mat *A, mat *B, mat *C) {
    rsvd_api_begin();   /* a failure of an EARLIER top-level call must not turn this one into a silent no-op */
    idx_t m = C->nrows, n = C->ncols, k = (ta == 'N') ? A->ncols : A->nrows;
    double *dA = rsvd_upload(A->d, (size_t)A->nrows * (size_t)A->ncols);
    double *dB = rsvd_upload(B->d, (size_t)B->nrows * (size_t)B->ncols);
    double *dC = rsvd_b200_dev_alloc((rsvd_i64)((size_t)m * (size_t)n + 1));
    if (dA && dB && dC) {
        rsvd_b200_gemm(ta, tb, m, n, k, 1.0, dA, A->nrows, dB, B->nrows, 0.0, dC, m);
        rsvd_download(C->d, dC, (size_t)m * (size_t)n);
    }
    rsvd_b200_dev_free(dA); rsvd_b200_dev_free(dB); rsvd_b200_dev_free(dC);
    rsvd_api_sync_error();
}
void matrix_matrix_mult(mat *A, mat *B, mat *C) { host_gemm('N', 'N', A, B, C); }
void matrix_transpose_matrix_mult(mat *A, mat *B, mat *C) { host_gemm('T', 'N', A, B, C); }
void matrix_matrix_transpose_mult(mat *A, mat *B, mat *C) { host_gemm('N', 'T', A, B, C); }

void matrix_vector_mult(mat *M, vec *x, vec *y) {
    mat X = {x->nrows, 1, x->d}, Y = {y->nrows, 1, y->d};
    host_gemm('N', 'N', M, &X, &Y);
}
void matrix_transpose_vector_mult(mat *M, vec *x, vec *y) {
    mat X = {x->nrows, 1, x->d}, Y = {y->nrows, 1, y->d};
    host_gemm('T', 'N', M, &X, &Y);
}

/* ---- rows / columns ---------------------------------------------------------------------------------------------- */
void matrix_get_col(mat *M, idx_t j, vec *c) { memcpy(c->d, &EL(M, 0, j), (size_t)M->nrows * sizeof(double)); }
void matrix_set_col(mat *M, idx_t j, vec *c) { memcpy(&EL(M, 0, j), c->d, (size_t)M->nrows * sizeof(double)); }
void matrix_get_row(mat *M, idx_t i, vec *r) { for (idx_t j = 0; j < M->ncols; ++j) r->d[j] = EL(M, i, j); }
void matrix_set_row(mat *M, idx_t i, vec *r) { for (idx_t j = 0; j < M->ncols; ++j) EL(M, i, j) = r->d[j]; }
void matrix_get_selected_columns(mat *M, idx_t *inds, mat *Mc) {
    for (idx_t j = 0; j < Mc->ncols; ++j) memcpy(&EL(Mc, 0, j), &EL(M, 0, inds[j]), (size_t)M->nrows * sizeof(double));
}
void matrix_set_selected_columns(mat *M, idx_t *inds, mat *Mc) {
    for (idx_t j = 0; j < Mc->ncols; ++j) memcpy(&EL(M, 0, inds[j]), &EL(Mc, 0, j), (size_t)M->nrows * sizeof(double));
}
void matrix_get_selected_rows(mat *M, idx_t *inds, mat *Mr) {
    for (idx_t j = 0; j < M->ncols; ++j)
        for (idx_t i = 0; i < Mr->nrows; ++i) EL(Mr, i, j) = EL(M, inds[i], j);
}
void matrix_set_selected_rows(mat *M, idx_t *inds, mat *Mr) {
    for (idx_t j = 0; j < M->ncols; ++j)
        for (idx_t i = 0; i < Mr->nrows; ++i) EL(M, inds[i], j) = EL(Mr, i, j);
}
void matrix_copy_symmetric(mat *S, mat *M) {
    for (idx_t j = 0; j < M->ncols; ++j)
        for (idx_t i = 0; i <= j && i < M->nrows; ++i) EL(S, i, j) = EL(M, i, j);
}
void matrix_keep_only_upper_triangular(mat *M) {
    for (idx_t j = 0; j < M->ncols; ++j)
        for (idx_t i = j + 1; i < M->nrows; ++i) EL(M, i, j) = 0.0;
}
void initialize_diagonal_matrix(mat *D, vec *data) { for (idx_t i = 0; i < D->nrows; ++i) EL(D, i, i) = data->d[i]; }
void initialize_identity_matrix(mat *D) {
    matrix_scale(D, 0.0);
    for (idx_t i = 0; i < D->nrows; ++i) EL(D, i, i) = 1.0;
}
void invert_diagonal_matrix(mat *Dinv, mat *D) { for (idx_t i = 0; i < D->nrows; ++i) EL(Dinv, i, i) = 1.0 / EL(D, i, i); }
void invert_upper_triangular_matrix(mat *Minv) {
    /* X = R^{-1} by solving R X = I on the device */
    idx_t n = Minv->nrows;
    mat *I = matrix_new(n, n);
    initialize_identity_matrix(I);
    mat *X = matrix_new(n, n);
    upper_triangular_system_solve(Minv, I, X, 1);
    matrix_copy(Minv, X);
    matrix_delete(I); matrix_delete(X);
}

/* ---- slicing ----------------------------------------------------------------------------------------------------- */
void fill_vector_from_row_list(vec *input, vec *inds, vec *output) {
    for (idx_t i = 0; i < inds->nrows; ++i) output->d[i] = input->d[(idx_t)inds->d[i]];
}
static void copy_block(mat *D, mat *S, idx_t r0, idx_t c0) {   /* D = S(r0:r0+D.nrows, c0:c0+D.ncols) */
    for (idx_t j = 0; j < D->ncols; ++j) memcpy(&EL(D, 0, j), &EL(S, r0, c0 + j), (size_t)D->nrows * sizeof(double));
}
void matrix_copy_first_rows(mat *M_out, mat *M) { copy_block(M_out, M, 0, 0); }
void matrix_copy_first_columns(mat *M_out, mat *M) { copy_block(M_out, M, 0, 0); }
void matrix_copy_first_columns_with_param(mat *D, mat *S, idx_t num_columns) {
    for (idx_t j = 0; j < num_columns; ++j) memcpy(&EL(D, 0, j), &EL(S, 0, j), (size_t)S->nrows * sizeof(double));
}
void matrix_copy_first_k_rows_and_columns(mat *M_out, mat *M) { copy_block(M_out, M, 0, 0); }
void matrix_copy_all_rows_and_last_columns_from_indexk(mat *M_out, mat *M, idx_t k) { copy_block(M_out, M, 0, k); }
void fill_matrix_from_first_rows(mat *M, idx_t k, mat *M_k) { (void)k; copy_block(M_k, M, 0, 0); }
void fill_matrix_from_last_rows(mat *M, idx_t k, mat *M_k) { copy_block(M_k, M, M->nrows - k, 0); }
void fill_matrix_from_first_columns(mat *M, idx_t k, mat *M_k) { (void)k; copy_block(M_k, M, 0, 0); }
void fill_matrix_from_last_columns(mat *M, idx_t k, mat *M_k) { copy_block(M_k, M, 0, M->ncols - k); }
void fill_matrix_from_last_columns_from_specified_one(mat *M, idx_t k, mat *M_k) { copy_block(M_k, M, 0, k); }
void fill_matrix_from_lower_right_corner(mat *M, idx_t k, mat *M_out) { copy_block(M_out, M, M->nrows - k, M->ncols - k); }
void fill_matrix_from_first_columns_from_list(mat *M, vec *I, idx_t k, mat *M_k) {
    for (idx_t j = 0; j < k; ++j) memcpy(&EL(M_k, 0, j), &EL(M, 0, (idx_t)I->d[j]), (size_t)M->nrows * sizeof(double));
}
void fill_matrix_from_first_rows_from_list(mat *M, vec *I, idx_t k, mat *M_k) {
    for (idx_t j = 0; j < M->ncols; ++j)
        for (idx_t i = 0; i < k; ++i) EL(M_k, i, j) = EL(M, (idx_t)I->d[i], j);
}
/* MVF:1046-1060: the columns listed in I[k .. ncols-1] (ncols - k of them), i.e. everything AFTER the first k of the list */
void fill_matrix_from_last_columns_from_list(mat *M, vec *I, idx_t k, mat *M_k) {
    idx_t n = M->ncols;
    for (idx_t j = k; j < n; ++j)
        memcpy(&EL(M_k, 0, j - k), &EL(M, 0, (idx_t)I->d[j]), (size_t)M->nrows * sizeof(double));
}
static void resize_to(mat **M, idx_t r0, idx_t c0, idx_t nr, idx_t nc) {
    mat *R = matrix_new(nr, nc);
    copy_block(R, *M, r0, c0);
    matrix_delete(*M);
    *M = R;
}
void resize_matrix_by_columns(mat **M, idx_t k) { resize_to(M, 0, 0, (*M)->nrows, k); }
void resize_matrix_by_columns_from_end(mat **M, idx_t k) { resize_to(M, 0, (*M)->ncols - k, (*M)->nrows, k); }
void resize_matrix_by_rows(mat **M, idx_t k) { resize_to(M, 0, 0, k, (*M)->ncols); }
void resize_matrix_by_rows_from_end(mat **M, idx_t k) { resize_to(M, (*M)->nrows - k, 0, k, (*M)->ncols); }
void append_matrices_horizontally(mat *A, mat *B, mat *C) {
    for (idx_t j = 0; j < A->ncols; ++j) memcpy(&EL(C, 0, j), &EL(A, 0, j), (size_t)A->nrows * sizeof(double));
    for (idx_t j = 0; j < B->ncols; ++j) memcpy(&EL(C, 0, A->ncols + j), &EL(B, 0, j), (size_t)B->nrows * sizeof(double));
}
void append_matrices_vertically(mat *A, mat *B, mat *C) {
    for (idx_t j = 0; j < C->ncols; ++j) {
        memcpy(&EL(C, 0, j), &EL(A, 0, j), (size_t)A->nrows * sizeof(double));
        memcpy(&EL(C, A->nrows, j), &EL(B, 0, j), (size_t)B->nrows * sizeof(double));
    }
}
void vector_build_rewrapped(vec *Iinv, vec *I) {
    for (idx_t i = 0; i < I->nrows; ++i) Iinv->d[(idx_t)I->d[i]] = (double)i;
}

/* ---- factorizations on the device -------------------------------------------------------------------------------- */
void compute_evals_and_evecs_of_symm_matrix(mat *S, vec *evals) {
    rsvd_api_begin();   /* a failure of an EARLIER top-level call must not turn this one into a silent no-op */
    idx_t n = S->nrows;
    /* dsyev 'U' reads only the upper triangle (callers fill it with matrix_copy_symmetric): mirror it first */
    for (idx_t j = 0; j < n; ++j)
        for (idx_t i = j + 1; i < n; ++i) EL(S, i, j) = EL(S, j, i);
    double *dS = rsvd_upload(S->d, (size_t)n * (size_t)n);
    double *dw = rsvd_b200_dev_alloc(n + 1);
    if (dS && dw) {
        rsvd_b200_eig_small(dS, n, n, dw);
        rsvd_download(S->d, dS, (size_t)n * (size_t)n);
        rsvd_download(evals->d, dw, (size_t)n);
    }
    rsvd_b200_dev_free(dS); rsvd_b200_dev_free(dw);
    rsvd_api_sync_error();
}

static void host_qr(mat *M, mat *Q, mat *R) {
    rsvd_api_begin();   /* a failure of an EARLIER top-level call must not turn this one into a silent no-op */
    idx_t m = M->nrows, n = M->ncols;
    if (m < n) { rsvd_api_error("QR of a %lld x %lld matrix: only tall panels (m >= n) are supported", (long long)m, (long long)n); return; }
    double *dY = rsvd_upload(M->d, (size_t)m * (size_t)n);
    double *dR = R ? rsvd_b200_dev_alloc((rsvd_i64)n * n + 1) : NULL;
    if (dY && (!R || dR)) {
        rsvd_b200_orthonormalize(dY, m, m, n, dR, n);
        rsvd_download(Q->d, dY, (size_t)m * (size_t)n);
        if (R) rsvd_download(R->d, dR, (size_t)n * (size_t)n);
    }
    rsvd_b200_dev_free(dY); rsvd_b200_dev_free(dR);
    rsvd_api_sync_error();
}
void compact_QR_factorization(mat *M, mat *Q, mat *R) { host_qr(M, Q, R); }
void QR_factorization_getQ(mat *M, mat *Q) { host_qr(M, Q, NULL); }

/* MVF:1270-1284 (dgesvd 'S','S'): U m x r, S r x r diagonal, Vt r x n, r = min(m,n).  Square inputs (the hot path's l x l
 * Rhat) go straight to the Jacobi kernel; other shapes through the QR-preconditioned full SVD. */
void singular_value_decomposition(mat *M, mat *U, mat *S, mat *Vt) {
    rsvd_api_begin();   /* a failure of an EARLIER top-level call must not turn this one into a silent no-op */
    idx_t m = M->nrows, n = M->ncols, r = min(m, n);
    double *dA = rsvd_upload(M->d, (size_t)m * (size_t)n);
    double *dU = rsvd_b200_dev_alloc((rsvd_i64)m * r + 1), *dV = rsvd_b200_dev_alloc((rsvd_i64)n * r + 1), *ds = rsvd_b200_dev_alloc(r + 1);
    if (dA && dU && dV && ds) {
        if (m == n) {
            rsvd_b200_svd_small(dA, n, n, dU, n, ds, dV, n);                  /* dV holds V^T */
            rsvd_download(Vt->d, dV, (size_t)n * (size_t)n);
        } else {
            rsvd_b200_svd_full_dev(dA, m, n, m, dU, m, ds, dV, n);            /* dV holds V (n x r) */
            double *dVt = rsvd_b200_dev_alloc((rsvd_i64)n * r + 1);
            if (dVt) { rsvd_b200_transpose(dV, n, dVt, r, n, r); rsvd_download(Vt->d, dVt, (size_t)n * (size_t)r); }
            rsvd_b200_dev_free(dVt);
        }
        rsvd_download(U->d, dU, (size_t)m * (size_t)r);
        double *s = (double *)malloc((size_t)(r ? r : 1) * sizeof(double));
        rsvd_download(s, ds, (size_t)r);
        for (idx_t i = 0; i < r; ++i) EL(S, i, i) = s[i];
        free(s);
    }
    rsvd_b200_dev_free(dA); rsvd_b200_dev_free(dU); rsvd_b200_dev_free(dV); rsvd_b200_dev_free(ds);
    rsvd_api_sync_error();
}

void form_svd_product_matrix(mat *U, mat *S, mat *V, mat *P) {
    rsvd_api_enter();
    mat *SVt = matrix_new(S->nrows, P->ncols);
    matrix_matrix_transpose_mult(S, V, SVt);
    matrix_matrix_mult(U, SVt, P);
    matrix_delete(SVt);
    rsvd_api_leave();
}
void form_cur_product_matrix(mat *C, mat *U, mat *R, mat *P) {
    rsvd_api_enter();
    mat *CU = matrix_new(P->nrows, U->nrows);
    matrix_matrix_mult(C, U, CU);
    matrix_matrix_mult(CU, R, P);
    matrix_delete(CU);
    rsvd_api_leave();
}

void upper_triangular_system_solve(mat *A, mat *B, mat *X, int solve_type) {
    rsvd_api_begin();   /* a failure of an EARLIER top-level call must not turn this one into a silent no-op */
    (void)solve_type;   /* the reference's variants 1-4 all compute A^{-1} B; one device solver serves them */
    idx_t k = A->nrows, nc = B->ncols;
    double *dA = rsvd_upload(A->d, (size_t)k * (size_t)k);
    double *dB = rsvd_upload(B->d, (size_t)k * (size_t)nc);
    if (dA && dB) {
        rsvd_b200_trsm_left_upper(dA, k, k, dB, k, nc);
        rsvd_download(X->d, dB, (size_t)k * (size_t)nc);
    }
    rsvd_b200_dev_free(dA); rsvd_b200_dev_free(dB);
    rsvd_api_sync_error();
}

void square_matrix_system_solve(mat *A, mat *X, mat *B) {
    rsvd_api_begin();   /* a failure of an EARLIER top-level call must not turn this one into a silent no-op */
    idx_t n = A->nrows, nc = B->ncols;
    double *dA = rsvd_upload(A->d, (size_t)n * (size_t)n);
    double *dB = rsvd_upload(B->d, (size_t)n * (size_t)nc);
    if (dA && dB) {
        rsvd_b200_lu_solve(dA, n, n, dB, n, nc);
        rsvd_download(B->d, dB, (size_t)n * (size_t)nc);   /* dgesv overwrites B with the solution (MVF:1528-1529) */
        matrix_copy(X, B);
    }
    rsvd_b200_dev_free(dA); rsvd_b200_dev_free(dB);
    rsvd_api_sync_error();
}

/* ---- legacy range-finder helpers (MVF:795-841, 1339-1467) ------------------------------------------ */
/* p = (v.u / ||u||^2) u   (MVF:795-802) */
void project_vector(vec *v, vec *u, vec *p) {
    double nu = vector_get2norm(u);
    double c = vector_dot_product(v, u) / (nu * nu);
    vector_copy(p, u);
    vector_scale(p, c);
}

/* MVF:817-841: two passes of modified Gram-Schmidt = the QR factor with positive diagonal R, which is what the
 * CholeskyQR2 device kernel returns. */
void build_orthonormal_basis_from_mat(mat *A, mat *Q) { host_qr(A, Q, NULL); }

/* MVF:1339-1400 */
void estimate_rank_and_buildQ(mat *M, double frac_of_max_rank, double TOL, mat **Q, idx_t *good_rank) {
    rsvd_api_begin();   /* a failure of an EARLIER top-level call must not turn this one into a silent no-op */
    idx_t m = M->nrows, n = M->ncols;
    idx_t maxdim = (idx_t)llround((double)min(m, n) * frac_of_max_rank);
    *Q = NULL; *good_rank = 0;
    if (maxdim <= 0 || maxdim > min(m, n)) { rsvd_api_error("estimate_rank_and_buildQ: frac_of_max_rank must give 0 < maxdim <= min(m,n)"); *Q = matrix_new(m, 0); return; }
    double *dA = rsvd_upload(M->d, (size_t)m * (size_t)n);
    double *dQ = rsvd_b200_dev_alloc((rsvd_i64)m * maxdim + 1);
    rsvd_i64 rank = 0;
    if (dA && dQ) rsvd_b200_estimate_rank1_dev(dA, m, n, m, maxdim, TOL, (uint64_t)rsvd_b200_get_option("seed"), dQ, m, &rank);
    rsvd_b200_dev_free(dA);
    *good_rank = (idx_t)rank;
    *Q = matrix_new(m, *good_rank);
    if (dQ) rsvd_download((*Q)->d, dQ, (size_t)m * (size_t)*good_rank);
    rsvd_b200_dev_free(dQ);
    rsvd_api_sync_error();
}

/* MVF:1404-1467 */
void estimate_rank_and_buildQ2(mat *M, idx_t kblock, double TOL, mat **Y, mat **Q, idx_t *good_rank) {
    rsvd_api_begin();   /* a failure of an EARLIER top-level call must not turn this one into a silent no-op */
    idx_t m = M->nrows, n = M->ncols, r = min(m, n);
    *Y = NULL; *Q = NULL; *good_rank = 0;
    if (kblock <= 0 || kblock > r) { rsvd_api_error("estimate_rank_and_buildQ2: need 0 < kblock <= min(m,n)"); *Y = matrix_new(m, 0); *Q = matrix_new(m, 0); return; }
    idx_t cap = (r / kblock) * kblock;
    double *dA = rsvd_upload(M->d, (size_t)m * (size_t)n);
    double *dY = rsvd_b200_dev_alloc((rsvd_i64)m * cap + 1), *dQ = rsvd_b200_dev_alloc((rsvd_i64)m * cap + 1);
    rsvd_i64 rank = 0;
    if (dA && dY && dQ)
        rsvd_b200_estimate_rank2_dev(dA, m, n, m, kblock, TOL, (uint64_t)rsvd_b200_get_option("seed"), dY, m, dQ, m, cap, &rank);
    rsvd_b200_dev_free(dA);
    *good_rank = (idx_t)rank;
    *Y = matrix_new(m, *good_rank); *Q = matrix_new(m, *good_rank);
    if (dY) rsvd_download((*Y)->d, dY, (size_t)m * (size_t)*good_rank);
    if (dQ) rsvd_download((*Q)->d, dQ, (size_t)m * (size_t)*good_rank);
    rsvd_b200_dev_free(dY); rsvd_b200_dev_free(dQ);
    rsvd_api_sync_error();
}

double get_seconds_frac(struct timeval t0, struct timeval t1) {
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-6 * (double)(t1.tv_usec - t0.tv_usec);
}
