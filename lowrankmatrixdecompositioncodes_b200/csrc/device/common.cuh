// common.cuh — shared declarations of the device layer (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <algorithm>

#include "../../../include/rsvd_b200.h"
#include "../../../include/rsvd_b200_rng.h"

namespace rsvd {

typedef long long i64;

// ---- error channel (the reference API is `void`; errors are reported out of band) -------------
void set_error(const char *fmt, ...);
extern int g_status;
extern int g_single_device;              // option "single_device": host-level calls ignore the worker pool (tails that need all of M on one GPU)

#define RSVD_CUDA(call)                                                                      \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            rsvd::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
        }                                                                                    \
    } while (0)

// ---- context -----------------------------------------------------------------------------------
struct Ctx {
    int device = -1;
    int sms = 148;
    cudaStream_t stream = nullptr;   // compute stream (all kernels)
    cudaStream_t copy_stream = nullptr;
    cudaStream_t aux_stream = nullptr;   // side stream for kernels that run next to the compute stream (Jacobi's live replay of V)
    cudaEvent_t aux_ev[2] = {nullptr, nullptr};
    void *block_cache = nullptr;     // runtime.cu: freed work buffers kept for the next request of the same size
    int no_block_cache = 0;          // option: every dfree goes straight back to the stream-ordered pool
    int no_live_replay = 0;          // option: rebuild V after the Jacobi kernel has finished instead of next to it
    bool inited = false;
    int *d_flag = nullptr;           // device int[8] scratch for status flags
    int *h_flag = nullptr;           // pinned mirror
    // distributed (row partition): filled by rsvd_b200_comm_init
    int rank = 0, world = 1;
    long long m_global = 0;          // total rows of the row-partitioned matrix (0 = not partitioned)
    long long row0 = 0;              // global index of this rank's first row (left sketches index Omega by global row)
    void *nccl_comm = nullptr;
    // peer mailboxes of the row/column-sharded pivoted QR (dist.cu: peer_mailbox): a block of this rank's HBM that every other
    // rank of the communicator has mapped (CUDA IPC between processes, peer access inside one process)
    struct PeerBox {
        int state = 0;                           // 0 not set up yet, 1 usable, -1 unavailable (the QR falls back to ncclAllGather)
        void *block = nullptr;                   // flags (2 x world u64, padded) followed by 2 x world records
        unsigned long long *flag = nullptr;      // this rank's flags   [parity][sender]
        double *cand = nullptr;                  // this rank's records [parity][sender][rec_max]
        unsigned long long **d_peer_flag = nullptr;   // device arrays [world]: the same two pointers on every rank
        double **d_peer_cand = nullptr;
        size_t rec_max = 0;
        unsigned long long token = 0;            // steps exchanged so far (identical on all ranks): step tokens and buffer parity
        int *err = nullptr;                      // device flag: a wait timed out
    } peer;
    // statistics
    unsigned long long launches = 0;
    int verbose = 0;
    int force_generic_gemm = 0;
    int force_qr_fallback = 0;
    int no_sketch_cluster = 0;       // option: sketch kernel without 2-CTA clusters (each CTA generates its own Omega stages)
    int qr_blocked_rows = 2048;      // inputs with at most this many rows go to the blocked pivoted QR (geqp3_blocked.cu)
    int force_unblocked_qr = 0;      // option: pivoted QR through the one-reflector-per-step kernel even for short-wide inputs
    int jacobi_transpose = 0;        // run the one-sided Jacobi on R^T (lower triangular) instead of R
    int no_chol_dataflow = 0;        // option: l x l Cholesky + inverse through the per-block launch sequence instead of cholinv.cu
    void *staging = nullptr;         // pinned staging buffers of this context (runtime.cu)
    int last_qr_path = 0;            // 1 = CholeskyQR2, 4 = shifted CholeskyQR3, 2 = TSQR-preconditioned fallback, 3 = Householder with explicit Q (singular panel)
    unsigned long long qr_fallbacks = 0;
};
Ctx &ctx();                          // the calling thread's context (primary context unless a worker bound its own)
void bind_ctx(Ctx *c);               // multi.cu: worker threads bind one context per device
int init_ctx(Ctx &c, int device);
void ensure_init();

double *dalloc(size_t n_doubles);
void *dalloc_bytes(size_t bytes);
void dfree(void *p);
void release_cached_blocks();              // return the calling context's cached work buffers to the pool
extern unsigned long long g_launches;            // all contexts together (approximate under concurrent workers: statistics only)
inline void count_launch(int n = 1) { ctx().launches += (unsigned long long)n; __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }

// RAII device buffer
struct DBuf {
    double *p = nullptr;
    size_t n = 0;
    DBuf() {}
    explicit DBuf(size_t n_) : p(dalloc(n_)), n(n_) {}
    DBuf(const DBuf &) = delete;
    DBuf &operator=(const DBuf &) = delete;
    DBuf(DBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DBuf &operator=(DBuf &&o) noexcept { if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; } return *this; }
    void alloc(size_t n_) { release(); p = dalloc(n_); n = n_; }
    void release() { if (p) dfree(p); p = nullptr; n = 0; }
    ~DBuf() { release(); }
    operator double *() const { return p; }
};

// ---- pipelined upload (hostapi.cu) ---------------------------------------------------------------
// A host matrix arriving in column chunks on the copy stream: ev[i] fires when columns [i*cw, (i+1)*cw) are in HBM.
// An empty event list means "already resident".  The first pass of every algorithm consumes the chunks as they land.
}  // namespace rsvd
#include <vector>
namespace rsvd {
struct Upload {
    i64 cw = 0;
    std::vector<cudaEvent_t> ev;
};
// h: host matrix block (rows x n, leading dimension ldh doubles) -> dA (rows x n, ld rows) on the calling context's copy stream
void upload_begin(const double *h, i64 ldh, double *dA, i64 rows, i64 n, Upload &up);
void upload_end(Upload &up);

// ---- GEMM engine (gemm.cu) ---------------------------------------------------------------------
// C(m x n) = alpha * op(A)(m x k) * op(B)(k x n) + beta * C, column-major everywhere.
// If `philox` is set, op(B)(kk, j) is not read from memory: it is normal(seed, ph_off + kk*ph_sk + j*ph_sc).
struct Gemm {
    char ta = 'N', tb = 'N';
    i64 m = 0, n = 0, k = 0;
    double alpha = 1.0, beta = 0.0;
    const double *A = nullptr; i64 lda = 0;
    const double *B = nullptr; i64 ldb = 0;
    double *C = nullptr; i64 ldc = 0;
    int batch = 1; i64 sA = 0, sB = 0, sC = 0;   // strided batch
    bool philox = false; uint64_t seed = 0; i64 ph_sk = 0, ph_sc = 0, ph_off = 0;
    bool b_upper = false;     // op(B) is upper triangular with exact zeros below the diagonal (hint: the streaming kernel stops each
                              // column tile's contraction at the diagonal; other paths ignore it)
    double *sumsq_out = nullptr;   // non-null: also return sum(C .^ 2) of the updated C (device scalar), fused into the streaming
                                   // kernel's epilogue when it serves the call, a separate pass otherwise
    bool sym_upper = false;   // C is symmetric (e.g. a Gram matrix) and only its upper triangle will be read: the streaming
                              // kernel may leave tiles entirely below the diagonal untouched (a hint; other paths ignore it)
};
void gemm(const Gemm &g);
// which kernel family served the last gemm() call: 0 generic, 1 tma
extern int g_last_gemm_path;
double dmma_peak_tflops(int iters, int use_dfma);

// ---- small kernels (small.cu) ------------------------------------------------------------------
void transpose(const double *A, i64 lda, double *B, i64 ldb, i64 m, i64 n);   // B(n x m) = A(m x n)^T
void copy_matrix(const double *A, i64 lda, double *B, i64 ldb, i64 m, i64 n);
void set_zero(double *A, size_t n);
void set_identity(double *A, i64 lda, i64 n);
void keep_upper(double *A, i64 lda, i64 n);                                  // zero strictly-lower part
void gather_cols(const double *A, i64 lda, i64 m, const double *idx, i64 k, double *B, i64 ldb);
void gather_rows(const double *A, i64 lda, i64 n, const double *idx, i64 k, double *B, i64 ldb);
void scale_cols(double *A, i64 lda, i64 m, i64 n, const double *s, int invert);
double frob_norm(const double *A, i64 lda, i64 m, i64 n);                     // syncs
void sumsq_async(const double *A, i64 lda, i64 m, i64 n, double *d_out);      // d_out[0] = sum of squares
void sum_array_async(const double *part, i64 n, double *d_out);               // d_out[0] = sum(part[0:n]) in a fixed order
void fill_normal(double *A, i64 n_entries, uint64_t seed, i64 first);
void trsm_left_upper(const double *R, i64 ldr, i64 k, double *B, i64 ldb, i64 ncols);  // B <- R^{-1} B
int lu_solve(double *A, i64 lda, i64 n, double *B, i64 ldb, i64 nrhs);        // dgesv semantics, in place
void gemv(char trans, i64 m, i64 n, double alpha, const double *A, i64 lda, const double *x, double beta, double *y);
void rank1_update(double *A, i64 lda, i64 m, i64 n, const double *q, const double *b);   // A -= q b^T
void scale_by_inv_norm(const double *y, i64 m, const double *sumsq, double *q);          // q = y / sqrt(*sumsq)
i64 mgs_rank_estimate(double *Q, i64 ldq, i64 m, i64 maxdim, double tol);                // MGS + stop rule of MVF:1366-1388; syncs

// ---- orthonormalisation (cholqr.cu) ------------------------------------------------------------
// Q (m x l, ld) <- orthonormal basis of range(Y); in place. If R != nullptr also returns R (l x l upper).
// sharded = true: Y is this rank's row block of a row-partitioned panel (Gram matrices are all-reduced);
// sharded = false: Y is replicated on every rank (no communication).
// loose = true: the caller only needs a well-conditioned basis of the same range (intermediate power-iteration steps).
void orthonormalize(double *Y, i64 ldy, i64 m, i64 l, double *R, i64 ldr, bool sharded = true, bool loose = false);
int potrf_upper(double *G, i64 ldg, i64 n);       // in-place upper Cholesky G = R^T R; returns 0 or failing column+1
void trtri_upper(const double *R, i64 ldr, i64 n, double *Rinv, i64 ldi); // Rinv = R^{-1}
// cholinv.cu: both in one dataflow kernel.  G <- R (upper, zeros below), Rinv <- R^{-1}; dminmax (host, optional) = min/max diag(R).
// 0 ok, > 0 failing column + 1, < 0 not run (caller uses the two functions above).  Synchronises the stream.
bool chol_inv_ok(i64 n);
int chol_inv_upper(double *G, i64 ldg, i64 n, double *Rinv, i64 ldi, double *dminmax);

// ---- Householder QR family (geqp3.cu) ----------------------------------------------------------
// In-place dgeqp3-compatible column-pivoted QR (R in the upper triangle, jpvt 0-based as doubles).
void geqp3(double *A, i64 lda, i64 m, i64 n, double *jpvt_out);
// geqp3_blocked.cu: dlaqps-style blocked kernel for rows <= 4096 (classic in-place layout; tau_out optional)
bool geqp3_blocked_ok(i64 m, i64 n);
void geqp3_blocked(double *A, i64 lda, i64 m, i64 n, double *jpvt_out, double *tau_out);
// pivoted QR + T = R11(0:k,0:k)^{-1} R(0:k, k:n) of an r x n_global matrix given by this rank's columns [col0, col0+nloc)
// (Y destroyed; `per` = columns per rank of the regular sharding; sharded == false: this rank holds all columns, no communication).
// I (n_global) and T (k x (n_global-k)) are produced on every rank.
void geqp3_id(double *Y, i64 ldy, i64 r, i64 nloc, i64 col0, i64 n_global, i64 per, bool sharded, i64 k, double *I, double *T, i64 ldt);
// dgeqp3 + dorgqr: A <- R (upper triangle/trapezoid), jpvt, and the explicit thin Q (m x min(m,n)) from the reflectors
void geqp3_q(double *A, i64 lda, i64 m, i64 n, double *jpvt_out, double *Q, i64 ldq);
// unpivoted Householder, R only (upper triangle of A on exit), used by the TSQR fallback
void geqrf_r(double *A, i64 lda, i64 m, i64 n);
// dgeqrf + dorgqr (m >= n): A <- explicit thin Q (LAPACK sign convention), R (n x n upper, optional) — any rank
void geqrf_q(double *A, i64 lda, i64 m, i64 n, double *R, i64 ldr);
// the reference's own partial pivoted QR (RRA:1012-1334); returns frank
i64 pqr_partial(double *A, i64 lda, i64 m, i64 n, i64 k, double tol, int zero_exact, double *I, double *Q, i64 ldq, double *R, i64 ldr);

// ---- small dense eigen/SVD (jacobi.cu) ---------------------------------------------------------
// A (n x n, overwritten) = U diag(s) V^T, s descending. U (n x n), Vt (n x n).
void jacobi_svd(double *A, i64 lda, i64 n, double *U, i64 ldu, double *s, double *Vt, i64 ldvt);
// symmetric eigendecomposition, ascending eigenvalues, eigenvectors in columns of A on exit
void jacobi_eig(double *A, i64 lda, i64 n, double *w);

// ---- collectives (dist.cu) ---------------------------------------------------------------------
void allreduce_sum(double *d, size_t count);      // no-op when world == 1
void allgather(const double *send, double *recv, size_t count);   // recv = [rank 0's count doubles | rank 1's | ...]
bool peer_mailbox(size_t rec_doubles);     // collective (every rank of the communicator calls it at the same point); true = ctx().peer usable
int nccl_unique_id(char id_out[128]);
int nccl_join(int rank, int world, const char id[128]);            // the calling thread's context joins a communicator
void nccl_leave();

// ---- single-process multi-GPU (multi.cu): one worker thread + context per device ----------------
int pool_size();                                  // devices driven by this process (1 = the calling thread's context only)
}  // namespace rsvd
#include <functional>
namespace rsvd {
void pool_run(const std::function<void(int)> &fn);   // fn(rank) on every worker (inline when pool_size() == 1); blocks

}  // namespace rsvd
