// pipeline.cu — the device-resident algorithms: randomized SVD, blocked randQB with its on-device tolerance
// test, randomized ID / two-sided ID / CUR.  A and every intermediate stay in HBM; the host only sequences
// kernels and reads back 4-byte status flags.  Control flow follows the reference line by line (citations are
// rank_revealing_algorithms_intel_mkl.c = RRA) so that results agree with it for the same Omega.
//
// Row partition (world > 1): `A` is this rank's block of rows; m is the local row count.  Products that contract
// over rows (A^T * X) are all-reduced; n x l panels are replicated, m x l panels are row-sharded.
#include "pipeline.cuh"
#include <vector>

namespace rsvd {

namespace {

// verbose == 2: synchronising phase timer; verbose == 3: non-synchronising (CUDA events, printed at the end of the call)
struct Phase {
    double t0 = 0; int mode;
    std::vector<std::pair<const char *, cudaEvent_t>> ev;
    Phase() : mode(ctx().verbose) {
        if (mode == 2) { cudaStreamSynchronize(ctx().stream); t0 = now(); }
        if (mode == 3) mark("start");
    }
    static double now() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
    void mark(const char *what) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, ctx().stream); ev.push_back({what, e}); }
    void lap(const char *what) {
        if (mode == 3) { mark(what); return; }
        if (mode != 2) return;
        cudaStreamSynchronize(ctx().stream);
        double t1 = now();
        fprintf(stderr, "[rsvd_b200]   %-28s %8.3f ms\n", what, (t1 - t0) * 1e3);
        t0 = t1;
    }
    ~Phase() {
        if (mode != 3 || ev.empty()) return;
        cudaStreamSynchronize(ctx().stream);
        char line[1024]; int n = snprintf(line, sizeof(line), "[rsvd_b200 r%d]", ctx().rank);
        for (size_t i = 1; i < ev.size(); ++i) {
            float ms = 0; cudaEventElapsedTime(&ms, ev[i - 1].second, ev[i].second);
            n += snprintf(line + n, sizeof(line) - n, " %s=%.1f", ev[i].first, ms);
        }
        fprintf(stderr, "%s\n", line);
        for (auto &p : ev) cudaEventDestroy(p.second);
    }
};

inline void mm(char ta, char tb, i64 m, i64 n, i64 k, double alpha, const double *A, i64 lda, const double *B, i64 ldb,
               double beta, double *C, i64 ldc) {
    Gemm g;
    g.ta = ta; g.tb = tb; g.m = m; g.n = n; g.k = k; g.alpha = alpha; g.beta = beta;
    g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc;
    gemm(g);
}
// C = op(A) * Omega, Omega(kk, j) = normal(seed, off + kk*sk + j*sc)
inline void sketch(char ta, i64 m, i64 n, i64 k, const double *A, i64 lda, uint64_t seed, i64 sk, i64 sc, i64 off,
                   double *C, i64 ldc) {
    Gemm g;
    g.ta = ta; g.tb = 'N'; g.m = m; g.n = n; g.k = k; g.A = A; g.lda = lda; g.B = nullptr; g.ldb = 0; g.C = C; g.ldc = ldc;
    g.philox = true; g.seed = seed; g.ph_sk = sk; g.ph_sc = sc; g.ph_off = off;
    gemm(g);
}

__global__ void sqrt_clamp_kernel(double *w, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) w[i] = sqrt(fmax(w[i], 0.0));
}

// V(Icol[i], j) = (i < k) ? delta_ij : T(j, i-k)   — V = [I_k ; T^T](Icol^{-1}, :)   (RRA:2200-2219)
__global__ void build_v_kernel(const double *Icol, const double *T, i64 ldt, i64 n, i64 k, double *V, i64 ldv) {
    i64 total = n * k;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
        i64 i = e % n, j = e / n;
        i64 dst = (i64)Icol[i];
        V[j * ldv + dst] = (i < k) ? (i == j ? 1.0 : 0.0) : T[(i - k) * ldt + j];
    }
}

// rows of the global matrix owned by this rank: R(i, :) = A(Irow[i] - row0, :) if owned else 0
__global__ void gather_rows_owned_kernel(const double *A, i64 lda, i64 mloc, i64 n, i64 row0, const double *idx, i64 k,
                                         double *B, i64 ldb) {
    i64 total = k * n;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
        i64 i = e % k, j = e / k;
        i64 r = (i64)idx[i] - row0;
        B[j * ldb + i] = (r >= 0 && r < mloc) ? A[j * lda + r] : 0.0;
    }
}

// in-place column reversal of A (m x n) and of the n-vector s
__global__ void reverse_cols_kernel(double *A, i64 lda, i64 m, i64 n, double *s) {
    const i64 j = blockIdx.x, jj = n - 1 - j;
    if (j >= jj) return;
    for (i64 r = threadIdx.x; r < m; r += blockDim.x) { double t = A[j * lda + r]; A[j * lda + r] = A[jj * lda + r]; A[jj * lda + r] = t; }
    if (threadIdx.x == 0 && s) { double t = s[j]; s[j] = s[jj]; s[jj] = t; }
}

// device-side tolerance test of randQB_pb_new (RRA:1771-1777): done = (sqrt(sumsq) < tol)
__global__ void tol_check_kernel(const double *sumsq, double tol, int *done, double *norm_out) {
    double nv = sqrt(sumsq[0]);
    norm_out[0] = nv;
    done[0] = (nv < tol) ? 1 : 0;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// SVD tail (RRA:133-225 and RRA:289-380)
// ---------------------------------------------------------------------------------------------------------
// factors from Bt = A^T Q (n x l, replicated; destroyed) and the orthonormal Q (m x l, row-sharded)
static int svd_tail(DBuf &Bt, i64 m, i64 n, double *Q, i64 ldq, i64 l, i64 k, int vnum, double *U, i64 ldu, double *S, double *V, i64 ldv, Phase &ph);

int svd_from_q(const double *A, i64 m, i64 n, i64 lda, double *Q, i64 ldq, i64 l, i64 k, int vnum,
               double *U, i64 ldu, double *S, double *V, i64 ldv) {
    DBuf Bt((size_t)n * l);
    Phase ph;
    mm('T', 'N', n, l, m, 1.0, A, lda, Q, ldq, 0.0, Bt.p, n);          // Bt = A^T Q   (RRA:139)  — last pass over A
    allreduce_sum(Bt.p, (size_t)n * l);
    ph.lap("Bt = A^T Q");
    return svd_tail(Bt, m, n, Q, ldq, l, k, vnum, U, ldu, S, V, ldv, ph);
}

// The same tail when only the QB RESIDUAL Ares = M - Q B is in HBM (randQB_pb_new overwrites its private copy, RRA:1750):
// M^T Q = Ares^T Q + B^T (Q^T Q), exact algebra — one pass over the residual (the pass the tail makes anyway), an l x l Gram
// matrix and an n x l x l product instead of re-uploading M from the host (80 GB at BASELINE configs[2]).
int svd_from_q_residual(const double *Ares, i64 m, i64 n, i64 lda, double *Q, i64 ldq, i64 l, const double *B, i64 ldb, i64 k, int vnum,
                        double *U, i64 ldu, double *S, double *V, i64 ldv) {
    DBuf Bt((size_t)n * l), G((size_t)l * l);
    Phase ph;
    mm('T', 'N', n, l, m, 1.0, Ares, lda, Q, ldq, 0.0, Bt.p, n);
    mm('T', 'N', l, l, m, 1.0, Q, ldq, Q, ldq, 0.0, G.p, l);
    if (ctx().world > 1) { allreduce_sum(Bt.p, (size_t)n * l); allreduce_sum(G.p, (size_t)l * l); }
    mm('T', 'N', n, l, l, 1.0, B, ldb, G.p, l, 1.0, Bt.p, n);          // Bt += B^T (Q^T Q)
    ph.lap("Bt = Ares^T Q + B^T QtQ");
    return svd_tail(Bt, m, n, Q, ldq, l, k, vnum, U, ldu, S, V, ldv, ph);
}

static int svd_tail(DBuf &Bt, i64 m, i64 n, double *Q, i64 ldq, i64 l, i64 k, int vnum, double *U, i64 ldu, double *S, double *V, i64 ldv, Phase &ph) {
    Ctx &c = ctx();
    if (vnum == 1 || vnum > 2) {
        DBuf Rhat((size_t)l * l), Uhat((size_t)l * l), Vhat_t((size_t)l * l), Vhat((size_t)l * l), sv((size_t)l);
        orthonormalize(Bt.p, n, n, l, Rhat.p, l, /*sharded=*/false);      // [Qhat, Rhat] = qr(Bt)  (RRA:146); Qhat overwrites Bt
        ph.lap("qr(Bt)");
        if (!c.jacobi_transpose) {
            jacobi_svd(Rhat.p, l, l, Uhat.p, l, sv.p, Vhat_t.p, l);       // Rhat = Uhat S Vhat^T   (RRA:152)
            transpose(Vhat_t.p, l, Vhat.p, l, l, l);
        } else {
            // Rhat^T = U' S V'^T  =>  Rhat = V' S U'^T : Uhat = V', Vhat = U'
            DBuf Rt((size_t)l * l);
            transpose(Rhat.p, l, Rt.p, l, l, l);
            jacobi_svd(Rt.p, l, l, Vhat.p, l, sv.p, Vhat_t.p, l);         // Vhat_t holds V'^T here
            transpose(Vhat_t.p, l, Uhat.p, l, l, l);
        }
        ph.lap("jacobi svd(Rhat)");
        mm('N', 'N', m, k, l, 1.0, Q, ldq, Vhat.p, l, 0.0, U, ldu);       // U = Q Vhat, first k columns (RRA:156,171)
        mm('N', 'N', n, k, l, 1.0, Bt.p, n, Uhat.p, l, 0.0, V, ldv);      // V = Qhat Uhat          (RRA:160,172)
        copy_matrix(sv.p, l, S, k, k, 1);
        ph.lap("form U, V");
    } else {
        // eig of B B^T (RRA:175-225): B = Q^T A = Bt^T, so B B^T = Bt^T Bt
        DBuf BBt((size_t)l * l), w((size_t)l), X((size_t)l * k);
        mm('T', 'N', l, l, n, 1.0, Bt.p, n, Bt.p, n, 0.0, BBt.p, l);
        jacobi_eig(BBt.p, l, l, w.p);                                     // ascending (RRA:190)
        sqrt_clamp_kernel<<<(unsigned)((l + 127) / 128), 128, 0, c.stream>>>(w.p, (int)l);   // RRA:196-198
        count_launch();
        const double *Uk = BBt.p + (l - k) * l;                           // keep the LAST k (RRA:220-223)
        mm('N', 'N', m, k, l, 1.0, Q, ldq, Uk, l, 0.0, U, ldu);           // U = Q Uhat             (RRA:203)
        copy_matrix(Uk, l, X.p, l, l, k);
        scale_cols(X.p, l, l, k, w.p + (l - k), 1);                       // Uhat S^{-1}            (RRA:208-211)
        mm('N', 'N', n, k, l, 1.0, Bt.p, n, X.p, l, 0.0, V, ldv);         // V = B^T Uhat S^{-1}    (RRA:212)
        copy_matrix(w.p + (l - k), k, S, k, k, 1);
    }
    return g_status;
}

// ---------------------------------------------------------------------------------------------------------
// low_rank_svd_rand_decomp_fixed_rank (RRA:73-234)
// ---------------------------------------------------------------------------------------------------------
// Y = A * Omega(:, block) for a matrix that may still be arriving: with `up` the sketch consumes the column chunks of A
// as their upload events fire (Y += A(:,chunk) * Omega(chunk,:), Omega generated in the kernel), so the first pass over A
// hides behind the PCIe transfer.  Omega(kk, j) = normal(seed, off + kk + j*n).
static void sketch_chunked(const double *A, i64 m, i64 n, i64 lda, i64 l, uint64_t seed, i64 off, double *Y, i64 ldy, const Upload *up) {
    Ctx &c = ctx();
    if (!up || up->ev.empty()) { sketch('N', m, l, n, A, lda, seed, 1, n, off, Y, ldy); return; }
    for (size_t i = 0; i < up->ev.size(); ++i) {
        const i64 c0 = (i64)i * up->cw, w = std::min(up->cw, n - c0);
        RSVD_CUDA(cudaStreamWaitEvent(c.stream, up->ev[i], 0));
        Gemm g;
        g.ta = 'N'; g.tb = 'N'; g.m = m; g.n = l; g.k = w; g.A = A + c0 * lda; g.lda = lda; g.C = Y; g.ldc = ldy;
        g.beta = (i == 0) ? 0.0 : 1.0;
        g.philox = true; g.seed = seed; g.ph_sk = 1; g.ph_sc = n; g.ph_off = off + c0;
        gemm(g);
    }
}

// `Y0` (optional): an already computed sketch A*Omega (m x l, ld m).  `up` (optional): A is still being uploaded in column chunks.
int svd_rand_impl(const double *A, i64 m, i64 n, i64 lda, i64 k, i64 p, int vnum, int q, int s, uint64_t seed,
                  const double *omega, DBuf *Y0, const Upload *up, double *U, i64 ldu, double *S, double *V, i64 ldv) {
    ensure_init();
    if (!ctx().inited) return 1;
    const i64 l = k + p;
    if (k <= 0 || p < 0 || l > n || s <= 0) { set_error("rsvd_b200: invalid parameters k=%lld p=%lld s=%d (need 0 < k+p <= n, s > 0)", (long long)k, (long long)p, s); return 1; }
    DBuf Y, Z((size_t)n * l);
    Phase ph;
    if (Y0) Y = std::move(*Y0);
    else {
        Y.alloc((size_t)m * l);
        if (omega) mm('N', 'N', m, l, n, 1.0, A, lda, omega, n, 0.0, Y.p, m);       // Y = M RN (RRA:95)
        else sketch_chunked(A, m, n, lda, l, seed, 0, Y.p, m, up);                  // RN generated in the B-operand producer
    }
    ph.lap("Y = A Omega (sketch)");
    for (int j = 1; j < q; ++j) {                                                   // NOTE j < q (RRA:101)
        if ((2 * j - 2) % s == 0) orthonormalize(Y.p, m, m, l, nullptr, 0, true, true);   // RRA:106 (stabilisation only)
        ph.lap("orth(Y)");
        mm('T', 'N', n, l, m, 1.0, A, lda, Y.p, m, 0.0, Z.p, n);                    // Z = M^T Y (RRA:108/112)
        allreduce_sum(Z.p, (size_t)n * l);
        ph.lap("Z = A^T Y");
        if ((2 * j - 1) % s == 0) orthonormalize(Z.p, n, n, l, nullptr, 0, false, true);  // RRA:118 (stabilisation only)
        ph.lap("orth(Z)");
        mm('N', 'N', m, l, n, 1.0, A, lda, Z.p, n, 0.0, Y.p, m);                    // Y = M Z (RRA:120/124)
        ph.lap("Y = A Z");
    }
    Z.release();
    orthonormalize(Y.p, m, m, l, nullptr, 0, true);                                 // Q (RRA:129-130)
    ph.lap("Q = orth(Y)");
    return svd_from_q(A, m, n, lda, Y.p, m, l, k, vnum, U, ldu, S, V, ldv);
}

int svd_rand(const double *A, i64 m, i64 n, i64 lda, i64 k, i64 p, int vnum, int q, int s, uint64_t seed,
             const double *omega, double *U, i64 ldu, double *S, double *V, i64 ldv) {
    return svd_rand_impl(A, m, n, lda, k, p, vnum, q, s, seed, omega, nullptr, nullptr, U, ldu, S, V, ldv);
}

// Same algorithm from a HOST matrix: the upload (hostapi.cu: upload_begin) runs in column chunks on the copy stream while
// the sketch pass consumes the chunks that have landed.  dA (m x n, ld m) receives the uploaded matrix.
int svd_rand_host(const double *hA, double *dA, i64 m, i64 n, i64 k, i64 p, int vnum, int q, int s, uint64_t seed,
                  double *U, i64 ldu, double *S, double *V, i64 ldv) {
    ensure_init();
    if (!ctx().inited) return 1;
    Upload up;
    upload_begin(hA, m, dA, m, n, up);
    int rc = svd_rand_impl(dA, m, n, m, k, p, vnum, q, s, seed, nullptr, nullptr, &up, U, ldu, S, V, ldv);
    upload_end(up);
    return rc;
}

// ---------------------------------------------------------------------------------------------------------
// randQB_pb_new (RRA:1576-1801)
// ---------------------------------------------------------------------------------------------------------
// legacy_reorth: re-orthogonalise Qp against Q(:, 0:c0) on EVERY step > 0 as randQB_pb does (RRA:1503-1528) instead of on
// even steps only (randQB_pb_new, RRA:1703).
int randqb(double *A, i64 m, i64 n, i64 lda, i64 kstep, i64 nstep, double tol, int q, int s, uint64_t seed,
           double *Q, i64 ldq, double *B, i64 ldb, i64 max_rank, i64 *frank_out, int legacy_reorth, const Upload *up) {
    ensure_init();
    Ctx &c = ctx();
    if (!c.inited) return 1;
    if (kstep <= 0 || s <= 0) { set_error("rsvd_b200: invalid kstep/s"); return 1; }
    const bool tolMode = nstep <= 0;
    if (tolMode) nstep = max_rank / kstep;
    if (kstep * nstep > max_rank) nstep = max_rank / kstep;
    DBuf Yp((size_t)m * kstep), W((size_t)n * kstep), sums(2);
    int *done = c.d_flag + 24;
    i64 frank = 0;
    for (i64 step = 0; step < nstep; ++step) {                                      // RRA:1635
        Phase ph;
        const i64 c0 = kstep * step;
        sketch_chunked(A, m, n, lda, kstep, seed, c0 * n, Yp.p, m, step == 0 ? up : nullptr);   // Yp = A RN(:,block) (RRA:1643-1644)
        ph.lap("sketch");
        for (int j = 1; j <= q; ++j) {                                              // NOTE j <= q (RRA:1652)
            if ((2 * j - 2) % s == 0) orthonormalize(Yp.p, m, m, kstep, nullptr, 0, true, true);   // RRA:1655
            ph.lap("orth");
            mm('T', 'N', n, kstep, m, 1.0, A, lda, Yp.p, m, 0.0, W.p, n);           // AtQp (RRA:1657-1658 / 1665)
            allreduce_sum(W.p, (size_t)n * kstep);
            ph.lap("AtY");
            if ((2 * j - 1) % s == 0) orthonormalize(W.p, n, n, kstep, nullptr, 0, false, true);   // RRA:1673
            ph.lap("orth");
            mm('N', 'N', m, kstep, n, 1.0, A, lda, W.p, n, 0.0, Yp.p, m);           // Yp = A AtQp2 (RRA:1674 / 1681)
            ph.lap("AZ");
        }
        orthonormalize(Yp.p, m, m, kstep, nullptr, 0, true);                        // Qp (RRA:1690)
        ph.lap("Qp=orth");
        if (step > 0 && (legacy_reorth || step % 2 == 0)) {                         // RRA:1703-1722
            DBuf T((size_t)c0 * kstep);
            mm('T', 'N', c0, kstep, m, 1.0, Q, ldq, Yp.p, m, 0.0, T.p, c0);
            allreduce_sum(T.p, (size_t)c0 * kstep);
            mm('N', 'N', m, kstep, c0, -1.0, Q, ldq, T.p, c0, 1.0, Yp.p, m);
            orthonormalize(Yp.p, m, m, kstep, nullptr, 0, true);
        }
        ph.lap("reorth");
        double *Bp = B + c0;                                                        // B(block,:) (RRA:1761)
        {   // Bp = Qp^T A (RRA:1741), computed as (A^T Qp)^T: with A as the streamed operand the n x kstep product tiles without
            // waste (kstep = 200 rows would fill 200 of 256 tile rows the other way round: measured 70.8 vs 57.1 ms), and the
            // all-reduce works on a contiguous buffer.  W is free again at this point.
            mm('T', 'N', n, kstep, m, 1.0, A, lda, Yp.p, m, 0.0, W.p, n);
            allreduce_sum(W.p, (size_t)n * kstep);
            transpose(W.p, n, Bp, ldb, n, kstep);
        }
        {   // A = A - Qp Bp (RRA:1750-1751); in tolerance mode the epilogue also accumulates ||A - Qp Bp||_F^2 (RRA:1771),
            // so the residual norm costs no extra 8mn-byte pass
            Gemm g;
            g.ta = 'N'; g.tb = 'N'; g.m = m; g.n = n; g.k = kstep; g.alpha = -1.0; g.beta = 1.0;
            g.A = Yp.p; g.lda = m; g.B = Bp; g.ldb = ldb; g.C = A; g.ldc = lda;
            if (tolMode) g.sumsq_out = sums.p;
            ph.lap("Bp=QptA");
            gemm(g);
            ph.lap("A-=QpBp");
        }
        copy_matrix(Yp.p, m, Q + c0 * ldq, ldq, m, kstep);                          // Q(:,block) (RRA:1760)
        frank = (step + 1) * kstep;                                                 // RRA:1770
        if (tolMode) {                                                              // RRA:1771-1777 (absolute Frobenius norm)
            allreduce_sum(sums.p, 1);
            tol_check_kernel<<<1, 1, 0, c.stream>>>(sums.p, tol, done, sums.p + 1); // the test itself is evaluated on the device;
            count_launch();                                                         // the host reads back 4 bytes to end the loop
            RSVD_CUDA(cudaMemcpyAsync(c.h_flag + 24, done, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
            RSVD_CUDA(cudaStreamSynchronize(c.stream));
            if (c.verbose) {
                double nv = 0;
                cudaMemcpy(&nv, sums.p + 1, 8, cudaMemcpyDeviceToHost);
                fprintf(stderr, "[rsvd_b200] randQB step %lld: ||A_res||_F = %g\n", (long long)step, nv);
            }
            if (c.h_flag[24]) break;
        }
        if (g_status) break;
    }
    *frank_out = frank;
    return g_status;
}

// ---------------------------------------------------------------------------------------------------------
// ID family
// ---------------------------------------------------------------------------------------------------------
// shared tail (RRA:1938-1956 / 1836-1850): pivoted QR of Y (r x n, destroyed), I, T = R11^{-1} R12 with k rows.
// Y holds ALL columns on this rank (replicated data): no communication.
static void id_tail(double *Y, i64 ldy, i64 r, i64 n, i64 k, double *I, double *T, i64 ldt) {
    if (geqp3_blocked_ok(r, n) && !ctx().force_unblocked_qr) { geqp3_id(Y, ldy, r, n, 0, n, n, false, k, I, T, ldt); return; }
    geqp3(Y, ldy, r, n, I);
    if (n > k) {
        copy_matrix(Y + k * ldy, ldy, T, ldt, k, n - k);         // Rk2 = R(0:k, k:n)
        trsm_left_upper(Y, ldy, k, T, ldt, n - k);               // T = triu(Rk1)^{-1} Rk2  (dtrsm reads only the upper triangle)
    }
}

// The same tail for a TALL transposed input Xt (n x r, ld ldx: column j of the r x n matrix is row j of Xt) that is either
//   replicated (rows0 < 0): every rank holds all n rows — with several ranks the columns are dealt out in equal ranges so the
//     HBM-bound per-step sweep is divided by the world size; or
//   row-sharded (rows0 >= 0): this rank holds rows [rows0, rows0 + nloc) of n_global — the row ID of the two-sided ID, whose
//     input MI = M(:, Icol(1:k)) is born row-partitioned (RRA:2071-2078): nothing is gathered, the pivoted QR runs on the
//     shards with one small all-gather per step.
static void id_tail_transposed(const double *Xt, i64 ldx, i64 nloc_in, i64 rows0, i64 n_global, i64 r, i64 k, double *I, double *T, i64 ldt) {
    Ctx &c = ctx();
    const int W = c.world;
    // A born-sharded input can only go through the blocked (shardable) kernel: it takes up to 4096 rows there whatever the
    // "qr_blocked_rows" preference for single-GPU inputs says.
    const bool sharded_in = W > 1 && rows0 >= 0;
    const bool blocked_ok = sharded_in ? (r >= 1 && r <= 4096 && n_global < (1ll << 31) - 64) : geqp3_blocked_ok(r, n_global);
    if (!blocked_ok || (c.force_unblocked_qr && !sharded_in) || W == 1) {
        if (sharded_in) { set_error("rsvd_b200: the sharded pivoted QR needs <= 4096 rows (got %lld)", (long long)r); return; }
        DBuf Y((size_t)r * n_global);
        transpose(Xt, ldx, Y.p, r, n_global, r);
        id_tail(Y.p, r, r, n_global, k, I, T, ldt);
        return;
    }
    i64 per, col0, nloc;
    const double *src;
    if (rows0 < 0) {                                   // replicated: deal out column ranges
        if (n_global < 32768) {                        // too small to amortise a per-step exchange: every rank does the whole thing
            DBuf Y((size_t)r * n_global);
            transpose(Xt, ldx, Y.p, r, n_global, r);
            id_tail(Y.p, r, r, n_global, k, I, T, ldt);
            return;
        }
        per = ((n_global + W - 1) / W + 7) / 8 * 8;
        col0 = std::min(n_global, per * c.rank);
        nloc = std::min(per, n_global - col0);
        src = Xt + col0;
    } else {                                           // born sharded: must be the regular row partition
        per = ((n_global + W - 1) / W + 15) / 16 * 16;
        col0 = rows0; nloc = nloc_in; src = Xt;
        if (col0 != std::min(n_global, per * c.rank) || nloc != std::min(per, n_global - col0)) {
            set_error("rsvd_b200: row-partitioned ID needs the regular partition of rsvd_b200_row_partition (rank %d holds rows %lld..%lld of %lld)",
                      c.rank, (long long)col0, (long long)(col0 + nloc), (long long)n_global);
            return;
        }
    }
    DBuf Y((size_t)r * std::max((i64)1, nloc));
    if (nloc > 0) transpose(src, ldx, Y.p, r, nloc, r);
    geqp3_id(Y.p, r, r, nloc, col0, n_global, per, true, k, I, T, ldt);
}

int id_rand(const double *A, i64 m, i64 n, i64 lda, i64 k, i64 p, int q, int s, uint64_t seed, const double *omega,
            double *I, double *T, i64 ldt, const Upload *up) {
    ensure_init();
    Ctx &c = ctx();
    if (!c.inited) return 1;
    const i64 l = k + p;
    if (k <= 0 || p < 0 || s <= 0 || l > n) { set_error("rsvd_b200: invalid parameters k=%lld p=%lld s=%d", (long long)k, (long long)p, s); return 1; }
    // The reference keeps Y as l x n and transposes around every QR (RRA:1889-1921); here the panels are held
    // transposed (tall) throughout: Yt = Y^T (n x l), Wt = (Z M^T)^T = M Z^T (m x l).
    DBuf Yt((size_t)n * l), Wt((size_t)m * l);
    Phase ph;
    if (omega) {   // omega is the reference's RN, l x m column-major (global rows; this rank uses columns row0..row0+m)
        mm('T', 'T', n, l, m, 1.0, A, lda, omega + c.row0 * l, l, 0.0, Yt.p, n);
    } else if (up && !up->ev.empty()) {
        // A is still arriving in column chunks: rows [c0, c0+w) of Yt = A(:, chunk)^T RN^T need only that chunk
        for (size_t i = 0; i < up->ev.size(); ++i) {
            const i64 c0 = (i64)i * up->cw, w = std::min(up->cw, n - c0);
            RSVD_CUDA(cudaStreamWaitEvent(c.stream, up->ev[i], 0));
            sketch('T', w, l, m, A + c0 * lda, lda, seed, l, 1, c.row0 * l, Yt.p + c0, n);
        }
    } else {
        sketch('T', n, l, m, A, lda, seed, l, 1, c.row0 * l, Yt.p, n);              // Y = RN M (RRA:1877), RN(c,i) at i*l + c
    }
    allreduce_sum(Yt.p, (size_t)n * l);
    ph.lap("Y = RN M (left sketch)");
    for (int j = 1; j <= q; ++j) {                                                  // NOTE j <= q (RRA:1882)
        if ((2 * j - 2) % s == 0) orthonormalize(Yt.p, n, n, l, nullptr, 0, false); // Z = qr(Y')' (RRA:1889-1894)
        mm('N', 'N', m, l, n, 1.0, A, lda, Yt.p, n, 0.0, Wt.p, m);                  // Y = Z M^T (RRA:1908)
        if ((2 * j - 1) % s == 0) orthonormalize(Wt.p, m, m, l, nullptr, 0, true);  // RRA:1912-1917
        mm('T', 'N', n, l, m, 1.0, A, lda, Wt.p, m, 0.0, Yt.p, n);                  // Y = Z M (RRA:1927)
        allreduce_sum(Yt.p, (size_t)n * l);
    }
    Wt.release();
    ph.lap("power iterations");
    id_tail_transposed(Yt.p, n, n, -1, n, l, k, I, T, ldt);                         // pivoted QR of Y (l x n), T = Rk1^{-1} Rk2 (RRA:1938-1956)
    ph.lap("pivoted QR + T");
    return g_status;
}

int id_full(const double *M, i64 k, i64 n, i64 ldm, double *I, double *T, i64 ldt) {
    ensure_init();
    if (!ctx().inited) return 1;
    DBuf W((size_t)k * n);
    copy_matrix(M, ldm, W.p, k, k, n);
    id_tail(W.p, k, k, n, k, I, T, ldt);
    return g_status;
}

// pivoted QR of an r x n matrix M (copied), I, T = R11(k x k)^{-1} R12   (id_rand_decomp_fromQB, oneapi_code/
// rank_revealing_algorithms_one_api.c:421-444; the B-factor step of id_blockrand_decomp_fixed_rank_or_prec, RRA:1996-2020)
int id_qr(const double *M, i64 r, i64 n, i64 ldm, i64 k, double *I, double *T, i64 ldt) {
    ensure_init();
    if (!ctx().inited) return 1;
    if (k > std::min(r, n) || k <= 0) { set_error("rsvd_b200: id_qr needs 0 < k <= min(rows, cols) (k=%lld, %lld x %lld)", (long long)k, (long long)r, (long long)n); return 1; }
    DBuf W((size_t)r * n);
    copy_matrix(M, ldm, W.p, r, r, n);
    id_tail(W.p, r, r, n, k, I, T, ldt);
    return g_status;
}

// row ID of MI = A(:, Icol(1:k))  (RRA:2071-2078 -> RRA:1830-1850): Irow (m_global), S k x (m_global - k)
int id_rows(const double *A, i64 m, i64 n, i64 lda, const double *Icol, i64 k, double *Irow, double *S, i64 lds, i64 m_global) {
    ensure_init();
    Ctx &c = ctx();
    if (!c.inited) return 1;
    (void)n;
    DBuf MI((size_t)std::max((i64)1, m) * k);
    Phase ph;
    gather_cols(A, lda, m, Icol, k, MI.p, m);                                        // MI = M(:, Icol(1:k)) (RRA:2073)
    // RRA:2074-2078 -> RRA:1830-1850: full pivoted QR of MI^T (k x m_global).  Row-partitioned: MI^T is column-sharded and
    // stays that way (no gather of the 8*k*m_global-byte matrix).
    id_tail_transposed(MI.p, std::max((i64)1, m), m, c.world > 1 ? c.row0 : -1, c.world > 1 ? m_global : m, k, k, Irow, S, lds);
    ph.lap("row ID: pivoted QR of MI^T + S");
    return g_status;
}

int id_two_sided_rand(const double *A, i64 m, i64 n, i64 lda, i64 k, i64 p, int q, int s, uint64_t seed,
                      double *Icol, double *Irow, double *T, i64 ldt, double *S, i64 lds, i64 m_global, const Upload *up) {
    if (id_rand(A, m, n, lda, k, p, q, s, seed, nullptr, Icol, T, ldt, up)) return 1;  // RRA:2068
    return id_rows(A, m, n, lda, Icol, k, Irow, S, lds, m_global);
}

// CUR from a two-sided ID (RRA:2200-2252): C = M(:,Icol(1:k)), R = M(Irow(1:k),:), U = ((R R^T)^{-1} R V)^T, V = [I;T^T](Icol^{-1},:)
int cur_from_id(const double *A, i64 m, i64 n, i64 lda, const double *Icol, const double *Irow, const double *T, i64 ldt, i64 k,
                double *Cm, i64 ldc, double *U, i64 ldu, double *R, i64 ldr) {
    ensure_init();
    Ctx &c = ctx();
    if (!c.inited) return 1;
    DBuf V((size_t)n * k);
    {
        i64 total = n * k;
        int blocks = (int)min((i64)c.sms * 8, (total + 255) / 256);
        build_v_kernel<<<max(blocks, 1), 256, 0, c.stream>>>(Icol, T, ldt, n, k, V.p, n);   // RRA:2200-2219
        count_launch();
    }
    {   // R = M(Irow(1:k), :)  (RRA:2230-2231)
        i64 total = k * n;
        int blocks = (int)min((i64)c.sms * 8, (total + 255) / 256);
        if (c.world == 1) gather_rows(A, lda, n, Irow, k, R, ldr);
        else {
            DBuf t((size_t)k * n);
            gather_rows_owned_kernel<<<max(blocks, 1), 256, 0, c.stream>>>(A, lda, m, n, c.row0, Irow, k, t.p, k);
            count_launch();
            allreduce_sum(t.p, (size_t)k * n);
            copy_matrix(t.p, k, R, ldr, k, n);
        }
    }
    gather_cols(A, lda, m, Icol, k, Cm, ldc);                                        // C = M(:, Icol(1:k)) (RRA:2236-2237)
    DBuf RRt((size_t)k * k), RV((size_t)k * k), Rt((size_t)n * k);
    transpose(R, ldr, Rt.p, n, k, n);
    mm('T', 'N', k, k, n, 1.0, Rt.p, n, Rt.p, n, 0.0, RRt.p, k);                     // R R^T (RRA:2247)
    mm('T', 'N', k, k, n, 1.0, Rt.p, n, V.p, n, 0.0, RV.p, k);                       // R V   (RRA:2248)
    // (R R^T) U^T = R V (RRA:2250, dgesv in the reference).  R R^T is symmetric positive definite whenever the k selected rows are
    // independent, so the solve goes through the blocked Cholesky factor G^T G and its inverse (three k^3 GEMM-speed steps)
    // instead of the one-CTA LU, which at k = 1000 took 0.3 s — as long as the whole two-sided ID; the LU with partial pivoting
    // remains the fallback when the Cholesky factorisation breaks down (numerically dependent rows).
    bool solved = false;
    {
        DBuf G((size_t)k * k), Ginv((size_t)k * k), Y((size_t)k * k), X((size_t)k * k), Res((size_t)k * k);
        copy_matrix(RRt.p, k, G.p, k, k, k);
        if (potrf_upper(G.p, k, k) == 0 && !g_status) {
            trtri_upper(G.p, k, k, Ginv.p, k);
            mm('T', 'N', k, k, k, 1.0, Ginv.p, k, RV.p, k, 0.0, Y.p, k);           // Y = G^{-T} (R V)
            mm('N', 'N', k, k, k, 1.0, Ginv.p, k, Y.p, k, 0.0, X.p, k);            // X = G^{-1} Y
            // one step of iterative refinement against the unfactored R R^T: the accuracy of a backward-stable solver
            copy_matrix(RV.p, k, Res.p, k, k, k);
            mm('N', 'N', k, k, k, -1.0, RRt.p, k, X.p, k, 1.0, Res.p, k);          // Res = R V - (R R^T) X
            mm('T', 'N', k, k, k, 1.0, Ginv.p, k, Res.p, k, 0.0, Y.p, k);
            mm('N', 'N', k, k, k, 1.0, Ginv.p, k, Y.p, k, 1.0, X.p, k);            // X += G^{-1} G^{-T} Res
            copy_matrix(X.p, k, RV.p, k, k, k);                                    // U^T
            solved = true;
        }
    }
    if (!solved) {
        int info = lu_solve(RRt.p, k, k, RV.p, k, k);
        if (info) set_error("rsvd_b200: CUR core solve hit a zero pivot at column %d", info);
    }
    transpose(RV.p, k, U, ldu, k, k);                                                // U = (U^T)^T (RRA:2252)
    return g_status;
}

int cur_rand(const double *A, i64 m, i64 n, i64 lda, i64 k, i64 p, int q, int s, uint64_t seed,
             double *Cm, i64 ldc, double *U, i64 ldu, double *R, i64 ldr, i64 m_global, const Upload *up) {
    ensure_init();
    if (!ctx().inited) return 1;
    DBuf Icol((size_t)n), Irow((size_t)m_global), T((size_t)k * max((i64)1, n - k)), S((size_t)k * max((i64)1, m_global - k));
    if (id_two_sided_rand(A, m, n, lda, k, p, q, s, seed, Icol.p, Irow.p, T.p, k, S.p, k, m_global, up)) return 1;   // RRA:2198
    S.release();
    return cur_from_id(A, m, n, lda, Icol.p, Irow.p, T.p, k, k, Cm, ldc, U, ldu, R, ldr);
}

// low_rank_svd_rand_decomp_fromQB (oneapi_code/rank_revealing_algorithms_one_api.c:244-304, restated in FP64):
// B B^T = Uhat S^2 Uhat^T (descending), U = Q Uhat, V = B^T Uhat S^{-1}; all l = rows(B) triplets are returned.
// ascending != 0: factors ordered by ascending singular value, the dsyev order randomized_low_rank_svd4 returns (RRA:664-688)
int svd_from_qb(const double *Q, i64 m, i64 ldq, const double *B, i64 l, i64 n, i64 ldb, double *U, i64 ldu, double *S,
                double *V, i64 ldv, int ascending) {
    ensure_init();
    Ctx &c = ctx();
    if (!c.inited) return 1;
    DBuf Bt((size_t)n * l), BBt((size_t)l * l), Uhat((size_t)l * l), Vt((size_t)l * l), X((size_t)l * l);
    transpose(B, ldb, Bt.p, n, l, n);
    mm('T', 'N', l, l, n, 1.0, Bt.p, n, Bt.p, n, 0.0, BBt.p, l);                     // B B^T (:262)
    jacobi_svd(BBt.p, l, l, Uhat.p, l, S, Vt.p, l);                                  // :268
    sqrt_clamp_kernel<<<(unsigned)((l + 127) / 128), 128, 0, c.stream>>>(S, (int)l); // :277-280
    count_launch();
    if (ascending) {
        reverse_cols_kernel<<<(unsigned)((l + 1) / 2 > 0 ? (l + 1) / 2 : 1), 128, 0, c.stream>>>(Uhat.p, l, l, l, S);
        count_launch();
    }
    mm('N', 'N', m, l, l, 1.0, Q, ldq, Uhat.p, l, 0.0, U, ldu);                      // U = Q Uhat (:286)
    copy_matrix(Uhat.p, l, X.p, l, l, l);
    scale_cols(X.p, l, l, l, S, 1);                                                  // Uhat S^{-1} (:293-295)
    mm('N', 'N', n, l, l, 1.0, Bt.p, n, X.p, l, 0.0, V, ldv);                        // V = B^T Uhat S^{-1} (:296)
    return g_status;
}

// ---------------------------------------------------------------------------------------------------------
// Deterministic baselines and legacy entry points
// ---------------------------------------------------------------------------------------------------------
// Full SVD behind low_rank_svd_decomp_fixed_rank_or_prec (RRA:7-69: dgesvd 'S','S' on a copy, then truncation by the
// caller).  A (m x n) = U diag(S) V^T with r = min(m,n): U m x r, S r descending, V n x r.  The tall orientation is
// QR-factored (CholeskyQR2, TSQR fallback) and the r x r triangular factor goes through the one-sided Jacobi kernel.
int svd_full(const double *A, i64 m, i64 n, i64 lda, double *U, i64 ldu, double *S, double *V, i64 ldv) {
    ensure_init();
    Ctx &c = ctx();
    if (!c.inited) return 1;
    const i64 r = min(m, n), t = max(m, n);
    if (r <= 0) return g_status;
    DBuf W((size_t)t * r), R((size_t)r * r), Uh((size_t)r * r), Vt((size_t)r * r);
    if (m >= n) copy_matrix(A, lda, W.p, t, m, n);
    else transpose(A, lda, W.p, t, m, n);                                  // W = A^T (n x m)
    orthonormalize(W.p, t, t, r, R.p, r, /*sharded=*/false);               // W = Qw, R upper (rank-deficient input: Householder completion)
    jacobi_svd(R.p, r, r, Uh.p, r, S, Vt.p, r);                            // R = Uh S Vt
    if (m >= n) {                                                          // A = (Qw Uh) S Vt
        mm('N', 'N', m, r, r, 1.0, W.p, t, Uh.p, r, 0.0, U, ldu);
        transpose(Vt.p, r, V, ldv, r, r);
    } else {                                                               // A^T = (Qw Uh) S Vt  =>  A = Vt^T S (Qw Uh)^T
        mm('N', 'N', n, r, r, 1.0, W.p, t, Uh.p, r, 0.0, V, ldv);
        transpose(Vt.p, r, U, ldu, r, r);
    }
    return g_status;
}

// randQB_p (RRA:1343-1421): single-vector randQB with p power steps.  A (m x n) is destroyed (the reference deflates a
// private copy).  The reference's sequential projection loop runs over i < j-1 (it skips the newest column); it is applied
// here as one classical Gram-Schmidt pass over the same columns — the coefficients are O(eps) because A is deflated.
int randqb_single(double *A, i64 m, i64 n, i64 lda, i64 k, i64 p, uint64_t seed, double *Q, i64 ldq, double *B, i64 ldb) {
    ensure_init();
    Ctx &c = ctx();
    if (!c.inited) return 1;
    if (k <= 0 || p < 0) { set_error("rsvd_b200: randQB_p needs k > 0 and p >= 0"); return 1; }
    DBuf r((size_t)n), y((size_t)m), pj((size_t)n), cj((size_t)k), nrm(1);
    for (i64 j = 0; j < k; ++j) {
        fill_normal(r.p, n, seed, j * n);                                  // RN(:, j), RN n x k column-major (RRA:1350,1377)
        gemv('N', m, n, 1.0, A, lda, r.p, 0.0, y.p);                       // yj = A rj
        for (i64 i = 0; i < p; ++i) {
            gemv('T', m, n, 1.0, A, lda, y.p, 0.0, pj.p);
            gemv('N', m, n, 1.0, A, lda, pj.p, 0.0, y.p);
        }
        if (j - 1 > 0) {                                                   // i < j-1 (RRA:1387)
            gemv('T', m, j - 1, 1.0, Q, ldq, y.p, 0.0, cj.p);
            gemv('N', m, j - 1, -1.0, Q, ldq, cj.p, 1.0, y.p);
        }
        double *qj = Q + j * ldq;
        sumsq_async(y.p, m, m, 1, nrm.p);
        scale_by_inv_norm(y.p, m, nrm.p, qj);                              // qj = yj / ||yj||
        gemv('T', m, n, 1.0, A, lda, qj, 0.0, pj.p);                       // bj = A^T qj
        copy_matrix(pj.p, 1, B + j, ldb, 1, n);                            // B(j, :) = bj
        rank1_update(A, lda, m, n, qj, pj.p);                              // A -= qj bj^T
        if (g_status) break;
    }
    return g_status;
}

// estimate_rank_and_buildQ (MVF:1339-1400): Y = A RN with RN n x maxdim, sequential modified Gram-Schmidt over the columns
// with the reference's stop rule (two consecutive projections shorter than TOL), then an orthonormal basis of the first
// good_rank columns.  Q: m x maxdim buffer, the first *rank_out columns are the result.
int estimate_rank1(const double *A, i64 m, i64 n, i64 lda, i64 maxdim, double tol, uint64_t seed, double *Q, i64 ldq, i64 *rank_out) {
    ensure_init();
    Ctx &c = ctx();
    if (!c.inited) return 1;
    *rank_out = 0;
    if (maxdim <= 0 || maxdim > n) { set_error("rsvd_b200: estimate_rank_and_buildQ needs 0 < maxdim <= n (got %lld)", (long long)maxdim); return 1; }
    sketch('N', m, maxdim, n, A, lda, seed, 1, n, 0, Q, ldq);
    const i64 good = mgs_rank_estimate(Q, ldq, m, maxdim, tol);
    if (good < 0 || g_status) return 1;
    if (good > 0) orthonormalize(Q, ldq, m, good, nullptr, 0, false);   // QR_factorization_getQ(Qsmall) (MVF:1393)
    *rank_out = good;
    return g_status;
}

// estimate_rank_and_buildQ2 (MVF:1404-1467): grow Y = A [RN_0 RN_1 ...] by kblock columns until
// ||Q Q^T A - A||_F / ||Q Q^T A||_F <= tol (get_percent_error_between_two_mats(QQtM, M)/100).  The reference reseeds each
// RN_b with time(NULL), i.e. draws the SAME block again within one second; here block b continues the Philox stream
// (columns b*kblock.. of one wide RN), which is what the algorithm intends.  Stops at max_cols.
int estimate_rank2(const double *A, i64 m, i64 n, i64 lda, i64 kblock, double tol, uint64_t seed, double *Y, i64 ldy, double *Q, i64 ldq,
                   i64 max_cols, i64 *rank_out) {
    ensure_init();
    Ctx &c = ctx();
    if (!c.inited) return 1;
    *rank_out = 0;
    if (kblock <= 0 || kblock > max_cols) { set_error("rsvd_b200: estimate_rank_and_buildQ2 needs 0 < kblock <= min(m,n)"); return 1; }
    DBuf Res((size_t)m * n), sums(2);
    i64 cols = 0;
    for (;;) {
        sketch('N', m, kblock, n, A, lda, seed, 1, n, cols * n, Y + cols * ldy, ldy);
        cols += kblock;
        copy_matrix(Y, ldy, Q, ldq, m, cols);
        orthonormalize(Q, ldq, m, cols, nullptr, 0, false);
        DBuf B((size_t)cols * n);
        mm('T', 'N', cols, n, m, 1.0, Q, ldq, A, lda, 0.0, B.p, cols);
        copy_matrix(A, lda, Res.p, m, m, n);
        mm('N', 'N', m, n, cols, -1.0, Q, ldq, B.p, cols, 1.0, Res.p, m);
        sumsq_async(Res.p, m, m, n, sums.p);
        sumsq_async(B.p, cols, cols, n, sums.p + 1);
        double h[2] = {0, 0};
        RSVD_CUDA(cudaMemcpyAsync(h, sums.p, 16, cudaMemcpyDeviceToHost, c.stream));
        RSVD_CUDA(cudaStreamSynchronize(c.stream));
        if (g_status) return 1;
        const double err = sqrt(h[0]) / sqrt(h[1]);
        if (c.verbose) fprintf(stderr, "[rsvd_b200] estimate_rank2: %lld columns, error_norm = %g\n", (long long)cols, err);
        if (!(err > tol) || cols + kblock > max_cols) break;
    }
    *rank_out = cols;
    return g_status;
}

// power iterations + SVD tail from an existing sketch Y (m x l, destroyed): randomized_low_rank_svd3_autorank2 (RRA:826-918)
int svd_rand_from_sketch(const double *A, i64 m, i64 n, i64 lda, double *Y, i64 ldy, i64 l, int q, int s, double *U, i64 ldu,
                         double *S, double *V, i64 ldv) {
    DBuf Y0((size_t)m * l);
    copy_matrix(Y, ldy, Y0.p, m, m, l);
    return svd_rand_impl(A, m, n, lda, l, 0, 1, q, s, 0, nullptr, &Y0, nullptr, U, ldu, S, V, ldv);
}

// streamed 100*||A - U diag(S) V^T||_F / ||A||_F
double svd_percent_error(const double *A, i64 m, i64 n, i64 lda, const double *U, i64 ldu, const double *S,
                         const double *V, i64 ldv, i64 k) {
    ensure_init();
    Ctx &c = ctx();
    if (!c.inited) return -1.0;
    DBuf SVt((size_t)k * n), Vs((size_t)n * k), sums(2);
    copy_matrix(V, ldv, Vs.p, n, n, k);
    scale_cols(Vs.p, n, n, k, S, 0);
    transpose(Vs.p, n, SVt.p, k, n, k);     // (V S)^T = S V^T, k x n
    Vs.release();
    i64 nb = max((i64)128, min(n, (i64)((1ull << 30) / (8ull * (size_t)m))));
    nb = (nb / 128) * 128; if (nb <= 0) nb = 128;
    DBuf D((size_t)m * min(nb, n));
    double res2 = 0.0, a2 = 0.0;
    for (i64 j0 = 0; j0 < n; j0 += nb) {
        i64 w = min(nb, n - j0);
        copy_matrix(A + j0 * lda, lda, D.p, m, m, w);
        double h[2];
        sumsq_async(D.p, m, m, w, sums.p);
        mm('N', 'N', m, w, k, -1.0, U, ldu, SVt.p + j0 * k, k, 1.0, D.p, m);
        sumsq_async(D.p, m, m, w, sums.p + 1);
        RSVD_CUDA(cudaMemcpyAsync(h, sums.p, 16, cudaMemcpyDeviceToHost, c.stream));
        RSVD_CUDA(cudaStreamSynchronize(c.stream));
        a2 += h[0]; res2 += h[1];
    }
    if (c.world > 1) {
        double h[2] = {a2, res2};
        RSVD_CUDA(cudaMemcpyAsync(sums.p, h, 16, cudaMemcpyHostToDevice, c.stream));
        allreduce_sum(sums.p, 2);
        RSVD_CUDA(cudaMemcpyAsync(h, sums.p, 16, cudaMemcpyDeviceToHost, c.stream));
        RSVD_CUDA(cudaStreamSynchronize(c.stream));
        a2 = h[0]; res2 = h[1];
    }
    return 100.0 * sqrt(res2) / sqrt(a2);
}

}  // namespace rsvd
