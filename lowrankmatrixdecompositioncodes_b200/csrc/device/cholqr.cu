// cholqr.cu — re-orthonormalisation of tall-skinny panels: CholeskyQR2 with an l x l Gram matrix, and a
// TSQR(R)-preconditioned fallback when the Gram matrix is too ill-conditioned for Cholesky.
// Replaces QR_factorization_getQ / compact_QR_factorization (dgeqrf + dorgqr Householder QR,
// matrix_vector_functions_intel_mkl.c:1251-1263, 1214-1245) as called from RRA:106,118,130,146,1655,1673,1690,1722,
// 1892,1915.  The basis differs from LAPACK's by an l x l orthogonal factor (signs of R's diagonal are positive
// here); everything downstream on the hot path is invariant under that (SURVEY.md §7 hard part 3).
//
// Row-partitioned multi-GPU: the only communication is the allreduce of the l x l Gram matrix (twice).
#include "common.cuh"

namespace rsvd {

constexpr int PB = 64;   // Cholesky / inverse block size

// ---- blocked upper Cholesky G = R^T R ------------------------------------------------------------------
// Diagonal block (<= 64 x 64): Cholesky factor AND its inverse in one right-looking sweep, 1024 threads = 16 per row.
// Alongside L (= R^T) the kernel carries E, the running right-hand side of L E = I: at column c row c of E is final
// (E(c,k) = Acc(c,k)/piv) and every later row takes the same rank-1 update as the trailing block of L, so after 64 steps
// E = L^{-1} = (R^{-1})^T.  Each step: 128 threads scale and publish column c of L / row c of E, barrier, all threads apply
// the rank-1 update while the owner of the next pivot takes its reciprocal square root, barrier: 0.6 us per column (39 us
// per block, ncu) against 1.5 us for the one-thread-per-row version with a separate back-substitution.  A variant holding
// the block in registers (shared memory only for the broadcasts) measured slower (50 us).
// Outputs: R_jj in place and W = R_jj^{-1} (PB x PB, ld PB) for the panel solve.  flag = failing column + 1.
constexpr int PT = 16;                       // threads per row
__global__ void __launch_bounds__(PB * PT) potf2_inv_kernel(double *G, i64 ldg, i64 j0, int jb, double *W, int *flag) {
    extern __shared__ double potf2_smem[];
    double (*L)[PB + 1] = reinterpret_cast<double (*)[PB + 1]>(potf2_smem);                    // L[r][c], r >= c : lower factor (R^T)
    double (*E)[PB + 1] = reinterpret_cast<double (*)[PB + 1]>(potf2_smem + PB * (PB + 1));    // running L^{-1}
    const int r = threadIdx.x / PT, g = threadIdx.x % PT;
#pragma unroll
    for (int q = 0; q < PB / PT; ++q) {
        const int c = g + q * PT;
        L[r][c] = (r < jb && c <= r) ? G[(j0 + r) * ldg + j0 + c] : 0.0;   // G(c, r) upper -> L(r, c)
        E[r][c] = (r == c) ? 1.0 : 0.0;
    }
    double *colc = potf2_smem + 2 * PB * (PB + 1);      // scaled column c of L (rows k > c, zero elsewhere)
    double *erow = colc + PB;                           // final row c of E (columns k <= c, zero elsewhere)
    double *pv = erow + PB;                             // pv[0] = 1/sqrt(pivot), pv[1] = sqrt(pivot) of the coming column
    __syncthreads();
    if (threadIdx.x == 0) {
        double d = L[0][0];
        if (!(d > 0.0)) { atomicCAS(flag, 0, (int)(j0 + 1)); d = 1.0; }   // keep going with finite numbers; the caller discards the result
        const double inv = rsqrt(d);
        pv[0] = inv; pv[1] = d * inv;
    }
    __syncthreads();
    for (int c = 0; c < jb; ++c) {
        // phase A (128 threads): scale column c of L and row c of E once, instead of in every thread that uses them
        if (threadIdx.x < 2 * PB) {
            const double inv = pv[0];
            const int k = threadIdx.x & (PB - 1);
            if (threadIdx.x < PB) {
                const double v = (k > c) ? L[k][c] * inv : 0.0;
                colc[k] = v;
                if (k > c) L[k][c] = v; else if (k == c) L[c][c] = pv[1];
            } else {
                const double v = (k <= c) ? E[c][k] * inv : 0.0;
                erow[k] = v;
                if (k <= c) E[c][k] = v;
            }
        }
        __syncthreads();
        // phase B (all threads): rank-1 update of the trailing rows of L and E; the owner of the next pivot prepares it
        if (r > c && r < jb) {
            const double lrc = colc[r];
#pragma unroll
            for (int q = 0; q < PB / PT; ++q) {
                const int k = g + q * PT;
                if (k > c && k <= r) {
                    const double v = fma(-lrc, colc[k], L[r][k]);
                    L[r][k] = v;
                    if (r == c + 1 && k == c + 1) {
                        double d = v;
                        if (!(d > 0.0)) { atomicCAS(flag, 0, (int)(j0 + c + 2)); d = 1.0; }
                        const double inv = rsqrt(d);
                        pv[0] = inv; pv[1] = d * inv;
                    }
                }
                if (k <= c) E[r][k] = fma(-lrc, erow[k], E[r][k]);
            }
        }
        __syncthreads();
    }
    // R_jj (upper) back in place: R(c, r) = L(r, c); W(i, j) = R^{-1}(i, j) = E(j, i), identity outside the jb x jb block
#pragma unroll
    for (int q = 0; q < PB / PT; ++q) {
        const int c = g + q * PT;
        if (r < jb && c < jb) G[(j0 + r) * ldg + j0 + c] = (c <= r) ? L[r][c] : 0.0;
        W[r * PB + c] = (r < jb) ? ((c <= r) ? E[r][c] : 0.0) : (c == r ? 1.0 : 0.0);
    }
}

int potrf_upper(double *G, i64 ldg, i64 n) {
    ensure_init();
    if (g_status) return -1;
    int *flag = ctx().d_flag;
    RSVD_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), ctx().stream));
    DBuf W((size_t)PB * PB), T((size_t)PB * (n > PB ? n - PB : 1));
    const size_t potf2_bytes = (2 * PB * (PB + 1) + 2 * PB + 2) * sizeof(double);
    RSVD_CUDA(cudaFuncSetAttribute(potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)potf2_bytes));   // per device: cheap enough per call
    for (i64 j0 = 0; j0 < n; j0 += PB) {
        int jb = (int)min((i64)PB, n - j0);
        potf2_inv_kernel<<<1, PB * PT, potf2_bytes, ctx().stream>>>(G, ldg, j0, jb, W.p, flag);
        count_launch();
        i64 rest = n - j0 - jb;
        if (rest > 0) {
            // row panel: R12 = R11^{-T} G12  (GEMM with the inverted diagonal block), then G22 -= R12^T R12
            double *G12 = G + (j0 + jb) * ldg + j0;
            Gemm t;
            t.ta = 'T'; t.tb = 'N'; t.m = jb; t.n = rest; t.k = jb; t.A = W.p; t.lda = PB; t.B = G12; t.ldb = ldg; t.C = T.p; t.ldc = PB;
            gemm(t);
            copy_matrix(T.p, PB, G12, ldg, jb, rest);
            Gemm g;
            g.ta = 'T'; g.tb = 'N'; g.m = rest; g.n = rest; g.k = jb; g.alpha = -1.0; g.beta = 1.0;
            g.A = G12; g.lda = ldg; g.B = G12; g.ldb = ldg;
            g.C = G + (j0 + jb) * ldg + j0 + jb; g.ldc = ldg;
            gemm(g);
        }
    }
    keep_upper(G, ldg, n);
    RSVD_CUDA(cudaMemcpyAsync(ctx().h_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx().stream));
    RSVD_CUDA(cudaStreamSynchronize(ctx().stream));
    return ctx().h_flag[0];
}

// ---- inverse of an upper-triangular matrix by recursive doubling -------------------------------------------
// diagonal PB x PB blocks: thread j back-substitutes column j of the inverse
__global__ void __launch_bounds__(PB) trtri_diag_kernel(const double *__restrict__ R, i64 ldr, i64 n, double *__restrict__ X, i64 ldx) {
    __shared__ double S[PB][PB + 1];
    const i64 o = (i64)blockIdx.x * PB;
    const int nb = (int)min((i64)PB, n - o);
    for (int e = threadIdx.x; e < PB * PB; e += blockDim.x) {
        int i = e % PB, j = e / PB;
        S[i][j] = (i < nb && j < nb && i <= j) ? R[(o + j) * ldr + o + i] : (i == j ? 1.0 : 0.0);
    }
    __syncthreads();
    const int j = threadIdx.x;
    if (j >= nb) return;
    double x[PB];
#pragma unroll 1
    for (int i = j; i >= 0; --i) {
        double s = (i == j) ? 1.0 : 0.0;
        for (int l = i + 1; l <= j; ++l) s -= S[i][l] * x[l];
        x[i] = s / S[i][i];
    }
    for (int i = 0; i <= j; ++i) X[(o + j) * ldx + o + i] = x[i];
}

void trtri_upper(const double *R, i64 ldr, i64 n, double *X, i64 ldx) {
    if (g_status) return;   // an earlier error (e.g. a failed allocation) is pending: launch nothing
    ensure_init();
    if (n <= 0) return;
    // X = 0, then diagonal blocks
    RSVD_CUDA(cudaMemset2DAsync(X, (size_t)ldx * 8, 0, (size_t)n * 8, (size_t)n, ctx().stream));
    int nblk = (int)((n + PB - 1) / PB);
    trtri_diag_kernel<<<nblk, PB, 0, ctx().stream>>>(R, ldr, n, X, ldx);
    count_launch();
    DBuf tmp((size_t)n * n);
    // level s: merge [X11 ?; 0 X22] of sizes (s, r): X12 = -X11 * R12 * X22
    for (i64 s = PB; s < n; s *= 2) {
        i64 npairs_full = n / (2 * s);                 // pairs with two full blocks
        i64 tail0 = npairs_full * 2 * s;               // start of a possibly partial pair
        i64 tail_r = (n - tail0 > s) ? (n - tail0 - s) : 0;   // size of the partial second block
        // T = R12 * X22   (s x r)
        if (npairs_full > 0) {
            Gemm g;
            g.ta = 'N'; g.tb = 'N'; g.m = s; g.n = s; g.k = s;
            g.A = R + s * ldr; g.lda = ldr; g.sA = 2 * s * (ldr + 1);
            g.B = X + s * ldx + s; g.ldb = ldx; g.sB = 2 * s * (ldx + 1);
            g.C = tmp.p; g.ldc = s; g.sC = s * s; g.batch = (int)npairs_full;
            gemm(g);
            Gemm h;
            h.ta = 'N'; h.tb = 'N'; h.m = s; h.n = s; h.k = s; h.alpha = -1.0;
            h.A = X; h.lda = ldx; h.sA = 2 * s * (ldx + 1);
            h.B = tmp.p; h.ldb = s; h.sB = s * s;
            h.C = X + s * ldx; h.ldc = ldx; h.sC = 2 * s * (ldx + 1); h.batch = (int)npairs_full;
            gemm(h);
        }
        if (tail_r > 0) {
            const double *R12 = R + (tail0 + s) * ldr + tail0;
            double *X11 = X + tail0 * ldx + tail0, *X22 = X + (tail0 + s) * ldx + tail0 + s, *X12 = X + (tail0 + s) * ldx + tail0;
            Gemm g;
            g.ta = 'N'; g.tb = 'N'; g.m = s; g.n = tail_r; g.k = tail_r; g.A = R12; g.lda = ldr; g.B = X22; g.ldb = ldx;
            g.C = tmp.p; g.ldc = s;
            gemm(g);
            Gemm h;
            h.ta = 'N'; h.tb = 'N'; h.m = s; h.n = tail_r; h.k = s; h.alpha = -1.0; h.A = X11; h.lda = ldx; h.B = tmp.p; h.ldb = s;
            h.C = X12; h.ldc = ldx;
            gemm(h);
        }
    }
}

// ---- diagonal statistics of R: out[0] = min diag, out[1] = max diag --------------------------------------------
__global__ void diag_minmax_kernel(const double *R, i64 ldr, i64 n, double *out) {
    __shared__ double smin[32], smax[32];
    double mn = INFINITY, mx = 0.0;
    for (i64 i = threadIdx.x; i < n; i += blockDim.x) {
        double d = fabs(R[i * ldr + i]);
        if (!(d == d)) d = 0.0;   // NaN -> treat as breakdown
        mn = fmin(mn, d); mx = fmax(mx, d);
    }
    for (int o = 16; o > 0; o >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = mn; smax[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { mn = fmin(mn, smin[w]); mx = fmax(mx, smax[w]); }
        out[0] = mn; out[1] = mx;
    }
}

// One Cholesky-QR pass: Qout = Yin * chol(Yin^T Yin)^{-1}.  Returns 0 ok, 1 = Gram not safely factorable.
// Rout (l x l) receives the Cholesky factor.  cond_limit: reject when max/min of diag(R) exceeds it.
static int cholqr_pass(const double *Yin, i64 ldy, i64 m, i64 l, double *Qout, i64 ldq, double *Rout, double cond_limit, bool sharded,
                       double *ratio_out = nullptr) {
    DBuf G((size_t)l * l), Rinv((size_t)l * l), stat(2);
    Gemm g;   // Gram = Yin^T Yin
    g.ta = 'T'; g.tb = 'N'; g.m = l; g.n = l; g.k = m; g.A = Yin; g.lda = ldy; g.B = Yin; g.ldb = ldy; g.C = G.p; g.ldc = l;
    g.sym_upper = true;                                   // potrf_upper reads the upper triangle only
    set_zero(G.p, (size_t)l * l);                         // skipped tiles must not feed NaNs into the all-reduce / trailing updates
    gemm(g);
    if (sharded) allreduce_sum(G.p, (size_t)l * l);
    double h[2];
    int info = ctx().no_chol_dataflow ? -1 : chol_inv_upper(G.p, l, l, Rinv.p, l, h);   // factor, inverse and diagonal statistics in one kernel
    if (info > 0 || g_status) return 1;
    const bool fused = (info == 0);
    if (!fused) {
        info = potrf_upper(G.p, l, l);
        if (info != 0) return 1;
        diag_minmax_kernel<<<1, 256, 0, ctx().stream>>>(G.p, l, l, stat.p);
        count_launch();
        RSVD_CUDA(cudaMemcpyAsync(h, stat.p, 16, cudaMemcpyDeviceToHost, ctx().stream));
        RSVD_CUDA(cudaStreamSynchronize(ctx().stream));
    }
    if (ratio_out) *ratio_out = (h[0] > 0.0) ? h[1] / h[0] : INFINITY;
    if (!(h[0] > 0.0) || h[1] / h[0] > cond_limit) return 1;
    if (!fused) trtri_upper(G.p, l, l, Rinv.p, l);
    Gemm q;   // Qout = Yin * Rinv
    q.ta = 'N'; q.tb = 'N'; q.m = m; q.n = l; q.k = l; q.A = Yin; q.lda = ldy; q.B = Rinv.p; q.ldb = l; q.C = Qout; q.ldc = ldq;
    q.b_upper = true;                                     // trtri_upper leaves exact zeros below the diagonal
    gemm(q);
    if (Rout) copy_matrix(G.p, l, Rout, l, l, l);
    return 0;
}

// G(i,i) += coef * trace(G)   (one CTA; the shift of the shifted Cholesky-QR pass)
__global__ void shift_diag_kernel(double *G, i64 ldg, i64 n, double coef) {
    __shared__ double sh[32];
    __shared__ double tr;
    double t = 0.0;
    for (i64 i = threadIdx.x; i < n; i += blockDim.x) t += G[i * ldg + i];
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
        tr = t;
    }
    __syncthreads();
    const double s = coef * tr;
    for (i64 i = threadIdx.x; i < n; i += blockDim.x) G[i * ldg + i] += s;
}

// One SHIFTED Cholesky-QR pass (Fukaya, Kannan, Nakatsukasa, Yamamoto, Yanagisawa: shifted CholeskyQR3, SIAM J. Sci. Comput. 2020):
// Qout = Yin * chol(Yin^T Yin + s I)^{-1} with s = 11 (m l + l (l+1)) u ||Yin||^2, ||.||_2^2 bounded by trace(Gram).  The shift makes
// the factorisation succeed for any cond(Yin) < 1/u and leaves cond(Qout) ~ sqrt(s) / sigma_min(Yin): a preconditioner at GEMM
// speed for panels the plain pass cannot take (cond > ~1e7) — where the Householder-based TSQR path costs two orders of magnitude
// more (measured 450-700 ms against 2 ms for a 100000 x 200 panel).  Returns 0 ok, 1 = not factorable (a zero / non-finite panel).
static int shifted_cholqr_pass(const double *Yin, i64 ldy, i64 m, i64 m_global, i64 l, double *Qout, i64 ldq, double *Rout, bool sharded) {
    DBuf G((size_t)l * l), Rinv((size_t)l * l), stat(2);
    Gemm g;
    g.ta = 'T'; g.tb = 'N'; g.m = l; g.n = l; g.k = m; g.A = Yin; g.lda = ldy; g.B = Yin; g.ldb = ldy; g.C = G.p; g.ldc = l;
    g.sym_upper = true;
    set_zero(G.p, (size_t)l * l);
    gemm(g);
    if (sharded) allreduce_sum(G.p, (size_t)l * l);
    const double coef = 11.0 * ((double)m_global * (double)l + (double)l * (double)(l + 1)) * 1.1102230246251565e-16;
    shift_diag_kernel<<<1, 256, 0, ctx().stream>>>(G.p, l, l, coef);
    count_launch();
    double h[2];
    int info = ctx().no_chol_dataflow ? -1 : chol_inv_upper(G.p, l, l, Rinv.p, l, h);
    if (info > 0 || g_status) return 1;
    if (info < 0) {
        if (potrf_upper(G.p, l, l) != 0) return 1;
        trtri_upper(G.p, l, l, Rinv.p, l);
    } else if (!(h[0] > 0.0) || !(h[1] < INFINITY)) return 1;
    Gemm q;
    q.ta = 'N'; q.tb = 'N'; q.m = m; q.n = l; q.k = l; q.A = Yin; q.lda = ldy; q.B = Rinv.p; q.ldb = l; q.C = Qout; q.ldc = ldq;
    q.b_upper = true;
    gemm(q);
    copy_matrix(G.p, l, Rout, l, l, l);
    return 0;
}

// R3 = R2 * R1 (upper * upper), result in R1buf
static void accumulate_r(double *R2, double *R1, i64 l) {
    DBuf T((size_t)l * l);
    Gemm g;
    g.ta = 'N'; g.tb = 'N'; g.m = l; g.n = l; g.k = l; g.A = R2; g.lda = l; g.B = R1; g.ldb = l; g.C = T.p; g.ldc = l;
    gemm(g);
    copy_matrix(T.p, l, R1, l, l, l);
}

// TSQR(R) fallback: Householder R of the local row block (no squaring of the condition number), stacked over the
// ranks and factored again; then Q1 = Y R^{-1} has cond(Q1) = O(1 + eps*cond(Y)) and one Cholesky-QR pass finishes.
static int tsqr_r(const double *Y, i64 ldy, i64 m, i64 l, double *R /* l x l */, bool sharded) {
    Ctx &c = ctx();
    // local Householder R.  Row blocks shorter than l are padded with zero rows.
    i64 mm = max(m, l);
    DBuf W((size_t)mm * l);
    if (mm > m) set_zero(W.p, (size_t)mm * l);
    copy_matrix(Y, ldy, W.p, mm, m, l);
    geqrf_r(W.p, mm, mm, l);
    if (c.world == 1 || !sharded) {
        copy_matrix(W.p, mm, R, l, l, l);
        keep_upper(R, l, l);
        return 0;
    }
    // stack the world's R factors (allgather emulated by a zero-padded sum), factor the (world*l) x l stack
    i64 rows = (i64)c.world * l;
    DBuf S((size_t)rows * l);
    set_zero(S.p, (size_t)rows * l);
    copy_matrix(W.p, mm, S.p + (i64)c.rank * l, rows, l, l);
    // zero the strictly-lower part of this rank's slot (Householder vectors live there)
    {
        DBuf T((size_t)l * l);
        copy_matrix(S.p + (i64)c.rank * l, rows, T.p, l, l, l);
        keep_upper(T.p, l, l);
        copy_matrix(T.p, l, S.p + (i64)c.rank * l, rows, l, l);
    }
    allreduce_sum(S.p, (size_t)rows * l);
    geqrf_r(S.p, rows, rows, l);
    copy_matrix(S.p, rows, R, l, l, l);
    keep_upper(R, l, l);
    return 0;
}

void orthonormalize(double *Y, i64 ldy, i64 m, i64 l, double *R, i64 ldr, bool sharded, bool loose) {
    if (g_status) return;   // an earlier error (e.g. a failed allocation) is pending: launch nothing
    ensure_init();
    if (m <= 0 || l <= 0) return;
    Ctx &c = ctx();
    DBuf Q1((size_t)m * l), R1((size_t)l * l), R2((size_t)l * l);
    int bad = 1;
    double ratio_dbg = -1.0;      // -1: the first Cholesky factorisation itself broke down
    if (!c.force_qr_fallback) {
        // CholeskyQR2: cond(Y) up to ~1e7 is safe (cond(Gram) = cond(Y)^2 must stay well below 1/eps)
        double ratio = INFINITY;
        bad = cholqr_pass(Y, ldy, m, l, Q1.p, m, R1.p, 1.0e7, sharded, &ratio);
        if (ratio < INFINITY) ratio_dbg = ratio;
        if (!bad && loose && !R && ratio <= 1.0e4) {
            // stabilisation-only call (an intermediate step of a power iteration, followed by another orthonormalisation
            // before anything is measured): one Cholesky-QR pass leaves ||Q^T Q - I|| ~ eps*cond(Y)^2 <~ 1e-8, the same
            // range, and is all the reference's `s > 1` skipping ever asks for.
            copy_matrix(Q1.p, m, Y, ldy, m, l);
            c.last_qr_path = 1;
            return;
        }
        if (!bad) {
            bad = cholqr_pass(Q1.p, m, m, l, Y, ldy, R2.p, 1.0e3, sharded);
            if (!bad) c.last_qr_path = 1;
        }
    }
    if (bad && c.force_qr_fallback != 1) {
        // Shifted CholeskyQR3: up to two shifted passes bring cond(Y) (anything below 1/u) under the plain pass's limit, then
        // CholeskyQR2 as above.  R = R_plain2 * R_plain1 * R_shifted...; everything at GEMM speed.
        c.qr_fallbacks++;
        const i64 mg = (sharded && c.world > 1) ? (c.m_global > 0 ? (i64)c.m_global : m * c.world) : m;
        DBuf P((size_t)m * l), Racc((size_t)l * l), Rs((size_t)l * l);
        const double *src = Y;
        i64 lds = ldy;
        for (int it = 0; it < 2 && bad; ++it) {
            double *dst = (it == 0) ? P.p : Q1.p;                                  // it = 1 reads P, writes Q1
            if (shifted_cholqr_pass(src, lds, m, mg, l, dst, m, Rs.p, sharded)) break;
            if (it == 0) copy_matrix(Rs.p, l, Racc.p, l, l, l); else accumulate_r(Rs.p, Racc.p, l);
            double *tmp = (it == 0) ? Q1.p : P.p;                                  // plain pass 1 writes the other buffer
            double ratio = INFINITY;
            int b = cholqr_pass(dst, m, m, l, tmp, m, R1.p, 1.0e7, sharded, &ratio);
            if (c.verbose) fprintf(stderr, "[rsvd_b200] orthonormalize %lld x %lld: shifted Cholesky-QR pass %d (first plain pass saw diag(R) max/min %.3g), then max/min %.3g\n",
                                   (long long)m, (long long)l, it + 1, ratio_dbg, ratio);
            if (!b) {
                b = cholqr_pass(tmp, m, m, l, Y, ldy, R2.p, 1.0e3, sharded);
                if (!b) {
                    bad = 0;
                    c.last_qr_path = 4;
                    if (R) {
                        accumulate_r(R1.p, Racc.p, l);       // R1 * Racc
                        accumulate_r(R2.p, Racc.p, l);       // R2 * R1 * Racc
                        copy_matrix(Racc.p, l, R, ldr, l, l);
                    }
                    return;
                }
                // (a refused pass returns before its apply GEMM, so Y is still the caller's panel for the paths below)
            }
            src = dst; lds = m;
        }
    }
    if (bad) {
        // TSQR-preconditioned path
        if (c.force_qr_fallback == 1) c.qr_fallbacks++;
        c.last_qr_path = 2;
        if (c.verbose) fprintf(stderr, "[rsvd_b200] orthonormalize %lld x %lld: Cholesky-QR refused the panel (diag(R) max/min %.3g), TSQR-preconditioned path\n",
                               (long long)m, (long long)l, ratio_dbg);
        tsqr_r(Y, ldy, m, l, R1.p, sharded);
        DBuf Rinv((size_t)l * l);
        trtri_upper(R1.p, l, l, Rinv.p, l);
        Gemm q;
        q.ta = 'N'; q.tb = 'N'; q.m = m; q.n = l; q.k = l; q.A = Y; q.lda = ldy; q.B = Rinv.p; q.ldb = l; q.C = Q1.p; q.ldc = m;
        gemm(q);
        int b2 = cholqr_pass(Q1.p, m, m, l, Y, ldy, R2.p, 1.0e7, sharded);
        if (b2) {
            // numerically rank-deficient panel: one more preconditioned sweep on Q1 (its R is well scaled now)
            DBuf R3((size_t)l * l), Q2((size_t)m * l);
            tsqr_r(Q1.p, m, m, l, R3.p, sharded);
            trtri_upper(R3.p, l, l, Rinv.p, l);
            Gemm q2 = q; q2.A = Q1.p; q2.lda = m; q2.C = Q2.p; q2.ldc = m;
            gemm(q2);
            accumulate_r(R3.p, R1.p, l);
            b2 = cholqr_pass(Q2.p, m, m, l, Y, ldy, R2.p, 1.0e12, sharded);
            if (b2) {
                // Numerically rank-deficient panel (e.g. an exactly low-rank A sketched with more columns than its rank): no
                // triangular solve can produce an orthonormal basis of all l columns.  Last resort = what the reference does
                // (dgeqrf + dorgqr, MVF:1251-1263): Householder QR with the explicit Q, whose extra columns are an orthonormal
                // completion.  BLAS-2 speed (one kernel per reflector), only ever reached on singular panels.
                c.last_qr_path = 3;
                if (sharded && c.world > 1) {
                    // Row-partitioned panel: TSQR with explicit factors.  Y_g = Q_g R_g (local Householder, explicit Q_g), the
                    // world's R_g are stacked by an all-gather and factored again with an explicit Q_s; Q = Q_g Q_s[g] is
                    // orthonormal for any rank (Householder completes the basis), R is the factor of the stack.
                    if (m < l) {
                        set_error("rsvd_b200: orthonormalisation of a numerically singular row-partitioned panel needs >= %lld rows per rank (this one holds %lld)", (long long)l, (long long)m);
                        return;
                    }
                    const i64 rows = (i64)c.world * l;
                    DBuf Rg((size_t)l * l), stack((size_t)rows * l), Rs((size_t)l * l), Qg((size_t)m * l), Qs((size_t)l * l);
                    copy_matrix(Y, ldy, Qg.p, m, m, l);
                    geqrf_q(Qg.p, m, m, l, Rg.p, l);                          // Qg <- Q_g, Rg <- R_g
                    DBuf all((size_t)rows * l);
                    allgather(Rg.p, all.p, (size_t)l * l);                    // [R_0 | R_1 | ...], each l x l column-major
                    for (int g = 0; g < c.world; ++g) copy_matrix(all.p + (i64)g * l * l, l, stack.p + (i64)g * l, rows, l, l);
                    geqrf_q(stack.p, rows, rows, l, Rs.p, l);                 // stack <- Q_s (rows x l), Rs <- R
                    copy_matrix(stack.p + (i64)c.rank * l, rows, Qs.p, l, l, l);
                    Gemm q3;
                    q3.ta = 'N'; q3.tb = 'N'; q3.m = m; q3.n = l; q3.k = l; q3.A = Qg.p; q3.lda = m; q3.B = Qs.p; q3.ldb = l; q3.C = Y; q3.ldc = ldy;
                    gemm(q3);
                    if (R) copy_matrix(Rs.p, l, R, ldr, l, l);
                    return;
                }
                geqrf_q(Y, ldy, m, l, R ? R1.p : nullptr, l);   // Y <- Q, R1 <- R
                if (R) copy_matrix(R1.p, l, R, ldr, l, l);
                return;
            }
        }
    }
    if (R) {
        accumulate_r(R2.p, R1.p, l);   // R = R2 * R1
        copy_matrix(R1.p, l, R, ldr, l, l);
    }
}

}  // namespace rsvd
