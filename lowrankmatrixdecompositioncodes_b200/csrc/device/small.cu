// small.cu — HBM-bound helper kernels of the hot path: transposes, slicing/gathers, scaling, Frobenius norms,
// Gaussian fill, blocked triangular solve, small LU solve.  They replace the reference's OpenMP element loops
// and per-column helper calls (matrix_vector_functions_intel_mkl.c:216-283,360-372,647-753,959-1042) and the
// dtrsm / dgesv calls (MVF:1487-1493, 1525-1531).
#include "common.cuh"

namespace rsvd {

// ---- transpose -------------------------------------------------------------------------------------
__global__ void transpose_kernel(const double *__restrict__ A, i64 lda, double *__restrict__ B, i64 ldb, i64 m, i64 n) {
    __shared__ double tile[32][33];
    i64 i0 = (i64)blockIdx.x * 32, j0 = (i64)blockIdx.y * 32;
    for (int jj = threadIdx.y; jj < 32; jj += 8) {
        i64 i = i0 + threadIdx.x, j = j0 + jj;
        if (i < m && j < n) tile[jj][threadIdx.x] = A[j * lda + i];
    }
    __syncthreads();
    for (int ii = threadIdx.y; ii < 32; ii += 8) {
        i64 j = j0 + threadIdx.x, i = i0 + ii;
        if (i < m && j < n) B[i * ldb + j] = tile[threadIdx.x][ii];
    }
}
void transpose(const double *A, i64 lda, double *B, i64 ldb, i64 m, i64 n) {
    if (g_status) return;   // an earlier error (e.g. a failed allocation) is pending: launch nothing
    if (m <= 0 || n <= 0) return;
    // grid.y is limited to 65535 blocks: loop over column slabs if needed
    const i64 slab = 65535ll * 32;
    for (i64 j0 = 0; j0 < n; j0 += slab) {
        i64 nn = min(slab, n - j0);
        dim3 grid((unsigned)((m + 31) / 32), (unsigned)((nn + 31) / 32));
        transpose_kernel<<<grid, dim3(32, 8), 0, ctx().stream>>>(A + j0 * lda, lda, B + j0, ldb, m, nn);
        count_launch();
    }
}

// ---- copy / fill ---------------------------------------------------------------------------------------
__global__ void copy_kernel(const double *__restrict__ A, i64 lda, double *__restrict__ B, i64 ldb, i64 m, i64 n) {
    i64 total = m * n;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
        i64 i = e % m, j = e / m;
        B[j * ldb + i] = A[j * lda + i];
    }
}
static inline int grid_for(i64 total, int per = 256) {
    i64 b = (total + per - 1) / per;
    i64 cap = (i64)ctx().sms * 16;
    return (int)max((i64)1, min(b, cap));
}
void copy_matrix(const double *A, i64 lda, double *B, i64 ldb, i64 m, i64 n) {
    if (g_status) return;   // an earlier error (e.g. a failed allocation) is pending: launch nothing
    if (m <= 0 || n <= 0) return;
    if (lda == m && ldb == m) {
        RSVD_CUDA(cudaMemcpyAsync(B, A, (size_t)m * n * 8, cudaMemcpyDeviceToDevice, ctx().stream));
        return;
    }
    copy_kernel<<<grid_for(m * n), 256, 0, ctx().stream>>>(A, lda, B, ldb, m, n);
    count_launch();
}
void set_zero(double *A, size_t n) { if (n) RSVD_CUDA(cudaMemsetAsync(A, 0, n * 8, ctx().stream)); }

__global__ void identity_kernel(double *A, i64 lda, i64 n) {
    i64 total = n * n;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
        i64 i = e % n, j = e / n;
        A[j * lda + i] = (i == j) ? 1.0 : 0.0;
    }
}
void set_identity(double *A, i64 lda, i64 n) {
    if (g_status) return;   // an earlier error (e.g. a failed allocation) is pending: launch nothing
    if (n <= 0) return;
    identity_kernel<<<grid_for(n * n), 256, 0, ctx().stream>>>(A, lda, n);
    count_launch();
}
__global__ void keep_upper_kernel(double *A, i64 lda, i64 n) {
    i64 total = n * n;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
        i64 i = e % n, j = e / n;
        if (i > j) A[j * lda + i] = 0.0;
    }
}
void keep_upper(double *A, i64 lda, i64 n) {
    if (g_status) return;   // an earlier error (e.g. a failed allocation) is pending: launch nothing
    if (n <= 0) return;
    keep_upper_kernel<<<grid_for(n * n), 256, 0, ctx().stream>>>(A, lda, n);
    count_launch();
}

// ---- gathers: B = A(:, idx(0:k)) and B = A(idx(0:k), :)  (indices stored as doubles, RRA:967-970) -----------
__global__ void gather_cols_kernel(const double *__restrict__ A, i64 lda, i64 m, const double *__restrict__ idx, i64 k,
                                   double *__restrict__ B, i64 ldb) {
    i64 total = m * k;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
        i64 i = e % m, j = e / m;
        B[j * ldb + i] = A[(i64)idx[j] * lda + i];
    }
}
void gather_cols(const double *A, i64 lda, i64 m, const double *idx, i64 k, double *B, i64 ldb) {
    if (g_status) return;   // an earlier error (e.g. a failed allocation) is pending: launch nothing
    if (m <= 0 || k <= 0) return;
    gather_cols_kernel<<<grid_for(m * k), 256, 0, ctx().stream>>>(A, lda, m, idx, k, B, ldb);
    count_launch();
}
__global__ void gather_rows_kernel(const double *__restrict__ A, i64 lda, i64 n, const double *__restrict__ idx, i64 k,
                                   double *__restrict__ B, i64 ldb) {
    i64 total = k * n;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
        i64 i = e % k, j = e / k;
        B[j * ldb + i] = A[j * lda + (i64)idx[i]];
    }
}
void gather_rows(const double *A, i64 lda, i64 n, const double *idx, i64 k, double *B, i64 ldb) {
    if (g_status) return;   // an earlier error (e.g. a failed allocation) is pending: launch nothing
    if (n <= 0 || k <= 0) return;
    gather_rows_kernel<<<grid_for(k * n), 256, 0, ctx().stream>>>(A, lda, n, idx, k, B, ldb);
    count_launch();
}

// ---- column scaling: A(:,j) *= s[j] or /= s[j] ---------------------------------------------------------
__global__ void scale_cols_kernel(double *A, i64 lda, i64 m, i64 n, const double *s, int invert) {
    i64 total = m * n;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
        i64 i = e % m, j = e / m;
        double f = invert ? 1.0 / s[j] : s[j];
        A[j * lda + i] *= f;
    }
}
void scale_cols(double *A, i64 lda, i64 m, i64 n, const double *s, int invert) {
    if (g_status) return;   // an earlier error (e.g. a failed allocation) is pending: launch nothing
    if (m <= 0 || n <= 0) return;
    scale_cols_kernel<<<grid_for(m * n), 256, 0, ctx().stream>>>(A, lda, m, n, s, invert);
    count_launch();
}

// ---- sum of squares (deterministic two-pass reduction) ---------------------------------------------------
__global__ void sumsq_partial_kernel(const double *__restrict__ A, i64 lda, i64 m, i64 n, double *__restrict__ part) {
    __shared__ double sh[32];
    i64 total = m * n;
    double s = 0.0;
    if (lda == m) {
        for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
            double v = A[e];
            s = fma(v, v, s);
        }
    } else {
        for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
            double v = A[(e / m) * lda + (e % m)];
            s = fma(v, v, s);
        }
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) part[blockIdx.x] = s;
    }
}
__global__ void sum_final_kernel(const double *__restrict__ part, i64 n, double *out) {
    __shared__ double sh[32];
    double s = 0.0;
    for (i64 i = threadIdx.x; i < n; i += blockDim.x) s += part[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) out[0] = s;
    }
}
void sumsq_async(const double *A, i64 lda, i64 m, i64 n, double *d_out) {
    if (g_status) return;   // an earlier error (e.g. a failed allocation) is pending: launch nothing
    if (m <= 0 || n <= 0) { set_zero(d_out, 1); return; }
    int blocks = (int)min((i64)ctx().sms * 8, (m * n + 1023) / 1024);
    if (blocks < 1) blocks = 1;
    DBuf part((size_t)blocks);
    sumsq_partial_kernel<<<blocks, 512, 0, ctx().stream>>>(A, lda, m, n, part.p);
    sum_final_kernel<<<1, 256, 0, ctx().stream>>>(part.p, blocks, d_out);
    count_launch(2);
}
void sum_array_async(const double *part, i64 n, double *d_out) {
    if (g_status) return;
    sum_final_kernel<<<1, 1024, 0, ctx().stream>>>(part, n, d_out);
    count_launch();
}
double frob_norm(const double *A, i64 lda, i64 m, i64 n) {
    ensure_init();
    if (g_status) return -1.0;
    DBuf out(1);
    sumsq_async(A, lda, m, n, out.p);
    double h = 0.0;
    RSVD_CUDA(cudaMemcpyAsync(&h, out.p, 8, cudaMemcpyDeviceToHost, ctx().stream));
    RSVD_CUDA(cudaStreamSynchronize(ctx().stream));
    if (ctx().world > 1) {   // row-partitioned: ||A||_F^2 = sum over ranks (one scalar allreduce)
        DBuf t(1);
        RSVD_CUDA(cudaMemcpyAsync(t.p, &h, 8, cudaMemcpyHostToDevice, ctx().stream));
        allreduce_sum(t.p, 1);
        RSVD_CUDA(cudaMemcpyAsync(&h, t.p, 8, cudaMemcpyDeviceToHost, ctx().stream));
        RSVD_CUDA(cudaStreamSynchronize(ctx().stream));
    }
    return sqrt(h);
}

// ---- Gaussian fill -------------------------------------------------------------------------------------
__global__ void fill_normal_kernel(double *A, i64 n, uint64_t seed, i64 first) {
    // one thread per Philox block of 4 consecutive linear entries
    i64 blk0 = first >> 2, blk1 = (first + n - 1) >> 2;
    for (i64 b = blk0 + (i64)blockIdx.x * blockDim.x + threadIdx.x; b <= blk1; b += (i64)gridDim.x * blockDim.x) {
        float z[4];
        rsvd_normal4(seed, (uint64_t)b, z);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            i64 lin = 4 * b + j;
            if (lin >= first && lin < first + n) A[lin - first] = (double)z[j];
        }
    }
}
void fill_normal(double *A, i64 n, uint64_t seed, i64 first) {
    if (n <= 0) return;
    fill_normal_kernel<<<grid_for((n + 3) / 4), 256, 0, ctx().stream>>>(A, n, seed, first);
    count_launch();
}

// ---- triangular solve B <- R^{-1} B (R k x k upper, B k x ncols), blocked with inverted diagonal blocks --------
constexpr int TB = 64;
// one CTA per diagonal block: Dinv[b] = inverse of R[b*TB.., b*TB..] (upper triangular, size <= TB)
__global__ void diag_inverse_kernel(const double *__restrict__ R, i64 ldr, i64 k, double *__restrict__ Dinv) {
    __shared__ double S[TB][TB + 1];
    const int b = blockIdx.x;
    const i64 o = (i64)b * TB;
    const int nb = (int)min((i64)TB, k - o);
    for (int e = threadIdx.x; e < TB * TB; e += blockDim.x) {
        int i = e % TB, j = e / TB;
        S[i][j] = (i < nb && j < nb && i <= j) ? R[(o + j) * ldr + o + i] : (i == j ? 1.0 : 0.0);
    }
    __syncthreads();
    // thread j solves R x = e_j by back substitution (column j of the inverse)
    const int j = threadIdx.x;
    if (j < TB) {
        double x[TB];
#pragma unroll 1
        for (int i = TB - 1; i >= 0; --i) {
            double s = (i == j) ? 1.0 : 0.0;
            if (i > j) { x[i] = 0.0; continue; }
            for (int l = i + 1; l <= j; ++l) s -= S[i][l] * x[l];
            x[i] = s / S[i][i];
        }
        double *D = Dinv + (i64)b * TB * TB;
        for (int i = 0; i < TB; ++i) D[j * TB + i] = x[i];
    }
}

void trsm_left_upper(const double *R, i64 ldr, i64 k, double *B, i64 ldb, i64 ncols) {
    if (g_status) return;   // an earlier error (e.g. a failed allocation) is pending: launch nothing
    if (k <= 0 || ncols <= 0) return;
    const int nblk = (int)((k + TB - 1) / TB);
    DBuf dinv((size_t)nblk * TB * TB);
    diag_inverse_kernel<<<nblk, TB, 0, ctx().stream>>>(R, ldr, k, dinv.p);
    count_launch();
    DBuf tmp((size_t)TB * ncols);
    for (int b = nblk - 1; b >= 0; --b) {
        const i64 o = (i64)b * TB;
        const i64 nb = min((i64)TB, k - o);
        const i64 rest = k - (o + nb);
        if (rest > 0) {   // B_b -= R[b, after] * B[after]
            Gemm g;
            g.ta = 'N'; g.tb = 'N'; g.m = nb; g.n = ncols; g.k = rest; g.alpha = -1.0; g.beta = 1.0;
            g.A = R + (o + nb) * ldr + o; g.lda = ldr; g.B = B + o + nb; g.ldb = ldb; g.C = B + o; g.ldc = ldb;
            gemm(g);
        }
        // B_b <- Dinv_b * B_b  (through a temporary: GEMM cannot run in place)
        copy_matrix(B + o, ldb, tmp.p, TB, nb, ncols);
        Gemm g;
        g.ta = 'N'; g.tb = 'N'; g.m = nb; g.n = ncols; g.k = nb; g.A = dinv.p + (i64)b * TB * TB; g.lda = TB;
        g.B = tmp.p; g.ldb = TB; g.C = B + o; g.ldc = ldb;
        gemm(g);
    }
}

// ---- small dense LU with partial pivoting (dgesv semantics), single CTA, matrix streamed from L2 ------------------
// n <= ~2048 on the hot path (CUR: k x k).  Right-looking, one column per step.
__global__ void __launch_bounds__(1024) lu_factor_kernel(double *A, i64 lda, int n, double *B, i64 ldb, int nrhs, int *flag);

int lu_solve(double *A, i64 lda, i64 n, double *B, i64 ldb, i64 nrhs) {
    ensure_init();
    if (g_status) return -1;
    if (n <= 0 || nrhs <= 0) return 0;
    int *flag = ctx().d_flag + 8;
    RSVD_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), ctx().stream));
    // single-CTA LU with partial pivoting + forward elimination of B, then the blocked upper-triangular solve
    lu_factor_kernel<<<1, 1024, 0, ctx().stream>>>(A, lda, (int)n, B, ldb, (int)nrhs, flag);
    count_launch();
    trsm_left_upper(A, lda, n, B, ldb, nrhs);
    RSVD_CUDA(cudaMemcpyAsync(ctx().h_flag + 8, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx().stream));
    RSVD_CUDA(cudaStreamSynchronize(ctx().stream));
    return ctx().h_flag[8];
}

// P A = L U by columns (first-max pivot like idamax); L's multipliers are applied to B on the fly
__global__ void __launch_bounds__(1024) lu_factor_kernel(double *A, i64 lda, int n, double *B, i64 ldb, int nrhs, int *flag) {
    __shared__ double red_v[32];
    __shared__ int red_i[32];
    __shared__ int piv_s;
    __shared__ double pivval_s;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int c = 0; c < n; ++c) {
        double best = -1.0; int bi = c;
        for (int i = c + tid; i < n; i += nt) {
            double v = fabs(A[(i64)c * lda + i]);
            if (v > best) { best = v; bi = i; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            double ov = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if ((tid & 31) == 0) { red_v[tid >> 5] = best; red_i[tid >> 5] = bi; }
        __syncthreads();
        if (tid < 32) {
            best = tid < (nt >> 5) ? red_v[tid] : -1.0;
            bi = tid < (nt >> 5) ? red_i[tid] : 0x7fffffff;
            for (int o = 16; o > 0; o >>= 1) {
                double ov = __shfl_xor_sync(0xffffffffu, best, o);
                int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (tid == 0) { piv_s = bi; pivval_s = best; if (best == 0.0) atomicExch(flag, c + 1); }
        }
        __syncthreads();
        const int pr = piv_s;
        if (pr != c) {
            for (int j = tid; j < n + nrhs; j += nt) {
                double *col = j < n ? A + (i64)j * lda : B + (i64)(j - n) * ldb;
                double t = col[c]; col[c] = col[pr]; col[pr] = t;
            }
        }
        __syncthreads();
        if (pivval_s == 0.0) continue;
        const double inv = 1.0 / A[(i64)c * lda + c];
        for (int i = c + 1 + tid; i < n; i += nt) A[(i64)c * lda + i] *= inv;
        __syncthreads();
        const int rows = n - c - 1;
        const int cols = (n - c - 1) + nrhs;
        for (i64 e = tid; e < (i64)rows * cols; e += nt) {
            int i = c + 1 + (int)(e % rows);
            int jj = (int)(e / rows);
            double *col = jj < n - c - 1 ? A + (i64)(c + 1 + jj) * lda : B + (i64)(jj - (n - c - 1)) * ldb;
            col[i] -= A[(i64)c * lda + i] * col[c];
        }
        __syncthreads();
    }
}


// ---- BLAS-2 pieces of the single-vector randQB (HBM-bound: every call streams A once) ----------------------------
// y = alpha * A x + beta * y : one thread per row, columns split over gridDim.y into partial sums that are reduced in a
// fixed order (deterministic)
__global__ void gemv_n_partial_kernel(const double *__restrict__ A, i64 lda, i64 m, i64 n, const double *__restrict__ x, double *part) {
    const i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    const i64 chunk = (n + gridDim.y - 1) / gridDim.y;
    const i64 j0 = (i64)blockIdx.y * chunk, j1 = min(n, j0 + chunk);
    double acc = 0.0;
    for (i64 j = j0; j < j1; ++j) acc = fma(A[j * lda + r], x[j], acc);
    part[(i64)blockIdx.y * m + r] = acc;
}
__global__ void gemv_n_reduce_kernel(const double *__restrict__ part, i64 m, int chunks, double alpha, double beta, double *y) {
    const i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    double acc = 0.0;
    for (int c = 0; c < chunks; ++c) acc += part[(i64)c * m + r];
    y[r] = alpha * acc + (beta != 0.0 ? beta * y[r] : 0.0);
}
// y = alpha * A^T x + beta * y : one warp per column
__global__ void gemv_t_kernel(const double *__restrict__ A, i64 lda, i64 m, i64 n, const double *__restrict__ x, double alpha, double beta, double *y) {
    const int lane = threadIdx.x & 31;
    i64 w = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const i64 nw = ((i64)gridDim.x * blockDim.x) >> 5;
    for (i64 j = w; j < n; j += nw) {
        const double *col = A + j * lda;
        double acc = 0.0;
        for (i64 r = lane; r < m; r += 32) acc = fma(col[r], x[r], acc);
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) y[j] = alpha * acc + (beta != 0.0 ? beta * y[j] : 0.0);
    }
}
void gemv(char trans, i64 m, i64 n, double alpha, const double *A, i64 lda, const double *x, double beta, double *y) {
    if (g_status || m <= 0 || n <= 0) return;
    Ctx &c = ctx();
    if (trans == 'N') {
        const int bx = (int)((m + 255) / 256);
        int chunks = (int)max((i64)1, min((i64)64, min(n / 32, (i64)(c.sms * 8) / bx)));
        DBuf part((size_t)chunks * m);
        gemv_n_partial_kernel<<<dim3(bx, chunks), 256, 0, c.stream>>>(A, lda, m, n, x, part.p);
        gemv_n_reduce_kernel<<<bx, 256, 0, c.stream>>>(part.p, m, chunks, alpha, beta, y);
        count_launch(2);
    } else {
        const int blocks = (int)max((i64)1, min((i64)c.sms * 8, (n + 7) / 8));
        gemv_t_kernel<<<blocks, 256, 0, c.stream>>>(A, lda, m, n, x, alpha, beta, y);
        count_launch();
    }
}
__global__ void rank1_update_kernel(double *A, i64 lda, i64 m, i64 n, const double *__restrict__ q, const double *__restrict__ b) {
    const i64 total = m * n;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
        const i64 r = e % m, j = e / m;
        A[j * lda + r] = fma(-q[r], b[j], A[j * lda + r]);
    }
}
void rank1_update(double *A, i64 lda, i64 m, i64 n, const double *q, const double *b) {
    if (g_status || m <= 0 || n <= 0) return;
    rank1_update_kernel<<<grid_for(m * n), 256, 0, ctx().stream>>>(A, lda, m, n, q, b);
    count_launch();
}
__global__ void scale_by_inv_norm_kernel(const double *__restrict__ y, i64 m, const double *__restrict__ sumsq, double *q) {
    const double inv = 1.0 / sqrt(sumsq[0]);
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < m; e += (i64)gridDim.x * blockDim.x) q[e] = y[e] * inv;
}
void scale_by_inv_norm(const double *y, i64 m, const double *sumsq, double *q) {
    if (g_status || m <= 0) return;
    scale_by_inv_norm_kernel<<<grid_for(m), 256, 0, ctx().stream>>>(y, m, sumsq, q);
    count_launch();
}


// Sequential modified Gram-Schmidt with the stop rule of estimate_rank_and_buildQ (MVF:1366-1388): column j is projected
// against the finished columns one at a time; when two consecutive projections (the second may belong to the previous
// column: p1 persists across j) are both shorter than tol the scan stops and j is the rank.  One CTA: the loop is a chain of
// dependent dot products.
__global__ void __launch_bounds__(1024) mgs_rank_kernel(double *Q, i64 ldq, i64 m, i64 maxdim, double tol, int *rank_out) {
    __shared__ double sh[2 * 32];
    __shared__ double bc[2];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5;
    double p1norm = 0.0;
    int good = (int)maxdim;
    for (i64 j = 0; j < maxdim; ++j) {
        double *vj = Q + j * ldq;
        bool stop = false;
        for (i64 i = 0; i < j; ++i) {
            const double *vi = Q + i * ldq;
            double d = 0.0, u = 0.0;
            for (i64 r = tid; r < m; r += nt) { const double a = vi[r]; d = fma(vj[r], a, d); u = fma(a, a, u); }
            for (int o = 16; o > 0; o >>= 1) { d += __shfl_xor_sync(0xffffffffu, d, o); u += __shfl_xor_sync(0xffffffffu, u, o); }
            if (lane == 0) { sh[2 * wid] = d; sh[2 * wid + 1] = u; }
            __syncthreads();
            if (tid == 0) {
                double dd = 0.0, uu = 0.0;
                for (int w = 0; w < (nt >> 5); ++w) { dd += sh[2 * w]; uu += sh[2 * w + 1]; }
                bc[0] = dd / uu; bc[1] = sqrt(uu);
            }
            __syncthreads();
            const double coef = bc[0], pnorm = fabs(coef) * bc[1];
            for (i64 r = tid; r < m; r += nt) vj[r] = fma(-coef, vi[r], vj[r]);
            __syncthreads();
            if (pnorm < tol && p1norm < tol) { good = (int)j; stop = true; break; }
            p1norm = pnorm;
        }
        if (stop) break;
        double u = 0.0;
        for (i64 r = tid; r < m; r += nt) u = fma(vj[r], vj[r], u);
        for (int o = 16; o > 0; o >>= 1) u += __shfl_xor_sync(0xffffffffu, u, o);
        if (lane == 0) sh[wid] = u;
        __syncthreads();
        if (tid == 0) { double uu = 0.0; for (int w = 0; w < (nt >> 5); ++w) uu += sh[w]; bc[0] = 1.0 / sqrt(uu); }
        __syncthreads();
        const double inv = bc[0];
        for (i64 r = tid; r < m; r += nt) vj[r] *= inv;
        __syncthreads();
    }
    if (tid == 0) *rank_out = good;
}
i64 mgs_rank_estimate(double *Q, i64 ldq, i64 m, i64 maxdim, double tol) {
    if (g_status) return -1;
    Ctx &c = ctx();
    mgs_rank_kernel<<<1, 1024, 0, c.stream>>>(Q, ldq, m, maxdim, tol, c.d_flag + 28);
    count_launch();
    RSVD_CUDA(cudaMemcpyAsync(c.h_flag + 28, c.d_flag + 28, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    RSVD_CUDA(cudaStreamSynchronize(c.stream));
    return g_status ? -1 : (i64)c.h_flag[28];
}

}  // namespace rsvd
