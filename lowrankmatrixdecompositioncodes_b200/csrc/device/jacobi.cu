// jacobi.cu — the small l x l SVD / symmetric eigen step as a one-sided (Hestenes) Jacobi kernel.
// Replaces singular_value_decomposition (LAPACKE_dgesvd 'S','S', matrix_vector_functions_intel_mkl.c:1270-1284,
// called at RRA:152 on the l x l factor Rhat) and compute_evals_and_evecs_of_symm_matrix (LAPACKE_dsyev,
// MVF:1206-1209, called at RRA:190 on B*B^T).
//
// One sweep = N-1 round-robin steps of N/2 disjoint column pairs; each pair is one CTA that keeps both columns in
// registers, forms the 2x2 Gram entries with a block reduction and rotates the columns of G and of the
// accumulated V.  The l x l working set (<= 2 x 8.8 MB at l = 1050) stays L2-resident.  A whole sweep is captured
// once in a CUDA graph and replayed until a sweep applies no rotation.
#include "common.cuh"
#include <algorithm>
#include <vector>

namespace rsvd {

namespace {

constexpr int JT = 128;     // threads per pair
constexpr int JR = 10;      // rows per thread: supports n <= 1280

__device__ __forceinline__ double block_sum3(double &a, double &b, double &c, double *sh) {
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sh[3 * w] = a; sh[3 * w + 1] = b; sh[3 * w + 2] = c; }
    __syncthreads();
    a = b = c = 0.0;
    for (int i = 0; i < JT / 32; ++i) { a += sh[3 * i]; b += sh[3 * i + 1]; c += sh[3 * i + 2]; }
    return 0.0;
}

__global__ void __launch_bounds__(JT) jacobi_step_kernel(double *G, i64 ldg, double *V, i64 ldv, int n, int N, int r,
                                                         double tol, int *rotated) {
    __shared__ double sh[3 * (JT / 32)];
    const int i = blockIdx.x;
    int p, q;
    if (i == 0) { p = N - 1; q = r; }
    else { p = (r + i) % (N - 1); q = (r - i + (N - 1)) % (N - 1); }
    if (p > q) { int t = p; p = q; q = t; }
    if (q >= n) return;   // padding index of an odd-sized problem
    double *gp = G + (i64)p * ldg, *gq = G + (i64)q * ldg;
    double xp[JR], xq[JR];
    double a = 0.0, b = 0.0, c = 0.0;
#pragma unroll
    for (int k = 0; k < JR; ++k) {
        int row = threadIdx.x + k * JT;
        xp[k] = row < n ? gp[row] : 0.0;
        xq[k] = row < n ? gq[row] : 0.0;
        a = fma(xp[k], xp[k], a);
        b = fma(xq[k], xq[k], b);
        c = fma(xp[k], xq[k], c);
    }
    block_sum3(a, b, c, sh);
    if (!(fabs(c) > tol * sqrt(a * b)) || a == 0.0 || b == 0.0) return;
    const double zeta = (b - a) / (2.0 * c);
    const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
    const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
    if (threadIdx.x == 0) *rotated = 1;
#pragma unroll
    for (int k = 0; k < JR; ++k) {
        int row = threadIdx.x + k * JT;
        if (row < n) {
            gp[row] = cs * xp[k] - sn * xq[k];
            gq[row] = sn * xp[k] + cs * xq[k];
        }
    }
    double *vp = V + (i64)p * ldv, *vq = V + (i64)q * ldv;
#pragma unroll
    for (int k = 0; k < JR; ++k) {
        int row = threadIdx.x + k * JT;
        if (row < n) {
            double y = vp[row], z = vq[row];
            vp[row] = cs * y - sn * z;
            vq[row] = sn * y + cs * z;
        }
    }
}

// ---- block version: one CTA owns two blocks of JB columns (2*JB = 16 columns staged in shared memory) and performs a
// full inner sweep over their 120 pairs (15 rounds of 8 disjoint pairs, one warp per pair) before the next global step.
// A sweep is then nblocks-1 global steps instead of n-1 (n = 520: 65 instead of 519), and the rotations of a step are
// replayed on the 16 matching columns of V row by row from registers.
constexpr int JB = 8;
constexpr int JC = 2 * JB;
constexpr int JROUNDS = JC - 1;
constexpr int JPAIRS = JC / 2;

__device__ __forceinline__ void local_pair(int round, int i, int &a, int &b) {
    if (i == 0) { a = JC - 1; b = round; }
    else { a = (round + i) % (JC - 1); b = (round - i + (JC - 1)) % (JC - 1); }
}

// RPL = rows per lane (n <= 32*RPL): the two columns of a pair live in registers between the dot products and the rotation
template <int RPL>
__global__ void __launch_bounds__(32 * JPAIRS) jacobi_block_step_kernel(double *G, i64 ldg, double *V, i64 ldv, int n, int nblk,
                                                                        int r, double tol, int *rotated) {
    extern __shared__ double sm[];
    double *Gs = sm;                               // [JC][n]
    double2 *rot = reinterpret_cast<double2 *>(sm + (size_t)JC * n);   // [JROUNDS][JPAIRS] (cs, sn)
    __shared__ int any_rot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // block pair of this CTA (round-robin over nblk blocks)
    const int i = blockIdx.x;
    int P, Q;
    if (i == 0) { P = nblk - 1; Q = r; }
    else { P = (r + i) % (nblk - 1); Q = (r - i + (nblk - 1)) % (nblk - 1); }
    if (P > Q) { int t = P; P = Q; Q = t; }
    if (tid == 0) any_rot = 0;
    // global column of local column c
    auto gcol = [&](int c) { return (c < JB ? P * JB + c : Q * JB + (c - JB)); };
    for (int c = 0; c < JC; ++c) {
        const int col = gcol(c);
        for (int row = tid; row < n; row += blockDim.x) Gs[(size_t)c * n + row] = (col < n) ? G[(i64)col * ldg + row] : 0.0;
    }
    __syncthreads();
    for (int round = 0; round < JROUNDS; ++round) {
        int a, b;
        local_pair(round, warp, a, b);
        double *ga = Gs + (size_t)a * n, *gb = Gs + (size_t)b * n;
        double x[RPL], y[RPL];
        double al0 = 0.0, be0 = 0.0, gm0 = 0.0, al1 = 0.0, be1 = 0.0, gm1 = 0.0;
#pragma unroll
        for (int k = 0; k < RPL; ++k) {
            const int row = lane + 32 * k;
            x[k] = row < n ? ga[row] : 0.0;
            y[k] = row < n ? gb[row] : 0.0;
        }
#pragma unroll
        for (int k = 0; k < RPL; ++k) {
            if (k & 1) { al1 = fma(x[k], x[k], al1); be1 = fma(y[k], y[k], be1); gm1 = fma(x[k], y[k], gm1); }
            else       { al0 = fma(x[k], x[k], al0); be0 = fma(y[k], y[k], be0); gm0 = fma(x[k], y[k], gm0); }
        }
        double al = al0 + al1, be = be0 + be1, gm = gm0 + gm1;
        for (int o = 16; o > 0; o >>= 1) {
            al += __shfl_xor_sync(0xffffffffu, al, o);
            be += __shfl_xor_sync(0xffffffffu, be, o);
            gm += __shfl_xor_sync(0xffffffffu, gm, o);
        }
        double cs = 1.0, sn = 0.0;
        if (fabs(gm) > tol * sqrt(al * be) && al != 0.0 && be != 0.0) {
            const double zeta = (be - al) / (2.0 * gm);
            const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            cs = rsqrt(1.0 + t * t); sn = cs * t;
#pragma unroll
            for (int k = 0; k < RPL; ++k) {
                const int row = lane + 32 * k;
                if (row < n) {
                    ga[row] = cs * x[k] - sn * y[k];
                    gb[row] = sn * x[k] + cs * y[k];
                }
            }
            if (lane == 0) any_rot = 1;
        }
        if (lane == 0) rot[round * JPAIRS + warp] = make_double2(cs, sn);
        __syncthreads();
    }
    if (!any_rot) return;     // nothing changed: G and V stay as they are
    if (tid == 0) *rotated = 1;
    for (int c = 0; c < JC; ++c) {
        const int col = gcol(c);
        if (col < n)
            for (int row = tid; row < n; row += blockDim.x) G[(i64)col * ldg + row] = Gs[(size_t)c * n + row];
    }
    // replay the rotations on the matching columns of V, one row per thread, all 16 values in registers
    for (int row = tid; row < n; row += blockDim.x) {
        double v[JC];
#pragma unroll
        for (int c = 0; c < JC; ++c) { const int col = gcol(c); v[c] = (col < n) ? V[(i64)col * ldv + row] : 0.0; }
#pragma unroll
        for (int round = 0; round < JROUNDS; ++round)
#pragma unroll
            for (int pi = 0; pi < JPAIRS; ++pi) {
                int a, b;
                local_pair(round, pi, a, b);
                const double2 cssn = rot[round * JPAIRS + pi];
                const double xv = v[a], yv = v[b];
                v[a] = cssn.x * xv - cssn.y * yv;
                v[b] = cssn.y * xv + cssn.x * yv;
            }
#pragma unroll
        for (int c = 0; c < JC; ++c) { const int col = gcol(c); if (col < n) V[(i64)col * ldv + row] = v[c]; }
    }
}

// sigma[j] = ||G(:,j)||, one warp per column
__global__ void colnorm_kernel(const double *G, i64 ldg, int n, double *sigma) {
    const int lane = threadIdx.x & 31;
    int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n) return;
    double s = 0.0;
    for (int r = lane; r < n; r += 32) { double v = G[(i64)w * ldg + r]; s = fma(v, v, s); }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) sigma[w] = sqrt(s);
}

// U(:,j) = G(:,perm[j]) / sigma[perm[j]];  Vt(j,:) = V(:,perm[j])^T;  s_out[j] = sigma[perm[j]]
__global__ void finalize_kernel(const double *G, i64 ldg, const double *V, i64 ldv, int n, const int *perm, const double *sigma,
                                double *U, i64 ldu, double *Vt, i64 ldvt, double *s_out) {
    const int j = blockIdx.x;
    const int src = perm[j];
    const double sg = sigma[src];
    const double inv = sg > 0.0 ? 1.0 / sg : 0.0;
    for (int r = threadIdx.x; r < n; r += blockDim.x) {
        if (U) U[(i64)j * ldu + r] = G[(i64)src * ldg + r] * inv;
        if (Vt) Vt[(i64)r * ldvt + j] = V[(i64)src * ldv + r];
    }
    if (threadIdx.x == 0 && s_out) s_out[j] = sg;
}

// core: G (n x n, destroyed) -> converged G = U*Sigma, V accumulated.  Returns sweeps used.
int jacobi_core(double *G, i64 ldg, double *V, i64 ldv, int n) {
    Ctx &c = ctx();
    set_identity(V, ldv, n);
    if (n > JT * JR) { set_error("rsvd_b200: Jacobi kernel supports n <= %d (got %d)", JT * JR, n); return -1; }
    if (n < 2) return 0;
    const double tol = 2.220446049250313e-16 * sqrt((double)n);
    int *flag = c.d_flag + 16;
    const int nblk = (((n + JB - 1) / JB) + 1) & ~1;      // even number of column blocks (the last may be padding)
    const size_t smem = (size_t)JC * n * sizeof(double) + (size_t)JROUNDS * JPAIRS * sizeof(double2);
    const bool use_block = nblk >= 4 && smem <= 200 * 1024 && !getenv("RSVD_B200_JACOBI_VECTOR");
    typedef void (*BlockKern)(double *, i64, double *, i64, int, int, int, double, int *);
    BlockKern bk = n <= 32 * 17 ? jacobi_block_step_kernel<17> : (n <= 32 * 33 ? jacobi_block_step_kernel<33> : jacobi_block_step_kernel<40>);
    if (use_block) RSVD_CUDA(cudaFuncSetAttribute(bk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int N = (n + 1) & ~1;
    // capture one sweep (dependent launches) in a graph
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    int launches_per_sweep = 0;
    RSVD_CUDA(cudaStreamBeginCapture(c.stream, cudaStreamCaptureModeThreadLocal));
    if (use_block) {
        for (int r = 0; r < nblk - 1; ++r)
            bk<<<nblk / 2, 32 * JPAIRS, smem, c.stream>>>(G, ldg, V, ldv, n, nblk, r, tol, flag);
        launches_per_sweep = nblk - 1;
    } else {
        for (int r = 0; r < N - 1; ++r)
            jacobi_step_kernel<<<N / 2, JT, 0, c.stream>>>(G, ldg, V, ldv, n, N, r, tol, flag);
        launches_per_sweep = N - 1;
    }
    RSVD_CUDA(cudaStreamEndCapture(c.stream, &graph));
    RSVD_CUDA(cudaGraphInstantiate(&exec, graph, 0));
    int sweeps = 0;
    const int max_sweeps = 40;
    if (exec) {
        for (; sweeps < max_sweeps; ++sweeps) {
            RSVD_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), c.stream));
            RSVD_CUDA(cudaGraphLaunch(exec, c.stream));
            count_launch(launches_per_sweep);
            RSVD_CUDA(cudaMemcpyAsync(c.h_flag + 16, flag, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
            RSVD_CUDA(cudaStreamSynchronize(c.stream));
            if (c.h_flag[16] == 0) { ++sweeps; break; }
        }
        cudaGraphExecDestroy(exec);
    }
    if (graph) cudaGraphDestroy(graph);
    if (c.verbose) fprintf(stderr, "[rsvd_b200] jacobi n=%d sweeps=%d (%s)\n", n, sweeps, use_block ? "block" : "vector");
    return sweeps;
}

}  // namespace

void jacobi_svd(double *A, i64 lda, i64 n_, double *U, i64 ldu, double *s, double *Vt, i64 ldvt) {
    ensure_init();
    Ctx &c = ctx();
    const int n = (int)n_;
    if (n <= 0) return;
    DBuf V((size_t)n * n), sigma((size_t)n);
    if (jacobi_core(A, lda, V.p, n, n) < 0) return;
    colnorm_kernel<<<(n * 32 + 255) / 256, 256, 0, c.stream>>>(A, lda, n, sigma.p);
    count_launch();
    std::vector<double> hs(n);
    RSVD_CUDA(cudaMemcpyAsync(hs.data(), sigma.p, (size_t)n * 8, cudaMemcpyDeviceToHost, c.stream));
    RSVD_CUDA(cudaStreamSynchronize(c.stream));
    std::vector<int> perm(n);
    for (int i = 0; i < n; ++i) perm[i] = i;
    std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return hs[a] > hs[b]; });
    int *dperm = (int *)dalloc_bytes((size_t)n * sizeof(int));
    RSVD_CUDA(cudaMemcpyAsync(dperm, perm.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, c.stream));
    finalize_kernel<<<n, 128, 0, c.stream>>>(A, lda, V.p, n, n, dperm, sigma.p, U, ldu, Vt, ldvt, s);
    count_launch();
    RSVD_CUDA(cudaStreamSynchronize(c.stream));   // perm (host vector) must outlive the copy
    dfree(dperm);
}

// eigenvalue sign: lambda_j = sigma_j * sign(u_j . v_j)
__global__ void eig_sign_kernel(const double *U, i64 ldu, const double *Vt, i64 ldvt, int n, double *w) {
    const int lane = threadIdx.x & 31;
    int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j >= n) return;
    double s = 0.0;
    for (int r = lane; r < n; r += 32) s = fma(U[(i64)j * ldu + r], Vt[(i64)r * ldvt + j], s);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0 && s < 0.0) w[j] = -w[j];
}
// ascending reorder: Aout(:, j) = Vt(n-1-j, :)^T ; wout[j] = w[n-1-j]   (valid for the PSD matrices of the hot path)
__global__ void eig_reverse_kernel(const double *Vt, i64 ldvt, const double *w, int n, double *A, i64 lda, double *wout) {
    const int j = blockIdx.x, src = n - 1 - j;
    for (int r = threadIdx.x; r < n; r += blockDim.x) A[(i64)j * lda + r] = Vt[(i64)r * ldvt + src];
    if (threadIdx.x == 0) wout[j] = w[src];
}

void jacobi_eig(double *A, i64 lda, i64 n_, double *w) {
    ensure_init();
    Ctx &c = ctx();
    const int n = (int)n_;
    if (n <= 0) return;
    DBuf U((size_t)n * n), Vt((size_t)n * n), s((size_t)n);
    jacobi_svd(A, lda, n, U.p, n, s.p, Vt.p, n);
    eig_sign_kernel<<<(n * 32 + 255) / 256, 256, 0, c.stream>>>(U.p, n, Vt.p, n, n, s.p);
    eig_reverse_kernel<<<n, 128, 0, c.stream>>>(Vt.p, n, s.p, n, A, lda, w);
    count_launch(2);
}

}  // namespace rsvd
