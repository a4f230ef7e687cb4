// geqp3.cu — column-pivoted Householder QR with LAPACK-dgeqp3-compatible pivoting, and the unpivoted R-only
// Householder factorisation used by the TSQR fallback.
// Replaces pivotedQR_mkl (rank_revealing_algorithms_intel_mkl.c:924-976: LAPACKE_dgeqp3 + dorgqr).  On the hot
// path only R and the permutation are consumed (RRA:1940-1949, 1839-1843), so Q is never formed.
//
// Pivot rule replicated from dgeqp3/dlaqps/dlaqp2 (LAPACK 3.9 as shipped in OpenBLAS 0.3.15):
//   vn1 = vn2 = ||A(:,j)||; at step i the pivot is the FIRST index of max vn1(i:n) (idamax);
//   after the reflector: t = |A(i,j)|/vn1(j); temp = max(0,(1+t)(1-t)) in the blocked (dlaqps) range of steps and
//   max(0,1-t^2) in the unblocked tail (last 128 steps; all steps when min(m,n) <= 128);
//   temp2 = temp*(vn1/vn2)^2; temp2 <= sqrt(eps) => recompute both norms from A(i+1:m,j), else vn1 *= sqrt(temp).
// Norms are accumulated in double-double so they are (nearly) correctly rounded like OpenBLAS's extended-precision
// dnrm2 kernel; pivot choice is by warp/block-reduced norms with a first-index tie-break.
//
// Structure: per Householder step one single-CTA kernel (pivot search, column swap, reflector) and one wide
// kernel applying H_i to all trailing columns and downdating their norms (each column is read and written once).
#include "common.cuh"

namespace rsvd {

namespace {

struct dd { double hi, lo; };
__device__ __forceinline__ dd dd_add_sq(dd a, double x) {   // a += x*x, error-free product + two-sum
    double p = x * x;
    double e = fma(x, x, -p);
    double s = a.hi + p;
    double bb = s - a.hi;
    double err = (a.hi - (s - bb)) + (p - bb);
    a.hi = s; a.lo += err + e;
    return a;
}
__device__ __forceinline__ dd dd_add(dd a, dd b) {
    double s = a.hi + b.hi;
    double bb = s - a.hi;
    double err = (a.hi - (s - bb)) + (b.hi - bb);
    dd r; r.hi = s; r.lo = a.lo + b.lo + err;
    return r;
}
__device__ __forceinline__ dd dd_warp_sum(dd a) {
    for (int o = 16; o > 0; o >>= 1) {
        dd b;
        b.hi = __shfl_xor_sync(0xffffffffu, a.hi, o);
        b.lo = __shfl_xor_sync(0xffffffffu, a.lo, o);
        a = dd_add(a, b);
    }
    return a;
}
__device__ __forceinline__ double dd_sqrt(dd a) { return sqrt(a.hi + a.lo); }

// block-wide dd sum; result valid in all threads. sh must hold 2*32 doubles.
__device__ double block_nrm2(dd a, double *sh) {
    a = dd_warp_sum(a);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 0) { sh[2 * w] = a.hi; sh[2 * w + 1] = a.lo; }
    __syncthreads();
    dd t; t.hi = 0.0; t.lo = 0.0;
    for (int i = 0; i < nw; ++i) { dd b; b.hi = sh[2 * i]; b.lo = sh[2 * i + 1]; t = dd_add(t, b); }
    return dd_sqrt(t);
}

// vn1[j] = vn2[j] = ||A(0:m, j)||, one warp per column
__global__ void colnorms_kernel(const double *__restrict__ A, i64 lda, i64 m, i64 n, double *vn1, double *vn2) {
    const int lane = threadIdx.x & 31;
    i64 w = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const i64 nw = ((i64)gridDim.x * blockDim.x) >> 5;
    for (i64 j = w; j < n; j += nw) {
        dd a; a.hi = 0.0; a.lo = 0.0;
        const double *col = A + j * lda;
        for (i64 r = lane; r < m; r += 32) a = dd_add_sq(a, col[r]);
        a = dd_warp_sum(a);
        if (lane == 0) { double v = dd_sqrt(a); vn1[j] = v; vn2[j] = v; }
    }
}

// Step kernel (single CTA): pivot + swap + reflector.  vbuf receives v = [1; x]; tau -> *tau_out.
__global__ void __launch_bounds__(1024) qr_step_kernel(double *A, i64 lda, i64 m, i64 n, i64 i, int pivoting,
                                                       double *vn1, double *vn2, int *jpvt, double *vbuf, double *tau_out) {
    __shared__ double shv[64];
    __shared__ i64 shi[32];
    __shared__ i64 pvt_s;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (pivoting) {
        double best = -1.0; i64 bi = i;
        for (i64 j = i + tid; j < n; j += nt) {
            double v = vn1[j];
            if (v > best) { best = v; bi = j; }   // strictly greater: keeps the first index within a thread's stride...
        }
        // ... and (value desc, index asc) ordering across threads
        for (int o = 16; o > 0; o >>= 1) {
            double ov = __shfl_xor_sync(0xffffffffu, best, o);
            i64 oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if ((tid & 31) == 0) { shv[tid >> 5] = best; shi[tid >> 5] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < (nt >> 5); ++w)
                if (shv[w] > best || (shv[w] == best && shi[w] < bi)) { best = shv[w]; bi = shi[w]; }
            pvt_s = bi;
        }
        __syncthreads();
        const i64 pvt = pvt_s;
        if (pvt != i) {
            for (i64 r = tid; r < m; r += nt) {
                double t = A[pvt * lda + r]; A[pvt * lda + r] = A[i * lda + r]; A[i * lda + r] = t;
            }
            if (tid == 0) {
                int tj = jpvt[pvt]; jpvt[pvt] = jpvt[i]; jpvt[i] = tj;
                vn1[pvt] = vn1[i]; vn2[pvt] = vn2[i];
            }
        }
        __syncthreads();
    }
    // reflector for A(i:m, i)   (dlarfg)
    double *col = A + i * lda;
    dd acc; acc.hi = 0.0; acc.lo = 0.0;
    for (i64 r = i + 1 + tid; r < m; r += nt) acc = dd_add_sq(acc, col[r]);
    const double xnorm = block_nrm2(acc, shv);
    const double alpha = col[i];
    double tau = 0.0, scal = 0.0, beta = alpha;
    if (xnorm != 0.0) {
        // beta = -sign(dlapy2(alpha, xnorm), alpha)
        double aa = fabs(alpha), xx = fabs(xnorm);
        double w = fmax(aa, xx), z = fmin(aa, xx);
        double h = (z == 0.0) ? w : w * sqrt(1.0 + (z / w) * (z / w));
        beta = (alpha >= 0.0) ? -h : h;
        tau = (beta - alpha) / beta;
        scal = 1.0 / (alpha - beta);
    }
    __syncthreads();
    for (i64 r = i + 1 + tid; r < m; r += nt) {
        double x = col[r] * scal;
        if (xnorm != 0.0) col[r] = x;
        vbuf[r - i] = (xnorm != 0.0) ? x : 0.0;
    }
    if (tid == 0) {
        vbuf[0] = 1.0;
        col[i] = beta;
        *tau_out = tau;
    }
}

// Apply H = I - tau v v^T to trailing columns j > i, rows i..m-1; one warp per column, the column lives in
// registers (<= 32*RPL rows).  Then downdate the partial norms (pivoting only).
template <int RPL>
__global__ void __launch_bounds__(256) qr_apply_short_kernel(double *A, i64 lda, i64 m, i64 n, i64 i, int pivoting, int ps_formula,
                                                             const double *__restrict__ vbuf, const double *__restrict__ tau_p,
                                                             double *vn1, double *vn2) {
    extern __shared__ double vs[];
    const int len = (int)(m - i);
    for (int r = threadIdx.x; r < len; r += blockDim.x) vs[r] = vbuf[r];
    __syncthreads();
    const double tau = *tau_p;
    const int lane = threadIdx.x & 31;
    i64 w = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const i64 nw = ((i64)gridDim.x * blockDim.x) >> 5;
    const double tol3z = 1.0536712127723509e-08;   // sqrt(dlamch('Epsilon')) = sqrt(2^-53)
    for (i64 j = i + 1 + w; j < n; j += nw) {
        double *col = A + j * lda + i;
        double a[RPL];
        double dot = 0.0;
#pragma unroll
        for (int q = 0; q < RPL; ++q) {
            int r = lane + 32 * q;
            a[q] = (r < len) ? col[r] : 0.0;
            dot = fma(a[q], (r < len) ? vs[r] : 0.0, dot);
        }
        for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        const double f = tau * dot;
        dd acc; acc.hi = 0.0; acc.lo = 0.0;
#pragma unroll
        for (int q = 0; q < RPL; ++q) {
            int r = lane + 32 * q;
            if (r < len) {
                a[q] = fma(-f, vs[r], a[q]);
                if (tau != 0.0) col[r] = a[q];
                if (r >= 1) acc = dd_add_sq(acc, a[q]);
            }
        }
        if (pivoting) {
            const double aij = __shfl_sync(0xffffffffu, a[0], 0);   // new A(i,j)
            const double v1 = vn1[j], v2 = vn2[j];
            if (v1 != 0.0) {
                double t = fabs(aij) / v1;
                double temp = ps_formula ? fmax(0.0, (1.0 + t) * (1.0 - t)) : fmax(1.0 - t * t, 0.0);
                double q2 = v1 / v2;
                double temp2 = temp * (q2 * q2);
                if (temp2 <= tol3z) {
                    acc = dd_warp_sum(acc);
                    if (lane == 0) {
                        double nv = (len > 1) ? dd_sqrt(acc) : 0.0;
                        vn1[j] = nv; vn2[j] = nv;
                    }
                } else if (lane == 0) {
                    vn1[j] = v1 * sqrt(temp);
                }
            }
        }
    }
}

// Tall columns: one CTA per column, two passes.  Unpivoted use only.
__global__ void __launch_bounds__(256) qr_apply_tall_kernel(double *A, i64 lda, i64 m, i64 n, i64 i,
                                                            const double *__restrict__ vbuf, const double *__restrict__ tau_p) {
    __shared__ double sh[32];
    __shared__ double dot_s;
    const double tau = *tau_p;
    if (tau == 0.0) return;
    const i64 len = m - i;
    for (i64 j = i + 1 + blockIdx.x; j < n; j += gridDim.x) {
        double *col = A + j * lda + i;
        double dot = 0.0;
        for (i64 r = threadIdx.x; r < len; r += blockDim.x) dot = fma(col[r], vbuf[r], dot);
        for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = dot;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sh[w];
            dot_s = s;
        }
        __syncthreads();
        const double f = tau * dot_s;
        for (i64 r = threadIdx.x; r < len; r += blockDim.x) col[r] = fma(-f, vbuf[r], col[r]);
    }
}

__global__ void iota_kernel(int *p, i64 n) {
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (i64)gridDim.x * blockDim.x) p[e] = (int)e;
}
__global__ void int_to_double_kernel(const int *p, double *d, i64 n) {
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (i64)gridDim.x * blockDim.x) d[e] = (double)p[e];
}

void householder_qr(double *A, i64 lda, i64 m, i64 n, int pivoting, double *jpvt_out) {
    if (g_status) return;   // an earlier error (e.g. a failed allocation) is pending: launch nothing
    ensure_init();
    Ctx &c = ctx();
    const i64 minmn = min(m, n);
    if (minmn <= 0) return;
    DBuf vn1((size_t)n), vn2((size_t)n), vbuf((size_t)m + 8), tau(1);
    int *jpvt = (int *)dalloc_bytes((size_t)n * sizeof(int));
    const int gsz = (int)min((i64)c.sms * 8, (n + 255) / 256);
    if (pivoting) {
        iota_kernel<<<max(gsz, 1), 256, 0, c.stream>>>(jpvt, n);
        int nb = (int)min((i64)c.sms * 8, (n + 7) / 8);
        colnorms_kernel<<<max(nb, 1), 256, 0, c.stream>>>(A, lda, m, n, vn1.p, vn2.p);
        count_launch(2);
    }
    const bool blocked_range = minmn > 128;   // dgeqp3: NB=32 < sminmn and NX=128 < sminmn
    for (i64 i = 0; i < minmn; ++i) {
        qr_step_kernel<<<1, 1024, 0, c.stream>>>(A, lda, m, n, i, pivoting, vn1.p, vn2.p, jpvt, vbuf.p, tau.p);
        count_launch();
        const i64 ncols = n - i - 1;
        if (ncols <= 0) continue;
        const i64 len = m - i;
        const int ps = (blocked_range && i < minmn - 128) ? 1 : 0;
        if (len <= 32 * 40) {
            int blocks = (int)min((i64)c.sms * 8, (ncols + 7) / 8);
            size_t sh = (size_t)len * 8;
            if (len <= 32 * 4) qr_apply_short_kernel<4><<<blocks, 256, sh, c.stream>>>(A, lda, m, n, i, pivoting, ps, vbuf.p, tau.p, vn1.p, vn2.p);
            else if (len <= 32 * 12) qr_apply_short_kernel<12><<<blocks, 256, sh, c.stream>>>(A, lda, m, n, i, pivoting, ps, vbuf.p, tau.p, vn1.p, vn2.p);
            else if (len <= 32 * 24) qr_apply_short_kernel<24><<<blocks, 256, sh, c.stream>>>(A, lda, m, n, i, pivoting, ps, vbuf.p, tau.p, vn1.p, vn2.p);
            else qr_apply_short_kernel<40><<<blocks, 256, sh, c.stream>>>(A, lda, m, n, i, pivoting, ps, vbuf.p, tau.p, vn1.p, vn2.p);
        } else {
            if (pivoting) { set_error("rsvd_b200: pivoted QR supports at most 1280 rows (got %lld)", (long long)m); break; }
            int blocks = (int)min(ncols, (i64)c.sms * 8);
            qr_apply_tall_kernel<<<blocks, 256, 0, c.stream>>>(A, lda, m, n, i, vbuf.p, tau.p);
        }
        count_launch();
    }
    if (pivoting && jpvt_out) {
        int_to_double_kernel<<<max(gsz, 1), 256, 0, c.stream>>>(jpvt, jpvt_out, n);
        count_launch();
    }
    RSVD_CUDA(cudaGetLastError());
    dfree(jpvt);
}

}  // namespace

void geqp3(double *A, i64 lda, i64 m, i64 n, double *jpvt_out) { householder_qr(A, lda, m, n, 1, jpvt_out); }
void geqrf_r(double *A, i64 lda, i64 m, i64 n) { householder_qr(A, lda, m, n, 0, nullptr); }

}  // namespace rsvd
