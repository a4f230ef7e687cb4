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
// A second pivot rule (mode 2) reproduces the reference's OWN partial Householder QR, pivoted_QR_of_specified_rank[_or_prec]
// (RRA:1012-1155, 1159-1334): squared column norms, plain downdate vn -= R(i,j)^2, strict '>' pivot search, reflector
// with R(i,i) = +||x|| (get_householder_matrix, RRA:982-1008), stop on a (near-)zero pivot norm or, in tolerance mode,
// when sqrt(sum of remaining squared norms)/n (refreshed every 5th step) drops below TOL.
//
// Structure: per Householder step one single-CTA kernel (pivot search, column swap, reflector) and one wide
// kernel applying H_i to all trailing columns and downdating their norms (each column is read and written once).
#include "common.cuh"
#include "ddsum.cuh"

namespace rsvd {

namespace {

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

// mode-2 control block (device memory): steps after `stop` is raised are no-ops
struct RefCtl { int stop; int frank; double r22; double tol; int tolmode; int zero_exact; };

// vn1[j] = sum of squares of A(0:m, j) (mode 2), one warp per column
__global__ void colnorms_sq_kernel(const double *__restrict__ A, i64 lda, i64 m, i64 n, double *vn1) {
    const int lane = threadIdx.x & 31;
    i64 w = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const i64 nw = ((i64)gridDim.x * blockDim.x) >> 5;
    for (i64 j = w; j < n; j += nw) {
        dd a; a.hi = 0.0; a.lo = 0.0;
        const double *col = A + j * lda;
        for (i64 r = lane; r < m; r += 32) a = dd_add_sq(a, col[r]);
        a = dd_warp_sum(a);
        if (lane == 0) vn1[j] = a.hi + a.lo;
    }
}

// Step kernel (single CTA): pivot + swap + reflector.  vbuf receives v = [1; x]; tau -> *tau_out.
// pivoting: 0 none, 1 dgeqp3 rule, 2 the reference's own rule (rc != nullptr).
__global__ void __launch_bounds__(1024) qr_step_kernel(double *A, i64 lda, i64 m, i64 n, i64 i, int pivoting,
                                                       double *vn1, double *vn2, int *jpvt, double *vbuf, double *tau_out,
                                                       RefCtl *rc) {
    __shared__ double shv[64];
    __shared__ i64 shi[32];
    __shared__ i64 pvt_s;
    __shared__ int stop_s;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (pivoting == 2) {
        if (rc->stop) return;
        if (rc->tolmode && i % 5 == 0) {           // RRA:1243-1266: R22norm refreshed every 5th step only
            double sacc = 0.0;
            for (i64 j = i + tid; j < n; j += nt) sacc += vn1[j];
            for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
            if ((tid & 31) == 0) shv[32 + (tid >> 5)] = sacc;
            __syncthreads();
            if (tid == 0) {
                double t = 0.0;
                for (int w = 0; w < (nt >> 5); ++w) t += shv[32 + w];
                rc->r22 = sqrt(t) / (double)n;
            }
            __syncthreads();
        }
    }
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
            stop_s = 0;
            if (pivoting == 2) {                    // RRA:1269-1280 (and 1066-1073 for the fixed-rank variant)
                const bool zero = rc->zero_exact ? (best == 0.0) : (fabs(best) < 1e-10);
                if (zero && i > 0) stop_s = 1;
                else if (rc->tolmode && rc->r22 < rc->tol) stop_s = 1;
                if (stop_s) rc->stop = 1; else rc->frank = (int)i + 1;
            }
        }
        __syncthreads();
        if (stop_s) return;
        const i64 pvt = pvt_s;
        if (pvt != i) {
            for (i64 r = tid; r < m; r += nt) {
                double t = A[pvt * lda + r]; A[pvt * lda + r] = A[i * lda + r]; A[i * lda + r] = t;
            }
            if (tid == 0) {
                int tj = jpvt[pvt]; jpvt[pvt] = jpvt[i]; jpvt[i] = tj;
                if (pivoting == 2) { double tv = vn1[pvt]; vn1[pvt] = vn1[i]; vn1[i] = tv; }
                else { vn1[pvt] = vn1[i]; vn2[pvt] = vn2[i]; }
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
        beta = (pivoting == 2) ? h : ((alpha >= 0.0) ? -h : h);   // mode 2: R(i,i) = +||x|| always (RRA:996)
        if (pivoting == 2 && alpha > 0.0) {
            // alpha - beta without cancellation: -(xnorm^2) / (alpha + beta)
            const double amb = -(xnorm / (alpha + beta)) * xnorm;
            tau = -amb / beta;
            scal = 1.0 / amb;
        } else {
            tau = (beta - alpha) / beta;
            scal = 1.0 / (alpha - beta);
        }
    } else if (pivoting == 2 && alpha < 0.0) {
        beta = -alpha; tau = 2.0;                  // v = 2 alpha e_i: a pure sign flip
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
                                                             double *vn1, double *vn2, const RefCtl *rc) {
    extern __shared__ double vs[];
    if (rc && rc->stop) return;
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
        if (pivoting == 2) {
            if (lane == 0) vn1[j] -= a[0] * a[0];                   // RRA:1316-1319: plain downdate, never recomputed
        } else if (pivoting) {
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

// Tall columns (more than 1280 rows): one CTA per column, two passes (dot, update).  The partial-norm downdate of either
// pivot rule rides on the update pass: the new A(i,j) and a double-double sum of squares of the rows below it.
__global__ void __launch_bounds__(256) qr_apply_tall_kernel(double *A, i64 lda, i64 m, i64 n, i64 i, int pivoting, int ps_formula,
                                                            const double *__restrict__ vbuf, const double *__restrict__ tau_p,
                                                            double *vn1, double *vn2, const RefCtl *rc) {
    __shared__ double sh[64];
    __shared__ double dot_s, a0_s;
    if (rc && rc->stop) return;
    const double tau = *tau_p;
    if (tau == 0.0 && !pivoting) return;
    const i64 len = m - i;
    const double tol3z = 1.0536712127723509e-08;   // sqrt(dlamch('Epsilon'))
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
        dd acc; acc.hi = 0.0; acc.lo = 0.0;
        for (i64 r = threadIdx.x; r < len; r += blockDim.x) {
            const double v = fma(-f, vbuf[r], col[r]);
            if (tau != 0.0) col[r] = v;
            if (r == 0) a0_s = v; else acc = dd_add_sq(acc, v);
        }
        if (pivoting == 2) {
            if (threadIdx.x == 0) vn1[j] -= a0_s * a0_s;             // thread 0 owns r == 0
        } else if (pivoting == 1) {
            __syncthreads();
            const double v1 = vn1[j], v2 = vn2[j];
            if (v1 != 0.0) {
                const double t = fabs(a0_s) / v1;
                const double temp = ps_formula ? fmax(0.0, (1.0 + t) * (1.0 - t)) : fmax(1.0 - t * t, 0.0);
                const double q2 = v1 / v2;
                const double temp2 = temp * (q2 * q2);
                if (temp2 <= tol3z) {                                // uniform across the CTA
                    const double nv = block_nrm2(acc, sh);
                    __syncthreads();
                    if (threadIdx.x == 0) { const double w = (len > 1) ? nv : 0.0; vn1[j] = w; vn2[j] = w; }
                } else {
                    __syncthreads();
                    if (threadIdx.x == 0) vn1[j] = v1 * sqrt(temp);
                }
            }
        }
    }
}

// Q(:, j) <- (I - tau v v^T) Q(:, j) for j = i..f-1, v = [1; A(i+1:m, i)] (dorg2r order: called for i = f-1 .. 0)
__global__ void __launch_bounds__(256) form_q_kernel(const double *__restrict__ A, i64 lda, i64 m, i64 i, const double *__restrict__ tau_p,
                                                     double *Q, i64 ldq) {
    __shared__ double sh[32];
    __shared__ double dot_s;
    const double tau = tau_p[i];
    if (tau == 0.0) return;
    const i64 j = i + blockIdx.x;
    double *q = Q + j * ldq;
    const double *v = A + i * lda;
    double dot = 0.0;
    for (i64 r = i + threadIdx.x; r < m; r += blockDim.x) dot = fma(r == i ? 1.0 : v[r], q[r], dot);
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = dot;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
        dot_s = t;
    }
    __syncthreads();
    const double f = tau * dot_s;
    for (i64 r = i + threadIdx.x; r < m; r += blockDim.x) q[r] = fma(-f, r == i ? 1.0 : v[r], q[r]);
}

// R(r, j) = A(r, j) for r <= j, 0 below the diagonal; r < f
__global__ void extract_r_kernel(const double *__restrict__ A, i64 lda, i64 f, i64 n, double *R, i64 ldr) {
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < f * n; e += (i64)gridDim.x * blockDim.x) {
        const i64 j = e / f, r = e % f;
        R[j * ldr + r] = (r <= j) ? A[j * lda + r] : 0.0;
    }
}

__global__ void iota_kernel(int *p, i64 n) {
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (i64)gridDim.x * blockDim.x) p[e] = (int)e;
}
__global__ void int_to_double_kernel(const int *p, double *d, i64 n) {
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (i64)gridDim.x * blockDim.x) d[e] = (double)p[e];
}

// steps: number of Householder steps (<= min(m,n)).  Mode 2 extras: rc_host (tolmode/tol/zero_exact in, frank out) and
// the explicit factors Qout (m x frank) / Rout (frank x n), written for the first frank steps only.
struct RefOpts { int tolmode; double tol; int zero_exact; i64 frank; double *Q; i64 ldq; double *R; i64 ldr; };

// qr_out (pivoting == 0 only): after the factorisation A is replaced by the explicit thin Q and R (n x n upper) goes to Rq.
void householder_qr(double *A, i64 lda, i64 m, i64 n, int pivoting, double *jpvt_out, i64 steps = -1, RefOpts *ro = nullptr,
                    bool q_out = false, double *Rq = nullptr, i64 ldrq = 0, double *tau_out = nullptr) {
    if (g_status) return;   // an earlier error (e.g. a failed allocation) is pending: launch nothing
    ensure_init();
    Ctx &c = ctx();
    const i64 minmn = min(m, n);
    if (minmn <= 0) return;
    if (steps < 0 || steps > minmn) steps = minmn;
    DBuf vn1((size_t)n), vn2((size_t)n), vbuf((size_t)m + 8), tau((size_t)steps + 1);
    int *jpvt = (int *)dalloc_bytes((size_t)n * sizeof(int));
    RefCtl *rc = nullptr;
    const int gsz = (int)min((i64)c.sms * 8, (n + 255) / 256);
    if (pivoting) {
        iota_kernel<<<max(gsz, 1), 256, 0, c.stream>>>(jpvt, n);
        int nb = (int)min((i64)c.sms * 8, (n + 7) / 8);
        if (pivoting == 2) colnorms_sq_kernel<<<max(nb, 1), 256, 0, c.stream>>>(A, lda, m, n, vn1.p);
        else colnorms_kernel<<<max(nb, 1), 256, 0, c.stream>>>(A, lda, m, n, vn1.p, vn2.p);
        count_launch(2);
    }
    if (pivoting == 2) {
        rc = (RefCtl *)dalloc_bytes(sizeof(RefCtl));
        RefCtl h; h.stop = 0; h.frank = 0; h.r22 = 0.0; h.tol = ro->tol; h.tolmode = ro->tolmode; h.zero_exact = ro->zero_exact;
        RSVD_CUDA(cudaMemcpyAsync(rc, &h, sizeof(h), cudaMemcpyHostToDevice, c.stream));
        RSVD_CUDA(cudaStreamSynchronize(c.stream));   // h is a stack object
        RSVD_CUDA(cudaMemsetAsync(tau.p, 0, (size_t)(steps + 1) * 8, c.stream));
    }
    const bool blocked_range = minmn > 128;   // dgeqp3: NB=32 < sminmn and NX=128 < sminmn
    for (i64 i = 0; i < steps; ++i) {
        qr_step_kernel<<<1, 1024, 0, c.stream>>>(A, lda, m, n, i, pivoting, vn1.p, vn2.p, jpvt, vbuf.p, tau.p + i, rc);
        count_launch();
        const i64 ncols = n - i - 1;
        if (ncols <= 0) continue;
        const i64 len = m - i;
        const int ps = (blocked_range && i < minmn - 128) ? 1 : 0;
        if (len <= 32 * 40) {
            int blocks = (int)min((i64)c.sms * 8, (ncols + 7) / 8);
            size_t sh = (size_t)len * 8;
            if (len <= 32 * 4) qr_apply_short_kernel<4><<<blocks, 256, sh, c.stream>>>(A, lda, m, n, i, pivoting, ps, vbuf.p, tau.p + i, vn1.p, vn2.p, rc);
            else if (len <= 32 * 12) qr_apply_short_kernel<12><<<blocks, 256, sh, c.stream>>>(A, lda, m, n, i, pivoting, ps, vbuf.p, tau.p + i, vn1.p, vn2.p, rc);
            else if (len <= 32 * 24) qr_apply_short_kernel<24><<<blocks, 256, sh, c.stream>>>(A, lda, m, n, i, pivoting, ps, vbuf.p, tau.p + i, vn1.p, vn2.p, rc);
            else qr_apply_short_kernel<40><<<blocks, 256, sh, c.stream>>>(A, lda, m, n, i, pivoting, ps, vbuf.p, tau.p + i, vn1.p, vn2.p, rc);
        } else {
            int blocks = (int)min(ncols, (i64)c.sms * 8);
            qr_apply_tall_kernel<<<blocks, 256, 0, c.stream>>>(A, lda, m, n, i, pivoting, ps, vbuf.p, tau.p + i, vn1.p, vn2.p, rc);
        }
        count_launch();
    }
    if (pivoting && jpvt_out) {
        int_to_double_kernel<<<max(gsz, 1), 256, 0, c.stream>>>(jpvt, jpvt_out, n);
        count_launch();
    }
    if (pivoting == 2) {
        RefCtl h;
        RSVD_CUDA(cudaMemcpyAsync(&h, rc, sizeof(h), cudaMemcpyDeviceToHost, c.stream));
        RSVD_CUDA(cudaStreamSynchronize(c.stream));
        const i64 f = g_status ? 0 : h.frank;
        ro->frank = f;
        if (f > 0 && ro->R) {
            extract_r_kernel<<<(int)min((i64)c.sms * 8, (f * n + 255) / 256), 256, 0, c.stream>>>(A, lda, f, n, ro->R, ro->ldr);
            count_launch();
        }
        if (f > 0 && ro->Q) {
            set_zero(ro->Q, (size_t)ro->ldq * f);
            set_identity(ro->Q, ro->ldq, f);
            for (i64 i = f - 1; i >= 0; --i) {
                form_q_kernel<<<(int)(f - i), 256, 0, c.stream>>>(A, lda, m, i, tau.p, ro->Q, ro->ldq);
                count_launch();
            }
        }
        dfree(rc);
    }
    if (q_out && pivoting == 0 && !g_status) {
        const i64 f = steps;
        if (Rq) {
            extract_r_kernel<<<(int)min((i64)c.sms * 8, (f * n + 255) / 256), 256, 0, c.stream>>>(A, lda, f, n, Rq, ldrq);
            count_launch();
        }
        DBuf Qb((size_t)m * f);
        set_zero(Qb.p, (size_t)m * f);
        set_identity(Qb.p, m, f);
        for (i64 i = f - 1; i >= 0; --i) {
            form_q_kernel<<<(int)(f - i), 256, 0, c.stream>>>(A, lda, m, i, tau.p, Qb.p, m);
            count_launch();
        }
        copy_matrix(Qb.p, m, A, lda, m, f);
    }
    if (tau_out) copy_matrix(tau.p, steps, tau_out, steps, steps, 1);
    RSVD_CUDA(cudaGetLastError());
    dfree(jpvt);
}

}  // namespace

// short-and-wide matrices: the blocked (dlaqps-style) kernel; tall ones: one reflector per step (dlaqp2-style)
void geqp3(double *A, i64 lda, i64 m, i64 n, double *jpvt_out) {
    if (geqp3_blocked_ok(m, n) && !ctx().force_unblocked_qr) { geqp3_blocked(A, lda, m, n, jpvt_out, nullptr); return; }
    householder_qr(A, lda, m, n, 1, jpvt_out);
}

// pivotedQR_mkl (RRA:924-976): dgeqp3 followed by dorgqr — the explicit Q is built from the Householder reflectors, so it
// is orthonormal whatever the conditioning (or rank) of the input.
void geqp3_q(double *A, i64 lda, i64 m, i64 n, double *jpvt_out, double *Q, i64 ldq) {
    if (g_status) return;
    ensure_init();
    Ctx &c = ctx();
    const i64 f = min(m, n);
    if (f <= 0) return;
    DBuf tau((size_t)f + 1);
    if (geqp3_blocked_ok(m, n) && !c.force_unblocked_qr) geqp3_blocked(A, lda, m, n, jpvt_out, tau.p);
    else householder_qr(A, lda, m, n, 1, jpvt_out, -1, nullptr, false, nullptr, 0, tau.p);
    if (g_status) return;
    set_zero(Q, (size_t)ldq * f);
    set_identity(Q, ldq, f);
    for (i64 i = f - 1; i >= 0; --i) {                    // dorg2r order
        form_q_kernel<<<(int)(f - i), 256, 0, c.stream>>>(A, lda, m, i, tau.p, Q, ldq);
        count_launch();
    }
}
void geqrf_r(double *A, i64 lda, i64 m, i64 n) { householder_qr(A, lda, m, n, 0, nullptr); }
void geqrf_q(double *A, i64 lda, i64 m, i64 n, double *R, i64 ldr) {
    if (m < n) { set_error("rsvd_b200: explicit-Q Householder QR needs m >= n (got %lld x %lld)", (long long)m, (long long)n); return; }
    householder_qr(A, lda, m, n, 0, nullptr, -1, nullptr, true, R, ldr);
}

// The reference's own partial pivoted QR (RRA:1012-1155 with zero_exact = 1, RRA:1159-1334 with zero_exact = 0).
// k > 0: rank mode (at most k steps); k <= 0: tolerance mode (at most min(m,n) steps, stop when R22norm < tol).
// A (m x n) is destroyed.  Returns frank; I (n doubles, 0-based), Q (m x frank), R (frank x n) may be null.
i64 pqr_partial(double *A, i64 lda, i64 m, i64 n, i64 k, double tol, int zero_exact, double *I, double *Q, i64 ldq, double *R, i64 ldr) {
    RefOpts ro;
    ro.tolmode = k <= 0; ro.tol = tol; ro.zero_exact = zero_exact; ro.frank = 0; ro.Q = Q; ro.ldq = ldq; ro.R = R; ro.ldr = ldr;
    const i64 steps = (k <= 0) ? min(m, n) : min(k, min(m, n));
    householder_qr(A, lda, m, n, 2, I, steps, &ro);
    return ro.frank;
}

}  // namespace rsvd
