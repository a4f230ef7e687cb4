// cholinv.cu — the l x l step of a Cholesky-QR pass in ONE kernel: upper Cholesky factor G = R^T R, its inverse R^{-1}, and
// min/max of diag(R) for the condition test.  Replaces the per-block sequence potf2 / panel GEMM / copy / trailing GEMM
// (9 x 4 launches at l = 520) + diag_minmax + trtri_diag + 8 recursive-doubling GEMMs of cholqr.cu: at l = 520 that sequence
// took ~1.1 ms per pass, six passes per randomized SVD (dgeqrf+dorgqr in the reference, MVF:1251-1263, called from
// RRA:106,118,130,146), almost all of it launch latency and single-CTA dependent chains.
//
// Dataflow formulation.  The matrix is cut into 32 x 32 blocks; block column c owns the tasks
//     C(r, c), r = 0..c :  acc = G(r,c) - sum_{j<r} R(j,r)^T R(j,c);   r == c: R(c,c) = chol(acc), W_c = R(c,c)^{-1} (one warp,
//                          registers + shuffles);   r < c: R(r,c) = W_r^T acc
//     X(r, c), r = c-1..0: X(r,c) = -( sum_{i=r}^{c-1} X(r,i) R(i,c) ) W_c          (X = R^{-1}, X(c,c) = W_c)
// Tasks are numbered column by column (c^2 + i), so every dependency of a task has a smaller number; CTA p of a co-resident
// (cooperative) grid of P CTAs executes tasks p, p+P, ... in order and waits on per-block ready flags (release/acquire through
// L2) — no device-wide barriers, no host round trips; everything that does not sit on the critical path
// (potf2 -> panel block -> next diagonal update) runs in its shadow, including the whole inverse.
// The diagonal task of column c additionally repeats the accumulation of C(c-1, c), so that the chain potf2 -> R(c-1,c) -> next
// diagonal block costs one hop through L2 per block column, not two.
// Deadlock-free: the smallest unfinished task is always being executed (its CTA has finished its earlier tasks) and all of
// its inputs are finished tasks.  Waits are bounded spins that raise an error flag instead of hanging the GPU.
#include "common.cuh"

namespace rsvd {

namespace {

constexpr int CB = 32;    // block size
constexpr int CS = 34;    // shared-memory column stride in doubles: 16-byte aligned columns, conflict-free LDS.128 across 8 columns
constexpr int CT = 128;   // threads per CTA: thread (ty, tx) = (tid / 8, tid % 8) owns outputs (ty + 16 i, tx + 8 j), i < 2, j < 4

struct CholStatus {
    int fail;       // first non-positive pivot met: global column + 1 (0 = none)
    int err;        // a wait timed out (grid not making progress)
    double dmin, dmax;   // min / max of diag(R)
};

__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

__device__ __forceinline__ double rsqrt1(double x) {      // 1/sqrt(x) to ~1 ulp for normal x: ~20-bit seed + one third-order step
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    const double e = fma(-x * y0, y0, 1.0);
    return fma(y0 * e, fma(0.375, e, 0.5), y0);            // y0 e and (0.5 + 0.375 e) in parallel: four dependent operations after the seed
}

// Threads 0 and 32 (two warps) poll one flag each; everybody leaves through the barrier.  false = give up.
__device__ __forceinline__ bool wait2(const int *fa, const int *fb, int *err, int *sh_ok) {
    if (threadIdx.x == 0 || threadIdx.x == 32) {
        const int *f = (threadIdx.x == 0) ? fa : fb;
        if (f) {
            long long spins = 0;
            while (ld_acquire(f) == 0) {
                if ((++spins & 255) == 0) {
                    if (spins > (1ll << 23) || *((volatile int *)err)) { *err = 1; *sh_ok = 0; break; }
                    __nanosleep(40);
                }
            }
        }
    }
    __syncthreads();
    return *((volatile int *)sh_ok) != 0;
}

// global block (br, bc) of the n x n column-major matrix M -> S (column-major with stride CS, or transposed).
// Entries outside the matrix are those of the identity (padded diagonal blocks stay trivially factorable).
__device__ __forceinline__ void load_block(double *S, const double *M, i64 ld, int br, int bc, int n, bool transposed) {
    const int i = threadIdx.x & 31, j0 = threadIdx.x >> 5;
    const int gi = br * CB + i;
#pragma unroll
    for (int e = 0; e < CB / 4; ++e) {
        const int j = j0 + 4 * e, gj = bc * CB + j;
        const double v = (gi < n && gj < n) ? __ldcg(M + (i64)gj * ld + gi) : (gi == gj ? 1.0 : 0.0);
        S[transposed ? (j + CS * i) : (i + CS * j)] = v;
    }
}
// S (column-major, stride CS) -> global block (br, bc); `zero_mirror`: also clear the mirrored block (bc, br) below the diagonal
__device__ __forceinline__ void store_block(const double *S, double *M, i64 ld, int br, int bc, int n, bool zero_mirror) {
    const int i = threadIdx.x & 31, j0 = threadIdx.x >> 5;
    const int gi = br * CB + i;
#pragma unroll
    for (int e = 0; e < CB / 4; ++e) {
        const int j = j0 + 4 * e, gj = bc * CB + j;
        if (gi < n && gj < n) M[(i64)gj * ld + gi] = S[i + CS * j];
    }
    if (zero_mirror) {
        const int mi = bc * CB + i;
#pragma unroll
        for (int e = 0; e < CB / 4; ++e) {
            const int mj = br * CB + j0 + 4 * e;
            if (mi < n && mj < n) M[(i64)mj * ld + mi] = 0.0;
        }
    }
}

// acc(a, b) += sgn * sum_k As[k + CS a] * Bs[k + CS b]
template <int SGN>
__device__ __forceinline__ void mma32(double (&acc)[2][4], const double *As, const double *Bs) {
    const int ty = threadIdx.x >> 3, tx = threadIdx.x & 7;
#pragma unroll 4
    for (int k = 0; k < CB; k += 2) {
        double2 a[2], b[4];
#pragma unroll
        for (int i = 0; i < 2; ++i) a[i] = *reinterpret_cast<const double2 *>(As + k + CS * (ty + 16 * i));
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const double2 *>(Bs + k + CS * (tx + 8 * j));
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (SGN > 0) { acc[i][j] = fma(a[i].x, b[j].x, acc[i][j]); acc[i][j] = fma(a[i].y, b[j].y, acc[i][j]); }
                else { acc[i][j] = fma(-a[i].x, b[j].x, acc[i][j]); acc[i][j] = fma(-a[i].y, b[j].y, acc[i][j]); }
            }
    }
}
__device__ __forceinline__ void acc_to_smem(const double (&acc)[2][4], double *S, bool transposed) {
    const int ty = threadIdx.x >> 3, tx = threadIdx.x & 7;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int a = ty + 16 * i, b = tx + 8 * j;
            S[transposed ? (b + CS * a) : (a + CS * b)] = acc[i][j];
        }
}

// One warp: Cholesky factor and inverse of the 32 x 32 block whose UPPER triangle is in D (column-major, stride CS).
// Factor: lane r carries row r of the lower factor L = R^T in registers.  Per column: one shuffle broadcasts the pivot, every
// lane publishes its scaled entry of column c in shared memory, one __syncwarp, and the rank-1 update reads the column back as
// broadcast LDS (double-buffered, so one sync per column).  The chain per column is pivot -> reciprocal square root -> scale ->
// update of the next pivot; nothing else sits on it.  Inverse: afterwards lane j solves L x = e_j by forward substitution
// (column j of L^{-1}; rows of L are broadcast reads from shared memory) — 32 light steps instead of carrying a second
// register matrix through the factorisation (fewer registers: 208 -> 128, so larger matrices keep 3-4 CTAs per SM: 2.59 -> 1.75 ms at n = 2048).
// A column step is a template so that every register index is a compile-time constant; failure handling is branch-free
// (a divergent branch makes ptxas wrap every later shuffle in WARPSYNC/ENDCOLLECTIVE pairs).
// Out: Rs = R (upper, zeros below), Ws = R^{-1} (upper, zeros below); both column-major with stride CS.
template <int c>
__device__ __forceinline__ void potf2_col(double (&L)[CB], double pv, int lane, double *buf, double *dinv, int &bad) {
    // pv: lane c holds the pivot (its L[c] after the updates of columns < c), computed WITHOUT the shared-memory round trip
    double d = __shfl_sync(0xffffffffu, pv, c);
    const bool neg = !(d > 0.0);                             // warp-uniform; keep going with finite numbers, the caller discards the result
    bad = (neg && bad == 0) ? c + 1 : bad;
    d = neg ? 1.0 : d;
    const double inv = rsqrt1(d);
    double *colb = buf + (c & 1) * CB;
    const double lrc = (lane > c) ? L[c] * inv : 0.0;
    L[c] = (lane == c) ? d * inv : lrc;
    colb[lane] = lrc;
    if (lane == 0) dinv[c] = inv;
    // the next pivot lives on lane c + 1 and only needs that lane's own lrc: keep it off the STS -> LDS path
    double pv_next = 0.0;
    if constexpr (c + 1 < CB) pv_next = fma(-lrc, lrc, L[c + 1]);
    __syncwarp();
#pragma unroll
    for (int k = c + 1; k < CB; ++k) L[k] = fma(-lrc, colb[k], L[k]);     // entries k > lane are never read
    if constexpr (c + 1 < CB) potf2_col<c + 1>(L, pv_next, lane, buf, dinv, bad);
}
template <int k>
__device__ __forceinline__ void trtri_col(double (&x)[CB], int lane, const double *Ls, const double *dinv) {
    // right-looking forward substitution for column `lane` of E = L^{-1}: x[k] holds delta_{k,lane} - sum_{j<k} L(k,j) E(j,lane) on
    // entry; finish it, then push its contribution into every later row (independent updates: the chain is one FMA per step)
    const double xk = (lane > k) ? 0.0 : x[k] * dinv[k];
    x[k] = xk;
#pragma unroll
    for (int i = k + 1; i < CB; ++i) x[i] = fma(-Ls[k + CS * i], xk, x[i]);          // Ls[k + CS*i] = L(i, k) = R(k, i)
    if constexpr (k + 1 < CB) trtri_col<k + 1>(x, lane, Ls, dinv);
}
__device__ __forceinline__ void potf2_inv_warp(const double *D, double *Rs, double *Ws, double *buf, int gcol0, int *fail) {
    const int lane = threadIdx.x & 31;
    double L[CB];
#pragma unroll
    for (int c = 0; c < CB; ++c) L[c] = (c <= lane) ? D[c + CS * lane] : 0.0;
    int bad = 0;
    double *dinv = buf + 2 * CB;
    potf2_col<0>(L, L[0], lane, buf, dinv, bad);
    if (bad && lane == 0) atomicCAS(fail, 0, gcol0 + bad);
#pragma unroll
    for (int c = 0; c < CB; ++c) Rs[c + CS * lane] = (c <= lane) ? L[c] : 0.0;       // R(c, lane) = L(lane, c)
    __syncwarp();
    double x[CB];
#pragma unroll
    for (int i = 0; i < CB; ++i) x[i] = (i == lane) ? 1.0 : 0.0;
    trtri_col<0>(x, lane, Rs, dinv);
#pragma unroll
    for (int i = 0; i < CB; ++i) Ws[lane + CS * i] = x[i];                           // R^{-1}(lane, i) = E(i, lane); zero for i < lane
}

__global__ void __launch_bounds__(CT) chol_inv_kernel(double *G, i64 ldg, int n, double *X, i64 ldx, int nb, int *flagC, int *flagX,
                                                      CholStatus *st) {
    __shared__ __align__(16) double As[CB * CS], Bs[CB * CS], Ds[CB * CS];
    __shared__ int sh_ok;
    __shared__ double red[2][CT / 32];
    __shared__ __align__(16) double pbuf[4 * CB];
    if (threadIdx.x == 0) sh_ok = 1;
    __syncthreads();
    const int ty = threadIdx.x >> 3, tx = threadIdx.x & 7;
    const int ntasks = nb * nb;
    for (int t = blockIdx.x; t < ntasks; t += gridDim.x) {
        int c = (int)sqrtf((float)t);
        while (c * c > t) --c;
        while ((c + 1) * (c + 1) <= t) ++c;
        const int idx = t - c * c;
        double acc[2][4];
        if (idx <= c) {
            // ---------------- C(r, c)
            const int r = idx;
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int gi = r * CB + ty + 16 * i, gj = c * CB + tx + 8 * j;
                    acc[i][j] = (gi < n && gj < n) ? G[(i64)gj * ldg + gi] : (gi == gj ? 1.0 : 0.0);
                }
            if (r == c && c > 0) {
                // Diagonal block: its last input, R(c-1, c), is one L2 hop behind W_{c-1} (task C(c-1, c) has to see the flag, load
                // W, multiply, store, publish).  The diagonal task keeps its own copy of that task's accumulation instead
                // (accP = G(c-1,c) - sum_{j<c-1} R(j,c-1)^T R(j,c): off the critical path, the inputs are old) and forms
                // R(c-1, c) = W_{c-1}^T accP itself as soon as W_{c-1} is published: one hop per block column instead of two.
                double accP[2][4];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int gi = (c - 1) * CB + ty + 16 * i, gj = c * CB + tx + 8 * j;
                        accP[i][j] = (gi < n && gj < n) ? G[(i64)gj * ldg + gi] : 0.0;
                    }
                for (int j = 0; j < c - 1; ++j) {
                    if (!wait2(flagC + j * nb + c, flagC + j * nb + (c - 1), &st->err, &sh_ok)) return;
                    load_block(As, G, ldg, j, c, n, false);          // R(j, c)
                    load_block(Bs, G, ldg, j, c - 1, n, false);      // R(j, c-1)
                    __syncthreads();
                    mma32<-1>(acc, As, As);
                    mma32<-1>(accP, Bs, As);
                    __syncthreads();
                }
                if (!wait2(flagC + (c - 1) * nb + (c - 1), nullptr, &st->err, &sh_ok)) return;
                load_block(As, X, ldx, c - 1, c - 1, n, false);      // W_{c-1}
                acc_to_smem(accP, Bs, false);
                __syncthreads();
                double rl[2][4] = {};
                mma32<1>(rl, As, Bs);                                 // R(c-1, c), the same arithmetic as task C(c-1, c)
                acc_to_smem(rl, Ds, false);
                __syncthreads();
                mma32<-1>(acc, Ds, Ds);
                __syncthreads();
            } else
            for (int j = 0; j < r; ++j) {
                if (!wait2(flagC + j * nb + r, flagC + j * nb + c, &st->err, &sh_ok)) return;
                load_block(As, G, ldg, j, r, n, false);
                if (r != c) load_block(Bs, G, ldg, j, c, n, false);
                __syncthreads();
                mma32<-1>(acc, As, (r != c) ? Bs : As);
                __syncthreads();
            }
            if (r == c) {
                acc_to_smem(acc, Ds, false);
                __syncthreads();
                if (threadIdx.x < 32) potf2_inv_warp(Ds, As, Bs, pbuf, c * CB, &st->fail);
                __syncthreads();
                store_block(As, G, ldg, c, c, n, false);
                store_block(Bs, X, ldx, c, c, n, false);
            } else {
                if (!wait2(flagC + r * nb + r, nullptr, &st->err, &sh_ok)) return;
                load_block(As, X, ldx, r, r, n, false);                    // W_r
                acc_to_smem(acc, Bs, false);
                __syncthreads();
                double out[2][4] = {};
                mma32<1>(out, As, Bs);                                      // R(r,c) = W_r^T acc
                acc_to_smem(out, Ds, false);
                __syncthreads();
                store_block(Ds, G, ldg, r, c, n, true);
            }
            __syncthreads();
            if (threadIdx.x == 0) { st_release(flagC + r * nb + c, 1); }
            if (r == c && c == nb - 1) {
                // the last diagonal block depends (transitively) on every other one: diag(R) is complete
                double mn = INFINITY, mx = 0.0;
                for (int i = threadIdx.x; i < n; i += CT) {
                    double d = fabs(__ldcg(G + (i64)i * ldg + i));
                    if (!(d == d)) d = 0.0;                                 // NaN -> breakdown
                    mn = fmin(mn, d); mx = fmax(mx, d);
                }
                for (int o = 16; o > 0; o >>= 1) {
                    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                }
                if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = mn; red[1][threadIdx.x >> 5] = mx; }
                __syncthreads();
                if (threadIdx.x == 0) {
                    for (int w = 1; w < CT / 32; ++w) { mn = fmin(mn, red[0][w]); mx = fmax(mx, red[1][w]); }
                    st->dmin = mn; st->dmax = mx;
                }
            }
        } else {
            // ---------------- X(r, c), r < c
            const int r = idx - c - 1;
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
            for (int i = r; i < c; ++i) {
                if (!wait2((i == r) ? flagC + r * nb + r : flagX + r * nb + i, flagC + i * nb + c, &st->err, &sh_ok)) return;
                load_block(As, X, ldx, r, i, n, true);                      // As[k + CS a] = X(r,i)(a, k)
                load_block(Bs, G, ldg, i, c, n, false);
                __syncthreads();
                mma32<1>(acc, As, Bs);
                __syncthreads();
            }
            if (!wait2(flagC + c * nb + c, nullptr, &st->err, &sh_ok)) return;
            acc_to_smem(acc, As, true);
            load_block(Bs, X, ldx, c, c, n, false);                         // W_c
            __syncthreads();
            double out[2][4] = {};
            mma32<-1>(out, As, Bs);
            acc_to_smem(out, Ds, false);
            __syncthreads();
            store_block(Ds, X, ldx, r, c, n, true);
            __syncthreads();
            if (threadIdx.x == 0) { st_release(flagX + r * nb + c, 1); }
        }
    }
}

}  // namespace

bool chol_inv_ok(i64 n) {
    if (n <= 0 || n > 8192) return false;
    static int coop = -1;
    if (coop < 0) {
        ensure_init();
        RSVD_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx().device));
    }
    return coop > 0;
}

// G (n x n, upper triangle read) <- R (upper, exact zeros below); Rinv <- R^{-1} (upper, exact zeros below).
// Returns 0 ok, > 0 = failing column + 1, < 0 = could not run (the caller falls back to potrf_upper/trtri_upper).
// dminmax (host, optional): min and max of diag(R).  Synchronises the stream once (24 bytes of status).
int chol_inv_upper(double *G, i64 ldg, i64 n, double *Rinv, i64 ldi, double *dminmax) {
    ensure_init();
    if (g_status) return -1;
    if (!chol_inv_ok(n)) return -1;
    Ctx &c = ctx();
    const int nb = (int)((n + CB - 1) / CB), ntasks = nb * nb;
    static int blocks_per_sm = -1;
    if (blocks_per_sm < 0) RSVD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, chol_inv_kernel, CT, 0));
    if (blocks_per_sm < 1) return -1;
    const int grid = std::min(ntasks, blocks_per_sm * c.sms);
    const size_t ws_bytes = sizeof(CholStatus) + (size_t)2 * ntasks * sizeof(int);
    char *ws = (char *)dalloc_bytes(ws_bytes);
    if (g_status) return -1;
    RSVD_CUDA(cudaMemsetAsync(ws, 0, ws_bytes, c.stream));
    CholStatus *st = (CholStatus *)ws;
    int *flagC = (int *)(ws + sizeof(CholStatus)), *flagX = flagC + ntasks;
    int ni = (int)n;
    void *args[] = {&G, &ldg, &ni, &Rinv, &ldi, (void *)&nb, &flagC, &flagX, &st};
    cudaError_t e = cudaLaunchCooperativeKernel((void *)chol_inv_kernel, dim3(grid), dim3(CT), args, 0, c.stream);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        dfree(ws);
        return -1;
    }
    count_launch();
    CholStatus *h = (CholStatus *)(c.h_flag + 32);          // pinned mirror (h_flag holds 64 ints)
    RSVD_CUDA(cudaMemcpyAsync(h, st, sizeof(CholStatus), cudaMemcpyDeviceToHost, c.stream));
    RSVD_CUDA(cudaStreamSynchronize(c.stream));
    dfree(ws);
    if (g_status) return -1;
    if (h->err) { set_error("rsvd_b200: Cholesky dataflow kernel timed out waiting for a block"); return -1; }
    if (dminmax) { dminmax[0] = h->dmin; dminmax[1] = h->dmax; }
    return h->fail;
}

}  // namespace rsvd
