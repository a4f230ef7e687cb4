// gemm.cu — FP64 GEMM engine, part 1: dispatch, the generic DMMA kernel (any layout / alignment / edge,
// strided batch, split-K with a deterministic reduction) and the FP64 peak micro-benchmark.
// Replaces cblas_dgemm NN/TN/NT (matrix_vector_functions_intel_mkl.c:538-561).
// The large streaming products (A*Omega, A*Z, A^T*Y) are served by the TMA-pipelined kernel in gemm_tma.cu.
#include "common.cuh"

namespace rsvd {

int g_last_gemm_path = 0;
bool gemm_tma_try(const Gemm &g);   // gemm_tma.cu

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

struct GenP {
    i64 m, n, k;
    const double *A; i64 lda; int ta;
    const double *B; i64 ldb; int tb;
    double *C; i64 ldc;
    double alpha, beta;
    int batch; i64 sA, sB, sC;
    int splits; i64 kchunk; double *part;   // split-K: partial tiles go to part[z][n][m]
    uint64_t seed; i64 ph_sk, ph_sc, ph_off;
};

constexpr int GT = 64;       // CTA tile (m and n)
constexpr int GK = 16;       // k per smem tile
constexpr int GLD = GT + 4;  // padded leading dimension: (GLD mod 16 == 4) makes the m8n8k4 fragment loads conflict-free

template <bool PHILOX>
__global__ void __launch_bounds__(128) gemm_generic_kernel(GenP p) {
    __shared__ double As[GK][GLD];
    __shared__ double Bs[GK][GLD];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp & 1) * 32, wn = (warp >> 1) * 32;
    const i64 m0 = (i64)blockIdx.x * GT, n0 = (i64)blockIdx.y * GT;
    const int zb = blockIdx.z / p.splits, zs = blockIdx.z % p.splits;
    const double *A = p.A + (i64)zb * p.sA;
    const double *B = p.B + (i64)zb * p.sB;
    const i64 kbeg = (i64)zs * p.kchunk;
    const i64 kend = min(p.k, kbeg + p.kchunk);

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    double ra[8], rb[8];
    auto load_tiles = [&](i64 k0) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            int e = tid + r * 128;
            {   // A tile element (i, kk)
                int i, kk;
                if (p.ta) { kk = e & 15; i = e >> 4; } else { i = e & 63; kk = e >> 6; }
                i64 gi = m0 + i, gk = k0 + kk;
                double v = 0.0;
                if (gi < p.m && gk < kend) v = p.ta ? A[gi * p.lda + gk] : A[gk * p.lda + gi];
                ra[r] = v;
            }
            {   // op(B) tile element (kk, j)
                int j, kk;
                if (p.tb) { j = e & 63; kk = e >> 6; } else { kk = e & 15; j = e >> 4; }
                i64 gj = n0 + j, gk = k0 + kk;
                double v = 0.0;
                if (gj < p.n && gk < kend) {
                    if (PHILOX) v = (double)rsvd_normal_at(p.seed, (uint64_t)(p.ph_off + gk * p.ph_sk + gj * p.ph_sc));
                    else v = p.tb ? B[gk * p.ldb + gj] : B[gj * p.ldb + gk];
                }
                rb[r] = v;
            }
        }
    };
    auto store_tiles = [&]() {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            int e = tid + r * 128;
            int i, kk, j, kb;
            if (p.ta) { kk = e & 15; i = e >> 4; } else { i = e & 63; kk = e >> 6; }
            if (p.tb) { j = e & 63; kb = e >> 6; } else { kb = e & 15; j = e >> 4; }
            As[kk][i] = ra[r];
            Bs[kb][j] = rb[r];
        }
    };

    if (kbeg < kend) load_tiles(kbeg);
    for (i64 k0 = kbeg; k0 < kend; k0 += GK) {
        store_tiles();
        __syncthreads();
        if (k0 + GK < kend) load_tiles(k0 + GK);
#pragma unroll
        for (int ks = 0; ks < GK / 4; ++ks) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[ks * 4 + t][wm + i * 8 + g];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[ks * 4 + t][wn + j * 8 + g];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncthreads();
    }

    if (p.splits > 1) {
        double *P = p.part + ((i64)blockIdx.z) * p.m * p.n;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    i64 gi = m0 + wm + i * 8 + g, gj = n0 + wn + j * 8 + 2 * t + c;
                    if (gi < p.m && gj < p.n) P[gj * p.m + gi] = acc[i][j][c];
                }
    } else {
        double *C = p.C + (i64)zb * p.sC;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    i64 gi = m0 + wm + i * 8 + g, gj = n0 + wn + j * 8 + 2 * t + c;
                    if (gi < p.m && gj < p.n) {
                        double v = p.alpha * acc[i][j][c];
                        if (p.beta != 0.0) v += p.beta * C[gj * p.ldc + gi];
                        C[gj * p.ldc + gi] = v;
                    }
                }
    }
}

// C = alpha * sum_z part[z] + beta * C   (fixed summation order => deterministic)
__global__ void splitk_reduce_kernel(const double *part, int splits, i64 m, i64 n, double alpha, double beta,
                                     double *C, i64 ldc, int batch, i64 sC) {
    i64 total = m * n;
    for (int b = 0; b < batch; ++b) {
        for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
            double s = 0.0;
            for (int z = 0; z < splits; ++z) s += part[((i64)(b * splits + z)) * total + e];
            i64 i = e % m, j = e / m;
            double *c = C + (i64)b * sC + j * ldc + i;
            double v = alpha * s;
            if (beta != 0.0) v += beta * (*c);
            *c = v;
        }
    }
}

void splitk_reduce(const double *part, int splits, i64 m, i64 n, double alpha, double beta, double *C, i64 ldc,
                   int batch, i64 sC) {
    i64 total = m * n;
    int blocks = (int)min((i64)ctx().sms * 8, (total + 255) / 256);
    splitk_reduce_kernel<<<blocks, 256, 0, ctx().stream>>>(part, splits, m, n, alpha, beta, C, ldc, batch, sC);
    count_launch();
}

static void gemm_generic(const Gemm &g) {
    GenP p;
    p.m = g.m; p.n = g.n; p.k = g.k;
    p.A = g.A; p.lda = g.lda; p.ta = (g.ta == 'T' || g.ta == 't');
    p.B = g.B; p.ldb = g.ldb; p.tb = (g.tb == 'T' || g.tb == 't');
    p.C = g.C; p.ldc = g.ldc; p.alpha = g.alpha; p.beta = g.beta;
    p.batch = g.batch; p.sA = g.sA; p.sB = g.sB; p.sC = g.sC;
    p.seed = g.seed; p.ph_sk = g.ph_sk; p.ph_sc = g.ph_sc; p.ph_off = g.ph_off;
    i64 tm = (g.m + GT - 1) / GT, tn = (g.n + GT - 1) / GT;
    i64 tiles = tm * tn * g.batch;
    int splits = 1;
    if (tiles < 2 * (i64)ctx().sms && g.k >= 1024) {
        i64 want = (2 * (i64)ctx().sms + tiles - 1) / tiles;
        i64 maxs = g.k / 512;
        splits = (int)max((i64)1, min(want, maxs));
        if (splits > 64) splits = 64;
    }
    i64 kchunk = ((g.k + splits - 1) / splits + GK - 1) / GK * GK;
    if (kchunk <= 0) kchunk = GK;
    splits = (int)((g.k + kchunk - 1) / kchunk);
    if (splits < 1) splits = 1;
    p.splits = splits; p.kchunk = kchunk; p.part = nullptr;
    DBuf part;
    if (splits > 1) { part.alloc((size_t)splits * g.batch * g.m * g.n); p.part = part.p; }
    if (g.m <= 0 || g.n <= 0) return;
    dim3 grid((unsigned)tm, (unsigned)tn, (unsigned)(g.batch * splits));
    if (g.philox) gemm_generic_kernel<true><<<grid, 128, 0, ctx().stream>>>(p);
    else gemm_generic_kernel<false><<<grid, 128, 0, ctx().stream>>>(p);
    count_launch();
    if (splits > 1) splitk_reduce(part.p, splits, g.m, g.n, g.alpha, g.beta, g.C, g.ldc, g.batch, g.sC);
    RSVD_CUDA(cudaGetLastError());
}

void gemm(const Gemm &g) {
    if (g_status) return;   // an earlier error (e.g. a failed allocation) is pending: launch nothing
    ensure_init();
    if (g.m <= 0 || g.n <= 0) return;
    if (!ctx().force_generic_gemm && gemm_tma_try(g)) { g_last_gemm_path = 1; return; }
    g_last_gemm_path = 0;
    gemm_generic(g);
    if (g.sumsq_out) sumsq_async(g.C, g.ldc, g.m, g.n, g.sumsq_out);   // the generic kernel has no fused norm: separate pass
}

// ---- FP64 peak micro-benchmark -----------------------------------------------------------------------
// Register-resident DMMA loop with the GEMM kernel's operand mix (2 x 8 accumulator blocks from 2 A- and 8 B-fragments
// per k-step, 8 warps per SM) and a DFMA loop for comparison.  MEASURED_PEAKS.json has no FP64 entry, so this is the
// roofline denominator bench.py reports.
__global__ void __launch_bounds__(256) dmma_peak_kernel(int iters, double *out) {
    double acc[2][8][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc[i][j][0] = threadIdx.x * 1e-9; acc[i][j][1] = (i + j) * 1e-9; }
    double a[2], b[8];
#pragma unroll
    for (int i = 0; i < 2; ++i) a[i] = 1.0 + (threadIdx.x + i) * 1e-12;
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] = 1.0 - (threadIdx.x + j) * 1e-12;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int i = 0; i < 2; ++i) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) s += acc[i][j][0] + acc[i][j][1];
    if (s == 123.456) out[0] = s;
}
__global__ void __launch_bounds__(256) dfma_peak_kernel(int iters, double *out) {
    double acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;
}

// mode 0: DMMA, best of 1/2/4 CTAs (8 warps each) per SM; mode 1: DFMA
double dmma_peak_tflops(int iters, int mode) {
    ensure_init();
    DBuf out(8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    const int occs[3] = {1, 2, 4};
    for (int oi = 0; oi < (mode ? 1 : 3); ++oi) {
        const int blocks = ctx().sms * (mode ? 4 : occs[oi]);
        float ms = 0;
        for (int rep = 0; rep < 3; ++rep) {   // rep 0 = warm-up
            cudaEventRecord(e0, ctx().stream);
            if (mode) dfma_peak_kernel<<<blocks, 256, 0, ctx().stream>>>(iters, out.p);
            else dmma_peak_kernel<<<blocks, 256, 0, ctx().stream>>>(iters, out.p);
            cudaEventRecord(e1, ctx().stream);
            cudaEventSynchronize(e1);
            float t = 0;
            cudaEventElapsedTime(&t, e0, e1);
            if (rep == 1 || (rep > 1 && t < ms)) ms = t;
        }
        count_launch(3);
        double flops = mode ? (double)blocks * 256 * (double)iters * 32 * 2 : (double)blocks * 8 * (double)iters * 16 * 512;
        best = fmax(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return best;
}

}  // namespace rsvd
