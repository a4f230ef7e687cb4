// gemm_tma.cu — FP64 GEMM engine, part 2: the streaming kernel for the large products of the hot path
//   Y = A*Omega (fused Philox Omega), Y = A*Z          ("NN": A is M-major, B is K-major)
//   Z = A^T*Y, B^T = A^T*Q, Gram = Y^T*Y               ("TN": both operands K-major)
// (cblas_dgemm NN / TN call sites matrix_vector_functions_intel_mkl.c:542,551, as used by RRA:95,108,120,139.)
//
// Design (sm_100a):
//   * tcgen05.mma has no f64 kind, so the FP64 tensor path is warp-level DMMA (mma.sync m8n8k4.f64).
//   * CTA tile 128 x 128 x 16, 5-stage smem ring filled by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B,
//     16-double-wide boxes => out-of-range rows/cols/k are zero-filled by hardware), mbarrier full/empty pairs.
//   * warp-specialised: warpgroup 0 = producer (one lane issues TMA; in sketch mode all 128 threads generate the
//     Omega tile with Philox4x32-10 + Box-Muller straight into the swizzled smem stage, Omega never exists in
//     HBM), warpgroups 1-2 = 8 DMMA consumer warps, each a 64 x 32 sub-tile (64 accumulator doubles/thread).
//     setmaxnreg moves registers from the producer to the consumers.
//   * fragment loads are LDS.128 and bank-conflict free under the 128B swizzle by construction:
//       K-major operand tile [row][16 k]: thread (g,t) reads chunk (t+4s') of physical row pi(g) = (g>>1)+4(g&1),
//         i.e. k = 2t+8s'+{0,1} (two consecutive DMMA k-steps per load);
//       M-major operand tile [k][16 m]:   thread (g,t) reads chunk g of row k = 2t+c+8s', i.e. rows 2g, 2g+1
//         (two DMMA row-blocks per load).
//     Both use the same k permutation kappa(t) = 2t+c+8s', which is legal because the k-sum is order-free.
//   * split-K over gridDim.y with a deterministic second-pass reduction when the tile count under-fills 148 SMs.
#include "common.cuh"

namespace rsvd {

void splitk_reduce(const double *part, int splits, i64 m, i64 n, double alpha, double beta, double *C, i64 ldc,
                   int batch, i64 sC);

namespace {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int STAGES = 5;
constexpr int A_STAGE_BYTES = BM * BK * 8;   // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 8;   // 16 KB
constexpr int SMEM_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int NTHREADS = 384;

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
        else (void)cudaGetLastError();
    }
    return fn;
}

// 2-D FP64 tensor map over a column-major array: inner (contiguous) extent `inner`, outer extent `outer`,
// outer stride ld doubles; box = {16, box_outer}; 128B swizzle.
bool make_map(CUtensorMap *map, const double *base, i64 inner, i64 outer, i64 ld, int box_outer) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    cuuint32_t box[2] = {16u, (cuuint32_t)box_outer};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

struct TmaP {
    i64 m, n, k;          // C is m x n, contraction length k
    double *C; i64 ldc;
    double alpha, beta;
    int tiles_n;          // n-tiles (fastest-varying in blockIdx.x)
    int splits; int iters_per_split; double *part;
    uint64_t seed; i64 ph_sk, ph_sc, ph_off;
};

// ---- PTX helpers -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void lds128(uint32_t addr, double &x, double &y) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(addr));
}
__device__ __forceinline__ void sts128(uint32_t addr, double x, double y) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// A_KMAJOR: op(A) = A^T with A stored k x m (TN); otherwise A stored m x k (NN).  PHILOX: B generated on the fly.
template <bool A_KMAJOR, bool PHILOX>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, TmaP p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = smem_base;
    const uint32_t sB = smem_base + STAGES * A_STAGE_BYTES;
    const uint32_t bars = sB + STAGES * B_STAGE_BYTES;   // full[STAGES], empty[STAGES]
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int tile_n = blockIdx.x % p.tiles_n, tile_m = blockIdx.x / p.tiles_n;
    const i64 m0 = (i64)tile_m * BM, n0 = (i64)tile_n * BN;
    const int total_iters = (int)((p.k + BK - 1) / BK);
    const int it0 = blockIdx.y * p.iters_per_split;
    const int it1 = min(total_iters, it0 + p.iters_per_split);
    const int niter = max(0, it1 - it0);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), PHILOX ? 129 : 1);
            mbar_init(empty_bar(s), 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp < 4) {
        // ===================== producer warpgroup =====================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
        if (PHILOX || tid == 0) {
            for (int it = 0; it < niter; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (uint32_t)((it / STAGES) & 1);
                mbar_wait(empty_bar(s), ph ^ 1u);
                const int kc = (it0 + it) * BK;
                if (tid == 0) {
                    const uint32_t fb = full_bar(s);
                    mbar_expect_tx(fb, PHILOX ? A_STAGE_BYTES : (A_STAGE_BYTES + B_STAGE_BYTES));
                    if (A_KMAJOR) {
                        tma_load_2d(sA + s * A_STAGE_BYTES, &mapA, kc, (int)m0, fb);
                    } else {
#pragma unroll
                        for (int b = 0; b < 8; ++b)
                            tma_load_2d(sA + s * A_STAGE_BYTES + b * 2048, &mapA, (int)m0 + 16 * b, kc, fb);
                    }
                    if (!PHILOX) tma_load_2d(sB + s * B_STAGE_BYTES, &mapB, kc, (int)n0, fb);
                }
                if (PHILOX) {
                    // Omega tile: element (col j, kk) = normal(seed, off + (kc+kk)*sk + (n0+j)*sc),
                    // stored at j*128 + (((kk>>1) ^ (j&7))<<4) + (kk&1)*8 (the TMA SWIZZLE_128B image of [col][16 k]).
                    const uint32_t bbase = sB + s * B_STAGE_BYTES;
                    uint64_t cached_blk = ~0ull;
                    float z[4];
                    if (p.ph_sk == 1) {
                        const int j = tid;   // 0..127 : one column, 16 consecutive linear entries
                        const uint64_t lin0 = (uint64_t)(p.ph_off + (i64)kc + (n0 + j) * p.ph_sc);
#pragma unroll
                        for (int kp = 0; kp < 8; ++kp) {
                            double v[2];
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                uint64_t lin = lin0 + (uint64_t)(2 * kp + h);
                                uint64_t blk = lin >> 2;
                                if (blk != cached_blk) { rsvd_normal4(p.seed, blk, z); cached_blk = blk; }
                                uint32_t sel = (uint32_t)lin & 3u;
                                float f = sel == 0 ? z[0] : (sel == 1 ? z[1] : (sel == 2 ? z[2] : z[3]));
                                v[h] = (double)f;
                            }
                            sts128(bbase + j * 128 + ((kp ^ (j & 7)) << 4), v[0], v[1]);
                        }
                    } else {
                        // consecutive linear entries run along the columns (ph_sc == 1) or arbitrary strides:
                        // thread owns k-pair kp = tid&7 and columns (tid>>3)*8 .. +7
                        const int kp = tid & 7, jb = (tid >> 3) * 8;
                        double v[8][2];
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const i64 linb = p.ph_off + ((i64)kc + 2 * kp + h) * p.ph_sk + (n0 + jb) * p.ph_sc;
#pragma unroll
                            for (int jj = 0; jj < 8; ++jj) {
                                uint64_t lin = (uint64_t)(linb + jj * p.ph_sc);
                                uint64_t blk = lin >> 2;
                                if (blk != cached_blk) { rsvd_normal4(p.seed, blk, z); cached_blk = blk; }
                                uint32_t sel = (uint32_t)lin & 3u;
                                float f = sel == 0 ? z[0] : (sel == 1 ? z[1] : (sel == 2 ? z[2] : z[3]));
                                v[jj][h] = (double)f;
                            }
                        }
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj)
                            sts128(bbase + (jb + jj) * 128 + ((kp ^ ((jb + jj) & 7)) << 4), v[jj][0], v[jj][1]);
                    }
                    mbar_arrive(full_bar(s));
                }
            }
        }
    } else {
        // ===================== consumer warpgroups =====================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
        const int cw = warp - 4;
        const int wm = cw & 1, wn = cw >> 1;      // 2 x 4 warps, warp tile 64 (m) x 32 (n)
        const int g = lane >> 2, t = lane & 3;
        const int pg = (g >> 1) + 4 * (g & 1);    // physical row of logical row g in a K-major 8-row group

        double acc[8][4][2];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        // loop-invariant byte offsets inside a stage
        uint32_t boff[4][2];   // B: [col-group][s']
#pragma unroll
        for (int nb = 0; nb < 4; ++nb)
#pragma unroll
            for (int sp = 0; sp < 2; ++sp)
                boff[nb][sp] = (uint32_t)((wn * 32 + nb * 8 + pg) * 128 + (((t + 4 * sp) ^ pg) << 4));

        for (int it = 0; it < niter; ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (uint32_t)((it / STAGES) & 1);
            mbar_wait(full_bar(s), ph);
            const uint32_t a_base = sA + s * A_STAGE_BYTES;
            const uint32_t b_base = sB + s * B_STAGE_BYTES;
#pragma unroll
            for (int sp = 0; sp < 2; ++sp) {
                double b[4][2];
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) lds128(b_base + boff[nb][sp], b[nb][0], b[nb][1]);
                if (A_KMAJOR) {
                    double a[8][2];
#pragma unroll
                    for (int rb = 0; rb < 8; ++rb)
                        lds128(a_base + (uint32_t)((wm * 64 + rb * 8 + pg) * 128 + (((t + 4 * sp) ^ pg) << 4)),
                               a[rb][0], a[rb][1]);
#pragma unroll
                    for (int c = 0; c < 2; ++c)
#pragma unroll
                        for (int rb = 0; rb < 8; ++rb)
#pragma unroll
                            for (int nb = 0; nb < 4; ++nb) dmma(acc[rb][nb][0], acc[rb][nb][1], a[rb][c], b[nb][c]);
                } else {
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const int kk = 2 * t + c + 8 * sp;
                        double a[4][2];
#pragma unroll
                        for (int mc = 0; mc < 4; ++mc)
                            lds128(a_base + (uint32_t)((wm * 4 + mc) * 2048 + kk * 128 + ((g ^ (kk & 7)) << 4)),
                                   a[mc][0], a[mc][1]);
#pragma unroll
                        for (int mc = 0; mc < 4; ++mc)
#pragma unroll
                            for (int xy = 0; xy < 2; ++xy)
#pragma unroll
                                for (int nb = 0; nb < 4; ++nb)
                                    dmma(acc[mc * 2 + xy][nb][0], acc[mc * 2 + xy][nb][1], a[mc][xy], b[nb][c]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_bar(s));
        }

        // ---- epilogue: accumulators -> C (or split-K partial) ----
        const bool partial = p.splits > 1;
        double *Cb = partial ? p.part + (i64)blockIdx.y * p.m * p.n : p.C;
        const i64 ldc = partial ? p.m : p.ldc;
        const double alpha = partial ? 1.0 : p.alpha, beta = partial ? 0.0 : p.beta;
#pragma unroll
        for (int rb = 0; rb < 8; ++rb) {
            // physical row of accumulator row-block rb, logical row g
            i64 row;
            if (A_KMAJOR) row = m0 + wm * 64 + rb * 8 + pg;
            else          row = m0 + wm * 64 + (rb >> 1) * 16 + 2 * g + (rb & 1);
            if (row >= p.m) continue;
#pragma unroll
            for (int nb = 0; nb < 4; ++nb)
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    // logical column 2t+cc -> physical column pi(2t+cc) = t + 4*cc
                    i64 col = n0 + wn * 32 + nb * 8 + t + 4 * cc;
                    if (col >= p.n) continue;
                    double v = alpha * acc[rb][nb][cc];
                    double *c = Cb + col * ldc + row;
                    if (beta != 0.0) v += beta * (*c);
                    *c = v;
                }
        }
    }
}

template <bool AK, bool PH>
bool launch(const CUtensorMap &ma, const CUtensorMap &mb, const TmaP &p, dim3 grid) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tma_kernel<AK, PH>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) { (void)cudaGetLastError(); return false; }
        configured = true;
    }
    gemm_tma_kernel<AK, PH><<<grid, NTHREADS, SMEM_BYTES, ctx().stream>>>(ma, mb, p);
    count_launch();
    return cudaGetLastError() == cudaSuccess;
}

}  // namespace

bool gemm_tma_try(const Gemm &g) {
    const bool ta = (g.ta == 'T' || g.ta == 't'), tb = (g.tb == 'T' || g.tb == 't');
    if (g.batch != 1) return false;
    if (!g.philox && tb) return false;                       // B must be K-major (stored k x n)
    if (g.m < 64 || g.n < 16 || g.k < 64) return false;
    if ((double)g.m * (double)g.n * (double)g.k < 4.0e7) return false;   // small products: generic kernel
    if (((uintptr_t)g.A & 15) || (g.lda & 1)) return false;  // TMA: 16-byte aligned base and strides
    if (!g.philox && (((uintptr_t)g.B & 15) || (g.ldb & 1))) return false;
    if (g.m >= (1ll << 31) || g.n >= (1ll << 31) || g.k >= (1ll << 31)) return false;

    CUtensorMap mapA, mapB;
    memset(&mapB, 0, sizeof(mapB));
    bool ok = ta ? make_map(&mapA, g.A, g.k, g.m, g.lda, BM)    // A stored k x m : inner k, box {16 k, 128 rows}
                 : make_map(&mapA, g.A, g.m, g.k, g.lda, BK);   // A stored m x k : inner m, box {16 m, 16 k}
    if (!ok) return false;
    if (!g.philox) {
        if (!make_map(&mapB, g.B, g.k, g.n, g.ldb, BN)) return false;
    } else {
        mapB = mapA;
    }

    TmaP p;
    p.m = g.m; p.n = g.n; p.k = g.k; p.C = g.C; p.ldc = g.ldc; p.alpha = g.alpha; p.beta = g.beta;
    p.seed = g.seed; p.ph_sk = g.ph_sk; p.ph_sc = g.ph_sc; p.ph_off = g.ph_off;
    const i64 tm = (g.m + BM - 1) / BM, tn = (g.n + BN - 1) / BN;
    const i64 tiles = tm * tn;
    if (tiles > 0x7fffffffll) return false;
    p.tiles_n = (int)tn;
    const int total_iters = (int)((g.k + BK - 1) / BK);
    // split-K: pick the split count with the best wave efficiency on `sms` SMs (ties -> fewer splits)
    int best = 1; double best_eff = 0.0;
    const int sms = ctx().sms;
    for (int s = 1; s <= 16; ++s) {
        if (total_iters / s < 64 && s > 1) break;
        double work = (double)tiles * s;
        double waves = ceil(work / sms);
        double eff = work / (waves * sms);
        if (s > 1) eff *= 0.97;      // partial write + reduction pass is not free
        if (eff > best_eff + 0.02) { best_eff = eff; best = s; }
    }
    int ips = (total_iters + best - 1) / best;
    int splits = (total_iters + ips - 1) / ips;
    p.splits = splits; p.iters_per_split = ips; p.part = nullptr;
    DBuf part;
    if (splits > 1) { part.alloc((size_t)splits * g.m * g.n); p.part = part.p; }
    dim3 grid((unsigned)tiles, (unsigned)splits, 1);
    bool launched;
    if (ta) launched = g.philox ? launch<true, true>(mapA, mapB, p, grid) : launch<true, false>(mapA, mapB, p, grid);
    else    launched = g.philox ? launch<false, true>(mapA, mapB, p, grid) : launch<false, false>(mapA, mapB, p, grid);
    if (!launched) return false;
    if (splits > 1) splitk_reduce(part.p, splits, g.m, g.n, g.alpha, g.beta, g.C, g.ldc, 1, 0);
    return true;
}

}  // namespace rsvd
