// gemm_tma.cu — FP64 GEMM engine, part 2: the streaming kernel for the large products of the hot path
//   Y = A*Omega (fused Philox Omega), Y = A*Z          ("NN": A is M-major, B is K-major)
//   Z = A^T*Y, B^T = A^T*Q, Gram = Y^T*Y               ("TN": both operands K-major)
// (cblas_dgemm NN / TN call sites matrix_vector_functions_intel_mkl.c:542,551, as used by RRA:95,108,120,139.)
//
// Design (sm_100a):
//   * tcgen05.mma has no f64 kind, so the FP64 tensor path is warp-level DMMA (mma.sync m8n8k4.f64 -> DMMA.8x8x4).
//   * CTA tile 128 x (8*NB) x 16 with NB <= 16 column groups chosen per call so that the n-tiles cover n = k+p
//     without padding waste (l = 520 -> 5 tiles of 104 columns, not 5 x 128).  5-stage smem ring filled by TMA
//     (cp.async.bulk.tensor.2d, SWIZZLE_128B, 16-double-wide boxes => out-of-range rows/cols/k are zero-filled by
//     hardware), mbarrier full/empty pairs.
//   * warp-specialised: warpgroup 0 = producer (one lane issues TMA; in sketch mode all 128 threads generate the
//     Omega tile with Philox4x32-10 + Box-Muller straight into the swizzled smem stage — Omega never exists in HBM),
//     warpgroups 1-2 = 8 DMMA consumer warps, each 16 rows x all 8*NB columns (<= 64 accumulator doubles/thread).
//     setmaxnreg moves registers from the producer to the consumers.
//   * fragment loads are LDS.128 and bank-conflict free under the 128B swizzle by construction (ncu: 1e4 conflicts
//     in 1.9e9 wavefronts):
//       K-major operand tile [row][16 k]: thread (g,t) reads chunk (t+4s') of physical row pi(g) = (g>>1)+4(g&1),
//         i.e. k = 2t+8s'+{0,1} (two consecutive DMMA k-steps per load);
//       M-major operand tile [k][16 m]:   thread (g,t) reads chunk g of row k = 2t+c+8s', i.e. rows 2g, 2g+1
//         (two DMMA row-blocks per load).
//     Both use the same k permutation kappa(t) = 2t+c+8s', which is legal because the k-sum is order-free.
//   * work units = (tile, k-split).  Under-filled grids split every tile along k; otherwise only the tiles of the
//     last partial wave are split so that all 148 SMs finish together.  Split tiles write partial sums to a
//     workspace and a second kernel adds them in fixed order (deterministic, no atomics).
#include "gemm_tma_kernel.cuh"

namespace rsvd {

using namespace tma;

namespace tma {
extern template bool launch_tma<4>(bool, bool, const CUtensorMap &, const CUtensorMap &, const TmaP &, unsigned);
extern template bool launch_tma<8>(bool, bool, const CUtensorMap &, const CUtensorMap &, const TmaP &, unsigned);
extern template bool launch_tma<12>(bool, bool, const CUtensorMap &, const CUtensorMap &, const TmaP &, unsigned);
extern template bool launch_tma<13>(bool, bool, const CUtensorMap &, const CUtensorMap &, const TmaP &, unsigned);
extern template bool launch_tma<14>(bool, bool, const CUtensorMap &, const CUtensorMap &, const TmaP &, unsigned);
extern template bool launch_tma<15>(bool, bool, const CUtensorMap &, const CUtensorMap &, const TmaP &, unsigned);
extern template bool launch_tma<16>(bool, bool, const CUtensorMap &, const CUtensorMap &, const TmaP &, unsigned);
extern template int max_sketch_clusters<4>(bool);
extern template int max_sketch_clusters<8>(bool);
extern template int max_sketch_clusters<12>(bool);
extern template int max_sketch_clusters<13>(bool);
extern template int max_sketch_clusters<14>(bool);
extern template int max_sketch_clusters<15>(bool);
extern template int max_sketch_clusters<16>(bool);
}

namespace {

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

// co-resident sketch clusters for this tile width on the current device (queried once per device / width / layout)
int sketch_cluster_slots(int nb_tile, bool ta) {
    static int cache[16][17][2];       // [device][nb][ta], 0 = not asked yet, -1 = unavailable
    const int dev = ctx().device & 15;
    int &c = cache[dev][nb_tile][ta ? 1 : 0];
    if (c == 0) {
        int n = 0;
        switch (nb_tile) {
            case 4: n = max_sketch_clusters<4>(ta); break;
            case 8: n = max_sketch_clusters<8>(ta); break;
            case 12: n = max_sketch_clusters<12>(ta); break;
            case 13: n = max_sketch_clusters<13>(ta); break;
            case 14: n = max_sketch_clusters<14>(ta); break;
            case 15: n = max_sketch_clusters<15>(ta); break;
            case 16: n = max_sketch_clusters<16>(ta); break;
        }
        c = n > 0 ? n : -1;
    }
    return c > 0 ? c : 0;
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

    // n-tiling: G column groups of 8 spread evenly over T tiles of at most 16 groups
    const i64 G = (g.n + 7) / 8;
    i64 T = (G + 15) / 16;
    int nb_tile = (int)((G + T - 1) / T);
    // instantiated tile widths (column groups): round up to the next one
    static const int kWidths[7] = {4, 8, 12, 13, 14, 15, 16};
    for (int i = 0; i < 7; ++i) if (kWidths[i] >= nb_tile) { nb_tile = kWidths[i]; break; }
    T = (G + nb_tile - 1) / nb_tile;

    CUtensorMap mapA, mapB;
    bool ok = ta ? make_map(&mapA, g.A, g.k, g.m, g.lda, BM)    // A stored k x m : inner k, box {16 k, 128 rows}
                 : make_map(&mapA, g.A, g.m, g.k, g.lda, BK);   // A stored m x k : inner m, box {16 m, 16 k}
    if (!ok) return false;
    if (!g.philox) {
        if (!make_map(&mapB, g.B, g.k, g.n, g.ldb, 8 * nb_tile)) return false;
    } else {
        mapB = mapA;
    }

    TmaP p;
    p.m = g.m; p.n = g.n; p.k = g.k; p.C = g.C; p.ldc = g.ldc; p.alpha = g.alpha; p.beta = g.beta;
    p.seed = g.seed; p.ph_sk = g.ph_sk; p.ph_sc = g.ph_sc; p.ph_off = g.ph_off;
    p.c_vec2 = (((uintptr_t)g.C & 15) == 0 && (g.ldc & 1) == 0) ? 1 : 0;
    const i64 tm = (g.m + BM - 1) / BM;
    i64 tiles = tm * T;
    // sketch: clusters of two row tiles share every generated Omega stage (work units are then PAIRS of row tiles)
    int slots = ctx().sms;
    p.cl = 0;
    if (g.philox && tm >= 2 && !ctx().no_sketch_cluster) {
        const int cs = sketch_cluster_slots(nb_tile, ta);
        if (cs >= 8) { p.cl = 1; slots = cs; tiles = ((tm + 1) / 2) * T; }
    }
    p.b_upper = (g.b_upper && !g.philox) ? 1 : 0;
    p.sym = 0;
    if (g.sym_upper && g.m == g.n && !g.philox) {           // only the tiles meeting the upper triangle
        p.sym = 1;
        tiles = 0;
        for (i64 tn = 0; tn < T; ++tn) tiles += sym_rows((int)tm, 8 * nb_tile, (int)tn);
    }
    if (tiles > 0x3fffffffll) return false;
    p.tiles_n = (int)T; p.nb_tile = nb_tile;
    p.total_iters = (int)((g.k + BK - 1) / BK);
    const int sms = slots;              // schedulable units per wave: SMs, or co-resident clusters
    int max_split = max(1, min(16, p.total_iters / 32));   // keep >= 32 k-iterations per unit
    if (g.sumsq_out) max_split = 1;                        // the fused norm lives in the direct-store epilogue
    p.main_tiles = (int)tiles; p.s_main = 1; p.s_tail = 1;
    if (tiles < sms) {
        // under-filled: split every tile so that ~all SMs are busy
        int s = (int)(sms / tiles);
        p.s_main = max(1, min(s, max_split));
    } else {
        const int rem = (int)(tiles % sms);
        if (rem != 0 && rem * 2 <= sms && max_split >= 2) {
            // only the last partial wave is split: `rem` tiles -> rem * s_tail units ~ one full wave of short units
            p.main_tiles = (int)(tiles - rem);
            p.s_tail = max(2, min(sms / rem, min(8, max_split)));
        }
    }
    const i64 tail_tiles = tiles - p.main_tiles;
    const i64 units = (i64)p.main_tiles * p.s_main + tail_tiles * p.s_tail;
    const i64 split_units = (p.s_main > 1 ? (i64)p.main_tiles * p.s_main : 0) + (p.s_tail > 1 ? tail_tiles * p.s_tail : 0);
    DBuf part;
    p.part = nullptr;
    if (split_units > 0) {
        // partial tiles are indexed by unit id; when only the tail is split the main units do not touch the buffer,
        // so shift the base instead of allocating their slots
        const i64 first_split_unit = (p.s_main > 1) ? 0 : (i64)p.main_tiles * p.s_main;
        const i64 per_unit = p.cl ? 2 : 1;           // a cluster unit writes one partial tile per member CTA
        part.alloc((size_t)((units - first_split_unit) * per_unit) * PART_TILE);
        p.part = part.p - first_split_unit * per_unit * PART_TILE;
    }
    DBuf ss_part;
    p.ss_part = nullptr;
    if (g.sumsq_out) { ss_part.alloc((size_t)units * 8); p.ss_part = ss_part.p; }
    const unsigned grid = (unsigned)(p.cl ? 2 * units : units);
    if (g.philox && ctx().verbose >= 2)
        fprintf(stderr, "[rsvd_b200] sketch %lld x %lld x %lld: %s, %d slots per wave, %lld units (%d main tiles x %d, tail x %d), tile width %d\n", (long long)g.m, (long long)g.n,
                (long long)g.k, p.cl ? "2-CTA clusters" : "single CTAs", slots, (long long)units, p.main_tiles, p.s_main, p.s_tail, 8 * nb_tile);
    bool launched = false;
    switch (nb_tile) {
        case 4: launched = launch_tma<4>(ta, g.philox, mapA, mapB, p, grid); break;
        case 8: launched = launch_tma<8>(ta, g.philox, mapA, mapB, p, grid); break;
        case 12: launched = launch_tma<12>(ta, g.philox, mapA, mapB, p, grid); break;
        case 13: launched = launch_tma<13>(ta, g.philox, mapA, mapB, p, grid); break;
        case 14: launched = launch_tma<14>(ta, g.philox, mapA, mapB, p, grid); break;
        case 15: launched = launch_tma<15>(ta, g.philox, mapA, mapB, p, grid); break;
        case 16: launched = launch_tma<16>(ta, g.philox, mapA, mapB, p, grid); break;
        default: return false;
    }
    if (!launched) return false;
    const i64 split_tiles = (p.s_main > 1 ? p.main_tiles : 0) + (p.s_tail > 1 ? tail_tiles : 0);
    if (split_tiles > 0) {
        const unsigned rt = (unsigned)(p.cl ? 2 * split_tiles : split_tiles);
        tile_reduce_kernel<<<dim3(rt, rt >= 64 ? 2 : 8), 256, 0, ctx().stream>>>(p, 0);
        count_launch();
    }
    if (g.sumsq_out) sum_array_async(ss_part.p, units * 8, g.sumsq_out);
    return cudaGetLastError() == cudaSuccess;
}

}  // namespace rsvd
