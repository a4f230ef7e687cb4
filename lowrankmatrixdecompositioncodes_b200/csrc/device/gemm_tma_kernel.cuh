// gemm_tma_kernel.cuh — kernel template of the streaming FP64 GEMM (see gemm_tma.cu for the design notes).
#pragma once
#include "common.cuh"

namespace rsvd {
namespace tma {


constexpr int BM = 128, BK = 16, BN_MAX = 128;
constexpr int STAGES = 5;
constexpr int A_STAGE_BYTES = BM * BK * 8;       // 16 KB
constexpr int B_STAGE_BYTES = BN_MAX * BK * 8;   // 16 KB
constexpr int SMEM_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int NTHREADS = 384;           // 1 producer + 2 consumer warpgroups
constexpr int NTHREADS_PHILOX = 512;    // sketch mode: 2 producer warpgroups (alternate stages) + 2 consumer warpgroups
constexpr int PART_TILE = BM * BN_MAX;           // doubles per partial tile

struct TmaP {
    i64 m, n, k;          // C is m x n, contraction length k
    double *C; i64 ldc;
    double alpha, beta;
    int tiles_n;          // n-tiles (fastest-varying in the tile index)
    int nb_tile;          // column groups (of 8) per n-tile
    int total_iters;      // ceil(k / BK)
    int main_tiles, s_main, s_tail;   // tiles [0,main_tiles) are split s_main ways, the rest s_tail ways
    double *part;         // partial tiles of split units
    int c_vec2;           // C rows can be stored as 16-byte pairs
    int sym;              // C is symmetric and only its upper triangle is wanted: tiles entirely below the diagonal are skipped
    int b_upper;          // op(B) (k x n) is upper triangular: column tile [n0, n0+w) only needs k < n0 + w
    uint64_t seed; i64 ph_sk, ph_sc, ph_off;
    int cl;               // sketch kernel launched as clusters of 2 CTAs (same column tile, adjacent row tiles) sharing the generated Omega
    double *ss_part;      // non-null: consumer warp w of unit u also writes the sum of squares of the C entries it stored to ss_part[8u + w]
};

// tile index -> (tile_m, tile_n).  Plain products: n-tiles vary fastest.  sym: only the tiles that meet the upper triangle,
// enumerated column tile by column tile (tile_n holds rows 0 .. upper_rows(tile_n)-1).
__host__ __device__ __forceinline__ int sym_rows(int tiles_m, int width, int tile_n) {   // row tiles meeting the upper triangle in column tile tile_n
    const int last_col = width * tile_n + width - 1;
    const int r = last_col / 128 + 1;          // BM = 128
    return r < tiles_m ? r : tiles_m;
}
__device__ __forceinline__ void tile_to_mn(const TmaP &p, int tile, int &tile_m, int &tile_n) {
    if (!p.sym) { tile_n = tile % p.tiles_n; tile_m = tile / p.tiles_n; return; }
    const int tiles_m = (int)((p.m + 127) / 128), width = 8 * p.nb_tile;
    int tn = 0;
    for (;; ++tn) {
        const int c = sym_rows(tiles_m, width, tn);
        if (tile < c || tn == p.tiles_n - 1) break;
        tile -= c;
    }
    tile_n = tn; tile_m = tile;
}

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
__device__ __forceinline__ void sts64(uint32_t addr, double x) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(x) : "memory");
}
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ---- cluster (DSMEM) helpers: the two CTAs of a sketch cluster write each other's Omega stage and signal each other's barriers ----
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa_peer(uint32_t addr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
// asynchronous DSMEM stores (STAS): the data lands in the peer CTA's shared memory and completes 16 / 8 bytes of the peer's
// transaction barrier — the same mechanism a TMA load uses, so the consumer needs neither a per-thread arrive nor a cluster fence
__device__ __forceinline__ void stas128(uint32_t addr, double x, double y, uint32_t bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" ::"r"(addr), "d"(x), "d"(y), "r"(bar) : "memory");
}
__device__ __forceinline__ void stas64(uint32_t addr, double x, uint32_t bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];" ::"r"(addr), "d"(x), "r"(bar) : "memory");
}
// "this warp is done READING the stage": nothing has to be published, and the reads have completed (their values fed the
// DMMAs already issued), so the remote arrive is relaxed — a release at cluster scope costs a MEMBAR.ALL.GPU per stage and warp
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {   // acquire at cluster scope
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAITC_%=:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONEC_%=;\n"
        "bra WAITC_%=;\n"
        "DONEC_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// byte offset of element (column j, k index kk) inside a K-major [col][16 k] stage under SWIZZLE_128B
__device__ __forceinline__ uint32_t bswz(int j, int kk) {
    return (uint32_t)(j * 128 + ((((kk >> 1) ^ (j & 7))) << 4) + (kk & 1) * 8);
}

// unit -> (tile, split, nsplit)
__device__ __forceinline__ void decode_unit(const TmaP &p, int unit, int &tile, int &split, int &nsplit) {
    const int main_units = p.main_tiles * p.s_main;
    if (unit < main_units) { tile = unit / p.s_main; split = unit - tile * p.s_main; nsplit = p.s_main; }
    else { int u = unit - main_units; int tt = u / p.s_tail; tile = p.main_tiles + tt; split = u - tt * p.s_tail; nsplit = p.s_tail; }
}

// A_KMAJOR: op(A) = A^T with A stored k x m (TN); otherwise A stored m x k (NN).  PHILOX: B generated on the fly.
// CL (sketch only): clusters of two CTAs working on the same column tile and adjacent row tiles.  Every Omega stage is
// generated ONCE per cluster — each CTA produces the half of the k-range given by its cluster rank, stores it into its own
// shared memory and ships it to the peer with asynchronous DSMEM stores that complete the peer's transaction barrier — so the
// Philox + Box-Muller work per CTA halves.  full = expect_tx arrive + one elected lane per generator warp; empty counts the
// consumer warps of both CTAs.
template <bool A_KMAJOR, bool PHILOX, int NB, bool CL = false>
__global__ void __launch_bounds__((PHILOX && !CL) ? NTHREADS_PHILOX : NTHREADS, 1)
gemm_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, TmaP p) {
    static_assert(!CL || PHILOX, "clusters are only used by the sketch kernel");
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = smem_base;
    const uint32_t sB = smem_base + STAGES * A_STAGE_BYTES;
    const uint32_t bars = sB + STAGES * B_STAGE_BYTES;   // full[STAGES], empty[STAGES]
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    int tile, split, nsplit;
    const uint32_t crank = CL ? cluster_ctarank() : 0u;
    decode_unit(p, CL ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, tile, split, nsplit);
    int tile_n, tile_m;
    if (CL) { tile_n = tile % p.tiles_n; tile_m = 2 * (tile / p.tiles_n) + (int)crank; }   // a row tile past the matrix computes zeros and stores nothing
    else tile_to_mn(p, tile, tile_m, tile_n);
    const i64 m0 = (i64)tile_m * BM, n0 = (i64)tile_n * (8 * NB);
    // NB = column groups per n-tile (compile time).  Columns beyond n in the last tile are zero-filled by TMA (or
    // generated and never stored), so every tile runs the same branch-free inner loop.
    const int titers = p.b_upper ? min(p.total_iters, (int)((n0 + 8 * NB + BK - 1) / BK)) : p.total_iters;   // rows of B below the diagonal are zero
    const int ips = (titers + nsplit - 1) / nsplit;
    const int it0 = split * ips;
    const int niter = max(0, min(titers, it0 + ips) - it0);
    constexpr uint32_t b_bytes = (uint32_t)(NB * 8 * BK * 8);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), PHILOX ? (CL ? 5 : 129) : 1);     // expect_tx arrive + the generator threads (CL: one lane per generator warp)
            mbar_init(empty_bar(s), CL ? 16 : 8);                    // the consumer warps (of both CTAs)
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (CL) cluster_sync_all();            // the peer's barriers exist before anything is signalled remotely
    const uint32_t peer = crank ^ 1u;

    constexpr int NPWG = (PHILOX && !CL) ? 2 : 1;     // producer warpgroups (a cluster CTA generates half a stage: one suffices, and the
                                                      // consumers keep the 208 registers of the plain kernel)
    if (warp < 4 * NPWG) {
        // ===================== producer warpgroup(s) =====================
        // sketch mode: generating one Omega tile costs ~1.3x the DMMA time of a stage for a single warpgroup (latency-
        // bound Philox + Box-Muller chains), so two warpgroups take alternate stages.
        if (PHILOX) asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
        else asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");     // (launch allocation: 168 with 384 threads, 128 with 512)
        const int pw = warp >> 2;                // which producer warpgroup
        const int ptid = tid & 127;              // thread index inside it
        if (PHILOX || tid == 0) {
            for (int it = pw; it < niter; it += NPWG) {
                const int s = it % STAGES;
                const uint32_t ph = (uint32_t)((it / STAGES) & 1);
                mbar_wait(empty_bar(s), ph ^ 1u);
                const int kc = (it0 + it) * BK;
                if (ptid == 0) {
                    const uint32_t fb = full_bar(s);
                    mbar_expect_tx(fb, PHILOX ? (CL ? A_STAGE_BYTES + b_bytes / 2 : A_STAGE_BYTES) : (A_STAGE_BYTES + b_bytes));   // CL: + the peer's half of Omega
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
                    // Omega tile: element (col j, kk) = normal(seed, off + (kc+kk)*sk + (n0+j)*sc), written at bswz(j, kk)
                    // (the image a TMA SWIZZLE_128B load of a stored Omega would have produced).
                    const uint32_t bbase = sB + s * B_STAGE_BYTES;
                    constexpr int ncols = 8 * NB;
                    if (CL) {
                        // this CTA's half of the stage: k in [8*crank, 8*crank + 8), written to both CTAs of the cluster
                        const uint32_t rbase = mapa_peer(bbase, peer), rbar = mapa_peer(full_bar(s), peer);
                        const int k0h = 8 * (int)crank;
                        if (p.ph_sk == 1) {
                            const int j = ptid;
                            if (j < ncols) {
                                const uint64_t lin0 = (uint64_t)(p.ph_off + (i64)kc + k0h + (n0 + j) * p.ph_sc);
                                if ((lin0 & 3u) == 0) {
                                    float z[2][4];
#pragma unroll
                                    for (int qd = 0; qd < 2; ++qd) rsvd_normal4(p.seed, (lin0 >> 2) + qd, z[qd]);
#pragma unroll
                                    for (int qd = 0; qd < 2; ++qd) {
                                        const int ch = 4 * (int)crank + 2 * qd;       // 16-byte chunk index of k = k0h + 4 qd
                                        const uint32_t o0 = (uint32_t)(j * 128 + ((ch ^ (j & 7)) << 4)), o1 = (uint32_t)(j * 128 + (((ch + 1) ^ (j & 7)) << 4));
                                        const double a0 = (double)z[qd][0], a1 = (double)z[qd][1], a2 = (double)z[qd][2], a3 = (double)z[qd][3];
                                        sts128(bbase + o0, a0, a1); sts128(bbase + o1, a2, a3);
                                        stas128(rbase + o0, a0, a1, rbar); stas128(rbase + o1, a2, a3, rbar);
                                    }
                                } else {
                                    uint64_t cached = ~0ull;
                                    float z[4];
                                    for (int kk = k0h; kk < k0h + 8; ++kk) {
                                        const uint64_t lin = lin0 + (uint64_t)(kk - k0h);
                                        if ((lin >> 2) != cached) { cached = lin >> 2; rsvd_normal4(p.seed, cached, z); }
                                        const uint32_t sel = (uint32_t)lin & 3u;
                                        const double f = (double)(sel == 0 ? z[0] : (sel == 1 ? z[1] : (sel == 2 ? z[2] : z[3])));
                                        sts64(bbase + bswz(j, kk), f); stas64(rbase + bswz(j, kk), f, rbar);
                                    }
                                }
                            }
                        } else if (p.ph_sc == 1 && ((p.ph_sk | (p.ph_off + n0)) & 3) == 0) {
                            constexpr int nitems = 8 * 2 * NB;     // 8 k  x  ncols/4 quads
#pragma unroll
                            for (int r = 0; r < 2; ++r) {
                                const int item = ptid + r * 128;
                                if (item < nitems) {
                                    const int kk = k0h + (item & 7), cq = item >> 3;
                                    const uint64_t lin = (uint64_t)(p.ph_off + ((i64)kc + kk) * p.ph_sk + n0 + 4 * cq);
                                    float z[4];
                                    rsvd_normal4(p.seed, lin >> 2, z);
#pragma unroll
                                    for (int i = 0; i < 4; ++i) { const uint32_t o = bswz(4 * cq + i, kk); sts64(bbase + o, (double)z[i]); stas64(rbase + o, (double)z[i], rbar); }
                                }
                            }
                        } else {
                            for (int e = ptid; e < ncols * 8; e += 128) {
                                const int kk = k0h + (e & 7), j = e >> 3;
                                const uint64_t lin = (uint64_t)(p.ph_off + ((i64)kc + kk) * p.ph_sk + (n0 + j) * p.ph_sc);
                                const double f = (double)rsvd_normal_at(p.seed, lin);
                                sts64(bbase + bswz(j, kk), f); stas64(rbase + bswz(j, kk), f, rbar);
                            }
                        }
                        __syncwarp();
                        if ((ptid & 31) == 0) mbar_arrive(full_bar(s));      // publishes this warp's local half (release, CTA scope)
                        continue;
                    }
                    if (p.ph_sk == 1) {
                        // linear index runs along k: thread = column, its 16 entries are 4 aligned Philox blocks
                        const int j = ptid;
                        if (j < ncols) {
                            const uint64_t lin0 = (uint64_t)(p.ph_off + (i64)kc + (n0 + j) * p.ph_sc);
                            if ((lin0 & 3u) == 0) {
                                float z[4][4];
#pragma unroll
                                for (int qd = 0; qd < 4; ++qd) rsvd_normal4(p.seed, (lin0 >> 2) + qd, z[qd]);   // 4 independent chains
#pragma unroll
                                for (int qd = 0; qd < 4; ++qd) {
                                    sts128(bbase + j * 128 + (((2 * qd) ^ (j & 7)) << 4), (double)z[qd][0], (double)z[qd][1]);
                                    sts128(bbase + j * 128 + (((2 * qd + 1) ^ (j & 7)) << 4), (double)z[qd][2], (double)z[qd][3]);
                                }
                            } else {
                                uint64_t cached = ~0ull;
                                float z[4];
                                for (int kk = 0; kk < 16; ++kk) {
                                    uint64_t lin = lin0 + (uint64_t)kk;
                                    if ((lin >> 2) != cached) { cached = lin >> 2; rsvd_normal4(p.seed, cached, z); }
                                    uint32_t sel = (uint32_t)lin & 3u;
                                    float f = sel == 0 ? z[0] : (sel == 1 ? z[1] : (sel == 2 ? z[2] : z[3]));
                                    sts64(bbase + bswz(j, kk), (double)f);
                                }
                            }
                        }
                    } else if (p.ph_sc == 1 && ((p.ph_sk | (p.ph_off + n0)) & 3) == 0) {
                        // linear index runs along the columns (left sketch of the ID): item = (k, column quad)
                        constexpr int nitems = 16 * 2 * NB;    // 16 k  x  ncols/4 quads
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            const int item = ptid + r * 128;
                            if (item < nitems) {
                                const int kk = item & 15, cq = item >> 4;
                                const uint64_t lin = (uint64_t)(p.ph_off + ((i64)kc + kk) * p.ph_sk + n0 + 4 * cq);
                                float z[4];
                                rsvd_normal4(p.seed, lin >> 2, z);
#pragma unroll
                                for (int i = 0; i < 4; ++i) sts64(bbase + bswz(4 * cq + i, kk), (double)z[i]);
                            }
                        }
                    } else {
                        // arbitrary strides: one Philox block per element
                        for (int e = ptid; e < ncols * 16; e += 128) {
                            const int kk = e & 15, j = e >> 4;
                            const uint64_t lin = (uint64_t)(p.ph_off + ((i64)kc + kk) * p.ph_sk + (n0 + j) * p.ph_sc);
                            sts64(bbase + bswz(j, kk), (double)rsvd_normal_at(p.seed, lin));
                        }
                    }
                    mbar_arrive(full_bar(s));
                }
            }
        }
    } else {
        // ===================== consumer warpgroups =====================
        if (PHILOX && !CL) asm volatile("setmaxnreg.inc.sync.aligned.u32 168;");
        else if (PHILOX) asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");     // 128 x 88 + 256 x 200 <= 64K registers
        else asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
        const int cw = warp - 4 * NPWG;           // rows [16*cw, 16*cw+16) of the tile, all columns
        const int g = lane >> 2, t = lane & 3;
        const int pg = (g >> 1) + 4 * (g & 1);    // physical row of logical row g in a K-major 8-row group

        double acc[2][NB][2];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < NB; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        // C <- C - A B on a full interior tile (the rank-kstep update of the blocked QB, RRA:1750: ~13 k-iterations per tile): the
        // accumulators start from -C, loaded while the TMA pipeline fills, so the old values cost no latency at all and the
        // epilogue only stores.  (-1) * (-C + sum) = C - sum exactly; only the summation order differs from alpha*acc + beta*C.
        const bool preload = !PHILOX && !A_KMAJOR && nsplit == 1 && p.beta == 1.0 && p.alpha == -1.0 && p.c_vec2 && m0 + BM <= p.m &&
                             n0 + 8 * NB <= p.n;
        if (preload) {
#pragma unroll
            for (int nb = 0; nb < NB; ++nb)
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const double2 o = __ldcs(reinterpret_cast<const double2 *>(p.C + (n0 + nb * 8 + t + 4 * cc) * p.ldc + m0 + cw * 16 + 2 * g));
                    acc[0][nb][cc] = -o.x; acc[1][nb][cc] = -o.y;
                }
        }

        // loop-invariant byte offsets inside a stage (s' = 0; s' = 1 adds 64 bytes before the XOR -> recomputed below)
        const uint32_t b_row = (uint32_t)(pg * 128);
        uint32_t b_ch[2], a_off[2][2];
#pragma unroll
        for (int sp = 0; sp < 2; ++sp) {
            b_ch[sp] = (uint32_t)(((t + 4 * sp) ^ pg) << 4);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                if (A_KMAJOR) a_off[sp][c] = (uint32_t)((cw * 16 + c * 8 + pg) * 128) + b_ch[sp];      // c = row-block here
                else { const int kk = 2 * t + c + 8 * sp; a_off[sp][c] = (uint32_t)(cw * 2048 + kk * 128 + ((g ^ (kk & 7)) << 4)); }
            }
        }

        for (int it = 0; it < niter; ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (uint32_t)((it / STAGES) & 1);
            mbar_wait(full_bar(s), ph);
            const uint32_t a_base = sA + s * A_STAGE_BYTES;
            const uint32_t b_base = sB + s * B_STAGE_BYTES + b_row;
#pragma unroll
            for (int sp = 0; sp < 2; ++sp) {
                // a[x][c]: x = row-block (0/1), c = k-step within the pair
                double a[2][2];
                if (A_KMAJOR) {
                    lds128(a_base + a_off[sp][0], a[0][0], a[0][1]);
                    lds128(a_base + a_off[sp][1], a[1][0], a[1][1]);
                } else {
                    lds128(a_base + a_off[sp][0], a[0][0], a[1][0]);   // k-step c=0: rows 2g (block 0), 2g+1 (block 1)
                    lds128(a_base + a_off[sp][1], a[0][1], a[1][1]);   // k-step c=1
                }
                // B fragments in chunks of 4 column groups, double-buffered so the LDS of chunk ch+1 are in flight
                // while the DMMAs of chunk ch issue
                constexpr int NCH = (NB + 3) / 4;
                double b[2][4][2];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (i < NB) lds128(b_base + (uint32_t)(i * 1024) + b_ch[sp], b[0][i][0], b[0][i][1]);
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) {
                    if (ch + 1 < NCH) {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if ((ch + 1) * 4 + i < NB)
                                lds128(b_base + (uint32_t)(((ch + 1) * 4 + i) * 1024) + b_ch[sp], b[(ch + 1) & 1][i][0], b[(ch + 1) & 1][i][1]);
                    }
#pragma unroll
                    for (int c = 0; c < 2; ++c)
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (ch * 4 + i < NB) {
                                dmma(acc[0][ch * 4 + i][0], acc[0][ch * 4 + i][1], a[0][c], b[ch & 1][i][c]);
                                dmma(acc[1][ch * 4 + i][0], acc[1][ch * 4 + i][1], a[1][c], b[ch & 1][i][c]);
                            }
                }
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(empty_bar(s));
                if (CL) mbar_arrive_remote_relaxed(mapa_peer(empty_bar(s), peer));
            }
        }

        // ---- epilogue: accumulators -> C, or the unit's partial tile ----
        // accumulator acc[x][nb][cc]: row = 16*cw + (M-major: 2g + x | K-major: 8x + pg), col = 8*nb + t + 4*cc
        if (nsplit > 1) {
            double *P = p.part + (i64)blockIdx.x * PART_TILE;
#pragma unroll
            for (int nb = 0; nb < NB; ++nb)
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) {
                        const int col = nb * 8 + t + 4 * cc;
                        if (A_KMAJOR) {
                            P[col * BM + cw * 16 + pg] = acc[0][nb][cc];
                            P[col * BM + cw * 16 + 8 + pg] = acc[1][nb][cc];
                        } else {
                            *reinterpret_cast<double2 *>(P + col * BM + cw * 16 + 2 * g) = make_double2(acc[0][nb][cc], acc[1][nb][cc]);
                        }
                    }
        } else {
            const double alpha = p.alpha, beta = preload ? 0.0 : p.beta;
            double ss = 0.0;          // fused Frobenius norm of the updated C (randQB_pb_new: ||A - Qp Bp||_F, RRA:1750-1751,1771)
            if (!A_KMAJOR && beta != 0.0 && p.c_vec2 && m0 + BM <= p.m && n0 + 8 * NB <= p.n) {
                // Read-modify-write of a full interior tile (the rank-kstep update A -= Qp Bp, RRA:1750, has only ~13 k-iterations
                // per tile, so its epilogue is a third of the work).  Old and new C alias, so the compiler keeps every load behind
                // the previous store: 32 dependent DRAM round trips per thread.  Here the loads of four column groups are issued
                // together (8 double2 in flight per thread: 4 round trips per tile instead of 32).
#pragma unroll
                for (int nb0 = 0; nb0 < NB; nb0 += 4) {
                    double2 o[4][2];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int cc = 0; cc < 2; ++cc)
                            if (nb0 + i < NB)
                                o[i][cc] = __ldcs(reinterpret_cast<const double2 *>(p.C + (n0 + (nb0 + i) * 8 + t + 4 * cc) * p.ldc + m0 + cw * 16 + 2 * g));
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int cc = 0; cc < 2; ++cc)
                            if (nb0 + i < NB) {
                                double2 v = make_double2(fma(beta, o[i][cc].x, alpha * acc[0][nb0 + i][cc]), fma(beta, o[i][cc].y, alpha * acc[1][nb0 + i][cc]));
                                *reinterpret_cast<double2 *>(p.C + (n0 + (nb0 + i) * 8 + t + 4 * cc) * p.ldc + m0 + cw * 16 + 2 * g) = v;
                                ss = fma(v.x, v.x, fma(v.y, v.y, ss));
                            }
                }
            } else
#pragma unroll
            for (int nb = 0; nb < NB; ++nb)
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) {
                        const i64 col = n0 + nb * 8 + t + 4 * cc;
                        if (col >= p.n) continue;
                        double *cp = p.C + col * p.ldc;
                        if (A_KMAJOR) {
#pragma unroll
                            for (int x = 0; x < 2; ++x) {
                                const i64 row = m0 + cw * 16 + 8 * x + pg;
                                if (row < p.m) {
                                    double v = alpha * acc[x][nb][cc];
                                    if (beta != 0.0) v += beta * cp[row];
                                    cp[row] = v;
                                    ss = fma(v, v, ss);
                                }
                            }
                        } else {
                            const i64 row = m0 + cw * 16 + 2 * g;
                            if (p.c_vec2 && row + 1 < p.m) {
                                double2 v = make_double2(alpha * acc[0][nb][cc], alpha * acc[1][nb][cc]);
                                double2 *dst = reinterpret_cast<double2 *>(cp + row);
                                if (beta != 0.0) { double2 o = *dst; v.x += beta * o.x; v.y += beta * o.y; }
                                *dst = v;
                                ss = fma(v.x, v.x, fma(v.y, v.y, ss));
                            } else {
#pragma unroll
                                for (int x = 0; x < 2; ++x)
                                    if (row + x < p.m) {
                                        double v = alpha * acc[x][nb][cc];
                                        if (beta != 0.0) v += beta * cp[row + x];
                                        cp[row + x] = v;
                                        ss = fma(v, v, ss);
                                    }
                            }
                        }
                    }
            if (p.ss_part) {          // fixed-order reduction: shuffle tree per warp, one slot per (unit, warp); summed by a second kernel
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
                if (lane == 0) p.ss_part[(i64)blockIdx.x * 8 + cw] = ss;
            }
        }
    }
    if (CL) cluster_sync_all();            // neither CTA leaves while the other may still signal its barriers
}

// sums the partial tiles of split units in fixed order: C = alpha * sum_s P[s] + beta * C.  One CTA per split tile.
static __global__ void __launch_bounds__(256) tile_reduce_kernel(TmaP p, int first_split_tile_is_main) {
    // split tiles: if s_main > 1 all main tiles (index 0..main_tiles-1) come first, then the tail tiles.
    // Cluster launches (p.cl): "tile" is a PAIR of row tiles; CTA 2u + r of unit u holds the partial of row tile 2*(tile / tiles_n) + r.
    const int member = p.cl ? (int)(blockIdx.x & 1) : 0;
    int st = p.cl ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, tile, nsplit, unit0;
    const int n_main_split = (p.s_main > 1) ? p.main_tiles : 0;
    if (st < n_main_split) { tile = st; nsplit = p.s_main; unit0 = tile * p.s_main; }
    else { int tt = st - n_main_split; tile = p.main_tiles + tt; nsplit = p.s_tail; unit0 = p.main_tiles * p.s_main + tt * p.s_tail; }
    (void)first_split_tile_is_main;
    int tile_n, tile_m;
    if (p.cl) { tile_n = tile % p.tiles_n; tile_m = 2 * (tile / p.tiles_n) + member; }
    else tile_to_mn(p, tile, tile_m, tile_n);
    const i64 m0 = (i64)tile_m * BM, n0 = (i64)tile_n * (8 * p.nb_tile);
    if (m0 >= p.m) return;
    const int ncols = (int)min((i64)(8 * p.nb_tile), p.n - n0);
    const i64 pstride = p.cl ? 2 * (i64)PART_TILE : (i64)PART_TILE;
    const double *P = p.part + (p.cl ? ((i64)unit0 * 2 + member) : (i64)unit0) * PART_TILE;
    // gridDim.y CTAs share one tile (a Gram matrix has 10-25 tiles: one CTA each would leave the reduction to a handful of SMs)
    for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < ncols * BM; e += blockDim.x * gridDim.y) {
        const int r = e % BM, c = e / BM;
        const i64 row = m0 + r, col = n0 + c;
        if (row >= p.m) continue;
        double s = 0.0;
        for (int u = 0; u < nsplit; ++u) s += P[(i64)u * pstride + c * BM + r];
        double *dst = p.C + col * p.ldc + row;
        double v = p.alpha * s;
        if (p.beta != 0.0) v += p.beta * (*dst);
        *dst = v;
    }
}


// one translation unit per NB instantiates this (gemm_tma_nb*.cu) so the variants compile in parallel
template <int NB>
bool launch_tma(bool a_kmajor, bool philox, const CUtensorMap &ma, const CUtensorMap &mb, const TmaP &p, unsigned grid) {
    auto go = [&](auto kern) -> bool {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) { (void)cudaGetLastError(); return false; }
        if (p.cl) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = ctx().stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            e = cudaLaunchKernelEx(&cfg, kern, ma, mb, p);
            count_launch();
            if (e != cudaSuccess) { (void)cudaGetLastError(); return false; }
            return true;
        }
        kern<<<grid, philox ? NTHREADS_PHILOX : NTHREADS, SMEM_BYTES, ctx().stream>>>(ma, mb, p);
        count_launch();
        return cudaGetLastError() == cudaSuccess;
    };
    if (p.cl) return a_kmajor ? go(gemm_tma_kernel<true, true, NB, true>) : go(gemm_tma_kernel<false, true, NB, true>);
    if (a_kmajor) return philox ? go(gemm_tma_kernel<true, true, NB>) : go(gemm_tma_kernel<true, false, NB>);
    return philox ? go(gemm_tma_kernel<false, true, NB>) : go(gemm_tma_kernel<false, false, NB>);
}

// co-resident 2-CTA clusters of the sketch kernel on this device (0: cluster launch not possible)
template <int NB>
int max_sketch_clusters(bool a_kmajor) {
    auto q = [&](auto kern) -> int {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * 148); cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = SMEM_BYTES;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
        return n;
    };
    return a_kmajor ? q(gemm_tma_kernel<true, true, NB, true>) : q(gemm_tma_kernel<false, true, NB, true>);
}

}  // namespace tma
}  // namespace rsvd
