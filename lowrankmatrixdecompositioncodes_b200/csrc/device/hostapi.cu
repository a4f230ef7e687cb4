// hostapi.cu — host-level entry points of the device layer: HOST pointers in, HOST pointers out, one call per
// reference API function.  This is what the C host code (csrc/host/rank_revealing_algorithms.c) calls for
// low_rank_svd_rand_decomp_fixed_rank (RRA:73-234), randQB_pb_new / low_rank_svd_blockrand_decomp_fixed_rank_or_prec
// (RRA:1576-1801, 239-381), id_rand_decomp_fixed_rank (RRA:1863-1965), id_two_sided_rand_decomp_fixed_rank (RRA:2060-2082)
// and cur_rand_decomp_fixed_rank (RRA:2191-2258).  Three things live here that the reference never needed:
//   * the upload of M is pipelined in column chunks with the first pass of EVERY algorithm (sketch / left sketch / first QB
//     block), not only the SVD's;
//   * single-process multi-GPU: with RSVD_B200_DEVICES / rsvd_b200_set_devices the host matrix is row-partitioned over one
//     worker per device (multi.cu); each worker uploads its row block straight from the caller's column-major matrix
//     (strided 2-D DMA), runs the row-partitioned pipeline with NCCL for the n x l / l x l sums, and downloads its row
//     block of U / C straight into the caller's output;
//   * residency across calls: a matrix registered with rsvd_b200_pin_matrix stays in HBM after its first upload, so a
//     driver that calls several routines on the same M (driver_multi_core_mkl3.c:51,82,92,112) uploads it once.
#include "pipeline.cuh"
#include <algorithm>
#include <map>
#include <mutex>
#include <vector>

namespace rsvd {

// ---- pipelined upload ---------------------------------------------------------------------------------------------
void upload_begin(const double *h, i64 ldh, double *dA, i64 rows, i64 n, Upload &up) {
    Ctx &c = ctx();
    up.ev.clear();
    if (rows <= 0 || n <= 0) return;
    i64 cw = ((i64)(256ll << 20) / (8 * rows)) / 16 * 16;     // ~256 MB column chunks, multiples of the GEMM's k-tile
    if (cw < 256) cw = 256;
    up.cw = cw;
    const int nchunks = (int)((n + cw - 1) / cw);
    up.ev.resize((size_t)nchunks);
    for (int i = 0; i < nchunks; ++i) {
        const i64 c0 = (i64)i * cw, w = std::min(cw, n - c0);
        RSVD_CUDA(cudaEventCreateWithFlags(&up.ev[i], cudaEventDisableTiming));
        if (ldh == rows)
            RSVD_CUDA(cudaMemcpyAsync(dA + c0 * rows, h + c0 * ldh, (size_t)w * rows * 8, cudaMemcpyHostToDevice, c.copy_stream));
        else   // a row block of a taller host matrix
            RSVD_CUDA(cudaMemcpy2DAsync(dA + c0 * rows, (size_t)rows * 8, h + c0 * ldh, (size_t)ldh * 8, (size_t)rows * 8, (size_t)w,
                                        cudaMemcpyHostToDevice, c.copy_stream));
        RSVD_CUDA(cudaEventRecord(up.ev[i], c.copy_stream));
    }
}
void upload_end(Upload &up) {
    if (up.ev.empty()) return;
    RSVD_CUDA(cudaStreamWaitEvent(ctx().stream, up.ev.back(), 0));   // whoever touches A next is ordered behind the whole upload
    for (auto &e : up.ev) cudaEventDestroy(e);
    up.ev.clear();
}

namespace {

// rows x cols block, device (ld ldd) -> host (ld ldh), on the compute stream; the caller synchronises
void d2h_block(double *h, i64 ldh, const double *d, i64 ldd, i64 rows, i64 cols) {
    if (rows <= 0 || cols <= 0) return;
    Ctx &c = ctx();
    if (ldh == rows && ldd == rows) RSVD_CUDA(cudaMemcpyAsync(h, d, (size_t)rows * cols * 8, cudaMemcpyDeviceToHost, c.stream));
    else RSVD_CUDA(cudaMemcpy2DAsync(h, (size_t)ldh * 8, d, (size_t)ldd * 8, (size_t)rows * 8, (size_t)cols, cudaMemcpyDeviceToHost, c.stream));
}

// ---- residency cache: host pointer -> device copies (one row block per worker) --------------------------------------
struct Resident {
    i64 m = 0, n = 0;
    int world = 0;
    bool valid = false;                 // device copies hold the matrix
    std::vector<double *> d;            // per worker (allocated and freed by that worker)
};
std::map<const double *, Resident> g_resident;
std::mutex g_res_mu;

Resident *resident_find(const double *h, i64 m, i64 n, int world) {
    std::lock_guard<std::mutex> lk(g_res_mu);
    auto it = g_resident.find(h);
    if (it == g_resident.end()) return nullptr;
    Resident &r = it->second;
    if (r.valid && (r.m != m || r.n != n || r.world != world)) return nullptr;   // registered with another shape: ignore
    return &r;
}

// how many workers a problem of m rows is split over: tiny problems stay on one device
int parts_for(i64 m) {
    const int W = pool_size();
    if (g_single_device || (W > 1 && m < 256ll * W)) return 1;
    return W;
}

// One worker's view of a host matrix: its row block, resident in HBM (from the cache, or uploaded now).
struct Block {
    i64 row0 = 0, rows = 0;
    double *dA = nullptr;
    bool owned = false;       // free at the end of the call
    Upload up;                // pending chunked upload (empty: resident)
};

// fan-out helper: runs fn(rank, world) on `parts` workers
template <class F>
void fan_out(int parts, F fn) {
    if (parts <= 1) {
        if (pool_size() > 1) {      // a small problem in multi-device mode: worker 0 alone, as a world of one
            pool_run([&](int rank) {
                if (rank != 0) return;
                Ctx &c = ctx();
                const int w = c.world, r = c.rank; void *comm = c.nccl_comm;
                c.world = 1; c.rank = 0; c.nccl_comm = nullptr;
                fn(0, 1);
                c.world = w; c.rank = r; c.nccl_comm = comm;
            });
        } else {
            ensure_init();
            fn(ctx().rank, 1);
        }
        return;
    }
    pool_run([&](int rank) { fn(rank, parts); });
}

// Sets the calling worker's partition state and makes its row block of the host matrix available.
// parts == 1 in a process-per-GPU job: the host matrix IS this rank's block; row0/m_global come from the options.
Block acquire_block(const double *hA, i64 m, i64 n, i64 ldh, int rank, int parts, bool will_modify) {
    Ctx &c = ctx();
    Block b;
    if (parts > 1) {
        rsvd_b200_row_partition(m, parts, rank, &b.row0, &b.rows);
        c.row0 = b.row0; c.m_global = m;
    } else {
        b.row0 = 0; b.rows = m;
        if (pool_size() > 1) { c.row0 = 0; c.m_global = m; }
    }
    Resident *r = resident_find(hA, m, n, parts);
    if (r && r->valid && !will_modify) { b.dA = r->d[(size_t)rank % r->d.size()]; return b; }
    if (r && r->valid && will_modify) {
        // the algorithm destroys its copy: duplicate the resident block if HBM allows, else take it over (and invalidate)
        size_t fr = 0, tot = 0;
        cudaMemGetInfo(&fr, &tot);
        const size_t need = (size_t)b.rows * n * 8;
        double *src = r->d[(size_t)rank % r->d.size()];
        if (fr > need + (8ull << 30)) {
            b.dA = dalloc((size_t)b.rows * n + 1); b.owned = true;
            if (b.dA) RSVD_CUDA(cudaMemcpyAsync(b.dA, src, need, cudaMemcpyDeviceToDevice, c.stream));
            return b;
        }
        b.dA = src; b.owned = true;
        std::lock_guard<std::mutex> lk(g_res_mu);
        r->d[(size_t)rank % r->d.size()] = nullptr;
        r->valid = false;
        return b;
    }
    b.dA = dalloc((size_t)b.rows * n + 1);
    if (!b.dA) return b;
    upload_begin(hA + b.row0, ldh, b.dA, b.rows, n, b.up);
    if (r && !will_modify) {      // registered for residency: keep this copy
        std::lock_guard<std::mutex> lk(g_res_mu);
        if ((int)r->d.size() != parts) r->d.assign((size_t)parts, nullptr);
        r->d[(size_t)(parts > 1 ? rank : 0)] = b.dA;
        r->m = m; r->n = n; r->world = parts;
    } else {
        b.owned = true;
    }
    return b;
}

void release_block(Block &b, const double *hA) {
    upload_end(b.up);
    if (b.owned) dfree(b.dA);
    b.dA = nullptr;
    (void)hA;
}

void mark_resident_valid(const double *hA, i64 m, i64 n, int parts) {
    std::lock_guard<std::mutex> lk(g_res_mu);
    auto it = g_resident.find(hA);
    if (it == g_resident.end()) return;
    Resident &r = it->second;
    if (r.m == m && r.n == n && r.world == parts && !r.d.empty()) {
        bool all = true;
        for (double *p : r.d) all = all && p != nullptr;
        if (all) r.valid = true;
    }
}

inline void stream_sync() { RSVD_CUDA(cudaStreamSynchronize(ctx().stream)); }
inline bool lead(int rank, int parts) { return parts == 1 || rank == 0; }   // who downloads the replicated outputs

}  // namespace

// ---- QB handle: the device-resident result of randQB_pb_new, consumed by the block-randomized SVD / ID / CUR tails -------
struct QB {
    i64 m = 0, n = 0, cap = 0, frank = 0, kstep = 0;
    int parts = 1;
    const double *hA = nullptr; i64 ldh = 0;
    std::vector<Block> blk;          // residual A - QB per worker
    std::vector<double *> dQ, dB;    // Q row block (rows x cap), B (cap x n, replicated) per worker
};

}  // namespace rsvd

using namespace rsvd;

extern "C" {

int rsvd_b200_pin_matrix(const double *h_A, rsvd_i64 m, rsvd_i64 n) {
    std::lock_guard<std::mutex> lk(g_res_mu);
    Resident &r = g_resident[h_A];
    if (r.valid && (r.m != m || r.n != n)) { set_error("rsvd_b200_pin_matrix: %p is already resident with another shape; unpin it first", (const void *)h_A); return 1; }
    r.m = m; r.n = n;
    return 0;
}

void rsvd_b200_unpin_matrix(const double *h_A) {
    Resident r;
    {
        std::lock_guard<std::mutex> lk(g_res_mu);
        auto it = g_resident.find(h_A);
        if (it == g_resident.end()) return;
        r = it->second;
        g_resident.erase(it);
    }
    if (r.d.empty()) return;
    const int parts = (int)r.d.size();
    fan_out(parts, [&](int rank, int) {
        double *p = r.d[(size_t)(parts > 1 ? rank : 0)];
        if (p) { dfree(p); stream_sync(); }
    });
}

int rsvd_b200_is_resident(const double *h_A) {
    std::lock_guard<std::mutex> lk(g_res_mu);
    auto it = g_resident.find(h_A);
    return it != g_resident.end() && it->second.valid;
}

/* low_rank_svd_rand_decomp_fixed_rank (RRA:73-234) from a host matrix to host factors. */
int rsvd_b200_svd_rand_h(const double *h_A, rsvd_i64 m, rsvd_i64 n, rsvd_i64 ldh, rsvd_i64 k, rsvd_i64 p, int vnum, int q, int s,
                         uint64_t seed, double *h_U, rsvd_i64 ldu, double *h_S, double *h_V, rsvd_i64 ldv) {
    const int parts = parts_for(m);
    fan_out(parts, [&](int rank, int) {
        if (!ctx().inited) return;
        Block b = acquire_block(h_A, m, n, ldh, rank, parts, false);
        DBuf U((size_t)b.rows * k + 1), S((size_t)k + 1), V((size_t)n * k + 1);
        if (b.dA && U.p && S.p && V.p) {
            svd_rand_impl(b.dA, b.rows, n, b.rows, k, p, vnum, q, s, seed, nullptr, nullptr, &b.up, U.p, b.rows, S.p, V.p, n);
            d2h_block(h_U + b.row0, ldu, U.p, b.rows, b.rows, k);
            if (lead(rank, parts)) { d2h_block(h_S, k, S.p, k, k, 1); d2h_block(h_V, ldv, V.p, n, n, k); }
        }
        release_block(b, h_A);
        stream_sync();
    });
    mark_resident_valid(h_A, m, n, parts);
    return g_status;
}

/* id_rand_decomp_fixed_rank (RRA:1863-1965): I (n doubles), T k x (n-k). */
int rsvd_b200_id_rand_h(const double *h_A, rsvd_i64 m, rsvd_i64 n, rsvd_i64 ldh, rsvd_i64 k, rsvd_i64 p, int q, int s, uint64_t seed,
                        double *h_I, double *h_T, rsvd_i64 ldt) {
    const int parts = parts_for(m);
    fan_out(parts, [&](int rank, int) {
        if (!ctx().inited) return;
        Block b = acquire_block(h_A, m, n, ldh, rank, parts, false);
        DBuf I((size_t)n + 1), T((size_t)k * std::max((i64)1, n - k));
        if (b.dA && I.p && T.p) {
            id_rand(b.dA, b.rows, n, b.rows, k, p, q, s, seed, nullptr, I.p, T.p, k, &b.up);
            if (lead(rank, parts)) { d2h_block(h_I, n, I.p, n, n, 1); d2h_block(h_T, ldt, T.p, k, k, n - k); }
        }
        release_block(b, h_A);
        stream_sync();
    });
    mark_resident_valid(h_A, m, n, parts);
    return g_status;
}

/* id_two_sided_rand_decomp_fixed_rank (RRA:2060-2082): Icol n, Irow m, T k x (n-k), S k x (m-k). */
int rsvd_b200_id_two_sided_rand_h(const double *h_A, rsvd_i64 m, rsvd_i64 n, rsvd_i64 ldh, rsvd_i64 k, rsvd_i64 p, int q, int s, uint64_t seed,
                                  double *h_Icol, double *h_Irow, double *h_T, rsvd_i64 ldt, double *h_S, rsvd_i64 lds) {
    const int parts = parts_for(m);
    fan_out(parts, [&](int rank, int) {
        Ctx &c = ctx();
        if (!c.inited) return;
        Block b = acquire_block(h_A, m, n, ldh, rank, parts, false);
        const i64 mg = (c.world > 1 && c.m_global > 0) ? c.m_global : m;
        DBuf Ic((size_t)n + 1), Ir((size_t)mg + 1), T((size_t)k * std::max((i64)1, n - k)), S((size_t)k * std::max((i64)1, mg - k));
        if (b.dA && Ic.p && Ir.p && T.p && S.p) {
            id_two_sided_rand(b.dA, b.rows, n, b.rows, k, p, q, s, seed, Ic.p, Ir.p, T.p, k, S.p, k, mg, &b.up);
            if (lead(rank, parts)) {
                d2h_block(h_Icol, n, Ic.p, n, n, 1); d2h_block(h_Irow, mg, Ir.p, mg, mg, 1);
                d2h_block(h_T, ldt, T.p, k, k, n - k); d2h_block(h_S, lds, S.p, k, k, mg - k);
            }
        }
        release_block(b, h_A);
        stream_sync();
    });
    mark_resident_valid(h_A, m, n, parts);
    return g_status;
}

/* cur_rand_decomp_fixed_rank (RRA:2191-2258): C m x k, U k x k, R k x n. */
int rsvd_b200_cur_rand_h(const double *h_A, rsvd_i64 m, rsvd_i64 n, rsvd_i64 ldh, rsvd_i64 k, rsvd_i64 p, int q, int s, uint64_t seed,
                         double *h_C, rsvd_i64 ldc, double *h_U, rsvd_i64 ldu, double *h_R, rsvd_i64 ldr) {
    const int parts = parts_for(m);
    fan_out(parts, [&](int rank, int) {
        Ctx &c = ctx();
        if (!c.inited) return;
        Block b = acquire_block(h_A, m, n, ldh, rank, parts, false);
        const i64 mg = (c.world > 1 && c.m_global > 0) ? c.m_global : m;
        DBuf Cm((size_t)b.rows * k + 1), U((size_t)k * k + 1), R((size_t)k * n + 1);
        if (b.dA && Cm.p && U.p && R.p) {
            cur_rand(b.dA, b.rows, n, b.rows, k, p, q, s, seed, Cm.p, b.rows, U.p, k, R.p, k, mg, &b.up);
            d2h_block(h_C + b.row0, ldc, Cm.p, b.rows, b.rows, k);
            if (lead(rank, parts)) { d2h_block(h_U, ldu, U.p, k, k, k); d2h_block(h_R, ldr, R.p, k, k, n); }
        }
        release_block(b, h_A);
        stream_sync();
    });
    mark_resident_valid(h_A, m, n, parts);
    return g_status;
}

/* randQB_pb_new (RRA:1576-1801) from a host matrix; Q, B and the residual stay in HBM behind the returned handle.
 * cap = kstep * (number of blocks that may be produced); *frank = columns produced.  NULL on failure. */
void *rsvd_b200_randqb_h(const double *h_A, rsvd_i64 m, rsvd_i64 n, rsvd_i64 ldh, rsvd_i64 kstep, rsvd_i64 nstep, rsvd_i64 cap, double tol,
                         int q, int s, uint64_t seed, rsvd_i64 *frank) {
    QB *qb = new QB();
    qb->m = m; qb->n = n; qb->cap = cap; qb->kstep = kstep; qb->hA = h_A; qb->ldh = ldh;
    const int parts = qb->parts = parts_for(m);
    qb->blk.resize((size_t)parts); qb->dQ.assign((size_t)parts, nullptr); qb->dB.assign((size_t)parts, nullptr);
    std::vector<i64> fr((size_t)parts, 0);
    fan_out(parts, [&](int rank, int) {
        if (!ctx().inited) return;
        const size_t slot = (size_t)(parts > 1 ? rank : 0);
        Block b = acquire_block(h_A, m, n, ldh, rank, parts, true);     // the private copy A = M (RRA:1630) lives only in HBM
        double *dQ = dalloc((size_t)b.rows * cap + 1), *dB = dalloc((size_t)cap * n + 1);
        if (b.dA && dQ && dB)
            randqb(b.dA, b.rows, n, b.rows, kstep, nstep, tol, q, s, seed, dQ, b.rows, dB, cap, cap, &fr[slot], 0, &b.up);
        upload_end(b.up);
        qb->blk[slot] = b; qb->dQ[slot] = dQ; qb->dB[slot] = dB;
        stream_sync();
    });
    qb->frank = fr[0];
    *frank = fr[0];
    return qb;
}

rsvd_i64 rsvd_b200_qb_parts(void *handle) { return handle ? ((QB *)handle)->parts : 0; }

/* single-device access for the tails that stay on one GPU (block-randomized ID / CUR): residual, Q, B device pointers */
int rsvd_b200_qb_dev_ptrs(void *handle, double **dA, double **dQ, double **dB) {
    QB *qb = (QB *)handle;
    if (!qb || qb->parts != 1) { set_error("rsvd_b200_qb_dev_ptrs: the QB result is partitioned over %d devices", qb ? qb->parts : 0); return 1; }
    *dA = qb->blk[0].dA; *dQ = qb->dQ[0]; *dB = qb->dB[0];
    return 0;
}

/* Q(:, 0:cols) -> h_Q (m x cols), B(0:cols, :) -> h_B (cols x n) */
int rsvd_b200_qb_download(void *handle, rsvd_i64 cols, double *h_Q, rsvd_i64 ldq, double *h_B, rsvd_i64 ldb) {
    QB *qb = (QB *)handle;
    if (!qb) return 1;
    const int parts = qb->parts;
    fan_out(parts, [&](int rank, int) {
        const size_t slot = (size_t)(parts > 1 ? rank : 0);
        Block &b = qb->blk[slot];
        if (!qb->dQ[slot] || !qb->dB[slot]) return;
        d2h_block(h_Q + b.row0, ldq, qb->dQ[slot], b.rows, b.rows, cols);
        if (lead(rank, parts)) d2h_block(h_B, ldb, qb->dB[slot], qb->cap, cols, qb->n);
        stream_sync();
    });
    return g_status;
}

/* SVD tail of low_rank_svd_blockrand_decomp_fixed_rank_or_prec (RRA:289-380) from the QB result, WITHOUT the original M:
 * M^T Q = Ares^T Q + B^T (Q^T Q) (pipeline.cu: svd_from_q_residual).  l = rows of B used, kk = rank kept. */
int rsvd_b200_qb_svd(void *handle, rsvd_i64 l, rsvd_i64 kk, int vnum, double *h_U, rsvd_i64 ldu, double *h_S, double *h_V, rsvd_i64 ldv) {
    QB *qb = (QB *)handle;
    if (!qb) return 1;
    const int parts = qb->parts;
    const i64 n = qb->n;
    fan_out(parts, [&](int rank, int) {
        const size_t slot = (size_t)(parts > 1 ? rank : 0);
        Block &b = qb->blk[slot];
        if (!b.dA || !qb->dQ[slot] || !qb->dB[slot]) return;
        DBuf U((size_t)b.rows * kk + 1), S((size_t)kk + 1), V((size_t)n * kk + 1);
        if (U.p && S.p && V.p) {
            svd_from_q_residual(b.dA, b.rows, n, b.rows, qb->dQ[slot], b.rows, l, qb->dB[slot], qb->cap, kk, vnum, U.p, b.rows, S.p, V.p, n);
            d2h_block(h_U + b.row0, ldu, U.p, b.rows, b.rows, kk);
            if (lead(rank, parts)) { d2h_block(h_S, kk, S.p, kk, kk, 1); d2h_block(h_V, ldv, V.p, n, n, kk); }
        }
        stream_sync();
    });
    return g_status;
}

/* hands the three device buffers over to the caller (after rsvd_b200_qb_dev_ptrs; free them with rsvd_b200_dev_free) */
void rsvd_b200_qb_release_handle(void *handle) { delete (QB *)handle; }

void rsvd_b200_qb_free(void *handle) {
    QB *qb = (QB *)handle;
    if (!qb) return;
    const int parts = qb->parts;
    fan_out(parts, [&](int rank, int) {
        const size_t slot = (size_t)(parts > 1 ? rank : 0);
        Block &b = qb->blk[slot];
        if (b.dA && b.owned) dfree(b.dA);
        dfree(qb->dQ[slot]); dfree(qb->dB[slot]);
        stream_sync();
    });
    delete qb;
}

}  // extern "C"
