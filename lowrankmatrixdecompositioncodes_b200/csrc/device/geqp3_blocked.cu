// geqp3_blocked.cu — blocked column-pivoted Householder QR with LAPACK-dgeqp3 pivoting for short-and-wide matrices
// (rows <= 4096: the l x n sketch of the column ID and the k x m transposed column subset of the row ID), optionally
// COLUMN-SHARDED over the ranks of a row-partitioned job.  Replaces pivotedQR_mkl -> LAPACKE_dgeqp3
// (rank_revealing_algorithms_intel_mkl.c:924-976, as called from RRA:1938 and RRA:1836) and the gather that precedes the
// row ID (RRA:2071-2078).
//
// dlaqps organisation (LAPACK 3.9): within a block of NB = 32 reflectors the trailing matrix is NOT updated; per step only
//   * the pivot column is brought up to date            A(k:m,p) -= V(k:m,0:j) F(p,0:j)^T
//   * one GEMV over the trailing matrix gives            F(:,j) = tau A(k:m,:)^T v  - tau F(:,0:j) (V^T v)
//   * the pivot ROW is updated                           A(k,:) -= V(k,0:j+1) F(:,0:j+1)^T
//   * the partial column norms are downdated from |A(k,c)| (dlaqps / dlaqp2 formulas, sqrt(eps) safeguard)
// and the block reflector is applied once per block with the DMMA GEMM.  Compared with the unblocked kernel (geqp3.cu) each
// step READS the trailing matrix once instead of reading and writing it — the HBM traffic that bounds this kernel halves.
// Two departures from dlaqps, both exact in real arithmetic:
//   * columns are never swapped in memory: a column keeps its storage slot and carries its current POSITION, the pivot is
//     the largest partial norm with the smallest position (idamax's first-index rule); this is what makes column sharding
//     free of data movement — a rank only ever ships its best candidate column;
//   * a column whose norm trips the sqrt(eps) safeguard does not end the block: its up-to-date tail is rebuilt on the spot
//     from the pending reflectors (m x j work for that one column) and its norm recomputed — the value dlaqps computes after
//     its early block exit.  (Sketch matrices trip the safeguard ~n times over the factorisation; ending the block each time
//     would degrade every step to BLAS-2.)
// Sharded mode: every rank owns a contiguous range of columns; per step one ncclAllGather of (norm, position, column id,
// up-to-date candidate column) replaces dlaqps' idamax + column swap; the reflector is then generated redundantly on every
// rank from the winning candidate, and everything else is local.
#include "common.cuh"
#include "ddsum.cuh"

namespace rsvd {

namespace {

constexpr int QNB = 32;                  // block size = warp width: lane i of a warp holds F(c, i)
constexpr int CAND_HDR = 4;              // candidate record: {norm, position, global column, spare} + the column (m doubles)
constexpr double TOL3Z = 1.0536712127723509e-08;   // sqrt(dlamch('Epsilon')) = sqrt(2^-53)

struct Best { double v; int pos; int c; };
__device__ __forceinline__ bool better(double v, int pos, double bv, int bpos) { return v > bv || (v == bv && pos < bpos); }
__device__ __forceinline__ Best warp_best(Best b) {
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, b.v, o);
        const int op = __shfl_xor_sync(0xffffffffu, b.pos, o), oc = __shfl_xor_sync(0xffffffffu, b.c, o);
        if (better(ov, op, b.v, b.pos)) { b.v = ov; b.pos = op; b.c = oc; }
    }
    return b;
}
// CTA-wide best -> part[blockIdx.x]; sh_* hold one entry per warp
__device__ void block_best_store(Best b, double *sh_v, int *sh_p, int *sh_c, double *part_v, int *part_pos, int *part_c) {
    b = warp_best(b);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (l == 0) { sh_v[w] = b.v; sh_p[w] = b.pos; sh_c[w] = b.c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < nw; ++i)
            if (better(sh_v[i], sh_p[i], b.v, b.pos)) { b.v = sh_v[i]; b.pos = sh_p[i]; b.c = sh_c[i]; }
        part_v[blockIdx.x] = b.v; part_pos[blockIdx.x] = b.pos; part_c[blockIdx.x] = b.c;
    }
}

// initial partial norms, positions, and the first candidate partials.  One warp per column.
__global__ void __launch_bounds__(256) qp_init_kernel(const double *__restrict__ A, i64 lda, int m, int nloc, int col0, double *vn1, double *vn2,
                                                      int *pos, double *part_v, int *part_pos, int *part_c) {
    __shared__ double sh_v[8];
    __shared__ int sh_p[8], sh_c[8];
    const int lane = threadIdx.x & 31;
    const int w0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    Best b; b.v = -2.0; b.pos = 0x7fffffff; b.c = -1;
    for (int c = w0; c < nloc; c += nw) {
        dd a; a.hi = 0.0; a.lo = 0.0;
        const double *col = A + (i64)c * lda;
        for (int r = lane; r < m; r += 32) a = dd_add_sq(a, col[r]);
        a = dd_warp_sum(a);
        const double v = dd_sqrt(a);
        if (lane == 0) { vn1[c] = v; vn2[c] = v; pos[c] = col0 + c; }
        if (better(v, col0 + c, b.v, b.pos)) { b.v = v; b.pos = col0 + c; b.c = c; }
    }
    block_best_store(b, sh_v, sh_p, sh_c, part_v, part_pos, part_c);
}

struct StepP {
    double *A; i64 lda;
    int m, nloc, col0, n_global, kmax;
    int k, j;                 // step, index inside the current block
    int world;
    double *vn1, *vn2; int *pos, *perm;
    double *Ft;               // QNB x nloc (ld QNB): F(c, i) at Ft[c*QNB + i]
    double *Vb;               // reflectors of the current block: Vb[i*m + r], i < QNB (a window of Vall)
    double *vglob;            // the current reflector v (m doubles, zeros above row k)
    double *Rpiv; int ldr;    // kmax x kmax: R columns of the pivots, in position order
    double *tau, *aux;        // tau[kmax]; aux[QNB] = -tau V^T v
    double *part_v; int *part_pos, *part_c; int nparts;
    double *cand_send, *cand_all;     // (CAND_HDR + m) doubles per rank
    i64 cand_stride;                  // doubles between two ranks' records in cand_all (0: CAND_HDR + m, the all-gather layout)
    int ps_formula, downdate;
};

// ---- phase 1 (per rank): local best candidate and its up-to-date column -> cand_send --------------------------------------
// ushared (one rank only): the candidate column also stays in shared memory, so the reflector phase does not read it back through L2
__device__ void candidate_phase(const StepP &p, double *ushared) {
    __shared__ double sh_v[32];
    __shared__ int sh_p[32], sh_c[32];
    __shared__ Best win;
    const int tid = threadIdx.x, nt = blockDim.x;
    Best b; b.v = -2.0; b.pos = 0x7fffffff; b.c = -1;
    for (int i = tid; i < p.nparts; i += nt)
        if (better(p.part_v[i], p.part_pos[i], b.v, b.pos)) { b.v = p.part_v[i]; b.pos = p.part_pos[i]; b.c = p.part_c[i]; }
    b = warp_best(b);
    if ((tid & 31) == 0) { sh_v[tid >> 5] = b.v; sh_p[tid >> 5] = b.pos; sh_c[tid >> 5] = b.c; }
    __syncthreads();
    if (tid == 0) {
        for (int i = 1; i < (nt >> 5); ++i)
            if (better(sh_v[i], sh_p[i], b.v, b.pos)) { b.v = sh_v[i]; b.pos = sh_p[i]; b.c = sh_c[i]; }
        win = b;
    }
    __syncthreads();
    const int c = win.c;
    double *cand = p.cand_send;
    if (tid == 0) {
        cand[0] = (c >= 0) ? win.v : -2.0;
        cand[1] = (double)win.pos;
        cand[2] = (c >= 0) ? (double)(p.col0 + c) : -1.0;
        cand[3] = 0.0;
    }
    const double *col = (c >= 0) ? p.A + (i64)c * p.lda : nullptr;
    __shared__ double fs[QNB];
    if (tid < QNB) fs[tid] = (c >= 0 && tid < p.j) ? p.Ft[(i64)c * QNB + tid] : 0.0;
    __syncthreads();
    for (int r = tid; r < p.m; r += nt) {
        double u = 0.0;
        if (c >= 0) {
            u = col[r];
            if (r >= p.k) {
                const double *vb = p.Vb + r;
#pragma unroll 8
                for (int i = 0; i < p.j; ++i) u = fma(-vb[(i64)i * p.m], fs[i], u);   // pending reflectors of this block
            }
        }
        if (ushared) ushared[r] = u; else cand[CAND_HDR + r] = u;
    }
}

// ---- phase 2 (identical on every rank): winner, bookkeeping, reflector, aux ------------------------------------------------
__device__ void reflect_phase(const StepP &p, const double *cand_all, const double *ushared) {
    __shared__ double shn[64];
    __shared__ int s_w;
    __shared__ double s_tau, s_scal, s_beta;
    const int tid = threadIdx.x, nt = blockDim.x;
    const i64 rec = p.cand_stride ? p.cand_stride : (i64)(CAND_HDR + p.m);
    if (tid == 0) {
        int w = 0; double bv = cand_all[0]; int bp = (int)cand_all[1];
        for (int g = 1; g < p.world; ++g) {
            const double v = cand_all[(i64)g * rec]; const int ps = (int)cand_all[(i64)g * rec + 1];
            if (better(v, ps, bv, bp)) { bv = v; bp = ps; w = g; }
        }
        s_w = w;
    }
    __syncthreads();
    if (tid == 32) {
        // bookkeeping on a warp of its own: nothing else in this kernel depends on it (the next kernel does)
        const int w = s_w, k = p.k;
        const int b = (int)cand_all[(i64)w * rec + 2], pb = (int)cand_all[(i64)w * rec + 1];
        const int a = p.perm[k];                       // the column sitting at position k moves to the pivot's old position
        p.perm[k] = b; p.perm[pb] = a;
        if (a != b && a >= p.col0 && a < p.col0 + p.nloc) p.pos[a - p.col0] = pb;
        if (b >= p.col0 && b < p.col0 + p.nloc) { p.pos[b - p.col0] = k; p.vn1[b - p.col0] = -1.0; }   // done: never a candidate again
    }
    const double *u = ushared ? ushared : cand_all + (i64)s_w * rec + CAND_HDR;
    const int k = p.k, m = p.m;
    // dlarfg on u(k:m)
    dd acc; acc.hi = 0.0; acc.lo = 0.0;
    for (int r = k + 1 + tid; r < m; r += nt) acc = dd_add_sq(acc, u[r]);
    acc = dd_warp_sum(acc);
    if ((tid & 31) == 0) { shn[2 * (tid >> 5)] = acc.hi; shn[2 * (tid >> 5) + 1] = acc.lo; }
    __syncthreads();
    if (tid == 0) {
        dd t; t.hi = 0.0; t.lo = 0.0;
        for (int i = 0; i < (nt >> 5); ++i) { dd b2; b2.hi = shn[2 * i]; b2.lo = shn[2 * i + 1]; t = dd_add(t, b2); }
        const double xnorm = dd_sqrt(t), alpha = u[k];
        double tau = 0.0, scal = 0.0, beta = alpha;
        if (xnorm != 0.0) {
            const double aa = fabs(alpha), xx = fabs(xnorm);
            const double w = fmax(aa, xx), z = fmin(aa, xx);
            const double h = (z == 0.0) ? w : w * sqrt(1.0 + (z / w) * (z / w));     // dlapy2
            beta = (alpha >= 0.0) ? -h : h;
            tau = (beta - alpha) / beta;
            scal = 1.0 / (alpha - beta);
        }
        s_tau = tau; s_scal = scal; s_beta = beta;
        p.tau[k] = tau;
    }
    __syncthreads();
    const double scal = s_scal, beta = s_beta, tau = s_tau;
    double *vb = p.Vb + (i64)p.j * m;
    for (int r = tid; r < m; r += nt) {
        const double ur = u[r];
        const double vv = (r < k) ? 0.0 : (r == k ? 1.0 : ur * scal);
        vb[r] = vv; p.vglob[r] = vv;
        if (r < p.kmax) p.Rpiv[(i64)k * p.ldr + r] = (r < k) ? ur : (r == k ? beta : 0.0);
    }
    __syncthreads();
    // aux[i] = -tau * V(k:m, i)^T v, i < j: one warp per previous reflector of the block
    const int w = tid >> 5, lane = tid & 31;
    for (int i = w; i < QNB; i += (nt >> 5)) {
        double s = 0.0;
        if (i < p.j) {
            const double *vi = p.Vb + (i64)i * m;
            for (int r = k + lane; r < m; r += 32) s = fma(vi[r], vb[r], s);
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        }
        if (lane == 0) p.aux[i] = (i < p.j) ? -tau * s : 0.0;
    }
}

__global__ void __launch_bounds__(1024) qp_candidate_kernel(StepP p) { candidate_phase(p, nullptr); }
__global__ void __launch_bounds__(1024) qp_reflect_kernel(StepP p) { reflect_phase(p, p.cand_all, nullptr); }
// Several ranks, exchange through the peer mailboxes (dist.cu: peer_mailbox): the candidate record goes straight into every
// peer's HBM over NVLink, a system-scope release store publishes the step's token, and the CTA waits for the tokens of all
// peers before the (replicated) reflector phase — one launch per step instead of candidate kernel + ncclAllGather + reflect
// kernel.  Two record buffers alternate with the parity of the token; a rank can be at most one step ahead of a peer (it
// needs the peer's record of the current step to finish it), so the buffer being read is never the one being written.
struct PeerArgs {
    double *const *peer_cand; unsigned long long *const *peer_flag;   // [world], the same layout on every rank
    double *cand; unsigned long long *flag;                            // this rank's mailbox
    unsigned long long token; i64 rec_max; int rank; int *err;
};
__global__ void __launch_bounds__(1024) qp_pivot_peer_kernel(StepP p, PeerArgs a) {
    __shared__ int s_bad;
    const int tid = threadIdx.x, nt = blockDim.x, world = p.world;
    const int parity = (int)(a.token & 1ull);
    const i64 slot = ((i64)parity * world + a.rank) * a.rec_max;
    if (tid == 0) s_bad = *((volatile int *)a.err);
    p.cand_send = a.cand + slot;                               // phase 1 writes the record into this rank's own slot
    candidate_phase(p, nullptr);
    __syncthreads();
    if (s_bad) return;                                         // an earlier step timed out: do not wait again (the host reports it)
    const int rec = CAND_HDR + p.m;
    const double *mine = a.cand + slot;
    for (int g = 0; g < world; ++g) {
        if (g == a.rank) continue;
        double *dst = a.peer_cand[g] + slot;
        for (int r = tid; r < rec; r += nt) dst[r] = mine[r];
    }
    __threadfence_system();                                    // every thread's remote stores before the token
    __syncthreads();
    if (tid < world) {
        unsigned long long *f = a.peer_flag[tid] + (parity * world + a.rank);
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(a.token) : "memory");
    }
    if (tid < world) {
        const unsigned long long *f = a.flag + (parity * world + tid);
        unsigned long long v;
        long long spins = 0;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
            if (v >= a.token) break;
            if (++spins > (1ll << 22)) { s_bad = 1; *a.err = 1; break; }
        }
    }
    __syncthreads();
    if (s_bad) return;
    p.cand_stride = a.rec_max;
    reflect_phase(p, a.cand + (i64)parity * world * a.rec_max, nullptr);
}
__global__ void __launch_bounds__(1024) qp_pivot_fused_kernel(StepP p) {   // one rank: no exchange between the phases
    extern __shared__ double ush[];                    // m doubles: the candidate column
    candidate_phase(p, ush);
    __threadfence_block();
    __syncthreads();
    reflect_phase(p, p.cand_send, ush);
}

// ---- wide kernel: F(:, j), pivot-row update, norm downdate, candidate partials.  One warp per (not yet chosen) column. ------
__global__ void __launch_bounds__(256) qp_wide_kernel(StepP p) {
    extern __shared__ double smem[];
    double *vs = smem;                           // v(k:m)
    double *auxs = smem + (p.m - p.k);           // QNB
    double *vrow = auxs + QNB;                   // V(k, 0:j+1) (the current reflector has a 1 there), zero beyond
    __shared__ double sh_v[8];
    __shared__ int sh_p[8], sh_c[8];
    const int k = p.k, m = p.m, j = p.j, len = m - k;
    for (int r = threadIdx.x; r < len; r += blockDim.x) vs[r] = p.vglob[k + r];
    if (threadIdx.x < QNB) {
        auxs[threadIdx.x] = (threadIdx.x < j) ? p.aux[threadIdx.x] : 0.0;
        vrow[threadIdx.x] = (threadIdx.x < j) ? p.Vb[(i64)threadIdx.x * m + k] : (threadIdx.x == j ? 1.0 : 0.0);
    }
    __syncthreads();
    const double tau = p.tau[k];
    const int lane = threadIdx.x & 31;
    const int w0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    Best best; best.v = -2.0; best.pos = 0x7fffffff; best.c = -1;
    // Software pipeline across columns: the scalars, the F row and the first 128 rows of column c + nw are requested before
    // column c is reduced, so a warp always has two columns' worth of loads in flight (the sweep is latency-, not issue-bound).
    struct Pre { double v1, v2, fi, x0, x1, x2, x3; int ps; };
    auto prefetch = [&](int c) {
        Pre q;
        const double *col = p.A + (i64)c * p.lda + k;
        q.v1 = p.vn1[c]; q.v2 = p.vn2[c]; q.ps = p.pos[c];
        q.fi = (lane < j) ? p.Ft[(i64)c * QNB + lane] : 0.0;
        q.x0 = (lane < len) ? col[lane] : 0.0; q.x1 = (lane + 32 < len) ? col[lane + 32] : 0.0;
        q.x2 = (lane + 64 < len) ? col[lane + 64] : 0.0; q.x3 = (lane + 96 < len) ? col[lane + 96] : 0.0;
        return q;
    };
    Pre cur;
    if (w0 < p.nloc) cur = prefetch(w0);
    for (int c = w0; c < p.nloc; c += nw) {
        Pre nxt = cur;
        if (c + nw < p.nloc) nxt = prefetch(c + nw);
        const double v1 = cur.v1, v2 = cur.v2;
        const int ps = cur.ps;
        double fi = cur.fi;
        double x0 = cur.x0, x1 = cur.x1, x2 = cur.x2, x3 = cur.x3;
        cur = nxt;
        if (v1 < 0.0) continue;                                   // already a pivot (warp-uniform)
        double *col = p.A + (i64)c * p.lda + k;
        double *f = p.Ft + (i64)c * QNB;
        double dot = 0.0;
        const double a0l = x0;
        for (int r = lane;;) {
            const int rn = r + 128;
            double y0 = 0.0, y1 = 0.0, y2 = 0.0, y3 = 0.0;
            if (rn < len) {                                       // next batch in flight while this one is reduced
                y0 = col[rn]; y1 = (rn + 32 < len) ? col[rn + 32] : 0.0; y2 = (rn + 64 < len) ? col[rn + 64] : 0.0; y3 = (rn + 96 < len) ? col[rn + 96] : 0.0;
            }
            if (r < len) dot = fma(x0, vs[r], dot);
            if (r + 32 < len) dot = fma(x1, vs[r + 32], dot);
            if (r + 64 < len) dot = fma(x2, vs[r + 64], dot);
            if (r + 96 < len) dot = fma(x3, vs[r + 96], dot);
            if (rn >= len) break;
            r = rn; x0 = y0; x1 = y1; x2 = y2; x3 = y3;
        }
        // one shuffle tree for the three sums: A(k:m,c)^T v, F(c,0:j) aux and V(k,0:j) F(c,0:j)^T
        double corr = fi * auxs[lane], rup = (lane < j) ? fi * vrow[lane] : 0.0;
        for (int o = 16; o > 0; o >>= 1) {
            dot += __shfl_xor_sync(0xffffffffu, dot, o);
            corr += __shfl_xor_sync(0xffffffffu, corr, o);
            rup += __shfl_xor_sync(0xffffffffu, rup, o);
        }
        const double a0 = __shfl_sync(0xffffffffu, a0l, 0);
        const double fj = fma(tau, dot, corr);                    // F(c, j) = tau A(k:m,c)^T v - tau F(c,0:j) V^T v
        if (lane == j) { fi = fj; f[j] = fj; }
        const double akc = a0 - (rup + fj);                       // A(k, c) -= V(k, 0:j+1) F(c, 0:j+1)^T   (V(k, j) = 1)
        if (lane == 0) col[0] = akc;
        double vnew = v1;
        if (p.downdate && v1 != 0.0) {
            const double t = fabs(akc) / v1;
            const double temp = p.ps_formula ? fmax(0.0, (1.0 + t) * (1.0 - t)) : fmax(1.0 - t * t, 0.0);
            const double q2 = v1 / v2;
            if (temp * (q2 * q2) <= TOL3Z) {
                // rebuild the column's up-to-date tail from the pending reflectors and recompute its norm (warp-uniform branch)
                dd acc; acc.hi = 0.0; acc.lo = 0.0;
                for (int r0 = 1; r0 < len; r0 += 32) {
                    const int r = r0 + lane;
                    double wv = (r < len) ? col[r] : 0.0;
                    for (int i = 0; i <= j; ++i) {
                        const double fc = __shfl_sync(0xffffffffu, fi, i);
                        if (r < len) wv = fma(-p.Vb[(i64)i * m + k + r], fc, wv);
                    }
                    if (r < len) acc = dd_add_sq(acc, wv);
                }
                acc = dd_warp_sum(acc);
                vnew = (len > 1) ? dd_sqrt(acc) : 0.0;
                if (lane == 0) { p.vn1[c] = vnew; p.vn2[c] = vnew; }
            } else {
                vnew = v1 * sqrt(temp);
                if (lane == 0) p.vn1[c] = vnew;
            }
        }
        if (better(vnew, ps, best.v, best.pos)) { best.v = vnew; best.pos = ps; best.c = c; }
    }
    block_best_store(best, sh_v, sh_p, sh_c, p.part_v, p.part_pos, p.part_c);
}

// ---- block end: A(r0:m, :) -= V(r0:m, 0:jb) F(:, 0:jb)^T, streamed once (read + write) with the reflector block in shared memory.
// CTA tile: 32 columns x all remaining rows in chunks of 128; a warp owns 4 columns, a lane 4 rows of each chunk.
constexpr int UR = 128, UC = 32;
__global__ void __launch_bounds__(256) qp_block_update_kernel(double *A, i64 lda, int m, int nloc, int r0, int jb, const double *__restrict__ Vb,
                                                              const double *__restrict__ Ft) {
    __shared__ double Vs[QNB * UR];          // Vs[i * UR + r]
    __shared__ double Fs[QNB * UC];          // Fs[i * UC + c]
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int ct = blockIdx.x; ct * UC < nloc; ct += gridDim.x) {
        const int c0 = ct * UC;
        __syncthreads();
        for (int e = tid; e < QNB * UC; e += 256) {
            const int i = e & (QNB - 1), cc = e / QNB;             // consecutive threads read consecutive i of one column (Ft is [c][i])
            Fs[i * UC + cc] = (i < jb && c0 + cc < nloc) ? Ft[(i64)(c0 + cc) * QNB + i] : 0.0;
        }
        for (int rc = r0; rc < m; rc += UR) {
            __syncthreads();
            for (int e = tid; e < QNB * UR; e += 256) {
                const int r = e & (UR - 1), i = e / UR;
                Vs[e] = (i < jb && rc + r < m) ? Vb[(i64)i * m + rc + r] : 0.0;
            }
            // this warp's 4 x 4 x 32-row block of C: loads in flight during the product
            double cv[4][4];
            double *cp[4];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const int c = c0 + 4 * w + cc;
                cp[cc] = A + (i64)c * lda + rc + lane;
#pragma unroll
                for (int q = 0; q < 4; ++q) cv[cc][q] = (c < nloc && rc + lane + 32 * q < m) ? cp[cc][32 * q] : 0.0;
            }
            __syncthreads();
            double acc[4][4];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[cc][q] = 0.0;
#pragma unroll 4
            for (int i = 0; i < QNB; ++i) {
                double v[4], f[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) v[q] = Vs[i * UR + lane + 32 * q];
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) f[cc] = Fs[i * UC + 4 * w + cc];
#pragma unroll
                for (int cc = 0; cc < 4; ++cc)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[cc][q] = fma(v[q], f[cc], acc[cc][q]);
            }
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const int c = c0 + 4 * w + cc;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (c < nloc && rc + lane + 32 * q < m) cp[cc][32 * q] = cv[cc][q] - acc[cc][q];
            }
        }
    }
}

__global__ void qp_iota_kernel(int *p, int n) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) p[e] = e;
}
__global__ void qp_perm_out_kernel(const int *perm, int n, double *jpvt, int *posg) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const int g = perm[e];
        if (jpvt) jpvt[e] = (double)g;
        if (posg) posg[g] = e;
    }
}
// T(:, p - k) for every position p >= k: pivot positions (p < kmax) from X = R11^{-1} Rpiv(0:k, k:kmax), the others from the
// gathered solved columns Tall (k rows, ld ldall; column of global column g at slot (g / per) * per_pad + g % per)
__global__ void qp_scatter_t_kernel(const int *perm, int n_global, int k, int kmax, const double *X, const double *Tall, i64 ldall, int per, int per_pad,
                                    double *T, i64 ldt) {
    const i64 total = (i64)(n_global - k) * k;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
        const int r = (int)(e % k);
        const int pp = (int)(e / k) + k;
        double v;
        if (pp < kmax) v = X[(i64)(pp - k) * k + r];
        else {
            const int g = perm[pp];
            const i64 slot = (i64)(g / per) * per_pad + (g % per);
            v = Tall[slot * ldall + r];
        }
        T[(i64)(pp - k) * ldt + r] = v;
    }
}
// classic dgeqp3 layout: out(:, p) = column at position p — R on and above the diagonal, reflectors below (p < kmax);
// R rows for the never-chosen columns
__global__ void qp_materialize_kernel(const double *A, i64 lda, int m, int n, int kmax, const int *perm, const double *Rpiv, int ldr, const double *Vall,
                                      double *out, i64 ldo) {
    const i64 total = (i64)m * n;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
        const int r = (int)(e % m), pp = (int)(e / m);
        double v;
        if (pp < kmax) v = (r <= pp) ? Rpiv[(i64)pp * ldr + r] : Vall[(i64)pp * m + r];
        else v = A[(i64)perm[pp] * lda + r];
        out[(i64)pp * ldo + r] = v;
    }
}

struct QpWork {
    DBuf vn1, vn2, Ft, Vall, Rpiv, tau, aux, vglob, part_v, cand_send, cand_all;
    int *pos = nullptr, *perm = nullptr, *part_pos = nullptr, *part_c = nullptr;
    int kmax = 0;
    ~QpWork() { dfree(pos); dfree(perm); dfree(part_pos); dfree(part_c); }
};

// The factorisation proper.  A (m x nloc, ld lda): this rank's columns [col0, col0 + nloc) of an m x n_global matrix
// (sharded == false: nloc == n_global, no communication).  On exit W holds perm (position -> global column), Rpiv, Vall,
// tau, and the never-chosen local columns of A hold their R entries in rows 0..kmax-1.
void qp_factor(double *A, i64 lda, int m, int nloc, int col0, int n_global, bool sharded, QpWork &W) {
    Ctx &c = ctx();
    const int world = sharded ? c.world : 1;
    const int kmax = std::min(m, n_global);
    W.kmax = kmax;
    int per_sm = 0;
    RSVD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, qp_wide_kernel, 256, (size_t)(m + 2 * QNB) * sizeof(double)));
    const int wide_blocks = std::max(1, std::min((nloc + 7) / 8, c.sms * std::max(1, per_sm)));   // one resident wave: no CTA waits for a slot
    W.vn1.alloc((size_t)nloc + 1); W.vn2.alloc((size_t)nloc + 1); W.Ft.alloc((size_t)QNB * nloc + 1);
    W.Vall.alloc((size_t)m * (kmax + QNB)); W.Rpiv.alloc((size_t)kmax * kmax + 1); W.tau.alloc((size_t)kmax + 1); W.aux.alloc(QNB);
    W.vglob.alloc((size_t)m + 1); W.part_v.alloc((size_t)wide_blocks);
    W.cand_send.alloc((size_t)(CAND_HDR + m)); W.cand_all.alloc((size_t)(CAND_HDR + m) * world);
    W.pos = (int *)dalloc_bytes((size_t)(nloc + 1) * sizeof(int));
    W.perm = (int *)dalloc_bytes((size_t)(n_global + 1) * sizeof(int));
    W.part_pos = (int *)dalloc_bytes((size_t)wide_blocks * sizeof(int));
    W.part_c = (int *)dalloc_bytes((size_t)wide_blocks * sizeof(int));
    if (g_status) return;
    set_zero(W.Rpiv.p, (size_t)kmax * kmax);
    set_zero(W.Ft.p, (size_t)QNB * nloc);
    qp_iota_kernel<<<std::max(1, std::min((n_global + 255) / 256, c.sms * 8)), 256, 0, c.stream>>>(W.perm, n_global);
    qp_init_kernel<<<wide_blocks, 256, 0, c.stream>>>(A, lda, m, nloc, col0, W.vn1.p, W.vn2.p, W.pos, W.part_v.p, W.part_pos, W.part_c);
    count_launch(2);
    StepP p;
    p.A = A; p.lda = lda; p.m = m; p.nloc = nloc; p.col0 = col0; p.n_global = n_global; p.kmax = kmax; p.world = world;
    p.vn1 = W.vn1.p; p.vn2 = W.vn2.p; p.pos = W.pos; p.perm = W.perm; p.Ft = W.Ft.p; p.vglob = W.vglob.p;
    p.Rpiv = W.Rpiv.p; p.ldr = kmax; p.tau = W.tau.p; p.aux = W.aux.p;
    p.part_v = W.part_v.p; p.part_pos = W.part_pos; p.part_c = W.part_c; p.nparts = wide_blocks;
    p.cand_send = W.cand_send.p; p.cand_all = W.cand_all.p;
    p.cand_stride = 0;
    // several ranks: exchange the candidate records through the peer mailboxes when every rank could map them (collective
    // set-up on first use), else through ncclAllGather
    const bool use_peer = world > 1 && peer_mailbox((size_t)(CAND_HDR + m));
    PeerArgs pa;
    if (use_peer) {
        pa.peer_cand = c.peer.d_peer_cand; pa.peer_flag = c.peer.d_peer_flag; pa.cand = c.peer.cand; pa.flag = c.peer.flag;
        pa.rec_max = (i64)c.peer.rec_max; pa.rank = c.rank; pa.err = c.peer.err; pa.token = 0;
    }
    const bool blocked_range = kmax > 128;            // dgeqp3: NB = 32 < sminmn and NX = 128 < sminmn
    int k0 = 0;
    for (int k = 0; k < kmax && !g_status; ++k) {
        p.k = k; p.j = k - k0; p.Vb = W.Vall.p + (i64)k0 * m;
        p.ps_formula = (blocked_range && k < kmax - 128) ? 1 : 0;   // dlaqps (1+t)(1-t) in the blocked range, dlaqp2 1-t^2 in the tail
        p.downdate = (k < kmax - 1) ? 1 : 0;                        // dlaqps/dlaqp2: no downdate at the last row
        if (world == 1) {
            qp_pivot_fused_kernel<<<1, 1024, (size_t)m * sizeof(double), c.stream>>>(p);
            count_launch();
        } else if (use_peer) {
            pa.token = c.peer.token + (unsigned long long)k + 1ull;
            qp_pivot_peer_kernel<<<1, 1024, 0, c.stream>>>(p, pa);
            count_launch();
        } else {
            qp_candidate_kernel<<<1, 1024, 0, c.stream>>>(p);
            allgather(W.cand_send.p, W.cand_all.p, (size_t)(CAND_HDR + m));
            qp_reflect_kernel<<<1, 1024, 0, c.stream>>>(p);
            count_launch(2);
        }
        if (k + 1 < n_global) {
            const size_t sh = (size_t)(m - k + 2 * QNB) * sizeof(double);
            qp_wide_kernel<<<wide_blocks, 256, sh, c.stream>>>(p);
            count_launch();
        }
        if (p.j == QNB - 1 || k == kmax - 1) {        // block end: apply the block reflector to the rows below the block
            const int r0 = k + 1, jb = p.j + 1;
            if (m - r0 > 0 && k + 1 < n_global) {
                const int ub = std::max(1, std::min((nloc + UC - 1) / UC, c.sms * 16));
                qp_block_update_kernel<<<ub, 256, 0, c.stream>>>(A, lda, m, nloc, r0, jb, p.Vb, W.Ft.p);
                count_launch();
            }
            k0 = k + 1;
        }
    }
    RSVD_CUDA(cudaGetLastError());
    if (use_peer) {
        c.peer.token += (unsigned long long)kmax;     // the same on every rank: tokens and buffer parity continue across factorisations
        int bad = 0;
        RSVD_CUDA(cudaMemcpyAsync(&bad, c.peer.err, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
        RSVD_CUDA(cudaStreamSynchronize(c.stream));
        if (bad) set_error("rsvd_b200: pivoted QR: a peer's candidate record did not arrive (peer mailbox wait timed out on rank %d)", c.rank);
    }
}

}  // namespace

// Row limit of the blocked kernel (one warp per column, the reflector in shared memory).  The default, 2048, is the widest
// sketch the benchmarks use; sketches of up to 4096 rows (the Jacobi kernel's limit on k + p) go to the one-reflector-per-step kernel
// unless option "qr_blocked_rows" raises the limit (up to 4096).
bool geqp3_blocked_ok(i64 m, i64 n) {
    const i64 lim = std::min<i64>(4096, std::max<i64>(1, ctx().qr_blocked_rows));
    return m >= 1 && m <= lim && n >= 1 && n < (1ll << 31) - 64;
}

// classic interface (dgeqp3 layout in place, jpvt 0-based as doubles, optional tau[min(m,n)])
void geqp3_blocked(double *A, i64 lda, i64 m, i64 n, double *jpvt_out, double *tau_out) {
    if (g_status) return;
    ensure_init();
    Ctx &c = ctx();
    QpWork W;
    qp_factor(A, lda, (int)m, (int)n, 0, (int)n, false, W);
    if (g_status) return;
    const int grid = (int)std::max((i64)1, std::min((n + 255) / 256, (i64)c.sms * 8));
    qp_perm_out_kernel<<<grid, 256, 0, c.stream>>>(W.perm, (int)n, jpvt_out, nullptr);
    count_launch();
    DBuf out((size_t)m * n);
    if (g_status) return;
    const i64 total = m * n;
    qp_materialize_kernel<<<(int)std::max((i64)1, std::min((total + 255) / 256, (i64)c.sms * 16)), 256, 0, c.stream>>>(A, lda, (int)m, (int)n, W.kmax, W.perm, W.Rpiv.p,
                                                                                                              W.kmax, W.Vall.p, out.p, m);
    count_launch();
    copy_matrix(out.p, m, A, lda, m, n);
    if (tau_out) copy_matrix(W.tau.p, W.kmax, tau_out, W.kmax, W.kmax, 1);
}

// Pivoted QR + interpolation matrix in one go (the tails RRA:1938-1956 and RRA:1836-1850):
//   Y (r x nloc, ld ldy, DESTROYED): this rank's columns [col0, col0 + nloc) of an r x n_global matrix; `per` = columns per
//   rank of the (regular) sharding, so that global column g lives on rank g / per (sharded == false: one rank holds all).
//   I (n_global doubles, 0-based permutation) and T = R11(0:k,0:k)^{-1} R(0:k, k:n) (k x (n_global - k)) are produced on EVERY rank.
void geqp3_id(double *Y, i64 ldy, i64 r, i64 nloc, i64 col0, i64 n_global, i64 per, bool sharded, i64 k, double *I, double *T, i64 ldt) {
    if (g_status) return;
    ensure_init();
    Ctx &c = ctx();
    const int world = sharded ? c.world : 1;
    QpWork W;
    qp_factor(Y, ldy, (int)r, (int)nloc, (int)col0, (int)n_global, sharded, W);
    if (g_status) return;
    const int kmax = W.kmax;
    const int grid = (int)std::max((i64)1, std::min((n_global + 255) / 256, (i64)c.sms * 8));
    qp_perm_out_kernel<<<grid, 256, 0, c.stream>>>(W.perm, (int)n_global, I, nullptr);
    count_launch();
    if (n_global <= k) return;
    // solved local columns: Y(0:k, :) <- R11^{-1} Y(0:k, :)   (in place, all local columns; chosen ones are ignored later)
    trsm_left_upper(W.Rpiv.p, kmax, k, Y, ldy, nloc);
    DBuf X;                                              // pivot positions k .. kmax-1
    if (kmax > k) {
        X.alloc((size_t)k * (kmax - k));
        copy_matrix(W.Rpiv.p + (i64)k * kmax, kmax, X.p, k, k, kmax - k);
        trsm_left_upper(W.Rpiv.p, kmax, k, X.p, k, kmax - k);
    }
    const i64 total = (n_global - k) * k;
    const int sgrid = (int)std::max((i64)1, std::min((total + 255) / 256, (i64)c.sms * 16));
    if (world == 1) {
        qp_scatter_t_kernel<<<sgrid, 256, 0, c.stream>>>(W.perm, (int)n_global, (int)k, kmax, X.p, Y, ldy, (int)n_global, (int)n_global, T, ldt);
        count_launch();
    } else {
        // every rank contributes its k x per block (short last blocks are padded), then scatters all of them by position
        DBuf mine((size_t)k * per), all((size_t)k * per * world);
        if (g_status) return;
        if (nloc < per) set_zero(mine.p, (size_t)k * per);
        copy_matrix(Y, ldy, mine.p, k, k, nloc);
        allgather(mine.p, all.p, (size_t)k * per);
        qp_scatter_t_kernel<<<sgrid, 256, 0, c.stream>>>(W.perm, (int)n_global, (int)k, kmax, X.p, all.p, k, (int)per, (int)per, T, ldt);
        count_launch();
    }
}

}  // namespace rsvd
