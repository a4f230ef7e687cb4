// runtime.cu — context, out-of-band error channel, stream-ordered device memory, pinned host memory and
// chunked host<->device staging.  The reference has none of this (its MKL build never leaves the host; its
// cuBLAS build re-uploads both operands on EVERY GEMM call, matrix_vector_functions_mkl_and_cublas.c:566-577);
// here one upload per API call keeps A resident for all 2q passes.
#include "common.cuh"
#include <mutex>
#include <unordered_map>
#include <time.h>
#include <algorithm>
#include <stdarg.h>
#include <mutex>

namespace rsvd {

int g_status = 0;
unsigned long long g_launches = 0;
int g_single_device = 0;
static char g_errbuf[1024] = "";
static std::mutex g_err_mu;

void set_error(const char *fmt, ...) {
    std::lock_guard<std::mutex> lk(g_err_mu);
    if (g_status == 0) {   // keep the first error
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(g_errbuf, sizeof(g_errbuf), fmt, ap);
        va_end(ap);
        g_status = 1;
        if (getenv("RSVD_B200_VERBOSE")) fprintf(stderr, "[rsvd_b200] error: %s\n", g_errbuf);
    }
}

// One context per (thread, device).  The calling thread of a C driver uses the primary context; in single-process
// multi-GPU mode (multi.cu) every worker thread binds its own context, so all of the device layer below ctx() is
// device-agnostic and re-entrant across workers.
static Ctx g_ctx0;
static thread_local Ctx *t_ctx = nullptr;
Ctx &ctx() { return t_ctx ? *t_ctx : g_ctx0; }
void bind_ctx(Ctx *c) {
    t_ctx = c;
    if (c && c->device >= 0) cudaSetDevice(c->device);
}

int init_ctx(Ctx &c, int device) {
    if (c.inited) return 0;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        (void)cudaGetLastError();
        set_error("rsvd_b200: no CUDA device available (%s); there is no CPU fallback", cudaGetErrorString(e));
        return 1;
    }
    if (device < 0) {
        const char *s = getenv("RSVD_B200_DEVICE");
        if (!s) s = getenv("LOCAL_RANK");
        if (!s) s = getenv("RSVD_B200_DEVICES");            // "0-7" / "2,3": the primary context sits on the first listed device
        device = (s && *s >= '0' && *s <= '9') ? atoi(s) : 0;
        if (device >= ndev) device = device % ndev;
    }
    RSVD_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    RSVD_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("rsvd_b200: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return 1;
    }
    c.device = device;
    c.sms = prop.multiProcessorCount;
    RSVD_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    RSVD_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
    RSVD_CUDA(cudaStreamCreateWithFlags(&c.aux_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) RSVD_CUDA(cudaEventCreateWithFlags(&c.aux_ev[i], cudaEventDisableTiming));
    // keep freed blocks in the pool: cudaMallocAsync then costs microseconds, not a driver round trip
    cudaMemPool_t pool;
    RSVD_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    unsigned long long thr = ~0ull;
    RSVD_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    RSVD_CUDA(cudaMalloc(&c.d_flag, 64 * sizeof(int)));
    RSVD_CUDA(cudaMemset(c.d_flag, 0, 64 * sizeof(int)));
    RSVD_CUDA(cudaHostAlloc(&c.h_flag, 64 * sizeof(int), cudaHostAllocDefault));
    const char *v = getenv("RSVD_B200_VERBOSE");
    c.verbose = v ? atoi(v) : 0;
    const char *nc = getenv("RSVD_B200_NO_SKETCH_CLUSTER");
    c.no_sketch_cluster = nc ? atoi(nc) : 0;
    const char *fg = getenv("RSVD_B200_FORCE_GENERIC_GEMM");
    c.force_generic_gemm = fg ? atoi(fg) : 0;
    c.inited = (g_status == 0);
    return g_status;
}
static int init_device(int device) { return init_ctx(ctx(), device); }

void ensure_init() {
    if (!ctx().inited) init_device(-1);
    else cudaSetDevice(ctx().device);
}

// ---- device memory ---------------------------------------------------------------------------------------------------
// cudaMallocAsync / cudaFreeAsync on the context's stream, with a small per-context cache of freed work buffers on top.
// Why the cache: the stream-ordered pool usually answers in microseconds, but now and then it goes back to the driver for
// fresh memory — measured on the 100000 x 50000 blocked QB: single cudaMallocAsync calls of a 153 MB panel buffer taking
// 18, 39, 59 ms on the host, and 100-700 ms holes in the GPU timeline when such a call fell between two short kernels.  The
// algorithms ask for the same few sizes over and over (m x l panels, l x l blocks), so a freed block of 1 MB .. 2 GB is kept
// (up to 6 GB per context, oldest dropped first) and handed to the next request of (nearly) the same size.  Reuse is safe
// because both uses are ordered on the same stream — exactly the guarantee cudaFreeAsync/cudaMallocAsync give themselves.
namespace {
struct BlockCache {
    std::mutex mu;
    std::unordered_map<void *, size_t> live;                 // blocks handed out by this context
    std::vector<std::pair<void *, size_t>> idle;             // freed, reusable (oldest first)
    size_t idle_bytes = 0;
};
constexpr size_t kCacheMin = (size_t)1 << 20, kCacheMax = (size_t)2 << 30, kCacheTotal = (size_t)6 << 30;
BlockCache *cache_of(Ctx &c) {
    if (!c.block_cache) c.block_cache = new BlockCache();
    return (BlockCache *)c.block_cache;
}
void cache_flush(Ctx &c) {
    BlockCache *bc = cache_of(c);
    std::lock_guard<std::mutex> lk(bc->mu);
    for (auto &b : bc->idle) cudaFreeAsync(b.first, c.stream);
    bc->idle.clear();
    bc->idle_bytes = 0;
}
}  // namespace

void *dalloc_bytes(size_t bytes) {
    ensure_init();
    Ctx &c = ctx();
    if (!c.inited) return nullptr;
    void *p = nullptr;
    if (bytes == 0) bytes = 8;
    BlockCache *bc = cache_of(c);
    if (bytes >= kCacheMin && bytes <= kCacheMax) {
        std::lock_guard<std::mutex> lk(bc->mu);
        for (size_t i = bc->idle.size(); i-- > 0;) {           // newest first
            const size_t sz = bc->idle[i].second;
            if (sz >= bytes && sz - bytes <= bytes / 8) {
                p = bc->idle[i].first;
                bc->idle_bytes -= sz;
                bc->idle.erase(bc->idle.begin() + (long)i);
                bc->live[p] = sz;
                return p;
            }
        }
    }
    struct timespec t0, t1;
    const bool timed = c.verbose != 0;
    if (timed) clock_gettime(CLOCK_MONOTONIC, &t0);
    cudaError_t e = cudaMallocAsync(&p, bytes, c.stream);
    if (e != cudaSuccess) {                                  // give the cached blocks back and try once more
        (void)cudaGetLastError();
        cache_flush(c);
        cudaStreamSynchronize(c.stream);
        e = cudaMallocAsync(&p, bytes, c.stream);
    }
    if (timed) {
        clock_gettime(CLOCK_MONOTONIC, &t1);
        const double ms = (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6;
        if (ms > 2.0) fprintf(stderr, "[rsvd_b200] cudaMallocAsync of %.1f MB took %.1f ms on the host\n", bytes / 1048576.0, ms);
    }
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        set_error("rsvd_b200: device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        return nullptr;
    }
    if (bytes >= kCacheMin && bytes <= kCacheMax) {
        std::lock_guard<std::mutex> lk(bc->mu);
        bc->live[p] = bytes;
    }
    return p;
}
double *dalloc(size_t n) { return (double *)dalloc_bytes(n * sizeof(double)); }
void dfree(void *p) {
    if (!p) return;
    Ctx &c = ctx();
    BlockCache *bc = cache_of(c);
    {
        std::lock_guard<std::mutex> lk(bc->mu);
        auto it = bc->live.find(p);
        if (it != bc->live.end()) {
            const size_t sz = it->second;
            bc->live.erase(it);
            if (!c.no_block_cache) {
                bc->idle.push_back({p, sz});
                bc->idle_bytes += sz;
                while (bc->idle_bytes > kCacheTotal && !bc->idle.empty()) {       // drop the oldest
                    RSVD_CUDA(cudaFreeAsync(bc->idle.front().first, c.stream));
                    bc->idle_bytes -= bc->idle.front().second;
                    bc->idle.erase(bc->idle.begin());
                }
                return;
            }
        }
    }
    RSVD_CUDA(cudaFreeAsync(p, c.stream));
}
void release_cached_blocks() { if (ctx().inited) cache_flush(ctx()); }

// ---- host <-> device -------------------------------------------------------------------------------
static bool host_is_pinned(const void *p) {
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

struct Staging {
    static constexpr size_t CHUNK = 64ull << 20;
    void *buf[2] = {nullptr, nullptr};
    cudaEvent_t ev[2];
    bool ok = false;
    void init() {
        if (ok) return;
        for (int i = 0; i < 2; ++i) {
            RSVD_CUDA(cudaHostAlloc(&buf[i], CHUNK, cudaHostAllocDefault));
            RSVD_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
        }
        ok = true;
    }
};
static Staging &staging() {
    Ctx &c = ctx();
    if (!c.staging) c.staging = new Staging();
    return *(Staging *)c.staging;
}
#define g_staging (staging())

static int copy_h2d(double *d, const double *h, size_t n) {
    ensure_init();
    if (!ctx().inited) return 1;
    size_t bytes = n * 8;
    cudaStream_t st = ctx().stream;
    if (bytes == 0) return 0;
    if (host_is_pinned(h) || bytes < (1u << 20)) {
        RSVD_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st));
        RSVD_CUDA(cudaStreamSynchronize(st));
        return g_status;
    }
    // pageable source: double-buffered memcpy -> pinned -> DMA, so the CPU copy of chunk i+1 overlaps the DMA of chunk i
    g_staging.init();
    size_t off = 0; int i = 0;
    bool used[2] = {false, false};
    while (off < bytes) {
        size_t c = bytes - off < Staging::CHUNK ? bytes - off : Staging::CHUNK;
        if (used[i]) RSVD_CUDA(cudaEventSynchronize(g_staging.ev[i]));
        memcpy(g_staging.buf[i], (const char *)h + off, c);
        RSVD_CUDA(cudaMemcpyAsync((char *)d + off, g_staging.buf[i], c, cudaMemcpyHostToDevice, st));
        RSVD_CUDA(cudaEventRecord(g_staging.ev[i], st));
        used[i] = true;
        off += c; i ^= 1;
    }
    RSVD_CUDA(cudaStreamSynchronize(st));
    return g_status;
}

static int copy_d2h(double *h, const double *d, size_t n) {
    ensure_init();
    if (!ctx().inited) return 1;
    size_t bytes = n * 8;
    cudaStream_t st = ctx().stream;
    if (bytes == 0) return 0;
    if (host_is_pinned(h) || bytes < (1u << 20)) {
        RSVD_CUDA(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, st));
        RSVD_CUDA(cudaStreamSynchronize(st));
        return g_status;
    }
    g_staging.init();
    size_t off = 0; int i = 0;
    size_t pend_off[2] = {0, 0}, pend_c[2] = {0, 0};
    bool used[2] = {false, false};
    while (off < bytes || used[0] || used[1]) {
        if (used[i]) {
            RSVD_CUDA(cudaEventSynchronize(g_staging.ev[i]));
            memcpy((char *)h + pend_off[i], g_staging.buf[i], pend_c[i]);
            used[i] = false;
        }
        if (off < bytes) {
            size_t c = bytes - off < Staging::CHUNK ? bytes - off : Staging::CHUNK;
            RSVD_CUDA(cudaMemcpyAsync(g_staging.buf[i], (const char *)d + off, c, cudaMemcpyDeviceToHost, st));
            RSVD_CUDA(cudaEventRecord(g_staging.ev[i], st));
            pend_off[i] = off; pend_c[i] = c; used[i] = true;
            off += c;
        }
        i ^= 1;
    }
    return g_status;
}

}  // namespace rsvd

using namespace rsvd;

extern "C" {

int rsvd_b200_init(int device) {
    const int rc = init_device(device);
    if (!rc) (void)pool_size();      // RSVD_B200_DEVICES: start the workers (and their NCCL ranks) now, not inside the first timed call
    return rc;
}
int rsvd_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    return n;
}
int rsvd_b200_status(void) { return g_status; }
const char *rsvd_b200_last_error(void) { return g_errbuf; }
void rsvd_b200_clear_error(void) { g_status = 0; g_errbuf[0] = 0; }
void *rsvd_b200_stream(void) { ensure_init(); return (void *)ctx().stream; }
void rsvd_b200_sync(void) { if (ctx().inited) RSVD_CUDA(cudaStreamSynchronize(ctx().stream)); }
unsigned long long rsvd_b200_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

static unsigned long long g_seed_opt = 777ull;
void rsvd_b200_set_option(const char *name, rsvd_i64 value) {
    if (!strcmp(name, "seed")) g_seed_opt = (unsigned long long)value;
    else if (!strcmp(name, "verbose")) ctx().verbose = (int)value;
    else if (!strcmp(name, "force_generic_gemm")) ctx().force_generic_gemm = (int)value;
    else if (!strcmp(name, "force_qr_fallback")) ctx().force_qr_fallback = (int)value;
    else if (!strcmp(name, "row0")) ctx().row0 = value;
    else if (!strcmp(name, "jacobi_transpose")) ctx().jacobi_transpose = (int)value;
    else if (!strcmp(name, "no_chol_dataflow")) ctx().no_chol_dataflow = (int)value;
    else if (!strcmp(name, "no_live_replay")) ctx().no_live_replay = (int)value;
    else if (!strcmp(name, "no_block_cache")) { ctx().no_block_cache = (int)value; if (value) release_cached_blocks(); }
    else if (!strcmp(name, "force_unblocked_qr")) ctx().force_unblocked_qr = (int)value;
    else if (!strcmp(name, "no_sketch_cluster")) ctx().no_sketch_cluster = (int)value;
    else if (!strcmp(name, "qr_blocked_rows")) ctx().qr_blocked_rows = (int)value;
    else if (!strcmp(name, "m_global")) ctx().m_global = value;
    else if (!strcmp(name, "single_device")) g_single_device = (int)value;
    else set_error("rsvd_b200_set_option: unknown option '%s'", name);
}
rsvd_i64 rsvd_b200_get_option(const char *name) {
    if (!strcmp(name, "seed")) return (rsvd_i64)g_seed_opt;
    if (!strcmp(name, "verbose")) return ctx().verbose;
    if (!strcmp(name, "force_generic_gemm")) return ctx().force_generic_gemm;
    if (!strcmp(name, "force_qr_fallback")) return ctx().force_qr_fallback;
    if (!strcmp(name, "last_gemm_path")) return g_last_gemm_path;
    if (!strcmp(name, "last_qr_path")) return ctx().last_qr_path;
    if (!strcmp(name, "qr_fallbacks")) return (rsvd_i64)ctx().qr_fallbacks;
    if (!strcmp(name, "sms")) { ensure_init(); return ctx().sms; }
    if (!strcmp(name, "rank")) return ctx().rank;
    if (!strcmp(name, "row0")) return ctx().row0;
    if (!strcmp(name, "m_global")) return ctx().m_global;
    if (!strcmp(name, "world")) return ctx().world;
    if (!strcmp(name, "single_device")) return g_single_device;
    if (!strcmp(name, "devices")) return pool_size();
    return -1;
}

double *rsvd_b200_dev_alloc(rsvd_i64 n) { return dalloc((size_t)n); }
void rsvd_b200_dev_free(double *d) { dfree(d); }
// ---- binary matrix files <-> device ------------------------------------------------------------
// The reference's file format (matrix_load_from_binary_file / matrix_write_to_binary_file, MVF:77-133, MVF64:78-135) is a
// header of two int32 (or int64) followed by ROW-major doubles.  A block of rows is a column-major (n x rows) matrix, so it
// is DMA'd as it lies in the file and transposed into the column-major device matrix by a kernel: the host never transposes
// and never holds the matrix.  File reads of block i+1 overlap the DMA + transpose of block i (two pinned staging buffers).
static int load_binary_dev(const char *path, int index_bits, double **dA_out, i64 *m_out, i64 *n_out) {
    ensure_init();
    Ctx &c = ctx();
    if (!c.inited) return 1;
    *dA_out = nullptr; *m_out = 0; *n_out = 0;
    FILE *fp = fopen(path, "rb");
    if (!fp) { set_error("rsvd_b200: cannot open %s", path); return 1; }
    i64 m = -1, n = -1;
    if (index_bits == 64) { long long h[2]; if (fread(h, 8, 2, fp) == 2) { m = h[0]; n = h[1]; } }
    else { int h[2]; if (fread(h, 4, 2, fp) == 2) { m = h[0]; n = h[1]; } }
    if (m < 0 || n < 0) { set_error("rsvd_b200: bad header in %s", path); fclose(fp); return 1; }
    const size_t row_bytes = (size_t)n * 8;
    if (row_bytes > Staging::CHUNK) { set_error("rsvd_b200: rows of %s exceed the %zu-byte staging buffer", path, Staging::CHUNK); fclose(fp); return 1; }
    g_staging.init();
    const i64 rb = row_bytes ? (i64)(Staging::CHUNK / row_bytes) : 1;
    double *A = dalloc((size_t)m * n + 1);
    double *ds[2] = {dalloc((size_t)rb * n + 1), dalloc((size_t)rb * n + 1)};
    bool used[2] = {false, false};
    int b = 0;
    for (i64 i0 = 0; i0 < m && !g_status; i0 += rb, b ^= 1) {
        const i64 rows = std::min(rb, m - i0);
        if (used[b]) RSVD_CUDA(cudaEventSynchronize(g_staging.ev[b]));
        if (fread(g_staging.buf[b], 8, (size_t)rows * n, fp) != (size_t)rows * n) { set_error("rsvd_b200: %s is truncated", path); break; }
        RSVD_CUDA(cudaMemcpyAsync(ds[b], g_staging.buf[b], (size_t)rows * row_bytes, cudaMemcpyHostToDevice, c.stream));
        RSVD_CUDA(cudaEventRecord(g_staging.ev[b], c.stream));
        used[b] = true;
        transpose(ds[b], n, A + i0, m, n, rows);          // (n x rows, ld n)^T -> rows i0.. of A (ld m)
    }
    fclose(fp);
    RSVD_CUDA(cudaStreamSynchronize(c.stream));
    dfree(ds[0]); dfree(ds[1]);
    if (g_status) { dfree(A); return 1; }
    *dA_out = A; *m_out = m; *n_out = n;
    return 0;
}

static int store_binary_dev(const char *path, int index_bits, const double *A, i64 lda, i64 m, i64 n) {
    ensure_init();
    Ctx &c = ctx();
    if (!c.inited) return 1;
    const size_t row_bytes = (size_t)n * 8;
    if (row_bytes > Staging::CHUNK) { set_error("rsvd_b200: rows of %lld doubles exceed the staging buffer", (long long)n); return 1; }
    FILE *fp = fopen(path, "wb");
    if (!fp) { set_error("rsvd_b200: cannot open %s for writing", path); return 1; }
    if (index_bits == 64) { long long h[2] = {m, n}; fwrite(h, 8, 2, fp); }
    else { int h[2] = {(int)m, (int)n}; fwrite(h, 4, 2, fp); }
    g_staging.init();
    const i64 rb = row_bytes ? (i64)(Staging::CHUNK / row_bytes) : 1;
    double *ds[2] = {dalloc((size_t)rb * n + 1), dalloc((size_t)rb * n + 1)};
    i64 pend_rows[2] = {0, 0};
    bool used[2] = {false, false};
    int b = 0;
    for (i64 i0 = 0; (i0 < m || used[0] || used[1]) && !g_status; i0 += rb, b ^= 1) {
        if (used[b]) {                                      // block issued two iterations ago has landed: write it out
            RSVD_CUDA(cudaEventSynchronize(g_staging.ev[b]));
            if (fwrite(g_staging.buf[b], 8, (size_t)pend_rows[b] * n, fp) != (size_t)pend_rows[b] * n) { set_error("rsvd_b200: short write to %s", path); break; }
            used[b] = false;
        }
        if (i0 < m) {
            const i64 rows = std::min(rb, m - i0);
            transpose(A + i0, lda, ds[b], n, rows, n);      // rows i0.. (rows x n, ld lda)^T -> (n x rows, ld n) = row-major block
            RSVD_CUDA(cudaMemcpyAsync(g_staging.buf[b], ds[b], (size_t)rows * row_bytes, cudaMemcpyDeviceToHost, c.stream));
            RSVD_CUDA(cudaEventRecord(g_staging.ev[b], c.stream));
            pend_rows[b] = rows; used[b] = true;
        }
    }
    fclose(fp);
    RSVD_CUDA(cudaStreamSynchronize(c.stream));
    dfree(ds[0]); dfree(ds[1]);
    return g_status;
}

int rsvd_b200_h2d(double *d, const double *h, rsvd_i64 n) { return copy_h2d(d, h, (size_t)n); }
int rsvd_b200_load_binary_dev(const char *path, int index_bits, double **dA, rsvd_i64 *m, rsvd_i64 *n) {
    i64 mm = 0, nn = 0;
    int rc = load_binary_dev(path, index_bits, dA, &mm, &nn);
    *m = mm; *n = nn;
    return rc;
}
int rsvd_b200_store_binary_dev(const char *path, int index_bits, const double *dA, rsvd_i64 lda, rsvd_i64 m, rsvd_i64 n) {
    return store_binary_dev(path, index_bits, dA, lda, m, n);
}
int rsvd_b200_d2h(double *h, const double *d, rsvd_i64 n) { return copy_d2h(h, d, (size_t)n); }

void *rsvd_b200_host_alloc(size_t bytes) {
    ensure_init();
    if (!ctx().inited) return nullptr;
    (void)pool_size();               // a driver's first large matrix_new is the natural place to bring the worker pool up
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 8, cudaHostAllocPortable) != cudaSuccess) {   // pinned for every device (multi-GPU uploads)
        (void)cudaGetLastError();
        return nullptr;
    }
    memset(p, 0, bytes);
    return p;
}
void rsvd_b200_host_free(void *p) { if (p) cudaFreeHost(p); }

}  // extern "C"
