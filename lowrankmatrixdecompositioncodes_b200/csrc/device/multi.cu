// multi.cu — single-process multi-GPU: the reference's caller is ONE single-threaded C process
// (multi_core_mkl_code_64bit/driver1.c:40-50, multi_core_mkl_code/driver_multi_core_mkl1.c:41), so a relinked driver
// reaches all GPUs of the box only if the library fans out by itself.  RSVD_B200_DEVICES=0-7 (or rsvd_b200_set_devices)
// starts one worker thread per device, each with its own context (stream, memory pool, flags) and its rank of one NCCL
// communicator; the host-level entry points (hostapi.cu) row-partition the host matrix over the workers.
// Without it (the default) pool_run() executes inline on the caller's context: one GPU, or one rank of a
// process-per-GPU job whose communicator was set up with rsvd_b200_comm_init.
#include "common.cuh"
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

namespace rsvd {

namespace {

struct Worker {
    Ctx ctx;
    int device = 0, rank = 0;
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    const std::function<void(int)> *job = nullptr;
    bool done = true, quit = false;
};

struct Pool {
    std::vector<Worker *> w;
    std::mutex run_mu;          // one fan-out at a time (the API is single-caller)
};
Pool *g_pool = nullptr;
bool g_env_checked = false;
std::mutex g_pool_mu;

void worker_main(Worker *w, int world, const char *nccl_id, int *init_rc) {
    bind_ctx(&w->ctx);
    int rc = init_ctx(w->ctx, w->device);
    bind_ctx(&w->ctx);
    if (!rc) rc = nccl_join(w->rank, world, nccl_id);
    {
        std::lock_guard<std::mutex> lk(w->mu);
        *init_rc = rc;
        w->done = true;
    }
    w->cv.notify_all();
    for (;;) {
        std::unique_lock<std::mutex> lk(w->mu);
        w->cv.wait(lk, [&] { return w->job != nullptr || w->quit; });
        if (w->quit) break;
        const std::function<void(int)> *job = w->job;
        lk.unlock();
        cudaSetDevice(w->device);
        (*job)(w->rank);
        lk.lock();
        w->job = nullptr;
        w->done = true;
        lk.unlock();
        w->cv.notify_all();
    }
    nccl_leave();
}

void pool_destroy() {
    if (!g_pool) return;
    for (Worker *w : g_pool->w) {
        { std::lock_guard<std::mutex> lk(w->mu); w->quit = true; }
        w->cv.notify_all();
        if (w->th.joinable()) w->th.join();
        delete w;
    }
    delete g_pool;
    g_pool = nullptr;
}

int pool_create(const std::vector<int> &ids) {
    pool_destroy();
    const int n = (int)ids.size();
    if (n <= 1) return 0;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { (void)cudaGetLastError(); set_error("rsvd_b200: no CUDA device available; there is no CPU fallback"); return 1; }
    for (int id : ids)
        if (id < 0 || id >= ndev) { set_error("rsvd_b200: device %d requested but the box has %d", id, ndev); return 1; }
    char nccl_id[128];
    if (nccl_unique_id(nccl_id)) return 1;
    Pool *p = new Pool();
    std::vector<int> rcs((size_t)n, -1);
    for (int r = 0; r < n; ++r) {
        Worker *w = new Worker();
        w->device = ids[r]; w->rank = r; w->done = false;
        p->w.push_back(w);
    }
    for (int r = 0; r < n; ++r) p->w[r]->th = std::thread(worker_main, p->w[r], n, nccl_id, &rcs[r]);   // ncclCommInitRank blocks until all ranks joined
    int bad = 0;
    for (int r = 0; r < n; ++r) {
        Worker *w = p->w[r];
        std::unique_lock<std::mutex> lk(w->mu);
        w->cv.wait(lk, [&] { return w->done; });
        bad |= rcs[r];
    }
    g_pool = p;
    if (bad) { pool_destroy(); return 1; }
    if (getenv("RSVD_B200_VERBOSE")) fprintf(stderr, "[rsvd_b200] single-process multi-GPU: %d workers\n", n);
    return 0;
}

// "0-7", "0,2,4", "all", "3"
std::vector<int> parse_devices(const char *s) {
    std::vector<int> ids;
    if (!s || !*s) return ids;
    if (!strcmp(s, "all")) {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess) { (void)cudaGetLastError(); ndev = 0; }
        for (int i = 0; i < ndev; ++i) ids.push_back(i);
        return ids;
    }
    const char *p = s;
    while (*p) {
        char *e = nullptr;
        long a = strtol(p, &e, 10);
        if (e == p) break;
        long b = a;
        p = e;
        if (*p == '-') { b = strtol(p + 1, &e, 10); p = e; }
        for (long i = a; i <= b; ++i) ids.push_back((int)i);
        if (*p == ',') ++p;
    }
    return ids;
}

void check_env() {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (g_env_checked) return;
    g_env_checked = true;
    const char *s = getenv("RSVD_B200_DEVICES");
    if (!s) return;
    std::vector<int> ids = parse_devices(s);
    if (ids.size() > 1) pool_create(ids);
}

}  // namespace

int pool_size() {
    if (!g_env_checked) check_env();
    return g_pool ? (int)g_pool->w.size() : 1;
}

void pool_run(const std::function<void(int)> &fn) {
    if (pool_size() <= 1) {
        ensure_init();
        fn(ctx().rank);
        return;
    }
    std::lock_guard<std::mutex> run_lk(g_pool->run_mu);
    for (Worker *w : g_pool->w) {
        { std::lock_guard<std::mutex> lk(w->mu); w->job = &fn; w->done = false; }
        w->cv.notify_all();
    }
    for (Worker *w : g_pool->w) {
        std::unique_lock<std::mutex> lk(w->mu);
        w->cv.wait(lk, [&] { return w->done; });
    }
    ensure_init();   // back on the caller's device
}

}  // namespace rsvd

using namespace rsvd;

extern "C" {

int rsvd_b200_set_devices(int n, const int *ids) {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    g_env_checked = true;
    std::vector<int> v;
    for (int i = 0; i < n; ++i) v.push_back(ids ? ids[i] : i);
    if (n <= 1) { pool_destroy(); return 0; }
    return pool_create(v);
}

int rsvd_b200_active_devices(void) { return pool_size(); }

}  // extern "C"
