// dist.cu — row-partitioned multi-GPU plumbing.  The reference is single-process, single-GPU, PCIe-copy-per-GEMM
// (SURVEY.md §5); here A is split by rows over the GPUs of one box, one process per GPU.  Y = A*Omega, Y = A*Z,
// Q = Y*R^{-1}, U = Q*Vhat need no communication; the transposed products (A^T*Y, A^T*Q, the ID's left sketch) and
// the l x l Gram matrices are summed with ncclAllReduce over NVLink.
// NCCL is resolved at run time (dlopen) so the library loads on hosts without NCCL and shares the copy a host
// framework may already have loaded.
#include "common.cuh"
#include <dlfcn.h>
#include <mutex>

namespace rsvd {

namespace {
typedef struct { char internal[128]; } NcclUniqueId;
typedef int (*GetUniqueIdFn)(NcclUniqueId *);
typedef int (*CommInitRankFn)(void **, int, NcclUniqueId, int);
typedef int (*AllReduceFn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*AllGatherFn)(const void *, void *, size_t, int, void *, cudaStream_t);
typedef int (*CommDestroyFn)(void *);
typedef const char *(*GetErrorStringFn)(int);

struct Nccl {
    void *h = nullptr;
    GetUniqueIdFn get_unique_id = nullptr;
    CommInitRankFn comm_init_rank = nullptr;
    AllReduceFn all_reduce = nullptr;
    AllGatherFn all_gather = nullptr;
    CommDestroyFn comm_destroy = nullptr;
    GetErrorStringFn error_string = nullptr;
    std::mutex mu;
    bool load() {
        std::lock_guard<std::mutex> lk(mu);
        if (h) return true;
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { set_error("rsvd_b200: cannot load libnccl.so.2: %s", dlerror()); return false; }
        get_unique_id = (GetUniqueIdFn)dlsym(h, "ncclGetUniqueId");
        comm_init_rank = (CommInitRankFn)dlsym(h, "ncclCommInitRank");
        all_reduce = (AllReduceFn)dlsym(h, "ncclAllReduce");
        all_gather = (AllGatherFn)dlsym(h, "ncclAllGather");
        comm_destroy = (CommDestroyFn)dlsym(h, "ncclCommDestroy");
        error_string = (GetErrorStringFn)dlsym(h, "ncclGetErrorString");
        if (!get_unique_id || !comm_init_rank || !all_reduce || !all_gather || !comm_destroy) {
            set_error("rsvd_b200: libnccl is missing required symbols");
            return false;
        }
        return true;
    }
};
Nccl g_nccl;
constexpr int kNcclFloat64 = 8, kNcclSum = 0;
}  // namespace

void allreduce_sum(double *d, size_t count) {
    Ctx &c = ctx();
    if (c.world <= 1 || count == 0) return;
    if (!c.nccl_comm) { set_error("rsvd_b200: world=%d but no communicator", c.world); return; }
    int r = g_nccl.all_reduce(d, d, count, kNcclFloat64, kNcclSum, c.nccl_comm, c.stream);
    if (r != 0) set_error("rsvd_b200: ncclAllReduce failed: %s", g_nccl.error_string ? g_nccl.error_string(r) : "?");
}

// recv (world * count doubles) = concatenation over ranks of each rank's `count` doubles at send.  world == 1: a copy.
void allgather(const double *send, double *recv, size_t count) {
    Ctx &c = ctx();
    if (count == 0) return;
    if (c.world <= 1) {
        if (send != recv) RSVD_CUDA(cudaMemcpyAsync(recv, send, count * 8, cudaMemcpyDeviceToDevice, c.stream));
        return;
    }
    if (!c.nccl_comm) { set_error("rsvd_b200: world=%d but no communicator", c.world); return; }
    int r = g_nccl.all_gather(send, recv, count, kNcclFloat64, c.nccl_comm, c.stream);
    if (r != 0) set_error("rsvd_b200: ncclAllGather failed: %s", g_nccl.error_string ? g_nccl.error_string(r) : "?");
}

int nccl_unique_id(char id_out[128]) {
    if (!g_nccl.load()) return 1;
    NcclUniqueId id;
    int r = g_nccl.get_unique_id(&id);
    if (r != 0) { set_error("rsvd_b200: ncclGetUniqueId failed (%d)", r); return 1; }
    memcpy(id_out, id.internal, 128);
    return 0;
}

// joins the calling thread's context (already initialised on its device) to a communicator of `world` ranks
int nccl_join(int rank, int world, const char id_in[128]) {
    Ctx &c = ctx();
    if (world <= 1) { c.rank = 0; c.world = 1; return 0; }
    if (!g_nccl.load()) return 1;
    NcclUniqueId id;
    memcpy(id.internal, id_in, 128);
    void *comm = nullptr;
    int r = g_nccl.comm_init_rank(&comm, world, id, rank);
    if (r != 0) { set_error("rsvd_b200: ncclCommInitRank failed: %s", g_nccl.error_string ? g_nccl.error_string(r) : "?"); return 1; }
    c.nccl_comm = comm; c.rank = rank; c.world = world;
    return 0;
}

void nccl_leave() {
    Ctx &c = ctx();
    if (c.nccl_comm) { cudaStreamSynchronize(c.stream); g_nccl.comm_destroy(c.nccl_comm); }
    c.nccl_comm = nullptr; c.rank = 0; c.world = 1;
}

}  // namespace rsvd

using namespace rsvd;

extern "C" {

int rsvd_b200_comm_unique_id(char id_out[128]) { return nccl_unique_id(id_out); }

int rsvd_b200_comm_init(int rank, int world, const char id_in[128]) {
    ensure_init();
    if (!ctx().inited) return 1;
    return nccl_join(rank, world, id_in);
}

void rsvd_b200_comm_destroy(void) { nccl_leave(); }

int rsvd_b200_allreduce_sum(double *d, rsvd_i64 count) { allreduce_sum(d, (size_t)count); return g_status; }

/* contiguous row blocks; block sizes are multiples of 16 rows (the TMA box height) except the last */
void rsvd_b200_row_partition(rsvd_i64 m, int world, int rank, rsvd_i64 *row0, rsvd_i64 *rows) {
    if (world < 1) world = 1;
    rsvd_i64 per = (m + world - 1) / world;
    per = (per + 15) / 16 * 16;
    rsvd_i64 r0 = per * rank;
    if (r0 > m) r0 = m;
    rsvd_i64 r1 = r0 + per;
    if (r1 > m) r1 = m;
    *row0 = r0; *rows = r1 - r0;
}

}  // extern "C"
