// dist.cu — row-partitioned multi-GPU plumbing.  The reference is single-process, single-GPU, PCIe-copy-per-GEMM
// (SURVEY.md §5); here A is split by rows over the GPUs of one box, one process per GPU.  Y = A*Omega, Y = A*Z,
// Q = Y*R^{-1}, U = Q*Vhat need no communication; the transposed products (A^T*Y, A^T*Q, the ID's left sketch) and
// the l x l Gram matrices are summed with ncclAllReduce over NVLink.
// NCCL is resolved at run time (dlopen) so the library loads on hosts without NCCL and shares the copy a host
// framework may already have loaded.
#include "common.cuh"
#include <dlfcn.h>
#include <mutex>
#include <unistd.h>
#include <vector>

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


// ---- peer mailboxes ------------------------------------------------------------------------------------------------------
// The sharded pivoted QR exchanges one small record per rank and step (a candidate column: <= 33 KB).  Through ncclAllGather
// that costs two extra kernel launches and ~25 us of collective latency per step — more than the step's arithmetic once the
// local columns fit in L2.  Instead every rank exposes a mailbox in its own HBM to all peers (cudaIpcOpenMemHandle between
// processes, plain peer access between the worker threads of one process); a step's kernel stores its record straight into
// the peers' mailboxes over NVLink, publishes a token with a system-scope release store and waits for the peers' tokens.
// Set-up is collective and all-or-nothing: the ranks all-reduce a failure count, so either every rank uses the mailboxes or
// every rank keeps the NCCL path.
namespace {
struct PeerRecord {            // 128 bytes = 16 doubles, all-gathered
    long long pid;
    long long dev;
    unsigned long long ptr;
    cudaIpcMemHandle_t handle; // 64 bytes
    char pad[128 - 24 - sizeof(cudaIpcMemHandle_t)];
};
static_assert(sizeof(PeerRecord) == 128, "PeerRecord is all-gathered as 16 doubles");
constexpr size_t kPeerFlagBytes = 1024;     // 2 x world u64 (world <= 64)
}  // namespace

bool peer_mailbox(size_t rec_doubles) {
    Ctx &c = ctx();
    Ctx::PeerBox &pb = c.peer;
    if (c.world <= 1 || c.world > 64) return false;
    if (pb.state != 0) return pb.state == 1 && rec_doubles <= pb.rec_max;
    pb.state = -1;
    const char *off = getenv("RSVD_B200_NO_PEER_MAILBOX");
    int failed = (off && atoi(off)) ? 1 : 0;
    const size_t rec_max = 4 + 4096;                       // CAND_HDR + the tallest panel of the blocked kernel
    const size_t bytes = kPeerFlagBytes + (size_t)2 * c.world * rec_max * sizeof(double);
    void *block = nullptr;
    PeerRecord mine;
    memset(&mine, 0, sizeof(mine));
    if (!failed) {
        if (cudaMalloc(&block, bytes) != cudaSuccess) { (void)cudaGetLastError(); failed = 1; block = nullptr; }
        else {
            RSVD_CUDA(cudaMemset(block, 0, bytes));
            mine.pid = (long long)getpid(); mine.dev = c.device; mine.ptr = (unsigned long long)(uintptr_t)block;
            if (cudaIpcGetMemHandle(&mine.handle, block) != cudaSuccess) { (void)cudaGetLastError(); failed = 1; }
        }
    }
    // all-gather the records (every rank takes part, also one that already failed: the set-up must stay collective)
    DBuf send(16), recv((size_t)16 * c.world), fsum(1);
    if (g_status) return false;
    std::vector<PeerRecord> all((size_t)c.world);
    RSVD_CUDA(cudaMemcpyAsync(send.p, &mine, sizeof(mine), cudaMemcpyHostToDevice, c.stream));
    allgather(send.p, recv.p, 16);
    RSVD_CUDA(cudaMemcpyAsync(all.data(), recv.p, sizeof(PeerRecord) * (size_t)c.world, cudaMemcpyDeviceToHost, c.stream));
    RSVD_CUDA(cudaStreamSynchronize(c.stream));
    std::vector<void *> base((size_t)c.world, nullptr);
    for (int g = 0; g < c.world && !failed; ++g) {
        if (g == c.rank) { base[g] = block; continue; }
        if (all[g].ptr == 0) { failed = 1; break; }
        if (all[g].pid == mine.pid) {                      // a worker thread of this process: peer access, same address
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, c.device, (int)all[g].dev) != cudaSuccess || !can) { (void)cudaGetLastError(); failed = 1; break; }
            cudaError_t e = cudaDeviceEnablePeerAccess((int)all[g].dev, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { failed = 1; }
            (void)cudaGetLastError();
            base[g] = (void *)(uintptr_t)all[g].ptr;
        } else {
            void *ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, all[g].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { (void)cudaGetLastError(); failed = 1; break; }
            base[g] = ptr;
        }
    }
    // all-or-nothing
    double f = failed ? 1.0 : 0.0;
    RSVD_CUDA(cudaMemcpyAsync(fsum.p, &f, 8, cudaMemcpyHostToDevice, c.stream));
    allreduce_sum(fsum.p, 1);
    RSVD_CUDA(cudaMemcpyAsync(&f, fsum.p, 8, cudaMemcpyDeviceToHost, c.stream));
    RSVD_CUDA(cudaStreamSynchronize(c.stream));
    if (g_status || f != 0.0) {
        if (c.verbose) fprintf(stderr, "[rsvd_b200 r%d] peer mailboxes unavailable (%d rank(s) could not map them): the sharded QR keeps ncclAllGather\n", c.rank, (int)f);
        return false;                                      // mapped blocks are simply left alone (a handful of KB per peer)
    }
    std::vector<unsigned long long *> pf((size_t)c.world);
    std::vector<double *> pc((size_t)c.world);
    for (int g = 0; g < c.world; ++g) {
        pf[g] = (unsigned long long *)base[g];
        pc[g] = (double *)((char *)base[g] + kPeerFlagBytes);
    }
    void *dpf = nullptr, *dpc = nullptr, *derr = nullptr;
    if (cudaMalloc(&dpf, sizeof(void *) * c.world) != cudaSuccess || cudaMalloc(&dpc, sizeof(void *) * c.world) != cudaSuccess ||
        cudaMalloc(&derr, sizeof(int)) != cudaSuccess) { (void)cudaGetLastError(); return false; }   // (cannot happen after the block itself was allocated)
    RSVD_CUDA(cudaMemcpy(dpf, pf.data(), sizeof(void *) * c.world, cudaMemcpyHostToDevice));
    RSVD_CUDA(cudaMemcpy(dpc, pc.data(), sizeof(void *) * c.world, cudaMemcpyHostToDevice));
    RSVD_CUDA(cudaMemset(derr, 0, sizeof(int)));
    pb.block = block; pb.flag = pf[c.rank]; pb.cand = pc[c.rank];
    pb.d_peer_flag = (unsigned long long **)dpf; pb.d_peer_cand = (double **)dpc;
    pb.rec_max = rec_max; pb.token = 0; pb.err = (int *)derr;
    pb.state = g_status ? -1 : 1;
    if (c.verbose) fprintf(stderr, "[rsvd_b200 r%d] peer mailboxes mapped on %d ranks\n", c.rank, c.world);
    return pb.state == 1 && rec_doubles <= pb.rec_max;
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
    c.peer.state = 0;                     // mailboxes belong to a communicator: a new one sets them up again (old blocks stay mapped)
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
