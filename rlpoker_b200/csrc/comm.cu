// Multi-GPU iteration loop with the NCCL all-reduce issued from C (no Python, no torch in the loop).
//
// One process per GPU; every rank sweeps its own deal subtrees into the [2Q] (regret || strategy-sum) delta buffer,
// ncclAllReduce(sum) runs in place on the caller's stream, and every rank applies the summed delta and recomputes
// the strategy (replicated).  The host cost per iteration is a handful of launches (~25 us), so the loop is bound by
// the GPUs rather than by the interpreter (a Python loop around the same three steps measured ~88 us / iteration).
// libnccl.so.2 is dlopen'ed: torch.distributed (which the caller uses to share the unique id) has already loaded it.
#include <dlfcn.h>

#include "cfr.cuh"

using rlp::fail;

namespace {
typedef struct { char internal[128]; } NcclUniqueId;
typedef void *NcclComm;
struct Nccl {
    void *handle = nullptr;
    bool tried = false;
    int (*get_unique_id)(NcclUniqueId *) = nullptr;
    int (*comm_init_rank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
    int (*all_reduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*comm_destroy)(NcclComm) = nullptr;
    const char *(*error_string)(int) = nullptr;
    bool load() {
        if (tried) return handle != nullptr;
        tried = true;
        handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!handle) return false;
#define RLP_SYM(field, sym) field = (decltype(field))dlsym(handle, #sym); if (!field) { handle = nullptr; return false; }
        RLP_SYM(get_unique_id, ncclGetUniqueId)
        RLP_SYM(comm_init_rank, ncclCommInitRank)
        RLP_SYM(all_reduce, ncclAllReduce)
        RLP_SYM(comm_destroy, ncclCommDestroy)
        RLP_SYM(error_string, ncclGetErrorString)
#undef RLP_SYM
        return true;
    }
};
Nccl g_nccl;
constexpr int kNcclDouble = 8, kNcclSum = 0;       // nccl.h: ncclFloat64 = 8, ncclSum = 0
}  // namespace

extern "C" int rlp_comm_unique_id(unsigned char *out128) {
    if (!out128) return fail(RLP_ERR_INVALID, "null argument");
    if (!g_nccl.load()) return fail(RLP_ERR_UNSUPPORTED, "libnccl.so.2 not available: %s", dlerror());
    NcclUniqueId id;
    int r = g_nccl.get_unique_id(&id);
    if (r != 0) return fail(RLP_ERR_CUDA, "ncclGetUniqueId: %s", g_nccl.error_string(r));
    memcpy(out128, id.internal, 128);
    return RLP_OK;
}

extern "C" int rlp_comm_init(rlp_cfr *c, int rank, int nranks, const unsigned char *id128) {
    if (!c || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(RLP_ERR_INVALID, "bad argument");
    if (!g_nccl.load()) return fail(RLP_ERR_UNSUPPORTED, "libnccl.so.2 not available");
    if (c->nccl_comm) { g_nccl.comm_destroy((NcclComm)c->nccl_comm); c->nccl_comm = nullptr; }
    NcclUniqueId id;
    memcpy(id.internal, id128, 128);
    NcclComm comm = nullptr;
    int r = g_nccl.comm_init_rank(&comm, nranks, id, rank);
    if (r != 0) return fail(RLP_ERR_CUDA, "ncclCommInitRank: %s", g_nccl.error_string(r));
    c->nccl_comm = comm;
    c->nccl_ranks = nranks;
    return RLP_OK;
}

extern "C" void rlp_comm_destroy(rlp_cfr *c) {
    if (c && c->nccl_comm && g_nccl.handle) g_nccl.comm_destroy((NcclComm)c->nccl_comm);
    if (c) c->nccl_comm = nullptr;
}

namespace {
constexpr int kShardGraphIters = 10;

int one_sharded_iteration(rlp_cfr *c, int64_t t, int linear_weight, cudaStream_t s) {
    const size_t count = 2 * (size_t)c->tree->Q;
    int rc = rlp_cfr_local_sweep(c, t, linear_weight, (void *)s);
    if (rc) return rc;
    if (c->nccl_ranks > 1) {
        int r = g_nccl.all_reduce(c->delta.p, c->delta.p, count, kNcclDouble, kNcclSum, (NcclComm)c->nccl_comm, s);
        if (r != 0) return fail(RLP_ERR_CUDA, "ncclAllReduce: %s", g_nccl.error_string(r));
    }
    return rlp_cfr_apply_delta(c, (void *)s);
}

// Capture kShardGraphIters iterations (NCCL collectives are capturable) so that the loop is not paced by host
// launches.  Every rank captures the same sequence.  Opt-in (RLP_SHARD_CGRAPH=1) until validated on 8 GPUs.
int ensure_shard_graph(rlp_cfr *c) {
    if (c->shard_exec) return RLP_OK;
    cudaStream_t cap;
    RLP_CUDA(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
    uint64_t before = rlp::g_launches.load();
    uint64_t touched = c->nodes_touched;
    cudaError_t e = cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal);
    int rc = e == cudaSuccess ? RLP_OK : RLP_ERR_CUDA;
    for (int u = 0; u < kShardGraphIters && rc == RLP_OK; ++u) rc = one_sharded_iteration(c, -1, 0, cap);
    int kernels = (int)(rlp::g_launches.load() - before);
    if (e == cudaSuccess) e = cudaStreamEndCapture(cap, &c->shard_graph);
    rlp::g_launches.store(before);
    c->nodes_touched = touched;                      // capture records, it does not run
    cudaStreamDestroy(cap);
    if (rc) return rc;
    if (e != cudaSuccess) return fail(RLP_ERR_CUDA, "capture of the sharded loop: %s", cudaGetErrorString(e));
    RLP_CUDA(cudaGraphInstantiate(&c->shard_exec, c->shard_graph, 0));
    c->shard_graph_kernels = kernels;
    return RLP_OK;
}
}  // namespace

extern "C" int rlp_cfr_sharded_iters(rlp_cfr *c, int64_t t0, int64_t n_iters, int linear_weight, void *stream_) {
    if (!c || n_iters < 0) return fail(RLP_ERR_INVALID, "bad argument");
    if (!c->nccl_comm) return fail(RLP_ERR_INVALID, "rlp_comm_init has not been called on this solver");
    cudaStream_t s = (cudaStream_t)stream_;
    const char *env = getenv("RLP_SHARD_CGRAPH");
    const bool use_graph = env && env[0] == '1';
    int64_t done = 0;
    const bool in_sync = c->dev_t == (long long)t0 && c->dev_linear == (linear_weight ? 1 : 0);
    if (n_iters > 0 && !in_sync) {                    // an eager iteration sets counter + weighting flag on the device
        int rc = one_sharded_iteration(c, t0, linear_weight, s);
        if (rc) return rc;
        done = 1;
    }
    if (use_graph && n_iters - done >= kShardGraphIters) {
        int rc = ensure_shard_graph(c);
        if (rc) return rc;
        for (; n_iters - done >= kShardGraphIters; done += kShardGraphIters) {
            RLP_CUDA(cudaGraphLaunch(c->shard_exec, s));
            rlp::g_launches.fetch_add(c->shard_graph_kernels, std::memory_order_relaxed);
            c->nodes_touched += 2ull * (uint64_t)c->tree->N * kShardGraphIters;
        }
    }
    for (; done < n_iters; ++done) {
        int rc = one_sharded_iteration(c, -1, linear_weight, s);
        if (rc) return rc;
    }
    c->dev_t = (long long)(t0 + n_iters);
    c->dev_linear = linear_weight ? 1 : 0;
    return RLP_OK;
}
