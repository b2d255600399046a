// Solver state shared by the CFR, best-response and sampled-traversal translation units.
#pragma once
#include "tree.cuh"

struct IterState {
    long long t;        // iteration about to run
    int linear;         // weight = t if linear else 1.0 (cfr.py:106)
    int pad;
    unsigned long long touched;   // device-side nodes_touched (sampled variants)
    unsigned long long *tl;       // optional timeline log (debug): records of {kind, start ns, end ns}
    unsigned int tl_count, tl_cap;
};

// Debug timeline: block 0 of a kernel appends {kind, globaltimer at start, at end}.  No-op when it->tl is null.
__device__ __forceinline__ unsigned long long rlp_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ int rlp_tl_begin(const IterState *it, int kind) {
    if (!it->tl || blockIdx.x != 0 || threadIdx.x != 0) return -1;
    IterState *m = const_cast<IterState *>(it);
    unsigned int slot = atomicAdd(&m->tl_count, 1u);
    if (slot >= m->tl_cap) return -1;
    m->tl[3 * slot] = (unsigned long long)kind;
    m->tl[3 * slot + 1] = rlp_now();
    return (int)slot;
}
__device__ __forceinline__ void rlp_tl_end(const IterState *it, int slot) {
    if (slot >= 0) const_cast<IterState *>(it)->tl[3 * slot + 2] = rlp_now();
}

struct rlp_cfr {
    rlp_tree *tree = nullptr;
    // [Q] internal pair order
    rlp::DevBuf<double> R, S, sigma, avg, cum_imm, tmpQ, delta;
    // action-major copy of sigma, [action][internal information set]: consecutive micro subtrees read consecutive
    // words (the [set][action] layout costs 3x the L2 sectors in the micro kernel)
    rlp::DevBuf<double> sigma_am;
    bool sigma_am_dirty = true;
    // group-fused persistent kernel (fused.cu): strategy in the fused numbering, [2 buffers][action][fused id]
    rlp::DevBuf<double> br_buf;              // scratch of the group-fused best-response kernel
    rlp::DevBuf<double> sigF;
    int sigF_cur = 0;                        // buffer holding the current strategy
    bool sigF_dirty = true;                  // c->sigma was changed by something other than the fused kernel

    rlp::DevBuf<uint8_t> flags;              // [3][I] internal order: in_R | in_S (or alpha) | in_sigma dict membership
    // per-node scratch
    rlp::DevBuf<double> pc, p1, p2, val1, val2, vbr;
    rlp::DevBuf<double> contrib_r, contrib_o;
    rlp::DevBuf<int32_t> astar;              // [I] best-response action per information set
    rlp::DevBuf<IterState> iter;
    // frontier records of the level-synchronous sampled lanes (sampled.cu), sized for `fr_lanes` lanes of `fr_variant`
    rlp::DevBuf<int32_t> fr_node, fr_trav, fr_child;
    rlp::DevBuf<double> fr_val, fr_reach;
    rlp::DevBuf<unsigned int> fr_count;
    rlp::DevBuf<double> lane_acc;            // zeroed accumulator copies of the batched lanes
    int64_t fr_lanes = 0;
    int fr_variant = -1;
    int fr_levels = 0;
    rlp::DevBuf<unsigned long long> timeline;
    rlp::DevBuf<double> scalars;             // small device scratch for reductions
    uint64_t nodes_touched = 0;              // host-side count (vanilla); sampled lanes count in iter->touched
    bool touched_dirty = false;
    bool external_state = false;             // S holds alpha of AverageStrategy
    // CUDA graph of one vanilla iteration
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    int graph_kernels = 0;
    // side stream + events for the two-branch iteration pipeline (accumulate of the micro-only information sets
    // overlaps the upper-tile kernels)
    // best-response sweeps captured as CUDA graphs, keyed by (strategy buffer, responder)
    struct BrGraph { const double *sigma; int player; cudaGraphExec_t exec; cudaGraph_t graph; int kernels; };
    std::vector<BrGraph> br_graphs;
    long long dev_t = -1;                       // iteration counter / weighting flag known to be on the device
    int dev_linear = -1;                        // (-1: unknown); lets rlp_cfr_sharded_iters skip its eager sync step
    cudaGraph_t shard_graph = nullptr;          // 10 iterations of sweep -> ncclAllReduce -> apply (comm.cu)
    cudaGraphExec_t shard_exec = nullptr;
    int shard_graph_kernels = 0;
    void *p2p = nullptr;                     // rlp::P2P of the solve sharded by public state (fused.cu)
    void *nccl_comm = nullptr;               // ncclComm_t of the deal-sharded solve (comm.cu)
    int nccl_ranks = 1;
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    ~rlp_cfr() {
        if (graph_exec) cudaGraphExecDestroy(graph_exec);
        if (graph) cudaGraphDestroy(graph);
        if (shard_exec) cudaGraphExecDestroy(shard_exec);
        if (shard_graph) cudaGraphDestroy(shard_graph);
        for (auto &b : br_graphs) { if (b.exec) cudaGraphExecDestroy(b.exec); if (b.graph) cudaGraphDestroy(b.graph); }
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
        if (side) cudaStreamDestroy(side);
    }
};

namespace rlp {
// implemented in cfr.cu, used by br.cu / sampled.cu
int launch_average(rlp_cfr *c, cudaStream_t s);                        // c->avg <- average strategy
int launch_best_response(rlp_cfr *c, const double *sigma_int, int player, double *value_dev, cudaStream_t s);
}  // namespace rlp
