// Best response / exploitability as a level sweep with a max over actions in place of the
// expectation (replaces rlpoker/best_response.py:9-141), plus expected_value_exact
// (rlpoker/extensive_game.py:382-423).
//
// forward : r(h) = product of chance probabilities and opponent strategy on the path (the responder's
//           own actions contribute a factor of exactly 1);  v(z) = r(z) * u_p(z) at terminals
// backward: parents not owned by the responder sum their children in child order; for the responder's
//           information sets Q[I,a] = sum over h in I (ascending) of v(h.a), a* = first arg-max,
//           v(h) = v(h.a*).
// The reference sums the same terms grouped by information-set node lists instead of by parent, so
// values agree to a few ulp (tests hold 1e-12), not bit for bit.
#include "cfr.cuh"
#include "fused.cuh"

using rlp::fail;

namespace {
constexpr int kBlock = 256;
inline unsigned grid_for(int64_t n) { return (unsigned)((n + kBlock - 1) / kBlock); }

struct BrView {
    const int32_t *first_child, *colbase, *krank;
    const int8_t *player;
    const double *cprob, *pay1, *pay2, *sigma;
    const int64_t *koff_col;
    double *reach, *vbr, *contrib;
    int responder;
    int zero_sum;
};

__global__ void __launch_bounds__(kBlock) k_br_fwd(BrView v, int64_t lo, int64_t hi) {
    int64_t i = lo + (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= hi) return;
    int pl = v.player[i];
    if (pl < 0) return;
    int fc = v.first_child[i], n = v.first_child[i + 1] - fc;
    double r = v.reach[i];
    const double *sg = (pl > 0 && pl != v.responder) ? v.sigma + v.colbase[i] : nullptr;
    for (int k = 0; k < n; ++k) {
        int j = fc + k;
        double rj = pl == v.responder ? r : r * (pl == 0 ? v.cprob[j] : sg[k]);
        if (v.player[j] < 0) {
            double u = v.responder == 1 ? v.pay1[j] : (v.zero_sum ? -v.pay1[j] : v.pay2[j]);
            v.vbr[j] = rj * u;
        } else {
            v.reach[j] = rj;
        }
    }
}

__global__ void __launch_bounds__(kBlock) k_br_bwd(BrView v, int64_t lo, int64_t hi) {
    int64_t i = lo + (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= hi) return;
    int pl = v.player[i];
    if (pl < 0) return;
    int fc = v.first_child[i], n = v.first_child[i + 1] - fc;
    if (pl != v.responder) {
        double acc = 0.0;
        for (int k = 0; k < n; ++k) acc = acc + v.vbr[fc + k];
        v.vbr[i] = acc;
    } else {
        double *dst = v.contrib + v.koff_col[v.krank[i]] + v.colbase[i];
        for (int k = 0; k < n; ++k) dst[k] = v.vbr[fc + k];
    }
}

// responder's information sets on `level`: ordered sums, first arg-max, write the chosen child value back
__global__ void __launch_bounds__(kBlock)
k_br_select(int32_t I, int level, int responder, const int32_t *__restrict__ iset_size,
            const int32_t *__restrict__ iset_level, const int8_t *__restrict__ iset_player,
            const int64_t *__restrict__ col_off, const int64_t *__restrict__ koff_col,
            const int64_t *__restrict__ iset_node_off, const int32_t *__restrict__ iset_nodes,
            const int32_t *__restrict__ first_child, const double *__restrict__ contrib, double *vbr, int32_t *astar) {
    int s = blockIdx.x * kBlock + threadIdx.x;
    if (s >= I || iset_level[s] != level || iset_player[s] != responder) return;
    int size = iset_size[s];
    int64_t c0 = col_off[s];
    int na = (int)(col_off[s + 1] - c0);
    double best = 0.0;
    int best_a = -1;
    for (int a = 0; a < na; ++a) {
        double q = 0.0;
        for (int k = 0; k < size; ++k) q = q + contrib[koff_col[k] + c0 + a];
        if (best_a < 0 || q > best) { best = q; best_a = a; }
    }
    astar[s] = best_a;
    const int32_t *nodes = iset_nodes + iset_node_off[s];
    for (int k = 0; k < size; ++k) {
        int h = nodes[k];
        vbr[h] = vbr[first_child[h] + best_a];
    }
}

// expected_value_exact: terminal contributions reach * u, summed in breadth-first order by ONE thread
// per chunk would lose the reference's order, so this kernel only computes reach; the ordered sum is
// done by k_ev_sum.
__global__ void __launch_bounds__(kBlock) k_ev_fwd(BrView v, int64_t lo, int64_t hi) {
    int64_t i = lo + (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= hi) return;
    int pl = v.player[i];
    if (pl < 0) return;
    int fc = v.first_child[i], n = v.first_child[i + 1] - fc;
    double r = v.reach[i];
    const double *sg = pl > 0 ? v.sigma + v.colbase[i] : nullptr;
    for (int k = 0; k < n; ++k) v.reach[fc + k] = r * (pl == 0 ? v.cprob[fc + k] : sg[k]);
}

// Ordered sum over terminals (extensive_game.py:404-406 adds reach * utility in breadth-first order into one running
// double, so the association is fixed: no tree reduction can reproduce it).  One CTA: all threads stage a chunk of
// products in shared memory with coalesced loads, thread 0 then adds the chunk's terminal terms in order.  The chain
// of dependent adds is the floor (~2.1 M adds at Leduc-13); the loads are off it.
constexpr int kEvChunk = 1024;
__global__ void __launch_bounds__(kEvChunk) k_ev_sum(BrView v, int64_t N, double *out2) {
    __shared__ double t1[kEvChunk], t2[kEvChunk];
    __shared__ unsigned char term[kEvChunk];
    double e1 = 0.0, e2 = 0.0;
    for (int64_t base = 0; base < N; base += kEvChunk) {
        const int64_t i = base + threadIdx.x;
        bool is_t = false;
        if (i < N && v.player[i] < 0) {
            is_t = true;
            const double u1 = v.pay1[i], u2 = v.zero_sum ? -u1 : v.pay2[i], r = v.reach[i];
            t1[threadIdx.x] = r * u1;
            t2[threadIdx.x] = r * u2;
        }
        term[threadIdx.x] = is_t ? 1 : 0;
        __syncthreads();
        if (threadIdx.x == 0) {
            const int n = (int)(N - base < kEvChunk ? N - base : kEvChunk);
            for (int k = 0; k < n; ++k)
                if (term[k]) { e1 += t1[k]; e2 += t2[k]; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out2[0] = e1; out2[1] = e2; }
}

// Node values under a strategy profile (cfr_metrics.py:181-254 compute_expected_utility_recursive):
// v(h) = ((0 + e_0 v(c_0)) + e_1 v(c_1)) + ... in child order, e = chance probability or the actor's strategy.
template <bool ZS>
__global__ void __launch_bounds__(kBlock) k_val_level(BrView v, double *val1, double *val2, int64_t lo, int64_t hi) {
    int64_t i = lo + (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= hi) return;
    int pl = v.player[i];
    if (pl < 0) return;
    int fc = v.first_child[i], n = v.first_child[i + 1] - fc;
    const double *sg = pl > 0 ? v.sigma + v.colbase[i] : nullptr;
    double a1 = 0.0, a2 = 0.0;
    for (int k = 0; k < n; ++k) {
        double e = pl == 0 ? v.cprob[fc + k] : sg[k];
        a1 = a1 + e * val1[fc + k];
        if (!ZS) a2 = a2 + e * val2[fc + k];
    }
    val1[i] = a1;
    if (!ZS) val2[i] = a2;
}

__global__ void k_set_one(double *p) { *p = 1.0; }
__global__ void k_fill_i32(int32_t *p, int64_t n, int32_t x) {
    int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i < n) p[i] = x;
}

BrView make_view(rlp_cfr *c, const double *sigma_int, int responder) {
    rlp_tree *t = c->tree;
    BrView v;
    v.first_child = t->first_child.p; v.colbase = t->colbase.p; v.krank = t->krank.p; v.player = t->player.p;
    v.cprob = t->cprob.p; v.pay1 = t->pay1.p; v.pay2 = t->pay2.p; v.sigma = sigma_int; v.koff_col = t->koff_col.p;
    v.reach = c->pc.p; v.vbr = c->vbr.p; v.contrib = c->contrib_r.p;
    v.responder = responder; v.zero_sum = t->zero_sum ? 1 : 0;
    return v;
}
}  // namespace

namespace rlp {
int permute(rlp_tree *t, const double *src, double *dst, int to_internal, cudaStream_t s);

static int enqueue_best_response(rlp_cfr *c, const double *sigma_int, int player, cudaStream_t s);

// value of the best response of `player` against sigma_int (internal pair order) -> *value_dev.
// The ~4 L launches are captured once per (strategy buffer, responder) and replayed as a CUDA graph.
int launch_best_response(rlp_cfr *c, const double *sigma_int, int player, double *value_dev, cudaStream_t s) {
    if (rlp::fused_br_available(c)) {       // both responders in one cooperative launch on the micro / upper groups
        int rc = rlp::fused_br(c, sigma_int, c->scalars.p + 8, s);
        if (rc) return rc;
        RLP_CUDA(cudaMemcpyAsync(value_dev, c->scalars.p + 8 + (player - 1), sizeof(double), cudaMemcpyDeviceToDevice, s));
        return RLP_OK;
    }
    rlp_cfr::BrGraph *hit = nullptr;
    for (auto &b : c->br_graphs) if (b.sigma == sigma_int && b.player == player) hit = &b;
    if (!hit) {
        cudaStream_t cap;
        RLP_CUDA(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
        uint64_t before = g_launches.load();
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        cudaError_t e = cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal);
        int rc = e == cudaSuccess ? enqueue_best_response(c, sigma_int, player, cap) : RLP_ERR_CUDA;
        int kernels = (int)(g_launches.load() - before);
        if (e == cudaSuccess) e = cudaStreamEndCapture(cap, &graph);
        g_launches.store(before);
        cudaStreamDestroy(cap);
        if (rc == RLP_OK && e == cudaSuccess && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
            c->br_graphs.push_back({sigma_int, player, exec, graph, kernels});
            hit = &c->br_graphs.back();
        } else {
            cudaGetLastError();
            if (graph) cudaGraphDestroy(graph);
        }
    }
    if (hit) {
        RLP_CUDA(cudaGraphLaunch(hit->exec, s));
        g_launches.fetch_add(hit->kernels, std::memory_order_relaxed);
    } else {
        int rc = enqueue_best_response(c, sigma_int, player, s);
        if (rc) return rc;
    }
    RLP_CUDA(cudaMemcpyAsync(value_dev, c->vbr.p, sizeof(double), cudaMemcpyDeviceToDevice, s));
    return RLP_OK;
}

static int enqueue_best_response(rlp_cfr *c, const double *sigma_int, int player, cudaStream_t s) {
    rlp_tree *t = c->tree;
    BrView v = make_view(c, sigma_int, player);
    // pc[0] is 1.0 and never overwritten (rlp_cfr_create); a single terminal root is degenerate
    if (t->N == 1) return fail(RLP_ERR_UNSUPPORTED, "tree has a single node");
    for (int l = 0; l + 1 < t->L; ++l) {
        int64_t lo = t->level_off[l], hi = t->level_off[l + 1];
        k_br_fwd<<<grid_for(hi - lo), kBlock, 0, s>>>(v, lo, hi); RLP_LAUNCH_CHECK();
    }
    for (int l = t->L - 2; l >= 0; --l) {
        int64_t lo = t->level_off[l], hi = t->level_off[l + 1];
        k_br_bwd<<<grid_for(hi - lo), kBlock, 0, s>>>(v, lo, hi); RLP_LAUNCH_CHECK();
        k_br_select<<<grid_for(t->I), kBlock, 0, s>>>(t->I, l, player, t->iset_size.p, t->iset_level.p,
                                                    t->iset_player.p, t->col_off.p, t->koff_col.p,
                                                    t->iset_node_off.p, t->iset_nodes.p, t->first_child.p,
                                                    c->contrib_r.p, c->vbr.p, c->astar.p);
        RLP_LAUNCH_CHECK();
    }
    return RLP_OK;
}
}  // namespace rlp

extern "C" int rlp_cfr_exploitability(rlp_cfr *c, int which, double *out3, void *stream_) {
    if (!c || !out3) return fail(RLP_ERR_INVALID, "null argument");
    cudaStream_t s = (cudaStream_t)stream_;
    const double *sg;
    if (which == RLP_AVERAGE) {
        int rc = rlp::launch_average(c, s);
        if (rc) return rc;
        sg = c->avg.p;
    } else if (which == RLP_SIGMA) {
        sg = c->sigma.p;
    } else {
        return fail(RLP_ERR_INVALID, "exploitability of array %d", which);
    }
    int rc;
    if (rlp::fused_br_available(c)) {
        rc = rlp::fused_br(c, sg, c->scalars.p + 0, s);          // {BR1, BR2} in one launch
        if (rc) return rc;
    } else {
        rc = rlp::launch_best_response(c, sg, 1, c->scalars.p + 0, s);
        if (rc) return rc;
        rc = rlp::launch_best_response(c, sg, 2, c->scalars.p + 1, s);
        if (rc) return rc;
    }
    double h[2];
    RLP_CUDA(cudaMemcpyAsync(h, c->scalars.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    RLP_CUDA(cudaStreamSynchronize(s));
    out3[0] = h[0] + h[1];
    out3[1] = h[0];
    out3[2] = h[1];
    return RLP_OK;
}

// Stand-alone entry points use a scratch solver owned by the tree (allocated on first use, freed with the tree).
static rlp_cfr *scratch_solver(rlp_tree *t, void *stream_) {
    if (!t->scratch && rlp_cfr_create(t, stream_, &t->scratch) != RLP_OK) t->scratch = nullptr;
    return t->scratch;
}

extern "C" int rlp_best_response(rlp_tree *t, const double *sigma, int player, double *value_out,
                                 int32_t *argmax_out, void *stream_) {
    if (!t || !sigma || !value_out || (player != 1 && player != 2)) return fail(RLP_ERR_INVALID, "bad argument");
    cudaStream_t s = (cudaStream_t)stream_;
    rlp_cfr *c = scratch_solver(t, stream_);
    if (!c) return RLP_ERR_CUDA;
    RLP_CUDA(cudaMemcpyAsync(c->tmpQ.p, sigma, sizeof(double) * (size_t)t->Q, cudaMemcpyHostToDevice, s));
    int rc = rlp::permute(t, c->tmpQ.p, c->avg.p, 1, s);
    if (rc) return rc;
    if (t->I > 0) { k_fill_i32<<<grid_for(t->I), kBlock, 0, s>>>(c->astar.p, t->I, -1); RLP_LAUNCH_CHECK(); }
    rc = rlp::launch_best_response(c, c->avg.p, player, c->scalars.p, s);
    if (rc) return rc;
    RLP_CUDA(cudaMemcpyAsync(value_out, c->scalars.p, sizeof(double), cudaMemcpyDeviceToHost, s));
    std::vector<int32_t> tmp;
    if (argmax_out) {
        tmp.resize((size_t)t->I);
        RLP_CUDA(cudaMemcpyAsync(tmp.data(), c->astar.p, sizeof(int32_t) * (size_t)t->I, cudaMemcpyDeviceToHost, s));
    }
    RLP_CUDA(cudaStreamSynchronize(s));
    if (argmax_out)        // -1 for the other player's information sets (the fused kernel answers for both responders)
        for (int32_t k = 0; k < t->I; ++k) argmax_out[t->iset_of_srank[k]] = t->iset_player_h[k] == player ? tmp[k] : -1;
    return RLP_OK;
}

extern "C" int rlp_expected_value(rlp_tree *t, const double *sigma, double *out2, void *stream_) {
    if (!t || !sigma || !out2) return fail(RLP_ERR_INVALID, "bad argument");
    cudaStream_t s = (cudaStream_t)stream_;
    rlp_cfr *c = scratch_solver(t, stream_);
    if (!c) return RLP_ERR_CUDA;
    RLP_CUDA(cudaMemcpyAsync(c->tmpQ.p, sigma, sizeof(double) * (size_t)t->Q, cudaMemcpyHostToDevice, s));
    int rc = rlp::permute(t, c->tmpQ.p, c->avg.p, 1, s);
    if (rc) return rc;
    BrView v = make_view(c, c->avg.p, 0);
    v.reach = c->vbr.p;            // reach for every node incl. terminals; keep pc[] untouched
    k_set_one<<<1, 1, 0, s>>>(c->vbr.p); RLP_LAUNCH_CHECK();
    for (int l = 0; l + 1 < t->L; ++l) {
        int64_t lo = t->level_off[l], hi = t->level_off[l + 1];
        k_ev_fwd<<<grid_for(hi - lo), kBlock, 0, s>>>(v, lo, hi); RLP_LAUNCH_CHECK();
    }
    k_ev_sum<<<1, kEvChunk, 0, s>>>(v, t->N, c->scalars.p + 2); RLP_LAUNCH_CHECK();
    RLP_CUDA(cudaMemcpyAsync(out2, c->scalars.p + 2, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    RLP_CUDA(cudaStreamSynchronize(s));
    return RLP_OK;
}

// Value of EVERY node for both players under `sigma` (host, [Q], canonical; player-1 rows are read at player-1
// nodes and player-2 rows at player-2 nodes, so sigma = sigma_1 merged with sigma_2):
// cfr_metrics.py:159-254 compute_expected_utility / compute_expected_utility_recursive.  v1_out / v2_out are host
// arrays [N] in breadth-first node order (terminals hold their utilities); either may be NULL.
extern "C" int rlp_node_values(rlp_tree *t, const double *sigma, double *v1_out, double *v2_out, void *stream_) {
    if (!t || !sigma) return fail(RLP_ERR_INVALID, "bad argument");
    cudaStream_t s = (cudaStream_t)stream_;
    rlp_cfr *c = scratch_solver(t, stream_);
    if (!c) return RLP_ERR_CUDA;
    RLP_CUDA(cudaMemcpyAsync(c->tmpQ.p, sigma, sizeof(double) * (size_t)t->Q, cudaMemcpyHostToDevice, s));
    int rc = rlp::permute(t, c->tmpQ.p, c->avg.p, 1, s);
    if (rc) return rc;
    BrView v = make_view(c, c->avg.p, 0);
    // val1 / val2 hold the payoffs at terminals permanently (rlp_cfr_create); interior nodes are rewritten here
    for (int l = t->L - 2; l >= 0; --l) {
        int64_t lo = t->level_off[l], hi = t->level_off[l + 1];
        if (t->zero_sum) k_val_level<true><<<grid_for(hi - lo), kBlock, 0, s>>>(v, c->val1.p, c->val2.p, lo, hi);
        else k_val_level<false><<<grid_for(hi - lo), kBlock, 0, s>>>(v, c->val1.p, c->val2.p, lo, hi);
        RLP_LAUNCH_CHECK();
    }
    const size_t bytes = sizeof(double) * (size_t)t->N;
    if (v1_out) RLP_CUDA(cudaMemcpyAsync(v1_out, c->val1.p, bytes, cudaMemcpyDeviceToHost, s));
    if (v2_out) RLP_CUDA(cudaMemcpyAsync(v2_out, t->zero_sum ? c->val1.p : c->val2.p, bytes, cudaMemcpyDeviceToHost, s));
    RLP_CUDA(cudaStreamSynchronize(s));
    if (v2_out && t->zero_sum)                       // v2 = -v1 exactly (negation commutes with rounding)
        for (int64_t i = 0; i < t->N; ++i) v2_out[i] = -v2_out[i];
    return RLP_OK;
}
