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

// Sequential (one thread) ordered sum over terminals: exact reproduction of the reference's
// accumulation order.  Only used by the expected_value API (not on the solver's hot path).
__global__ void k_ev_sum(BrView v, int64_t N, double *out2) {
    double e1 = 0.0, e2 = 0.0;
    for (int64_t i = 0; i < N; ++i) {
        if (v.player[i] >= 0) continue;
        double u1 = v.pay1[i], u2 = v.zero_sum ? -u1 : v.pay2[i];
        e1 += v.reach[i] * u1;
        e2 += v.reach[i] * u2;
    }
    out2[0] = e1;
    out2[1] = e2;
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
    int rc = rlp::launch_best_response(c, sg, 1, c->scalars.p + 0, s);
    if (rc) return rc;
    rc = rlp::launch_best_response(c, sg, 2, c->scalars.p + 1, s);
    if (rc) return rc;
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
    if (argmax_out)
        for (int32_t k = 0; k < t->I; ++k) argmax_out[t->iset_of_srank[k]] = tmp[k];
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
    k_ev_sum<<<1, 1, 0, s>>>(v, t->N, c->scalars.p + 2); RLP_LAUNCH_CHECK();
    RLP_CUDA(cudaMemcpyAsync(out2, c->scalars.p + 2, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    RLP_CUDA(cudaStreamSynchronize(s));
    return RLP_OK;
}
