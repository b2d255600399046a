// Vanilla CFR as a depth-ordered sweep (replaces rlpoker/cfr/cfr.py:105-122,158-255 and
// rlpoker/cfr/cfr_util.py:14-58 of the reference).
//
// One iteration = forward reach propagation level by level, backward counterfactual values level by
// level with the regret / own-reach contributions of every decision node scattered into a
// jagged-diagonal scratch, then one accumulate kernel that adds each information set's contributions
// in ascending node order (the order the reference's recursion visits them), and regret matching.
// Every product and sum is written in the reference's operand order and the library is compiled with
// -fmad=false, so regrets, action counts and strategies are bit-identical to the reference.
#include <memory>

#include "cfr.cuh"
#include "fused.cuh"

using rlp::fail;

namespace rlp {
int launch_tile_sweep(rlp_cfr *c, cudaStream_t s, const IterState *iter, int *launches, cudaEvent_t *marks);
int launch_tile_stage(rlp_cfr *c, cudaStream_t s, const IterState *iter, int stage, int *launches);
}

namespace {

constexpr int kBlock = 256;
constexpr int kTopBlock = 1024;

struct View {
    const int32_t *first_child, *colbase, *srank, *krank;
    const int8_t *player;
    const double *cprob, *sigma;
    const int64_t *koff_col, *koff_iset;
    double *pc, *p1, *p2, *val1, *val2, *contrib_r, *contrib_o;
    const IterState *iter;
    long long ND;       // contrib_r is action-major: [action][ND], indexed like contrib_o
};

// ---- forward: reach of the children of node i (cfr.py:180,208,217) -----------------------------
__device__ __forceinline__ void fwd_node(const View &v, int64_t i) {
    int pl = v.player[i];
    if (pl < 0) return;
    int fc = v.first_child[i], n = v.first_child[i + 1] - fc;
    double c = v.pc[i], a = v.p1[i], b = v.p2[i];
    const double *sg = pl > 0 ? v.sigma + v.colbase[i] : nullptr;
    for (int k = 0; k < n; ++k) {
        int j = fc + k;
        if (v.player[j] < 0) continue;                 // terminals need no reach
        double e = pl == 0 ? v.cprob[j] : sg[k];
        v.pc[j] = pl == 0 ? e * c : c;
        v.p1[j] = pl == 1 ? e * a : a;
        v.p2[j] = pl == 2 ? e * b : b;
    }
}

// ---- backward: value of node i from its children, and its contributions (cfr.py:177-186,204-242) --
template <bool ZS>
__device__ __forceinline__ void bwd_node(const View &v, int64_t i, double w) {
    int pl = v.player[i];
    if (pl < 0) return;
    int fc = v.first_child[i], n = v.first_child[i + 1] - fc;
    const double *sg = pl > 0 ? v.sigma + v.colbase[i] : nullptr;
    double a1 = 0.0, a2 = 0.0;
    for (int k = 0; k < n; ++k) {
        double e = pl == 0 ? v.cprob[fc + k] : sg[k];
        a1 = a1 + e * v.val1[fc + k];
        if (!ZS) a2 = a2 + e * v.val2[fc + k];
    }
    v.val1[i] = a1;
    if (!ZS) v.val2[i] = a2;
    if (pl == 0) return;
    double c = v.pc[i], x1 = v.p1[i], x2 = v.p2[i];
    double opp = pl == 1 ? c * x2 : c * x1;            // pi_c * pi_{-i}
    double own = pl == 1 ? x1 : x2;
    double vp = pl == 1 ? a1 : (ZS ? -a1 : a2);
    int kk = v.krank[i];
    const long long slot = v.koff_iset[kk] + v.srank[i];
    for (int k = 0; k < n; ++k) {
        double va = pl == 1 ? v.val1[fc + k] : (ZS ? -v.val1[fc + k] : v.val2[fc + k]);
        v.contrib_r[(long long)k * v.ND + slot] = w * (va - vp) * opp;      // weight * (v_a - v) * pi_minus_i
    }
    v.contrib_o[slot] = w * own;                      // (weight * pi_i), times sigma in accumulate
}

__device__ __forceinline__ double iter_weight(const IterState *it) { return it->linear ? (double)it->t : 1.0; }

__global__ void __launch_bounds__(kBlock) k_fwd_level(View v, int64_t lo, int64_t hi) {
    int64_t i = lo + (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i < hi) fwd_node(v, i);
}

template <bool ZS>
__global__ void __launch_bounds__(kBlock) k_bwd_level(View v, int64_t lo, int64_t hi) {
    int64_t i = lo + (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i < hi) bwd_node<ZS>(v, i, iter_weight(v.iter));
}

// The shallow top of the tree (a few thousand nodes over several levels) in ONE block.
struct TopLevels { long long off[18]; int n; };

__global__ void __launch_bounds__(kTopBlock) k_fwd_top(View v, TopLevels tl) {
    for (int l = 0; l < tl.n; ++l) {
        for (int64_t i = tl.off[l] + threadIdx.x; i < tl.off[l + 1]; i += kTopBlock) fwd_node(v, i);
        __syncthreads();
    }
}

template <bool ZS>
__global__ void __launch_bounds__(kTopBlock) k_bwd_top(View v, TopLevels tl) {
    double w = iter_weight(v.iter);
    for (int l = tl.n - 1; l >= 0; --l) {
        for (int64_t i = tl.off[l] + threadIdx.x; i < tl.off[l + 1]; i += kTopBlock) bwd_node<ZS>(v, i, w);
        __syncthreads();
    }
}

// ---- accumulate + regret matching (cfr.py:244-252, cfr_util.py:14-58) -------------------------------
// One thread per information set (internal, size-sorted order => neighbouring threads stream
// neighbouring columns of every jagged diagonal).
__global__ void __launch_bounds__(kBlock)
k_accumulate_rm(int32_t I, const int32_t *__restrict__ iset_size, const int64_t *__restrict__ col_off,
                const int64_t *__restrict__ koff_col, const int64_t *__restrict__ koff_iset,
                const double *__restrict__ contrib_r, const double *__restrict__ contrib_o, double *R, double *S,
                double *sigma, uint8_t *flags, IterState *iter, int advance, int64_t ND, double *sigma_am) {
    int s = blockIdx.x * kBlock + threadIdx.x;
    if (s == 0 && advance) iter->t += 1;       // every reader of iter->t ran in an earlier kernel
    if (s >= I) return;
    int size = iset_size[s];
    int64_t c0 = col_off[s];
    int na = (int)(col_off[s + 1] - c0);
    if (na <= 4) {
        double r[4], sa[4], sg[4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
            if (a < na) { r[a] = R[c0 + a]; sa[a] = S[c0 + a]; sg[a] = sigma[c0 + a]; }
        for (int k = 0; k < size; ++k) {
            const int64_t slot = koff_iset[k] + s;
            double wo = contrib_o[slot];
#pragma unroll
            for (int a = 0; a < 4; ++a)
                if (a < na) { r[a] = r[a] + contrib_r[(int64_t)a * ND + slot]; sa[a] = sa[a] + wo * sg[a]; }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
            if (a < na) { R[c0 + a] = r[a]; S[c0 + a] = sa[a]; }
    } else {
        for (int a = 0; a < na; ++a) {
            double r = R[c0 + a], sa = S[c0 + a], sg = sigma[c0 + a];
            for (int k = 0; k < size; ++k) {
                r = r + contrib_r[(int64_t)a * ND + koff_iset[k] + s];
                sa = sa + contrib_o[koff_iset[k] + s] * sg;
            }
            R[c0 + a] = r;
            S[c0 + a] = sa;
        }
    }
    rlp::regret_matching(R + c0, na, sigma + c0);
    for (int a = 0; a < na; ++a) sigma_am[(int64_t)a * I + s] = sigma[c0 + a];
    flags[s] = 1; flags[I + s] = 1; flags[2 * I + s] = 1;
}

// Same, four lanes per information set (one per action; information sets of <= 4 actions): 4x the loads in
// flight, and the regret-matching normaliser is formed from the four lanes' regrets by shuffles.
__global__ void __launch_bounds__(kBlock)
k_accumulate_rm4(int32_t I, const int32_t *__restrict__ iset_size, const int64_t *__restrict__ col_off,
                 const int64_t *__restrict__ koff_col, const int64_t *__restrict__ koff_iset,
                 const double *__restrict__ contrib_r, const double *__restrict__ contrib_o, double *R, double *S,
                 double *sigma, uint8_t *flags, IterState *iter, int advance, const int *__restrict__ list, int count,
                 int64_t ND, double *sigma_am) {
    constexpr int kTab = 128;                      // diagonals whose offsets are staged in shared memory
    __shared__ int64_t s_kc[kTab], s_ki[kTab];
    rlp::pdl_launch_dependents();
    const int gid = blockIdx.x * kBlock + threadIdx.x;
    const int g = gid >> 2, a = gid & 3;
    const bool valid = g < count;
    const int s = valid ? (list ? list[g] : g) : 0;
    if (threadIdx.x < kTab) {                      // the offset tables are tiny and shared by every thread
        const int kmax = iset_size[0];             // sizes are sorted descending
        if ((int)threadIdx.x < kmax) { s_kc[threadIdx.x] = koff_col[threadIdx.x]; s_ki[threadIdx.x] = koff_iset[threadIdx.x]; }
    }
    const int size = valid ? iset_size[s] : 0;
    const int64_t c0 = valid ? col_off[s] : 0;
    const int na = valid ? (int)(col_off[s + 1] - c0) : 0;
    const bool active = a < na;
    double r = 0.0, sa = 0.0, sg = 0.0;
    __syncthreads();
    // Only static tables above: under programmatic launch this block may start while ANY earlier kernel of the
    // chain is still running, so regrets / sums / strategy / contributions are read after the wait.
    rlp::pdl_wait();
    const int tl_slot = rlp_tl_begin(iter, advance ? 4 : 3);
    if (gid == 0 && advance) iter->t += 1;
    if (active) {
        r = R[c0 + a]; sa = S[c0 + a];
        sg = sigma[c0 + a];
        const double *cr = contrib_r + (int64_t)a * ND + s;      // action-major: same slots as contrib_o
        const double *co = contrib_o + s;
        int k = 0;
        const int fast = size < kTab ? size : kTab;
        // batches of 8 diagonals, the last one predicated: all of a batch's loads are in flight before the ordered adds
        for (; k < fast; k += 8) {
            double x[8], y[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const bool in = k + j < fast;
                x[j] = in ? cr[s_ki[k + j]] : 0.0;
                y[j] = in ? co[s_ki[k + j]] : 0.0;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (k + j < fast) { r = r + x[j]; sa = sa + y[j] * sg; }
        }
        for (k = fast; k < size; ++k) {
            r = r + cr[koff_iset[k]];
            sa = sa + co[koff_iset[k]] * sg;
        }
        R[c0 + a] = r;
        S[c0 + a] = sa;
    }
    rlp_tl_end(iter, tl_slot);
    const unsigned lane = threadIdx.x & 31u, base = lane & ~3u;
    double rr[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) rr[j] = __shfl_sync(0xffffffffu, r, base + j);
    if (!active) return;
    double mx = rr[0];
#pragma unroll
    for (int j = 1; j < 4; ++j) if (j < na) mx = rr[j] > mx ? rr[j] : mx;
    double out;
    if (mx <= 0.0) {
        out = 1.0 / (double)na;
    } else {
        rlp::Neumaier acc;
        double mine = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j < na) {
                double p = rr[j] > 0.0 ? rr[j] : 0.0;
                p = p > rlp::kEpsilon ? p : rlp::kEpsilon;
                acc.add(p);
                if (j == a) mine = p;
            }
        }
        out = mine / acc.total();
    }
    sigma[c0 + a] = out;
    sigma_am[(int64_t)a * I + s] = out;
    if (a == 0) { flags[s] = 1; flags[I + s] = 1; flags[2 * I + s] = 1; }
}

// Multi-GPU split: local deltas into delta[0..Q) (regret) and delta[Q..2Q) (strategy sum), canonical order.
__global__ void __launch_bounds__(kBlock)
k_local_delta(int32_t I, int64_t Q, const int32_t *__restrict__ iset_size, const int64_t *__restrict__ col_off,
              const int64_t *__restrict__ koff_col, const int64_t *__restrict__ koff_iset,
              const int64_t *__restrict__ canon_of_col, const double *__restrict__ contrib_r,
              const double *__restrict__ contrib_o, const double *__restrict__ sigma, double *delta, int64_t ND) {
    int s = blockIdx.x * kBlock + threadIdx.x;
    if (s >= I) return;
    int size = iset_size[s];
    int64_t c0 = col_off[s];
    int na = (int)(col_off[s + 1] - c0);
    for (int a = 0; a < na; ++a) {
        double r = 0.0, sa = 0.0, sg = sigma[c0 + a];
        for (int k = 0; k < size; ++k) {
            r = r + contrib_r[(int64_t)a * ND + koff_iset[k] + s];
            sa = sa + contrib_o[koff_iset[k] + s] * sg;
        }
        int64_t q = canon_of_col[c0 + a];
        delta[q] = r;
        delta[Q + q] = sa;
    }
}

// Four lanes per information set (<= 4 actions): the deltas of one action per lane, loads batched like
// k_accumulate_rm4.  delta[q] (regret) and delta[Q + q] (strategy sum) in canonical pair order.
__global__ void __launch_bounds__(kBlock)
k_local_delta4(int32_t I, int64_t Q, const int32_t *__restrict__ iset_size, const int64_t *__restrict__ col_off,
               const int64_t *__restrict__ koff_col, const int64_t *__restrict__ koff_iset,
               const int64_t *__restrict__ canon_of_col, const double *__restrict__ contrib_r,
               const double *__restrict__ contrib_o, const double *__restrict__ sigma, double *delta, int64_t ND) {
    const int gid = blockIdx.x * kBlock + threadIdx.x;
    const int s = gid >> 2, a = gid & 3;
    if (s >= I) return;
    const int size = iset_size[s];
    const int64_t c0 = col_off[s];
    const int na = (int)(col_off[s + 1] - c0);
    if (a >= na) return;
    const double sg = sigma[c0 + a];
    const double *cr = contrib_r + (int64_t)a * ND + s;
    const double *co = contrib_o + s;
    double r = 0.0, sa = 0.0;
    int k = 0;
    for (; k + 8 <= size; k += 8) {
        double x[8], y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { x[j] = cr[koff_iset[k + j]]; y[j] = co[koff_iset[k + j]]; }
#pragma unroll
        for (int j = 0; j < 8; ++j) { r = r + x[j]; sa = sa + y[j] * sg; }
    }
    for (; k < size; ++k) { r = r + cr[koff_iset[k]]; sa = sa + co[koff_iset[k]] * sg; }
    const int64_t q = canon_of_col[c0 + a];
    delta[q] = r;
    delta[Q + q] = sa;
}

__global__ void __launch_bounds__(kBlock)
k_apply_delta(int32_t I, int64_t Q, const int64_t *__restrict__ col_off, const int64_t *__restrict__ canon_of_col,
              const double *__restrict__ delta, double *R, double *S, double *sigma, uint8_t *flags, IterState *iter,
              double *sigma_am) {
    int s = blockIdx.x * kBlock + threadIdx.x;
    if (s == 0) iter->t += 1;
    if (s >= I) return;
    int64_t c0 = col_off[s];
    int na = (int)(col_off[s + 1] - c0);
    for (int a = 0; a < na; ++a) {
        int64_t q = canon_of_col[c0 + a];
        R[c0 + a] = R[c0 + a] + delta[q];
        S[c0 + a] = S[c0 + a] + delta[Q + q];
    }
    rlp::regret_matching(R + c0, na, sigma + c0);
    for (int a = 0; a < na; ++a) sigma_am[(int64_t)a * I + s] = sigma[c0 + a];
    flags[s] = 1; flags[I + s] = 1; flags[2 * I + s] = 1;
}

// ---- average strategy (cfr.py:32-41 / extensive_game.py:67-77) -----------------------------------
__global__ void __launch_bounds__(kBlock)
k_average(int32_t I, const int64_t *__restrict__ col_off, const double *__restrict__ S, double *avg) {
    int s = blockIdx.x * kBlock + threadIdx.x;
    if (s >= I) return;
    int64_t c0 = col_off[s];
    int na = (int)(col_off[s + 1] - c0);
    rlp::Neumaier acc;
    for (int a = 0; a < na; ++a) acc.add(S[c0 + a]);
    double tot = acc.total();
    for (int a = 0; a < na; ++a) avg[c0 + a] = tot > 0.0 ? S[c0 + a] / tot : 1.0 / (double)na;
}

__global__ void __launch_bounds__(kBlock) k_uniform(int32_t I, const int64_t *__restrict__ col_off, double *sigma) {
    int s = blockIdx.x * kBlock + threadIdx.x;
    if (s >= I) return;
    int64_t c0 = col_off[s];
    int na = (int)(col_off[s + 1] - c0);
    for (int a = 0; a < na; ++a) sigma[c0 + a] = 1.0 / (double)na;
}

// sigma_am[action][set] <- sigma[set][action]
__global__ void __launch_bounds__(kBlock)
k_sigma_am(int32_t I, const int64_t *__restrict__ col_off, const double *__restrict__ sigma, double *sigma_am) {
    int s = blockIdx.x * kBlock + threadIdx.x;
    if (s >= I) return;
    int64_t c0 = col_off[s];
    int na = (int)(col_off[s + 1] - c0);
    for (int a = 0; a < na; ++a) sigma_am[(int64_t)a * I + s] = sigma[c0 + a];
}

__global__ void k_set_iter(IterState *it, long long t, int linear) {
    it->t = t;
    it->linear = linear;
}

__global__ void __launch_bounds__(kBlock) k_permute(int64_t Q, const int64_t *__restrict__ canon_of_col, const double *src,
                                                    double *dst, int to_internal) {
    int64_t q = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (q >= Q) return;
    if (to_internal) dst[q] = src[canon_of_col[q]]; else dst[canon_of_col[q]] = src[q];
}

inline unsigned grid_for(int64_t n) { return (unsigned)((n + kBlock - 1) / kBlock); }

View make_view(rlp_cfr *c) {
    rlp_tree *t = c->tree;
    View v;
    v.first_child = t->first_child.p; v.colbase = t->colbase.p; v.srank = t->srank.p; v.krank = t->krank.p;
    v.player = t->player.p; v.cprob = t->cprob.p; v.sigma = c->sigma.p;
    v.koff_col = t->koff_col.p; v.koff_iset = t->koff_iset.p;
    v.pc = c->pc.p; v.p1 = c->p1.p; v.p2 = c->p2.p; v.val1 = c->val1.p; v.val2 = c->val2.p;
    v.contrib_r = c->contrib_r.p; v.contrib_o = c->contrib_o.p; v.iter = c->iter.p;
    v.ND = t->ND;
    return v;
}

// forward + backward sweep of one iteration on stream s; returns #kernels via *count
template <bool ZS>
int enqueue_sweep(rlp_cfr *c, cudaStream_t s, int *count, const IterState *iter = nullptr, cudaEvent_t *marks = nullptr) {
    rlp_tree *t = c->tree;
    if (t->tiles.enabled) {                 // shared-memory tiled sweep: one launch instead of one per level
        return rlp::launch_tile_sweep(c, s, iter, count, marks);
    }
    View v = make_view(c);
    if (iter) v.iter = iter;
    int L = t->L, top = std::min<int>(t->top_levels, L - 1);   // parents live on levels 0 .. L-2
    TopLevels tl;
    tl.n = std::min(top, 17);
    top = tl.n;
    for (int l = 0; l <= tl.n; ++l) tl.off[l] = t->level_off[l];
    int n = 0;
    if (top > 0) { k_fwd_top<<<1, kTopBlock, 0, s>>>(v, tl); RLP_LAUNCH_CHECK(); ++n; }
    for (int l = top; l + 1 < L; ++l) {
        int64_t lo = t->level_off[l], hi = t->level_off[l + 1];
        k_fwd_level<<<grid_for(hi - lo), kBlock, 0, s>>>(v, lo, hi); RLP_LAUNCH_CHECK(); ++n;
    }
    for (int l = L - 2; l >= top; --l) {
        int64_t lo = t->level_off[l], hi = t->level_off[l + 1];
        k_bwd_level<ZS><<<grid_for(hi - lo), kBlock, 0, s>>>(v, lo, hi); RLP_LAUNCH_CHECK(); ++n;
    }
    if (top > 0) { k_bwd_top<ZS><<<1, kTopBlock, 0, s>>>(v, tl); RLP_LAUNCH_CHECK(); ++n; }
    *count = n;
    return RLP_OK;
}

int enqueue_iteration(rlp_cfr *c, cudaStream_t s, int *count, cudaEvent_t *marks = nullptr) {
    rlp_tree *t = c->tree;
    int n = 0;
    if (marks && !t->tiles.enabled) RLP_CUDA(cudaEventRecord(marks[0], s));
    int rc = t->zero_sum ? enqueue_sweep<true>(c, s, &n, nullptr, marks) : enqueue_sweep<false>(c, s, &n, nullptr, marks);
    if (rc) return rc;
    if (marks && !t->tiles.enabled) { for (int k = 1; k <= 3; ++k) RLP_CUDA(cudaEventRecord(marks[k], s)); }
    if (t->I > 0) {
        if (t->max_act <= 4)
            RLP_CUDA(rlp::launch_pdl(k_accumulate_rm4, dim3(grid_for(4 * (int64_t)t->I)), dim3(kBlock), 0, s, t->I,
                                     t->iset_size.p, t->col_off.p, t->koff_col.p, t->koff_iset.p, c->contrib_r.p,
                                     c->contrib_o.p, c->R.p, c->S.p, c->sigma.p, c->flags.p, c->iter.p, 1,
                                     (const int *)nullptr, (int)t->I, (int64_t)t->ND, c->sigma_am.p));
        else
            k_accumulate_rm<<<grid_for(t->I), kBlock, 0, s>>>(t->I, t->iset_size.p, t->col_off.p, t->koff_col.p,
                                                            t->koff_iset.p, c->contrib_r.p, c->contrib_o.p, c->R.p,
                                                            c->S.p, c->sigma.p, c->flags.p, c->iter.p, 1, (int64_t)t->ND,
                                                            c->sigma_am.p);
        RLP_LAUNCH_CHECK(); ++n;
    }
    if (marks) RLP_CUDA(cudaEventRecord(marks[4], s));
    *count = n;
    return RLP_OK;
}

// Two-branch pipeline of `iters` iterations (tiled trees): after the micro kernel, the accumulate + regret matching
// of the information sets that live only in micro subtrees (A, the bulk) runs on a side stream while the main stream
// finishes the upper tiles, accumulates the few information sets they touch (B) and already pushes the next
// iteration's reach down to the portals; the branches join before the next micro kernel.
bool can_pipeline(rlp_cfr *c) {
    rlp_tree *t = c->tree;
    const TileSet &ts = t->tiles;
    return ts.enabled && t->max_act <= 4 && ts.num_tiles > 0 && ts.num_micro > 0 && ts.acc_a_count > 0 &&
           ts.acc_b_count > 0 && c->side && c->ev_fork && c->ev_join;
}

int launch_acc(rlp_cfr *c, cudaStream_t s, const int *list, int count, int advance) {
    rlp_tree *t = c->tree;
    RLP_CUDA(rlp::launch_pdl(k_accumulate_rm4, dim3(grid_for(4 * (int64_t)count)), dim3(kBlock), 0, s, t->I,
                             t->iset_size.p, t->col_off.p, t->koff_col.p, t->koff_iset.p, c->contrib_r.p, c->contrib_o.p,
                             c->R.p, c->S.p, c->sigma.p, c->flags.p, c->iter.p, advance, list, count, (int64_t)t->ND,
                             c->sigma_am.p));
    rlp::g_launches.fetch_add(1, std::memory_order_relaxed);
    return RLP_OK;
}

int enqueue_pipelined(rlp_cfr *c, cudaStream_t s, int iters, int *count) {
    rlp_tree *t = c->tree;
    TileSet &ts = t->tiles;
    int n = 0, rc;
    if ((rc = rlp::launch_tile_stage(c, s, nullptr, 0, &n))) return rc;
    for (int u = 0; u < iters; ++u) {
        if ((rc = rlp::launch_tile_stage(c, s, nullptr, 1, &n))) return rc;
        RLP_CUDA(cudaEventRecord(c->ev_fork, s));
        RLP_CUDA(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
        if ((rc = launch_acc(c, c->side, ts.acc_list_a.p, ts.acc_a_count, 0))) return rc;
        RLP_CUDA(cudaEventRecord(c->ev_join, c->side));
        if ((rc = rlp::launch_tile_stage(c, s, nullptr, 2, &n))) return rc;
        if ((rc = launch_acc(c, s, ts.acc_list_b.p, ts.acc_b_count, 1))) return rc;
        n += 2;
        if (u + 1 < iters && (rc = rlp::launch_tile_stage(c, s, nullptr, 0, &n))) return rc;
        RLP_CUDA(cudaStreamWaitEvent(s, c->ev_join, 0));
    }
    *count = n;
    return RLP_OK;
}

constexpr int kGraphUnroll = 10;     // = the evaluation cadence of cfr() (cfr.py:125), so a 10-iteration call is one graph launch
}  // namespace

// -------------------------------------------------------------------------------------------------
namespace rlp {
int launch_average(rlp_cfr *c, cudaStream_t s) {
    rlp_tree *t = c->tree;
    if (t->I == 0) return RLP_OK;
    k_average<<<grid_for(t->I), kBlock, 0, s>>>(t->I, t->col_off.p, c->S.p, c->avg.p);
    RLP_LAUNCH_CHECK();
    return RLP_OK;
}
int permute(rlp_tree *t, const double *src, double *dst, int to_internal, cudaStream_t s) {
    if (t->Q == 0) return RLP_OK;
    k_permute<<<grid_for(t->Q), kBlock, 0, s>>>(t->Q, t->d_canon_of_col.p, src, dst, to_internal);
    RLP_LAUNCH_CHECK();
    return RLP_OK;
}
}  // namespace rlp

extern "C" void rlp_comm_destroy(rlp_cfr *c);
namespace rlp { void p2p_release(rlp_cfr *c); }
extern "C" void rlp_cfr_free(rlp_cfr *c) {
    if (c) rlp::p2p_release(c);
    rlp_comm_destroy(c);
    delete c;
}

extern "C" int rlp_cfr_reset(rlp_cfr *c, void *stream_) {
    if (!c) return fail(RLP_ERR_INVALID, "null solver");
    cudaStream_t s = (cudaStream_t)stream_;
    rlp_tree *t = c->tree;
    RLP_CUDA(cudaMemsetAsync(c->R.p, 0, c->R.bytes(), s));
    RLP_CUDA(cudaMemsetAsync(c->S.p, 0, c->S.bytes(), s));
    RLP_CUDA(cudaMemsetAsync(c->cum_imm.p, 0, c->cum_imm.bytes(), s));
    RLP_CUDA(cudaMemsetAsync(c->flags.p, 0, c->flags.bytes(), s));
    RLP_CUDA(cudaMemsetAsync(c->iter.p, 0, 2 * sizeof(IterState), s));
    if (t->I > 0) { k_uniform<<<grid_for(t->I), kBlock, 0, s>>>(t->I, t->col_off.p, c->sigma.p); RLP_LAUNCH_CHECK(); }
    c->sigma_am_dirty = true;
    c->sigF_dirty = true;
    c->nodes_touched = 0;
    c->touched_dirty = false;
    c->external_state = false;
    c->dev_t = -1;
    return RLP_OK;
}

extern "C" int rlp_cfr_create(rlp_tree *t, void *stream_, rlp_cfr **out) {
    if (!t || !out) return fail(RLP_ERR_INVALID, "null argument");
    *out = nullptr;
    cudaStream_t s = (cudaStream_t)stream_;
    std::unique_ptr<rlp_cfr> c(new (std::nothrow) rlp_cfr);
    if (!c) return fail(RLP_ERR_NOMEM, "out of host memory");
    c->tree = t;
    size_t Q = (size_t)t->Q, N = (size_t)t->N;
    RLP_CUDA(c->R.alloc(Q)); RLP_CUDA(c->S.alloc(Q)); RLP_CUDA(c->sigma.alloc(Q)); RLP_CUDA(c->avg.alloc(Q));
    RLP_CUDA(c->cum_imm.alloc(Q)); RLP_CUDA(c->tmpQ.alloc(Q)); RLP_CUDA(c->delta.alloc(2 * Q));
    RLP_CUDA(c->flags.alloc(3 * (size_t)t->I));
    RLP_CUDA(c->pc.alloc(N)); RLP_CUDA(c->p1.alloc(N)); RLP_CUDA(c->p2.alloc(N));
    RLP_CUDA(c->val1.alloc(N)); RLP_CUDA(c->vbr.alloc(N));
    if (!t->zero_sum) RLP_CUDA(c->val2.alloc(N));
    // regret contributions are action-major [max_act][ND]; the best-response sweep reuses the buffer with its own
    // [diagonal][column] indexing (NC entries)
    RLP_CUDA(c->contrib_r.alloc(std::max((size_t)t->NC, (size_t)t->max_act * (size_t)t->ND)));
    RLP_CUDA(c->contrib_o.alloc((size_t)t->ND));
    RLP_CUDA(c->sigma_am.alloc((size_t)std::max(1, t->max_act) * (size_t)t->I));
    RLP_CUDA(c->astar.alloc((size_t)t->I));
    {
        // the side branch (bulk accumulate) yields to the main chain, which is the critical path
        int lo = 0, hi = 0;
        RLP_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        RLP_CUDA(cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, lo));
    }
    RLP_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    RLP_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    RLP_CUDA(c->iter.alloc(2));    // [1] stays {t=0, linear=0}: weight 1.0 for metric sweeps
    RLP_CUDA(c->scalars.alloc(16));
    // terminal values never change: val = payoff everywhere, interior nodes are rewritten every sweep
    RLP_CUDA(cudaMemcpyAsync(c->val1.p, t->pay1.p, sizeof(double) * N, cudaMemcpyDeviceToDevice, s));
    if (!t->zero_sum) RLP_CUDA(cudaMemcpyAsync(c->val2.p, t->pay2.p, sizeof(double) * N, cudaMemcpyDeviceToDevice, s));
    const double one = 1.0;
    RLP_CUDA(cudaMemcpyAsync(c->pc.p, &one, sizeof(double), cudaMemcpyHostToDevice, s));
    RLP_CUDA(cudaMemcpyAsync(c->p1.p, &one, sizeof(double), cudaMemcpyHostToDevice, s));
    RLP_CUDA(cudaMemcpyAsync(c->p2.p, &one, sizeof(double), cudaMemcpyHostToDevice, s));
    RLP_CUDA(cudaStreamSynchronize(s));
    int rc = rlp_cfr_reset(c.get(), stream_);
    if (rc) return rc;
    *out = c.release();
    return RLP_OK;
}

namespace rlp {
// Bring the action-major strategy copy up to date (after a reset, a host write or the sampled variants; the vanilla
// kernels keep it current themselves).
int ensure_sigma_am(rlp_cfr *c, cudaStream_t s) {
    if (!c->sigma_am_dirty) return RLP_OK;
    rlp_tree *t = c->tree;
    if (t->I > 0) { k_sigma_am<<<grid_for(t->I), kBlock, 0, s>>>(t->I, t->col_off.p, c->sigma.p, c->sigma_am.p); RLP_LAUNCH_CHECK(); }
    c->sigma_am_dirty = false;
    return RLP_OK;
}
}  // namespace rlp

static int build_graph(rlp_cfr *c);
static int ensure_graph(rlp_cfr *c) {
    if (c->graph_exec) return RLP_OK;
    int rc = build_graph(c);
    if (rc != RLP_OK && rlp::pdl_enabled()) {       // retry without programmatic dependent launch
        cudaGetLastError();
        rlp::pdl_disable();
        if (c->graph) { cudaGraphDestroy(c->graph); c->graph = nullptr; }
        rc = build_graph(c);
    }
    return rc;
}

static int build_graph(rlp_cfr *c) {
    cudaStream_t cap;
    RLP_CUDA(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
    cudaError_t e = cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) { cudaStreamDestroy(cap); return fail(RLP_ERR_CUDA, "begin capture: %s", cudaGetErrorString(e)); }
    uint64_t before = rlp::g_launches.load();
    int total = 0, rc = RLP_OK;
    if (can_pipeline(c)) {
        rc = enqueue_pipelined(c, cap, kGraphUnroll, &total);
    } else {
        for (int u = 0; u < kGraphUnroll && rc == RLP_OK; ++u) {
            int n = 0;
            rc = enqueue_iteration(c, cap, &n);
            total += n;
        }
    }
    e = cudaStreamEndCapture(cap, &c->graph);
    rlp::g_launches.store(before);      // capture records, it does not launch
    cudaStreamDestroy(cap);
    if (rc) return rc;
    if (e != cudaSuccess) return fail(RLP_ERR_CUDA, "end capture: %s", cudaGetErrorString(e));
    RLP_CUDA(cudaGraphInstantiate(&c->graph_exec, c->graph, 0));
    c->graph_kernels = total;
    return RLP_OK;
}

extern "C" int rlp_cfr_vanilla_iters(rlp_cfr *c, int64_t t0, int64_t n_iters, int linear_weight, void *stream_) {
    if (!c || n_iters < 0) return fail(RLP_ERR_INVALID, "bad argument");
    cudaStream_t s = (cudaStream_t)stream_;
    if (rlp::fused_available(c)) {          // one cooperative launch runs all n_iters iterations (fused.cu)
        c->dev_t = -1;
        if (n_iters > 0) { int rc = rlp::fused_run(c, t0, n_iters, linear_weight, s, nullptr, 0); if (rc) return rc; }
        c->nodes_touched += 2ull * (uint64_t)c->tree->N * (uint64_t)n_iters;
        return RLP_OK;
    }
    k_set_iter<<<1, 1, 0, s>>>(c->iter.p, (long long)t0, linear_weight);
    RLP_LAUNCH_CHECK();
    c->dev_t = -1;
    c->sigF_dirty = true;
    { int rc = rlp::ensure_sigma_am(c, s); if (rc) return rc; }
    int64_t left = n_iters;
    if (left >= kGraphUnroll) {
        int rc = ensure_graph(c);
        if (rc) return rc;
        for (; left >= kGraphUnroll; left -= kGraphUnroll) {
            RLP_CUDA(cudaGraphLaunch(c->graph_exec, s));
            rlp::g_launches.fetch_add(c->graph_kernels, std::memory_order_relaxed);
        }
    }
    for (; left > 0; --left) {
        int n = 0;
        int rc = enqueue_iteration(c, s, &n);
        if (rc) return rc;
    }
    c->nodes_touched += 2ull * (uint64_t)c->tree->N * (uint64_t)n_iters;
    return RLP_OK;
}

extern "C" uint64_t rlp_nodes_touched(rlp_cfr *c) {
    if (!c) return 0;
    unsigned long long dev = 0;
    if (c->touched_dirty) {      // sampled lanes count on the device
        cudaDeviceSynchronize();
        cudaMemcpy(&dev, &c->iter.p->touched, sizeof(dev), cudaMemcpyDeviceToHost);
        if (c->fr_count.p) {    // the frontier form of the lanes flags a record overflow instead of writing out of bounds
            unsigned int over = 0;
            cudaMemcpy(&over, c->fr_count.p + c->fr_levels, sizeof(over), cudaMemcpyDeviceToHost);
            if (over) { fail(RLP_ERR_CUDA, "sampled lanes: frontier records overflowed their bound"); return 0; }
        }
    }
    return c->nodes_touched + dev;
}

extern "C" int rlp_cfr_local_sweep(rlp_cfr *c, int64_t t_iter, int linear_weight, void *stream_) {
    if (!c) return fail(RLP_ERR_INVALID, "null solver");
    cudaStream_t s = (cudaStream_t)stream_;
    rlp_tree *t = c->tree;
    if (t_iter >= 0) {          // t < 0: keep the device-side counter (lets the caller capture the call in a CUDA graph)
        k_set_iter<<<1, 1, 0, s>>>(c->iter.p, (long long)t_iter, linear_weight);
        RLP_LAUNCH_CHECK();
        c->dev_t = -1;          // callers that track the counter (comm.cu) set it again themselves
    }
    if (c->sigma_am_dirty) { int rc0 = rlp::ensure_sigma_am(c, s); if (rc0) return rc0; }
    int n = 0;
    int rc = t->zero_sum ? enqueue_sweep<true>(c, s, &n) : enqueue_sweep<false>(c, s, &n);
    if (rc) return rc;
    if (t->I > 0) {
        if (t->max_act <= 4)
            k_local_delta4<<<grid_for(4 * (int64_t)t->I), kBlock, 0, s>>>(t->I, t->Q, t->iset_size.p, t->col_off.p,
                                                                     t->koff_col.p, t->koff_iset.p, t->d_canon_of_col.p,
                                                                     c->contrib_r.p, c->contrib_o.p, c->sigma.p, c->delta.p,
                                                                     (int64_t)t->ND);
        else
            k_local_delta<<<grid_for(t->I), kBlock, 0, s>>>(t->I, t->Q, t->iset_size.p, t->col_off.p, t->koff_col.p,
                                                          t->koff_iset.p, t->d_canon_of_col.p, c->contrib_r.p,
                                                          c->contrib_o.p, c->sigma.p, c->delta.p, (int64_t)t->ND);
        RLP_LAUNCH_CHECK();
    }
    return RLP_OK;
}

extern "C" void *rlp_cfr_delta_ptr(rlp_cfr *c) { return c ? (void *)c->delta.p : nullptr; }

extern "C" int rlp_cfr_apply_delta(rlp_cfr *c, void *stream_) {
    if (!c) return fail(RLP_ERR_INVALID, "null solver");
    cudaStream_t s = (cudaStream_t)stream_;
    rlp_tree *t = c->tree;
    if (t->I > 0) {
        k_apply_delta<<<grid_for(t->I), kBlock, 0, s>>>(t->I, t->Q, t->col_off.p, t->d_canon_of_col.p, c->delta.p,
                                                      c->R.p, c->S.p, c->sigma.p, c->flags.p, c->iter.p, c->sigma_am.p);
        RLP_LAUNCH_CHECK();
    }
    c->sigF_dirty = true;
    c->nodes_touched += 2ull * (uint64_t)t->N;
    return RLP_OK;
}

// ---- I/O in canonical order -----------------------------------------------------------------------
namespace rlp { int permute(rlp_tree *t, const double *src, double *dst, int to_internal, cudaStream_t s); }

static double *which_buffer(rlp_cfr *c, int which) {
    switch (which) {
        case RLP_REGRET: return c->R.p;
        case RLP_STRAT_SUM: return c->S.p;
        case RLP_SIGMA: return c->sigma.p;
        case RLP_AVERAGE: return c->avg.p;
        case RLP_CUM_IMM_REGRET: return c->cum_imm.p;
        default: return nullptr;
    }
}

extern "C" int rlp_cfr_read(rlp_cfr *c, int which, double *dst, void *stream_) {
    if (!c || !dst) return fail(RLP_ERR_INVALID, "null argument");
    cudaStream_t s = (cudaStream_t)stream_;
    double *src = which_buffer(c, which);
    if (!src) return fail(RLP_ERR_INVALID, "unknown array %d", which);
    if (which == RLP_AVERAGE) { int rc = rlp::launch_average(c, s); if (rc) return rc; }
    int rc = rlp::permute(c->tree, src, c->tmpQ.p, 0, s);
    if (rc) return rc;
    RLP_CUDA(cudaMemcpyAsync(dst, c->tmpQ.p, sizeof(double) * (size_t)c->tree->Q, cudaMemcpyDeviceToHost, s));
    RLP_CUDA(cudaStreamSynchronize(s));
    return RLP_OK;
}

extern "C" int rlp_cfr_write(rlp_cfr *c, int which, const double *src, void *stream_) {
    if (!c || !src) return fail(RLP_ERR_INVALID, "null argument");
    cudaStream_t s = (cudaStream_t)stream_;
    double *dst = which_buffer(c, which);
    if (!dst || which == RLP_AVERAGE) return fail(RLP_ERR_INVALID, "array %d is not writable", which);
    RLP_CUDA(cudaMemcpyAsync(c->tmpQ.p, src, sizeof(double) * (size_t)c->tree->Q, cudaMemcpyHostToDevice, s));
    int rc = rlp::permute(c->tree, c->tmpQ.p, dst, 1, s);
    if (rc) return rc;
    if (which == RLP_SIGMA) { c->sigma_am_dirty = true; c->sigF_dirty = true; }
    // No synchronisation: a copy from pageable memory has been staged by the time cudaMemcpyAsync returns, and a
    // pinned source must simply stay untouched until the stream reaches this point (it is ordered before every later
    // call on the same stream, including rlp_cfr_read, which does synchronise).
    return RLP_OK;
}

extern "C" int rlp_cfr_read_flags(rlp_cfr *c, uint8_t *dst, void *stream_) {
    if (!c || !dst) return fail(RLP_ERR_INVALID, "null argument");
    cudaStream_t s = (cudaStream_t)stream_;
    rlp_tree *t = c->tree;
    size_t I = (size_t)t->I;
    std::vector<uint8_t> tmp(3 * I);
    RLP_CUDA(cudaMemcpyAsync(tmp.data(), c->flags.p, 3 * I, cudaMemcpyDeviceToHost, s));
    RLP_CUDA(cudaStreamSynchronize(s));
    for (size_t k = 0; k < I; ++k)
        dst[t->iset_of_srank[k]] = (uint8_t)((tmp[k] ? 1 : 0) | (tmp[I + k] ? 2 : 0) | (tmp[2 * I + k] ? 4 : 0));
    return RLP_OK;
}

// ---- immediate regret (cfr_metrics.py:13-78), one strategy per call ------------------------------------
extern "C" int rlp_cfr_immediate_regret_detail(rlp_cfr *c, int64_t step, const int32_t *order, double *out,
                                               double *per_iset_out, void *stream_);
namespace {
__global__ void __launch_bounds__(kBlock)
k_imm_regret(int32_t I, const int32_t *__restrict__ iset_size, const int64_t *__restrict__ col_off,
             const int64_t *__restrict__ koff_iset, const double *__restrict__ contrib_r, double *cum, double *per_iset,
             double inv_steps, int64_t ND) {
    int s = blockIdx.x * kBlock + threadIdx.x;
    if (s >= I) return;
    int size = iset_size[s];
    int64_t c0 = col_off[s];
    int na = (int)(col_off[s + 1] - c0);
    double mx = 0.0;
    for (int a = 0; a < na; ++a) {
        double r = 0.0;
        for (int k = 0; k < size; ++k) r = r + contrib_r[(int64_t)a * ND + koff_iset[k] + s];
        double cu = cum[c0 + a] + r;
        cum[c0 + a] = cu;
        mx = (a == 0 || cu > mx) ? cu : mx;
    }
    per_iset[s] = inv_steps * mx;
}
}  // namespace

extern "C" int rlp_cfr_immediate_regret_step(rlp_cfr *c, int64_t step, const int32_t *order, double *out, void *stream_) {
    return rlp_cfr_immediate_regret_detail(c, step, order, out, nullptr, stream_);
}

// Same, and additionally the per-information-set terms 1/(1+step) * max_a total[I,a] (the t-th dictionary of
// cfr_metrics.compute_immediate_regret's third return value, cfr_metrics.py:68-73) into per_iset[I] (canonical ids).
extern "C" int rlp_cfr_immediate_regret_detail(rlp_cfr *c, int64_t step, const int32_t *order, double *out,
                                               double *per_iset_out, void *stream_) {
    if (!c || !out || step < 0) return fail(RLP_ERR_INVALID, "bad argument");
    cudaStream_t s = (cudaStream_t)stream_;
    rlp_tree *t = c->tree;
    int n = 0;
    { int rc0 = rlp::ensure_sigma_am(c, s); if (rc0) return rc0; }
    int rc = t->zero_sum ? enqueue_sweep<true>(c, s, &n, c->iter.p + 1) : enqueue_sweep<false>(c, s, &n, c->iter.p + 1);
    if (rc) return rc;
    double inv = 1.0 / (double)(1 + step);
    // per-information-set values land in the (otherwise idle) own-reach scratch: ND >= I doubles
    double *per_iset = c->contrib_o.p;
    if (t->I > 0) {
        k_imm_regret<<<grid_for(t->I), kBlock, 0, s>>>(t->I, t->iset_size.p, t->col_off.p, t->koff_iset.p, c->contrib_r.p,
                                                     c->cum_imm.p, per_iset, inv, (int64_t)t->ND);
        RLP_LAUNCH_CHECK();
    }
    std::vector<double> h((size_t)t->I);
    RLP_CUDA(cudaMemcpyAsync(h.data(), per_iset, sizeof(double) * (size_t)t->I, cudaMemcpyDeviceToHost, s));
    RLP_CUDA(cudaStreamSynchronize(s));
    // builtin sum() over the reference's dict order (cfr_metrics.py:74) -- I scalars, host side
    rlp::Neumaier acc;
    for (int32_t k = 0; k < t->I; ++k) acc.add(h[t->srank_of_iset[order ? order[k] : k]]);
    *out = acc.total();
    if (per_iset_out)
        for (int32_t k = 0; k < t->I; ++k) per_iset_out[k] = h[t->srank_of_iset[k]];
    return RLP_OK;
}

extern "C" int rlp_cfr_write_flags(rlp_cfr *c, const uint8_t *src, void *stream_) {
    if (!c || !src) return fail(RLP_ERR_INVALID, "null argument");
    cudaStream_t s = (cudaStream_t)stream_;
    rlp_tree *t = c->tree;
    size_t I = (size_t)t->I;
    std::vector<uint8_t> tmp(3 * I);
    for (size_t k = 0; k < I; ++k) {
        uint8_t f = src[t->iset_of_srank[k]];
        tmp[k] = f & 1; tmp[I + k] = (f >> 1) & 1; tmp[2 * I + k] = (f >> 2) & 1;
    }
    RLP_CUDA(cudaMemcpyAsync(c->flags.p, tmp.data(), 3 * I, cudaMemcpyHostToDevice, s));
    RLP_CUDA(cudaStreamSynchronize(s));
    return RLP_OK;
}

// Per-stage device time of `reps` iterations (ms, summed): tiled sweep -> {upper tiles phase 0, micro subtrees,
// upper tiles phase 1, accumulate + regret matching}; level sweep -> {0, 0, forward + backward levels, accumulate}.
extern "C" int rlp_cfr_profile_stages(rlp_cfr *c, int reps, float *out_ms4, void *stream_) {
    if (!c || !out_ms4 || reps < 1) return fail(RLP_ERR_INVALID, "bad argument");
    cudaStream_t s = (cudaStream_t)stream_;
    cudaEvent_t ev[5];
    for (auto &e : ev) RLP_CUDA(cudaEventCreate(&e));
    for (int k = 0; k < 4; ++k) out_ms4[k] = 0.f;
    c->sigF_dirty = true;
    int rc = rlp::ensure_sigma_am(c, s);
    if (rc) return rc;
    for (int r = 0; r < reps + 2 && rc == RLP_OK; ++r) {
        int n = 0;
        rc = enqueue_iteration(c, s, &n, ev);
        if (rc) break;
        RLP_CUDA(cudaEventSynchronize(ev[4]));
        if (r < 2) continue;                      // warm-up
        for (int k = 0; k < 4; ++k) {
            float ms = 0.f;
            RLP_CUDA(cudaEventElapsedTime(&ms, ev[k], ev[k + 1]));
            out_ms4[k] += ms;
        }
    }
    for (auto &e : ev) cudaEventDestroy(e);
    c->nodes_touched += 2ull * (uint64_t)c->tree->N * (uint64_t)(reps + 2);
    return rc;
}

// Debug: record the start/end of block 0 of every sweep kernel for the next `iters` iterations (in-graph schedule).
extern "C" int rlp_cfr_timeline(rlp_cfr *c, int iters, unsigned long long *out, int cap_records, int *n_out, void *stream_) {
    if (!c || !out || !n_out || cap_records < 1) return fail(RLP_ERR_INVALID, "bad argument");
    cudaStream_t s = (cudaStream_t)stream_;
    RLP_CUDA(c->timeline.alloc(3 * (size_t)cap_records));
    RLP_CUDA(cudaMemsetAsync(c->timeline.p, 0, c->timeline.bytes(), s));
    IterState h;
    RLP_CUDA(cudaMemcpyAsync(&h, c->iter.p, sizeof(h), cudaMemcpyDeviceToHost, s));
    RLP_CUDA(cudaStreamSynchronize(s));
    long long t0 = h.t;
    h.tl = c->timeline.p; h.tl_count = 0; h.tl_cap = (unsigned)cap_records;
    RLP_CUDA(cudaMemcpyAsync(c->iter.p, &h, sizeof(h), cudaMemcpyHostToDevice, s));
    int rc = rlp_cfr_vanilla_iters(c, t0, iters, h.linear, s);
    RLP_CUDA(cudaStreamSynchronize(s));
    IterState h2;
    RLP_CUDA(cudaMemcpy(&h2, c->iter.p, sizeof(h2), cudaMemcpyDeviceToHost));
    int n = (int)std::min<unsigned>(h2.tl_count, (unsigned)cap_records);
    RLP_CUDA(cudaMemcpy(out, c->timeline.p, sizeof(unsigned long long) * 3 * (size_t)n, cudaMemcpyDeviceToHost));
    h2.tl = nullptr; h2.tl_count = 0; h2.tl_cap = 0;
    RLP_CUDA(cudaMemcpy(c->iter.p, &h2, sizeof(h2), cudaMemcpyHostToDevice));
    *n_out = n;
    return rc;
}

// Measurement aid for the group-fused persistent kernel: runs `iters` iterations (continuing at iteration t0 with unit
// weights) in one launch while one CTA (the first "leader", which also owns upper unit 0) stamps %globaltimer
// {iteration start, its micro phase done, first grid barrier passed, upper unit: loads issued, fan-in done, tile
// evaluated, accumulate + regret matching done, reach written, its upper phase done}; out_ns receives 9 * iters
// nanosecond stamps.  RLP_ERR_UNSUPPORTED when the tree does not qualify.
extern "C" int rlp_cfr_fused_timeline(rlp_cfr *c, int64_t t0, int iters, unsigned long long *out_ns, void *stream_) {
    if (!c || !out_ns || iters < 1) return fail(RLP_ERR_INVALID, "bad argument");
    if (!rlp::fused_available(c)) return fail(RLP_ERR_UNSUPPORTED, "the fused iteration kernel is not in use for this tree");
    cudaStream_t s = (cudaStream_t)stream_;
    RLP_CUDA(c->timeline.alloc(9 * (size_t)iters));
    RLP_CUDA(cudaMemsetAsync(c->timeline.p, 0, c->timeline.bytes(), s));
    int rc = rlp::fused_run(c, t0, iters, 0, s, c->timeline.p, 9ll * iters);
    if (rc) return rc;
    RLP_CUDA(cudaMemcpyAsync(out_ns, c->timeline.p, c->timeline.bytes(), cudaMemcpyDeviceToHost, s));
    RLP_CUDA(cudaStreamSynchronize(s));
    c->nodes_touched += 2ull * (uint64_t)c->tree->N * (uint64_t)iters;
    return RLP_OK;
}
