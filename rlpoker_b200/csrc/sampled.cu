// Sampled CFR as independent traversal lanes (replaces the recursions of rlpoker/cfr/cfr.py:158-255 with
// use_chance_sampling=True and rlpoker/cfr/external_cfr.py:82-173) plus the full-tree average-strategy
// walk of rlpoker/cfr/cfr_util.py:109-198 as a forward level sweep.
//
// One thread = one traversal (lane, traversing player): an explicit-stack depth-first walk over the
// breadth-first arrays against the strategy frozen for the iteration; chance (and, for external
// sampling, opponent) actions are drawn with the reference's sampler (rlpoker/util.py:17-27 ->
// numpy RandomState.choice: normalise, cumsum, renormalise, searchsorted-right) from a counter-based
// Philox-4x32-10 stream keyed (seed; iteration, lane, player, draw index), so a CPU run fed the same
// stream reproduces every lane exactly.  lanes == 1 is the reference algorithm bit for bit (plain adds in
// visit order); lanes > 1 is mini-batch MCCFR with fp64 atomics (order of adds not deterministic).
#include "cfr.cuh"

using rlp::fail;

namespace {
constexpr int kBlock = 128;
constexpr int kMaxDepth = 32;
constexpr int kMaxAct = 8;        // widest information set the lane kernels handle
constexpr int kMaxChance = 128;   // widest chance node (numpy's pairwise sum is restated up to here)
inline unsigned grid_for(int64_t n, int b = kBlock) { return (unsigned)((n + b - 1) / b); }

struct LaneView {
    const int4 *rec;                       // packed node records (tree.cuh)
    const int32_t *first_child, *colbase, *srank;
    const int8_t *player;
    const double *cprob, *pay1, *pay2, *sigma;
    double *R, *S;
    uint8_t *in_R, *in_S, *in_next;       // [I] each
    IterState *iter;
    uint64_t seed;
    int zero_sum;
    int atomic;                            // lanes > 1
};

// np.sum over a short contiguous double array: sequential below 8 elements, 8 running sums above
// (numpy/core/src/umath/loops_utils.h pairwise sum, n <= 128).
__device__ inline double numpy_sum(const double *p, int n) {
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r += p[i];
        return r;
    }
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = p[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; ++j) r[j] += p[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += p[i];
    return res;
}

// RandomState.choice(len(p), p=p/np.sum(p)) given the uniform u.
__device__ inline int draw(const double *p, int n, double u) {
    double tot = numpy_sum(p, n);
    double last = 0.0;
    for (int a = 0; a < n; ++a) last += p[a] / tot;
    double run = 0.0;
    int idx = 0;
    for (int a = 0; a < n; ++a) {
        run += p[a] / tot;
        if (run / last <= u) idx = a + 1; else break;
    }
    return idx < n ? idx : n - 1;
}

// One stack frame of the depth-first walk.  External sampling needs no reach probabilities (its regret increment is
// unweighted, external_cfr.py:143-152), and the action-value array is as wide as the game's widest information set: the
// per-thread stack is what limits this kernel (it lives in local memory), so it is kept as small as the game allows
// (Leduc external: 16 frames x 48 B instead of 32 x 104 B).
template <int VARIANT, int MAXA>
struct Frame {
    int node;
    int k;            // next child to visit
    double value;     // running expectation
    double va[MAXA];
    double reach[VARIANT == 0 ? 3 : 1];     // chance-sampled: pc, p1, p2
};

__device__ inline void add(double *p, double x, int atomic) {
    if (atomic) atomicAdd(p, x); else *p = *p + x;
}

// variant 0: chance-sampled CFR (cfr.py:158-255); variant 1: external sampling (external_cfr.py:82-173)
// The walk keeps a stack frame only for the decision nodes whose actions are ALL explored (external sampling: the
// traverser's own nodes; chance sampling: every player node).  Sampled nodes (chance, and the opponent under external
// sampling) pass their child's value up unchanged, so they need no frame: the walk simply moves on to the drawn child.
// Leaves are valued where they are met.  The walk alternates two phases per warp.  Phase 1: every thread advances its
// own walk until it stands on a node where an action has to be DRAWN (or the traversal is over) -- cheap bookkeeping,
// divergent.  Phase 2: the threads that need a draw perform it TOGETHER (Philox + numpy's sampler with its divisions
// is the longest straight piece of the walk).
template <int VARIANT, int MAXA, int DEPTH>
__global__ void __launch_bounds__(kBlock) k_lanes(LaneView v, int64_t lanes, int64_t lane_base) {
    int64_t tid = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const uint32_t lane = (uint32_t)((tid >> 1) + lane_base);     // global lane id: results do not depend on the sharding
    const int me = (int)(tid & 1) + 1;                  // traversing player
    const uint64_t iteration = (uint64_t)v.iter->t;
    const double w = v.iter->linear ? (double)v.iter->t : 1.0;
    uint32_t ndraw = 0;
    unsigned long long touched = 0;
    typedef Frame<VARIANT, MAXA> F;
    F st[DEPTH];
    enum { ENTER = 0, RET = 1, NEXT = 2 };
    int sp = -1, cur = 0, mode = ENTER;
    bool done = tid >= 2 * lanes;                       // idle threads keep running: the warp votes below need every lane
    double pr0 = 1.0, pr1 = 1.0, pr2 = 1.0;             // reach of `cur` (chance-sampled variant)
    if (!done) touched++;
    double ret = 0.0;                                   // value delivered to the top frame in mode RET
    auto payoff = [&](int node) { return me == 1 ? v.pay1[node] : (v.zero_sum ? -v.pay1[node] : v.pay2[node]); };
    for (;;) {
        // ---- phase 1: walk until a draw is needed ----------------------------------------------------------
        bool need = false;
        int d_fc = 0, d_n = 0, d_pl = 0, d_col = 0;
        while (!done) {
            if (mode == ENTER) {
                const int4 rec = __ldg(v.rec + cur);        // {first child, #children | (player + 1) << 8, colbase, srank}
                const int pl = (rec.y >> 8) - 1;
                if (pl < 0) { ret = payoff(cur); mode = RET; continue; }
                if (pl == 0 || (VARIANT == 1 && pl != me)) {
                    need = true; d_fc = rec.x; d_n = rec.y & 0xff; d_pl = pl; d_col = rec.z;
                    break;
                }
                F &f = st[++sp];                            // a node whose actions are all explored
                f.node = cur; f.k = 0; f.value = 0.0;
                if (VARIANT == 0) { f.reach[0] = pr0; f.reach[1] = pr1; f.reach[2] = pr2; }
                mode = NEXT;
                continue;
            }
            if (mode == RET) {                              // `ret` is the value of the child the top frame visited last
                if (sp < 0) { done = true; break; }
                F &f = st[sp];
                const int a = f.k - 1;
                const double sga = v.sigma[__ldg(v.rec + f.node).z + a];
                f.va[a] = ret;
                f.value += sga * ret;
                mode = NEXT;
                continue;
            }
            // NEXT: the top frame visits its next child, or finishes
            F &f = st[sp];
            const int4 rec = __ldg(v.rec + f.node);
            const int pl = (rec.y >> 8) - 1, fc = rec.x, n = rec.y & 0xff;
            const double *sg = v.sigma + rec.z;
            if (f.k < n) {
                const int a = f.k++;
                const int child = fc + a;
                touched++;
                if (v.player[child] < 0) {                  // a leaf: valued in place
                    const double u = payoff(child);
                    f.va[a] = u;
                    f.value += sg[a] * u;
                    continue;
                }
                cur = child;
                if (VARIANT == 0) {
                    pr0 = f.reach[0];
                    pr1 = pl == 1 ? sg[a] * f.reach[1] : f.reach[1];
                    pr2 = pl == 2 ? sg[a] * f.reach[2] : f.reach[2];
                }
                mode = ENTER;
                continue;
            }
            // all children done: update (cfr.py:229-252 / external_cfr.py:143-157)
            const int s = rec.w;
            if (VARIANT == 0) {
                v.in_R[s] = 1;
                if (pl == me) {
                    double pi_minus = me == 2 ? f.reach[0] * f.reach[1] : f.reach[0] * f.reach[2];
                    double pi_own = me == 1 ? f.reach[1] : f.reach[2];
                    double *R = v.R + rec.z, *S = v.S + rec.z;
                    for (int a = 0; a < n; ++a) {
                        add(R + a, w * (f.va[a] - f.value) * pi_minus, v.atomic);
                        add(S + a, w * pi_own * sg[a], v.atomic);
                    }
                    v.in_S[s] = 1;
                    v.in_next[s] = 1;
                }
            } else {
                double *R = v.R + rec.z;
                for (int a = 0; a < n; ++a) add(R + a, f.va[a] - f.value, v.atomic);
                v.in_R[s] = 1;
                v.in_next[s] = 1;
            }
            ret = f.value;
            --sp;
            mode = RET;
        }
        if (!__any_sync(0xffffffffu, need)) break;          // every lane of the warp has finished its traversal
        // ---- phase 2: the draws of the warp, together ---------------------------------------------------------
        if (need) {
            const double *p = d_pl == 0 ? v.cprob + d_fc : v.sigma + d_col;
            const int a = draw(p, d_n, rlp::philox_uniform(v.seed, iteration, lane, (uint32_t)(me - 1), ndraw++));
            const int child = d_fc + a;
            touched++;
            if (v.player[child] < 0) { ret = payoff(child); mode = RET; }     // the drawn child is a leaf
            else { cur = child; mode = ENTER; }                             // (its reach is that of the sampled node)
        }
    }
    if (touched) atomicAdd(&v.iter->touched, touched);
}

__global__ void k_set_iter2(IterState *it, long long t, int linear) {
    it->t = t;
    it->linear = linear;
}

// sigma <- regret matching of R for every information set; advance the iteration counter.
__global__ void __launch_bounds__(kBlock)
k_rm_all(int32_t I, const int64_t *__restrict__ col_off, const double *__restrict__ R, double *sigma, IterState *iter) {
    int s = blockIdx.x * kBlock + threadIdx.x;
    if (s == 0) iter->t += 1;
    if (s >= I) return;
    int64_t c0 = col_off[s];
    rlp::regret_matching(R + c0, (int)(col_off[s + 1] - c0), sigma + c0);
}

// ---- AverageStrategy walk (cfr_util.py:109-198) as a forward level sweep ------------------------------
struct WalkView {
    const int32_t *first_child, *colbase, *srank, *krank;
    const int8_t *player;
    const double *sigma;
    const uint8_t *present;     // strategy membership (in_next)
    const int64_t *koff_iset;
    double *alive, *p1, *p2, *contrib_o;
    uint8_t *in_alpha;
    double weight;
};

__global__ void __launch_bounds__(kBlock) k_walk_level(WalkView v, int64_t lo, int64_t hi) {
    int64_t i = lo + (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= hi) return;
    int pl = v.player[i];
    if (pl < 0) return;
    int fc = v.first_child[i], n = v.first_child[i + 1] - fc;
    bool alive = v.alive[i] != 0.0;
    double a = v.p1[i], b = v.p2[i];
    const double *sg = nullptr;
    if (pl > 0) {
        int s = v.srank[i];
        bool ok = alive && v.present[s];
        // pruned / unreached nodes add +0.0, which leaves alpha unchanged bit for bit
        v.contrib_o[v.koff_iset[v.krank[i]] + s] = ok ? (pl == 1 ? a : b) * v.weight : 0.0;
        if (ok) v.in_alpha[s] = 1;
        alive = ok;
        sg = v.sigma + v.colbase[i];
    }
    for (int k = 0; k < n; ++k) {
        int j = fc + k;
        if (v.player[j] < 0) continue;
        v.alive[j] = alive ? 1.0 : 0.0;
        v.p1[j] = pl == 1 ? a * sg[k] : a;
        v.p2[j] = pl == 2 ? b * sg[k] : b;
    }
}

__global__ void __launch_bounds__(kBlock)
k_walk_accumulate(int32_t I, const int32_t *__restrict__ iset_size, const int64_t *__restrict__ col_off,
                  const int64_t *__restrict__ koff_iset, const double *__restrict__ contrib_o,
                  const double *__restrict__ sigma, double *alpha) {
    int s = blockIdx.x * kBlock + threadIdx.x;
    if (s >= I) return;
    int size = iset_size[s];
    int64_t c0 = col_off[s];
    int na = (int)(col_off[s + 1] - c0);
    for (int a = 0; a < na; ++a) {
        double acc = alpha[c0 + a], sg = sigma[c0 + a];
        for (int k = 0; k < size; ++k) acc = acc + sg * contrib_o[koff_iset[k] + s];
        alpha[c0 + a] = acc;
    }
}
}  // namespace

namespace rlp {
int launch_average_walk(rlp_cfr *c, double weight, cudaStream_t s) {
    rlp_tree *t = c->tree;
    WalkView v;
    v.first_child = t->first_child.p; v.colbase = t->colbase.p; v.srank = t->srank.p; v.krank = t->krank.p;
    v.player = t->player.p; v.sigma = c->sigma.p; v.present = c->flags.p + 2 * (size_t)t->I;
    v.koff_iset = t->koff_iset.p; v.alive = c->pc.p; v.p1 = c->p1.p; v.p2 = c->p2.p; v.contrib_o = c->contrib_o.p;
    v.in_alpha = c->flags.p + (size_t)t->I; v.weight = weight;
    // root: alive, both reaches 1.0 -- pc[0]/p1[0]/p2[0] hold 1.0 permanently (rlp_cfr_create)
    for (int l = 0; l + 1 < t->L; ++l) {
        int64_t lo = t->level_off[l], hi = t->level_off[l + 1];
        k_walk_level<<<grid_for(hi - lo), kBlock, 0, s>>>(v, lo, hi); RLP_LAUNCH_CHECK();
    }
    if (t->I > 0) {
        k_walk_accumulate<<<grid_for(t->I), kBlock, 0, s>>>(t->I, t->iset_size.p, t->col_off.p, t->koff_iset.p,
                                                          c->contrib_o.p, c->sigma.p, c->S.p);
        RLP_LAUNCH_CHECK();
    }
    return RLP_OK;
}
}  // namespace rlp

extern "C" int rlp_cfr_update_average(rlp_cfr *c, double weight, void *stream_) {
    if (!c) return fail(RLP_ERR_INVALID, "null solver");
    c->external_state = true;
    return rlp::launch_average_walk(c, weight, (cudaStream_t)stream_);
}

namespace {
// Pick the smallest frame layout the game fits (the stack is per thread, in local memory).
template <int VARIANT>
int launch_lanes(const LaneView &v, rlp_tree *t, int64_t lanes, int64_t lane_base, cudaStream_t s) {
    const unsigned g = grid_for(2 * lanes);
    if (t->max_act <= 4 && t->L <= 16) k_lanes<VARIANT, 4, 16><<<g, kBlock, 0, s>>>(v, lanes, lane_base);
    else if (t->max_act <= 4) k_lanes<VARIANT, 4, kMaxDepth><<<g, kBlock, 0, s>>>(v, lanes, lane_base);
    else k_lanes<VARIANT, kMaxAct, kMaxDepth><<<g, kBlock, 0, s>>>(v, lanes, lane_base);
    RLP_LAUNCH_CHECK();
    return RLP_OK;
}

int check_lanes(rlp_cfr *c, int variant, int64_t lanes) {
    if (variant != RLP_CHANCE_SAMPLED && variant != RLP_EXTERNAL_SAMPLED) return fail(RLP_ERR_INVALID, "variant %d", variant);
    rlp_tree *t = c->tree;
    if (t->L > kMaxDepth) return fail(RLP_ERR_UNSUPPORTED, "tree deeper than %d levels", kMaxDepth);
    if (t->max_act > kMaxAct) return fail(RLP_ERR_UNSUPPORTED, "more than %d actions in an information set", kMaxAct);
    if (t->max_chance > kMaxChance) return fail(RLP_ERR_UNSUPPORTED, "chance node wider than %d", kMaxChance);
    if (lanes > (int64_t)1 << 31) return fail(RLP_ERR_INVALID, "too many lanes");
    return RLP_OK;
}

LaneView lane_view(rlp_cfr *c, uint64_t seed, double *R, double *S, int atomic) {
    rlp_tree *t = c->tree;
    LaneView v;
    v.rec = t->node_rec.p;
    v.first_child = t->first_child.p; v.colbase = t->colbase.p; v.srank = t->srank.p; v.player = t->player.p;
    v.cprob = t->cprob.p; v.pay1 = t->pay1.p; v.pay2 = t->pay2.p; v.sigma = c->sigma.p; v.R = R; v.S = S;
    v.in_R = c->flags.p; v.in_S = c->flags.p + (size_t)t->I; v.in_next = c->flags.p + 2 * (size_t)t->I;
    v.iter = c->iter.p; v.seed = seed; v.zero_sum = t->zero_sum ? 1 : 0; v.atomic = atomic;
    return v;
}

__global__ void __launch_bounds__(kBlock) k_add_delta(int64_t Q, const double *__restrict__ delta, double *R, double *S) {
    int64_t q = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (q >= Q) return;
    R[q] = R[q] + delta[q];
    S[q] = S[q] + delta[Q + q];
}
}  // namespace

extern "C" int rlp_cfr_sampled_iters(rlp_cfr *c, int variant, int64_t lanes, uint64_t seed, int64_t t0, int64_t n_iters,
                                     int linear_weight, void *stream_) {
    if (!c || lanes < 1 || n_iters < 0) return fail(RLP_ERR_INVALID, "bad argument");
    { int rc = check_lanes(c, variant, lanes); if (rc) return rc; }
    rlp_tree *t = c->tree;
    cudaStream_t s = (cudaStream_t)stream_;
    LaneView v = lane_view(c, seed, c->R.p, c->S.p, lanes > 1);
    k_set_iter2<<<1, 1, 0, s>>>(c->iter.p, (long long)t0, linear_weight);     // keeps the running touched count
    RLP_LAUNCH_CHECK();
    if (variant == RLP_EXTERNAL_SAMPLED) c->external_state = true;
    for (int64_t it = 0; it < n_iters; ++it) {
        int rc = variant == RLP_CHANCE_SAMPLED ? launch_lanes<0>(v, t, lanes, 0, s) : launch_lanes<1>(v, t, lanes, 0, s);
        if (rc) return rc;
        if (t->I > 0) { k_rm_all<<<grid_for(t->I), kBlock, 0, s>>>(t->I, t->col_off.p, c->R.p, c->sigma.p, c->iter.p); RLP_LAUNCH_CHECK(); }
        if (variant == RLP_EXTERNAL_SAMPLED) {
            rc = rlp::launch_average_walk(c, 1.0, s);
            if (rc) return rc;
        }
    }
    c->touched_dirty = true;
    c->sigma_am_dirty = true;
    c->sigF_dirty = true;
    c->dev_t = -1;
    return RLP_OK;
}

// ---- sampled lanes sharded over ranks ("shard lanes"; one exchange per iteration) -----------------------------
// Every rank holds the full tree and identical regrets / sums / strategy.  Per iteration: rlp_cfr_sampled_local runs
// this rank's lanes [lane_base, lane_base + lanes) -- the Philox counter carries the GLOBAL lane id, so the draws do not
// depend on the number of ranks -- into the zeroed [2Q] delta buffer (rlp_cfr_delta_ptr: regret deltas | sum deltas,
// internal pair order, the same on every rank) and sets the dict-membership flags (rlp_cfr_flags_ptr, [3 I] bytes);
// the caller sum-all-reduces the delta and max-all-reduces the flags; rlp_cfr_sampled_finish adds the delta, runs regret
// matching and, for external sampling, the average-strategy walk (replicated).
extern "C" int rlp_cfr_sampled_local(rlp_cfr *c, int variant, int64_t lane_base, int64_t lanes, uint64_t seed, int64_t t,
                                     int linear_weight, void *stream_) {
    if (!c || lanes < 0 || lane_base < 0) return fail(RLP_ERR_INVALID, "bad argument");
    { int rc = check_lanes(c, variant, lane_base + lanes); if (rc) return rc; }
    rlp_tree *tr = c->tree;
    cudaStream_t s = (cudaStream_t)stream_;
    RLP_CUDA(cudaMemsetAsync(c->delta.p, 0, c->delta.bytes(), s));
    k_set_iter2<<<1, 1, 0, s>>>(c->iter.p, (long long)t, linear_weight);
    RLP_LAUNCH_CHECK();
    if (lanes > 0) {
        LaneView v = lane_view(c, seed, c->delta.p, c->delta.p + tr->Q, 1);
        int rc = variant == RLP_CHANCE_SAMPLED ? launch_lanes<0>(v, tr, lanes, lane_base, s) : launch_lanes<1>(v, tr, lanes, lane_base, s);
        if (rc) return rc;
    }
    c->touched_dirty = true;
    return RLP_OK;
}

extern "C" int rlp_cfr_sampled_finish(rlp_cfr *c, int variant, void *stream_) {
    if (!c) return fail(RLP_ERR_INVALID, "null solver");
    if (variant != RLP_CHANCE_SAMPLED && variant != RLP_EXTERNAL_SAMPLED) return fail(RLP_ERR_INVALID, "variant %d", variant);
    rlp_tree *t = c->tree;
    cudaStream_t s = (cudaStream_t)stream_;
    if (t->Q > 0) { k_add_delta<<<grid_for(t->Q), kBlock, 0, s>>>(t->Q, c->delta.p, c->R.p, c->S.p); RLP_LAUNCH_CHECK(); }
    if (t->I > 0) { k_rm_all<<<grid_for(t->I), kBlock, 0, s>>>(t->I, t->col_off.p, c->R.p, c->sigma.p, c->iter.p); RLP_LAUNCH_CHECK(); }
    if (variant == RLP_EXTERNAL_SAMPLED) {
        c->external_state = true;
        int rc = rlp::launch_average_walk(c, 1.0, s);
        if (rc) return rc;
    }
    c->sigma_am_dirty = true;
    c->sigF_dirty = true;
    c->dev_t = -1;
    return RLP_OK;
}

extern "C" void *rlp_cfr_flags_ptr(rlp_cfr *c) { return c ? (void *)c->flags.p : nullptr; }
