// Sampled CFR as independent traversal lanes (replaces the recursions of rlpoker/cfr/cfr.py:158-255 with
// use_chance_sampling=True and rlpoker/cfr/external_cfr.py:82-173) plus the full-tree average-strategy
// walk of rlpoker/cfr/cfr_util.py:109-198 as a forward level sweep.
//
// One traversal per (lane, traversing player) against the strategy frozen for the iteration; chance (and, for
// external sampling, opponent) actions are drawn with the reference's sampler (rlpoker/util.py:17-27 ->
// numpy RandomState.choice: normalise, cumsum, renormalise, searchsorted-right) from a counter-based
// Philox-4x32-10 stream keyed (seed; iteration, lane, player, node of the draw), so a CPU run fed the same
// stream reproduces every lane exactly.  lanes == 1 is the reference algorithm bit for bit (plain adds in
// visit order); lanes > 1 is mini-batch MCCFR with fp64 atomics (order of adds not deterministic).
// Two forms of the same traversals: depth first, one thread per traversal with an explicit stack (k_lanes), and --
// for many lanes -- level by level over frontier records (k_front_*); DESIGN.md section 6 has the measurements.
#include "cfr.cuh"
#include <cstring>

using rlp::fail;

namespace {
constexpr int kBlock = 128;
constexpr int kMaxDepth = 32;
constexpr int kMaxAct = 8;        // widest information set the lane kernels handle
constexpr int kMaxChance = 128;   // widest chance node (numpy's pairwise sum is restated up to here)
inline unsigned grid_for(int64_t n, int b = kBlock) { return (unsigned)((n + b - 1) / b); }

struct LaneView {
    const int4 *rec;                       // packed node records (tree.cuh)
    const int8_t *player;
    const double *cprob, *pay1, *pay2, *sigma;
    double *R, *S;                         // lanes == 1: the solver's arrays; else copy 0 of the accumulators below
    int64_t copy_stride;                   // batched lanes add into one of (copy_mask + 1) zeroed copies, picked by the
    unsigned copy_mask;                    //   block index: the 10^6 adds of a level to the few root sets do not serialise
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
__device__ inline int draw_wide(const double *p, int n, double u) {
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

// The same for up to four outcomes, from registers.  IEEE division by exactly 1.0 returns its operand, so the two
// normalisations are skipped when the sums are exactly 1.0 (uniform thirds, most regret-matched strategies): the same
// bits as draw_wide with a fraction of its divisions.
__device__ inline int draw(const double *p, int n, double u) {
    if (n > 4) return draw_wide(p, n, u);
    double q0 = p[0], q1 = n > 1 ? p[1] : 0.0, q2 = n > 2 ? p[2] : 0.0, q3 = n > 3 ? p[3] : 0.0;
    double tot = 0.0 + q0;
    if (n > 1) tot += q1;
    if (n > 2) tot += q2;
    if (n > 3) tot += q3;
    if (tot != 1.0) { q0 = q0 / tot; q1 = q1 / tot; q2 = q2 / tot; q3 = q3 / tot; }
    double last = 0.0 + q0;
    if (n > 1) last += q1;
    if (n > 2) last += q2;
    if (n > 3) last += q3;
    const bool unit = last == 1.0;
    double run = 0.0 + q0;
    if (!((unit ? run : run / last) <= u)) return 0;
    if (n == 1) return 0;
    run += q1;
    if (!((unit ? run : run / last) <= u)) return 1;
    if (n == 2) return 1;
    run += q2;
    if (!((unit ? run : run / last) <= u)) return 2;
    if (n == 3) return 2;
    run += q3;
    if (!((unit ? run : run / last) <= u)) return 3;
    return 3;
}

// One saved frame of the depth-first walk (the frame being worked on is in registers, see k_lanes).  External sampling
// needs no reach probabilities (its regret increment is unweighted, external_cfr.py:143-152), and the action-value array is
// as wide as the game's widest information set, so the local-memory stack stays as small as the game allows.
template <int VARIANT, int MAXA>
struct Frame {
    int node;
    int k;            // next child to visit
    double value;     // running expectation
    double va[MAXA];
    double reach[VARIANT == 0 ? 3 : 1];     // chance-sampled: pc, p1, p2
};

// Dictionary-membership flags only ever go 0 -> 1: test before storing, or every traversal of every iteration rewrites
// the same few cache lines (the stores of 10^6 lanes to one line serialise in its L2 slice).
__device__ inline void mark(uint8_t *flag) {
    if (*flag == 0) *flag = 1;
}

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
// The frame being worked on lives in REGISTERS (its action values are written with unrolled selects, never indexed);
// local memory only holds the frames below it, written when a deeper explored node is entered and read back when that
// node is finished -- a few times per traversal instead of at every child.
template <int VARIANT, int MAXA, int DEPTH, int MINB>
__global__ void __launch_bounds__(kBlock, MINB) k_lanes(LaneView v, int64_t lanes, int64_t lane_base) {
    int64_t tid = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const uint32_t lane = (uint32_t)((tid >> 1) + lane_base);     // global lane id: results do not depend on the sharding
    const int me = (int)(tid & 1) + 1;                  // traversing player
    const uint64_t iteration = (uint64_t)v.iter->t;
    const double w = v.iter->linear ? (double)v.iter->t : 1.0;
    const int64_t copy = (int64_t)(blockIdx.x & v.copy_mask) * v.copy_stride;
    unsigned long long touched = 0;
    typedef Frame<VARIANT, MAXA> F;
    F st[DEPTH];                                        // the frames BELOW the one in registers
    enum { ENTER = 0, RET = 1, NEXT = 2 };
    int sp = 0, cur = 0, mode = ENTER;                  // sp: live frames (the top one in t_*)
    int t_node = 0, t_fc = 0, t_n = 0, t_pl = 0, t_col = 0, t_s = 0, t_k = 0;
    uint32_t t_leaf = 0;                                // leaf mask of the top frame's children (explored nodes: <= kMaxAct)
    double t_val = 0.0, t_va[MAXA], t_r0 = 1.0, t_r1 = 1.0, t_r2 = 1.0;
#pragma unroll
    for (int i = 0; i < MAXA; ++i) t_va[i] = 0.0;
    bool done = tid >= 2 * lanes;                       // idle threads keep running: the warp votes below need every lane
    double pr0 = 1.0, pr1 = 1.0, pr2 = 1.0;             // reach of `cur` (chance-sampled variant)
    if (!done) touched++;
    double ret = 0.0;                                   // value delivered to the top frame in mode RET
    auto payoff = [&](int node) { return me == 1 ? v.pay1[node] : (v.zero_sum ? -v.pay1[node] : v.pay2[node]); };
    auto set_va = [&](int a, double x) {
#pragma unroll
        for (int i = 0; i < MAXA; ++i) if (i == a) t_va[i] = x;
    };
    for (;;) {
        // ---- phase 1: walk until a draw is needed ----------------------------------------------------------
        bool need = false;
        int d_fc = 0, d_n = 0, d_pl = 0, d_col = 0;
        uint32_t d_leaf = 0;
        while (!done) {
            if (mode == ENTER) {
                const int4 rec = __ldg(v.rec + cur);        // {first child, #children | (player + 1) << 8, colbase, srank}
                const int pl = ((rec.y >> 8) & 0xff) - 1;
                if (pl < 0) { ret = payoff(cur); mode = RET; continue; }
                if (pl == 0 || (VARIANT == 1 && pl != me)) {
                    need = true; d_fc = rec.x; d_n = rec.y & 0xff; d_pl = pl; d_col = rec.z; d_leaf = (uint32_t)rec.y >> 16;
                    break;
                }
                if (sp > 0) {                               // a node whose actions are all explored: new top frame
                    F &f = st[sp - 1];
                    f.node = t_node; f.k = t_k; f.value = t_val;
#pragma unroll
                    for (int i = 0; i < MAXA; ++i) f.va[i] = t_va[i];
                    if (VARIANT == 0) { f.reach[0] = t_r0; f.reach[1] = t_r1; f.reach[2] = t_r2; }
                }
                ++sp;
                t_node = cur; t_fc = rec.x; t_n = rec.y & 0xff; t_pl = pl; t_col = rec.z; t_s = rec.w;
                t_leaf = (uint32_t)rec.y >> 16;
                t_k = 0; t_val = 0.0;
                if (VARIANT == 0) { t_r0 = pr0; t_r1 = pr1; t_r2 = pr2; }
                mode = NEXT;
                continue;
            }
            if (mode == RET) {                              // `ret` is the value of the child the top frame visited last
                if (sp == 0) { done = true; break; }
                const int a = t_k - 1;
                const double sga = v.sigma[t_col + a];
                set_va(a, ret);
                t_val += sga * ret;
                mode = NEXT;
                continue;
            }
            // NEXT: the top frame visits its next child, or finishes
            if (t_k < t_n) {
                const int a = t_k++;
                const int child = t_fc + a;
                touched++;
                if ((t_leaf >> a) & 1) {                    // a leaf: valued in place
                    const double u = payoff(child);
                    set_va(a, u);
                    t_val += v.sigma[t_col + a] * u;
                    continue;
                }
                cur = child;
                if (VARIANT == 0) {
                    const double sga = v.sigma[t_col + a];
                    pr0 = t_r0;
                    pr1 = t_pl == 1 ? sga * t_r1 : t_r1;
                    pr2 = t_pl == 2 ? sga * t_r2 : t_r2;
                }
                mode = ENTER;
                continue;
            }
            // all children done: update (cfr.py:229-252 / external_cfr.py:143-157)
            if (VARIANT == 0) {
                mark(v.in_R + t_s);
                if (t_pl == me) {
                    double pi_minus = me == 2 ? t_r0 * t_r1 : t_r0 * t_r2;
                    double pi_own = me == 1 ? t_r1 : t_r2;
                    double *R = v.R + copy + t_col, *S = v.S + copy + t_col;
#pragma unroll
                    for (int a = 0; a < MAXA; ++a)
                        if (a < t_n) {
                            add(R + a, w * (t_va[a] - t_val) * pi_minus, v.atomic);
                            add(S + a, w * pi_own * v.sigma[t_col + a], v.atomic);
                        }
                    mark(v.in_S + t_s);
                    mark(v.in_next + t_s);
                }
            } else {
                double *R = v.R + copy + t_col;
#pragma unroll
                for (int a = 0; a < MAXA; ++a)
                    if (a < t_n) add(R + a, t_va[a] - t_val, v.atomic);
                mark(v.in_R + t_s);
                mark(v.in_next + t_s);
            }
            ret = t_val;
            --sp;
            if (sp > 0) {                                   // the frame below becomes the top frame again
                const F &f = st[sp - 1];
                t_node = f.node; t_k = f.k; t_val = f.value;
#pragma unroll
                for (int i = 0; i < MAXA; ++i) t_va[i] = f.va[i];
                if (VARIANT == 0) { t_r0 = f.reach[0]; t_r1 = f.reach[1]; t_r2 = f.reach[2]; }
                const int4 rec = __ldg(v.rec + t_node);
                t_fc = rec.x; t_n = rec.y & 0xff; t_pl = ((rec.y >> 8) & 0xff) - 1; t_col = rec.z; t_s = rec.w;
                t_leaf = (uint32_t)rec.y >> 16;
            }
            mode = RET;
        }
        if (!__any_sync(0xffffffffu, need)) break;          // every lane of the warp has finished its traversal
        // ---- phase 2: the draws of the warp, together ---------------------------------------------------------
        if (need) {
            const double *p = d_pl == 0 ? v.cprob + d_fc : v.sigma + d_col;
            const int a = draw(p, d_n, rlp::philox_uniform(v.seed, iteration, lane, (uint32_t)(me - 1), (uint32_t)cur));
            const int child = d_fc + a;
            touched++;
            const bool leaf = a < 16 ? (d_leaf >> a) & 1 : v.player[child] < 0;
            if (leaf) { ret = payoff(child); mode = RET; }                  // the drawn child is a leaf
            else { cur = child; mode = ENTER; }                             // (its reach is that of the sampled node)
        }
    }
    if (touched) atomicAdd(&v.iter->touched, touched);
}

// ---- the same traversals, level by level ---------------------------------------------------------------------------
// With many lanes the depth-first kernel above is bound by divergence: the threads of a warp stand at different places of
// their walks.  The frontier form keeps one RECORD per EXPLORED node of every traversal (a node whose actions are all
// followed: the traverser's own nodes under external sampling, every player node under chance sampling), grouped by
// the number of explored ancestors.  Sampled nodes never become records: whoever steps onto one keeps drawing until
// the walk stands on an explored node (-> a record at the next level) or on a terminal node (-> a record that only
// carries the payoff).  A forward pass per level expands the explored nodes (a block reserves its children with one
// atomic on the level counter), a backward pass per level turns children values into node values and regret /
// strategy-sum increments.  Every thread of a launch does the same step of the walk, the values are the depth-first
// kernel's bit for bit (same operands, same order within a node), terminal children of explored nodes never become
// records, and the Philox key (node of the draw) does not depend on the processing order.
constexpr int kFBlock = 256;
constexpr int kTouchedParts = 64;
struct Frontier {
    int32_t *node, *trav, *child;      // node: the explored node, or ~terminal where a sampled chain ended;
                                       // child: first record of the node's non-terminal children at the next level
    double *val, *reach;               // reach: [2][capacity], the players' own reach (chance-sampled variant; the
                                       //   chance factor of a chance-SAMPLED walk stays 1.0 and 1.0 * x == x)
    unsigned int *count;               // [levels + 1] records per level; the last word flags a capacity overflow
    unsigned long long *touched;       // [kTouchedParts] visit counts, folded into IterState::touched at the end
    int64_t off[kMaxDepth + 2];        // first record of each level
    int64_t total;
    int levels;
};

// From `cur`, draw through sampled nodes until the walk stands on an explored or a terminal node; returns that node
// (~node for a terminal one) and counts the nodes stepped onto.
template <int VARIANT>
__device__ __forceinline__ int follow_chain(const LaneView &v, int cur, int me, uint32_t lane, uint64_t iteration, uint32_t &touched) {
    for (;;) {
        const int4 rec = __ldg(v.rec + cur);
        const int pl = ((rec.y >> 8) & 0xff) - 1;
        if (pl < 0) return ~cur;
        if (!(pl == 0 || (VARIANT == 1 && pl != me))) return cur;
        const int fc = rec.x, n = rec.y & 0xff;
        const double *p = pl == 0 ? v.cprob + fc : v.sigma + rec.z;
        cur = fc + draw(p, n, rlp::philox_uniform(v.seed, iteration, lane, (uint32_t)(me - 1), (uint32_t)cur));
        touched++;
    }
}

__device__ __forceinline__ void count_touched(const Frontier &f, uint32_t touched, unsigned int *s_touched) {
#pragma unroll
    for (int d = 16; d; d >>= 1) touched += __shfl_xor_sync(0xffffffffu, touched, d);
    if ((threadIdx.x & 31) == 0) s_touched[threadIdx.x >> 5] = touched;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int tt = 0;
        for (int k = 0; k < kFBlock / 32; ++k) tt += s_touched[k];
        if (tt) atomicAdd(f.touched + (blockIdx.x & (kTouchedParts - 1)), (unsigned long long)tt);
    }
}

template <int VARIANT>
__global__ void __launch_bounds__(kFBlock) k_front_root(LaneView v, Frontier f, int64_t lanes, int64_t lane_base) {
    __shared__ unsigned int s_touched[kFBlock / 32];
    const int64_t tid = (int64_t)blockIdx.x * kFBlock + threadIdx.x;
    if (tid == 0) f.count[0] = (unsigned int)(2 * lanes);
    uint32_t touched = 0;
    if (tid < 2 * lanes) {
        // traversal id = 2 * lane + (traversing player - 1).  Player 1's traversals first, then player 2's: which nodes
        // are explored depends on the traverser, and a block's children stay together, so warps stay of one kind.
        const int tr = (int)(tid < lanes ? 2 * tid : 2 * (tid - lanes) + 1);
        touched = 1;
        f.node[tid] = follow_chain<VARIANT>(v, 0, (tr & 1) + 1, (uint32_t)((tr >> 1) + lane_base), (uint64_t)v.iter->t, touched);
        f.trav[tid] = tr;
        if (VARIANT == 0) { f.reach[tid] = 1.0; f.reach[f.total + tid] = 1.0; }
    }
    count_touched(f, touched, s_touched);
}

__global__ void k_front_touched(Frontier f, IterState *iter) {
    unsigned long long x = f.touched[threadIdx.x] + f.touched[threadIdx.x + 32];
    for (int d = 16; d; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
    if (threadIdx.x == 0) iter->touched += x;
}

template <int VARIANT>
__global__ void __launch_bounds__(kFBlock) k_front_expand(LaneView v, Frontier f, int level, int64_t lane_base) {
    __shared__ int s_new[kFBlock / 32];
    __shared__ unsigned int s_touched[kFBlock / 32], s_base;
    const unsigned int count = f.count[level];
    if ((int64_t)blockIdx.x * kFBlock >= count) return;          // the whole block is past the end
    const int64_t i = (int64_t)blockIdx.x * kFBlock + threadIdx.x;
    const int64_t at = f.off[level] + i, next = f.off[level + 1];
    const int64_t cap = f.off[level + 2] - next;
    int node = -1, tr = 0, fc = 0, n = 0, pl = 0, col = 0, n_new = 0;
    uint32_t mask = 0, touched = 0;
    if (i < count) {
        node = f.node[at]; tr = f.trav[at];
        if (node >= 0) {                                        // (a negative node is the terminal end of a sampled chain)
            const int4 rec = __ldg(v.rec + node);
            fc = rec.x; n = rec.y & 0xff; pl = ((rec.y >> 8) & 0xff) - 1; col = rec.z;
            mask = ~((uint32_t)rec.y >> 16) & ((1u << n) - 1u);  // explored nodes have at most kMaxAct children
            n_new = __popc(mask);
            touched = (uint32_t)n;
        }
    }
    // the block reserves its children records with ONE atomic on the level counter (a counter shared by 10^5 warps
    // would take an atomic per warp at about one per clock)
    int incl = n_new;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int o = __shfl_up_sync(0xffffffffu, incl, d);
        if ((int)(threadIdx.x & 31) >= d) incl += o;
    }
    const int wid = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 31) s_new[wid] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int k = 0; k < kFBlock / 32; ++k) { int x = s_new[k]; s_new[k] = tot; tot += x; }
        s_base = tot ? atomicAdd(f.count + level + 1, (unsigned int)tot) : 0u;
    }
    __syncthreads();
    if (n_new > 0) {
        const int64_t c0 = (int64_t)s_base + s_new[wid] + incl - n_new;
        if (c0 + n_new > cap) { f.count[f.levels] = 1; f.child[at] = 0; n_new = 0; }     // cannot happen: the bounds are exact
        else {
            f.child[at] = (int32_t)c0;
            const int me = (tr & 1) + 1;
            const uint32_t lane = (uint32_t)((tr >> 1) + lane_base);
            const uint64_t iteration = (uint64_t)v.iter->t;
            double r1 = 0.0, r2 = 0.0;
            if (VARIANT == 0) { r1 = f.reach[at]; r2 = f.reach[f.total + at]; }
            int j = 0;
            for (int a = 0; a < n; ++a) {
                if (!((mask >> a) & 1)) continue;
                const int64_t to = next + c0 + j++;
                f.node[to] = follow_chain<VARIANT>(v, fc + a, me, lane, iteration, touched);
                f.trav[to] = tr;
                if (VARIANT == 0) {
                    const double sga = v.sigma[col + a];
                    f.reach[to] = pl == 1 ? sga * r1 : r1;
                    f.reach[f.total + to] = pl == 2 ? sga * r2 : r2;
                }
            }
        }
    }
    count_touched(f, touched, s_touched);
}

template <int VARIANT, int MAXA>
__global__ void __launch_bounds__(kFBlock) k_front_collect(LaneView v, Frontier f, int level) {
    const int64_t i = (int64_t)blockIdx.x * kFBlock + threadIdx.x;
    if (i >= f.count[level]) return;
    const int64_t at = f.off[level] + i, next = f.off[level + 1];
    const int node = f.node[at], tr = f.trav[at];
    const int me = (tr & 1) + 1;
    auto payoff = [&](int x) { return me == 1 ? v.pay1[x] : (v.zero_sum ? -v.pay1[x] : v.pay2[x]); };
    if (node < 0) { f.val[at] = payoff(~node); return; }        // a sampled chain that ended on a terminal node
    const int c0 = f.child[at];
    const int4 rec = __ldg(v.rec + node);
    const int fc = rec.x, n = rec.y & 0xff, pl = ((rec.y >> 8) & 0xff) - 1, col = rec.z, s = rec.w;
    const uint32_t leaf = (uint32_t)rec.y >> 16;
    double va[MAXA], value = 0.0;
    int j = 0;
#pragma unroll
    for (int a = 0; a < MAXA; ++a) {
        va[a] = 0.0;
        if (a < n) {
            va[a] = (leaf >> a) & 1 ? payoff(fc + a) : f.val[next + c0 + j++];
            value += v.sigma[col + a] * va[a];
        }
    }
    f.val[at] = value;
    const double w = v.iter->linear ? (double)v.iter->t : 1.0;
    const int64_t copy = (int64_t)(blockIdx.x & v.copy_mask) * v.copy_stride;
    if (VARIANT == 0) {                                         // cfr.py:229-252
        mark(v.in_R + s);
        if (pl == me) {
            const double r1 = f.reach[at], r2 = f.reach[f.total + at];
            const double pi_minus = me == 2 ? r1 : r2, pi_own = me == 1 ? r1 : r2;
#pragma unroll
            for (int a = 0; a < MAXA; ++a)
                if (a < n) {
                    atomicAdd(v.R + copy + col + a, w * (va[a] - value) * pi_minus);
                    atomicAdd(v.S + copy + col + a, w * pi_own * v.sigma[col + a]);
                }
            mark(v.in_S + s);
            mark(v.in_next + s);
        }
    } else {                                                    // external_cfr.py:143-157
#pragma unroll
        for (int a = 0; a < MAXA; ++a)
            if (a < n) atomicAdd(v.R + copy + col + a, va[a] - value);
        mark(v.in_R + s);
        mark(v.in_next + s);
    }
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

namespace { const std::vector<int64_t> &lane_widths(rlp_tree *t, int variant); }
namespace rlp {
const std::vector<int64_t> &lane_record_bounds(rlp_tree *t, int variant) { return lane_widths(t, variant); }
}  // namespace rlp

extern "C" int rlp_cfr_update_average(rlp_cfr *c, double weight, void *stream_) {
    if (!c) return fail(RLP_ERR_INVALID, "null solver");
    c->external_state = true;
    return rlp::launch_average_walk(c, weight, (cudaStream_t)stream_);
}

namespace {
// Most frontier records per level (= number of explored ancestors) that the two traversals of ONE lane can hold.
// chain(x) = records of a walk that steps onto x: a terminal node leaves one payoff record at the current level; a
// sampled node takes the largest of its children's (exact per level); an explored node leaves one record and its
// non-terminal children start chains one level further down.
const std::vector<int64_t> &lane_widths(rlp_tree *t, int variant) {
    std::vector<int64_t> &w = t->lane_width[variant];
    if (!w.empty()) return w;
    const int L = t->L;
    w.assign((size_t)L + 1, 0);
    const std::vector<int32_t> &fc = t->first_child_h;
    const std::vector<int8_t> &pl = t->player_h;
    const int E = L + 1;                        // explored ancestors < L
    for (int me = 1; me <= 2; ++me) {
        std::vector<int32_t> below, cur;        // [node of the tree level][explored level, relative to the node's own]
        for (int l = L - 1; l >= 0; --l) {
            const int64_t lo = t->level_off[l], hi = t->level_off[l + 1];
            cur.assign((size_t)(hi - lo) * E, 0);
            for (int64_t x = lo; x < hi; ++x) {
                int32_t *row = &cur[(size_t)(x - lo) * E];
                if (pl[x] < 0) { row[0] = 1; continue; }
                const bool explored = variant == 0 ? pl[x] > 0 : pl[x] == me;
                if (explored) row[0] = 1;
                for (int32_t ch = fc[x]; ch < fc[x + 1]; ++ch) {
                    const int32_t *crow = &below[(size_t)(ch - hi) * E];
                    if (explored) {
                        if (pl[ch] < 0) continue;                       // terminal children of explored nodes: no record
                        for (int d = 0; d + 1 < E; ++d) row[d + 1] += crow[d];
                    } else {
                        for (int d = 0; d < E; ++d) row[d] = std::max(row[d], crow[d]);
                    }
                }
            }
            below.swap(cur);
        }
        for (int d = 0; d < E; ++d) w[d] += below[d];
    }
    return w;
}

constexpr int64_t kFrontierMinLanes = 4096;     // below this the depth-first kernel is launch-count cheaper
constexpr size_t kFrontierMaxBytes = (size_t)32 << 30;

// Sizes (and on first use allocates) the frontier records for `lanes` lanes; false: use the depth-first kernel.
bool frontier_for(rlp_cfr *c, int variant, int64_t lanes, Frontier *f) {
    // External sampling: 2.1 ms against 2.9 ms per 2^20 lanes on Leduc-3.  Chance sampling explores every player node:
    // three times the records, whose traffic costs more than the divergence it removes (8.0 against 7.4 ms), so it
    // stays depth first unless RLP_LANES_FRONTIER=all asks otherwise (tests/test_gpu_sampled.py runs it that way).
    static const int mode = [] {
        if (getenv("RLP_LANES_DFS")) return 0;
        const char *e = getenv("RLP_LANES_FRONTIER");
        return e && !strcmp(e, "all") ? 2 : 1;
    }();
    if (mode == 0 || (mode == 1 && variant == 0) || lanes < kFrontierMinLanes || rlp::g_dry_run) return false;
    rlp_tree *t = c->tree;
    const std::vector<int64_t> &w = lane_widths(t, variant);
    int levels = 0;
    int64_t per_lane = 0;
    for (int d = 0; d < (int)w.size() && d < kMaxDepth; ++d) if (w[d] > 0) { levels = d + 1; per_lane += w[d]; }
    if (levels == 0) return false;              // the root is terminal: nothing to record
    for (int d = 0; d < levels; ++d) if (w[d] * lanes >= ((int64_t)1 << 31)) return false;
    const size_t total = (size_t)per_lane * (size_t)lanes;
    const size_t bytes = total * (12 + 8 + (variant == 0 ? 16 : 0));
    if (bytes > kFrontierMaxBytes) return false;
    if (c->fr_lanes < lanes || c->fr_variant != variant) {
        c->fr_lanes = 0;
        c->fr_reach.release();
        if (c->fr_node.alloc(total) != cudaSuccess || c->fr_trav.alloc(total) != cudaSuccess ||
            c->fr_child.alloc(total) != cudaSuccess || c->fr_val.alloc(total) != cudaSuccess ||
            (variant == 0 && c->fr_reach.alloc(2 * total) != cudaSuccess) ||
            c->fr_count.alloc(64 + 2 * kTouchedParts) != cudaSuccess) {
            cudaGetLastError();
            c->fr_node.release(); c->fr_trav.release(); c->fr_child.release(); c->fr_val.release(); c->fr_reach.release();
            return false;                       // not enough memory for the records: the depth-first kernel needs none
        }
        c->fr_lanes = lanes; c->fr_variant = variant;
    }
    // the buffers may be larger than this call needs (fewer lanes than at allocation): lay the levels out for `lanes`
    f->node = c->fr_node.p; f->trav = c->fr_trav.p; f->child = c->fr_child.p; f->val = c->fr_val.p; f->reach = c->fr_reach.p;
    f->count = c->fr_count.p; f->levels = levels;
    f->touched = (unsigned long long *)(c->fr_count.p + 64);     // count words first (kMaxDepth + 2 <= 64), then the visit counts
    int64_t at = 0;
    for (int d = 0; d <= kMaxDepth + 1; ++d) { f->off[d] = at; if (d < levels) at += w[d] * lanes; }
    f->total = (int64_t)(c->fr_node.n);         // stride of the two reach planes
    c->fr_levels = levels;
    return true;
}

template <int VARIANT>
int launch_frontier(const LaneView &v, const Frontier &f, rlp_tree *t, int64_t lanes, int64_t lane_base, cudaStream_t s) {
    RLP_CUDA(cudaMemsetAsync(f.count, 0, sizeof(unsigned int) * (size_t)(64 + 2 * kTouchedParts), s));
    k_front_root<VARIANT><<<grid_for(2 * lanes, kFBlock), kFBlock, 0, s>>>(v, f, lanes, lane_base); RLP_LAUNCH_CHECK();
    for (int l = 0; l < f.levels; ++l) {
        const int64_t cap = f.off[l + 1] - f.off[l];
        k_front_expand<VARIANT><<<grid_for(cap, kFBlock), kFBlock, 0, s>>>(v, f, l, lane_base); RLP_LAUNCH_CHECK();
    }
    for (int l = f.levels - 1; l >= 0; --l) {
        const int64_t cap = f.off[l + 1] - f.off[l];
        if (t->max_act <= 4) k_front_collect<VARIANT, 4><<<grid_for(cap, kFBlock), kFBlock, 0, s>>>(v, f, l);
        else k_front_collect<VARIANT, kMaxAct><<<grid_for(cap, kFBlock), kFBlock, 0, s>>>(v, f, l);
        RLP_LAUNCH_CHECK();
    }
    k_front_touched<<<1, 32, 0, s>>>(f, v.iter); RLP_LAUNCH_CHECK();
    return RLP_OK;
}

// Many lanes: level by level (frontier records).  Few lanes, or no room for the records: depth first, with the smallest
// frame layout the game fits (that stack is per thread, in local memory).
template <int VARIANT>
int launch_lanes(rlp_cfr *c, const LaneView &v, rlp_tree *t, int64_t lanes, int64_t lane_base, cudaStream_t s) {
    Frontier f;
    if (v.atomic && frontier_for(c, VARIANT, lanes, &f)) return launch_frontier<VARIANT>(v, f, t, lanes, lane_base, s);
    const unsigned g = grid_for(2 * lanes);
    if (t->max_act <= 4 && t->L <= 16) k_lanes<VARIANT, 4, 16, 10><<<g, kBlock, 0, s>>>(v, lanes, lane_base);   // 10 CTAs / SM: measured best of 1 / 8 / 10
    else if (t->max_act <= 4) k_lanes<VARIANT, 4, kMaxDepth, 1><<<g, kBlock, 0, s>>>(v, lanes, lane_base);
    else k_lanes<VARIANT, kMaxAct, kMaxDepth, 1><<<g, kBlock, 0, s>>>(v, lanes, lane_base);
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
    v.player = t->player.p;
    v.cprob = t->cprob.p; v.pay1 = t->pay1.p; v.pay2 = t->pay2.p; v.sigma = c->sigma.p; v.R = R; v.S = S;
    v.copy_stride = 0; v.copy_mask = 0;
    v.in_R = c->flags.p; v.in_S = c->flags.p + (size_t)t->I; v.in_next = c->flags.p + 2 * (size_t)t->I;
    v.iter = c->iter.p; v.seed = seed; v.zero_sum = t->zero_sum ? 1 : 0; v.atomic = atomic;
    return v;
}

// Batched lanes: the adds go to `copies` zeroed accumulator copies ([copy][regret increments | sum increments]), folded
// into the target afterwards.  Without the copies the adds of one level to the handful of information sets near the root
// (10^6 lanes -> 3 x 10^6 adds on one cache line) serialise in one L2 slice and dominate the iteration.
int lane_accumulators(rlp_cfr *c, LaneView *v, cudaStream_t s) {
    rlp_tree *t = c->tree;
    const int64_t stride = (2 * t->Q + 15) / 16 * 16;
    int copies = 64;
    while (copies > 1 && (size_t)copies * (size_t)stride * sizeof(double) > ((size_t)64 << 20)) copies >>= 1;
    if (c->lane_acc.n != (size_t)copies * (size_t)stride) RLP_CUDA(c->lane_acc.alloc((size_t)copies * (size_t)stride));
    RLP_CUDA(cudaMemsetAsync(c->lane_acc.p, 0, c->lane_acc.bytes(), s));
    v->R = c->lane_acc.p; v->S = c->lane_acc.p + t->Q;
    v->copy_stride = stride; v->copy_mask = (unsigned)(copies - 1);
    return RLP_OK;
}

__global__ void __launch_bounds__(kBlock)
k_fold_copies(int64_t Q, int copies, int64_t stride, const double *__restrict__ acc, double *R, double *S) {
    int64_t q = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (q >= 2 * Q) return;
    double sum = 0.0;
    for (int k = 0; k < copies; ++k) sum += acc[(int64_t)k * stride + q];
    double *to = q < Q ? R + q : S + (q - Q);
    *to = *to + sum;
}

int fold_accumulators(rlp_cfr *c, const LaneView &v, double *R, double *S, cudaStream_t s) {
    const int64_t Q = c->tree->Q;
    if (Q > 0) {
        k_fold_copies<<<grid_for(2 * Q), kBlock, 0, s>>>(Q, (int)v.copy_mask + 1, v.copy_stride, c->lane_acc.p, R, S);
        RLP_LAUNCH_CHECK();
    }
    return RLP_OK;
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
        if (lanes > 1) { int rc = lane_accumulators(c, &v, s); if (rc) return rc; }
        int rc = variant == RLP_CHANCE_SAMPLED ? launch_lanes<0>(c, v, t, lanes, 0, s) : launch_lanes<1>(c, v, t, lanes, 0, s);
        if (rc) return rc;
        if (lanes > 1) { rc = fold_accumulators(c, v, c->R.p, c->S.p, s); if (rc) return rc; }
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
    if (t >= 0) {           // t < 0: the device-side counter goes on (graph replay)
        k_set_iter2<<<1, 1, 0, s>>>(c->iter.p, (long long)t, linear_weight);
        RLP_LAUNCH_CHECK();
    }
    if (lanes > 0) {
        LaneView v = lane_view(c, seed, c->delta.p, c->delta.p + tr->Q, 1);
        int rc = lane_accumulators(c, &v, s);
        if (rc) return rc;
        rc = variant == RLP_CHANCE_SAMPLED ? launch_lanes<0>(c, v, tr, lanes, lane_base, s) : launch_lanes<1>(c, v, tr, lanes, lane_base, s);
        if (rc) return rc;
        rc = fold_accumulators(c, v, c->delta.p, c->delta.p + tr->Q, s);
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
