/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Nothing under rlpoker_b200/ may import, link or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs use it, as the checker and the CPU baseline.
 *
 * Plain-C restatement, on flat breadth-first arrays, of the reference's tabular CFR path:
 *   rlpoker/cfr/cfr.py:105-122,158-255      (vanilla + chance-sampled CFR, cfr_recursive)
 *   rlpoker/cfr/cfr_util.py:14-58            (regret matching, epsilon floor)
 *   rlpoker/cfr/cfr.py:32-41                 (average strategy from action counts)
 *   rlpoker/cfr/external_cfr.py:47-56,82-173 (external-sampling MCCFR)
 *   rlpoker/cfr/cfr_util.py:61-198           (AverageStrategy full-tree walk with pruning)
 *   rlpoker/best_response.py:9-141           (best response over information-set node lists)
 *   rlpoker/util.py:17-27                    (sample_action -> numpy RandomState.choice)
 *   rlpoker/cfr/cfr_metrics.py:13-156        (immediate regret)
 *   rlpoker/extensive_game.py:382-423        (expected_value_exact)
 *
 * Parity is PINNED: tests/test_oracle_golden.py checks this file bit-for-bit against
 * the JSON / npz files under tests/golden, which oracle/gen_golden.py produced by running the reference itself.
 *
 * Floating point: compile with -ffp-contract=off (no FMA); the two normaliser sums use
 * Neumaier compensation because CPython >= 3.12's builtin sum() does.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAX_ACT 64

typedef struct {
    int64_t num_nodes;
    int32_t num_infosets;
    int32_t num_levels;
    const int32_t *parent;        /* [N] */
    const int32_t *first_child;   /* [N+1] children of i are first_child[i] .. first_child[i+1]-1 */
    const int8_t *player;         /* [N] -1 terminal, 0 chance, 1, 2 */
    const int32_t *infoset;       /* [N] canonical information-set index or -1 */
    const int32_t *label;         /* [N] id of the action label on the edge into the node */
    const uint8_t *hidden;        /* [N] bit0: actions here hidden from player 1, bit1: from player 2 */
    const double *cprob;          /* [N] chance probability of the edge into the node */
    const double *u1, *u2;        /* [N] terminal utilities */
    const int64_t *act_off;       /* [I+1] */
    const int64_t *iset_node_off; /* [I+1] */
    const int32_t *iset_nodes;    /* nodes of each information set, ascending */
    const int64_t *level_off;     /* [L+1] */
} orc_tree;

/* ------------------------------------------------------------------ sums */

/* CPython >= 3.12 builtin sum() over floats (Neumaier compensation). */
static double neumaier(const double *x, int n) {
    double s = 0.0, c = 0.0;
    for (int i = 0; i < n; ++i) {
        double t = s + x[i];
        if (fabs(s) >= fabs(x[i])) c += (s - t) + x[i]; else c += (x[i] - t) + s;
        s = t;
    }
    return s + c;
}

/* numpy's pairwise float64 add.reduce for a contiguous 1-d array (np.sum). */
static double numpy_sum(const double *a, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res += a[i];
        return res;
    }
    if (n <= 128) {
        double r[8];
        int i;
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return numpy_sum(a, n2) + numpy_sum(a + n2, n - n2);
}

/* ---------------------------------------------------------- regret matching */

/* cfr_util.py:47-58 */
void orc_normalise_probs(const double *p, int n, double eps, double *out) {
    double q[ORC_MAX_ACT];
    for (int a = 0; a < n; ++a) q[a] = p[a] > eps ? p[a] : eps;      /* max(prob, eps) */
    double tot = neumaier(q, n);
    for (int a = 0; a < n; ++a) out[a] = q[a] / tot;
}

/* cfr_util.py:14-44 */
void orc_regret_matching(const double *r, int n, double eps, int highest_regret, double *out) {
    double mx = r[0];
    for (int a = 1; a < n; ++a) if (r[a] > mx) mx = r[a];
    if (mx <= 0.0) {
        if (highest_regret) {
            double q[ORC_MAX_ACT];
            int best = 0;
            for (int a = 1; a < n; ++a) if (r[a] > r[best]) best = a;
            for (int a = 0; a < n; ++a) q[a] = eps;
            q[best] = 1.0;
            orc_normalise_probs(q, n, eps, out);
        } else {
            for (int a = 0; a < n; ++a) out[a] = 1.0 / n;
        }
        return;
    }
    double q[ORC_MAX_ACT];
    for (int a = 0; a < n; ++a) q[a] = r[a] > 0.0 ? r[a] : 0.0;       /* max(0.0, v) */
    orc_normalise_probs(q, n, eps, out);
}

/* cfr.py:32-41 : avg = counts / sum(counts) where the sum is positive; else absent. */
void orc_average_strategy(const orc_tree *t, const double *S, double *avg, uint8_t *present) {
    for (int32_t i = 0; i < t->num_infosets; ++i) {
        int64_t lo = t->act_off[i];
        int n = (int)(t->act_off[i + 1] - lo);
        double tot = neumaier(S + lo, n);
        present[i] = tot > 0.0;
        for (int a = 0; a < n; ++a) avg[lo + a] = tot > 0.0 ? S[lo + a] / tot : 1.0 / n;
    }
}

/* extensive_game.py:67-77 (ActionFloat.normalise) applied to alpha: -1 on a negative entry. */
int orc_normalise_alpha(const orc_tree *t, const double *alpha, double *out) {
    for (int32_t i = 0; i < t->num_infosets; ++i) {
        int64_t lo = t->act_off[i];
        int n = (int)(t->act_off[i + 1] - lo);
        for (int a = 0; a < n; ++a) if (alpha[lo + a] < 0.0) return -1;
        double tot = neumaier(alpha + lo, n);
        for (int a = 0; a < n; ++a) out[lo + a] = tot == 0.0 ? 1.0 / n : alpha[lo + a] / tot;
    }
    return 0;
}

/* ------------------------------------------------------------------ RNG */

typedef struct {
    int kind;               /* 0: consume `stream` in order (numpy MT19937 doubles); 1: Philox */
    const double *stream;
    int64_t stream_len, stream_pos;
    uint64_t seed;
    uint64_t iteration;
    uint32_t lane, trav, draw;
    int exhausted;
} orc_rng;

static inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

/* Philox-4x32-10; counter = (node of the draw, traversing player, lane, iteration), key = seed. */
double orc_philox_uniform(uint64_t seed, uint64_t iteration, uint32_t lane, uint32_t trav, uint32_t draw) {
    uint32_t c[4] = {draw, trav, lane, (uint32_t)iteration};
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(iteration >> 32)};
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
    }
    /* 53-bit double in [0,1) like genrand_res53 */
    uint32_t a = c[0] >> 5, b = c[1] >> 6;
    return ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
}

/* The uniform for the draw at `node`.  kind 0 consumes the recorded stream in visit order (the reference's numpy
 * generator); kind 1 is counter based and keyed by the NODE the draw happens at -- a traversal visits a node at most
 * once, so the key is unique, and it does not depend on the order the traversal's nodes are processed in (the device
 * processes a traversal level by level, not depth first). */
static double next_uniform(orc_rng *g, int32_t node) {
    if (g->kind == 0) {
        if (g->stream_pos >= g->stream_len) { g->exhausted = 1; return 0.0; }
        return g->stream[g->stream_pos++];
    }
    g->draw++;
    return orc_philox_uniform(g->seed, g->iteration, g->lane, g->trav, (uint32_t)node);
}

/* util.py:17-27 -> RandomState.choice(p=...): normalise, cumsum, renormalise, searchsorted right. */
int orc_draw(const double *p, int n, double u) {
    double q[ORC_MAX_ACT * 4];
    double tot = numpy_sum(p, n);
    double run = 0.0;
    for (int a = 0; a < n; ++a) { run += p[a] / tot; q[a] = run; }
    double last = q[n - 1];
    int idx = 0;
    for (int a = 0; a < n; ++a) if (q[a] / last <= u) idx = a + 1; else break;
    return idx < n ? idx : n - 1;
}

/* ------------------------------------------------ faithful recursion (cfr.py:158-255) */

typedef struct {
    const orc_tree *t;
    double *R, *S;             /* [Q] regrets, action counts */
    double *sigma_t;           /* [Q] strategy_t, complete (absent == uniform) */
    double *sigma_next;        /* [Q] strategy_t_1 */
    uint8_t *in_R, *in_S, *in_next;   /* [I] dict-membership flags */
    double weight;
    int sampling;
    orc_rng *rng;
    int64_t touched;
} cfr_ctx;

static double cfr_rec(cfr_ctx *c, int32_t node, int i, double pi_c, double pi_1, double pi_2) {
    const orc_tree *t = c->t;
    c->touched++;
    int pl = t->player[node];
    if (pl == -1) return i == 1 ? t->u1[node] : t->u2[node];
    int32_t fc = t->first_child[node];
    int n = t->first_child[node + 1] - fc;
    if (pl == 0) {
        if (c->sampling) {
            int a = orc_draw(t->cprob + fc, n, next_uniform(c->rng, node));
            return cfr_rec(c, fc + a, i, pi_c, pi_1, pi_2);
        }
        double value = 0.0;
        for (int a = 0; a < n; ++a) {
            double cp = t->cprob[fc + a];
            value += cp * cfr_rec(c, fc + a, i, cp * pi_c, pi_1, pi_2);
        }
        return value;
    }
    int32_t I = t->infoset[node];
    const double *sg = c->sigma_t + t->act_off[I];
    double va[ORC_MAX_ACT];
    double value = 0.0;
    for (int a = 0; a < n; ++a) {
        if (pl == 1) va[a] = cfr_rec(c, fc + a, i, pi_c, sg[a] * pi_1, pi_2);
        else         va[a] = cfr_rec(c, fc + a, i, pi_c, pi_1, sg[a] * pi_2);
        value += sg[a] * va[a];
    }
    c->in_R[I] = 1;
    if (pl == i) {
        double *R = c->R + t->act_off[I], *S = c->S + t->act_off[I];
        double pi_minus = i == 2 ? pi_c * pi_1 : pi_c * pi_2;
        double pi_own = i == 1 ? pi_1 : pi_2;
        for (int a = 0; a < n; ++a) {
            double radd = c->weight * (va[a] - value) * pi_minus;
            double sadd = c->weight * pi_own * sg[a];
            R[a] = R[a] + radd;
            S[a] = S[a] + sadd;
        }
        c->in_S[I] = 1;
        orc_regret_matching(R, n, 1e-7, 0, c->sigma_next + t->act_off[I]);
        c->in_next[I] = 1;
    }
    return value;
}

/*
 * cfr.py:105-122 for iterations t0 .. t0+n_iters-1 by the reference's own recursion.
 * sigma holds strategy_t (complete; pass uniform at t0 == 0).  flags[3*I]: in_R | in_S | in_next.
 * lanes > 1 (sampled only) = that many traversal pairs per iteration against the same frozen
 * strategy_t, accumulated in lane order (mini-batch; lanes == 1 is the reference algorithm).
 * Returns nodes touched, or -1 if the uniform stream ran dry.
 */
int64_t orc_cfr_recursive_iters(const orc_tree *t, double *R, double *S, double *sigma, uint8_t *flags,
                                int64_t t0, int64_t n_iters, int linear_weight, int sampling,
                                int64_t lanes, orc_rng *rng) {
    int64_t Q = t->act_off[t->num_infosets];
    cfr_ctx c;
    c.t = t; c.R = R; c.S = S; c.sigma_t = sigma;
    c.sigma_next = (double *)malloc(sizeof(double) * (Q ? Q : 1));
    memcpy(c.sigma_next, sigma, sizeof(double) * Q);
    c.in_R = flags; c.in_S = flags + t->num_infosets; c.in_next = flags + 2 * (int64_t)t->num_infosets;
    c.sampling = sampling; c.rng = rng; c.touched = 0;
    if (!sampling) lanes = 1;
    for (int64_t it = t0; it < t0 + n_iters; ++it) {
        c.weight = linear_weight ? (double)it : 1.0;
        for (int64_t lane = 0; lane < lanes; ++lane)
            for (int i = 1; i <= 2; ++i) {
                if (rng) { rng->iteration = (uint64_t)it; rng->lane = (uint32_t)lane; rng->trav = (uint32_t)(i - 1); rng->draw = 0; }
                cfr_rec(&c, 0, i, 1.0, 1.0, 1.0);
            }
        memcpy(sigma, c.sigma_next, sizeof(double) * Q);          /* strategy_t = strategy_t_1.copy() */
    }
    free(c.sigma_next);
    if (rng && rng->exhausted) return -1;
    return c.touched;
}

/* ------------------------------------------- level sweep (VANILLA_ITER) */

/*
 * The depth-ordered restatement the CUDA path implements; bit-identical to the recursion above
 * (tests/test_oracle_golden.py proves it).  `threads` > 1 splits every level over a pool of
 * pthreads with a barrier between levels (every node / information set is written by exactly
 * one thread, so results do not depend on the thread count).  This is also the multi-core CPU
 * baseline bench.py times.
 */
#include <pthread.h>

typedef struct {
    const orc_tree *t;
    double *R, *S, *sigma;
    double *pc, *p1, *p2, *v1, *v2;
    int64_t t0, n_iters;
    int linear_weight, threads;
    pthread_barrier_t bar;
} lvl_shared;

typedef struct { lvl_shared *sh; int tid; } lvl_arg;

static void split(int64_t lo, int64_t hi, int tid, int nth, int64_t *a, int64_t *b) {
    int64_t n = hi - lo;
    *a = lo + n * tid / nth;
    *b = lo + n * (tid + 1) / nth;
}

static void *lvl_worker(void *argp) {
    lvl_arg *arg = (lvl_arg *)argp;
    lvl_shared *s = arg->sh;
    const orc_tree *t = s->t;
    int L = t->num_levels, nth = s->threads, tid = arg->tid;
    double *pc = s->pc, *p1 = s->p1, *p2 = s->p2, *v1 = s->v1, *v2 = s->v2;
    double *R = s->R, *S = s->S, *sigma = s->sigma;
    int64_t a, b;
    for (int64_t it = s->t0; it < s->t0 + s->n_iters; ++it) {
        double w = s->linear_weight ? (double)it : 1.0;
        if (tid == 0) pc[0] = p1[0] = p2[0] = 1.0;
        pthread_barrier_wait(&s->bar);
        for (int d = 0; d + 1 < L; ++d) {
            split(t->level_off[d], t->level_off[d + 1], tid, nth, &a, &b);
            for (int64_t i = a; i < b; ++i) {
                int pl = t->player[i];
                if (pl == -1) continue;
                const double *sg = pl > 0 ? sigma + t->act_off[t->infoset[i]] : 0;
                int32_t fc = t->first_child[i];
                int n = t->first_child[i + 1] - fc;
                for (int k = 0; k < n; ++k) {
                    int32_t j = fc + k;
                    double e = pl == 0 ? t->cprob[j] : sg[k];
                    pc[j] = pl == 0 ? e * pc[i] : pc[i];
                    p1[j] = pl == 1 ? e * p1[i] : p1[i];
                    p2[j] = pl == 2 ? e * p2[i] : p2[i];
                }
            }
            pthread_barrier_wait(&s->bar);
        }
        for (int d = L - 1; d >= 0; --d) {
            split(t->level_off[d], t->level_off[d + 1], tid, nth, &a, &b);
            for (int64_t i = a; i < b; ++i) {
                int pl = t->player[i];
                if (pl == -1) { v1[i] = t->u1[i]; v2[i] = t->u2[i]; continue; }
                const double *sg = pl > 0 ? sigma + t->act_off[t->infoset[i]] : 0;
                int32_t fc = t->first_child[i];
                int n = t->first_child[i + 1] - fc;
                double a1 = 0.0, a2 = 0.0;
                for (int k = 0; k < n; ++k) {
                    double e = pl == 0 ? t->cprob[fc + k] : sg[k];
                    a1 = a1 + e * v1[fc + k];
                    a2 = a2 + e * v2[fc + k];
                }
                v1[i] = a1; v2[i] = a2;
            }
            pthread_barrier_wait(&s->bar);
        }
        split(0, t->num_infosets, tid, nth, &a, &b);
        for (int64_t I = a; I < b; ++I) {
            int64_t q = t->act_off[I];
            int n = (int)(t->act_off[I + 1] - q);
            for (int64_t k = t->iset_node_off[I]; k < t->iset_node_off[I + 1]; ++k) {
                int32_t h = t->iset_nodes[k];
                int pl = t->player[h];
                const double *v = pl == 1 ? v1 : v2;
                double opp = pl == 1 ? pc[h] * p2[h] : pc[h] * p1[h];
                double own = pl == 1 ? p1[h] : p2[h];
                int32_t fc = t->first_child[h];
                for (int x = 0; x < n; ++x) {
                    R[q + x] = R[q + x] + w * (v[fc + x] - v[h]) * opp;
                    S[q + x] = S[q + x] + w * own * sigma[q + x];
                }
            }
        }
        pthread_barrier_wait(&s->bar);      /* all S updates read the frozen sigma before it moves */
        for (int64_t I = a; I < b; ++I) {
            int64_t q = t->act_off[I];
            orc_regret_matching(R + q, (int)(t->act_off[I + 1] - q), 1e-7, 0, sigma + q);
        }
        pthread_barrier_wait(&s->bar);
    }
    return 0;
}

int64_t orc_vanilla_level_iters(const orc_tree *t, double *R, double *S, double *sigma,
                                int64_t t0, int64_t n_iters, int linear_weight, int threads) {
    int64_t N = t->num_nodes;
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    lvl_shared sh;
    sh.t = t; sh.R = R; sh.S = S; sh.sigma = sigma; sh.t0 = t0; sh.n_iters = n_iters;
    sh.linear_weight = linear_weight; sh.threads = threads;
    sh.pc = (double *)malloc(sizeof(double) * N); sh.p1 = (double *)malloc(sizeof(double) * N);
    sh.p2 = (double *)malloc(sizeof(double) * N);
    sh.v1 = (double *)malloc(sizeof(double) * N); sh.v2 = (double *)malloc(sizeof(double) * N);
    pthread_barrier_init(&sh.bar, 0, (unsigned)threads);
    pthread_t th[256];
    lvl_arg args[256];
    for (int i = 0; i < threads; ++i) { args[i].sh = &sh; args[i].tid = i; }
    for (int i = 1; i < threads; ++i) pthread_create(&th[i], 0, lvl_worker, &args[i]);
    lvl_worker(&args[0]);
    for (int i = 1; i < threads; ++i) pthread_join(th[i], 0);
    pthread_barrier_destroy(&sh.bar);
    free(sh.pc); free(sh.p1); free(sh.p2); free(sh.v1); free(sh.v2);
    return 2 * N * n_iters;
}

/* ------------------------------------- external sampling (external_cfr.py:82-173) */

static double ext_rec(cfr_ctx *c, int32_t node, int player) {
    const orc_tree *t = c->t;
    c->touched++;
    int pl = t->player[node];
    if (pl == -1) return player == 1 ? t->u1[node] : t->u2[node];
    int32_t fc = t->first_child[node];
    int n = t->first_child[node + 1] - fc;
    if (pl == 0) {
        int a = orc_draw(t->cprob + fc, n, next_uniform(c->rng, node));
        return ext_rec(c, fc + a, player);
    }
    int32_t I = t->infoset[node];
    const double *sg = c->sigma_t + t->act_off[I];
    if (pl == player) {
        double u[ORC_MAX_ACT], eu = 0.0;
        for (int a = 0; a < n; ++a) {
            u[a] = ext_rec(c, fc + a, player);
            eu += sg[a] * u[a];
        }
        double *R = c->R + t->act_off[I];
        for (int a = 0; a < n; ++a) R[a] = R[a] + (u[a] - eu);
        c->in_R[I] = 1;
        orc_regret_matching(R, n, 1e-7, 0, c->sigma_next + t->act_off[I]);
        c->in_next[I] = 1;
        return eu;
    }
    int a = orc_draw(sg, n, next_uniform(c->rng, node));
    return ext_rec(c, fc + a, player);
}

/* cfr_util.py:109-198 with pruning below information sets absent from the strategy. */
static void avg_walk(const orc_tree *t, int32_t node, const double *sigma, const uint8_t *present,
                     double *alpha, uint8_t *in_alpha, double pi1, double pi2, double weight) {
    int pl = t->player[node];
    if (pl == -1) return;
    int32_t fc = t->first_child[node];
    int n = t->first_child[node + 1] - fc;
    if (pl == 0) {
        for (int a = 0; a < n; ++a) avg_walk(t, fc + a, sigma, present, alpha, in_alpha, pi1, pi2, weight);
        return;
    }
    int32_t I = t->infoset[node];
    if (!present[I]) return;
    const double *sg = sigma + t->act_off[I];
    double *al = alpha + t->act_off[I];
    double uw = (pl == 1 ? pi1 : pi2) * weight;
    for (int a = 0; a < n; ++a) al[a] = al[a] + sg[a] * uw;
    in_alpha[I] = 1;
    for (int a = 0; a < n; ++a) {
        if (pl == 1) avg_walk(t, fc + a, sigma, present, alpha, in_alpha, pi1 * sg[a], pi2, weight);
        else         avg_walk(t, fc + a, sigma, present, alpha, in_alpha, pi1, pi2 * sg[a], weight);
    }
}

void orc_update_average_strategy(const orc_tree *t, const double *sigma, const uint8_t *present,
                                 double *alpha, uint8_t *in_alpha, double weight) {
    avg_walk(t, 0, sigma, present, alpha, in_alpha, 1.0, 1.0, weight);
}

/*
 * external_cfr.py:47-56.  flags[3*I]: in_R | in_alpha | in_next (== membership of strategy_t).
 * sigma: strategy_t (complete).  alpha: AverageStrategy.alpha.
 */
int64_t orc_external_iters(const orc_tree *t, double *R, double *alpha, double *sigma, uint8_t *flags,
                           int64_t t0, int64_t n_iters, int64_t lanes, orc_rng *rng) {
    int64_t Q = t->act_off[t->num_infosets];
    cfr_ctx c;
    c.t = t; c.R = R; c.S = 0; c.sigma_t = sigma;
    c.sigma_next = (double *)malloc(sizeof(double) * (Q ? Q : 1));
    memcpy(c.sigma_next, sigma, sizeof(double) * Q);
    c.in_R = flags; c.in_S = flags + t->num_infosets; c.in_next = flags + 2 * (int64_t)t->num_infosets;
    c.sampling = 1; c.rng = rng; c.touched = 0; c.weight = 1.0;
    for (int64_t it = t0; it < t0 + n_iters; ++it) {
        for (int64_t lane = 0; lane < lanes; ++lane)
            for (int p = 1; p <= 2; ++p) {
                rng->iteration = (uint64_t)it; rng->lane = (uint32_t)lane; rng->trav = (uint32_t)(p - 1); rng->draw = 0;
                ext_rec(&c, 0, p);
            }
        memcpy(sigma, c.sigma_next, sizeof(double) * Q);
        avg_walk(t, 0, sigma, c.in_next, alpha, c.in_S, 1.0, 1.0, 1.0);
    }
    free(c.sigma_next);
    if (rng->exhausted) return -1;
    return c.touched;
}

/* ------------------------------------------ best response (best_response.py:9-100) */

typedef struct {
    const orc_tree *t;
    const double *sigma;
    double *reach;       /* [N] */
    int32_t *argmax;     /* [I] or NULL */
    int player;
    int num_labels;
} br_ctx;

static double br_rec(br_ctx *c, const int32_t *set, int64_t n) {
    const orc_tree *t = c->t;
    int32_t first = set[0];
    int pl = t->player[first];
    if (pl == -1) {
        double util = 0.0;
        const double *u = c->player == 1 ? t->u1 : t->u2;
        for (int64_t k = 0; k < n; ++k) util += c->reach[set[k]] * u[set[k]];
        return util;
    }
    if (pl != c->player) {
        /* partition the children by action label, groups in first-appearance order */
        int64_t total = 0;
        for (int64_t k = 0; k < n; ++k) total += t->first_child[set[k] + 1] - t->first_child[set[k]];
        int32_t *group_of_label = (int32_t *)malloc(sizeof(int32_t) * c->num_labels);
        for (int l = 0; l < c->num_labels; ++l) group_of_label[l] = -1;
        int64_t *count = (int64_t *)calloc(c->num_labels + 1, sizeof(int64_t));
        int ngroups = 0;
        for (int64_t k = 0; k < n; ++k)
            for (int32_t j = t->first_child[set[k]]; j < t->first_child[set[k] + 1]; ++j) {
                int l = t->label[j];
                if (group_of_label[l] < 0) group_of_label[l] = ngroups++;
                count[group_of_label[l]]++;
            }
        int64_t *start = (int64_t *)malloc(sizeof(int64_t) * (ngroups + 1));
        start[0] = 0;
        for (int g = 0; g < ngroups; ++g) start[g + 1] = start[g] + count[g];
        int64_t *fill = (int64_t *)malloc(sizeof(int64_t) * (ngroups + 1));
        memcpy(fill, start, sizeof(int64_t) * (ngroups + 1));
        int32_t *kids = (int32_t *)malloc(sizeof(int32_t) * (total ? total : 1));
        for (int64_t k = 0; k < n; ++k) {
            int32_t nd = set[k];
            int32_t fc = t->first_child[nd];
            int npl = t->player[nd];
            for (int32_t j = fc; j < t->first_child[nd + 1]; ++j) {
                double p = npl == 0 ? t->cprob[j] : c->sigma[t->act_off[t->infoset[nd]] + (j - fc)];
                c->reach[j] = c->reach[nd] * p;
                kids[fill[group_of_label[t->label[j]]]++] = j;
            }
        }
        double br_sum = 0.0;
        if ((t->hidden[first] >> (c->player - 1)) & 1) {
            if (total) br_sum += br_rec(c, kids, total);
        } else {
            for (int g = 0; g < ngroups; ++g) br_sum += br_rec(c, kids + start[g], start[g + 1] - start[g]);
        }
        free(group_of_label); free(count); free(start); free(fill); free(kids);
        return br_sum;
    }
    int32_t fc0 = t->first_child[first];
    int na = t->first_child[first + 1] - fc0;
    int32_t *next = (int32_t *)malloc(sizeof(int32_t) * n);
    double best = 0.0;
    int best_a = -1;
    for (int a = 0; a < na; ++a) {
        for (int64_t k = 0; k < n; ++k) {
            next[k] = t->first_child[set[k]] + a;
            c->reach[next[k]] = c->reach[set[k]];
        }
        double v = br_rec(c, next, n);
        if (best_a < 0 || v > best) { best = v; best_a = a; }   /* python max(): first maximum wins */
    }
    if (c->argmax) c->argmax[t->infoset[first]] = best_a;
    free(next);
    return best;
}

/* best_response.py:103-119.  sigma must be complete.  argmax (nullable): best action per
 * information set of `player`, -1 where never reached. */
double orc_best_response(const orc_tree *t, const double *sigma, int player, int num_labels, int32_t *argmax) {
    br_ctx c;
    c.t = t; c.sigma = sigma; c.player = player; c.argmax = argmax; c.num_labels = num_labels;
    c.reach = (double *)malloc(sizeof(double) * t->num_nodes);
    c.reach[0] = 1.0;
    if (argmax) for (int32_t i = 0; i < t->num_infosets; ++i) argmax[i] = -1;
    int32_t root = 0;
    double v = br_rec(&c, &root, 1);
    free(c.reach);
    return v;
}

/* best_response.py:122-141 */
double orc_exploitability(const orc_tree *t, const double *sigma, int num_labels) {
    double e1 = orc_best_response(t, sigma, 1, num_labels, 0);
    double e2 = orc_best_response(t, sigma, 2, num_labels, 0);
    return e1 + e2;
}

/* --------------------------- expected_value_exact (extensive_game.py:382-423) */

void orc_expected_value_exact(const orc_tree *t, const double *sigma, double *out) {
    int64_t N = t->num_nodes;
    double *reach = (double *)malloc(sizeof(double) * N);
    double e1 = 0.0, e2 = 0.0;
    reach[0] = 1.0;
    for (int64_t i = 0; i < N; ++i) {
        int pl = t->player[i];
        if (pl == -1) { e1 += reach[i] * t->u1[i]; e2 += reach[i] * t->u2[i]; continue; }
        int32_t fc = t->first_child[i];
        for (int32_t j = fc; j < t->first_child[i + 1]; ++j) {
            double p = pl == 0 ? t->cprob[j] : sigma[t->act_off[t->infoset[i]] + (j - fc)];
            reach[j] = reach[i] * p;
        }
    }
    free(reach);
    out[0] = e1; out[1] = e2;
}

/* ------------------------------ immediate regret (cfr_metrics.py:13-156) */

static void node_regret_rec(const orc_tree *t, int32_t node, const double *sigma, double *acc,
                            double pi_1, double pi_2, double pi_c, double *o1, double *o2) {
    int pl = t->player[node];
    if (pl == -1) { *o1 = t->u1[node]; *o2 = t->u2[node]; return; }
    int32_t fc = t->first_child[node];
    int n = t->first_child[node + 1] - fc;
    double v1 = 0.0, v2 = 0.0;
    if (pl == 0) {
        for (int a = 0; a < n; ++a) {
            double cp = t->cprob[fc + a], a1, a2;
            node_regret_rec(t, fc + a, sigma, acc, pi_1, pi_2, pi_c * cp, &a1, &a2);
            v1 += cp * a1; v2 += cp * a2;
        }
        *o1 = v1; *o2 = v2;
        return;
    }
    int32_t I = t->infoset[node];
    const double *sg = sigma + t->act_off[I];
    double w1[ORC_MAX_ACT], w2[ORC_MAX_ACT];
    for (int a = 0; a < n; ++a) {
        node_regret_rec(t, fc + a, sigma, acc, pl == 1 ? pi_1 * sg[a] : pi_1, pl == 2 ? pi_2 * sg[a] : pi_2, pi_c,
                        &w1[a], &w2[a]);
        v1 += sg[a] * w1[a]; v2 += sg[a] * w2[a];
    }
    double *dst = acc + t->act_off[I];
    for (int a = 0; a < n; ++a)
        dst[a] += pl == 1 ? pi_c * pi_2 * (w1[a] - v1) : pi_c * pi_1 * (w2[a] - v2);
    *o1 = v1; *o2 = v2;
}

/*
 * One step of compute_immediate_regret's outer loop: this_iter[Q] (zeroed here) receives the
 * per-information-set sums of node regrets under `sigma` (complete), then cumulative += this_iter.
 * Returns sum_I 1/(1+t) * max_a cumulative[I,a] with the builtin-sum (Neumaier) outer sum taken
 * in `order` (the reference's dict order; canonical when NULL).
 */
double orc_immediate_regret_detail(const orc_tree *t, const double *sigma, double *cumulative, double *scratch,
                                   int64_t step, const int32_t *order, double *per_iset);
double orc_immediate_regret_step(const orc_tree *t, const double *sigma, double *cumulative, double *scratch,
                                 int64_t step, const int32_t *order) {
    return orc_immediate_regret_detail(t, sigma, cumulative, scratch, step, order, NULL);
}

/* Same; per_iset[I] (nullable) receives the step's dictionary of cfr_metrics.py:68-73 in canonical order. */
double orc_immediate_regret_detail(const orc_tree *t, const double *sigma, double *cumulative, double *scratch,
                                   int64_t step, const int32_t *order, double *per_iset) {
    int64_t Q = t->act_off[t->num_infosets];
    memset(scratch, 0, sizeof(double) * Q);
    double a1, a2;
    node_regret_rec(t, 0, sigma, scratch, 1.0, 1.0, 1.0, &a1, &a2);
    for (int64_t q = 0; q < Q; ++q) cumulative[q] += scratch[q];
    double s = 0.0, comp = 0.0;
    for (int32_t k = 0; k < t->num_infosets; ++k) {
        int32_t I = order ? order[k] : k;
        const double *cu = cumulative + t->act_off[I];
        int n = (int)(t->act_off[I + 1] - t->act_off[I]);
        double mx = cu[0];
        for (int a = 1; a < n; ++a) if (cu[a] > mx) mx = cu[a];
        double x = 1.0 / (double)(1 + step) * mx;
        if (per_iset) per_iset[I] = x;
        double tt = s + x;
        if (fabs(s) >= fabs(x)) comp += (s - tt) + x; else comp += (x - tt) + s;
        s = tt;
    }
    return s + comp;
}

/* ------------------------------ node values (cfr_metrics.py:159-254) */

/* compute_expected_utility_recursive: v_i(h) = sum over children, in child order from 0.0, of (chance probability or
 * the actor's strategy) * v_i(child); terminals return their utilities.  sigma[Q] holds player-1 rows at player-1
 * information sets and player-2 rows at player-2 ones.  v1 / v2: [N]. */
static void node_value_rec(const orc_tree *t, int32_t node, const double *sigma, double *v1, double *v2) {
    int pl = t->player[node];
    if (pl == -1) { v1[node] = t->u1[node]; v2[node] = t->u2[node]; return; }
    int32_t fc = t->first_child[node];
    int n = t->first_child[node + 1] - fc;
    double a1 = 0.0, a2 = 0.0;
    for (int a = 0; a < n; ++a) {
        node_value_rec(t, fc + a, sigma, v1, v2);
        double p = pl == 0 ? t->cprob[fc + a] : sigma[t->act_off[t->infoset[node]] + a];
        a1 += p * v1[fc + a];
        a2 += p * v2[fc + a];
    }
    v1[node] = a1; v2[node] = a2;
}

void orc_node_values(const orc_tree *t, const double *sigma, double *v1, double *v2) {
    node_value_rec(t, 0, sigma, v1, v2);
}
