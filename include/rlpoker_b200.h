/*
 * rlpoker_b200 -- C ABI of the B200-native tabular CFR solver.
 *
 * The reference (cgnicholls/rlpoker) is pure Python and has no FFI: its boundary for this path is
 * the Python call surface
 *     rlpoker/cfr/cfr.py:64            cfr(experiment, game, num_iters, use_chance_sampling, ...)
 *     rlpoker/cfr/external_cfr.py:14   external_sampling_cfr(experiment, game, num_iters, ...)
 *     rlpoker/best_response.py:103,122 compute_best_response / compute_exploitability
 * which rlpoker_b200/{cfr,best_response}.py keep.  Those Python functions call the entry points
 * below through ctypes (rlpoker_b200/_lib.py); each comment names the reference code the entry
 * point replaces.  INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions: every function returns 0 on success or a negative rlp_status and never throws;
 * rlp_last_error() describes the last failure on the calling thread.  All arrays crossing the ABI
 * are in CANONICAL order: nodes breadth-first (children in insertion order), information sets by
 * first appearance in that order, (information set, action) pairs at act_off[I] + child slot.
 * Unless a parameter is documented as a device pointer, pointers are HOST pointers and the
 * call copies.  The library owns all device memory it allocates.  `stream` is a cudaStream_t
 * (NULL = the legacy default stream); calls are asynchronous with respect to the host unless they
 * return data to host memory.  Handles are not thread-safe; use one host thread per handle.
 * There is no CPU fallback: without a CUDA device every call fails with RLP_ERR_CUDA.
 */
#ifndef RLPOKER_B200_H
#define RLPOKER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    RLP_OK = 0,
    RLP_ERR_INVALID = -1, /* bad argument / malformed tree */
    RLP_ERR_CUDA = -2,    /* CUDA runtime error (message has the detail) */
    RLP_ERR_NOMEM = -3,
    RLP_ERR_UNSUPPORTED = -4
} rlp_status;

typedef struct rlp_tree rlp_tree; /* a flattened game resident in HBM */
typedef struct rlp_cfr rlp_cfr;   /* solver state (regrets, action counts, strategy, scratch) */

/*
 * Flattened game (host arrays; the product of rlpoker_b200.flatten / rlp_leduc_generate).
 * Replaces the object tree of rlpoker/extensive_game.py:196-335 walked by every reference routine.
 */
typedef struct {
    int64_t num_nodes;            /* N */
    int32_t num_infosets;         /* I */
    int32_t num_levels;           /* depth of the tree + 1 */
    const int32_t *first_child;   /* [N+1] children of node i: first_child[i] .. first_child[i+1]-1 */
    const int8_t *player;         /* [N] -1 terminal, 0 chance, 1, 2 */
    const int32_t *infoset;       /* [N] information-set index of decision nodes, -1 otherwise */
    const double *cprob;          /* [N] probability of the chance edge INTO node i (parent is chance) */
    const double *u1;             /* [N] terminal utility of player 1 */
    const double *u2;             /* [N] terminal utility of player 2 */
    const int64_t *act_off;       /* [I+1] pair offsets: pair (I, a) lives at act_off[I] + a */
    const int64_t *level_off;     /* [num_levels+1] node ranges of each depth */
} rlp_tree_desc;

/* ---- library ------------------------------------------------------------------------- */
const char *rlp_last_error(void);
int rlp_version(void);
/* Number of CUDA devices visible; negative status when the runtime is unusable. */
int rlp_device_count(void);
/* Total kernels this library has launched in this process (bench.py reports it). */
uint64_t rlp_kernel_launches(void);

/* ---- game tree ------------------------------------------------------------------------ */
/* Validate, derive the device layout and copy to the current CUDA device. */
int rlp_tree_upload(const rlp_tree_desc *desc, void *stream, rlp_tree **out);
void rlp_tree_free(rlp_tree *tree);
/* Device bytes held by the tree. */
int64_t rlp_tree_bytes(const rlp_tree *tree);
/* Layout facts: which = 0 tiled sweep in use, 1 #upper tiles, 2 #micro subtrees, 3 #micro shapes, 4 #shapes with a
 * compiled straight-line kernel, 5 zero-sum, 6 largest information set, 7 upper-tile block size, 8 its shared memory,
 * 9 upper tiles compiled to straight-line code, 10 group-fused persistent iteration kernel in use, 11 its micro batches
 * per iteration, 12 its upper groups, 13 its grid (co-resident CTAs), 14 public-state components (the sharding unit). */
int64_t rlp_tree_stat(const rlp_tree *tree, int which);

/*
 * Native Leduc generator: writes the same arrays flatten() produces from
 * rlpoker/games/leduc.py:44-278 `Leduc(get_deck(num_values, num_suits))`, without building Python
 * objects.  Call once with all output pointers NULL to get sizes in *num_nodes / *num_infosets,
 * then with buffers of those sizes.  parent/label/hidden are [N]; labels are 0,1,2 for fold, call,
 * bet and 16 + deck index for cards.
 */
int rlp_leduc_generate(int32_t num_values, int32_t num_suits, int32_t max_raises, int32_t raise_amount,
                       int64_t *num_nodes, int32_t *num_infosets,
                       int32_t *parent, int8_t *player, int32_t *infoset, double *cprob,
                       double *u1, double *u2, int32_t *label, uint8_t *hidden);

/* Host-only planning (no GPU needed): runs the validation, layout derivation, two-tier tiling and NVRTC compilation
 * that rlp_tree_upload performs and reports stats16 = {tiled, upper tiles, micro subtrees, micro shapes, nodes and
 * decision nodes and leaves of the largest micro shape, information sets only in micro subtrees, the rest, nodes and
 * non-terminals of the largest upper tile, its shared memory, kernels compiled, kernels that failed to compile,
 * largest information set, zero-sum}. */
int rlp_tree_plan(const rlp_tree_desc *desc, int64_t *stats16);
/* The same dry run, returning also the CUDA C++ source generated for the group-fused persistent iteration kernel
 * (empty string when the tree does not qualify): *len = bytes needed incl. NUL; buf (nullable) gets up to cap bytes. */
int rlp_tree_plan_fused_source(const rlp_tree_desc *desc, int64_t *stats16, char *buf, int64_t cap, int64_t *len);

/*
 * Host only.  Record capacities of the level-by-level sampled lanes (rlp_cfr_sampled_iters with many lanes) for ONE lane,
 * both traversals: out33[e] = most records a lane can hold whose node has e explored ancestors (explored = every action
 * followed: the traverser's own nodes for RLP_EXTERNAL_SAMPLED, every player node for RLP_CHANCE_SAMPLED), zero padded.
 */
int rlp_tree_plan_lane_records(const rlp_tree_desc *desc, int variant, int64_t *out33);

/* ---- solver --------------------------------------------------------------------------- */
typedef enum { RLP_REGRET = 0, RLP_STRAT_SUM = 1, RLP_SIGMA = 2, RLP_AVERAGE = 3, RLP_CUM_IMM_REGRET = 4 } rlp_which;
typedef enum { RLP_CHANCE_SAMPLED = 0, RLP_EXTERNAL_SAMPLED = 1 } rlp_variant;

/* Allocates regret / strategy-sum / strategy [Q] (zero, zero, uniform) and the per-node scratch.
 * Replaces the dict initialisation at rlpoker/cfr/cfr.py:78-99. */
int rlp_cfr_create(rlp_tree *tree, void *stream, rlp_cfr **out);
void rlp_cfr_free(rlp_cfr *cfr);
/* Back to iteration 0 (regrets and sums zero, uniform strategy). */
int rlp_cfr_reset(rlp_cfr *cfr, void *stream);

/*
 * n_iters iterations t0 .. t0+n_iters-1 of full-tree CFR for both players:
 * rlpoker/cfr/cfr.py:105-122 with use_chance_sampling=False (cfr_recursive, cfr.py:158-255, and
 * compute_regret_matching, cfr_util.py:14-58).  weight = t if linear_weight else 1.0 (cfr.py:106).
 * Bit-identical to the reference on one GPU.
 */
int rlp_cfr_vanilla_iters(rlp_cfr *cfr, int64_t t0, int64_t n_iters, int linear_weight, void *stream);

/*
 * Sampled CFR: `lanes` independent traversal pairs (one per player) per iteration against the frozen
 * strategy, Philox-4x32-10 draws keyed by (seed; iteration, lane, player, draw).  lanes == 1 is the
 * reference algorithm (cfr.py:167-175 chance-sampled; external_cfr.py:82-173 external-sampled,
 * followed by the full-tree average-strategy walk of cfr_util.py:109-198).
 */
int rlp_cfr_sampled_iters(rlp_cfr *cfr, int variant, int64_t lanes, uint64_t seed, int64_t t0,
                          int64_t n_iters, int linear_weight, void *stream);

/*
 * Sampled lanes sharded over ranks (every rank holds the whole tree and identical state).  Per iteration:
 * rlp_cfr_sampled_local runs lanes [lane_base, lane_base + lanes) of the iteration -- the Philox counter carries the
 * global lane id, so the draws do not depend on how the lanes are split -- into the zeroed delta buffer
 * (rlp_cfr_delta_ptr: [2Q] doubles, regret deltas | sum deltas, internal pair order = the same on every rank) and sets
 * the dict-membership flags (rlp_cfr_flags_ptr: [3 I] bytes); the caller sum-all-reduces the delta and max-all-reduces the
 * flags (NCCL); rlp_cfr_sampled_finish adds the delta, runs regret matching and, for external sampling, the
 * average-strategy walk.  rlp_nodes_touched counts this rank's lanes only.  t < 0 keeps the iteration counter and
 * weighting flag on the device (rlp_cfr_sampled_finish advances the counter), so that a local / all-reduce / finish
 * sequence captured in a CUDA graph can be replayed for the following iterations.
 */
int rlp_cfr_sampled_local(rlp_cfr *cfr, int variant, int64_t lane_base, int64_t lanes, uint64_t seed, int64_t t,
                          int linear_weight, void *stream);
int rlp_cfr_sampled_finish(rlp_cfr *cfr, int variant, void *stream);
void *rlp_cfr_flags_ptr(rlp_cfr *cfr);

/* AverageStrategy update with the solver's current strategy (cfr_util.py:109-198): full-tree walk, pruned
 * below information sets that have no strategy entry; alpha lives in the RLP_STRAT_SUM array. */
int rlp_cfr_update_average(rlp_cfr *cfr, double weight, void *stream);

/* Copy a [Q] array (canonical order) to / from host memory.  RLP_AVERAGE (read only) is
 * cfr.py:32-41 for vanilla / chance-sampled state and AverageStrategy.compute_strategy
 * (cfr_util.py:95-106) for external-sampled state; absent information sets come out uniform. */
int rlp_cfr_read(rlp_cfr *cfr, int which, double *dst, void *stream);
int rlp_cfr_write(rlp_cfr *cfr, int which, const double *src, void *stream);   /* asynchronous if src is pinned */
/* [I] flags: bit0 information set has a regret entry, bit1 an action-count / alpha entry,
 * bit2 a strategy entry in the reference's dicts (matters for the sampled variants only). */
int rlp_cfr_read_flags(rlp_cfr *cfr, uint8_t *dst, void *stream);
int rlp_cfr_write_flags(rlp_cfr *cfr, const uint8_t *src, void *stream);
/* Measurement aid: runs `reps` (+2 warm-up) vanilla iterations with CUDA events between the stages of the sweep and
 * returns the summed device milliseconds of {upper tiles phase 0, micro subtrees, upper tiles phase 1,
 * accumulate + regret matching} (level-sweep fallback: {0, 0, all level kernels, accumulate}). */
int rlp_cfr_profile_stages(rlp_cfr *cfr, int reps, float *out_ms4, void *stream);
/* Debug aid: runs `iters` vanilla iterations while block 0 of every sweep kernel logs {kind, start ns, end ns}
 * (kind 0 upper phase 0, 1 micro, 2 upper phase 1, 3 accumulate A, 4 accumulate B / all) -- the in-graph schedule. */
int rlp_cfr_timeline(rlp_cfr *cfr, int iters, unsigned long long *out, int cap_records, int *n_out, void *stream);
/* Measurement aid for the group-fused persistent kernel (one cooperative launch per rlp_cfr_vanilla_iters call): runs
 * `iters` unit-weight iterations continuing at t0 while one CTA stamps %globaltimer {iteration start, its micro phase
 * done, grid barrier passed, then inside its upper unit: loads issued, fan-in done, tile evaluated, accumulate + regret
 * matching done, reach written, and its upper phase done}; out_ns receives 9 * iters stamps.  RLP_ERR_UNSUPPORTED when
 * the kernel is not in use. */
int rlp_cfr_fused_timeline(rlp_cfr *cfr, int64_t t0, int iters, unsigned long long *out_ns, void *stream);
/* CFRState.nodes_touched (rlpoker/cfr/cfr_util.py:5-11). */
uint64_t rlp_nodes_touched(rlp_cfr *cfr);

/* Exploitability of the solver's current average (which = RLP_AVERAGE) or current (RLP_SIGMA)
 * strategy without leaving the device: best_response.py:122-141.  out[0..2] = total, BR1, BR2. */
int rlp_cfr_exploitability(rlp_cfr *cfr, int which, double *out3, void *stream);

/* One step of cfr_metrics.compute_immediate_regret (cfr_metrics.py:13-78) for the current strategy:
 * adds this strategy's per-information-set regrets to the running totals and returns
 * sum_I max_a total[I,a] / (step+1), the I terms added in `order` (canonical information-set ids;
 * NULL = canonical order) with the compensated sum CPython's sum() uses. */
int rlp_cfr_immediate_regret_step(rlp_cfr *cfr, int64_t step, const int32_t *order, double *out, void *stream);
/* Same, and the per-information-set terms 1/(1+step) * max_a total[I,a] -- the step-th dictionary of the third value
 * compute_immediate_regret returns (cfr_metrics.py:68-73,78) -- into per_iset[I] (canonical ids; nullable). */
int rlp_cfr_immediate_regret_detail(rlp_cfr *cfr, int64_t step, const int32_t *order, double *out, double *per_iset,
                                    void *stream);

/*
 * Best response of `player` against `sigma` ([Q], host, canonical, complete):
 * rlpoker/best_response.py:103-119.  value_out: best-response value; argmax_out (nullable, [I]):
 * chosen action slot per information set of `player`, -1 elsewhere.
 */
int rlp_best_response(rlp_tree *tree, const double *sigma, int player, double *value_out,
                      int32_t *argmax_out, void *stream);

/* (u1, u2) expected under `sigma` for both players: extensive_game.py:382-423. out2 = {u1, u2}. */
int rlp_expected_value(rlp_tree *tree, const double *sigma, double *out2, void *stream);

/* Value of every node for both players under `sigma` ([Q], host, canonical, complete; player-p rows are used at
 * player-p nodes): cfr_metrics.py:159-254 compute_expected_utility(_recursive).  v1_out / v2_out: host [N] in
 * breadth-first node order (terminals hold their utilities); either may be NULL. */
int rlp_node_values(rlp_tree *tree, const double *sigma, double *v1_out, double *v2_out, void *stream);

/* ---- multi-GPU (deal-sharded trees; one process per GPU) ---------------------------------- */
/*
 * Split iteration for sharded solves: `rlp_cfr_local_sweep` runs the forward/backward sweep over
 * this rank's nodes and leaves the per-pair deltas (regret || strategy-sum, 2*Q doubles, canonical
 * order) in a device buffer whose address `rlp_cfr_delta_ptr` returns; the caller all-reduces
 * that buffer in place (NCCL via torch.distributed) on the same stream; `rlp_cfr_apply_delta`
 * adds it and runs regret matching.  weight as in rlp_cfr_vanilla_iters.  t < 0 keeps the device-side iteration
 * counter (which rlp_cfr_apply_delta advances), so a sweep / all-reduce / apply sequence can be captured once in a
 * CUDA graph and replayed.
 */
int rlp_cfr_local_sweep(rlp_cfr *cfr, int64_t t, int linear_weight, void *stream);
/* The same loop entirely in C: rank 0 calls rlp_comm_unique_id, shares the 128 bytes (e.g. torch.distributed
 * broadcast), every rank calls rlp_comm_init (ncclCommInitRank), then rlp_cfr_sharded_iters runs n iterations of
 * sweep -> ncclAllReduce(sum, in place, caller's stream) -> apply without returning to the host language. */
int rlp_comm_unique_id(unsigned char *out128);
int rlp_comm_init(rlp_cfr *cfr, int rank, int nranks, const unsigned char *id128);
void rlp_comm_destroy(rlp_cfr *cfr);
int rlp_cfr_sharded_iters(rlp_cfr *cfr, int64_t t0, int64_t n_iters, int linear_weight, void *stream);
void *rlp_cfr_delta_ptr(rlp_cfr *cfr);
int rlp_cfr_apply_delta(rlp_cfr *cfr, void *stream);


/*
 * Sharding by PUBLIC STATE (trees that run the group-fused kernel, e.g. every Leduc-n): the micro subtrees fall into
 * components that share no information set (Leduc: one per board card and round-1 betting line); rank r sweeps its
 * components and alone holds their regrets / action counts / strategy.  Per iteration the ranks exchange only the VALUE
 * of every micro subtree (1.1 MB in total at Leduc-13), stored straight into every peer's buffer over NVLink by the
 * micro phase of the one persistent kernel, with the phase's grid barrier doubling as the cross-GPU handshake.  All
 * sums keep the single-GPU order: the result is bit-identical to one GPU and to the reference.
 *   every rank: rlp_cfr_p2p_export -> 64-byte CUDA IPC handle of its exchange buffer; share the handles (e.g.
 *   torch.distributed all_gather); every rank: rlp_cfr_p2p_connect with all handles in rank order (<= 8 ranks on one
 *   node); then rlp_cfr_vanilla_iters with the SAME arguments on every rank.  rlp_cfr_read returns this rank's view:
 *   merge with rlp_cfr_p2p_owned (1 = this rank's entry is the current one; every set is owned by exactly one rank).
 */
int rlp_cfr_p2p_export(rlp_cfr *cfr, unsigned char *handle64);
int rlp_cfr_p2p_connect(rlp_cfr *cfr, int rank, int nranks, const unsigned char *handles);
int rlp_cfr_p2p_owned(rlp_cfr *cfr, unsigned char *owned);
int rlp_cfr_p2p_failed(rlp_cfr *cfr);

#ifdef __cplusplus
}
#endif
#endif /* RLPOKER_B200_H */
