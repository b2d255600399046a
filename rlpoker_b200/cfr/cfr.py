"""Counterfactual regret minimisation: the reference's driver API (rlpoker/cfr/cfr.py:64-152) on top of the
CUDA level sweep.

Both call shapes of the reference are accepted:

    cfr(experiment, game, num_iters=10000, use_chance_sampling=True, linear_weight=False, experiment_writer=False)
    cfr(game, num_iters=..., use_chance_sampling=...)          # README / tests/test_best_response.py

and the same things happen at the same cadence: every 10 iterations the exploitability of the average
strategy and the immediate regret are printed / logged and the best strategy so far is pickled.  The
per-iteration tree walk of the reference (cfr_recursive, cfr.py:158-255) is rlp_cfr_vanilla_iters or, with
chance sampling, rlp_cfr_sampled_iters (one Philox-driven lane per player and iteration; ``lanes`` > 1 runs a
mini-batch of traversals against each frozen strategy).
"""
import pickle
import time
from collections.abc import Sequence

import numpy as np

from rlpoker_b200 import _lib
from rlpoker_b200._lib import RLP_AVERAGE, RLP_CHANCE_SAMPLED, RLP_SIGMA, RLP_STRAT_SUM
from rlpoker_b200.experiment import Experiment, SolverCheckpoint, StrategySaver
from rlpoker_b200.extensive_game import ActionFloat, Strategy
from rlpoker_b200.flatten import array_to_strategy_dict, flatten, postorder_infoset_order
from rlpoker_b200.solver import DeviceCFR, device_tree


class CFRStrategySaver(StrategySaver):
    """Pickles the strategy object (cfr.py:17-29)."""

    def _save_strategy(self, strategy, file_name):
        with open(file_name, 'wb') as f:
            pickle.dump(strategy, f)

    def _load_strategy(self, file_name):
        with open(file_name, 'rb') as f:
            return pickle.load(f)


def compute_average_strategy(action_counts):
    """counts / sum(counts) per information set, skipping sets whose total is not positive (cfr.py:32-41)."""
    average = {}
    for info_set, counts in action_counts.items():
        total = sum(counts.values())
        if total > 0:
            average[info_set] = {a: float(c) / float(total) for a, c in counts.items()}
    return Strategy(average)


def compare_strategies(s1, s2):
    """Mean (over common information sets) root-mean-square difference of the action distributions."""
    common = [k for k in s1.keys() if k in s2.keys()]
    return np.mean([np.sqrt(np.mean([float(s1[k][a] - s2[k][a]) ** 2 for a in s1[k]])) for k in common])


class DeviceStrategy(Strategy):
    """A Strategy whose probabilities live in a flat [Q] array (canonical pair order); the dict view is built
    on first use.  compute_exploitability / compute_best_response take the array directly (no marshalling)."""

    def __init__(self, flat_tree, sigma, present=None, plain_dicts=False):
        self._flat = flat_tree
        self.sigma_array = np.ascontiguousarray(sigma, np.float64)
        self._present = present
        self._plain = plain_dicts
        self._dict = None

    @property
    def strategy(self):
        if self._dict is None:
            self._dict = array_to_strategy_dict(self._flat, self.sigma_array, self._present,
                                                wrap=None if self._plain else ActionFloat)
        return self._dict

    @strategy.setter
    def strategy(self, value):
        self._dict = value

    def copy(self):
        return Strategy({k: (v.copy() if hasattr(v, 'copy') else dict(v)) for k, v in self.strategy.items()})

    def __reduce__(self):
        return (Strategy, (self.strategy,))


class StrategyHistory(Sequence):
    """The reference returns one Strategy copy per iteration (cfr.py:121-122,152); this is the same sequence
    kept as arrays and materialised per index.  Also carries the immediate-regret series when the driver
    computed it, so cfr_metrics.compute_immediate_regret(game, history) costs nothing."""

    def __init__(self, flat_tree, limit_bytes=1 << 30):
        self._flat = flat_tree
        self._sigmas, self._present = [], []
        self._limit = limit_bytes
        self._count = 0
        self._base = 0              # iterations before a resume point: counted, but their strategies are not here
        self.immediate_regrets = []

    def _append(self, sigma, present):
        self._count += 1
        if (len(self._sigmas) + 1) * sigma.nbytes <= self._limit:
            self._sigmas.append(sigma)
            self._present.append(present)

    def __len__(self):
        return self._count

    def __getitem__(self, index):
        if isinstance(index, slice):
            return [self[i] for i in range(*index.indices(self._count))]
        if index < 0:
            index += self._count
        if not 0 <= index < self._count:
            raise IndexError(index)
        if index < self._base:
            raise IndexError("strategy %d predates the resume point (iteration %d); only later strategies are held"
                             % (index, self._base))
        if index - self._base >= len(self._sigmas):
            raise IndexError("strategy %d was not kept (history limited to %d bytes; pass keep_strategies=...)"
                             % (index, self._limit))
        return DeviceStrategy(self._flat, self._sigmas[index - self._base], self._present[index - self._base])


def _unpack(args, kwargs):
    names = ['experiment', 'game', 'num_iters', 'use_chance_sampling', 'linear_weight', 'experiment_writer']
    defaults = {'experiment': None, 'num_iters': 10000, 'use_chance_sampling': True, 'linear_weight': False,
                'experiment_writer': False}
    args = list(args)
    legacy = bool(args) and not isinstance(args[0], Experiment) and args[0] is not None and \
        (hasattr(args[0], 'root') or hasattr(args[0], '_flat_tree'))
    if legacy or ('game' in kwargs and 'experiment' not in kwargs and not args):
        names = names[1:]
    bound = dict(defaults)
    if len(args) > len(names):
        raise TypeError("cfr() takes at most %d positional arguments" % len(names))
    bound.update(zip(names, args))
    for k in list(kwargs):
        if k in defaults or k == 'game':
            if k in names[:len(args)]:
                raise TypeError("cfr() got multiple values for argument %r" % k)
            bound[k] = kwargs.pop(k)
    if 'game' not in bound:
        raise TypeError("cfr() missing required argument: 'game'")
    return bound, kwargs


def cfr(*args, **kwargs):
    """Returns (average_strategy, exploitabilities, strategies) like the reference.

    Extra keyword arguments (all optional): ``seed`` (Philox seed for chance sampling; default drawn from
    numpy's global RNG so ``np.random.seed`` makes runs repeatable), ``lanes`` (sampled traversal pairs per
    iteration), ``immediate_regret`` (default True: maintain and report cfr_metrics' immediate regret),
    ``keep_strategies`` (byte budget of the returned per-iteration history, default 1 GiB), ``verbose``,
    ``checkpoint`` (save the solver state into the experiment directory at every evaluation) and ``resume`` (continue
    from that state; ``num_iters`` stays the total iteration count).
    """
    bound, extra = _unpack(args, kwargs)
    experiment, game = bound['experiment'], bound['game']
    num_iters, sampling = int(bound['num_iters']), bool(bound['use_chance_sampling'])
    linear_weight, writer = bool(bound['linear_weight']), bound['experiment_writer']
    seed = extra.pop('seed', None)
    lanes = int(extra.pop('lanes', 1))
    want_imm = bool(extra.pop('immediate_regret', True))
    keep = extra.pop('keep_strategies', 1 << 30)
    verbose = bool(extra.pop('verbose', True))
    checkpoint = bool(extra.pop('checkpoint', False))
    resume = bool(extra.pop('resume', False))
    if extra:
        raise TypeError("cfr() got unexpected keyword arguments %s" % sorted(extra))
    ft = flatten(game)
    solver = DeviceCFR(device_tree(game))
    ckpt = SolverCheckpoint(experiment) if experiment is not None and (checkpoint or resume) else None
    resumed = {}
    if resume and ckpt is not None and ckpt.exists():
        saved_seed, resumed = ckpt.load(solver)
        seed = saved_seed if seed is None else seed
    if sampling and seed is None:
        seed = int(np.random.randint(0, 2 ** 31 - 1))
    saver = CFRStrategySaver(experiment=experiment) if experiment is not None else None
    if saver is not None and resumed.get('best_exploitability') is not None:      # never overwrite a better best_strategy.pkl
        saver._best_exploitability = float(resumed['best_exploitability'])
        saver._best_exploitability_t = int(resumed.get('best_exploitability_t', 0))
    order = postorder_infoset_order(ft) if want_imm else None
    history = StrategyHistory(ft, limit_bytes=int(keep) if keep else 0)
    if want_imm and solver.t > 0:
        series = resumed.get('immediate_regrets')
        if series is None or len(series) != solver.t:
            raise ValueError("cannot resume with immediate_regret=True: the checkpoint does not hold the immediate-regret "
                             "series of its %d iterations (it was written with immediate_regret=False)" % solver.t)
        history.immediate_regrets = [float(x) for x in series]
    exploitabilities = []
    average_strategy = None
    every_iter = want_imm or bool(keep)
    all_present = np.ones(ft.num_infosets, bool)

    def advance(n):
        if sampling:
            solver.sampled(RLP_CHANCE_SAMPLED, n, lanes=lanes, seed=seed, linear_weight=linear_weight)
        else:
            solver.vanilla(n, linear_weight=linear_weight)

    def current_average():
        S = solver.read(RLP_STRAT_SUM)
        present = np.add.reduceat(S, ft.act_off[:-1]) > 0 if ft.num_infosets else np.zeros(0, bool)
        return DeviceStrategy(ft, solver.read(RLP_AVERAGE), present, plain_dicts=True)

    start = time.time()
    t = solver.t
    history._count = history._base = t         # iterations before a resume point are counted, not held
    while t < num_iters:
        n = 1 if every_iter else min(num_iters - t, 10 - t % 10 if t % 10 else 1)
        advance(n)
        t_last = t + n - 1                       # the reference's loop variable after these iterations
        if every_iter:
            if keep:
                present = (solver.flags() & 4) != 0 if sampling else all_present
                history._append(solver.read(RLP_SIGMA), present)
            else:
                history._count += 1
            if want_imm:
                history.immediate_regrets.append(solver.immediate_regret_step(t_last, order))
        else:
            history._count += n
        if t_last % 10 == 0:
            if verbose:
                print("t: {}. Time since last evaluation: {:.4f} s".format(t_last, time.time() - start))
            start = time.time()
            exploitability = solver.exploitability(RLP_AVERAGE)[0]
            exploitabilities.append((t_last, exploitability))
            if verbose:
                print("t: {}, nodes touched: {}, exploitability: {} mbb/h".format(t_last, solver.nodes_touched,
                                                                                  exploitability * 1000))
            log_data = {"nodes_touched": solver.nodes_touched, "exploitability_mbbh": exploitability * 1000}
            if want_imm:
                log_data["immediate_regret"] = history.immediate_regrets[-1]
                if verbose:
                    print("Immediate regret: {}".format(history.immediate_regrets[-1]))
            if writer:
                writer.log(data=log_data, global_step=t_last)
            if saver is not None:
                average_strategy = current_average()
                saver.save_best_strategy(average_strategy, t_last, exploitability)
            if checkpoint and ckpt is not None:
                ckpt.save(solver, seed=seed, use_chance_sampling=sampling, linear_weight=linear_weight,
                          immediate_regrets=history.immediate_regrets if want_imm else None,
                          best_exploitability=saver._best_exploitability if saver is not None else None,
                          best_exploitability_t=saver._best_exploitability_t if saver is not None else None)
        t += n
    if num_iters > 0:
        average_strategy = current_average()
    return average_strategy, exploitabilities, history


def evaluate_strategies(game, strategy, num_iters=500):
    return game.expected_value(strategy, strategy, num_iters)


# ---------------------------------------------------------------------------------------------------
# The reference's recursion entry points under their own names (cfr.py:158, external_cfr.py:82).  The device has no
# node-by-node recursion to offer: these run ONE traversal of the whole tree for one player on the device and write
# the result into the caller's dictionaries exactly as the reference's recursion leaves them, so code that drives
# the reference's inner loop by hand (`for i in [1, 2]: cfr_recursive(game, game.root, i, t, 1.0, 1.0, 1.0, ...)`)
# keeps working.  Only the root call is supported (a call on an inner node raises).


def _shim_solver(game):
    solver = getattr(game, '_shim_solver', None)
    if solver is None:
        solver = DeviceCFR(device_tree(game))
        try:
            game._shim_solver = solver
        except AttributeError:
            pass
    return solver


def _dict_to_array(ft, d, default):
    out = np.empty(ft.num_pairs, np.float64)
    keys = ft.info_set_keys
    for i in range(ft.num_infosets):
        lo, k = int(ft.act_off[i]), int(ft.nact[i])
        entry = d[keys[i]] if keys[i] in d else None
        if entry is None:
            out[lo:lo + k] = default(k)
        else:
            out[lo:lo + k] = [entry[a] for a in ft.actions_of(i)]
    return out


def cfr_recursive(game, node, i, t, pi_c, pi_1, pi_2, regrets, action_counts, strategy_t, strategy_t_1, cfr_state,
                  use_chance_sampling=False, weight=1.0):
    """One traversal for player ``i`` against the frozen ``strategy_t`` (cfr.py:158-255): updates ``regrets``,
    ``action_counts`` and ``strategy_t_1`` for player i's information sets, fills ``strategy_t`` lazily with uniform
    entries, counts the nodes and returns player i's value of the root."""
    from rlpoker_b200._lib import RLP_REGRET
    from rlpoker_b200.cfr.cfr_metrics import node_values
    if node is not game.root or (pi_c, pi_1, pi_2) != (1.0, 1.0, 1.0):
        raise ValueError("cfr_recursive is only available as a whole-tree traversal: call it on game.root with unit reaches")
    if use_chance_sampling:
        raise NotImplementedError("chance-sampled traversals run as device lanes: use cfr(..., use_chance_sampling=True)")
    if weight not in (1.0, float(t)):
        raise ValueError("weight must be 1.0 or t (cfr.py:106)")
    ft = flatten(game)
    keys = ft.info_set_keys
    for idx in range(ft.num_infosets):                      # lazily-uniform strategy_t (cfr.py:198-199)
        if keys[idx] not in strategy_t:
            strategy_t[keys[idx]] = ActionFloat.initialise_uniform(ft.actions_of(idx))
    sigma = _dict_to_array(ft, strategy_t, lambda k: 1.0 / float(k))
    R0 = _dict_to_array(ft, regrets, lambda k: 0.0)
    S0 = _dict_to_array(ft, action_counts, lambda k: 0.0)
    solver = _shim_solver(game)
    solver.reset()
    solver.write(RLP_REGRET, R0)
    solver.write(RLP_STRAT_SUM, S0)
    solver.write(RLP_SIGMA, sigma)
    solver.t = int(t)
    solver.vanilla(1, linear_weight=(weight != 1.0))
    R1, S1, sig1 = solver.read(RLP_REGRET), solver.read(RLP_STRAT_SUM), solver.read(RLP_SIGMA)
    for idx in range(ft.num_infosets):
        lo, k = int(ft.act_off[idx]), int(ft.nact[idx])
        acts = ft.actions_of(idx)
        key = keys[idx]
        if key not in regrets:                              # every visited information set gets a regret entry (cfr.py:229-230)
            regrets[key] = ActionFloat.initialise_zero(acts)
        if int(ft.iset_player[idx]) != i:
            continue
        regrets[key] = ActionFloat({a: float(R1[lo + j]) for j, a in enumerate(acts)})
        action_counts[key] = ActionFloat({a: float(S1[lo + j]) for j, a in enumerate(acts)})
        strategy_t_1[key] = ActionFloat({a: float(sig1[lo + j]) for j, a in enumerate(acts)})
    cfr_state.nodes_touched += ft.num_nodes
    v1, v2 = node_values(game, DeviceStrategy(ft, sigma), DeviceStrategy(ft, sigma))
    return float(v1[0] if i == 1 else v2[0])
