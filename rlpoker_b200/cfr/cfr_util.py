"""CFR utilities with the reference's API (rlpoker/cfr/cfr_util.py:5-198).

``compute_regret_matching`` / ``normalise_probs`` are the scalar host versions (handy for single information
sets and pinned by the reference's own unit tests); the solver runs the same arithmetic for all information
sets at once in csrc/common.cuh::regret_matching.  ``AverageStrategy`` keeps alpha in a device array and
``update_average_strategy`` is the forward level sweep of csrc/sampled.cu."""
import numpy as np

from rlpoker_b200 import extensive_game
from rlpoker_b200._lib import RLP_AVERAGE, RLP_SIGMA, RLP_STRAT_SUM
from rlpoker_b200.flatten import array_to_strategy_dict, flatten, strategy_to_array


class CFRState:
    def __init__(self):
        self.nodes_touched = 0

    def node_touched(self):
        self.nodes_touched += 1


def normalise_probs(probs, epsilon=1e-7):
    """Floor every probability at epsilon, then divide by the (builtin, compensated) sum."""
    assert min(probs.values()) >= 0.0
    floored = {a: max(p, epsilon) for a, p in probs.items()}
    total = sum(floored.values())
    return extensive_game.ActionFloat({a: p / total for a, p in floored.items()})


def compute_regret_matching(action_regrets, epsilon=1e-7, highest_regret=False):
    """Play proportionally to positive regret (floored at epsilon); uniform when none is positive, or, with
    ``highest_regret``, (almost) all mass on the largest regret."""
    if max(action_regrets.values()) <= 0.0:
        if not highest_regret:
            return extensive_game.ActionFloat.initialise_uniform(action_regrets.action_list)
        probs = dict.fromkeys(action_regrets, epsilon)
        probs[max(action_regrets, key=action_regrets.get)] = 1.0
        return normalise_probs(extensive_game.ActionFloat(probs), epsilon=epsilon)
    positive = {a: max(0.0, r) for a, r in action_regrets.items()}
    return normalise_probs(extensive_game.ActionFloat(positive), epsilon=epsilon)


class AverageStrategy:
    """alpha[I][a] = sum_t sum_{h in I} pi_i^{sigma^t}(h) sigma^t(I, a), held on the device."""

    def __init__(self, game, solver=None):
        from rlpoker_b200.solver import DeviceCFR, device_tree
        self.game = game
        self._flat = flatten(game)
        self._solver = solver if solver is not None else DeviceCFR(device_tree(game))

    @property
    def alpha(self):
        """{information set: ActionFloat} for the sets updated so far."""
        present = (self._solver.flags() & 2) != 0
        return array_to_strategy_dict(self._flat, self._solver.read(RLP_STRAT_SUM), present,
                                      wrap=extensive_game.ActionFloat)

    def compute_strategy(self):
        present = (self._solver.flags() & 2) != 0
        return extensive_game.Strategy(array_to_strategy_dict(self._flat, self._solver.read(RLP_AVERAGE), present,
                                                              wrap=extensive_game.ActionFloat))


def update_average_strategy(game, average_strategy, strategy, weight=1.0):
    """alpha += walk(strategy): full-tree, pruned below information sets missing from ``strategy``."""
    import ctypes as C
    from rlpoker_b200 import _lib
    from rlpoker_b200.solver import _stream_ptr
    solver = average_strategy._solver
    sigma, present = strategy_to_array(flatten(game), strategy)
    solver.write(RLP_SIGMA, sigma)
    solver.write_flags(present, bit=4)
    _lib.check(_lib.load().rlp_cfr_update_average(solver.handle, C.c_double(weight), _stream_ptr(None)))
    return average_strategy


def update_average_strategy_recursive(game, node, average_strategy, strategy, pi1=1.0, pi2=1.0, pic=1.0, weight=1.0):
    """The reference's recursion entry (cfr_util.py:129-198); available as the whole-tree walk only."""
    if node is not game.root or (pi1, pi2) != (1.0, 1.0):
        raise ValueError("update_average_strategy_recursive is only available on game.root with unit reaches")
    update_average_strategy(game, average_strategy, strategy, weight=weight)
