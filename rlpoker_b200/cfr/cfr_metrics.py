"""CFR metrics with the reference's API (rlpoker/cfr/cfr_metrics.py:13-254).

``compute_immediate_regret`` runs one device sweep per strategy (rlp_cfr_immediate_regret_step) instead of the
reference's Python recursion; when given the StrategyHistory a driver returned it reuses the series the
driver already computed."""
import collections

import numpy as np

from rlpoker_b200._lib import RLP_SIGMA
from rlpoker_b200.flatten import flatten, postorder_infoset_order, strategy_to_array
from rlpoker_b200.solver import DeviceCFR, device_tree


def compute_immediate_regret(game, strategies):
    """(immediate regret at the last step, [cumulative immediate regret per step], per-step detail or None).

    The third element of the reference (a dict per step and information set) is produced only on request via
    ``compute_immediate_regret_detail``; here it is None."""
    series = getattr(strategies, 'immediate_regrets', None)
    if series is not None and len(series) == len(strategies) and len(series) > 0:
        return series[-1], list(series), None
    ft = flatten(game)
    solver = DeviceCFR(device_tree(game))
    order = postorder_infoset_order(ft)
    out = []
    for step, strategy in enumerate(strategies):
        arr = getattr(strategy, 'sigma_array', None)
        if arr is None:
            arr = strategy_to_array(ft, strategy)[0]
        solver.write(RLP_SIGMA, arr)
        out.append(solver.immediate_regret_step(step, order))
    return out[-1], out, None


def compute_expected_utility(game, sigma_1, sigma_2):
    """(expected_utility_1, expected_utility_2): value of every player-1 node to player 1 and of every player-2
    node to player 2 under (sigma_1, sigma_2) (cfr_metrics.py:159-254).  Host recursion over node objects: this
    is an inspection helper keyed by node objects, not part of the solve."""
    eu1, eu2 = collections.defaultdict(float), collections.defaultdict(float)

    def walk(node):
        if node.player == -1:
            return node.utility[1], node.utility[2]
        v1 = v2 = 0.0
        if node.player == 0:
            for action, child in node.children.items():
                a1, a2 = walk(child)
                p = node.chance_probs[action]
                v1 += p * a1
                v2 += p * a2
            return v1, v2
        info_set = game.info_set_ids[node]
        probs = (sigma_1 if node.player == 1 else sigma_2)[info_set]
        for action, child in node.children.items():
            a1, a2 = walk(child)
            v1 += probs[action] * a1
            v2 += probs[action] * a2
        if node.player == 1:
            eu1[node] += v1
        else:
            eu2[node] += v2
        return v1, v2

    walk(game.root)
    return eu1, eu2
