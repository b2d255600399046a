"""CFR metrics with the reference's API (rlpoker/cfr/cfr_metrics.py).

``compute_immediate_regret`` (cfr_metrics.py:13-78) runs one device sweep per strategy
(rlp_cfr_immediate_regret_detail) instead of the reference's Python recursion; when given the StrategyHistory a driver
returned it reuses the series the driver already computed.  ``compute_expected_utility`` (cfr_metrics.py:159-254) is a
backward value sweep on the device (rlp_node_values) read out per node.
"""
import collections
from collections.abc import Sequence

import numpy as np

from rlpoker_b200 import _lib
from rlpoker_b200._lib import RLP_SIGMA
from rlpoker_b200.flatten import flatten, postorder_infoset_order, strategy_to_array
from rlpoker_b200.solver import DeviceCFR, device_tree


def _sigma_of(ft, strategy):
    arr = getattr(strategy, 'sigma_array', None)
    return arr if arr is not None else strategy_to_array(ft, strategy)[0]


class _ImmediateRegretDetail(Sequence):
    """The reference's ``all_immediate_regrets``: element t maps information set -> 1/(1+t) max_a cumulative regret
    (cfr_metrics.py:68-73).  Held as one [I] array per step; the dictionaries (keyed by information-set tuples, in the
    order the reference's dicts have) are built per index on demand."""

    def __init__(self, ft, order, arrays=None, recompute=None):
        self._ft, self._order, self._arrays, self._recompute = ft, order, arrays, recompute

    def _ensure(self):
        if self._arrays is None:
            self._arrays = self._recompute()
        return self._arrays

    def __len__(self):
        return len(self._ensure())

    def array(self, t):
        """[I] values of step t in canonical information-set order."""
        return self._ensure()[t]

    def __getitem__(self, t):
        if isinstance(t, slice):
            return [self[i] for i in range(*t.indices(len(self)))]
        values = self._ensure()[t]
        keys = self._ft.info_set_keys
        return {keys[i]: float(values[i]) for i in self._order}


def _run_steps(game, ft, strategies, order, first_step=0):
    solver = DeviceCFR(device_tree(game))
    series, detail = [], []
    for step, strategy in enumerate(strategies, start=first_step):
        solver.write(RLP_SIGMA, _sigma_of(ft, strategy))
        value, per_iset = solver.immediate_regret_step(step, order, detail=True)
        series.append(value)
        detail.append(per_iset)
    return series, detail


def compute_immediate_regret(game, strategies):
    """(immediate regret at the last step, [cumulative immediate regret per step], [per-step {information set:
    immediate regret}]) like cfr_metrics.py:13-78.  The third value is a lazy sequence of dictionaries."""
    ft = flatten(game)
    order = postorder_infoset_order(ft)
    series = getattr(strategies, 'immediate_regrets', None)
    if series is not None and len(series) == len(strategies) and len(series) > 0:
        def recompute():
            if getattr(strategies, '_base', 0) != 0:
                raise ValueError("per-information-set immediate regrets need every strategy since iteration 0; this "
                                 "history starts at a resume point")
            return _run_steps(game, ft, [strategies[i] for i in range(len(strategies))], order)[1]
        return series[-1], list(series), _ImmediateRegretDetail(ft, order, recompute=recompute)
    out, detail = _run_steps(game, ft, strategies, order)
    return out[-1], out, _ImmediateRegretDetail(ft, order, arrays=detail)


def node_values(game, sigma_1, sigma_2):
    """(v1[N], v2[N]): value of every node (breadth-first order) for both players when player 1 plays sigma_1 and
    player 2 plays sigma_2 (dict-like strategies; missing information sets are uniform)."""
    ft = flatten(game)
    s1, s2 = _sigma_of(ft, sigma_1), _sigma_of(ft, sigma_2)
    mine = np.repeat(ft.iset_player == 1, ft.nact)
    sigma = np.where(mine, s1, s2)
    tree = device_tree(game)
    v1, v2 = np.empty(ft.num_nodes), np.empty(ft.num_nodes)
    _lib.check(_lib.load().rlp_node_values(tree.handle, _lib.ptr(np.ascontiguousarray(sigma), _lib._dp),
                                           _lib.ptr(v1, _lib._dp), _lib.ptr(v2, _lib._dp), None))
    return v1, v2


def _nodes_breadth_first(game):
    nodes = [game.root]
    queue = collections.deque(nodes)
    while queue:
        node = queue.popleft()
        for child in node.children.values():
            nodes.append(child)
            queue.append(child)
    return nodes


def compute_expected_utility(game, sigma_1, sigma_2):
    """(expected_utility_1, expected_utility_2): defaultdicts mapping player-1 nodes to v_1 and player-2 nodes to v_2
    (cfr_metrics.py:159-180).  Games that exist only as arrays (FlatGame) have no node objects: use ``node_values``."""
    if not hasattr(game, 'root'):
        raise TypeError("compute_expected_utility needs a game with node objects; use node_values() for a FlatGame")
    ft = flatten(game)
    v1, v2 = node_values(game, sigma_1, sigma_2)
    eu1, eu2 = collections.defaultdict(float), collections.defaultdict(float)
    for i, node in enumerate(_nodes_breadth_first(game)):
        if ft.player[i] == 1:
            eu1[node] += float(v1[i])
        elif ft.player[i] == 2:
            eu2[node] += float(v2[i])
    return eu1, eu2


def compute_expected_utility_recursive(game, node, sigma_1, sigma_2, expected_utility_1, expected_utility_2,
                                       pi_1=1.0, pi_2=1.0, pi_c=1.0):
    """(v_1, v_2) of ``node`` and, as in cfr_metrics.py:183-254, the values of the player nodes below it added into
    the two dictionaries.  The reach arguments are accepted for signature compatibility; like in the reference they do
    not enter the values."""
    ft = flatten(game)
    v1, v2 = node_values(game, sigma_1, sigma_2)
    nodes = _nodes_breadth_first(game)
    index = {id(n): i for i, n in enumerate(nodes)}
    start = index[id(node)]
    stack = [start]
    while stack:
        i = stack.pop()
        if ft.player[i] == 1:
            expected_utility_1[nodes[i]] += float(v1[i])
        elif ft.player[i] == 2:
            expected_utility_2[nodes[i]] += float(v2[i])
        stack.extend(range(int(ft.first_child[i]), int(ft.first_child[i + 1])))
    return float(v1[start]), float(v2[start])
