"""CFR metrics with the reference's API (rlpoker/cfr/cfr_metrics.py:13-156).

``compute_immediate_regret`` runs one device sweep per strategy (rlp_cfr_immediate_regret_step) instead of the
reference's Python recursion; when given the StrategyHistory a driver returned it reuses the series the
driver already computed.

Not provided: ``compute_expected_utility`` (cfr_metrics.py:159-254), which returns per-node value dictionaries keyed by
node objects -- an inspection helper outside the solve.  The root values it would report are available from the device
through ``rlpoker_b200.solver.DeviceTree.expected_value`` (rlp_expected_value)."""
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
