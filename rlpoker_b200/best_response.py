"""Best response and exploitability (reference API: rlpoker/best_response.py:103-141), computed by the
device level sweep (csrc/br.cu) through rlp_best_response."""
import numpy as np

from rlpoker_b200.flatten import flatten, strategy_to_array
from rlpoker_b200.solver import device_tree


def _sigma_of(game, strategy):
    """Complete strategy array for ``strategy`` (a Strategy / dict, or a DeviceStrategy that already is one)."""
    arr = getattr(strategy, 'sigma_array', None)
    if arr is not None:
        return arr
    return strategy_to_array(flatten(game), strategy)[0]            # uniform where missing (extensive_game.py:425)


def compute_best_response(game, strategy, player):
    """(value, br_strategy): ``player``'s best response to ``strategy`` (completed uniformly).
    br_strategy maps each of the player's information sets to a one-hot dict; ties go to the first action."""
    ft = flatten(game)
    value, argmax = device_tree(game).best_response(_sigma_of(game, strategy), player, want_argmax=True)
    keys = ft.info_set_keys
    br_strategy = {}
    for i in np.flatnonzero(ft.iset_player == player):
        acts = ft.actions_of(i)
        br_strategy[keys[i]] = {a: (1.0 if j == argmax[i] else 0.0) for j, a in enumerate(acts)}
    return value, br_strategy


def compute_exploitability(game, strategy):
    """BR_1 + BR_2 against ``strategy`` completed uniformly (not halved, as in the reference)."""
    dt = device_tree(game)
    sigma = _sigma_of(game, strategy)
    return dt.best_response(sigma, 1) + dt.best_response(sigma, 2)
