"""cfr_metrics on the device against the reference's own outputs (tests/golden/reference_metrics.*): node values of
compute_expected_utility bit for bit (incl. the exact-equality KATs of rlpoker/cfr/tests/test_cfr_metrics.py:11-99),
and all three return values of compute_immediate_regret."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, make_game

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def metrics():
    with open(os.path.join(GOLDEN_DIR, 'reference_metrics.json')) as f:
        return json.load(f), dict(np.load(os.path.join(GOLDEN_DIR, 'reference_metrics.npz')))


def _strategy(ft, sigma):
    from rlpoker_b200.cfr.cfr import DeviceStrategy
    return DeviceStrategy(ft, sigma)


@pytest.mark.parametrize('name', ['rps', 'ocp4', 'leduc2'])
def test_node_values_bit_exact(metrics, name):
    from rlpoker_b200.cfr.cfr_metrics import compute_expected_utility, node_values
    from rlpoker_b200.flatten import flatten
    J, arr = metrics
    game = make_game(name)
    ft = flatten(game)
    for seed in (0, 1):
        s = _strategy(ft, arr['%s_s%d_sigma' % (name, seed)])
        v1, v2 = node_values(game, s, s)
        assert np.array_equal(v1, arr['%s_s%d_v1' % (name, seed)])
        assert np.array_equal(v2, arr['%s_s%d_v2' % (name, seed)])
        eu1, eu2 = compute_expected_utility(game, s, s)
        assert [len(eu1), len(eu2)] == J['%s_s%d_counts' % (name, seed)]


def test_reference_unit_test_of_expected_utility():
    """rlpoker/cfr/tests/test_cfr_metrics.py, both tests, restated against this package."""
    import collections
    from rlpoker_b200 import extensive_game
    from rlpoker_b200.cfr import cfr_metrics
    from rlpoker_b200.games import rock_paper_scissors
    game = rock_paper_scissors.create_rock_paper_scissors()
    info_set_1 = game.info_set_ids[game.root]
    info_set_2 = game.info_set_ids[game.root.children['R']]
    sigma_1 = extensive_game.Strategy({info_set_1: extensive_game.ActionFloat({'R': 0.5, 'P': 0.3, 'S': 0.2})})
    sigma_2 = extensive_game.Strategy({info_set_2: extensive_game.ActionFloat({'R': 0.3, 'P': 0.3, 'S': 0.4})})
    for a1 in 'RPS':
        for a2 in 'RPS':
            node = game.get_node((a1, a2))
            v1, v2 = cfr_metrics.compute_expected_utility_recursive(
                game, node, sigma_1, sigma_2, collections.defaultdict(float), collections.defaultdict(float), 1.0, 1.0, 1.0)
            assert v1 == node.utility[1] and v2 == node.utility[2]
    v1, v2 = cfr_metrics.compute_expected_utility_recursive(
        game, game.get_node(('R',)), sigma_1, sigma_2, collections.defaultdict(float), collections.defaultdict(float),
        0.5, 1.0, 1.0)
    assert v1 == 0 * 0.3 + -1 * 0.3 + 1 * 0.4 and v2 == 0 * 0.3 + 1 * 0.3 + -1 * 0.4
    eu1, eu2 = cfr_metrics.compute_expected_utility(game, sigma_1, sigma_2)
    root = (0.5 * (0 * 0.3 + -1 * 0.3 + 1 * 0.4) + 0.3 * (1 * 0.3 + 0 * 0.3 + -1 * 0.4) +
            0.2 * (-1 * 0.3 + 1 * 0.3 + 0 * 0.4))
    assert eu1[game.get_node(())] == root
    assert eu2[game.get_node(('R',))] == 0 * 0.3 + 1 * 0.3 + -1 * 0.4
    assert eu2[game.get_node(('P',))] == -1 * 0.3 + 0 * 0.3 + 1 * 0.4
    assert eu2[game.get_node(('S',))] == 1 * 0.3 + -1 * 0.3 + 0 * 0.4


def test_immediate_regret_all_three_returns(metrics):
    from rlpoker_b200.cfr.cfr import cfr
    from rlpoker_b200.cfr.cfr_metrics import compute_immediate_regret
    from rlpoker_b200.flatten import flatten
    J, arr = metrics
    game = make_game('leduc2')
    ft = flatten(game)
    _, _, history = cfr(game, num_iters=5, use_chance_sampling=False, verbose=False)
    for strategies in (history, [history[i] for i in range(5)]):       # the driver's series / a plain list
        last, series, detail = compute_immediate_regret(game, strategies)
        assert series == J['leduc2_imm_series'] and last == J['leduc2_imm_last']
        assert len(detail) == 5
        for step in range(5):
            assert np.array_equal(detail.array(step), arr['leduc2_imm_detail_%d' % step])
        d = detail[4]
        keys = ft.info_set_keys
        assert [repr(k) for k in list(d.keys())[:6]] == J['leduc2_imm_order_head']
        assert all(d[keys[i]] == arr['leduc2_imm_detail_4'][i] for i in range(ft.num_infosets))


def test_expected_value_large_tree_ordered_sum():
    """rlp_expected_value's staged ordered sum at Leduc-5 (151 121 nodes, 148 chunks) equals the oracle bit for bit."""
    from oracle import flat, oracle as orc
    from rlpoker_b200.games.leduc_native import leduc_flat_tree
    from rlpoker_b200.solver import DeviceTree
    ft = leduc_flat_tree(5)
    ot = flat.from_arrays(ft.num_nodes, ft.parent, ft.first_child, ft.player, ft.infoset, ft.label, ft.hidden,
                          ft.cprob, ft.u1, ft.u2)
    o = orc.Oracle(ot)
    rng = np.random.RandomState(3)
    x = rng.random_sample(o.Q) + 0.05
    sigma = x / np.repeat(np.add.reduceat(x, ot.act_off[:-1]), ot.nact)
    dt = DeviceTree(ft)
    assert dt.expected_value(sigma) == o.expected_value_exact(sigma)
    v1, _ = o.node_values(sigma)
    assert abs(dt.expected_value(sigma)[0] - v1[0]) < 1e-12
    dt.close()
