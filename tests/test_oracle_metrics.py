"""The oracle's metric restatements (node values, per-information-set immediate regret) bit for bit against what the
reference's cfr_metrics.py produced (tests/golden/reference_metrics.*, oracle/gen_golden_metrics.py)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from oracle import flat, oracle as orc


@pytest.fixture(scope='module')
def metrics():
    with open(os.path.join(GOLDEN_DIR, 'reference_metrics.json')) as f:
        return json.load(f), dict(np.load(os.path.join(GOLDEN_DIR, 'reference_metrics.npz')))


@pytest.mark.parametrize('name', ['rps', 'ocp4', 'leduc2'])
def test_node_values_match_reference(metrics, otrees, name):
    _, arr = metrics
    o = orc.Oracle(otrees(name))
    for seed in (0, 1):
        v1, v2 = o.node_values(arr['%s_s%d_sigma' % (name, seed)])
        assert np.array_equal(v1, arr['%s_s%d_v1' % (name, seed)])
        assert np.array_equal(v2, arr['%s_s%d_v2' % (name, seed)])


def test_rps_kat_of_the_reference_test(metrics, otrees):
    """rlpoker/cfr/tests/test_cfr_metrics.py:11-99: exact equality with the written-out sums."""
    o = orc.Oracle(otrees('rps'))
    v1, v2 = o.node_values(np.array([0.5, 0.3, 0.2, 0.3, 0.3, 0.4]))
    assert v1[1] == 0 * 0.3 + -1 * 0.3 + 1 * 0.4 and v2[1] == 0 * 0.3 + 1 * 0.3 + -1 * 0.4
    assert v2[2] == -1 * 0.3 + 0 * 0.3 + 1 * 0.4 and v2[3] == 1 * 0.3 + -1 * 0.3 + 0 * 0.4
    assert v1[0] == (0.5 * (0 * 0.3 + -1 * 0.3 + 1 * 0.4) + 0.3 * (1 * 0.3 + 0 * 0.3 + -1 * 0.4) +
                     0.2 * (-1 * 0.3 + 1 * 0.3 + 0 * 0.4))


def test_immediate_regret_detail_matches_reference(metrics, otrees):
    J, arr = metrics
    t = otrees('leduc2')
    o = orc.Oracle(t)
    order = flat.postorder_infoset_order(t)
    series = []
    for step in range(5):
        o.vanilla_levels(1)
        value, per_iset = o.immediate_regret_step(o.sigma, step, order, detail=True)
        series.append(value)
        assert np.array_equal(per_iset, arr['leduc2_imm_detail_%d' % step]), step
    assert series == J['leduc2_imm_series'] and series[-1] == J['leduc2_imm_last']
    assert [repr(t.keys[i]) for i in order[:6]] == J['leduc2_imm_order_head']
