"""Second golden set (tests/golden/reference_variants.*, produced by oracle/gen_golden_variants.py from the
unmodified reference): Leduc with other `max_raises` / `raise_amount` / suit counts and other OneCardPoker sizes.
Checks, on the CPU, that this repo's builders (Python objects and the native generator) emit the reference's trees and
that the oracle reproduces the reference's best responses and a short vanilla CFR run bit for bit."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from oracle import flat, oracle as orc
from oracle.gen_golden import sha, tree_record
from oracle.gen_golden_variants import VARIANTS


@pytest.fixture(scope='module')
def variants():
    with open(os.path.join(GOLDEN_DIR, 'reference_variants.json')) as f:
        return json.load(f)


@pytest.fixture(scope='module')
def variant_arrays():
    return dict(np.load(os.path.join(GOLDEN_DIR, 'reference_variants.npz')))


_BUILT = {}


def build(name):
    if name not in _BUILT:
        from rlpoker_b200.games.card import get_deck
        from rlpoker_b200.games.leduc import Leduc
        from rlpoker_b200.games.one_card_poker import OneCardPoker
        kind, a = VARIANTS[name]
        if kind == 'leduc':
            g = Leduc(get_deck(a['values'], a['suits']), max_raises=a['max_raises'], raise_amount=a['raise_amount'])
        else:
            g = OneCardPoker.create_game(a['n'])
        _BUILT[name] = (g, flat.flatten_game(g))
    return _BUILT[name]


@pytest.mark.parametrize('name', sorted(VARIANTS))
def test_tree_layout(variants, name):
    want = variants['games'][name]['tree']
    got = tree_record(build(name)[1], full=False)
    for k in want:
        assert got[k] == want[k], (name, k)


@pytest.mark.parametrize('name', sorted(k for k, v in VARIANTS.items() if v[0] == 'leduc'))
def test_native_generator_emits_the_same_tree(name):
    from rlpoker_b200.flatten import flatten
    from rlpoker_b200.games.leduc_native import leduc_flat_tree
    a = VARIANTS[name][1]
    py = flatten(build(name)[0])
    nat = leduc_flat_tree(a['values'], a['suits'], max_raises=a['max_raises'], raise_amount=a['raise_amount'])
    assert py.num_nodes == nat.num_nodes and py.num_infosets == nat.num_infosets
    for field in ('parent', 'player', 'infoset', 'cprob', 'u1', 'u2', 'hidden', 'first_child', 'level_off', 'act_off',
                  'iset_nodes', 'iset_node_off', 'nact'):
        assert np.array_equal(getattr(py, field), getattr(nat, field)), (name, field)
    assert py.info_set_keys == nat.info_set_keys


@pytest.mark.parametrize('name', sorted(VARIANTS))
def test_best_response_and_short_cfr_run(variants, variant_arrays, name):
    want = variants['games'][name]
    t = build(name)[1]
    o = orc.Oracle(t)
    assert [o.best_response(o.uniform(), p) for p in (1, 2)] == want['uniform_br']
    assert o.exploitability(o.uniform()) == want['uniform_exploitability']
    for mode in ('recursive', 'levels'):
        o = orc.Oracle(t)
        o.cfr_recursive(want['cfr']['T']) if mode == 'recursive' else o.vanilla_levels(want['cfr']['T'], threads=2)
        avg, present = o.average_strategy()
        assert np.array_equal(o.R, variant_arrays[name + '_R']), (name, mode)
        assert sha(o.S) == want['cfr']['sha_S']
        assert np.array_equal(np.where(np.repeat(present, t.nact), avg, 0.0), variant_arrays[name + '_avg'])
        assert np.array_equal(present.astype(np.uint8), variant_arrays[name + '_avg_present'])
        assert o.exploitability(avg) == want['cfr']['exploitability']
        assert o.nodes_touched == want['cfr']['nodes_touched']


@pytest.mark.parametrize('name', sorted(VARIANTS))
def test_device_layout_plans_and_compiles(name):
    """rlp_tree_plan on the host: the variant trees tile and every generated kernel compiles (no GPU)."""
    from rlpoker_b200.flatten import flatten
    from rlpoker_b200.solver import plan_tree
    p = plan_tree(flatten(build(name)[0]))
    assert p['tiled'] == 1 and p['zero_sum'] == 1 and p['kernels_failed'] == 0
    assert p['micro_subtrees'] > 0
