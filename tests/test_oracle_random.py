"""Third golden set (tests/golden/reference_random.*, oracle/gen_golden_random.py): random non-poker trees -- up to
five actions, uneven chance probabilities, float payoffs, leaves at several depths, odd seeds not zero-sum -- solved by
the unmodified reference.  This repo's `ExtensiveGame` must flatten to the same arrays and the oracle (recursion and
level sweep) must reproduce the reference's regrets, action counts, average strategy and best responses bit for bit.
The CUDA path is compared with the oracle on the same trees in tests/test_gpu_generic_trees.py."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, random_game
from oracle import flat, oracle as orc
from oracle.gen_golden import sha, tree_record
from oracle.gen_golden_random import SEEDS, zero_sum


@pytest.fixture(scope='module')
def want():
    with open(os.path.join(GOLDEN_DIR, 'reference_random.json')) as f:
        return json.load(f)


@pytest.fixture(scope='module')
def arrays():
    return dict(np.load(os.path.join(GOLDEN_DIR, 'reference_random.npz')))


@pytest.mark.parametrize('seed', SEEDS)
def test_random_tree_against_reference(want, arrays, seed):
    w = want['games'][str(seed)]
    t = flat.flatten_game(random_game(seed, zero_sum=zero_sum(seed)))
    got = tree_record(t, full=False)
    for k in w['tree']:
        assert got[k] == w['tree'][k], (seed, k)
    o = orc.Oracle(t)
    assert [o.best_response(o.uniform(), p) for p in (1, 2)] == w['uniform_br']
    assert list(o.expected_value_exact(o.uniform())) == w['uniform_value']
    for mode in ('recursive', 'levels', 'levels_mt'):
        o = orc.Oracle(t)
        if mode == 'recursive':
            o.cfr_recursive(w['cfr']['T'])
        else:
            o.vanilla_levels(w['cfr']['T'], threads=1 if mode == 'levels' else 3)
        avg, present = o.average_strategy()
        assert np.array_equal(o.R, arrays['seed%d_R' % seed]), (seed, mode)
        assert np.array_equal(o.S, arrays['seed%d_S' % seed]), (seed, mode)
        assert np.array_equal(o.sigma, arrays['seed%d_sigma' % seed]), (seed, mode)
        assert np.array_equal(np.where(np.repeat(present, t.nact), avg, 0.0), arrays['seed%d_avg' % seed])
        assert o.exploitability(avg) == w['cfr']['exploitability']
        assert o.nodes_touched == w['cfr']['nodes_touched']


@pytest.mark.parametrize('seed', SEEDS)
def test_random_tree_sampled_variants_numpy_stream(want, arrays, seed):
    """Chance-sampled (cfr.py:167-175) and external-sampling (external_cfr.py:82-173) runs of the reference driven by
    numpy's seeded legacy generator; the oracle consumes the same MT19937 doubles through the restated sampler."""
    w = want['games'][str(seed)]
    t = flat.flatten_game(random_game(seed, zero_sum=zero_sum(seed)))
    cs = w['chance_sampled']
    o = orc.Oracle(t)
    np.random.seed(cs['seed'])
    o.cfr_recursive(cs['T'], sampling=True, rng=orc.rng_stream(np.random.random_sample(cs['T'] * 2 * 4000)))
    assert o.nodes_touched == cs['nodes_touched']
    assert np.array_equal(o.R, arrays['seed%d_cs_R' % seed]) and np.array_equal(o.S, arrays['seed%d_cs_S' % seed])
    assert np.array_equal(o.in_R, arrays['seed%d_cs_in_R' % seed]) and np.array_equal(o.in_S, arrays['seed%d_cs_in_S' % seed])
    assert o.exploitability(o.average_strategy()[0]) == cs['exploitability']
    ex = w['external']
    o = orc.Oracle(t)
    np.random.seed(ex['seed'])
    o.external(ex['T'], orc.rng_stream(np.random.random_sample(ex['T'] * 2 * 4000)))
    assert o.nodes_touched == ex['nodes_touched']
    assert np.array_equal(o.R, arrays['seed%d_ex_R' % seed]) and np.array_equal(o.S, arrays['seed%d_ex_alpha' % seed])
    assert np.array_equal(o.in_R, arrays['seed%d_ex_in_R' % seed]) and np.array_equal(o.in_S, arrays['seed%d_ex_in_alpha' % seed])
    assert o.exploitability(o.alpha_strategy()) == ex['exploitability']
