"""Generic ExtensiveGame trees (random shapes, non-zero-sum payoffs, up to 5 actions, chance anywhere) and the
level-sweep fallback path, against the oracle's faithful recursion."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, make_game, random_game

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('seed,zero_sum', [(0, False), (1, False), (2, True), (3, False), (4, True), (5, False)])
def test_random_tree_bit_exact(seed, zero_sum):
    from oracle import flat, oracle as orc
    from rlpoker_b200.solver import DeviceCFR, device_tree
    from rlpoker_b200._lib import RLP_CHANCE_SAMPLED, RLP_EXTERNAL_SAMPLED
    game = random_game(seed, zero_sum=zero_sum)
    t = flat.flatten_game(game)
    o = orc.Oracle(t)
    s = DeviceCFR(device_tree(game))
    o.cfr_recursive(7)
    s.vanilla(7)
    assert np.array_equal(s.regrets, o.R) and np.array_equal(s.strategy_sum, o.S) and np.array_equal(s.sigma, o.sigma)
    avg = o.average_strategy()[0]
    for p in (1, 2):
        assert abs(s.tree.best_response(avg, p) - o.best_response(avg, p)) < 1e-12
    assert s.tree.expected_value(avg) == o.expected_value_exact(avg)
    # sampled variants on the same tree
    for variant, run in ((RLP_CHANCE_SAMPLED, lambda oo: oo.cfr_recursive(40, sampling=True, rng=orc.rng_philox(7))),
                         (RLP_EXTERNAL_SAMPLED, lambda oo: oo.external(40, orc.rng_philox(7)))):
        s.reset()
        o2 = orc.Oracle(t)
        run(o2)
        s.sampled(variant, 40, lanes=1, seed=7)
        assert np.array_equal(s.regrets, o2.R) and np.array_equal(s.strategy_sum, o2.S)
        assert s.nodes_touched == o2.nodes_touched


def test_level_sweep_fallback_path():
    """RLP_NO_TILES=1 forces the per-level kernels (the path taken when a tree cannot be tiled)."""
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
from conftest import make_game, oracle_tree, random_game
from oracle import flat, oracle as orc
from rlpoker_b200.solver import DeviceCFR, device_tree
for name in ('leduc3', 'rps'):
    o = orc.Oracle(oracle_tree(name)); s = DeviceCFR(device_tree(make_game(name)))
    o.vanilla_levels(12); s.vanilla(12)
    assert np.array_equal(s.regrets, o.R) and np.array_equal(s.strategy_sum, o.S)
g = random_game(11, zero_sum=False)
o = orc.Oracle(flat.flatten_game(g)); s = DeviceCFR(device_tree(g))
o.cfr_recursive(5); s.vanilla(5)
assert np.array_equal(s.regrets, o.R)
print('fallback ok')
''' % (ROOT, os.path.join(ROOT, 'tests'))
    env = dict(os.environ, RLP_NO_TILES='1')
    out = subprocess.check_output([sys.executable, '-c', code], env=env, text=True)
    assert 'fallback ok' in out


def test_interpreter_kernels_path():
    """RLP_NO_JIT=1: the hand-written interpreter kernels (k_micro, k_tile_sweep) instead of the NVRTC ones."""
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
from conftest import make_game, oracle_tree
from oracle import oracle as orc
from rlpoker_b200.solver import DeviceCFR, device_tree
for name in ('leduc3', 'ocp4', 'rps'):
    o = orc.Oracle(oracle_tree(name)); s = DeviceCFR(device_tree(make_game(name)))
    assert s.tree.layout['jit_shapes'] == 0 and s.tree.layout['upper_jit'] == 0 and s.tree.layout['tiled'] == 1
    o.vanilla_levels(23); s.vanilla(23)
    assert np.array_equal(s.regrets, o.R) and np.array_equal(s.strategy_sum, o.S) and np.array_equal(s.sigma, o.sigma)
print('interpreter ok')
''' % (ROOT, os.path.join(ROOT, 'tests'))
    env = dict(os.environ, RLP_NO_JIT='1')
    out = subprocess.check_output([sys.executable, '-c', code], env=env, text=True)
    assert 'interpreter ok' in out
