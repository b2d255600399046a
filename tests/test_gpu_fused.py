"""The group-fused persistent iteration kernel (csrc/fused.cu) against the CPU oracle and against the five-kernel
pipeline it replaces, incl. switching between the two on one solver (the strategy copies each path keeps must be
rebuilt when the other one ran)."""
import os

import numpy as np
import pytest

from conftest import make_game, oracle_tree

pytestmark = pytest.mark.gpu


def _env(**kw):
    class Ctx:
        def __enter__(self):
            self.old = {k: os.environ.get(k) for k in kw}
            os.environ.update({k: v for k, v in kw.items()})

        def __exit__(self, *a):
            for k, v in self.old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
    return Ctx()


@pytest.mark.parametrize('name', ['leduc2', 'leduc3'])
def test_fused_is_in_use_and_bit_exact(name):
    from oracle import oracle as orc
    from rlpoker_b200.flatten import flatten
    from rlpoker_b200.solver import DeviceCFR, DeviceTree
    dt = DeviceTree(flatten(make_game(name)))
    lay = dt.layout
    assert lay['fused'] == 1 and lay['fused_batches'] > 0 and lay['fused_grid'] >= 148, lay
    s = DeviceCFR(dt)
    o = orc.Oracle(oracle_tree(name))
    done = 0
    for T in (1, 2, 3, 10, 31):                # odd and even launch lengths: both strategy buffers end up current
        s.vanilla(T - done)
        o.vanilla_levels(T - done)
        done = T
        assert np.array_equal(s.regrets, o.R), (name, T)
        assert np.array_equal(s.strategy_sum, o.S), (name, T)
        assert np.array_equal(s.sigma, o.sigma), (name, T)
        assert np.array_equal(s.average, o.average_strategy()[0]), (name, T)
        assert s.nodes_touched == o.nodes_touched
        assert (s.flags() == 7).all()
    s.vanilla(9, linear_weight=True)
    o.vanilla_levels(9, linear_weight=True)
    assert np.array_equal(s.regrets, o.R) and np.array_equal(s.strategy_sum, o.S)
    del s
    dt.close()


def test_switching_between_fused_and_pipeline():
    from oracle import oracle as orc
    from rlpoker_b200._lib import RLP_SIGMA
    from rlpoker_b200.flatten import flatten
    from rlpoker_b200.solver import DeviceCFR, DeviceTree
    dt = DeviceTree(flatten(make_game('leduc3')))
    s = DeviceCFR(dt)
    o = orc.Oracle(oracle_tree('leduc3'))
    s.vanilla(7)
    with _env(RLP_NO_FUSED='1'):
        s.vanilla(5)                            # five-kernel pipeline (eager)
        s.vanilla(12)                           # ... and its 10-iteration graph
    s.vanilla(11)
    o.vanilla_levels(35)
    assert np.array_equal(s.regrets, o.R) and np.array_equal(s.strategy_sum, o.S) and np.array_equal(s.sigma, o.sigma)
    # a host write of the strategy must reach the fused kernel's copy
    sig = o.uniform()
    s.write(RLP_SIGMA, sig)
    o.sigma[:] = sig
    s.vanilla(3)
    o.vanilla_levels(3)
    assert np.array_equal(s.regrets, o.R) and np.array_equal(s.sigma, o.sigma)
    # sampled iterations in between (they rewrite the strategy by regret matching)
    s.reset()
    o.reset()
    s.vanilla(4)
    s.sampled(0, 3, lanes=1, seed=5)
    s.vanilla(4)
    o.vanilla_levels(4)
    o.cfr_recursive(3, sampling=True, rng=orc.rng_philox(5))
    o.vanilla_levels(4)
    assert np.array_equal(s.regrets, o.R) and np.array_equal(s.strategy_sum, o.S)
    del s
    dt.close()


def test_fused_timeline_reports_phases():
    from rlpoker_b200.flatten import flatten
    from rlpoker_b200.solver import DeviceCFR, DeviceTree
    dt = DeviceTree(flatten(make_game('leduc3')))
    s = DeviceCFR(dt)
    s.vanilla(3)
    tl = s.fused_timeline(12)
    assert {'micro_phase', 'barrier_1', 'upper_phase', 'barrier_2', 'iteration'} <= set(tl)
    assert all(v >= 0 for k, v in tl.items() if k != 'upper_split') and 0 < tl['iteration'] < 1000
    assert s.t == 15
    del s
    dt.close()


def test_trees_without_the_group_structure_keep_the_pipeline():
    from rlpoker_b200.flatten import flatten
    from rlpoker_b200.solver import DeviceTree
    for name in ('ocp4', 'rps'):
        dt = DeviceTree(flatten(make_game(name)))
        assert dt.layout['fused'] in (0, 1)      # whichever it is, test_gpu_vanilla checks it bit for bit
        dt.close()
