"""GPU parity at the sizes the numbers are quoted on: native Leduc-5 / -9 / -13 trees (rlp_leduc_generate), the CUDA
sweep through the C ABI against the CPU oracle's level sweep on all host threads.  23 iterations = two replays of the
10-iteration CUDA graph plus an eager tail, so the pipelined graph, the fan-in kernel, the NVRTC kernels and the
eager path are all on the compared trajectory.  north_star tolerance: average strategy within 1e-9, exploitability
within 1e-7; the single-GPU path is expected to be bit-identical (regrets, action counts, strategy)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

_TREES = {}


def native_tree(n):
    if n not in _TREES:
        from oracle import flat
        from rlpoker_b200.games.leduc_native import leduc_flat_tree
        ft = leduc_flat_tree(n)
        ot = flat.from_arrays(ft.num_nodes, ft.parent, ft.first_child, ft.player, ft.infoset, ft.label, ft.hidden,
                              ft.cprob, ft.u1, ft.u2)
        _TREES[n] = (ft, ot)
    return _TREES[n]


def _check(s, o, tag):
    assert np.array_equal(s.regrets, o.R), tag
    assert np.array_equal(s.strategy_sum, o.S), tag
    assert np.array_equal(s.sigma, o.sigma), tag
    avg = o.average_strategy()[0]
    assert np.abs(s.average - avg).max() <= 1e-9, tag
    assert s.nodes_touched == o.nodes_touched, tag
    return avg


@pytest.mark.parametrize('n', [5, 9, 13])
def test_vanilla_bit_exact_native_leduc(n):
    from oracle import oracle as orc
    from rlpoker_b200.solver import DeviceCFR, DeviceTree
    ft, ot = native_tree(n)
    threads = os.cpu_count() or 1
    dt = DeviceTree(ft)
    lay = dt.layout
    assert lay['tiled'] == 1 and lay['jit_shapes'] >= 1 and lay['upper_jit'] == 1, lay
    s = DeviceCFR(dt)
    o = orc.Oracle(ot)
    done = 0
    for T in (3, 23):                       # eager only, then graph x2 + eager tail
        s.vanilla(T - done)
        o.vanilla_levels(T - done, threads=threads)
        done = T
        avg = _check(s, o, ('leduc%d' % n, T))
    assert abs(s.exploitability()[0] - o.exploitability(avg)) <= 1e-7
    del s
    dt.close()


@pytest.mark.parametrize('n', [5, 13])
def test_linear_weight_through_the_graph(n):
    from oracle import oracle as orc
    from rlpoker_b200.solver import DeviceCFR, DeviceTree
    ft, ot = native_tree(n)
    threads = os.cpu_count() or 1
    dt = DeviceTree(ft)
    s = DeviceCFR(dt)
    o = orc.Oracle(ot)
    s.vanilla(23, linear_weight=True)
    o.vanilla_levels(23, linear_weight=True, threads=threads)
    avg = _check(s, o, ('leduc%d linear' % n,))
    s.vanilla(11)                           # back to unit weights on the same state
    o.vanilla_levels(11, threads=threads)
    _check(s, o, ('leduc%d mixed' % n,))
    assert abs(s.exploitability()[0] - o.exploitability(o.average_strategy()[0])) <= 1e-7
    del s
    dt.close()


def test_best_response_leduc13_vs_oracle():
    """Best response at the quoted size: uniform and a CFR strategy, both responders, 1e-9 (the level sweep sums the
    same terms in another association than br())."""
    from oracle import oracle as orc
    from rlpoker_b200.solver import DeviceCFR, DeviceTree
    ft, ot = native_tree(13)
    dt = DeviceTree(ft)
    o = orc.Oracle(ot)
    for p in (1, 2):
        assert abs(dt.best_response(o.uniform(), p) - o.best_response(o.uniform(), p)) < 1e-9
    s = DeviceCFR(dt)
    s.vanilla(5)
    avg = s.average
    for p in (1, 2):
        assert abs(dt.best_response(avg, p) - o.best_response(avg, p)) < 1e-9
    del s
    dt.close()
