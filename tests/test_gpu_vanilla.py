"""GPU parity: the CUDA level sweep (through the C ABI) against the CPU oracle on the same games."""
import numpy as np
import pytest

from conftest import make_game, oracle_tree

pytestmark = pytest.mark.gpu


def _solver(name):
    from rlpoker_b200.solver import DeviceCFR, device_tree
    return DeviceCFR(device_tree(make_game(name)))


@pytest.mark.parametrize('name', ['rps', 'ocp4', 'leduc2', 'leduc3'])
def test_vanilla_bit_exact(name):
    from oracle import oracle as orc
    o = orc.Oracle(oracle_tree(name))
    s = _solver(name)
    done = 0
    for T in (1, 2, 5, 13, 40):          # 40 crosses the 8-iteration CUDA-graph path
        s.vanilla(T - done)
        o.vanilla_levels(T - done)
        done = T
        assert np.array_equal(s.regrets, o.R), (name, T)
        assert np.array_equal(s.strategy_sum, o.S), (name, T)
        assert np.array_equal(s.sigma, o.sigma), (name, T)
        assert np.array_equal(s.average, o.average_strategy()[0]), (name, T)
        assert s.nodes_touched == o.nodes_touched
        e_gpu = s.exploitability()
        assert abs(e_gpu[0] - o.exploitability(o.average_strategy()[0])) < 1e-12


def test_vanilla_matches_reference_golden(golden, golden_arrays):
    """Straight against numbers the reference itself produced (tests/golden)."""
    s = _solver('leduc3')
    s.vanilla(10)
    assert np.array_equal(s.regrets, golden_arrays['vanilla_leduc3_T10_R'])
    assert np.array_equal(s.strategy_sum, golden_arrays['vanilla_leduc3_T10_S'])
    assert np.array_equal(s.average, golden_arrays['vanilla_leduc3_T10_avg'])
    assert abs(s.exploitability()[0] - golden['vanilla']['leduc3']['10']['exploitability']) < 1e-12


def test_linear_weight(golden_arrays):
    from oracle import oracle as orc
    s = _solver('leduc2')
    s.vanilla(6, linear_weight=True)
    assert np.array_equal(s.regrets, golden_arrays['vanilla_linear_leduc2_T6_R'])
    o = orc.Oracle(oracle_tree('leduc2'))
    o.vanilla_levels(6, linear_weight=True)
    assert np.array_equal(s.strategy_sum, o.S)


@pytest.mark.parametrize('name', ['rps', 'ocp4', 'leduc2', 'leduc3'])
def test_best_response_uniform(golden, golden_arrays, name):
    from rlpoker_b200.solver import device_tree
    from oracle import oracle as orc
    dt = device_tree(make_game(name))
    o = orc.Oracle(oracle_tree(name))
    want = golden['uniform'][name]
    v1, a1 = dt.best_response(o.uniform(), 1, want_argmax=True)
    v2, a2 = dt.best_response(o.uniform(), 2, want_argmax=True)
    assert abs(v1 - want['br1']) < 1e-12 and abs(v2 - want['br2']) < 1e-12
    t = oracle_tree(name)
    amax = np.where(t.iset_player == 1, a1, a2)
    assert np.array_equal(amax, golden_arrays[name + '_uniform_br_argmax'])
    ev = dt.expected_value(o.uniform())
    assert list(ev) == want['expected_value_exact']


def test_best_response_random_strategies():
    from rlpoker_b200.solver import device_tree
    from oracle import oracle as orc
    rng = np.random.RandomState(1)
    for name in ('ocp4', 'leduc3'):
        t = oracle_tree(name)
        o = orc.Oracle(t)
        dt = device_tree(make_game(name))
        for _ in range(3):
            x = rng.random_sample(o.Q) + 0.05
            sigma = x / np.repeat(np.add.reduceat(x, t.act_off[:-1]), t.nact)
            for p in (1, 2):
                assert abs(dt.best_response(sigma, p) - o.best_response(sigma, p)) < 1e-12
            assert dt.expected_value(sigma) == o.expected_value_exact(sigma)


def test_immediate_regret_series(golden):
    from oracle import flat
    want = golden['cfr_driver_leduc3']['cumulative_immediate_regrets']
    order = flat.postorder_infoset_order(oracle_tree('leduc3'))
    s = _solver('leduc3')
    got = []
    for it in range(len(want)):
        s.vanilla(1)
        got.append(s.immediate_regret_step(it, order))
    assert got == want


def test_write_read_roundtrip_and_reset():
    s = _solver('leduc2')
    q = s.tree.num_pairs
    x = np.arange(q, dtype=np.float64)
    s.write(0, x)
    assert np.array_equal(s.read(0), x)
    s.reset()
    assert not s.read(0).any() and s.nodes_touched == 0


def test_long_horizon_parity_leduc3():
    """BASELINE config 2 at test scale: 20 000 vanilla iterations on Leduc-3, CUDA vs the CPU oracle (which
    tests/test_oracle_golden.py pins to the reference at T <= 100).  north_star tolerance: average strategy within
    1e-9 abs, exploitability within 1e-7; the single-GPU path is expected to be bit-identical."""
    from oracle import oracle as orc
    o = orc.Oracle(oracle_tree('leduc3'))
    s = _solver('leduc3')
    T = 20000
    s.vanilla(T)
    o.vanilla_levels(T, threads=4)
    assert np.array_equal(s.regrets, o.R) and np.array_equal(s.strategy_sum, o.S)
    avg = o.average_strategy()[0]
    assert np.abs(s.average - avg).max() <= 1e-9
    assert abs(s.exploitability()[0] - o.exploitability(avg)) <= 1e-7
    assert s.nodes_touched == 2 * 25537 * T
