"""Pins the CPU oracle (oracle/cfr_oracle.c) bit-for-bit to outputs of the reference itself
(tests/golden/, produced by oracle/gen_golden.py) and checks this repo's game builders produce
the reference's trees.  CPU only."""
import hashlib

import numpy as np
import pytest

from oracle import flat, oracle as orc
from oracle.gen_golden import tree_record, sha


def test_regret_matching_kats(golden):
    for case in golden['regret_matching']:
        out = orc.regret_matching(case['r'], eps=case['eps'], highest_regret=case['hr'])
        assert out.tolist() == case['out'], case
    for case in golden['normalise_probs']:
        assert orc.normalise_probs(case['p'], eps=case['eps']).tolist() == case['out']


def test_sampler_matches_numpy_choice(golden):
    g = golden['sample_action']
    np.random.seed(0)
    u = np.random.random_sample(20)
    assert [[1, 2][orc.draw([0.2, 0.8], x)] for x in u] == g['a']
    assert [['R', 'P', 'S'][orc.draw([0.2, 0.7, 0.1], x)] for x in u] == g['b']
    np.random.seed(g['c_seed'])
    u = np.random.random_sample(len(g['c']))
    assert [orc.draw(g['c_probs'], x) for x in u] == g['c']      # 10 outcomes: numpy's pairwise sum path


@pytest.mark.parametrize('name', ['rps', 'ocp4', 'leduc2', 'leduc3', 'leduc4'])
def test_tree_layout_matches_reference(golden, otrees, name):
    if name not in golden['trees']:
        pytest.skip('not in golden set')
    want = golden['trees'][name]
    got = tree_record(otrees(name), full='parent' in want)
    for k in want:
        assert got[k] == want[k], (name, k)


@pytest.mark.parametrize('name', ['rps', 'ocp4', 'leduc2', 'leduc3'])
def test_uniform_best_response(golden, golden_arrays, otrees, name):
    t = otrees(name)
    o = orc.Oracle(t)
    want = golden['uniform'][name]
    v1, a1 = o.best_response(o.uniform(), 1, want_argmax=True)
    v2, a2 = o.best_response(o.uniform(), 2, want_argmax=True)
    assert v1 == want['br1'] and v2 == want['br2']
    assert o.exploitability(o.uniform()) == want['exploitability']
    amax = np.where(t.iset_player == 1, a1, a2)
    assert np.array_equal(amax, golden_arrays[name + '_uniform_br_argmax'])
    assert list(o.expected_value_exact(o.uniform())) == want['expected_value_exact']


def test_rps_expected_values(golden, otrees):
    t = otrees('rps')
    o = orc.Oracle(t)
    sigma = np.array([0.5, 0.2, 0.3, 0.2, 0.3, 0.5])
    assert list(o.expected_value_exact(sigma)) == golden['rps_expected_value_exact']


@pytest.mark.parametrize('name', ['rps', 'ocp4', 'leduc2', 'leduc3'])
@pytest.mark.parametrize('mode', ['recursive', 'levels', 'levels_mt'])
def test_vanilla_trajectory(golden, golden_arrays, otrees, name, mode):
    t = otrees(name)
    o = orc.Oracle(t)
    done = 0
    for T in sorted(int(k) for k in golden['vanilla'][name]):
        want = golden['vanilla'][name][str(T)]
        if mode == 'recursive':
            o.cfr_recursive(T - done)
        else:
            o.vanilla_levels(T - done, threads=1 if mode == 'levels' else 3)
        done = T
        avg, present = o.average_strategy()
        assert sha(o.R) == want['sha_R'], (name, T, 'R')
        assert sha(o.S) == want['sha_S'], (name, T, 'S')
        assert sha(np.where(np.repeat(present, t.nact), avg, 0.0)) == want['sha_avg'], (name, T, 'avg')
        assert o.exploitability(avg) == want['exploitability']
        assert o.nodes_touched == want['nodes_touched']
        key = 'vanilla_%s_T%d_sigma' % (name, T)
        if key in golden_arrays:
            assert np.array_equal(o.sigma, golden_arrays[key])


def test_vanilla_linear_weight(golden, golden_arrays, otrees):
    for mode in ('recursive', 'levels'):
        o = orc.Oracle(otrees('leduc2'))
        o.cfr_recursive(6, linear_weight=True) if mode == 'recursive' else o.vanilla_levels(6, linear_weight=True)
        avg, present = o.average_strategy()
        assert np.array_equal(o.R, golden_arrays['vanilla_linear_leduc2_T6_R'])
        got = np.where(np.repeat(present, otrees('leduc2').nact), avg, 0.0)
        assert np.array_equal(got, golden_arrays['vanilla_linear_leduc2_T6_avg'])
        assert o.exploitability(avg) == golden['vanilla_linear']['leduc2']['6']['exploitability']


def test_chance_sampled_numpy_stream(golden, golden_arrays, otrees):
    want = golden['chance_sampled_leduc3']
    t = otrees('leduc3')
    o = orc.Oracle(t)
    np.random.seed(want['seed'])
    rng = orc.rng_stream(np.random.random_sample(want['T'] * 2 * 16))
    o.cfr_recursive(want['T'], sampling=True, rng=rng)
    assert rng.stream_pos == want['T'] * 2 * 11           # 2 hole cards + 9 boards per traversal
    assert np.array_equal(o.R, golden_arrays['chance_leduc3_R'])
    assert np.array_equal(o.S, golden_arrays['chance_leduc3_S'])
    assert np.array_equal(o.in_R, golden_arrays['chance_leduc3_in_R'])
    assert np.array_equal(o.in_S, golden_arrays['chance_leduc3_in_S'])
    assert np.array_equal(o.in_sigma, golden_arrays['chance_leduc3_in_sigma'])
    assert o.nodes_touched == want['nodes_touched']
    avg, present = o.average_strategy()
    assert int(present.sum()) == want['num_present']
    assert o.exploitability(avg) == want['exploitability']
    # one-card poker, other seed
    want = golden['chance_sampled_ocp4']
    o = orc.Oracle(otrees('ocp4'))
    np.random.seed(want['seed'])
    o.cfr_recursive(want['T'], sampling=True, rng=orc.rng_stream(np.random.random_sample(want['T'] * 8)))
    assert np.array_equal(o.R, golden_arrays['chance_ocp4_R']) and np.array_equal(o.S, golden_arrays['chance_ocp4_S'])
    assert o.exploitability(o.average_strategy()[0]) == want['exploitability']


def test_external_sampling_numpy_stream(golden, golden_arrays, otrees):
    want = golden['external_leduc3']
    o = orc.Oracle(otrees('leduc3'))
    np.random.seed(want['seed'])
    o.external(want['T'], orc.rng_stream(np.random.random_sample(want['T'] * 200)))
    assert o.nodes_touched == want['nodes_touched']
    assert np.array_equal(o.R, golden_arrays['external_leduc3_R'])
    assert np.array_equal(o.S, golden_arrays['external_leduc3_alpha'])
    assert np.array_equal(o.in_R, golden_arrays['external_leduc3_in_R'])
    assert np.array_equal(o.in_S, golden_arrays['external_leduc3_in_alpha'])
    assert np.array_equal(o.in_sigma, golden_arrays['external_leduc3_in_sigma'])
    avg = o.alpha_strategy()
    assert np.array_equal(np.where(np.repeat(o.in_S, otrees('leduc3').nact) > 0, avg, 0.0),
                          golden_arrays['external_leduc3_avg'])
    assert o.exploitability(avg) == want['exploitability']
    want = golden['external_rps']
    o = orc.Oracle(otrees('rps'))
    np.random.seed(want['seed'])
    o.external(want['T'], orc.rng_stream(np.random.random_sample(want['T'] * 8)))
    assert o.R.tolist() == want['R'] and o.S.tolist() == want['alpha']
    assert o.exploitability(o.alpha_strategy()) == want['exploitability']


def test_cfr_driver_and_immediate_regret(golden, golden_arrays, otrees):
    """cfr(experiment, game, num_iters=11, use_chance_sampling=False): exploitability at t=0,10 and the
    immediate-regret series (cfr_metrics.py:13-78) over strategies[0..10]."""
    want = golden['cfr_driver_leduc3']
    t = otrees('leduc3')
    o = orc.Oracle(t)
    order = flat.postorder_infoset_order(t)
    series, expl = [], []
    for it in range(11):
        o.vanilla_levels(1)
        if it % 10 == 0:
            expl.append([it, o.exploitability(o.average_strategy()[0])])
        series.append(o.immediate_regret_step(o.sigma, it, order))
    assert expl == want['exploitabilities']
    assert series == want['cumulative_immediate_regrets']
    assert series[-1] == want['immediate_regret_T11']
    assert np.array_equal(o.average_strategy()[0], golden_arrays['cfr_driver_leduc3_avg'])


def test_philox_reference_vector():
    """Philox-4x32-10 known-answer test (Random123 kat_vectors: counter/key all zero and all ones)."""
    import ctypes as C
    lib = orc.lib()
    # zero counter/key -> 6627e8d5 e169c58d bc57ac4c 9b00dbd8 ; our uniform uses words 0 and 1
    u = orc.philox_uniform(0, 0, 0, 0, 0)
    a, b = 0x6627e8d5 >> 5, 0xe169c58d >> 6
    assert u == (a * 67108864.0 + b) / 9007199254740992.0


def test_external_driver_exploitability_trace(golden, otrees):
    """external_sampling_cfr(experiment, game, num_iters=401) after np.random.seed(0): exploitability of
    AverageStrategy.compute_strategy() at t = 0, 200, 400 (external_cfr.py:59-63)."""
    if 'external_driver_leduc3' not in golden:
        pytest.skip('quick golden set')
    want = golden['external_driver_leduc3']
    o = orc.Oracle(otrees('leduc3'))
    np.random.seed(want['seed'])
    rng = orc.rng_stream(np.random.random_sample(401 * 200))
    got = []
    for t in range(401):
        o.external(1, rng)
        if t % 200 == 0:
            got.append([t, o.exploitability(o.alpha_strategy())])
    assert got == want['exploitabilities']


def test_chance_sampled_driver_exploitability_trace(golden, otrees):
    """cfr(experiment, game, num_iters=101, use_chance_sampling=True) after np.random.seed(0): exploitability of the
    average strategy every 10 iterations (cfr.py:125-129)."""
    if 'chance_driver_leduc3' not in golden:
        pytest.skip('quick golden set')
    want = golden['chance_driver_leduc3']
    o = orc.Oracle(otrees('leduc3'))
    np.random.seed(want['seed'])
    rng = orc.rng_stream(np.random.random_sample(101 * 2 * 16))
    got = []
    for t in range(101):
        o.cfr_recursive(1, sampling=True, rng=rng)
        if t % 10 == 0:
            got.append([t, o.exploitability(o.average_strategy()[0])])
    assert got == want['exploitabilities']


def test_vanilla_leduc3_300_iterations_survey_values(otrees):
    """SURVEY.md section 8(c) lists what the reference itself produced at T = 200 and 300 on Leduc-3 (an 8 minute
    run in the survey session): exploitability, nodes_touched and the first 16 hex digits of the SHA-256 of the
    average strategy.  The committed golden set stops at T = 100; this pins the longer horizon."""
    o = orc.Oracle(otrees('leduc3'))
    o.vanilla_levels(200, threads=4)
    assert o.exploitability(o.average_strategy()[0]) == 0.2504793684655993
    o.vanilla_levels(100, threads=4)
    avg, present = o.average_strategy()
    assert present.all()
    assert o.exploitability(avg) == 0.1813807003021225
    assert sha(avg)[:16] == '95ae4634f6f5513c'
    assert o.nodes_touched == 15322200


@pytest.mark.parametrize('n', [2, 3, 4])
def test_oracle_leduc_generator_matches_reference_layout(golden, otrees, n):
    """oracle/leduc_oracle.c (the generator behind bench.py's CPU arms) against the layout hashes the reference
    itself produced, and array for array against the flattened object tree."""
    from oracle.gen_golden import sha
    name = 'leduc%d' % n
    t = orc.leduc_tree(n)
    if name in golden['trees']:
        want = golden['trees'][name]
        assert t.num_nodes == want['num_nodes'] and t.num_infosets == want['num_infosets']
        assert [int(x) for x in t.level_off] == want['level_off']
        for k in ('parent', 'player', 'infoset', 'cprob', 'u1', 'u2', 'hidden', 'first_child'):
            assert sha(getattr(t, k)) == want['sha_' + k], (name, k)
    a = otrees(name)
    for k in ('parent', 'first_child', 'player', 'infoset', 'hidden', 'cprob', 'u1', 'u2', 'act_off', 'iset_nodes'):
        assert np.array_equal(getattr(a, k), getattr(t, k)), (name, k)
    # labels: 0 fold, 1 call, 2 bet, 16 + deck index for cards
    for i in range(1, a.num_nodes):
        lab = a.labels[a.label[i]]
        code = int(t.label[i])
        assert (code == lab) if isinstance(lab, int) else (code == 16 + lab.value * 2 + lab.suit), i
