"""The drop-in Python API (reference call shapes) end to end on the GPU, against reference goldens."""
import os
import pickle

import numpy as np
import pytest

from conftest import make_game, oracle_tree

pytestmark = pytest.mark.gpu


def test_legacy_call_shape_matches_reference_run(golden, golden_arrays, capsys):
    """cfr(game, num_iters=11, use_chance_sampling=False)  (README.md:19, tests/test_best_response.py:16)."""
    from rlpoker_b200.cfr.cfr import cfr
    from rlpoker_b200.best_response import compute_exploitability
    from rlpoker_b200.games.card import Card
    want = golden['cfr_driver_leduc3']
    game = make_game('leduc3')
    strategy, exploitabilities, strategies = cfr(game, num_iters=11, use_chance_sampling=False)
    assert [t for t, _ in exploitabilities] == [t for t, _ in want['exploitabilities']]
    for (_, got), (_, ref) in zip(exploitabilities, want['exploitabilities']):
        assert abs(got - ref) < 1e-12
    assert strategies.immediate_regrets == want['cumulative_immediate_regrets']
    assert len(strategies) == 11
    key = (Card(0, 0), -1, 1, 1, Card(1, 0), 1, 2, 2)
    assert repr(key) == want['sample_entry']['key']
    assert {str(a): p for a, p in strategy[key].items()} == want['sample_entry']['value']
    assert np.array_equal(strategy.sigma_array, golden_arrays['cfr_driver_leduc3_avg'])
    assert abs(compute_exploitability(game, strategy) - want['exploitabilities'][-1][1]) < 1e-12
    out = capsys.readouterr().out
    assert 'nodes touched: 51074,' in out and 'Immediate regret:' in out
    # a per-iteration strategy behaves like the reference's Strategy objects
    s5 = strategies[5]
    assert key in s5 and abs(sum(s5[key].values()) - 1.0) < 1e-12 and len(s5.get_info_sets()) == 2208


def test_experiment_call_shape_logs_and_saves(tmp_path, golden):
    from rlpoker_b200.cfr.cfr import cfr
    from rlpoker_b200.experiment import Experiment, ExperimentWriter

    class Writer(ExperimentWriter):
        rows = []

        def log(self, data, global_step=None):
            self.rows.append((global_step, data))

    exp = Experiment('run', base_save_path=str(tmp_path))
    w = Writer()
    strategy, expl, _ = cfr(exp, make_game('leduc2'), num_iters=21, use_chance_sampling=False, experiment_writer=w,
                            verbose=False)
    assert [s for s, _ in w.rows] == [0, 10, 20]
    assert set(w.rows[0][1]) == {'nodes_touched', 'exploitability_mbbh', 'immediate_regret'}
    with open(os.path.join(exp.save_path, 'best_strategy.pkl'), 'rb') as f:
        best = pickle.load(f)
    assert best == strategy.copy() or len(best.get_info_sets()) == len(strategy.get_info_sets())
    assert abs(expl[-1][1] - golden['vanilla']['leduc2']['20']['exploitability']) > 0   # t=20 is after 21 iterations


def test_chance_sampling_and_external_drivers(tmp_path):
    from rlpoker_b200.cfr.cfr import cfr
    from rlpoker_b200.cfr.external_cfr import external_sampling_cfr
    from rlpoker_b200.best_response import compute_exploitability
    from rlpoker_b200.experiment import Experiment
    from rlpoker_b200.extensive_game import ActionFloat
    game = make_game('leduc3')
    np.random.seed(0)
    s1, e1, h1 = cfr(game, num_iters=301, use_chance_sampling=True, verbose=False, immediate_regret=False)
    np.random.seed(0)
    s2, e2, _ = cfr(game, num_iters=301, use_chance_sampling=True, verbose=False, immediate_regret=False)
    assert e1 == e2 and e1[-1][1] < e1[0][1]
    assert 0 < len(h1[0].get_info_sets()) < 2208            # only visited information sets, like the reference
    exp = Experiment('ext', base_save_path=str(tmp_path))
    s3, e3, h3 = external_sampling_cfr(exp, game, num_iters=401, verbose=False, seed=3)
    assert [t for t, _ in e3] == [0, 200, 400]
    some = next(iter(s3.strategy.values()))
    assert isinstance(some, ActionFloat)
    assert abs(compute_exploitability(game, s3) - e3[-1][1]) < 1e-12
    # batched lanes converge much further in the same number of iterations
    s4, e4, _ = external_sampling_cfr(None, game, num_iters=201, verbose=False, seed=3, lanes=256,
                                      immediate_regret=False, keep_strategies=0)
    assert e4[-1][1] < 0.6 * e3[1][1]


def test_best_response_api(golden):
    from rlpoker_b200.best_response import compute_best_response, compute_exploitability
    from rlpoker_b200.extensive_game import Strategy
    for name in ('rps', 'ocp4', 'leduc3'):
        game = make_game(name)
        want = golden['uniform'][name]
        assert abs(compute_exploitability(game, Strategy.initialise()) - want['exploitability']) < 1e-12
        v1, br1 = compute_best_response(game, game.complete_strategy_uniformly(Strategy.initialise())[0], 1)
        assert abs(v1 - want['br1']) < 1e-12
        assert all(sorted(d.values())[-1] == 1.0 and sum(d.values()) == 1.0 for d in br1.values())
        n1 = sum(1 for n, k in game.info_set_ids.items() if n.player == 1)
        assert len(br1) == len({k for n, k in game.info_set_ids.items() if n.player == 1}) and n1 >= len(br1)


def test_average_strategy_object_and_metrics(golden):
    from rlpoker_b200.cfr import cfr_metrics, cfr_util
    from rlpoker_b200.extensive_game import Strategy
    from oracle import oracle as orc
    game = make_game('leduc2')
    t = oracle_tree('leduc2')
    o = orc.Oracle(t)
    avg = cfr_util.AverageStrategy(game)
    # a partial strategy: only player-1 information sets, random probabilities
    rng = np.random.RandomState(0)
    strat, present = {}, np.zeros(t.num_infosets, np.uint8)
    sigma = o.uniform()
    from rlpoker_b200.flatten import flatten
    ft = flatten(game)
    for i in range(t.num_infosets):
        if t.iset_player[i] == 1:
            p = rng.random_sample(int(t.nact[i])) + 0.1
            p /= p.sum()
            strat[ft.info_set_keys[i]] = dict(zip(ft.actions_of(i), p))
            sigma[t.act_off[i]:t.act_off[i + 1]] = p
            present[i] = 1
    cfr_util.update_average_strategy(game, avg, Strategy(strat), weight=2.0)
    import ctypes as C
    alpha = np.zeros(o.Q)
    in_alpha = np.zeros(t.num_infosets, np.uint8)
    orc.lib().orc_update_average_strategy(C.byref(o.c), orc._p(sigma, orc._dp), orc._p(present, orc._u8p),
                                          orc._p(alpha, orc._dp), orc._p(in_alpha, orc._u8p), C.c_double(2.0))
    got = avg.alpha
    assert len(got) == int(in_alpha.sum())
    for i in np.flatnonzero(in_alpha):
        k = ft.info_set_keys[i]
        assert list(got[k].values()) == alpha[t.act_off[i]:t.act_off[i + 1]].tolist()
    # immediate regret of explicit Strategy objects
    imm, series, _ = cfr_metrics.compute_immediate_regret(game, [Strategy(strat), Strategy.initialise()])
    o2 = orc.Oracle(t)
    from oracle import flat
    order = flat.postorder_infoset_order(t)
    want = [o2.immediate_regret_step(sigma, 0, order), o2.immediate_regret_step(o.uniform(), 1, order)]
    assert series == want and imm == want[-1]


def test_flat_game_drop_in(golden):
    """A natively generated Leduc (no node objects) through the same driver, keys rebuilt from the arrays."""
    from rlpoker_b200.cfr.cfr import cfr
    from rlpoker_b200.best_response import compute_exploitability
    from rlpoker_b200.games.card import Card
    from rlpoker_b200.games.leduc_native import create_leduc
    want = golden['cfr_driver_leduc3']
    game = create_leduc(3)
    strategy, expl, _ = cfr(game, num_iters=11, use_chance_sampling=False, verbose=False, immediate_regret=False,
                            keep_strategies=0)
    assert abs(expl[-1][1] - want['exploitabilities'][-1][1]) < 1e-12
    key = (Card(0, 0), -1, 1, 1, Card(1, 0), 1, 2, 2)
    assert {str(a): p for a, p in strategy[key].items()} == want['sample_entry']['value']
    assert abs(compute_exploitability(game, strategy) - expl[-1][1]) < 1e-12


def test_checkpoint_and_resume(tmp_path):
    from rlpoker_b200.cfr.cfr import cfr
    from rlpoker_b200.experiment import Experiment, SolverCheckpoint
    game = make_game('leduc2')
    full, _, _ = cfr(Experiment('full', base_save_path=str(tmp_path)), game, num_iters=41, use_chance_sampling=False,
                     verbose=False, immediate_regret=False, keep_strategies=0)
    exp = Experiment('parts', base_save_path=str(tmp_path))
    cfr(exp, game, num_iters=21, use_chance_sampling=False, verbose=False, immediate_regret=False, keep_strategies=0,
        checkpoint=True)
    assert SolverCheckpoint(exp).exists()
    resumed, expl, hist = cfr(exp, game, num_iters=41, use_chance_sampling=False, verbose=False, immediate_regret=False,
                              keep_strategies=0, resume=True)
    assert np.array_equal(resumed.sigma_array, full.sigma_array)          # bit-identical continuation
    assert [t for t, _ in expl] == [30, 40] and len(hist) == 41


def test_resume_with_default_arguments(tmp_path):
    """Resume with immediate_regret=True and the strategy history on (the defaults): the immediate-regret series, the
    nodes-touched counter, the history indexing and best_strategy.pkl must continue as if never interrupted."""
    import pickle
    from rlpoker_b200.cfr.cfr import cfr
    from rlpoker_b200.cfr.cfr_metrics import compute_immediate_regret
    from rlpoker_b200.experiment import Experiment

    class Log:
        def __init__(self):
            self.rows = []

        def log(self, data, global_step=None):
            self.rows.append((global_step, dict(data)))

    game = make_game('leduc2')
    w_full, w_a, w_b = Log(), Log(), Log()
    full, expl_full, hist_full = cfr(Experiment('full', base_save_path=str(tmp_path)), game, num_iters=41,
                                     use_chance_sampling=False, verbose=False, experiment_writer=w_full)
    exp = Experiment('parts', base_save_path=str(tmp_path))
    cfr(exp, game, num_iters=21, use_chance_sampling=False, verbose=False, checkpoint=True, experiment_writer=w_a)
    with open(os.path.join(exp.save_path, 'best_strategy.pkl'), 'rb') as f:
        best_before = pickle.load(f)
    resumed, expl, hist = cfr(exp, game, num_iters=41, use_chance_sampling=False, verbose=False, resume=True,
                              experiment_writer=w_b)
    assert np.array_equal(resumed.sigma_array, full.sigma_array)
    assert hist.immediate_regrets == hist_full.immediate_regrets
    assert w_a.rows + w_b.rows == w_full.rows                 # nodes_touched, exploitability, immediate_regret per step
    assert len(hist) == 41 and hist[40] == hist_full[40] and hist[-1] == hist_full[-1] and hist[21] == hist_full[21]
    with pytest.raises(IndexError, match='resume point'):
        hist[0]
    assert compute_immediate_regret(game, hist)[:2] == compute_immediate_regret(game, hist_full)[:2]
    # exploitability falls over this horizon, so the resumed run must have replaced the pickle with a better strategy
    with open(os.path.join(exp.save_path, 'best_strategy.pkl'), 'rb') as f:
        best_after = pickle.load(f)
    assert expl[-1][1] < expl_full[2][1] and best_after != best_before


def test_command_line_front_end(tmp_path):
    import subprocess
    import sys
    from conftest import ROOT
    out = subprocess.check_output(
        [sys.executable, os.path.join(ROOT, 'examples', 'cfr.py'), '--exp_name', 'cli', '--cfr_algorithm', 'vanilla',
         '--num_iters', '11', '--game_specifier', 'Leduc:values_2:suits_2'], cwd=str(tmp_path), text=True)
    assert 't: 10, nodes touched:' in out and 'Immediate regret:' in out and 'Exploitability of saved strategy:' in out
    assert os.path.exists(os.path.join(str(tmp_path), 'experiments', 'cli', 'best_strategy.pkl'))
    assert os.path.exists(os.path.join(str(tmp_path), 'experiments', 'cli', 'metrics.jsonl'))
    out = subprocess.check_output(
        [sys.executable, os.path.join(ROOT, 'examples', 'cfr.py'), '--exp_name', 'cli2', '--cfr_algorithm', 'external',
         '--num_iters', '201', '--lanes', '64', '--seed', '1', '--no_immediate_regret'], cwd=str(tmp_path), text=True)
    assert 't: 200, nodes touched:' in out


def test_cfr_recursive_entry_point_drives_the_reference_loop(golden_arrays):
    """The reference's inner loop driven by hand through `cfr_recursive` (cfr.py:105-122) gives the reference's numbers."""
    from oracle import oracle as orc
    from rlpoker_b200.cfr import cfr_util
    from rlpoker_b200.cfr.cfr import cfr_recursive, compute_average_strategy
    from rlpoker_b200.extensive_game import Strategy
    from rlpoker_b200.flatten import flatten, strategy_to_array
    game = make_game('leduc2')
    regrets, action_counts = {}, {}
    state = cfr_util.CFRState()
    strategy_t, strategy_t_1 = Strategy.initialise(), Strategy.initialise()
    for t in range(6):
        values = [cfr_recursive(game, game.root, i, t, 1.0, 1.0, 1.0, regrets, action_counts, strategy_t, strategy_t_1,
                                state, use_chance_sampling=False, weight=float(t)) for i in (1, 2)]
        assert values[0] == -values[1]
        strategy_t = strategy_t_1.copy()
    ft = flatten(game)
    got_R = np.array([regrets[k][a] for i, k in enumerate(ft.info_set_keys) for a in ft.actions_of(i)])
    assert np.array_equal(got_R, golden_arrays['vanilla_linear_leduc2_T6_R'])
    avg = strategy_to_array(ft, compute_average_strategy(action_counts))[0]
    assert np.array_equal(avg, golden_arrays['vanilla_linear_leduc2_T6_avg'])
    assert state.nodes_touched == 2 * 6 * ft.num_nodes
    with pytest.raises(ValueError):
        cfr_recursive(game, game.root.children[next(iter(game.root.children))], 1, 0, 1.0, 1.0, 1.0, regrets, action_counts,
                      strategy_t, strategy_t_1, state)
