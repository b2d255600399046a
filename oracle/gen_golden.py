"""ORACLE (test infrastructure): generate tests/golden/* by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.gen_golden [--quick]

The reference imports after a two-line shim: a stub ``tensorboardX`` module
and ``np.float = float`` (rlpoker/experiment.py:8,71).  Nothing else is altered; every number
written here comes out of the reference's own functions.  The GPU box never sees
/root/reference, so the outputs are committed.
"""
import argparse
import hashlib
import io
import json
import os
import sys
import time
import types
import contextlib

import numpy as np

REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def import_reference():
    if 'tensorboardX' not in sys.modules:
        stub = types.ModuleType('tensorboardX')
        stub.SummaryWriter = type('SummaryWriter', (), {'__init__': lambda self, *a, **k: None})
        sys.modules['tensorboardX'] = stub
    if 'wandb' not in sys.modules:
        try:
            import wandb  # noqa: F401
        except Exception:
            sys.modules['wandb'] = types.ModuleType('wandb')
    if not hasattr(np, 'float'):
        np.float = float
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import rlpoker.cfr.cfr as ref_cfr
    import rlpoker.cfr.external_cfr as ref_ext
    import rlpoker.cfr.cfr_util as ref_util
    import rlpoker.cfr.cfr_metrics as ref_metrics
    import rlpoker.best_response as ref_br
    import rlpoker.extensive_game as ref_eg
    import rlpoker.util as ref_sample
    from rlpoker.games import leduc, one_card_poker, rock_paper_scissors, card
    from rlpoker.experiment import Experiment
    return types.SimpleNamespace(cfr=ref_cfr, ext=ref_ext, util=ref_util, metrics=ref_metrics, br=ref_br,
                                 eg=ref_eg, sample=ref_sample, leduc=leduc, ocp=one_card_poker,
                                 rps=rock_paper_scissors, card=card, Experiment=Experiment)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def tree_record(t, full):
    rec = {
        'num_nodes': int(t.num_nodes), 'num_infosets': int(t.num_infosets), 'num_pairs': int(t.act_off[-1]),
        'level_off': [int(x) for x in t.level_off],
        'sha_parent': sha(t.parent), 'sha_player': sha(t.player), 'sha_infoset': sha(t.infoset),
        'sha_cprob': sha(t.cprob), 'sha_u1': sha(t.u1), 'sha_u2': sha(t.u2), 'sha_hidden': sha(t.hidden),
        'sha_first_child': sha(t.first_child),
        'sha_keys': hashlib.sha256(repr(t.keys).encode()).hexdigest(),
        'sha_labels': hashlib.sha256(repr([repr(t.labels[l]) for l in t.label[1:]]).encode()).hexdigest(),
        'labels': [repr(x) for x in t.labels],
        'keys_head': [repr(k) for k in t.keys[:12]],
    }
    if full:
        rec.update(parent=t.parent.tolist(), player=t.player.tolist(), infoset=t.infoset.tolist(),
                   cprob=t.cprob.tolist(), u1=t.u1.tolist(), u2=t.u2.tolist(), hidden=t.hidden.tolist(),
                   keys=[repr(k) for k in t.keys])
    return rec


def dict_to_array(t, d, default=0.0):
    """{key: mapping action->float} -> array[Q] in canonical order + present flags."""
    out = np.full(int(t.act_off[-1]), default, np.float64)
    present = np.zeros(t.num_infosets, np.uint8)
    for i, key in enumerate(t.keys):
        if key in d:
            h = int(t.iset_nodes[t.iset_node_off[i]])
            acts = [t.labels[t.label[j]] for j in range(t.first_child[h], t.first_child[h + 1])]
            lo = int(t.act_off[i])
            for j, a in enumerate(acts):
                out[lo + j] = d[key][a]
            present[i] = 1
    return out, present


def run_vanilla(ref, game, t, checkpoints, linear=False, sampling=False):
    """Drive cfr_recursive exactly as cfr() does (cfr.py:105-122) and snapshot at checkpoints."""
    regrets, counts = {}, {}
    state = ref.util.CFRState()
    s_t, s_t1 = ref.eg.Strategy.initialise(), ref.eg.Strategy.initialise()
    snaps = {}
    for it in range(max(checkpoints)):
        w = it if linear else 1.0
        for i in (1, 2):
            ref.cfr.cfr_recursive(game, game.root, i, it, 1.0, 1.0, 1.0, regrets, counts, s_t, s_t1, state,
                                  use_chance_sampling=sampling, weight=w)
        s_t = s_t1.copy()
        T = it + 1
        if T in checkpoints:
            avg = ref.cfr.compute_average_strategy(counts)
            a, present = dict_to_array(t, avg.strategy)
            snaps[T] = {
                'avg': a, 'avg_present': present,
                'R': dict_to_array(t, regrets)[0], 'S': dict_to_array(t, counts)[0],
                'in_R': dict_to_array(t, regrets)[1], 'in_S': dict_to_array(t, counts)[1],
                'sigma': dict_to_array(t, s_t.strategy)[0], 'in_sigma': dict_to_array(t, s_t.strategy)[1],
                'exploitability': ref.br.compute_exploitability(game, avg),
                'nodes_touched': state.nodes_touched,
            }
    return snaps


def run_external(ref, game, t, T):
    """external_cfr.py:47-56 loop, keeping regrets/alpha for inspection."""
    regrets = {}
    s_t, s_t1 = ref.eg.Strategy.initialise(), ref.eg.Strategy.initialise()
    state = ref.util.CFRState()
    average = ref.util.AverageStrategy(game)
    for it in range(T):
        for p in (1, 2):
            ref.ext.external_sampling_cfr_recursive(game, game.root, p, regrets, s_t, s_t1, state)
        s_t = s_t1.copy()
        ref.util.update_average_strategy(game, average, s_t)
    strat = average.compute_strategy()
    return {
        'R': dict_to_array(t, regrets)[0], 'in_R': dict_to_array(t, regrets)[1],
        'alpha': dict_to_array(t, average.alpha)[0], 'in_alpha': dict_to_array(t, average.alpha)[1],
        'sigma': dict_to_array(t, s_t.strategy)[0], 'in_sigma': dict_to_array(t, s_t.strategy)[1],
        'avg': dict_to_array(t, strat.strategy)[0],
        'exploitability': ref.br.compute_exploitability(game, strat),
        'nodes_touched': state.nodes_touched,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--quick', action='store_true', help='short horizons only (seconds instead of minutes)')
    args = ap.parse_args()
    from oracle import flat
    ref = import_reference()
    os.makedirs(OUT, exist_ok=True)
    J = {'generated_by': 'oracle/gen_golden.py', 'numpy': np.__version__, 'python': sys.version.split()[0]}
    arrays = {}
    Card = ref.card.Card

    def leduc(n):
        return ref.leduc.Leduc(ref.card.get_deck(n, 2))

    games = {
        'rps': ref.rps.create_rock_paper_scissors(),
        'ocp4': ref.ocp.OneCardPoker.create_game(4),
        'leduc2': leduc(2),
        'leduc3': leduc(3),
    }
    if not args.quick:
        games['leduc4'] = leduc(4)
    trees = {k: flat.flatten_game(g) for k, g in games.items()}

    # 1. tree layout
    J['trees'] = {k: tree_record(trees[k], full=k in ('rps', 'ocp4')) for k in trees}

    # 2. reference unit-test KATs restated as data (rlpoker/cfr/tests/test_cfr.py:10-55)
    AF = ref.eg.ActionFloat
    rm = ref.util.compute_regret_matching
    J['regret_matching'] = [
        {'r': [0.5, 1.0, -1.0], 'eps': 0.0, 'hr': False, 'out': list(rm(AF({1: 0.5, 2: 1.0, 3: -1.0}), epsilon=0.0).values())},
        {'r': [0.5, 1.0, -1.0], 'eps': 1e-7, 'hr': False, 'out': list(rm(AF({1: 0.5, 2: 1.0, 3: -1.0}), epsilon=1e-7).values())},
        {'r': [-1.0, -3.0, -1.0], 'eps': 1e-7, 'hr': False, 'out': list(rm(AF({1: -1.0, 2: -3.0, 3: -1.0})).values())},
        {'r': [0.0, 0.0, 0.0], 'eps': 1e-7, 'hr': False, 'out': list(rm(AF({1: 0.0, 2: 0.0, 3: 0.0})).values())},
        {'r': [-3.0, -1.0, -4.0], 'eps': 1e-6, 'hr': True,
         'out': list(rm(AF({1: -3.0, 2: -1.0, 3: -4.0}), epsilon=1e-6, highest_regret=True).values())},
        {'r': [1e-9, 3.0, 2.5e-8], 'eps': 1e-7, 'hr': False, 'out': list(rm(AF({1: 1e-9, 2: 3.0, 3: 2.5e-8})).values())},
        {'r': [0.1, 0.2], 'eps': 1e-7, 'hr': False, 'out': list(rm(AF({1: 0.1, 2: 0.2})).values())},
    ]
    rng = np.random.RandomState(7)
    for _ in range(40):
        n = int(rng.randint(2, 6))
        r = (rng.standard_normal(n) * 10 ** rng.uniform(-9, 2)).tolist()
        J['regret_matching'].append({'r': r, 'eps': 1e-7, 'hr': False,
                                     'out': list(rm(AF(dict(enumerate(r)))).values())})
    J['normalise_probs'] = [
        {'p': p, 'eps': e, 'out': list(ref.util.normalise_probs(AF(dict(enumerate(p))), epsilon=e).values())}
        for p, e in [([1.0, 2.0], 1e-7), ([1.0, 0.0], 1e-7), ([1.0, 0.0], 0.0), ([0.3, 0.3, 0.4], 1e-7)]]

    # 3. sampler (rlpoker/tests/test_util.py:12-31 + longer sequences)
    np.random.seed(0)
    seq_a = [ref.sample.sample_action({1: 0.2, 2: 0.8}) for _ in range(20)]
    np.random.seed(0)
    seq_b = [ref.sample.sample_action({'R': 0.2, 'P': 0.7, 'S': 0.1}) for _ in range(20)]
    np.random.seed(3)
    p10 = {i: (i + 1.0) for i in range(10)}
    seq_c = [ref.sample.sample_action(p10) for _ in range(200)]
    J['sample_action'] = {'a': seq_a, 'b': seq_b, 'c_seed': 3, 'c_probs': list(p10.values()), 'c': seq_c}

    # 4. exploitability / expected value of the uniform strategy
    J['uniform'] = {}
    for k, g in games.items():
        empty = ref.eg.Strategy.initialise()
        full, _ = g.complete_strategy_uniformly(empty)
        v1, bs1 = ref.br.compute_best_response(g, full, 1)
        v2, bs2 = ref.br.compute_best_response(g, full, 2)
        t = trees[k]
        amax = np.full(t.num_infosets, -1, np.int32)
        for bs in (bs1, bs2):
            arr, pres = dict_to_array(t, bs)
            for i in range(t.num_infosets):
                if pres[i]:
                    lo = int(t.act_off[i])
                    amax[i] = int(np.argmax(arr[lo:lo + int(t.nact[i])]))
        arrays[k + '_uniform_br_argmax'] = amax
        J['uniform'][k] = {
            'exploitability': ref.br.compute_exploitability(g, empty), 'br1': v1, 'br2': v2,
            'expected_value_exact': list(g.expected_value_exact(ref.eg.Strategy(full.strategy), ref.eg.Strategy(full.strategy))),
        }

    # 5. RPS expected values with asymmetric strategies (test_extensive_game.py:9-36, test_cfr_metrics.py:11-99)
    g = games['rps']
    i1, i2 = g.info_set_ids[g.root], g.info_set_ids[g.root.children['R']]
    s = ref.eg.Strategy({i1: AF({'R': 0.5, 'P': 0.2, 'S': 0.3}), i2: AF({'R': 0.2, 'P': 0.3, 'S': 0.5})})
    J['rps_expected_value_exact'] = list(g.expected_value_exact(s, s))
    s1 = ref.eg.Strategy({i1: AF({'R': 0.5, 'P': 0.3, 'S': 0.2})})
    s2 = ref.eg.Strategy({i2: AF({'R': 0.3, 'P': 0.3, 'S': 0.4})})
    eu1, eu2 = ref.metrics.compute_expected_utility(g, s1, s2)
    J['rps_expected_utility'] = {'root_u1': eu1[g.get_node(())], 'u2_after': [eu2[g.get_node((a,))] for a in 'RPS']}

    # 6. vanilla CFR trajectories (cfr.py:105-122 driven through cfr_recursive)
    plan = {
        'rps': [1, 2, 5, 20], 'ocp4': [1, 5, 30], 'leduc2': [1, 5, 20],
        'leduc3': [1, 2, 5, 10] if args.quick else [1, 2, 5, 8, 10, 20, 50, 100],
    }
    J['vanilla'] = {}
    for k, cps in plan.items():
        t0 = time.time()
        snaps = run_vanilla(ref, games[k], trees[k], cps)
        J['vanilla'][k] = {}
        for T, sn in snaps.items():
            J['vanilla'][k][str(T)] = {'exploitability': sn['exploitability'], 'nodes_touched': sn['nodes_touched'],
                                       'sha_avg': sha(sn['avg']), 'sha_R': sha(sn['R']), 'sha_S': sha(sn['S'])}
            if T in (cps[-1], 10):
                for name in ('avg', 'R', 'S', 'sigma', 'avg_present'):
                    arrays['vanilla_%s_T%d_%s' % (k, T, name)] = sn[name]
        print('vanilla', k, cps, '%.1fs' % (time.time() - t0), flush=True)
    # linear weighting (cfr.py:106)
    snaps = run_vanilla(ref, games['leduc2'], trees['leduc2'], [6], linear=True)
    J['vanilla_linear'] = {'leduc2': {'6': {'exploitability': snaps[6]['exploitability'], 'sha_avg': sha(snaps[6]['avg'])}}}
    arrays['vanilla_linear_leduc2_T6_avg'] = snaps[6]['avg']
    arrays['vanilla_linear_leduc2_T6_R'] = snaps[6]['R']

    # 7. the public drivers end to end
    def quiet(fn, *a, **k):
        with contextlib.redirect_stdout(io.StringIO()) as buf:
            out = fn(*a, **k)
        return out, buf.getvalue()

    class Writer:
        def __init__(self):
            self.rows = []

        def log(self, data, global_step=None):
            self.rows.append((global_step, dict(data)))

    tmp = '/tmp/rlpoker_golden_experiments'
    exp = ref.Experiment('g_vanilla', base_save_path=tmp)
    w = Writer()
    (avg, expl, strategies), _ = quiet(ref.cfr.cfr, exp, games['leduc3'], num_iters=11, use_chance_sampling=False,
                                       experiment_writer=w)
    imm, cum_imm, _ = ref.metrics.compute_immediate_regret(games['leduc3'], strategies)
    J['cfr_driver_leduc3'] = {'exploitabilities': [[int(a), b] for a, b in expl], 'log': [[s, d] for s, d in w.rows],
                              'immediate_regret_T11': imm, 'cumulative_immediate_regrets': cum_imm,
                              'sha_avg': sha(dict_to_array(trees['leduc3'], avg.strategy)[0])}
    arrays['cfr_driver_leduc3_avg'] = dict_to_array(trees['leduc3'], avg.strategy)[0]
    key = (Card(0, 0), -1, 1, 1, Card(1, 0), 1, 2, 2)
    J['cfr_driver_leduc3']['sample_entry'] = {'key': repr(key), 'value': {str(a): v for a, v in avg[key].items()}}

    # 8. chance-sampled CFR with numpy's legacy global RNG
    np.random.seed(0)
    Tcs = 200 if args.quick else 2000
    t0 = time.time()
    snaps = run_vanilla(ref, games['leduc3'], trees['leduc3'], [Tcs], sampling=True)
    sn = snaps[Tcs]
    J['chance_sampled_leduc3'] = {'seed': 0, 'T': Tcs, 'exploitability': sn['exploitability'],
                                  'nodes_touched': sn['nodes_touched'], 'sha_R': sha(sn['R']), 'sha_S': sha(sn['S']),
                                  'num_present': int(sn['avg_present'].sum())}
    for name in ('avg', 'R', 'S', 'in_R', 'in_S', 'in_sigma', 'sigma'):
        arrays['chance_leduc3_%s' % name] = sn[name]
    print('chance-sampled %.1fs' % (time.time() - t0), flush=True)
    np.random.seed(5)
    sn = run_vanilla(ref, games['ocp4'], trees['ocp4'], [300], sampling=True)[300]
    J['chance_sampled_ocp4'] = {'seed': 5, 'T': 300, 'exploitability': sn['exploitability'], 'sha_R': sha(sn['R']),
                                'sha_S': sha(sn['S']), 'nodes_touched': sn['nodes_touched']}
    arrays['chance_ocp4_R'] = sn['R']
    arrays['chance_ocp4_S'] = sn['S']

    # 9. external sampling
    np.random.seed(0)
    Tex = 60 if args.quick else 600
    t0 = time.time()
    ex = run_external(ref, games['leduc3'], trees['leduc3'], Tex)
    J['external_leduc3'] = {'seed': 0, 'T': Tex, 'exploitability': ex['exploitability'],
                            'nodes_touched': ex['nodes_touched'], 'sha_R': sha(ex['R']), 'sha_alpha': sha(ex['alpha']),
                            'n_in_R': int(ex['in_R'].sum()), 'n_in_alpha': int(ex['in_alpha'].sum())}
    for name in ('R', 'alpha', 'in_R', 'in_alpha', 'sigma', 'in_sigma', 'avg'):
        arrays['external_leduc3_%s' % name] = ex[name]
    print('external %.1fs' % (time.time() - t0), flush=True)
    np.random.seed(11)
    ex = run_external(ref, games['rps'], trees['rps'], 50)
    J['external_rps'] = {'seed': 11, 'T': 50, 'exploitability': ex['exploitability'], 'R': ex['R'].tolist(),
                         'alpha': ex['alpha'].tolist(), 'nodes_touched': ex['nodes_touched']}
    if not args.quick:
        np.random.seed(0)
        exp = ref.Experiment('g_external', base_save_path=tmp)
        (avg, expl, _), _ = quiet(ref.ext.external_sampling_cfr, exp, games['leduc3'], num_iters=401,
                                  experiment_writer=Writer())
        J['external_driver_leduc3'] = {'seed': 0, 'exploitabilities': [[int(a), b] for a, b in expl]}

    # 10. the chance-sampled driver end to end (cfr.py:64-152 with use_chance_sampling=True)
    if not args.quick:
        np.random.seed(0)
        exp = ref.Experiment('g_chance', base_save_path=tmp)
        (avg, expl, _), _ = quiet(ref.cfr.cfr, exp, games['leduc3'], num_iters=101, use_chance_sampling=True,
                                  experiment_writer=Writer())
        J['chance_driver_leduc3'] = {'seed': 0, 'exploitabilities': [[int(a), b] for a, b in expl]}

    with open(os.path.join(OUT, 'reference_golden.json'), 'w') as f:
        json.dump(J, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(OUT, 'reference_arrays.npz'), **arrays)
    print('wrote', OUT)


if __name__ == '__main__':
    main()
