"""Third golden set: random non-poker trees.  TEST INFRASTRUCTURE.

Builds the random extensive games of tests/conftest.py::random_game with the REFERENCE's own node classes
(rlpoker/extensive_game.py:196-264), runs the unmodified reference on them (vanilla CFR through cfr_recursive,
cfr.py:158-255; best responses, best_response.py:103-141) and records the results.  These trees have up to five
actions per node, uneven chance probabilities, float payoffs, leaves at several depths and, for odd seeds, payoffs
that are not zero-sum -- none of which the poker goldens exercise.
Output: tests/golden/reference_random.json + reference_random.npz.

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.gen_golden_random
"""
import json
import os
import sys

import numpy as np

from oracle import flat
from oracle.gen_golden import OUT, import_reference, run_external, run_vanilla, sha, tree_record

SEEDS = list(range(8))
CFR_T = 5
SAMPLED_T = 80           # iterations of the numpy-seeded chance-sampled / external-sampled runs (util.py:6-27 sampler)


def zero_sum(seed):
    return seed % 2 == 0


def main():
    ref = import_reference()
    sys.path.insert(0, os.path.join(os.path.dirname(OUT)))          # tests/ for conftest.random_game
    from conftest import random_game
    J = {'generated_by': 'oracle/gen_golden_random.py', 'cfr_T': CFR_T, 'games': {}}
    arrays = {}
    for seed in SEEDS:
        game = random_game(seed, zero_sum=zero_sum(seed), classes=(ref.eg.ExtensiveGame, ref.eg.ExtensiveGameNode))
        t = flat.flatten_game(game)
        full, _ = game.complete_strategy_uniformly(ref.eg.Strategy.initialise())
        rec = {'tree': tree_record(t, full=False),
               'uniform_br': [ref.br.compute_best_response(game, full, p)[0] for p in (1, 2)],
               'uniform_value': list(game.expected_value_exact(ref.eg.Strategy(full.strategy), ref.eg.Strategy(full.strategy)))}
        sn = run_vanilla(ref, game, t, [CFR_T])[CFR_T]
        rec['cfr'] = {'T': CFR_T, 'exploitability': sn['exploitability'], 'nodes_touched': sn['nodes_touched'],
                      'sha_R': sha(sn['R']), 'sha_S': sha(sn['S']), 'sha_avg': sha(sn['avg'])}
        for k in ('R', 'S', 'avg', 'avg_present', 'sigma'):
            arrays['seed%d_%s' % (seed, k)] = sn[k]
        np.random.seed(1000 + seed)
        cs = run_vanilla(ref, game, t, [SAMPLED_T], sampling=True)[SAMPLED_T]
        rec['chance_sampled'] = {'seed': 1000 + seed, 'T': SAMPLED_T, 'nodes_touched': cs['nodes_touched'],
                                 'exploitability': cs['exploitability']}
        for k in ('R', 'S', 'in_R', 'in_S'):
            arrays['seed%d_cs_%s' % (seed, k)] = cs[k]
        np.random.seed(2000 + seed)
        ex = run_external(ref, game, t, SAMPLED_T)
        rec['external'] = {'seed': 2000 + seed, 'T': SAMPLED_T, 'nodes_touched': ex['nodes_touched'],
                           'exploitability': ex['exploitability']}
        for k in ('R', 'alpha', 'in_R', 'in_alpha'):
            arrays['seed%d_ex_%s' % (seed, k)] = ex[k]
        J['games'][str(seed)] = rec
        print(seed, t.num_nodes, t.num_infosets, rec['uniform_br'], sn['exploitability'], flush=True)
    with open(os.path.join(OUT, 'reference_random.json'), 'w') as f:
        json.dump(J, f, indent=1)
    np.savez_compressed(os.path.join(OUT, 'reference_random.npz'), **arrays)


if __name__ == '__main__':
    main()
