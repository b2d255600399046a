"""Fourth golden set: the reference's seed distribution of exploitability-vs-iteration for the sampled variants.
TEST INFRASTRUCTURE.

For numpy seeds 0..15 runs the unmodified reference on Leduc-3 (`Leduc(get_deck(3, 2))`):
  * chance-sampled CFR exactly as cfr() drives it (cfr.py:105-122, use_chance_sampling=True), exploitability of
    compute_average_strategy (cfr.py:32-41) at T = 250, 500, 1000, 2000;
  * external-sampling CFR as external_sampling_cfr() drives it (external_cfr.py:47-56), exploitability of
    AverageStrategy.compute_strategy (cfr_util.py:95-106) at T = 100, 200, 400.
Output: tests/golden/reference_seed_curves.json.  About five minutes on eight cores.

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.gen_golden_curves
"""
import json
import multiprocessing as mp
import os

import numpy as np

from oracle.gen_golden import OUT, import_reference

SEEDS = list(range(16))
CHANCE_T = [250, 500, 1000, 2000]
EXTERNAL_T = [100, 200, 400]


def chance_curve(seed):
    ref = import_reference()
    game = ref.leduc.Leduc(ref.card.get_deck(3, 2))
    np.random.seed(seed)
    regrets, counts = {}, {}
    state = ref.util.CFRState()
    s_t, s_t1 = ref.eg.Strategy.initialise(), ref.eg.Strategy.initialise()
    out = []
    for it in range(max(CHANCE_T)):
        for i in (1, 2):
            ref.cfr.cfr_recursive(game, game.root, i, it, 1.0, 1.0, 1.0, regrets, counts, s_t, s_t1, state,
                                  use_chance_sampling=True, weight=1.0)
        s_t = s_t1.copy()
        if it + 1 in CHANCE_T:
            out.append(ref.br.compute_exploitability(game, ref.cfr.compute_average_strategy(counts)))
    return seed, out


def external_curve(seed):
    ref = import_reference()
    game = ref.leduc.Leduc(ref.card.get_deck(3, 2))
    np.random.seed(seed)
    regrets = {}
    s_t, s_t1 = ref.eg.Strategy.initialise(), ref.eg.Strategy.initialise()
    state = ref.util.CFRState()
    average = ref.util.AverageStrategy(game)
    out = []
    for it in range(max(EXTERNAL_T)):
        for p in (1, 2):
            ref.ext.external_sampling_cfr_recursive(game, game.root, p, regrets, s_t, s_t1, state)
        s_t = s_t1.copy()
        ref.util.update_average_strategy(game, average, s_t)
        if it + 1 in EXTERNAL_T:
            out.append(ref.br.compute_exploitability(game, average.compute_strategy()))
    return seed, out


def main():
    with mp.Pool(min(8, os.cpu_count() or 1)) as pool:
        chance = dict(pool.map(chance_curve, SEEDS))
        external = dict(pool.map(external_curve, SEEDS))
    J = {'generated_by': 'oracle/gen_golden_curves.py', 'game': 'Leduc(get_deck(3, 2))', 'seeds': SEEDS,
         'chance_sampled': {'T': CHANCE_T, 'exploitability': [chance[s] for s in SEEDS]},
         'external': {'T': EXTERNAL_T, 'exploitability': [external[s] for s in SEEDS]}}
    with open(os.path.join(OUT, 'reference_seed_curves.json'), 'w') as f:
        json.dump(J, f, indent=1)
    for k in ('chance_sampled', 'external'):
        a = np.array(J[k]['exploitability'])
        print(k, J[k]['T'], 'mean', a.mean(0), 'sd', a.std(0, ddof=1))


if __name__ == '__main__':
    main()
