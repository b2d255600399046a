"""Fifth golden set: the metrics of rlpoker/cfr/cfr_metrics.py.  TEST INFRASTRUCTURE.

Runs the UNMODIFIED reference from /root/reference (same import shim as gen_golden.py):

* ``compute_expected_utility`` (cfr_metrics.py:159-254) on rock-paper-scissors with the strategies of the reference's
  own test (rlpoker/cfr/tests/test_cfr_metrics.py:11-99), on OneCardPoker-4 and on Leduc-2 with seeded random
  strategies: the value of EVERY node for both players (compute_expected_utility_recursive's return values, recorded
  through a wrapper that stores what each call returns), in breadth-first node order;
* ``compute_immediate_regret`` (cfr_metrics.py:13-78) on the first 5 vanilla strategies of Leduc-2: the third return
  value (per step, per information set), in canonical information-set order.

Output: tests/golden/reference_metrics.npz (+ .json with the scalar series).

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.gen_golden_metrics
"""
import collections
import json
import os

import numpy as np

from oracle import flat
from oracle.gen_golden import OUT, import_reference


def random_sigma(t, seed):
    rng = np.random.RandomState(seed)
    x = rng.random_sample(int(t.act_off[-1])) + 0.05
    return x / np.repeat(np.add.reduceat(x, t.act_off[:-1]), t.nact)


def strategy_from_array(ref, t, sigma):
    d = {}
    for i, key in enumerate(t.keys):
        h = int(t.iset_nodes[t.iset_node_off[i]])
        acts = [t.labels[t.label[j]] for j in range(t.first_child[h], t.first_child[h + 1])]
        lo = int(t.act_off[i])
        d[key] = ref.eg.ActionFloat({a: float(sigma[lo + k]) for k, a in enumerate(acts)})
    return ref.eg.Strategy(d)


def all_node_values(ref, game, s1, s2):
    """Value of every node: call the reference's recursion at every node of the tree (it returns that node's (v1, v2))."""
    nodes = [game.root]
    queue = collections.deque(nodes)
    while queue:
        nd = queue.popleft()
        for ch in nd.children.values():
            nodes.append(ch)
            queue.append(ch)
    v1, v2 = np.zeros(len(nodes)), np.zeros(len(nodes))
    for i, nd in enumerate(nodes):
        v1[i], v2[i] = ref.metrics.compute_expected_utility_recursive(
            game, nd, s1, s2, collections.defaultdict(float), collections.defaultdict(float), 1.0, 1.0, 1.0)
    return v1, v2


def main():
    ref = import_reference()
    arrays, J = {}, {'generated_by': 'oracle/gen_golden_metrics.py'}
    games = {'rps': ref.rps.create_rock_paper_scissors(), 'ocp4': ref.ocp.OneCardPoker.create_game(4),
             'leduc2': ref.leduc.Leduc(ref.card.get_deck(2, 2))}
    for name, game in games.items():
        t = flat.flatten_game(game)
        for seed in (0, 1):
            sigma = random_sigma(t, 100 + seed)
            if name == 'rps' and seed == 0:          # the strategies of test_cfr_metrics.py
                sigma = np.array([0.5, 0.3, 0.2, 0.3, 0.3, 0.4])
            s = strategy_from_array(ref, t, sigma)
            v1, v2 = all_node_values(ref, game, s, s)
            arrays['%s_s%d_sigma' % (name, seed)] = sigma
            arrays['%s_s%d_v1' % (name, seed)] = v1
            arrays['%s_s%d_v2' % (name, seed)] = v2
            # the dictionaries compute_expected_utility itself returns: only player nodes
            eu1, eu2 = ref.metrics.compute_expected_utility(game, s, s)
            assert all(eu1[n] == v for n, v in eu1.items())
            J['%s_s%d_counts' % (name, seed)] = [len(eu1), len(eu2)]
        print(name, t.num_nodes, flush=True)
    # immediate regret detail on Leduc-2
    game = games['leduc2']
    t = flat.flatten_game(game)
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        _, _, strategies = ref.cfr.cfr(ref.Experiment('golden_metrics', base_save_path='/tmp/rlp_golden_metrics'), game,
                                       num_iters=5, use_chance_sampling=False)
    last, series, detail = ref.metrics.compute_immediate_regret(game, strategies)
    index = {k: i for i, k in enumerate(t.keys)}
    for step, d in enumerate(detail):
        arr = np.full(t.num_infosets, np.nan)
        for k, v in d.items():
            arr[index[k]] = v
        assert not np.isnan(arr).any()
        arrays['leduc2_imm_detail_%d' % step] = arr
    J['leduc2_imm_series'] = series
    J['leduc2_imm_last'] = last
    J['leduc2_imm_order_head'] = [repr(k) for k in list(detail[0].keys())[:6]]
    with open(os.path.join(OUT, 'reference_metrics.json'), 'w') as f:
        json.dump(J, f, indent=1)
    np.savez_compressed(os.path.join(OUT, 'reference_metrics.npz'), **arrays)
    print('wrote', len(arrays), 'arrays')


if __name__ == '__main__':
    main()
