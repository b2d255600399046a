"""Second golden set: non-default game parameters.  TEST INFRASTRUCTURE (see oracle/README in DESIGN.md section 2).

Runs the UNMODIFIED reference from /root/reference (same import shim as gen_golden.py) on Leduc variants
(`max_raises`, `raise_amount`, three suits, one suit; rlpoker/games/leduc.py:29-42) and other OneCardPoker
sizes (rlpoker/games/one_card_poker.py:52-57), and records the tree layout, the uniform best responses and a short
vanilla CFR run of each.  Output: tests/golden/reference_variants.json + reference_variants.npz.

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.gen_golden_variants
"""
import json
import os

import numpy as np

from oracle import flat
from oracle.gen_golden import OUT, import_reference, run_vanilla, sha, tree_record

# name -> (kind, args)
VARIANTS = {
    'leduc3_r2_a3': ('leduc', dict(values=3, suits=2, max_raises=2, raise_amount=3)),
    'leduc3_r1_a2': ('leduc', dict(values=3, suits=2, max_raises=1, raise_amount=2)),
    'leduc2_s3': ('leduc', dict(values=2, suits=3, max_raises=4, raise_amount=2)),
    'leduc4_s1': ('leduc', dict(values=4, suits=1, max_raises=4, raise_amount=2)),
    'leduc3_r6_a1': ('leduc', dict(values=3, suits=2, max_raises=6, raise_amount=1)),
    'ocp3': ('ocp', dict(n=3)),
    'ocp6': ('ocp', dict(n=6)),
}
CFR_T = 6


def build(ref, kind, a):
    if kind == 'leduc':
        return ref.leduc.Leduc(ref.card.get_deck(a['values'], a['suits']), max_raises=a['max_raises'],
                               raise_amount=a['raise_amount'])
    return ref.ocp.OneCardPoker.create_game(a['n'])


def main():
    ref = import_reference()
    J = {'generated_by': 'oracle/gen_golden_variants.py', 'variants': {k: list(v) for k, v in VARIANTS.items()},
         'cfr_T': CFR_T, 'games': {}}
    arrays = {}
    for name, (kind, a) in VARIANTS.items():
        game = build(ref, kind, a)
        t = flat.flatten_game(game)
        rec = {'tree': tree_record(t, full=False)}
        full, _ = game.complete_strategy_uniformly(ref.eg.Strategy.initialise())
        rec['uniform_br'] = [ref.br.compute_best_response(game, full, p)[0] for p in (1, 2)]
        rec['uniform_exploitability'] = ref.br.compute_exploitability(game, ref.eg.Strategy.initialise())
        sn = run_vanilla(ref, game, t, [CFR_T])[CFR_T]
        rec['cfr'] = {'T': CFR_T, 'exploitability': sn['exploitability'], 'nodes_touched': sn['nodes_touched'],
                      'sha_R': sha(sn['R']), 'sha_S': sha(sn['S']), 'sha_avg': sha(sn['avg'])}
        arrays[name + '_R'] = sn['R']
        arrays[name + '_avg'] = sn['avg']
        arrays[name + '_avg_present'] = sn['avg_present']
        J['games'][name] = rec
        print(name, t.num_nodes, t.num_infosets, rec['uniform_exploitability'], sn['exploitability'], flush=True)
    with open(os.path.join(OUT, 'reference_variants.json'), 'w') as f:
        json.dump(J, f, indent=1)
    np.savez_compressed(os.path.join(OUT, 'reference_variants.npz'), **arrays)


if __name__ == '__main__':
    main()
