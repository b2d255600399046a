"""Command-line front end for the tabular solvers, with the reference's flags (examples/cfr.py:19-29):

    python -m examples.cfr --exp_name run1 --cfr_algorithm vanilla  --game_specifier Leduc:values_3:suits_2
    python -m examples.cfr --exp_name run2 --cfr_algorithm external --num_iters 2000 --lanes 4096

Beyond the reference: the game name in --game_specifier is matched case-insensitively (so the default builds),
``Leduc:values_N`` with N >= 6 is generated natively instead of as Python objects, metrics go to
``<experiment>/metrics.jsonl`` unless --wandb is given (the reference always opens a W&B run), --lanes / --seed expose
the batched sampled traversals and --no_immediate_regret skips the extra sweep per iteration that metric costs.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ALGORITHMS = ('vanilla', 'external')


def build_parser():
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument('--exp_name', required=True, type=str, help='The name of the experiment.')
    p.add_argument('--num_iters', default=10000, type=int, help='The number of iterations to run CFR for.')
    p.add_argument('--game_specifier', default='leduc:values_3:suits_2')
    p.add_argument('--use_chance_sampling', action='store_true',
                   help="Pass this option to use chance sampling. By default, we don't use chance sampling.")
    p.add_argument('--cfr_algorithm', choices=ALGORITHMS, required=True,
                   help='Which cfr algorithm to use. Choose between vanilla and external.')
    p.add_argument('--wandb', action='store_true', help='log to Weights & Biases like the reference')
    p.add_argument('--lanes', type=int, default=1, help='sampled traversal pairs per iteration')
    p.add_argument('--seed', type=int, default=None, help='Philox seed of the sampled variants')
    p.add_argument('--no_immediate_regret', action='store_true', help='skip the immediate-regret metric')
    return p


def build_game(spec):
    from rlpoker_b200.games.game_builder import buildExtensiveGame
    from rlpoker_b200.games.util import parse_params
    name, _, rest = spec.partition(':')
    if name.lower() == 'leduc':
        params = parse_params(rest)
        values, suits = int(params.get('values', 3)), int(params.get('suits', 2))
        if values >= 6:                           # Python node objects get heavy: ~1.6 kB and ~20 us per node
            from rlpoker_b200.games.leduc_native import create_leduc
            return create_leduc(values, suits)
    return buildExtensiveGame(spec)


def main(argv=None):
    from rlpoker_b200.best_response import compute_exploitability
    from rlpoker_b200.cfr import cfr, cfr_metrics, external_cfr
    from rlpoker_b200.experiment import Experiment, JSONLinesExperimentWriter, WANDBExperimentWriter
    args = build_parser().parse_args(argv)
    game = build_game(args.game_specifier)
    experiment = Experiment(args.exp_name)
    writer = (WANDBExperimentWriter if args.wandb else JSONLinesExperimentWriter)(experiment)
    common = dict(num_iters=args.num_iters, experiment_writer=writer, seed=args.seed, lanes=args.lanes,
                  immediate_regret=not args.no_immediate_regret)
    if args.cfr_algorithm == 'vanilla':
        strategy, exploitabilities, strategies = cfr.cfr(experiment, game, use_chance_sampling=args.use_chance_sampling,
                                                         **common)
    else:
        strategy, exploitabilities, strategies = external_cfr.external_sampling_cfr(experiment, game, **common)
    if not args.no_immediate_regret:
        immediate_regret, _, _ = cfr_metrics.compute_immediate_regret(game, strategies)
        print("Immediate regret: {}".format(immediate_regret))
    print("Exploitability of saved strategy: {}".format(compute_exploitability(game, strategy)))
    return strategy, exploitabilities


if __name__ == "__main__":
    main()
