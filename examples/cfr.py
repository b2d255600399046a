"""Command-line front end with the reference's flags (examples/cfr.py:19-29):

    python -m examples.cfr --exp_name run1 --cfr_algorithm vanilla  --game_specifier Leduc:values_3:suits_2
    python -m examples.cfr --exp_name run2 --cfr_algorithm external --num_iters 2000

Differences: the game name in --game_specifier is case-insensitive (so the default builds), metrics go to
<experiment>/metrics.jsonl unless --wandb is given (the reference always opens a W&B run), and
--lanes / --seed expose the batched sampled traversals.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from rlpoker_b200.best_response import compute_exploitability
from rlpoker_b200.cfr import cfr, cfr_metrics, external_cfr
from rlpoker_b200.experiment import Experiment, JSONLinesExperimentWriter, WANDBExperimentWriter
from rlpoker_b200.games.game_builder import buildExtensiveGame

if __name__ == "__main__":
    cfr_algorithms = ['vanilla', 'external']
    parser = argparse.ArgumentParser()
    parser.add_argument('--exp_name', required=True, type=str, help='The name of the experiment.')
    parser.add_argument('--num_iters', default=10000, type=int, help='The number of iterations to run CFR for.')
    parser.add_argument('--game_specifier', default='leduc:values_3:suits_2')
    parser.add_argument('--use_chance_sampling', action='store_true',
                        help="Pass this option to use chance sampling. By default, we don't use chance sampling.")
    parser.add_argument('--cfr_algorithm', choices=cfr_algorithms, required=True,
                        help='Which cfr algorithm to use. Choose between vanilla and external.')
    parser.add_argument('--wandb', action='store_true', help='log to Weights & Biases like the reference')
    parser.add_argument('--lanes', type=int, default=1, help='sampled traversal pairs per iteration')
    parser.add_argument('--seed', type=int, default=None, help='Philox seed of the sampled variants')
    args = parser.parse_args()

    game = buildExtensiveGame(spec=args.game_specifier)
    experiment = Experiment(args.exp_name)
    experiment_writer = WANDBExperimentWriter(experiment) if args.wandb else JSONLinesExperimentWriter(experiment)

    if args.cfr_algorithm == 'vanilla':
        strategy, exploitabilities, strategies = cfr.cfr(
            experiment, game, num_iters=args.num_iters, use_chance_sampling=args.use_chance_sampling,
            experiment_writer=experiment_writer, seed=args.seed, lanes=args.lanes)
    else:
        strategy, exploitabilities, strategies = external_cfr.external_sampling_cfr(
            experiment, game, num_iters=args.num_iters, experiment_writer=experiment_writer, seed=args.seed,
            lanes=args.lanes)

    immediate_regret, _, _ = cfr_metrics.compute_immediate_regret(game, strategies)
    print("Immediate regret: {}".format(immediate_regret))
    exploitability = compute_exploitability(game, strategy)
    print("Exploitability of saved strategy: {}".format(exploitability))
