"""External-sampling MCCFR (Lanctot et al.) with the reference's driver API
(rlpoker/cfr/external_cfr.py:14-79) on the device lanes of csrc/sampled.cu: per iteration one sampled
traversal per player (``lanes`` > 1: a mini-batch), regret matching, then the full-tree AverageStrategy walk
with the new strategy."""
import numpy as np

from rlpoker_b200._lib import RLP_AVERAGE, RLP_EXTERNAL_SAMPLED, RLP_SIGMA
from rlpoker_b200.cfr.cfr import CFRStrategySaver, DeviceStrategy, StrategyHistory
from rlpoker_b200.flatten import flatten, postorder_infoset_order
from rlpoker_b200.solver import DeviceCFR, device_tree


def external_sampling_cfr(experiment, game, num_iters=1000, experiment_writer=None, seed=None, lanes=1,
                          immediate_regret=True, keep_strategies=1 << 30, verbose=True, eval_every=200):
    """Returns (average_strategy, exploitabilities, strategies) like the reference; evaluation every 200
    iterations.  ``experiment_writer=None`` is tolerated (the reference would crash at t == 0)."""
    if seed is None:
        seed = int(np.random.randint(0, 2 ** 31 - 1))
    ft = flatten(game)
    solver = DeviceCFR(device_tree(game))
    saver = CFRStrategySaver(experiment=experiment) if experiment is not None else None
    order = postorder_infoset_order(ft) if immediate_regret else None
    history = StrategyHistory(ft, limit_bytes=int(keep_strategies) if keep_strategies else 0)
    exploitabilities = []
    every_iter = immediate_regret or bool(keep_strategies)

    def average():
        present = (solver.flags() & 2) != 0
        return DeviceStrategy(ft, solver.read(RLP_AVERAGE), present)

    t = 0
    while t < num_iters:
        n = 1 if every_iter else min(num_iters - t, eval_every - t % eval_every if t % eval_every else 1)
        solver.sampled(RLP_EXTERNAL_SAMPLED, n, lanes=lanes, seed=seed)
        t_last = t + n - 1
        if every_iter:
            if keep_strategies:
                history._append(solver.read(RLP_SIGMA), (solver.flags() & 4) != 0)
            else:
                history._count += 1
            if immediate_regret:
                history.immediate_regrets.append(solver.immediate_regret_step(t_last, order))
        else:
            history._count += n
        if t_last % eval_every == 0:
            exploitability = solver.exploitability(RLP_AVERAGE)[0]
            exploitabilities.append((t_last, exploitability))
            if verbose:
                print("t: {}, nodes touched: {}, exploitability: {:.3f} mbb/h".format(
                    t_last, solver.nodes_touched, exploitability * 1000))
            log_data = {'nodes_touched': solver.nodes_touched, 'exploitability_mbbh': exploitability * 1000}
            if immediate_regret:
                log_data['immediate_regret'] = history.immediate_regrets[-1]
                if verbose:
                    print("Immediate regret: {}".format(history.immediate_regrets[-1]))
            if experiment_writer:
                experiment_writer.log(log_data, global_step=t_last)
            if saver is not None:
                saver.save_best_strategy(average(), t_last, exploitability)
        t += n
    return average(), exploitabilities, history


def external_sampling_cfr_recursive(game, node, player, regrets, strategy_t, strategy_t_1, cfr_state):
    """The reference's recursion entry (external_cfr.py:82-173).  A sampled traversal draws from numpy's global RNG in
    the reference and from a counter-based Philox stream here, so a single call cannot reproduce the reference's walk;
    the traversals run as device lanes through ``external_sampling_cfr``."""
    raise NotImplementedError("external-sampling traversals run as device lanes: use external_sampling_cfr(...)")
