"""Run directory, metric sinks, best-strategy saving and solver checkpoints.

Keeps the public names of the reference's ``rlpoker/experiment.py`` (``Experiment`` :11-28, ``ExperimentWriter`` :31-35,
``WANDBExperimentWriter`` :38-47, ``ExperimentSummaryWriter`` :50-64, ``StrategySaver`` :67-100) because drivers and
user code construct them, but none of this is on the hot path: it is host-side bookkeeping around the device solver.

What is different on purpose
  * heavy optional dependencies (``wandb``, ``tensorboardX``) are imported by the writer that needs them, not at
    module import, and ``float('inf')`` replaces the ``np.float`` alias that numpy >= 1.24 dropped (experiment.py:71),
    so the module imports on a bare machine;
  * ``JSONLinesExperimentWriter`` needs nothing but the standard library;
  * ``SolverCheckpoint`` adds what the reference cannot do at all -- stop and resume a solve (it only ever saved
    the best average strategy, never the regrets).
"""
import abc
import json
import os
import pathlib
import time

import numpy as np


class Experiment:
    """A named directory ``<base_save_path>/<exp_name>`` (created on construction)."""

    def __init__(self, exp_name=None, base_save_path='experiments'):
        if exp_name is None:
            exp_name = time.strftime("%Y-%m-%d-%H:%M:%S", time.gmtime())
        self._exp_name = exp_name
        self._base_save_path = base_save_path
        pathlib.Path(self.save_path).mkdir(parents=True, exist_ok=True)

    exp_name = property(lambda self: self._exp_name)
    save_path = property(lambda self: os.path.join(self._base_save_path, self._exp_name))


class ExperimentWriter(abc.ABC):
    """Sink for ``{metric name: value}`` dictionaries, one call per evaluation step."""

    @abc.abstractmethod
    def log(self, data, global_step=None):
        ...


class JSONLinesExperimentWriter(ExperimentWriter):
    """Appends one JSON object per call to ``<save_path>/metrics.jsonl``."""

    def __init__(self, experiment):
        self.path = pathlib.Path(experiment.save_path) / 'metrics.jsonl'

    def log(self, data, global_step=None):
        record = {'global_step': global_step}
        record.update(data)
        with self.path.open('a') as sink:
            sink.write(json.dumps(record) + '\n')


class WANDBExperimentWriter(ExperimentWriter):
    """Weights & Biases run tagged with the experiment name (same defaults as the reference)."""

    def __init__(self, experiment, project="rlpoker", entity="cgnicholls", tags=None):
        import wandb
        self._wandb = wandb
        self._tags = {**(tags or {}), 'exp_name': experiment.exp_name}
        wandb.init(project=project, entity=entity, tags=self._tags)

    def log(self, data, global_step=None):
        self._wandb.log(data, step=global_step)


class ExperimentSummaryWriter(ExperimentWriter):
    """tensorboardX scalars under ``<save_path>/logs``; like the reference it refuses to reuse that directory."""

    def __init__(self, experiment, flush_secs=120):
        from tensorboardX import SummaryWriter
        logdir = os.path.join(experiment.save_path, 'logs')
        print("To run tensorboard: tensorboard --logdir {}".format(os.path.join(os.getcwd(), experiment.save_path)))
        if os.path.exists(logdir):
            raise ValueError(f"Experiment already exists at: {logdir}.")
        self._save_path = experiment.save_path
        self._writer = SummaryWriter(logdir=logdir, flush_secs=flush_secs)

    def log(self, data, global_step=None):
        for tag in data:
            self._writer.add_scalar(tag, data[tag], global_step=global_step)


class StrategySaver:
    """Per-step strategy files plus "best exploitability so far"; subclasses supply the (de)serialisation."""

    def __init__(self, experiment):
        self._save_path = experiment.save_path
        self._best_exploitability, self._best_exploitability_t = float('inf'), 0

    # -- to be provided by subclasses (rlpoker_b200.cfr.cfr.CFRStrategySaver pickles) --
    def _save_strategy(self, strategy, file_name):
        raise NotImplementedError()

    def _load_strategy(self, file_name):
        raise NotImplementedError()

    # -- file naming -------------------------------------------------------------------
    def file_name(self, t):
        return os.path.join(self._save_path, 'strategy_%05d.pkl' % t)

    @property
    def file_name_best_exploitability(self):
        return os.path.join(self._save_path, 'best_strategy.pkl')

    # -- operations ----------------------------------------------------------------------
    def save_strategy_at_step(self, strategy, t):
        self._save_strategy(strategy, self.file_name(t))

    def load_strategy_from_step(self, t):
        return self._load_strategy(self.file_name(t))

    def save_best_strategy(self, strategy, t, exploitability):
        if not exploitability < self._best_exploitability:
            return
        self._best_exploitability, self._best_exploitability_t = exploitability, t
        self._save_strategy(strategy, self.file_name_best_exploitability)

    def load_best_strategy(self):
        return self._load_strategy(self.file_name_best_exploitability)


class SolverCheckpoint:
    """Stop / resume for a device solver: regrets, strategy sums, current strategy, the running immediate-regret totals
    (cfr_metrics.py:71-76), membership flags, iteration counter, nodes touched and the sampler seed in one ``.npz``
    (flat arrays in canonical pair order, so the file is independent of the device layout).  ``extra`` carries what the
    driver needs to continue its own series (best exploitability so far, the immediate-regret series)."""

    FILE = 'solver_state.npz'

    def __init__(self, experiment_or_dir):
        root = getattr(experiment_or_dir, 'save_path', experiment_or_dir)
        self.path = os.path.join(root, self.FILE)

    def save(self, solver, seed=None, **extra):
        from rlpoker_b200._lib import RLP_CUM_IMM_REGRET, RLP_REGRET, RLP_SIGMA, RLP_STRAT_SUM
        tmp = self.path + '.tmp.npz'
        np.savez(tmp, regrets=solver.read(RLP_REGRET), sums=solver.read(RLP_STRAT_SUM), sigma=solver.read(RLP_SIGMA),
                 cum_imm=solver.read(RLP_CUM_IMM_REGRET), flags=solver.flags(), t=np.int64(solver.t),
                 nodes_touched=np.uint64(solver.nodes_touched), seed=np.int64(-1 if seed is None else seed),
                 extra=json.dumps(extra))
        os.replace(tmp, self.path)
        return self.path

    def exists(self):
        return os.path.exists(self.path)

    def load(self, solver):
        """Restore ``solver`` in place; returns (seed or None, extra dict)."""
        from rlpoker_b200._lib import RLP_CUM_IMM_REGRET, RLP_REGRET, RLP_SIGMA, RLP_STRAT_SUM
        with np.load(self.path) as state:
            if state['regrets'].shape != (solver.tree.num_pairs,):
                raise ValueError("checkpoint is for a different game (%d pairs, solver has %d)"
                                 % (state['regrets'].shape[0], solver.tree.num_pairs))
            solver.write(RLP_REGRET, state['regrets'])
            solver.write(RLP_STRAT_SUM, state['sums'])
            solver.write(RLP_SIGMA, state['sigma'])
            solver.write_flags(state['flags'])
            if 'cum_imm' in state:
                solver.write(RLP_CUM_IMM_REGRET, state['cum_imm'])
            if 'nodes_touched' in state:
                solver.set_nodes_touched(int(state['nodes_touched']))
            solver.t = int(state['t'])
            seed = int(state['seed'])
            extra = json.loads(str(state['extra']))
        return (None if seed < 0 else seed), extra
