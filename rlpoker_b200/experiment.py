"""Experiment directory, metric writers and best-strategy saving (reference: rlpoker/experiment.py:11-100).

Host-side side effects only.  Differences from the reference, all of them import-time robustness:
``tensorboardX`` and ``wandb`` are optional (imported when a writer that needs them is constructed), and
``float('inf')`` replaces the ``np.float`` alias that numpy >= 1.24 no longer has (experiment.py:71)."""
import abc
import os
import time


class Experiment:
    def __init__(self, exp_name=None, base_save_path='experiments'):
        self._base_save_path = base_save_path
        self._exp_name = exp_name if exp_name is not None else time.strftime("%Y-%m-%d-%H:%M:%S", time.gmtime())
        os.makedirs(self.save_path, exist_ok=True)

    @property
    def save_path(self):
        return os.path.join(self._base_save_path, self._exp_name)

    @property
    def exp_name(self):
        return self._exp_name


class ExperimentWriter(abc.ABC):
    @abc.abstractmethod
    def log(self, data, global_step=None):
        pass


class WANDBExperimentWriter(ExperimentWriter):
    def __init__(self, experiment, project="rlpoker", entity="cgnicholls", tags=None):
        import wandb
        self._wandb = wandb
        self._tags = dict(tags or {})
        self._tags['exp_name'] = experiment.exp_name
        wandb.init(project=project, entity=entity, tags=self._tags)

    def log(self, data, global_step=None):
        self._wandb.log(data, step=global_step)


class ExperimentSummaryWriter(ExperimentWriter):
    """tensorboardX scalars under <save_path>/logs (refuses to reuse a log directory, like the reference)."""

    def __init__(self, experiment, flush_secs=120):
        from tensorboardX import SummaryWriter
        self._save_path = experiment.save_path
        print("To run tensorboard: tensorboard --logdir {}".format(os.path.join(os.getcwd(), self._save_path)))
        logdir = os.path.join(self._save_path, 'logs')
        if os.path.exists(logdir):
            raise ValueError(f"Experiment already exists at: {logdir}.")
        self._writer = SummaryWriter(logdir=logdir, flush_secs=flush_secs)

    def log(self, data, global_step=None):
        for tag, val in data.items():
            self._writer.add_scalar(tag, val, global_step=global_step)


class JSONLinesExperimentWriter(ExperimentWriter):
    """Dependency-free writer: one JSON object per log call in <save_path>/metrics.jsonl."""

    def __init__(self, experiment):
        self._path = os.path.join(experiment.save_path, 'metrics.jsonl')

    def log(self, data, global_step=None):
        import json
        with open(self._path, 'a') as f:
            f.write(json.dumps(dict(data, global_step=global_step)) + '\n')


class StrategySaver:
    def __init__(self, experiment):
        self._save_path = experiment.save_path
        self._best_exploitability = float('inf')
        self._best_exploitability_t = 0

    def file_name(self, t):
        return os.path.join(self._save_path, f'strategy_{t:05d}.pkl')

    def _save_strategy(self, strategy, file_name):
        raise NotImplementedError()

    def _load_strategy(self, file_name):
        raise NotImplementedError()

    def save_strategy_at_step(self, strategy, t):
        self._save_strategy(strategy, self.file_name(t))

    def load_strategy_from_step(self, t):
        return self._load_strategy(self.file_name(t))

    @property
    def file_name_best_exploitability(self):
        return os.path.join(self._save_path, 'best_strategy.pkl')

    def save_best_strategy(self, strategy, t, exploitability):
        if exploitability < self._best_exploitability:
            self._save_strategy(strategy, self.file_name_best_exploitability)
            self._best_exploitability = exploitability
            self._best_exploitability_t = t

    def load_best_strategy(self):
        return self._load_strategy(self.file_name_best_exploitability)
