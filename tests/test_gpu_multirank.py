"""Real multi-process runs: one process per GPU, launched with torch.distributed.run on 127.0.0.1.  Skipped on
boxes with fewer than two GPUs (the single-GPU emulation of the same logic is tests/test_gpu_sharded.py)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _gpus():
    import torch
    return torch.cuda.device_count()


def _torchrun(world, script, args, env=None, timeout=420):
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
           '--master-addr', '127.0.0.1', '--master-port', str(29500 + os.getpid() % 2000),
           os.path.join(ROOT, 'tests', 'workers', script)] + [str(a) for a in args]
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run(cmd, env=e, text=True, capture_output=True, timeout=timeout)
    assert out.returncode == 0, out.stdout[-3000:] + '\n' + out.stderr[-3000:]
    return out.stdout


@pytest.mark.parametrize('graph', ['0', '1'])
def test_two_rank_nccl_deal_sharded_vs_oracle(graph):
    if _gpus() < 2:
        pytest.skip('needs 2 GPUs')
    out = _torchrun(2, 'nccl_sharded_worker.py', [5, 23], env={'RLP_SHARD_CGRAPH': graph})
    assert 'nccl sharded ok' in out


@pytest.mark.parametrize('n', [3, 6])
def test_two_rank_public_state_sharding_bit_exact(n):
    if _gpus() < 2:
        pytest.skip('needs 2 GPUs')
    out = _torchrun(2, 'p2p_sharded_worker.py', [n])
    assert 'p2p sharded ok' in out


def test_public_state_sharding_world_one():
    """The same code path with a single rank (no peers): the group ranges, the owned mask and the merge."""
    out = _torchrun(1, 'p2p_sharded_worker.py', [3])
    assert 'p2p sharded ok' in out


def test_sampled_lanes_world_one_and_two():
    """Lane sharding of both sampled variants: one rank always, two ranks when the box has them."""
    out = _torchrun(1, 'sampled_sharded_worker.py', [])
    assert 'sampled sharded ok: world=1' in out
    if _gpus() >= 2:
        out = _torchrun(2, 'sampled_sharded_worker.py', [])
        assert 'sampled sharded ok: world=2' in out
