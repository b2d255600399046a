"""Deal-sharded solve on the GPU.  Two ranks are emulated on ONE device (two solver objects, deltas summed
with a torch add in place of the NCCL all-reduce); a real multi-process NCCL run is exercised by
``torchrun ... bench.py --gpus N``."""
import numpy as np
import pytest

from conftest import make_game, oracle_tree

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name,world', [('leduc3', 2), ('leduc3', 3), ('ocp4', 2)])
def test_emulated_ranks_match_oracle(name, world):
    import torch
    from oracle import oracle as orc
    from rlpoker_b200.flatten import flatten
    from rlpoker_b200.sharding import ShardedCFR
    ft = flatten(make_game(name))
    ranks = [ShardedCFR(ft, r, world) for r in range(world)]
    o = orc.Oracle(oracle_tree(name))
    iters = 12
    for _ in range(iters):
        for r in ranks:
            r.solver.local_sweep()
        total = torch.stack([r.delta for r in ranks]).sum(0)
        for r in ranks:
            r.delta.copy_(total)
            r.solver.apply_delta()
    o.vanilla_levels(iters)
    for r in ranks:
        np.testing.assert_allclose(r.solver.regrets, o.R, rtol=0, atol=1e-11)
        np.testing.assert_allclose(r.solver.strategy_sum, o.S, rtol=0, atol=1e-11)
        np.testing.assert_allclose(r.solver.average, o.average_strategy()[0], rtol=0, atol=1e-9)
        assert np.array_equal(r.solver.regrets, ranks[0].solver.regrets)      # replicas stay identical
    assert ranks[0].solver.nodes_touched > 0


def test_world_one_split_path_equals_fused_path():
    from rlpoker_b200.flatten import flatten
    from rlpoker_b200.sharding import ShardedCFR
    from rlpoker_b200.solver import DeviceCFR, device_tree
    ft = flatten(make_game('leduc3'))
    a = ShardedCFR(ft, 0, 1)
    a.iterate(9)
    b = DeviceCFR(device_tree(make_game('leduc3')))
    b.vanilla(9)
    np.testing.assert_allclose(a.solver.regrets, b.regrets, rtol=0, atol=1e-11)
    assert a.nodes_touched == b.nodes_touched


def test_world_one_graph_replay_path():
    """iterate() >= 10 iterations replays a CUDA graph of (local sweep, all-reduce, apply) x 10."""
    from rlpoker_b200.flatten import flatten
    from rlpoker_b200.sharding import ShardedCFR
    from rlpoker_b200.solver import DeviceCFR, device_tree
    ft = flatten(make_game('leduc3'))
    a = ShardedCFR(ft, 0, 1)
    a.iterate(25, use_graph=True)
    a.iterate(10, linear_weight=True, use_graph=True)
    b = DeviceCFR(device_tree(make_game('leduc3')))
    b.vanilla(25)
    b.vanilla(10, linear_weight=True)
    np.testing.assert_allclose(a.solver.regrets, b.regrets, rtol=0, atol=1e-9)
    np.testing.assert_allclose(a.solver.strategy_sum, b.strategy_sum, rtol=0, atol=1e-9)
    assert a.solver.t == 35 and a._cuda_graph is not None


def test_world_one_native_nccl_loop():
    """rlp_cfr_sharded_iters: the sweep / ncclAllReduce / apply loop in C (a 1-rank communicator here)."""
    from rlpoker_b200.flatten import flatten
    from rlpoker_b200.sharding import ShardedCFR
    from rlpoker_b200.solver import DeviceCFR, device_tree
    ft = flatten(make_game('leduc3'))
    a = ShardedCFR(ft, 0, 1)
    assert a._native
    a.iterate(12)
    a.iterate(5, linear_weight=True)
    b = DeviceCFR(device_tree(make_game('leduc3')))
    b.vanilla(12)
    b.vanilla(5, linear_weight=True)
    np.testing.assert_allclose(a.solver.regrets, b.regrets, rtol=0, atol=1e-9)
    np.testing.assert_allclose(a.solver.strategy_sum, b.strategy_sum, rtol=0, atol=1e-9)


def test_world_one_native_loop_as_cuda_graph():
    """RLP_SHARD_CGRAPH=1: the C loop (sweep, ncclAllReduce, apply) captured as a CUDA graph of 10 iterations."""
    import os, subprocess, sys
    from conftest import ROOT
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
from conftest import make_game
from rlpoker_b200.flatten import flatten
from rlpoker_b200.sharding import ShardedCFR
from rlpoker_b200.solver import DeviceCFR, device_tree
ft = flatten(make_game('leduc3'))
a = ShardedCFR(ft, 0, 1)
assert a._native
a.iterate(34); a.iterate(10); a.iterate(12, linear_weight=True)
b = DeviceCFR(device_tree(make_game('leduc3')))
b.vanilla(44); b.vanilla(12, linear_weight=True)
np.testing.assert_allclose(a.solver.regrets, b.regrets, rtol=1e-11, atol=1e-9)
np.testing.assert_allclose(a.solver.strategy_sum, b.strategy_sum, rtol=1e-11, atol=1e-9)
print('cgraph ok')
''' % (ROOT, os.path.join(ROOT, 'tests'))
    out = subprocess.check_output([sys.executable, '-c', code], env=dict(os.environ, RLP_SHARD_CGRAPH='1'), text=True)
    assert 'cgraph ok' in out
