"""Deal sharding: structure checks and a world_size-2 gloo run of the per-iteration exchange on CPU.

The CPU run uses the oracle's level sweep for each rank's local subtree sweep (the CUDA kernels need a GPU);
what is under test is the product's host logic: rlpoker_b200.sharding.shard_flat_tree / allreduce_delta."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, make_game


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize('name,world', [('leduc3', 2), ('leduc3', 4), ('leduc3', 8), ('ocp4', 2)])
def test_shards_partition_the_tree(name, world):
    from rlpoker_b200.flatten import flatten
    from rlpoker_b200.sharding import shard_flat_tree, shard_ranges
    ft = flatten(make_game(name))
    cut, ranges = shard_ranges(ft, world)
    assert ranges[0][0] == ft.level_off[cut] and ranges[-1][1] == ft.level_off[cut + 1]
    assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    seen = np.zeros(ft.num_nodes, int)
    for r in range(world):
        sh = shard_flat_tree(ft, r, world)
        seen[sh.global_nodes] += 1
        g = sh.global_nodes
        assert np.array_equal(sh.player, ft.player[g]) and np.array_equal(sh.infoset, ft.infoset[g])
        assert np.array_equal(sh.cprob, ft.cprob[g]) and np.array_equal(sh.u1, ft.u1[g])
        assert np.array_equal(g[sh.parent[1:]], ft.parent[g[1:]])           # same edges
        assert sh.num_infosets == ft.num_infosets and sh.num_levels == ft.num_levels
    below = np.arange(ft.num_nodes) >= ft.level_off[cut]
    assert np.all(seen[below] == 1)                                           # every deep node on exactly one rank
    assert np.all(seen[~below] >= 1)
    sizes = [shard_flat_tree(ft, r, world).num_nodes for r in range(world)]
    assert max(sizes) <= 1.35 * min(sizes)


def test_root_player_game_does_not_shard():
    from rlpoker_b200.flatten import flatten
    from rlpoker_b200.sharding import shard_ranges
    with pytest.raises(ValueError):
        shard_ranges(flatten(make_game('rps')), 2)


def _rank_main(rank, world, port, name, iters, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch
    import torch.distributed as dist
    from conftest import make_game
    from oracle import flat, oracle as orc
    from rlpoker_b200.flatten import flatten
    from rlpoker_b200.sharding import allreduce_delta, shard_flat_tree
    dist.init_process_group('gloo', rank=rank, world_size=world)
    ft = flatten(make_game(name))
    sh = shard_flat_tree(ft, rank, world)
    local = orc.Oracle(flat.from_arrays(sh.num_nodes, sh.parent, sh.first_child, sh.player, sh.infoset, sh.label,
                                        sh.hidden, sh.cprob, sh.u1, sh.u2, nact=ft.nact))
    Q = ft.num_pairs
    assert local.Q == Q
    R, S = np.zeros(Q), np.zeros(Q)
    sigma = np.repeat(1.0 / ft.nact, ft.nact)
    for _ in range(iters):
        q = local.Q
        local.R[:] = 0.0
        local.S[:] = 0.0
        local.sigma[:] = sigma[:q]
        local.vanilla_levels(1)
        delta = torch.zeros(2 * Q, dtype=torch.float64)
        delta[:q] = torch.from_numpy(local.R.copy())
        delta[Q:Q + q] = torch.from_numpy(local.S.copy())
        allreduce_delta(delta)
        R += delta[:Q].numpy()
        S += delta[Q:].numpy()
        for i in range(ft.num_infosets):
            lo, hi = int(ft.act_off[i]), int(ft.act_off[i + 1])
            sigma[lo:hi] = orc.regret_matching(R[lo:hi])
    if rank == 0:
        np.savez(out, R=R, S=S, sigma=sigma)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_run_matches_single_process():
    import torch.multiprocessing as mp
    from oracle import oracle as orc
    from conftest import oracle_tree
    name, iters, world = 'leduc2', 6, 2
    out = os.path.join('/tmp', 'rlp_gloo_%d.npz' % os.getpid())
    mp.start_processes(_rank_main, args=(world, _free_port(), name, iters, out), nprocs=world, join=True,
                       start_method='spawn')
    got = np.load(out)
    o = orc.Oracle(oracle_tree(name))
    o.vanilla_levels(iters)
    np.testing.assert_allclose(got['R'], o.R, rtol=0, atol=1e-12)
    np.testing.assert_allclose(got['S'], o.S, rtol=0, atol=1e-12)
    np.testing.assert_allclose(got['sigma'], o.sigma, rtol=0, atol=1e-10)
    os.remove(out)


@pytest.mark.parametrize('world', [2, 4, 8])
def test_every_shard_plans_and_compiles(world):
    """The device layout of every rank's subtree (tiling, shapes, NVRTC compilation) derived on the host."""
    from rlpoker_b200.games.leduc_native import leduc_flat_tree
    from rlpoker_b200.sharding import shard_flat_tree
    from rlpoker_b200.solver import plan_tree
    ft = leduc_flat_tree(5)
    deals = 0
    for r in range(world):
        p = plan_tree(shard_flat_tree(ft, r, world))
        assert p['tiled'] == 1 and p['micro_shapes'] == 1 and p['kernels_failed'] == 0 and p['kernels_compiled'] == 2
        assert p['micro_subtrees'] == 9 * 8 * p['upper_tiles']
        deals += p['upper_tiles']
    assert deals == 10 * 9


def test_lane_ranges_partition_the_lanes():
    from rlpoker_b200.sharding import ShardedSampledCFR
    for lanes in (1, 7, 64, 1 << 20):
        for world in (1, 2, 3, 8):
            r = [ShardedSampledCFR.lane_range(lanes, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == lanes
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
