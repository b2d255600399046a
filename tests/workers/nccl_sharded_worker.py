"""Worker of tests/test_gpu_multirank.py: one process per GPU (torchrun), real NCCL ranks.

Runs the deal-sharded solve (rlp_cfr_sharded_iters: local sweep -> ncclAllReduce -> apply, in C) on a native Leduc-n
tree and compares EVERY rank's replicated state with the CPU oracle: regrets / action counts within 1e-9 (the
all-reduce adds the ranks' partial sums in another association than the reference), average strategy within 1e-9,
exploitability within 1e-7 (north_star tolerances)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 23
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from oracle import flat, oracle as orc
    from rlpoker_b200.games.leduc_native import leduc_flat_tree
    from rlpoker_b200.sharding import ShardedCFR
    ft = leduc_flat_tree(n)
    sh = ShardedCFR(ft, rank, world)
    assert sh._native, 'library-owned NCCL communicator was not created'
    sh.iterate(iters)
    sh.iterate(7, linear_weight=True)
    torch.cuda.synchronize()
    ot = flat.from_arrays(ft.num_nodes, ft.parent, ft.first_child, ft.player, ft.infoset, ft.label, ft.hidden,
                          ft.cprob, ft.u1, ft.u2)
    o = orc.Oracle(ot)
    o.vanilla_levels(iters, threads=4)
    o.vanilla_levels(7, linear_weight=True, threads=4)
    s = sh.solver
    np.testing.assert_allclose(s.regrets, o.R, rtol=0, atol=1e-9)
    np.testing.assert_allclose(s.strategy_sum, o.S, rtol=0, atol=1e-9)
    avg = o.average_strategy()[0]
    np.testing.assert_allclose(s.average, avg, rtol=0, atol=1e-9)
    # best response needs the WHOLE tree (a rank's solver only holds its deals): evaluate the replicated average there
    from rlpoker_b200.solver import DeviceTree
    full = DeviceTree(ft)
    assert abs(full.exploitability(s.average) - o.exploitability(avg)) <= 1e-7
    full.close()
    # replicas must be bitwise identical across ranks
    mine = torch.from_numpy(s.regrets).cuda()
    ref = mine.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(mine, ref), 'replicated regrets diverged between ranks'
    dist.barrier()
    if rank == 0:
        print('nccl sharded ok: leduc%d world=%d iters=%d graph=%s' % (n, world, iters + 7,
                                                                      os.environ.get('RLP_SHARD_CGRAPH', '0')))
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
