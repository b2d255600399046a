"""Worker of tests/test_gpu_multirank.py: sharding by public state (peer stores inside the persistent kernel), one
process per GPU.  Every rank's merged state must equal the CPU oracle BIT FOR BIT (the exchange moves subtree values,
no floating-point sum changes its order), and the exploitability to 1e-12."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from oracle import flat, oracle as orc
    from rlpoker_b200._lib import RLP_REGRET, RLP_SIGMA, RLP_STRAT_SUM
    from rlpoker_b200.games.leduc_native import leduc_flat_tree
    from rlpoker_b200.sharding import PublicStateShardedCFR
    ft = leduc_flat_tree(n)
    sh = PublicStateShardedCFR(ft, rank, world)
    assert sh.owned_isets.any()
    total = torch.tensor([int(sh.owned_isets.sum())], device='cuda')
    dist.all_reduce(total)
    assert int(total.item()) == ft.num_infosets, 'every information set must have exactly one owner'
    ot = flat.from_arrays(ft.num_nodes, ft.parent, ft.first_child, ft.player, ft.infoset, ft.label, ft.hidden,
                          ft.cprob, ft.u1, ft.u2)
    o = orc.Oracle(ot)
    done = 0
    for T, lin in ((1, False), (4, False), (23, False), (30, True)):
        sh.iterate(T - done, linear_weight=lin)
        sh.check()
        o.vanilla_levels(T - done, linear_weight=lin, threads=4)
        done = T
        assert np.array_equal(sh.merged(RLP_REGRET), o.R), ('regrets', T)
        assert np.array_equal(sh.merged(RLP_STRAT_SUM), o.S), ('action counts', T)
        assert np.array_equal(sh.merged(RLP_SIGMA), o.sigma), ('strategy', T)
    sh.gather_state()
    avg = o.average_strategy()[0]
    assert np.array_equal(sh.solver.average, avg)
    assert abs(sh.solver.exploitability()[0] - o.exploitability(avg)) <= 1e-12
    sh.iterate(5)                                   # the solve continues after a gather
    sh.check()
    o.vanilla_levels(5, threads=4)
    assert np.array_equal(sh.merged(RLP_REGRET), o.R)
    dist.barrier()
    if rank == 0:
        print('p2p sharded ok: leduc%d world=%d' % (n, world))
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
