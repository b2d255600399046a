"""Worker of tests/test_gpu_multirank.py: sampled lanes sharded over real NCCL ranks vs the oracle's lane loop fed the
same Philox stream (tolerance 1e-9: the lanes' contributions are added by floating-point atomics in no fixed order)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from conftest import make_game, oracle_tree
    from oracle import oracle as orc
    from rlpoker_b200.flatten import flatten
    from rlpoker_b200.sharding import ShardedSampledCFR
    ft = flatten(make_game('leduc3'))
    # 9000 lanes: the level-by-level lane kernels; two calls: the second one replays the CUDA graph the first captured
    for variant, lanes, iters in ((0, 37, 6), (1, 101, 8), (1, 9000, 7)):
        sh = ShardedSampledCFR(ft, rank, world)
        sh.iterate(variant, iters - 3, lanes, seed=11)
        sh.iterate(variant, 3, lanes, seed=11)
        assert world == 1 or sh._graph is not None
        torch.cuda.synchronize()
        o = orc.Oracle(oracle_tree('leduc3'))
        if variant == 0:
            o.cfr_recursive(iters, sampling=True, lanes=lanes, rng=orc.rng_philox(11))
        else:
            o.external(iters, orc.rng_philox(11), lanes=lanes)
        s = sh.solver
        np.testing.assert_allclose(s.regrets, o.R, rtol=0, atol=1e-9)
        np.testing.assert_allclose(s.strategy_sum, o.S, rtol=0, atol=1e-9)
        np.testing.assert_allclose(s.sigma, o.sigma, rtol=0, atol=1e-9)
        flags = s.flags()
        assert np.array_equal((flags & 1) != 0, o.in_R != 0) and np.array_equal((flags & 4) != 0, o.in_sigma != 0)
        assert sh.nodes_touched() == o.nodes_touched, (sh.nodes_touched(), o.nodes_touched)
        sh.close()
    dist.barrier()
    if rank == 0:
        print('sampled sharded ok: world=%d' % world)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
