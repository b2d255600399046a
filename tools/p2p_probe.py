"""Measurement aid (torchrun, one process per GPU): phase timeline of the public-state-sharded solve on Leduc-13."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', rank)))
dist.init_process_group('nccl', device_id=torch.device('cuda', torch.cuda.current_device()))
from rlpoker_b200.games.leduc_native import leduc_flat_tree  # noqa: E402
from rlpoker_b200.sharding import PublicStateShardedCFR  # noqa: E402
sh = PublicStateShardedCFR(leduc_flat_tree(int(sys.argv[1]) if len(sys.argv) > 1 else 13), rank, world)
sh.iterate(20)
torch.cuda.synchronize()
dist.barrier()
for rep in range(3):
    tl = sh.solver.fused_timeline(100)
    dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); dist.barrier()
a.record()
for _ in range(30):
    sh.iterate(10)
b.record()
torch.cuda.synchronize()
print(json.dumps({'rank': rank, 'debug': os.environ.get('RLP_P2P_DEBUG', '0'), 'timeline': tl, 'us_per_iter_events': a.elapsed_time(b) * 1e3 / 300}), flush=True)
dist.barrier()
dist.destroy_process_group()
