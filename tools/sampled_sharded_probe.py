"""Measurement aid (torchrun, one process per GPU): BASELINE config 3 with the 2^20 lanes split over the ranks."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', rank)))
dist.init_process_group('nccl', device_id=torch.device('cuda', torch.cuda.current_device()))
import bench  # noqa: E402
out = bench.sampled_config3_sharded(rank, world)
if rank == 0:
    out['n_gpus'] = world
    print(json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
