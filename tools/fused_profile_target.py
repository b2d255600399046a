"""ncu target: a few launches of the group-fused persistent kernel on Leduc-n (default 13), 10 iterations each."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from rlpoker_b200.games.leduc_native import leduc_flat_tree  # noqa: E402
from rlpoker_b200.solver import DeviceCFR, DeviceTree  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 13
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 6
s = DeviceCFR(DeviceTree(leduc_flat_tree(n)))
for _ in range(launches):
    s.vanilla(10)
torch.cuda.synchronize()
print('done', s.t)
