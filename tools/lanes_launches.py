"""Measurement aid: one warm iteration of 2^20 external lanes on Leduc-3, for an ncu launch list."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from rlpoker_b200.games.leduc_native import leduc_flat_tree  # noqa: E402
from rlpoker_b200.solver import DeviceCFR, DeviceTree  # noqa: E402

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 1
lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
s = DeviceCFR(DeviceTree(leduc_flat_tree(3)))
s.sampled(variant, 2, lanes=lanes, seed=1)
torch.cuda.synchronize()
print('touched', s.nodes_touched)
