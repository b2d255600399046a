"""Measurement aid: BASELINE config 3 -- 2^20 traversal lanes per iteration on Leduc-3, both sampled variants."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from rlpoker_b200.games.leduc_native import leduc_flat_tree  # noqa: E402
from rlpoker_b200.solver import DeviceCFR, DeviceTree  # noqa: E402

lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
dt = DeviceTree(leduc_flat_tree(3))
for variant, name in ((1, 'external'), (0, 'chance')):
    s = DeviceCFR(dt)
    s.sampled(variant, 2, lanes=lanes, seed=1)
    torch.cuda.synchronize()
    n0 = s.nodes_touched
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s.sampled(variant, 5, lanes=lanes, seed=1)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(json.dumps({'variant': name, 'lanes': lanes, 'ms_per_iteration': ms,
                      'node_updates_per_s': (s.nodes_touched - n0) / 5 / (ms * 1e-3), 'exploitability': s.exploitability()[0]}), flush=True)
    del s
