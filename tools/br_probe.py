"""Measurement aid: time of one exploitability evaluation (both best responses) on Leduc-n, fused groups vs level sweep,
with the oracle's best response timed on the host beside it."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import oracle as orc  # noqa: E402
from rlpoker_b200.games.leduc_native import leduc_flat_tree  # noqa: E402
from rlpoker_b200.solver import DeviceCFR, DeviceTree  # noqa: E402


def gpu_time(s, reps=20):
    s.exploitability()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        v = s.exploitability()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / reps, v


for n in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else '3,5,13').split(',')]:
    ft = leduc_flat_tree(n)
    dt = DeviceTree(ft)
    s = DeviceCFR(dt)
    s.vanilla(10)
    us, v = gpu_time(s)
    os.environ['RLP_NO_FUSED_BR'] = '1'
    us_lvl, v_lvl = gpu_time(s)
    del os.environ['RLP_NO_FUSED_BR']
    o = orc.Oracle(orc.leduc_tree(n))
    avg = s.average
    t0 = time.perf_counter()
    e_cpu = o.exploitability(avg)
    cpu_s = time.perf_counter() - t0
    print(json.dumps({'n': n, 'nodes': ft.num_nodes, 'fused_br': dt.layout['fused_br'], 'gpu_us': us, 'gpu_level_sweep_us': us_lvl,
                      'cpu_oracle_s': cpu_s, 'expl_gpu': v[0], 'expl_level': v_lvl[0], 'expl_cpu': e_cpu,
                      'abs_diff_cpu': abs(v[0] - e_cpu), 'abs_diff_level': abs(v[0] - v_lvl[0])}), flush=True)
    del s
    dt.close()
