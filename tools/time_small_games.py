"""Scratch timing of the first CUDA path (not a bench): vanilla iterations on Leduc-3/5."""
import time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from rlpoker_b200.games.leduc import Leduc
from rlpoker_b200.flatten import flatten
from rlpoker_b200.solver import DeviceCFR, device_tree
from rlpoker_b200.games.leduc_native import create_leduc
for n in (3, 5, 13):
    t0 = time.time(); g = Leduc.create_game(n) if n < 6 else create_leduc(n); ft = flatten(g); print('build+flatten leduc%d %.2fs N=%d' % (n, time.time() - t0, ft.num_nodes))
    s = DeviceCFR(device_tree(g)); print(s.tree.layout)
    s.vanilla(64); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 2048
    a.record(); s.vanilla(iters); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print('leduc%d: %.2f us/iter, %.3e node updates/s' % (n, 1e3 * ms / iters, 2 * ft.num_nodes * iters / (ms * 1e-3)))
    print('stages us', s.profile_stages(50))
    print('exploitability', s.exploitability())
