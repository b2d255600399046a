"""Scratch: in-graph schedule of the sweep kernels (Leduc-13)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlpoker_b200.games.leduc_native import create_leduc
from rlpoker_b200.solver import DeviceCFR, device_tree
n = int(sys.argv[1]) if len(sys.argv) > 1 else 13
s = DeviceCFR(device_tree(create_leduc(n)))
s.vanilla(64)
tl = s.timeline(16)
names = {0: 'phase0', 1: 'micro', 2: 'phase1', 3: 'accA', 4: 'accB'}
for k, a, b in tl[16:48]:
    print('%-7s start %8.2f  end %8.2f  dur %6.2f' % (names.get(k, k), a, b, b - a))
