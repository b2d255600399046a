"""Experiment driver for the group-fused persistent kernel: one subprocess per configuration (the CTA shape is fixed at
tree-upload time through environment variables), Leduc-n, prints one JSON line per configuration with the phase
timeline (CTA 0's view) and the CUDA-event time per iteration, plus a hash of the regrets after a fixed number of
iterations (all configurations must agree bit for bit).

    python tools/fused_sweep.py [--ranks 13] [--configs 'RLP_FUSED_BLOCK=192,RLP_FUSED_MINB=3;RLP_FUSED_BLOCK=256']
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import hashlib, json, os, sys, time
sys.path.insert(0, %r)
import numpy as np, torch
from rlpoker_b200.games.leduc_native import leduc_flat_tree
from rlpoker_b200.solver import DeviceCFR, DeviceTree
n = int(sys.argv[1])
ft = leduc_flat_tree(n)
dt = DeviceTree(ft)
s = DeviceCFR(dt)
s.vanilla(20)
torch.cuda.synchronize()
digest = hashlib.sha256(s.regrets.tobytes()).hexdigest()[:16]
out = {'layout': {k: dt.layout[k] for k in ('fused', 'fused_batches', 'fused_upper_groups', 'fused_grid')}, 'sha_R20': digest}
if dt.layout['fused'] == 1 and os.environ.get('RLP_NO_FUSED') != '1':
    out['timeline_us'] = s.fused_timeline(60)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for rep in range(5):
    a.record()
    for _ in range(20):
        s.vanilla(10)
    b.record()
    torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b) * 1e3 / 200)
out['us_per_iteration'] = best
out['updates_per_s'] = 2.0 * ft.num_nodes / (best * 1e-6)
print(json.dumps(out))
''' % ROOT


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--ranks', type=int, default=13)
    ap.add_argument('--configs', default='')
    args = ap.parse_args()
    configs = [c for c in args.configs.split(';')] if args.configs else ['']
    for cfg in configs:
        env = dict(os.environ)
        for kv in filter(None, cfg.split(',')):
            k, v = kv.split('=')
            env[k] = v
        try:
            out = subprocess.run([sys.executable, '-c', WORKER, str(args.ranks)], env=env, text=True, capture_output=True,
                                 timeout=300)
            line = out.stdout.strip().splitlines()[-1] if out.returncode == 0 and out.stdout.strip() else \
                json.dumps({'error': out.stderr[-800:]})
        except subprocess.TimeoutExpired:
            line = json.dumps({'error': 'timeout'})
        print(json.dumps({'config': cfg or 'default', 'result': json.loads(line)}), flush=True)


if __name__ == '__main__':
    main()
