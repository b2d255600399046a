"""Measurements for the BASELINE.json configs other than the headline (one GPU):
   config 2  Leduc-3 vanilla CFR, 100k iterations: time + parity vs the CPU oracle
   config 3  Leduc-3 external / chance sampling with 2^20 lanes per iteration
   config 5  exploitability (best-response sweep) on Leduc-N
Writes one JSON object to stdout / gpurun_out/configs.json."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from rlpoker_b200._lib import RLP_AVERAGE, RLP_CHANCE_SAMPLED, RLP_EXTERNAL_SAMPLED
from rlpoker_b200.games.leduc_native import create_leduc, leduc_flat_tree
from rlpoker_b200.solver import DeviceCFR, DeviceTree, device_tree

out = {}
what = set(sys.argv[1:]) or {'c2', 'c3', 'c5', 'l25'}


def timed(fn):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3


if 'c2' in what:
    from oracle import flat, oracle as orc
    ft = leduc_flat_tree(3)
    s = DeviceCFR(DeviceTree(ft))
    s.vanilla(100)
    s.reset()
    T = 100000
    sec = timed(lambda: s.vanilla(T))
    o = orc.Oracle(flat.from_arrays(ft.num_nodes, ft.parent, ft.first_child, ft.player, ft.infoset, ft.label, ft.hidden,
                                    ft.cprob, ft.u1, ft.u2))
    t0 = time.time(); o.vanilla_levels(T, threads=min(8, os.cpu_count() or 1)); cpu_sec = time.time() - t0
    avg_gpu, avg_cpu = s.average, o.average_strategy()[0]
    out['config2_leduc3_100k'] = {
        'iterations': T, 'gpu_seconds': sec, 'gpu_updates_per_s': 2.0 * ft.num_nodes * T / sec,
        'cpu_oracle_seconds': cpu_sec, 'max_abs_diff_average_strategy': float(np.abs(avg_gpu - avg_cpu).max()),
        'regrets_bit_identical': bool(np.array_equal(s.regrets, o.R)),
        'exploitability_gpu': s.exploitability()[0], 'exploitability_oracle': o.exploitability(avg_cpu)}

if 'c3' in what:
    ft = leduc_flat_tree(3)
    for name, variant in (('external', RLP_EXTERNAL_SAMPLED), ('chance', RLP_CHANCE_SAMPLED)):
        s = DeviceCFR(DeviceTree(ft))
        lanes = 1 << 20
        s.sampled(variant, 2, lanes=lanes, seed=1)
        n0 = s.nodes_touched
        iters = 5
        sec = timed(lambda: s.sampled(variant, iters, lanes=lanes, seed=1))
        out['config3_leduc3_%s_1M_lanes' % name] = {
            'lanes': lanes, 'iterations': iters, 'seconds_per_iteration': sec / iters,
            'node_updates_per_s': (s.nodes_touched - n0) / sec, 'exploitability_after': s.exploitability()[0]}

if 'c5' in what:
    res = {}
    for n in (3, 5, 9, 13):
        ft = leduc_flat_tree(n)
        s = DeviceCFR(DeviceTree(ft))
        s.vanilla(20)
        s.exploitability()
        sec = timed(lambda: [s.exploitability() for _ in range(5)]) / 5
        res[str(n)] = {'nodes': ft.num_nodes, 'seconds_per_exploitability': sec, 'nodes_per_s': 2.0 * ft.num_nodes / sec}
    out['config5_exploitability'] = res

if 'l25' in what:
    t0 = time.time(); ft = leduc_flat_tree(25); gen = time.time() - t0
    t0 = time.time(); tree = DeviceTree(ft); up = time.time() - t0
    s = DeviceCFR(tree)
    s.vanilla(20)
    sec = timed(lambda: s.vanilla(100)) / 100
    out['leduc25_vanilla'] = {'nodes': ft.num_nodes, 'generate_s': gen, 'upload_tile_jit_s': up, 'layout': tree.layout,
                              'us_per_iteration': sec * 1e6, 'node_updates_per_s': 2.0 * ft.num_nodes / sec,
                              'device_bytes_tree': tree.device_bytes}

print(json.dumps(out, indent=1))
os.makedirs('gpurun_out', exist_ok=True)
json.dump(out, open('gpurun_out/configs.json', 'w'), indent=1)
