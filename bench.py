#!/usr/bin/env python
"""Headline benchmark: vanilla CFR tree-node updates/s on Leduc-13 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One STEP = ``--iters-per-step`` (default 10, the reference's evaluation cadence, cfr.py:125) full-tree CFR
iterations for both players on ``Leduc(get_deck(13, 2))`` (3 244 177 nodes, 47 008 information sets).  A node
update is one increment of the reference's ``CFRState.nodes_touched`` (cfr_util.py:10-11): 2 N per iteration.

* ``value``  : updates/s with all state resident in HBM (CUDA events around exactly K steps, max over ranks).
* ``e2e``    : the same through the public solver object with HOST buffers: every step uploads the regret and
               action-count tables from pinned host memory (rlp_cfr_write), runs the iterations, and reads
               the average strategy back (rlp_cfr_read).
* ``roofline``: algorithmic bytes per iteration (SURVEY.md section 8d formula) / measured iteration time
               against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
* ``cpu_baseline``: this repo's C port of the reference algorithm (oracle/, bit-identical to the reference's
               Python) timed on the host cores of the same box.
N > 1 (torchrun): the depth-2 deal subtrees are sharded over ranks (strong scaling of the same tree) with
one NCCL all-reduce of the [2Q] deltas per iteration.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'cfr_tree_node_updates_per_s'
UNIT = 'node-updates/s'
RANKS = 13


def algorithmic_bytes_per_iteration(ft):
    """fp64 state, i32 topology, u8 slot/player; each array element charged once per
    pass that must read or write it."""
    import numpy as np
    N, Nt, Nd, Q = ft.num_nodes, ft.num_terminals, ft.num_decisions, ft.num_pairs
    nchild = np.diff(ft.first_child.astype(np.int64))
    Ec = int(nchild[ft.player == 0].sum())
    F = 5 * (N - 1) + 8 * Ec + 8 * Q + 4 * Nd + 24 * (N - Nt) + 24 * N
    B = 8 * Nt + 16 * N + 16 * (N - 1) + 5 * (N - 1) + 8 * Ec + 8 * Q + 4 * Nd
    U = 24 * Nd + 40 * Q
    return F + B + U


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


def traffic_per_iteration():
    """dram__bytes_read+write per iteration summed over the sweep kernels, from the committed ncu capture."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
            return float(json.load(f)['dram_bytes_per_iteration'])
    except Exception:
        return None


# One Leduc micro subtree = the 23-node betting round after the board card: 8 decision nodes, 15 leaves, 22 edges.
# Bytes one `micro_k` thread moves for it (the generated source, csrc/jit.cu generate()): per decision node its
# information-set column and its contribution slot (2 x i32), one f64 payoff per leaf, static chance reach (f64),
# portal and reach indices (2 x i32), the two players' reaches in (2 x f64); out: the subtree value (f64), 22 regret
# terms and 8 own-reach terms (f64).  The 22 strategy probabilities come from the 1 MB strategy table shared by all
# subtrees; it is charged once per launch.
MICRO_BYTES_IN = 8 * 2 * 4 + 15 * 8 + 8 + 2 * 4 + 2 * 8
MICRO_BYTES_OUT = 8 + (22 + 8) * 8


def dominant_kernel_roofline(layout, stages, peak, strategy_bytes):
    """Roofline of the single longest kernel of the iteration (`micro_k`): algorithmic bytes of one launch over
    its live CUDA-event duration.  None when the tree is not tiled or the stage times are missing."""
    try:
        n = int(layout['micro_subtrees'])
        us = float(stages['micro'])
        if n <= 0 or us <= 0:
            return None
        nbytes = n * (MICRO_BYTES_IN + MICRO_BYTES_OUT) + strategy_bytes
        achieved = nbytes / (us * 1e-6) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
                k = json.load(f)['kernels']['micro_k']
            traffic = (float(k['dram_read_MB']) + float(k['dram_write_MB'])) * 1e6
        except Exception:
            pass
        return {'kernel': 'micro_k', 'units': n, 'bytes_per_unit': MICRO_BYTES_IN + MICRO_BYTES_OUT,
                'algorithmic_bytes_per_launch': nbytes, 'us_per_launch': us, 'achieved': achieved, 'peak': peak,
                'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
                'note': 'stores of one launch stay in the 126 MB L2 for the accumulate kernel, so DRAM traffic is '
                        'below the algorithmic bytes'}
    except Exception:
        return None


class ClockSampler:
    QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.proc = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.QUERY,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ''
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in out.strip().splitlines():
            parts = [x.strip() for x in line.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


def build_tree():
    from rlpoker_b200.games.leduc_native import leduc_flat_tree
    return leduc_flat_tree(RANKS)


def oracle_for(ft):
    """The CPU port (oracle/) on the same arrays -- used only as the reported CPU baseline."""
    from oracle import flat, oracle as orc
    t = flat.from_arrays(ft.num_nodes, ft.parent, ft.first_child, ft.player, ft.infoset, ft.label, ft.hidden,
                         ft.cprob, ft.u1, ft.u2)
    return orc.Oracle(t)


def time_cpu(ft, budget_s, threads):
    o = oracle_for(ft)
    t0 = time.perf_counter()
    o.vanilla_levels(1, threads=threads)
    one = time.perf_counter() - t0
    n = max(1, min(5000, int(budget_s / max(one, 1e-3))))
    t0 = time.perf_counter()
    o.vanilla_levels(n, threads=threads)
    dt = time.perf_counter() - t0
    return 2.0 * ft.num_nodes * n / dt, n, dt


def run_reference(args, rank, world):
    """--impl reference: the reference algorithm's CPU port on all host cores, same tree / metric."""
    if rank != 0:
        return
    ft = build_tree()
    threads = os.cpu_count() or 1
    o = oracle_for(ft)
    for _ in range(args.warmup):
        o.vanilla_levels(1, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.vanilla_levels(1, threads=threads)
    dt = time.perf_counter() - t0
    value = 2.0 * ft.num_nodes * args.steps / dt
    sample = '1 CFR iteration per step (of the %d the GPU arm runs per step), %d steps' % (args.iters_per_step, args.steps)
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'leduc13-vanilla-cfr', 'nodes': ft.num_nodes, 'infosets': ft.num_infosets,
                   'iters_per_step': 1},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }))


def arm_watchdog(seconds):
    """A hung collective must not hang the box: after `seconds` of wall clock the process reports and exits non-zero
    (torchrun then tears the other ranks down)."""
    import threading

    def fire():
        sys.stderr.write('bench.py: watchdog fired after %.0f s, aborting\n' % seconds)
        sys.stderr.flush()
        os._exit(3)

    t = threading.Timer(seconds, fire)
    t.daemon = True
    t.start()
    return t


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--iters-per-step', type=int, default=10)
    ap.add_argument('--cpu-budget-s', type=float, default=12.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--watchdog-s', type=float, default=float(os.environ.get('RLP_BENCH_WATCHDOG_S', '480')))
    args = ap.parse_args()
    if args.watchdog_s > 0:
        arm_watchdog(args.watchdog_s)
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from rlpoker_b200 import _lib
    from rlpoker_b200._lib import RLP_AVERAGE, RLP_REGRET, RLP_STRAT_SUM
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    ft = build_tree()
    N, Q = ft.num_nodes, ft.num_pairs
    ipS = args.iters_per_step

    if world == 1:
        from rlpoker_b200.solver import DeviceCFR, DeviceTree
        solver = DeviceCFR(DeviceTree(ft))
        step = lambda: solver.vanilla(ipS)
        core = solver
    else:
        from rlpoker_b200.sharding import ShardedCFR
        sh = ShardedCFR(ft, rank, world)
        step = lambda: sh.iterate(ipS)
        core = sh.solver

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device='cuda')      # 4x the 126 MB L2

    def timed(fn, steps, flush=True):
        """Device time of exactly `steps` steps (sum of per-step CUDA-event intervals; with flush=True a 512 MB
        write evicts L2 between steps, outside the timed intervals); max over ranks."""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for a, b in evs:
            if flush:
                flush_buf.fill_(1)
            a.record()
            fn()
            b.record()
        barrier()
        ms = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], device='cuda')
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident throughput -----------------------------------------------------------------
    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = _lib.load().rlp_kernel_launches()
    ms = timed(step, args.steps, flush=True)
    launches = _lib.load().rlp_kernel_launches() - launches0
    clocks = sampler.stop() if sampler else None
    updates = 2.0 * N * ipS * args.steps
    value = updates / (ms * 1e-3)
    ms_warm = timed(step, args.steps, flush=False)           # back-to-back steps, L2 left alone (how a solve runs)

    # ---- end to end with host buffers ----------------------------------------------------------------
    host_R = torch.zeros(Q, dtype=torch.float64).pin_memory()
    host_S = torch.zeros(Q, dtype=torch.float64).pin_memory()
    host_avg = torch.zeros(Q, dtype=torch.float64).pin_memory()
    host_R.numpy()[:] = core.read(RLP_REGRET)
    host_S.numpy()[:] = core.read(RLP_STRAT_SUM)
    lib = _lib.load()
    from rlpoker_b200.solver import _stream_ptr

    def e2e_step():
        _lib.check(lib.rlp_cfr_write(core.handle, RLP_REGRET, _lib.ptr(host_R.numpy(), _lib._dp), _stream_ptr(None)))
        _lib.check(lib.rlp_cfr_write(core.handle, RLP_STRAT_SUM, _lib.ptr(host_S.numpy(), _lib._dp), _stream_ptr(None)))
        step()
        _lib.check(lib.rlp_cfr_read(core.handle, RLP_AVERAGE, _lib.ptr(host_avg.numpy(), _lib._dp), _stream_ptr(None)))

    for _ in range(3):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps, flush=True)
    e2e_value = updates / (ms_e2e * 1e-3)

    # ---- per-stage device time, live (CUDA events between the stages; serialised, so their sum exceeds the
    # pipelined iteration time -- shares are what matters) -------------------------------------------------
    stages = core.profile_stages(30) if world == 1 else None
    layout = core.tree.layout

    # ---- report -------------------------------------------------------------------------------------------
    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        bytes_iter = algorithmic_bytes_per_iteration(ft)
        iter_ms = ms / (args.steps * ipS)
        achieved = (bytes_iter / world) / (iter_ms * 1e-3) / 1e9          # per GPU
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': 'leduc13-vanilla-cfr', 'game': 'Leduc(get_deck(13, 2))', 'nodes': N,
                       'infosets': ft.num_infosets, 'pairs': Q, 'iters_per_step': ipS,
                       'parallelism': 'deal-sharded x%d + allreduce' % world if world > 1 else 'single GPU',
                       'l2_policy': 'L2 flushed (512 MB write) between timed steps, outside the timed intervals; one '
                                    'iteration touches ~76 MB of HBM-resident state (ncu), less than the 126 MB L2'},
            'steady_state': {'value': updates / (ms_warm * 1e-3), 'us_per_iteration': 1e3 * ms_warm / (args.steps * ipS),
                             'note': 'same K steps without the L2 flush between steps'},
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': 16 * Q, 'd2h_bytes_per_step': 8 * Q,
                    'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': int(launches),
            'clocks': clocks,
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': traffic_per_iteration(), 'peak_source': peak_src,
                         'kernel': 'whole CFR iteration (upper_k0 -> micro_k -> k_fanin -> upper_k1 -> k_accumulate_rm4 x2)',
                         'algorithmic_bytes_per_iteration': bytes_iter,
                         'bytes_per_node_update': bytes_iter / (2.0 * N), 'us_per_iteration': iter_ms * 1e3,
                         'stage_us_serialised': stages, 'layout': layout,
                         'dominant_kernel': dominant_kernel_roofline(layout, stages, peak, 8 * Q),
                         'note': 'algorithmic bytes = SURVEY 8d formula (level sweep with reach/value arrays in HBM); the '
                                 'tiled sweep keeps those on chip, so DRAM traffic (ncu) is ~4x lower than the formula'},
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v1, n1, d1 = time_cpu(ft, args.cpu_budget_s / 2, 1)
            vN, nN, dN = time_cpu(ft, args.cpu_budget_s / 2, threads)
            line['cpu_baseline'] = {'value': vN, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                                    'sample': '%d Leduc-13 CFR iterations in %.1f s on %d threads (C port of the '
                                              'reference algorithm, oracle/cfr_oracle.c)' % (nN, dN, threads),
                                    'single_thread_value': v1}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
