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


def traffic_per_iteration(fused=False):
    """dram__bytes_read+write per iteration summed over the sweep kernels, from the committed ncu capture
    (profiles/traffic.json: five-kernel pipeline; profiles/r2_fused_traffic.json: the fused persistent kernel)."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'r2_fused_traffic.json' if fused else 'traffic.json')) as f:
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


# Group-fused persistent kernel (csrc/fused.cu): bytes one iteration must move when every table element is charged
# once per pass that reads or writes it (the storage format differs from SURVEY 8d's level sweep, so the formula is
# restated here as the survey asks; both are reported).  Leduc shapes: 8 decision nodes / 15 leaves per micro
# subtree, 8 decision nodes and 9 board-deal chance nodes per upper tile.
#   micro subtree, per pass (it is evaluated once per player): 4 opponent information-set ids (i32), payoff pattern id,
#     opponent-reach index (i32), static chance reach (f64), the opponent's reach at its root (f64) [the own reach, the
#     own ids and the 16-byte payoff pattern are per GROUP of 2n-2 subtrees]; every pass writes the subtree value for
#     its player (f64, block layout [pass][upper group][portal][member]) -- to every rank's buffer when sharded
#   upper tile, per pass: the values of its 9 x (2n-2) micro subtrees (f64) and 22 strategy probabilities
#   (information set, action) pair: regret and action count read + written, strategy read, written twice
#     (row-major for the API, fused numbering for the next iteration)
FUSED_MICRO_BYTES_PER_PASS = 4 * 4 + 4 + 4 + 8 + 8
FUSED_MICRO_BYTES = 2 * (FUSED_MICRO_BYTES_PER_PASS + 8)
FUSED_PAIR_BYTES = 4 * 8 + 8 + 2 * 8


def fused_bytes_per_iteration(ft, layout, world=1):
    """Per GPU: the micro subtrees and their information sets are divided over the ranks, the upper tiles replicated."""
    boards = RANKS * 2 - 2
    tile = 2 * (9 * boards * 8 + 22 * 8)
    micro = layout['micro_subtrees'] * (FUSED_MICRO_BYTES + 2 * 8 * (world - 1))
    return micro / world + layout['upper_tiles'] * tile + ft.num_pairs * FUSED_PAIR_BYTES / world


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
    """The CPU port (oracle/) on the product's arrays -- the checker of the parity block."""
    from oracle import flat, oracle as orc
    t = flat.from_arrays(ft.num_nodes, ft.parent, ft.first_child, ft.player, ft.infoset, ft.label, ft.hidden,
                         ft.cprob, ft.u1, ft.u2)
    return orc.Oracle(t)


def oracle_own_tree():
    """The CPU port on a Leduc-13 tree built by the oracle's OWN generator (oracle/leduc_oracle.c): the CPU arms do not
    touch the product package."""
    from oracle import oracle as orc
    return orc.Oracle(orc.leduc_tree(RANKS))


def time_cpu(budget_s, threads, o=None):
    o = o or oracle_own_tree()
    N = o.t.num_nodes
    t0 = time.perf_counter()
    o.vanilla_levels(1, threads=threads)
    one = time.perf_counter() - t0
    n = max(1, min(5000, int(budget_s / max(one, 1e-3))))
    t0 = time.perf_counter()
    o.vanilla_levels(n, threads=threads)
    dt = time.perf_counter() - t0
    return 2.0 * N * n / dt, n, dt


def run_reference(args, rank, world):
    """--impl reference: the reference algorithm's CPU port on all host cores, same tree / metric / step (the
    reference itself is Python and cannot travel to the GPU box; BASELINE.md has its container figure).  One step =
    ``iters_per_step`` iterations in ONE call, exactly like the GPU arm."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    o = oracle_own_tree()
    N, I = o.t.num_nodes, o.t.num_infosets
    ipS = args.iters_per_step
    for _ in range(args.warmup):
        o.vanilla_levels(ipS, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.vanilla_levels(ipS, threads=threads)
    dt = time.perf_counter() - t0
    value = 2.0 * N * ipS * args.steps / dt
    sample = '%d steps of %d Leduc-13 CFR iterations on %d threads (C port of the reference algorithm)' % (
        args.steps, ipS, threads)
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'leduc13-vanilla-cfr', 'game': 'Leduc(get_deck(13, 2))', 'nodes': N, 'infosets': I,
                   'pairs': int(o.Q), 'iters_per_step': ipS},
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
    ap.add_argument('--steps', type=int, default=None)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--iters-per-step', type=int, default=10)
    ap.add_argument('--cpu-budget-s', type=float, default=12.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--watchdog-s', type=float, default=float(os.environ.get('RLP_BENCH_WATCHDOG_S', '480')))
    ap.add_argument('--repeats', type=int, default=10, help='extra repeats of the K-step block for the spread')
    ap.add_argument('--parity-iters', type=int, default=23)
    ap.add_argument('--no-extras', action='store_true', help='skip the e2e_api / extra.* legs')
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 50 if args.impl == 'b200' else 5
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
        run_iters = solver.vanilla
    else:
        # default: sharding by public state (peer stores inside the persistent kernel, bit-exact); RLP_BENCH_SHARDING=deal
        # selects the deal-sharded solve with one NCCL all-reduce of the regret / action-count deltas per iteration
        sharding = os.environ.get('RLP_BENCH_SHARDING', 'public')
        if sharding == 'deal':
            from rlpoker_b200.sharding import ShardedCFR
            sh = ShardedCFR(ft, rank, world)
        else:
            from rlpoker_b200.sharding import PublicStateShardedCFR
            sh = PublicStateShardedCFR(ft, rank, world)
        step = lambda: sh.iterate(ipS)
        core = sh.solver
        run_iters = sh.iterate
    layout = core.tree.layout
    public = world > 1 and os.environ.get('RLP_BENCH_SHARDING', 'public') != 'deal'
    if os.environ.get('RLP_NO_JIT') != '1' and os.environ.get('RLP_NO_TILES') != '1':
        # the numbers below are those of the straight-line NVRTC kernels; anything else is a different program
        assert layout['tiled'] == 1 and layout['jit_shapes'] >= 1 and layout['upper_jit'] == 1, layout
        if world == 1 and os.environ.get('RLP_NO_FUSED') != '1':
            assert layout['fused'] == 1, layout

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
    updates = 2.0 * N * ipS * args.steps
    value = updates / (ms * 1e-3)
    # the same K-step block again, `repeats` times: spread of the headline number
    rep = [updates / (timed(step, args.steps, flush=True) * 1e-3) for _ in range(max(0, args.repeats))]
    clocks = sampler.stop() if sampler else None
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
    e2e_rep = [updates / (timed(e2e_step, args.steps, flush=True) * 1e-3) for _ in range(min(5, max(0, args.repeats)))]

    # ---- per-stage device time, live (CUDA events between the stages; serialised, so their sum exceeds the
    # pipelined iteration time -- shares are what matters) -------------------------------------------------
    fused = layout.get('fused') == 1 and os.environ.get('RLP_NO_FUSED') != '1'
    if world > 1 and not public:
        stages = None
    elif fused:
        stages = core.fused_timeline(40)          # CTA 0's view of the phases inside the persistent kernel
    else:
        stages = core.profile_stages(30)

    # ---- parity: the program that was just timed, against the CPU oracle on the same tree ------------------
    # single GPU: regrets / action counts / strategy bit for bit; N > 1 (all-reduce changes the association): 1e-9.
    T = args.parity_iters
    core.reset()
    if world > 1:
        sh.solver.t = 0
    run_iters(T)
    torch.cuda.synchronize()
    if public:
        sh.check()
        got_R, got_S = sh.merged(RLP_REGRET), sh.merged(RLP_STRAT_SUM)
        sh.gather_state()
    else:
        got_R, got_S = core.read(RLP_REGRET), core.read(RLP_STRAT_SUM)
    got_avg = core.read(RLP_AVERAGE)
    got_expl = core.exploitability()[0]
    parity = None
    if rank == 0:
        o = oracle_for(ft)
        o.vanilla_levels(T, threads=os.cpu_count() or 1)
        want_avg = o.average_strategy()[0]
        want_expl = o.exploitability(want_avg)
        bit = bool(np.array_equal(got_R, o.R) and np.array_equal(got_S, o.S))
        parity = {'iterations': T, 'game': 'Leduc(get_deck(13, 2))', 'oracle': 'oracle/cfr_oracle.c level sweep',
                  'bit_exact': bit, 'max_abs_regret': float(np.abs(got_R - o.R).max()),
                  'max_abs_avg': float(np.abs(got_avg - want_avg).max()),
                  'abs_exploitability': abs(got_expl - want_expl), 'exploitability': got_expl,
                  'tolerance': {'avg': 1e-9, 'exploitability': 1e-7, 'bit_exact_required': world == 1 or public}}
        parity['ok'] = bool(parity['max_abs_avg'] <= 1e-9 and parity['abs_exploitability'] <= 1e-7 and
                            (bit or (world > 1 and not public)) and parity['max_abs_regret'] <= 1e-9)

    # ---- the public Python API, wall clock (cfr() as a user calls it, FlatGame = no node objects) -----------
    e2e_api = None
    extra = None
    if not args.no_extras:
        # free the Leduc-13 solver first: the extras build their own trees (Leduc-25 needs ~3 GB)
        if world == 1:
            e2e_api = time_python_api(ft, N)
        extra = {}
        if world == 1:
            extra['br_sweep'] = br_sweep()
            extra['config3'] = sampled_config3()
            extra['exploitability_vs_time'] = exploitability_vs_time()
            extra['config1'] = readme_config1()
        else:
            extra['config3'] = sampled_config3_sharded(rank, world)
        extra['leduc25'] = leduc25(rank, world, public)

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
                       'parallelism': ('single GPU' if world == 1 else
                                       'public-state components sharded x%d, subtree values exchanged by peer stores inside the '
                                       'persistent kernel' % world if public else 'deal-sharded x%d + NCCL allreduce' % world),
                       'l2_policy': 'L2 flushed (512 MB write) between timed steps, outside the timed intervals; one '
                                    'iteration touches ~76 MB of HBM-resident state (ncu), less than the 126 MB L2'},
            'spread': ({'repeats': len(rep), 'median': statistics.median(rep), 'min': min(rep), 'max': max(rep),
                        'note': 'the K-step timed block repeated; `value` is the first block'} if rep else None),
            'steady_state': {'value': updates / (ms_warm * 1e-3), 'us_per_iteration': 1e3 * ms_warm / (args.steps * ipS),
                             'note': 'same K steps without the L2 flush between steps'},
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': 16 * Q, 'd2h_bytes_per_step': 8 * Q,
                    'ms_per_step': ms_e2e / args.steps,
                    'spread': ({'repeats': len(e2e_rep), 'median': statistics.median(e2e_rep), 'min': min(e2e_rep),
                                'max': max(e2e_rep)} if e2e_rep else None),
                    'path': 'C ABI with pinned host buffers: rlp_cfr_write x2 -> rlp_cfr_vanilla_iters -> rlp_cfr_read'},
            'e2e_api': e2e_api,
            'extra': extra,
            'gpu_launches': int(launches),
            'clocks': clocks,
            'parity': parity,
            'roofline': roofline_block(ft, layout, stages, fused, world, bytes_iter, iter_ms, achieved, peak, peak_src),
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v1, n1, d1 = time_cpu(args.cpu_budget_s / 2, 1)
            vN, nN, dN = time_cpu(args.cpu_budget_s / 2, threads)
            line['cpu_baseline'] = {'value': vN, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                                    'sample': '%d Leduc-13 CFR iterations in %.1f s on %d threads (C port of the '
                                              'reference algorithm, oracle/cfr_oracle.c)' % (nN, dN, threads),
                                    'single_thread_value': v1}
        print(json.dumps(line))
        l25 = (extra or {}).get('leduc25') or {}
        if l25.get('parity') is not None and not l25['parity']['bit_exact']:
            parity['ok'] = False
            parity['leduc25_failed'] = True
        if not parity['ok']:
            sys.stderr.write('bench.py: PARITY FAILED: %s\n' % json.dumps(parity))
            sys.stderr.flush()
            if world > 1:
                dist.destroy_process_group()
            sys.exit(4)
    if world > 1:
        dist.destroy_process_group()


def roofline_block(ft, layout, stages, fused, world, bytes_iter, iter_ms, achieved, peak, peak_src):
    N, Q = ft.num_nodes, ft.num_pairs
    r = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
         'traffic': traffic_per_iteration(fused) if world == 1 else None, 'peak_source': peak_src,
         'algorithmic_bytes_per_iteration': bytes_iter, 'bytes_per_node_update': bytes_iter / (2.0 * N),
         'us_per_iteration': iter_ms * 1e3, 'layout': layout}
    if fused:
        fb = fused_bytes_per_iteration(ft, layout, world)
        r.update({
            'kernel': 'cfr_fused_k: ONE cooperative persistent launch per call; per iteration a micro phase (subtree '
                      'groups: evaluate + ordered accumulate + regret matching) and an upper phase (round-1 tiles), two '
                      'grid barriers',
            'stage_us': stages,
            'restated': {'algorithmic_bytes_per_iteration': fb, 'achieved': fb / (iter_ms * 1e-3) / 1e9,
                         'frac': fb / (iter_ms * 1e-3) / 1e9 / peak,
                         'bytes_per_micro_subtree': FUSED_MICRO_BYTES, 'bytes_per_pair': FUSED_PAIR_BYTES, 'per_gpu': True,
                         'note': 'bytes of the fused storage format (no reach / value / contribution arrays leave the '
                                 'SM); the ~30 MB working set is L2-resident, so the iteration is bound by FP64 issue, '
                                 'shared-memory ordered sums and two grid barriers, not by HBM'},
            'note': 'achieved/frac = SURVEY 8d formula (level sweep with reach/value arrays in HBM) over the measured '
                    'iteration time, as the survey defines the metric; `restated` charges what this design moves'})
    else:
        r.update({'kernel': 'whole CFR iteration (upper_k0 -> micro_k -> k_fanin -> upper_k1 -> k_accumulate_rm4 x2)',
                  'stage_us_serialised': stages,
                  'dominant_kernel': dominant_kernel_roofline(layout, stages, peak, 8 * Q),
                  'note': 'algorithmic bytes = SURVEY 8d formula (level sweep with reach/value arrays in HBM); the tiled '
                          'sweep keeps those on chip, so DRAM traffic (ncu) is ~4x lower than the formula'})
    return r


def br_sweep(ranks=(3, 5, 7, 9, 11, 13, 17, 21, 25)):
    """BASELINE config 5: exploitability (both best responses) of a 10-iteration CFR average on Leduc-n, the device call
    (rlp_cfr_exploitability: one cooperative launch on the micro / upper groups + the host read-back) timed with CUDA
    events, the oracle's best response (oracle/cfr_oracle.c, one host thread, recursion over information-set node
    lists like best_response.py:9-100) timed beside it.  Best response runs as replicas on N > 1."""
    import torch
    from oracle import oracle as orc
    from rlpoker_b200.games.leduc_native import leduc_flat_tree
    from rlpoker_b200.solver import DeviceCFR, DeviceTree
    rows = []
    for n in ranks:
        ft = leduc_flat_tree(n)
        dt = DeviceTree(ft)
        s = DeviceCFR(dt)
        s.vanilla(10)
        s.exploitability()
        torch.cuda.synchronize()
        reps = 20 if n <= 13 else 5
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            v = s.exploitability()
        b.record()
        torch.cuda.synchronize()
        avg = s.average
        o = orc.Oracle(orc.leduc_tree(n))
        t0 = time.perf_counter()
        want = o.exploitability(avg)
        cpu_s = time.perf_counter() - t0
        rows.append({'n': n, 'nodes': ft.num_nodes, 'gpu_us': a.elapsed_time(b) * 1e3 / reps, 'cpu_s': cpu_s, 'cpu_threads': 1,
                     'exploitability': v[0], 'abs_diff_vs_cpu': abs(v[0] - want), 'fused_br': dt.layout.get('fused_br')})
        del s, o
        dt.close()
    return rows


def exploitability_vs_time(checkpoints=(10, 30, 100, 300, 1000, 3000)):
    """The second half of BASELINE's metric: exploitability of the average strategy against wall-clock solve time, vanilla
    CFR on Leduc-13 from scratch, one GPU.  `seconds` is host wall clock from the first iteration to the moment the
    exploitability of that checkpoint is on the host (every evaluation so far included).  Beside it the CPU port on all
    host threads up to the checkpoint it reaches within ~6 s; equal iteration counts must give equal exploitability."""
    import torch
    from oracle import oracle as orc
    from rlpoker_b200.games.leduc_native import leduc_flat_tree
    from rlpoker_b200.solver import DeviceCFR, DeviceTree
    dt = DeviceTree(leduc_flat_tree(RANKS))
    s = DeviceCFR(dt)
    s.vanilla(10)
    s.exploitability()
    s.reset()
    torch.cuda.synchronize()
    points, done = [], 0
    t0 = time.perf_counter()
    for T in checkpoints:
        s.vanilla(T - done)
        done = T
        e = s.exploitability()[0]           # synchronises: the value is on the host
        points.append({'iterations': T, 'seconds': time.perf_counter() - t0, 'exploitability': e})
    del s
    dt.close()
    threads = os.cpu_count() or 1
    o = orc.Oracle(orc.leduc_tree(RANKS))
    cpu, done = [], 0
    t0 = time.perf_counter()
    for T in checkpoints:
        if cpu and (time.perf_counter() - t0) / done * T > 6.0:
            break
        o.vanilla_levels(T - done, threads=threads)
        done = T
        e = o.exploitability(o.average_strategy()[0])
        cpu.append({'iterations': T, 'seconds': time.perf_counter() - t0, 'exploitability': e})
    for c, g in zip(cpu, points):
        c['abs_diff_vs_gpu'] = abs(c['exploitability'] - g['exploitability'])
    return {'game': 'Leduc(get_deck(13, 2))', 'algorithm': 'vanilla CFR, average strategy', 'gpu': points,
            'cpu_port': {'threads': threads, 'points': cpu}}


def sampled_config3(lanes=1 << 20):
    """BASELINE config 3: 2^20 traversal lanes per iteration on Leduc-3 (external sampling; chance sampling beside it)."""
    import torch
    from rlpoker_b200.games.leduc_native import leduc_flat_tree
    from rlpoker_b200.solver import DeviceCFR, DeviceTree
    dt = DeviceTree(leduc_flat_tree(3))
    out = {'lanes': lanes, 'game': 'Leduc(get_deck(3, 2))'}
    for variant, name in ((1, 'external'), (0, 'chance')):
        s = DeviceCFR(dt)
        s.sampled(variant, 2, lanes=lanes, seed=1)
        torch.cuda.synchronize()
        n0 = s.nodes_touched
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        s.sampled(variant, 4, lanes=lanes, seed=1)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 4
        out[name] = {'ms_per_iteration': ms, 'node_updates_per_s': (s.nodes_touched - n0) / 4 / (ms * 1e-3)}
        del s
    dt.close()
    return out


def sampled_config3_sharded(rank, world, lanes=1 << 20):
    """BASELINE config 3 at N > 1: the 2^20 lanes of every iteration split over the ranks (ShardedSampledCFR: one NCCL
    sum of the [2Q] increments and one max of the membership flags per iteration); device time, max over ranks."""
    import torch
    import torch.distributed as dist
    from rlpoker_b200.games.leduc_native import leduc_flat_tree
    from rlpoker_b200.sharding import ShardedSampledCFR
    out = {'lanes': lanes, 'game': 'Leduc(get_deck(3, 2))', 'sharding': 'lanes x%d + NCCL all-reduce per iteration' % world}
    for variant, name in ((1, 'external'), (0, 'chance')):
        sh = ShardedSampledCFR(leduc_flat_tree(3), rank, world)
        sh.iterate(variant, 3, lanes, seed=1)
        torch.cuda.synchronize()
        n0 = sh.nodes_touched()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        sh.iterate(variant, 8, lanes, seed=1)
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / 8], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        out[name] = {'ms_per_iteration': ms, 'node_updates_per_s': (sh.nodes_touched() - n0) / 8 / (ms * 1e-3)}
        sh.close()
    return out


def leduc25(rank, world, public):
    """Leduc-25 (24 399 601 nodes): generation, upload, time per iteration on this many GPUs, and bit-exact parity of two
    more iterations against the oracle (rank 0)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from rlpoker_b200._lib import RLP_REGRET, RLP_STRAT_SUM
    from rlpoker_b200.games.leduc_native import leduc_flat_tree
    from rlpoker_b200.solver import DeviceCFR, DeviceTree
    t0 = time.perf_counter()
    ft = leduc_flat_tree(25)
    gen_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    if world == 1:
        dt = DeviceTree(ft)
        core = DeviceCFR(dt)
        run = core.vanilla
        sh = None
    else:
        if not public:
            return {'skipped': 'deal sharding is benchmarked on Leduc-13 only'}
        from rlpoker_b200.sharding import PublicStateShardedCFR
        sh = PublicStateShardedCFR(ft, rank, world)
        core, dt, run = sh.solver, sh.tree, sh.iterate
    torch.cuda.synchronize()
    upload_s = time.perf_counter() - t0
    run(10)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        run(10)
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b)], device='cuda')
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    us_iter = float(ms.item()) * 1e3 / 50
    core.reset()
    run(2)
    torch.cuda.synchronize()
    if sh is not None:
        sh.check()
        got_R, got_S = sh.merged(RLP_REGRET), sh.merged(RLP_STRAT_SUM)
    else:
        got_R, got_S = core.read(RLP_REGRET), core.read(RLP_STRAT_SUM)
    out = {'game': 'Leduc(get_deck(25, 2))', 'nodes': ft.num_nodes, 'infosets': ft.num_infosets, 'generate_s': gen_s,
           'upload_tile_compile_s': upload_s, 'tree_device_bytes': dt.device_bytes, 'n_gpus': world,
           'us_per_iteration': us_iter, 'node_updates_per_s': 2.0 * ft.num_nodes / (us_iter * 1e-6), 'layout': dt.layout}
    if rank == 0:
        o = oracle_for(ft)
        o.vanilla_levels(2, threads=os.cpu_count() or 1)
        out['parity'] = {'iterations': 2, 'bit_exact': bool(np.array_equal(got_R, o.R) and np.array_equal(got_S, o.S)),
                         'max_abs_regret': float(np.abs(got_R - o.R).max())}
    return out


def time_python_api(ft, N):
    """Wall clock of the reference-shaped call ``cfr(game, num_iters=100, use_chance_sampling=False)`` on Leduc-13, with
    the driver's default keyword arguments (immediate regret + the per-iteration strategy history, i.e. one Python
    trip with two device->host reads per iteration) and with the lean ones (no history, no immediate regret: the
    iterations run in blocks of 10 between the evaluations)."""
    from rlpoker_b200.cfr.cfr import cfr
    from rlpoker_b200.games.leduc_native import FlatGame
    game = FlatGame(ft, 'Leduc values=13 suits=2')
    out = {'call': 'cfr(game, num_iters=100, use_chance_sampling=False)', 'unit': UNIT}
    cfr(game, num_iters=1, use_chance_sampling=False, verbose=False)          # upload + tiling + NVRTC: not timed
    for name, kw in (('default', {}), ('lean', {'immediate_regret': False, 'keep_strategies': 0})):
        t0 = time.perf_counter()
        avg, expl, hist = cfr(game, num_iters=100, use_chance_sampling=False, verbose=False, **kw)
        dt = time.perf_counter() - t0
        out[name] = {'seconds': dt, 'value': 2.0 * N * 100 / dt, 'evaluations': len(expl),
                     'final_exploitability': expl[-1][1], 'kwargs': kw}
    return out


def readme_config1(iters=2000):
    """BASELINE config 1: the README's call -- chance-sampling ``cfr()`` on ``Leduc.create_game(3)`` plus the
    best-response exploitability -- through the Python API, one traversal lane per iteration like the reference
    (wall clock; game construction and upload not timed).  BASELINE.md section 2 measured the reference at 29.7 ms per
    chance-sampled iteration and 0.10 s per compute_exploitability call (container CPU, one thread)."""
    from rlpoker_b200.best_response import compute_exploitability
    from rlpoker_b200.cfr.cfr import cfr
    from rlpoker_b200.games.leduc import Leduc
    game = Leduc.create_game(3)
    cfr(game, num_iters=10, use_chance_sampling=True, verbose=False)
    out = {'call': 'cfr(Leduc.create_game(3), num_iters=%d, use_chance_sampling=True) + compute_exploitability' % iters}
    for name, kw in (('default', {}), ('lean', {'immediate_regret': False, 'keep_strategies': 0})):
        t0 = time.perf_counter()
        avg, expl, hist = cfr(game, num_iters=iters, use_chance_sampling=True, verbose=False, **kw)
        t1 = time.perf_counter()
        e = compute_exploitability(game, avg)
        t2 = time.perf_counter()
        out[name] = {'cfr_seconds': t1 - t0, 'ms_per_iteration': 1e3 * (t1 - t0) / iters, 'exploitability_seconds': t2 - t1,
                     'exploitability': e, 'kwargs': kw}
    out['reference'] = {'ms_per_iteration': 29.7, 'exploitability_seconds': 0.10, 'source': 'BASELINE.md section 2'}
    return out


if __name__ == '__main__':
    main()
