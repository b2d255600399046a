"""profiles/r2_ncu_summary.json and profiles/r2_fused_traffic.json from the raw pages of the round-2 ncu captures.

    ncu -i gpurun_out/r2_fused_full.ncu-rep --page raw --csv > profiles/r2_fused_full_raw.csv      (same for the others)
    python tools/ncu_summary_r2.py

One captured launch per file: cfr_fused_k (10 iterations of Leduc-13), br_fused_k, and one mid-level k_front_expand /
k_front_collect of 2^20 external-sampling lanes on Leduc-3 (tools/gpu_profile_r2.sh)."""
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
METRICS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
    'launch__grid_size', 'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
    'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors.sum',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
]
FILES = {'cfr_fused_k': 'r2_fused_full_raw.csv', 'br_fused_k': 'r2_br_full_raw.csv',
         'k_front_expand': 'r2_lanes_expand_raw.csv', 'k_front_collect': 'r2_lanes_collect_raw.csv'}


def first_launch(path):
    rows = list(csv.reader(open(path)))
    hdr, units, row = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    out = {'kernel': row[col['Kernel Name']][:90]}
    for m in METRICS:
        out[m] = [row[col[m]], units[col[m]]] if m in col else [None, None]
    return out


def to_bytes(entry):
    return float(entry[0]) * SCALE[entry[1]]


def main():
    summary = {}
    for name, f in FILES.items():
        path = os.path.join(ROOT, 'profiles', f)
        if os.path.exists(path):
            summary[name] = first_launch(path)
    with open(os.path.join(ROOT, 'profiles', 'r2_ncu_summary.json'), 'w') as f:
        json.dump(summary, f, indent=1)
    k = summary['cfr_fused_k']
    rd, wr = to_bytes(k['dram__bytes_read.sum']), to_bytes(k['dram__bytes_write.sum'])
    traffic = {
        'source': 'profiles/r2_fused_full_raw.csv (ncu --set full, one launch of cfr_fused_k = 10 iterations, Leduc-13)',
        'iterations_per_launch': 10, 'dram_read_bytes_per_launch': rd, 'dram_write_bytes_per_launch': wr,
        'dram_bytes_per_iteration': (rd + wr) / 10, 'kernel_us_per_launch_under_ncu': float(k['gpu__time_duration.sum'][0]),
        'note': 'the ~30 MB working set of an iteration (static group tables, strategy, regrets, subtree values) stays in '
                'the 126 MB L2; DRAM sees ~1.5 MB per iteration',
    }
    with open(os.path.join(ROOT, 'profiles', 'r2_fused_traffic.json'), 'w') as f:
        json.dump(traffic, f, indent=1)
    return summary, traffic


if __name__ == '__main__':
    s, t = main()
    for name, k in s.items():
        print(name, {m.split('.')[0].replace('smsp__average_warps_issue_stalled_', 'stall_'): k[m][0] for m in METRICS})
    print(t)
