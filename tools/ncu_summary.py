"""profiles/traffic.json from an `ncu --set full` capture.

    ncu -i gpurun_out/full_TAG.ncu-rep --page raw --csv > profiles/rN_TAG_full_raw.csv
    python tools/ncu_summary.py profiles/rN_TAG_full_raw.csv > profiles/traffic.json

One row per captured launch; the capture is expected to hold exactly one iteration's worth of sweep kernels
(upper_k0, micro_k, k_accumulate_rm4 [A], k_fanin, upper_k1, k_accumulate_rm4 [B]) -- the second occurrence of a kernel
name gets the suffix `_B`.  `dram_bytes_per_iteration` is the sum of dram__bytes_read + dram__bytes_write over them
(what bench.py reports as roofline.traffic)."""
import csv
import json
import re
import sys

SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}


def short_name(full):
    m = re.search(r'(\w+(?:<[^>]*>)?)\s*\(', full)
    return m.group(1) if m else full


def summarise(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name, to=None):
        x = float(r[col[name]])
        return x * SCALE[units[col[name]]] / (SCALE[to] if to else 1.0) if to else x

    kernels, total = {}, 0.0
    for r in data:
        name = short_name(r[col['Kernel Name']])
        if name in kernels:
            name += '_B'
        rd, wr = val(r, 'dram__bytes_read.sum', 'Mbyte'), val(r, 'dram__bytes_write.sum', 'Mbyte')
        total += (rd + wr) * 1e6
        kernels[name] = {
            'us': val(r, 'gpu__time_duration.sum', 'us'), 'dram_read_MB': rd, 'dram_write_MB': wr,
            'regs': int(float(r[col['launch__registers_per_thread']])), 'grid': r[col['launch__grid_size']],
            'warps_active_pct': val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'),
            'lts_pct': val(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'),
            'dram_pct': val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
            'ld_sectors': r[col['l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum']],
            'st_sectors': r[col['l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum']],
            'inst': r[col['smsp__inst_executed.sum']],
        }
    return {'source': '%s (ncu --set full --clock-control none, cold caches, one launch of each sweep kernel)' % path,
            'dram_bytes_per_iteration': total, 'kernels': kernels}


if __name__ == '__main__':
    print(json.dumps(summarise(sys.argv[1]), indent=1))
