"""Host-side arithmetic of bench.py (no GPU): algorithmic byte counts and the per-kernel roofline."""
import json
import os

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_algorithmic_bytes_match_survey_8d():
    from rlpoker_b200.games.leduc_native import leduc_flat_tree
    assert bench.algorithmic_bytes_per_iteration(leduc_flat_tree(3)) == 2678272        # SURVEY 8(d), n = 3
    assert bench.algorithmic_bytes_per_iteration(leduc_flat_tree(5)) == 14834464       # n = 5


def test_dominant_kernel_roofline_from_the_committed_run():
    with open(os.path.join(ROOT, 'profiles', 'r1_v5_bench.json')) as f:
        r = json.load(f)['roofline']
    d = bench.dominant_kernel_roofline(r['layout'], r['stage_us_serialised'], r['peak'], 8 * 129272)
    assert d['kernel'] == 'micro_k' and d['units'] == 140400 and d['bytes_per_unit'] == 464
    assert d['algorithmic_bytes_per_launch'] == 140400 * 464 + 8 * 129272
    assert abs(d['frac'] - d['achieved'] / r['peak']) < 1e-15 and 0.3 < d['frac'] < 1.0
    assert d['traffic'] is not None and d['traffic'] < d['algorithmic_bytes_per_launch']
    assert bench.dominant_kernel_roofline(None, None, 1.0, 0) is None
    assert bench.dominant_kernel_roofline({'micro_subtrees': 0}, {'micro': 1.0}, 1.0, 0) is None


def test_traffic_json_is_reproducible_from_the_committed_capture():
    import sys
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import ncu_summary
    got = ncu_summary.summarise(os.path.join(ROOT, 'profiles', 'r1_v5_full_raw.csv'))
    with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
        want = json.load(f)
    assert got['dram_bytes_per_iteration'] == want['dram_bytes_per_iteration'] == bench.traffic_per_iteration()
    assert got['kernels'] == want['kernels']
