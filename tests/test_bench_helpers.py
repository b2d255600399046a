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


def test_round2_summaries_are_reproducible_from_the_committed_captures(tmp_path):
    """profiles/r2_ncu_summary.json and r2_fused_traffic.json (what bench.py reports as roofline.traffic for the fused
    kernel) are derived from the committed raw pages by tools/ncu_summary_r2.py."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import ncu_summary_r2
    with open(os.path.join(ROOT, 'profiles', 'r2_ncu_summary.json')) as f:
        want_summary = json.load(f)
    with open(os.path.join(ROOT, 'profiles', 'r2_fused_traffic.json')) as f:
        want_traffic = json.load(f)
    got = {name: ncu_summary_r2.first_launch(os.path.join(ROOT, 'profiles', fn)) for name, fn in ncu_summary_r2.FILES.items()}
    assert json.loads(json.dumps(got)) == want_summary
    k = got['cfr_fused_k']
    per_iter = (ncu_summary_r2.to_bytes(k['dram__bytes_read.sum']) + ncu_summary_r2.to_bytes(k['dram__bytes_write.sum'])) / 10
    assert per_iter == want_traffic['dram_bytes_per_iteration'] == bench.traffic_per_iteration(fused=True)
    assert per_iter < 0.2 * bench.FUSED_MICRO_BYTES * 140400          # 1.5 MB against 13.5 MB: the working set is L2-resident
