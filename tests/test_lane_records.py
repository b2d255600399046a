"""Host logic of the level-by-level sampled lanes: the per-level record capacities (csrc/sampled.cu::lane_widths through
rlp_tree_plan_lane_records, no GPU) against a plain restatement, and against simulated traversals -- a capacity that is
too small would make the device drop records (it flags that instead of writing out of bounds), one that is too large only
wastes memory."""
import numpy as np
import pytest

from conftest import make_game


def _chain(ft, x, me, variant, memo):
    """Records (by relative explored level) of a walk that steps onto node x: the restated recurrence."""
    key = (x, me)
    if key in memo:
        return memo[key]
    pl = int(ft.player[x])
    kids = range(int(ft.first_child[x]), int(ft.first_child[x + 1]))
    if pl < 0:
        out = [1]
    elif (pl > 0) if variant == 0 else (pl == me):
        below = [_chain(ft, c, me, variant, memo) for c in kids if ft.player[c] >= 0]
        depth = max([len(b) for b in below], default=0)
        out = [1] + [sum(b[d] for b in below if d < len(b)) for d in range(depth)]
    else:
        below = [_chain(ft, c, me, variant, memo) for c in kids]
        out = [max(b[d] for b in below if d < len(b)) for d in range(max(len(b) for b in below))]
    memo[key] = out
    return out


def _simulate(ft, me, variant, rng):
    """One sampled traversal: number of records per explored level."""
    counts = {}

    def step(x, level):
        while True:
            pl = int(ft.player[x])
            if pl < 0 or ((pl > 0) if variant == 0 else (pl == me)):
                break
            x = int(rng.integers(ft.first_child[x], ft.first_child[x + 1]))
        counts[level] = counts.get(level, 0) + 1
        if ft.player[x] < 0:
            return
        for c in range(int(ft.first_child[x]), int(ft.first_child[x + 1])):
            if ft.player[c] >= 0:
                step(c, level + 1)

    step(0, 0)
    return counts


@pytest.mark.parametrize('name', ['leduc2', 'leduc3', 'ocp4', 'rps'])
@pytest.mark.parametrize('variant', [0, 1])
def test_record_capacities(name, variant):
    from rlpoker_b200.flatten import flatten
    from rlpoker_b200.solver import plan_lane_records
    ft = flatten(make_game(name))
    got = plan_lane_records(ft, variant)
    want = np.zeros(33, np.int64)
    for me in (1, 2):
        rows = _chain(ft, 0, me, variant, {})
        want[:len(rows)] += rows
    assert np.array_equal(got, want)
    assert got[0] == 2                                  # one root record per traversal
    rng = np.random.default_rng(5)
    for me in (1, 2):
        bound = _chain(ft, 0, me, variant, {})
        for _ in range(200):
            for level, n in _simulate(ft, me, variant, rng).items():
                assert n <= bound[level]
