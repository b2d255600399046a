"""CUDA path against the second golden set (Leduc with other betting rules / suit counts, other OneCardPoker sizes;
tests/golden/reference_variants.*): regrets and average strategy bit for bit, best responses to 1e-12.

The same trees are covered on the CPU by tests/test_oracle_variants.py (layout, oracle vs reference, device layout
planning + NVRTC compilation)."""
import numpy as np
import pytest

from test_oracle_variants import build, variant_arrays, variants  # noqa: F401  (fixtures)
from oracle.gen_golden_variants import VARIANTS

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name', sorted(VARIANTS))
def test_variant_parity(variants, variant_arrays, name):  # noqa: F811
    from rlpoker_b200.flatten import flatten
    from rlpoker_b200.solver import DeviceCFR, DeviceTree
    want = variants['games'][name]
    game, t = build(name)
    dt = DeviceTree(flatten(game))
    uniform = np.repeat(1.0 / t.nact, t.nact)
    for p in (1, 2):
        assert abs(dt.best_response(uniform, p) - want['uniform_br'][p - 1]) < 1e-12
    s = DeviceCFR(dt)
    s.vanilla(want['cfr']['T'])
    assert np.array_equal(s.regrets, variant_arrays[name + '_R'])
    assert np.array_equal(s.average, variant_arrays[name + '_avg'])
    assert abs(s.exploitability()[0] - want['cfr']['exploitability']) < 1e-12
    assert s.nodes_touched == want['cfr']['nodes_touched']
