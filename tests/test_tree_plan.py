"""Host-only coverage of the device layout derivation (rlp_tree_plan: validation, two-tier tiling, shape classes,
NVRTC compilation of the generated straight-line kernels).  No GPU needed."""
import numpy as np
import pytest

from conftest import make_game, random_game


def _plan(ft):
    from rlpoker_b200.solver import plan_tree
    return plan_tree(ft)


@pytest.mark.parametrize('n', [2, 3, 5, 13])
def test_leduc_tiling_closed_forms(n):
    """Leduc-n: one upper tile per (hole1, hole2) deal, one micro subtree per (deal, round-1 ending, board card),
    every micro subtree the 23-node betting round, one shape, both generated kernels compile."""
    from rlpoker_b200.games.leduc_native import leduc_flat_tree
    ft = leduc_flat_tree(n)
    c = 2 * n
    deals, boards = c * (c - 1), c - 2
    p = _plan(ft)
    assert p['tiled'] == 1 and p['zero_sum'] == 1
    assert p['upper_tiles'] == deals
    assert p['micro_subtrees'] == 9 * deals * boards
    assert (p['micro_shapes'], p['micro_nodes'], p['micro_decisions'], p['micro_leaves']) == (1, 23, 8, 15)
    assert p['tile_nodes'] == 23 + 9 * boards and p['tile_nonterminals'] == 17
    assert p['infosets_upper'] == 8 * c and p['infosets_micro_only'] == ft.num_infosets - 8 * c
    assert p['largest_infoset'] == c - 1
    assert p['kernels_failed'] == 0 and p['kernels_compiled'] == 3        # micro_k, upper_k0/1, cfr_fused_k


def test_small_games_and_fallbacks():
    from rlpoker_b200.flatten import flatten
    p = _plan(flatten(make_game('rps')))
    assert (p['tiled'], p['upper_tiles'], p['micro_subtrees'], p['micro_nodes']) == (1, 0, 1, 13)
    p = _plan(flatten(make_game('ocp4')))
    assert p['tiled'] == 1 and p['micro_subtrees'] == 12 and p['upper_tiles'] == 0 and p['kernels_failed'] == 0
    for seed in range(6):
        ft = flatten(random_game(seed, zero_sum=seed % 2 == 0))
        p = _plan(ft)                                   # any outcome is fine as long as planning succeeds
        assert p['zero_sum'] == seed % 2 == 0 or p['zero_sum'] in (0, 1)
        assert p['kernels_failed'] == 0


def test_plan_rejects_malformed_trees():
    from rlpoker_b200._lib import RlpError
    from rlpoker_b200.flatten import flatten
    import copy
    ft = copy.copy(flatten(make_game('leduc2')))
    bad = copy.copy(ft)
    bad.player = ft.player.copy()
    bad.player[int(np.flatnonzero(ft.player == -1)[0])] = 1           # a decision node without children
    with pytest.raises(RlpError):
        _plan(bad)
    bad = copy.copy(ft)
    bad.infoset = ft.infoset.copy()
    dec = np.flatnonzero(ft.player > 0)
    bad.infoset[dec[0]] = ft.infoset[dec[-1]]                          # information set spanning depths
    with pytest.raises(RlpError):
        _plan(bad)


def test_fused_kernel_source_is_generated_and_compiles():
    """The group-fused persistent kernel: generated for every Leduc-n (one micro shape, one upper shape, complete
    groups) and compiled by NVRTC on the host; not generated for a deal shard (incomplete groups) nor for trees without
    the structure."""
    import ctypes as C
    from conftest import make_game
    from rlpoker_b200 import _lib
    from rlpoker_b200.flatten import flatten
    from rlpoker_b200.games.leduc_native import leduc_flat_tree
    from rlpoker_b200.sharding import shard_flat_tree
    from rlpoker_b200.solver import _tree_desc

    def source_of(ft):
        desc, keep = _tree_desc(ft)
        stats, ln = np.zeros(16, np.int64), C.c_int64()
        _lib.check(_lib.load().rlp_tree_plan_fused_source(C.byref(desc), _lib.ptr(stats, _lib._i64p), None, 0, C.byref(ln)))
        buf = C.create_string_buffer(ln.value)
        _lib.check(_lib.load().rlp_tree_plan_fused_source(C.byref(desc), _lib.ptr(stats, _lib._i64p), buf, ln.value, C.byref(ln)))
        return buf.value.decode(), stats

    for n in (2, 4, 18):                    # 18 ranks: groups of 34 subtrees (wider than a warp)
        src, stats = source_of(leduc_flat_tree(n))
        assert 'cfr_fused_k' in src and 'grid_barrier_x' in src and '__grid_constant__' in src
        assert stats[12] == 3 and stats[13] == 0
    src, stats = source_of(shard_flat_tree(leduc_flat_tree(4), 0, 2))
    assert src == '' and stats[12] == 2
    src, _ = source_of(flatten(make_game('ocp4')))
    assert 'cfr_fused_k' not in src or src == ''
