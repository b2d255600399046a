"""The native Leduc generator (host C++ inside the CUDA library; no GPU needed) against the Python builder
that tests/test_oracle_golden.py pins to the reference."""
import numpy as np
import pytest

from conftest import make_game


@pytest.mark.parametrize('n', [2, 3, 4, 5])
def test_native_equals_python_builder(n):
    from rlpoker_b200.flatten import flatten
    from rlpoker_b200.games.leduc_native import leduc_flat_tree
    a = flatten(make_game('leduc%d' % n))
    b = leduc_flat_tree(n)
    assert a.num_nodes == b.num_nodes and a.num_infosets == b.num_infosets
    for name in ('parent', 'player', 'infoset', 'cprob', 'u1', 'u2', 'hidden', 'first_child', 'level_off', 'act_off',
                 'iset_nodes', 'iset_node_off', 'nact', 'depth', 'slot'):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
    assert [a.labels[i] for i in a.label[1:]] == [b.labels[i] for i in b.label[1:]]
    assert a.info_set_keys == b.info_set_keys
    assert [a.actions_of(i) for i in range(a.num_infosets)] == [b.actions_of(i) for i in range(b.num_infosets)]
    assert a.zero_sum and b.zero_sum


def test_closed_form_sizes():
    """N = 1 + C + 23 D + 207 D B, I = 8 C + 72 D, Q = 2.75 I."""
    from rlpoker_b200.games.leduc_native import leduc_flat_tree
    for n in (3, 7, 13):
        ft = leduc_flat_tree(n)
        c = 2 * n
        d, b = c * (c - 1), c - 2
        assert ft.num_nodes == 1 + c + 23 * d + 207 * d * b
        assert ft.num_infosets == 8 * c + 72 * d
        assert ft.num_pairs * 4 == 11 * ft.num_infosets
        assert ft.num_levels == 14


def test_other_rules():
    from rlpoker_b200.flatten import flatten
    from rlpoker_b200.games.card import get_deck
    from rlpoker_b200.games.leduc import Leduc
    from rlpoker_b200.games.leduc_native import leduc_flat_tree
    a = flatten(Leduc(get_deck(3, 1), max_raises=2, raise_amount=3))
    b = leduc_flat_tree(3, 1, max_raises=2, raise_amount=3)
    for name in ('parent', 'player', 'infoset', 'cprob', 'u1', 'u2', 'hidden'):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name


def test_create_game_switches_to_the_native_generator():
    from rlpoker_b200.games.leduc import Leduc
    from rlpoker_b200.games.leduc_native import FlatGame
    small = Leduc.create_game(2)
    assert isinstance(small, Leduc) and small.root.player == 0
    big = Leduc.create_game(8)
    assert isinstance(big, FlatGame) and big.flat.num_nodes == 1 + 16 + 23 * 240 + 207 * 240 * 14
    forced = Leduc.create_game(3, native=True)
    assert isinstance(forced, FlatGame) and forced.flat.num_infosets == 2208
