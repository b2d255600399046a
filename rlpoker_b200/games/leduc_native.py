"""Leduc-n straight into flat arrays via the native generator (rlp_leduc_generate), for decks where
building Python node objects is impractical (the reference builder, rlpoker/games/leduc.py:44-234, needs
~5 GB / 75 s at 13 ranks).  The arrays are identical to ``flatten(Leduc(get_deck(n, suits)))``
(tests/test_leduc_native.py checks that for small decks)."""
import ctypes as C

import numpy as np

from rlpoker_b200 import _lib
from rlpoker_b200.flatten import FlatTree
from rlpoker_b200.games.card import get_deck

CARD_LABEL = 16


class FlatGame:
    """A game that exists only as a FlatTree (no node objects).  Accepted wherever the solvers take a game;
    ``info_set_ids`` style keys are rebuilt lazily from the arrays when a Strategy dict is requested."""

    def __init__(self, flat_tree, description=''):
        self._flat_tree = flat_tree
        self.description = description

    @property
    def flat(self):
        return self._flat_tree

    def __repr__(self):
        return 'FlatGame(%s, %d nodes)' % (self.description, self._flat_tree.num_nodes)


def leduc_flat_tree(num_values, num_suits=2, max_raises=4, raise_amount=2):
    lib = _lib.load()
    n = C.c_int64(0)
    ni = C.c_int32(0)
    none = None
    _lib.check(lib.rlp_leduc_generate(num_values, num_suits, max_raises, raise_amount, C.byref(n), C.byref(ni),
                                      none, none, none, none, none, none, none, none))
    N = n.value
    ft = FlatTree()
    ft.num_nodes = N
    ft.parent = np.empty(N, np.int32)
    ft.player = np.empty(N, np.int8)
    ft.infoset = np.empty(N, np.int32)
    ft.cprob = np.empty(N, np.float64)
    ft.u1 = np.empty(N, np.float64)
    ft.u2 = np.empty(N, np.float64)
    ft.label = np.empty(N, np.int32)
    ft.hidden = np.empty(N, np.uint8)
    p = _lib.ptr
    _lib.check(lib.rlp_leduc_generate(
        num_values, num_suits, max_raises, raise_amount, C.byref(n), C.byref(ni),
        p(ft.parent, _lib._i32p), p(ft.player, _lib._i8p), p(ft.infoset, _lib._i32p), p(ft.cprob, _lib._dp),
        p(ft.u1, _lib._dp), p(ft.u2, _lib._dp), p(ft.label, _lib._i32p), p(ft.hidden, _lib._u8p)))
    deck = get_deck(num_values, num_suits)
    ft.labels = [0, 1, 2] + [None] * (CARD_LABEL - 3) + deck
    ft._keys = None
    ft._key_fn = lambda: [ft.info_set_key_of(i) for i in range(ft.num_infosets)]
    ft.finalize()
    assert ft.num_infosets == ni.value
    return ft


def create_leduc(num_values, num_suits=2, max_raises=4, raise_amount=2):
    return FlatGame(leduc_flat_tree(num_values, num_suits, max_raises, raise_amount),
                    'Leduc values=%d suits=%d' % (num_values, num_suits))
