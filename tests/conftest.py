import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden():
    with open(os.path.join(GOLDEN_DIR, 'reference_golden.json')) as f:
        return json.load(f)


@pytest.fixture(scope='session')
def golden_arrays():
    return dict(np.load(os.path.join(GOLDEN_DIR, 'reference_arrays.npz')))


_GAMES = {}


def make_game(name):
    """Games built by THIS repo's builders (cached)."""
    if name not in _GAMES:
        from rlpoker_b200.games.leduc import Leduc
        from rlpoker_b200.games.card import get_deck
        from rlpoker_b200.games.one_card_poker import OneCardPoker
        from rlpoker_b200.games.rock_paper_scissors import create_rock_paper_scissors
        if name == 'rps':
            _GAMES[name] = create_rock_paper_scissors()
        elif name.startswith('ocp'):
            _GAMES[name] = OneCardPoker.create_game(int(name[3:]))
        elif name.startswith('leduc'):
            _GAMES[name] = Leduc(get_deck(int(name[5:]), 2))
        else:
            raise KeyError(name)
    return _GAMES[name]


_OTREES = {}


def oracle_tree(name):
    if name not in _OTREES:
        from oracle import flat
        _OTREES[name] = flat.flatten_game(make_game(name))
    return _OTREES[name]


@pytest.fixture(scope='session')
def games():
    return make_game


@pytest.fixture(scope='session')
def otrees():
    return oracle_tree
