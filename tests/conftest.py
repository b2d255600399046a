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


def pytest_sessionstart(session):
    """Build the native pieces when they are missing (a fresh checkout): the CUDA library cross-compiles without a
    GPU, the oracle needs gcc only.  Up-to-date builds are left alone."""
    import subprocess
    lib = os.path.join(ROOT, 'rlpoker_b200', 'librlpoker_b200.so')
    if not os.path.exists(lib):
        subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'rlpoker_b200', 'csrc')])
    if not os.path.exists(os.path.join(ROOT, 'oracle', '_build', 'libcfr_oracle.so')):
        subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'oracle')])


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


def random_game(seed, depth=6, zero_sum=False, max_act=5, classes=None):
    """A random extensive game with perfect recall: the acting player, #actions and hidden_from depend only on
    the depth, so nodes of one information set agree on their actions; where an action is visible to both
    players its first child may be a leaf (so leaves sit at several depths while every set of nodes a player
    cannot tell apart is all-terminal or all-internal, which the reference's br() assumes); random chance
    probabilities, float payoffs.  `classes` = (ExtensiveGame, ExtensiveGameNode) of another implementation
    (oracle/gen_golden_random.py passes the reference's to build the identical tree there)."""
    if classes is None:
        from rlpoker_b200.extensive_game import ExtensiveGame, ExtensiveGameNode
    else:
        ExtensiveGame, ExtensiveGameNode = classes
    rng = np.random.RandomState(seed)
    who = [int(rng.randint(0, 3)) for _ in range(depth)]
    who[0] = 0 if seed % 2 == 0 else 1
    nact = [int(rng.randint(2, max_act + 1)) for _ in range(depth)]
    hidden = [[p for p in (1, 2) if rng.rand() < 0.4 and p != who[d]] for d in range(depth)]   # perfect recall
    early_leaf = [d >= 1 and not hidden[d] and rng.rand() < 0.7 for d in range(depth)]

    def build(d, path, leaf):
        if d == depth or leaf:
            u1 = float(np.round(rng.standard_normal() * 3, 3))
            u2 = -u1 if zero_sum else float(np.round(rng.standard_normal() * 3, 3))
            return ExtensiveGameNode(-1, action_list=path, utility={1: u1, 2: u2})
        node = ExtensiveGameNode(who[d], action_list=path, hidden_from=list(hidden[d]))
        acts = ['a%d_%d' % (d, k) for k in range(nact[d])]
        if who[d] == 0:
            p = rng.rand(len(acts)) + 0.1
            p /= p.sum()
            node.chance_probs = dict(zip(acts, p.tolist()))
        for k, a in enumerate(acts):
            node.children[a] = build(d + 1, path + (a,), early_leaf[d] and k == 0)
        return node

    return ExtensiveGame(build(0, (), False))
