"""CPU-side checks: the C ABI library loads and exports every declared symbol, fails loudly without a GPU,
and the host data model behaves like the reference's (restating rlpoker/tests/test_extensive_game.py and
rlpoker/cfr/tests/test_cfr.py)."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT, make_game


def test_library_exports_every_declared_symbol():
    from rlpoker_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'rlpoker_b200.h')).read()
    declared = set(re.findall(r'\b(rlp_[a-z0-9_]+)\s*\(', header))
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert lib.rlp_version() >= 1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from rlpoker_b200._lib import RlpError
    from rlpoker_b200.best_response import compute_exploitability
    from rlpoker_b200.cfr.cfr import cfr
    from rlpoker_b200.extensive_game import Strategy
    with pytest.raises(RlpError):
        cfr(make_game('rps'), num_iters=1, use_chance_sampling=False)
    with pytest.raises(RlpError):
        compute_exploitability(make_game('rps'), Strategy.initialise())


def test_product_never_imports_the_oracle():
    for base, _, files in os.walk(os.path.join(ROOT, 'rlpoker_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.cpp')):
                src = open(os.path.join(base, f)).read()
                assert 'oracle' not in src.replace('# oracle', ''), os.path.join(base, f)


def test_regret_matching_host_kats(golden):
    from rlpoker_b200.cfr.cfr_util import compute_regret_matching, normalise_probs
    from rlpoker_b200.extensive_game import ActionFloat
    for case in golden['regret_matching']:
        out = compute_regret_matching(ActionFloat(dict(enumerate(case['r']))), epsilon=case['eps'],
                                      highest_regret=case['hr'])
        assert list(out.values()) == case['out']
    for case in golden['normalise_probs']:
        assert list(normalise_probs(ActionFloat(dict(enumerate(case['p']))), epsilon=case['eps']).values()) == case['out']
    zero = ActionFloat.initialise_zero(['a', 'b'])
    assert isinstance(zero, ActionFloat) and zero['a'] == 0.0 and set(zero.keys()) == {'a', 'b'}


def test_sampler_golden_sequences(golden):
    from rlpoker_b200.util import sample_action
    from rlpoker_b200.games.card import Card
    g = golden['sample_action']
    np.random.seed(0)
    assert [sample_action({1: 0.2, 2: 0.8}) for _ in range(20)] == g['a']
    np.random.seed(0)
    assert [sample_action({'R': 0.2, 'P': 0.7, 'S': 0.1}) for _ in range(20)] == g['b']
    assert sample_action({Card(3, 4): 0.2, Card(2, 2): 0.8}) in (Card(3, 4), Card(2, 2))
    assert {sample_action({1: 0.2, 2: 0.8}, available_actions=[1]) for _ in range(100)} == {1}


def test_extensive_game_behaviour(golden):
    from rlpoker_b200 import extensive_game as eg
    game = make_game('rps')
    i1, i2 = game.info_set_ids[game.get_node(())], game.info_set_ids[game.get_node(('R',))]
    assert i1 == () and i2 == (-1,)
    s1 = eg.Strategy({i1: eg.ActionFloat({'R': 0.5, 'P': 0.2, 'S': 0.3})})
    s2 = eg.Strategy({i2: eg.ActionFloat({'R': 0.2, 'P': 0.3, 'S': 0.5})})
    assert list(game.expected_value_exact(s1, s2)) == golden['rps_expected_value_exact']
    assert game.get_node(('R', 'P')) is game.root.children['R'].children['P'] and game.get_node(('R', 'P', 'A')) is None
    assert not game.is_strategy_complete(s1)
    both = eg.Strategy({**s1.strategy, **s2.strategy})
    assert game.is_strategy_complete(both)
    c = both.copy()
    c[i1] = eg.ActionFloat({'R': 1.0, 'P': 0.0, 'S': 0.0})
    assert both[i1]['R'] == 0.5 and c != both and both.copy() == both
    a = eg.ActionFloat.sum(eg.ActionFloat({'x': 1.0}), eg.ActionFloat({'x': 2.0, 'y': 3.0}))
    assert dict(a) == {'x': 3.0, 'y': 3.0}
    assert dict(eg.ActionFloat.normalise(eg.ActionFloat({'x': 0.0, 'y': 0.0}))) == {'x': 0.5, 'y': 0.5}
    with pytest.raises(ValueError):
        eg.ActionFloat.normalise(eg.ActionFloat({'x': -1.0, 'y': 2.0}))
    completed, missing = game.complete_strategy_uniformly(s1)
    assert missing == 1 and completed[i2] == {'R': 1 / 3, 'P': 1 / 3, 'S': 1 / 3}
    w = eg.compute_weighted_strategy({i1: [(1.0, eg.ActionFloat({'R': 1.0, 'P': 0.0})), (3.0, eg.ActionFloat({'R': 0.0, 'P': 1.0}))]})
    assert dict(w[i1]) == {'R': 0.25, 'P': 0.75}


def test_leduc_hand_walk_and_showdown():
    """rlpoker/tests/test_leduc.py:10-89 restated."""
    from rlpoker_b200.games.card import Card
    from rlpoker_b200.games.leduc import Leduc
    game = make_game('leduc3')
    cards = [Card(v, s) for v in range(3) for s in range(2)]
    node = game.root
    assert set(node.children) == set(cards)
    node = node.children[Card(0, 0)].children[Card(1, 0)].children[1].children[2].children[1]
    assert set(node.children) == {Card(2, 0), Card(0, 1), Card(1, 1), Card(2, 1)}
    node = node.children[Card(1, 1)]
    assert node.player == 2
    for _ in range(4):
        node = node.children[2]
    assert node.utility == {1: -11, 2: 11}
    pot = {1: 10, 2: 10}
    assert Leduc.compute_showdown(Card(2, 3), Card(2, 4), Card(1, 0), pot) == {1: 0, 2: 0}
    assert Leduc.compute_showdown(Card(2, 3), Card(1, 4), Card(2, 0), pot) == {1: 10, 2: -10}
    assert Leduc.compute_showdown(Card(1, 3), Card(2, 4), Card(2, 0), pot) == {1: -10, 2: 10}
    assert Leduc.compute_showdown(Card(1, 3), Card(0, 4), Card(2, 0), pot) == {1: 10, 2: -10}
    assert Leduc.compute_showdown(Card(1, 3), Card(3, 4), Card(2, 0), pot) == {1: -10, 2: 10}


def test_game_builder_and_cfr_game():
    from rlpoker_b200.cfr import cfr_game
    from rlpoker_b200.games.game_builder import buildExtensiveGame
    assert len(buildExtensiveGame('Leduc:values_3:suits_2').info_set_ids) == 8880
    assert buildExtensiveGame('OneCardPoker:values_4').root.player == 0
    assert buildExtensiveGame('leduc:values_2:suits_2').root.player == 0
    with pytest.raises(ValueError):
        buildExtensiveGame('Chess:values_3')
    game = make_game('rps')
    with pytest.raises(ValueError):
        cfr_game.sample_chance_action(game.root)


def test_install_as_rlpoker_aliases():
    import subprocess
    import sys
    code = ("import rlpoker_b200; rlpoker_b200.install_as_rlpoker();"
            "from rlpoker.cfr import cfr; from rlpoker.games.leduc import Leduc; from rlpoker.games.card import Card;"
            "from rlpoker.best_response import compute_exploitability; import rlpoker.cfr.external_cfr as e;"
            "print(cfr.cfr.__module__, Leduc.create_game.__name__)")
    out = subprocess.check_output([sys.executable, '-c', code], cwd=ROOT, text=True)
    assert out.split() == ['rlpoker_b200.cfr.cfr', 'create_game']


def test_cfr_argument_shapes():
    from rlpoker_b200.cfr.cfr import _unpack
    from rlpoker_b200.experiment import Experiment
    game = make_game('rps')
    b, _ = _unpack((game,), {'num_iters': 10, 'use_chance_sampling': False})
    assert b['game'] is game and b['experiment'] is None and b['num_iters'] == 10 and b['use_chance_sampling'] is False
    exp = Experiment.__new__(Experiment)
    b, _ = _unpack((exp, game, 5), {})
    assert b['experiment'] is exp and b['game'] is game and b['num_iters'] == 5 and b['use_chance_sampling'] is True
    b, extra = _unpack((), {'experiment': None, 'game': game, 'seed': 3})
    assert b['game'] is game and extra == {'seed': 3}
    b, _ = _unpack((None, game), {'linear_weight': True})
    assert b['game'] is game and b['linear_weight'] is True
