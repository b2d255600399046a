"""rlpoker_b200: a B200-native tabular CFR solver behind the rlpoker Python API.

    from rlpoker_b200.games.leduc import Leduc
    from rlpoker_b200.cfr.cfr import cfr
    from rlpoker_b200.best_response import compute_exploitability

``install_as_rlpoker()`` additionally registers this package under the reference's import names
(``rlpoker.cfr.cfr``, ``rlpoker.best_response``, ``rlpoker.games.leduc`` ...), so code written against
cgnicholls/rlpoker runs unchanged.
"""
import importlib
import sys

__version__ = '0.1.0'

_ALIASES = ['extensive_game', 'best_response', 'experiment', 'util', 'cfr', 'cfr.cfr', 'cfr.external_cfr',
            'cfr.cfr_util', 'cfr.cfr_game', 'cfr.cfr_metrics', 'games', 'games.card', 'games.leduc',
            'games.one_card_poker', 'games.rock_paper_scissors', 'games.game_builder', 'games.util']


def install_as_rlpoker():
    """Make ``import rlpoker...`` resolve to this package (refuses to shadow an already imported rlpoker)."""
    if 'rlpoker' in sys.modules and sys.modules['rlpoker'] is not sys.modules[__name__]:
        raise ImportError("a different 'rlpoker' package is already imported")
    sys.modules['rlpoker'] = sys.modules[__name__]
    for name in _ALIASES:
        sys.modules['rlpoker.' + name] = importlib.import_module(__name__ + '.' + name)
