"""``'Name:key_val:...'`` -> ExtensiveGame (reference: rlpoker/games/game_builder.py:7-19).

Unlike the reference the game name is matched case-insensitively, so the default
spec of ``examples/cfr.py`` (``'leduc:values_3:suits_2'``) builds."""
from rlpoker_b200.games.leduc import LeducBuilder
from rlpoker_b200.games.one_card_poker import OneCardPokerBuilder
from rlpoker_b200.games.rock_paper_scissors import RockPaperScissorsBuilder

_BUILDERS = {'leduc': LeducBuilder, 'onecardpoker': OneCardPokerBuilder, 'rockpaperscissors': RockPaperScissorsBuilder}


def buildExtensiveGame(spec):
    name, _, game_spec = spec.partition(':')
    try:
        builder = _BUILDERS[name.lower()]
    except KeyError:
        raise ValueError(f"Undefined game: {name}") from None
    return builder.build(game_spec)
