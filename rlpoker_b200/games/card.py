"""Playing cards (reference: rlpoker/games/card.py:3-7)."""
from collections import namedtuple

Card = namedtuple('Card', ['value', 'suit'])


def get_deck(num_values, num_suits):
    """All (value, suit) cards, value-major."""
    return [Card(v, s) for v in range(num_values) for s in range(num_suits)]
