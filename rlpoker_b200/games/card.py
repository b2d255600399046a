"""Playing cards.

A card is the ``(value, suit)`` named tuple of the reference (rlpoker/games/card.py:3), so cards hash, compare and
print exactly like the reference's and can be used as action labels / information-set key elements interchangeably.
Decks are value-major: ``get_deck(n, s)[value * s + suit]`` -- the native Leduc generator (csrc/leduc_gen.cu) numbers
cards the same way.
"""
import collections

Card = collections.namedtuple('Card', ['value', 'suit'])


def get_deck(num_values, num_suits):
    """The ``num_values * num_suits`` distinct cards, lowest value first, suits innermost."""
    deck = []
    for value in range(num_values):
        deck.extend(Card(value, suit) for suit in range(num_suits))
    return deck


def card_index(card, num_suits):
    """Position of ``card`` in ``get_deck(*, num_suits)``."""
    return card.value * num_suits + card.suit
