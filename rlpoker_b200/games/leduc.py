"""Leduc hold'em with an arbitrary deck (reference: rlpoker/games/leduc.py:16-288, 504-510).

Rules as implemented by the reference builder: both players ante 1; round 1
(player 1 opens) then a board card then round 2 (player 2 opens); a bet puts
``raise_amount`` (doubled in round 2) into the bettor's pot, calling a bet puts
in the same amount; facing a bet the options are call / fold / bet, in that
order; the ``max_raises``-th bet of a round ends the round unanswered; pair
beats high card.

The tree is built depth-first here; the breadth-first node numbering the solver
relies on is produced by ``rlpoker_b200.flatten`` from ``children`` insertion
order, which is what this module has to get (and gets) identical to the
reference: chance children in deck order of the remaining cards, betting
children in the order call(1), fold(0), bet(2).
"""
from collections import Counter

from rlpoker_b200.extensive_game import ExtensiveGame, ExtensiveGameNode
from rlpoker_b200.games.card import Card, get_deck
from rlpoker_b200.games.util import ExtensiveGameBuilder, parse_params

CALL, FOLD, BET = 1, 0, 2


def _without_first(cards, card):
    i = cards.index(card)
    return cards[:i] + cards[i + 1:]


class Leduc(ExtensiveGame):

    def __init__(self, cards, max_raises=4, raise_amount=2):
        assert len(set(cards)) == len(cards), "cards must be distinct"
        self.cards = cards
        self.max_raises = max_raises
        self.raise_amount = raise_amount
        super().__init__(Leduc.create_tree(cards, max_raises=max_raises, raise_amount=raise_amount))

    @classmethod
    def create_game(cls, num_values, num_suits=2, native=None, **kwargs):
        """``Leduc.create_game(n)`` of the reference README (README.md:6): n ranks x 2 suits.

        ``native=True`` (default from 8 ranks up, where node objects cost gigabytes) returns a
        ``rlpoker_b200.games.leduc_native.FlatGame``: the same tree as flat arrays straight from the native generator,
        accepted by every solver entry point but without ``root`` / node objects."""
        if native is None:
            native = num_values >= 8
        if native:
            from rlpoker_b200.games.leduc_native import create_leduc
            return create_leduc(num_values, num_suits, **kwargs)
        return cls(get_deck(num_values, num_suits), **kwargs)

    # ------------------------------------------------------------------ tree
    @staticmethod
    def create_tree(cards, max_raises, raise_amount):
        stakes = {1: raise_amount, 2: 2 * raise_amount}

        def node(player, path, pot, hidden_from=None):
            n = ExtensiveGameNode(player, action_list=path, hidden_from=hidden_from)
            n.extra_info['pot'] = pot
            return n

        def deal(parent, remaining, make_child):
            for card, count in Counter(remaining).items():
                parent.chance_probs[card] = count / len(remaining)
                parent.children[card] = make_child(card, _without_first(remaining, card))

        def end_of_round(rnd, path, pot, remaining, holes):
            if rnd == 1:
                board_node = node(0, path, pot)
                deal(board_node, remaining,
                     lambda c, rest: betting(2, 2, (), path + (c,), dict(pot), rest, holes + (c,)))
                return board_node
            leaf = node(-1, path, pot)
            leaf.utility = Leduc.compute_showdown(holes[0], holes[1], holes[2], pot)
            return leaf

        def betting(rnd, player, history, path, pot, remaining, holes):
            """The decision node of ``player`` after ``history`` in round ``rnd``."""
            this = node(player, path, pot)
            other = 3 - player
            raised = dict(pot)
            raised[player] += stakes[rnd]
            if not history:
                this.children[CALL] = betting(rnd, other, (CALL,), path + (CALL,), dict(pot), remaining, holes)
                this.children[BET] = betting(rnd, other, (BET,), path + (BET,), raised, remaining, holes)
            elif history[-1] == CALL:
                this.children[CALL] = end_of_round(rnd, path + (CALL,), dict(pot), remaining, holes)
                this.children[BET] = betting(rnd, other, history + (BET,), path + (BET,), raised, remaining, holes)
            else:
                this.children[CALL] = end_of_round(rnd, path + (CALL,), dict(raised), remaining, holes)
                folded = node(-1, path + (FOLD,), dict(pot))
                lost = pot[player]
                folded.utility = {1: -lost, 2: lost} if player == 1 else {1: lost, 2: -lost}
                this.children[FOLD] = folded
                if history.count(BET) < max_raises - 1:
                    this.children[BET] = betting(rnd, other, history + (BET,), path + (BET,), raised, remaining, holes)
                else:
                    this.children[BET] = end_of_round(rnd, path + (BET,), dict(raised), remaining, holes)
            return this

        ante = {1: 1, 2: 1}
        root = node(0, (), ante, hidden_from=[2])

        def second_hole(c1, rest1):
            n = node(0, (c1,), ante, hidden_from=[1])
            deal(n, rest1, lambda c2, rest2: betting(1, 1, (), (c1, c2), ante, rest2, (c1, c2)))
            return n

        deal(root, tuple(cards), second_hole)
        return root

    @staticmethod
    def compute_showdown(hole1, hole2, board, pot):
        """{1: reward, 2: reward}: pair with the board wins, else the higher rank, else a draw."""
        r1, r2, rb = hole1.value, hole2.value, board.value
        strength1 = (r1 == rb, r1)
        strength2 = (r2 == rb, r2)
        if strength1 == strength2 or (r1 == rb and r2 == rb):
            return {1: 0, 2: 0}
        if strength1 > strength2:
            return {1: pot[2], 2: -pot[2]}
        return {1: -pot[1], 2: pot[1]}

    def get_node_from_actions(self, action_list):
        node = self.root
        for action in action_list:
            node = node.children[action]
        return node


class LeducBuilder(ExtensiveGameBuilder):
    @staticmethod
    def build(spec):
        params = parse_params(spec)
        return Leduc(get_deck(num_values=int(params['values']), num_suits=int(params['suits'])))
