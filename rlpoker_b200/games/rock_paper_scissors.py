"""Rock-paper-scissors as a two-level tree whose root is a player node
(reference: rlpoker/games/rock_paper_scissors.py:33-66).  Player 2 does not see player 1's move."""
from rlpoker_b200 import extensive_game
from rlpoker_b200.games.util import ExtensiveGameBuilder

MOVES = ('R', 'P', 'S')
_BEATS = {('R', 'S'), ('P', 'R'), ('S', 'P')}


def create_rock_paper_scissors():
    Node = extensive_game.ExtensiveGameNode
    root = Node(1, action_list=(), hidden_from={2})
    for first in MOVES:
        reply = Node(2, action_list=(first,), hidden_from={1})
        for second in MOVES:
            u1 = 0 if first == second else (1 if (first, second) in _BEATS else -1)
            reply.children[second] = Node(-1, action_list=(first, second), utility={1: u1, 2: -u1})
        root.children[first] = reply
    return extensive_game.ExtensiveGame(root)


class RockPaperScissorsBuilder(ExtensiveGameBuilder):
    @staticmethod
    def build(spec):
        return create_rock_paper_scissors()
