"""One-card poker (http://www.cs.cmu.edu/~ggordon/poker/; reference: rlpoker/games/one_card_poker.py:9-128).

Each player antes 1 and is dealt one private card from ``n_cards`` distinct
cards.  Player 1 bets 0 or 1; player 2 answers 0 or 1; after (0, 1) player 1
answers once more.  Answering 0 to a 1 is a fold.  Utilities are floats.
"""
from rlpoker_b200 import extensive_game
from rlpoker_b200.games.util import ExtensiveGameBuilder, parse_params

Node = extensive_game.ExtensiveGameNode


class OneCardPoker(extensive_game.ExtensiveGame):

    @staticmethod
    def compute_utility(betting_actions, hole_cards):
        staked = {1: 1.0, 2: 1.0}
        for turn, amount in enumerate(betting_actions):
            staked[turn % 2 + 1] += amount
        if list(betting_actions[-2:]) == [1, 0]:
            winner = 1 if len(betting_actions) % 2 == 0 else 2       # the bettor wins a fold
        else:
            winner = 1 if hole_cards[1] > hole_cards[2] else 2
        loser = 3 - winner
        return {winner: staked[loser], loser: -staked[loser]}

    @staticmethod
    def create_one_card_tree(action_list, cards):
        """Subtree below ``action_list`` (cards first, then bets); the root call returns the game."""
        depth = len(action_list)
        if depth == 0:
            root = Node(0)
            root.hidden_from = [2]
            for card in cards:
                root.children[card] = OneCardPoker.create_one_card_tree([card], cards)
                root.chance_probs[card] = 1.0 / len(cards)
            return extensive_game.ExtensiveGame(root)
        if depth == 1:
            node = Node(0)
            node.hidden_from = [1]
            for card in cards:
                if card != action_list[0]:
                    node.children[card] = OneCardPoker.create_one_card_tree(action_list + [card], cards)
                    node.chance_probs[card] = 1.0 / (len(cards) - 1.0)
            return node
        bets = action_list[2:]
        over = (len(bets) == 2 and (bets[1] == 0 or bets[0] == bets[1])) or len(bets) == 3
        if over:
            node = Node(-1)
            node.utility = OneCardPoker.compute_utility(bets, {1: action_list[0], 2: action_list[1]})
            return node
        assert len(bets) < 3
        node = Node(1 if len(bets) % 2 == 0 else 2)
        for bet in (0, 1):
            node.children[bet] = OneCardPoker.create_one_card_tree(action_list + [bet], cards)
        return node

    @staticmethod
    def create_game(n_cards):
        return OneCardPoker.create_one_card_tree([], range(1, n_cards + 1))


class OneCardPokerBuilder(ExtensiveGameBuilder):
    @staticmethod
    def build(spec):
        return OneCardPoker.create_game(int(parse_params(spec)['values']))
