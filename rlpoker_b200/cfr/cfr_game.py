"""Node accessors kept for API compatibility (reference: rlpoker/cfr/cfr_game.py:8-51)."""
from rlpoker_b200.util import sample_action


def payoffs(node):
    return node.utility


def is_terminal(node):
    return node.player == -1


def which_player(node):
    return node.player


def get_available_actions(node):
    return list(node.children)


def sample_chance_action(node):
    if node.player != 0:
        raise ValueError("Node must be a chance node, but was player {}".format(node.player))
    return sample_action(node.chance_probs)


def get_information_set(game, node):
    return game.info_set_ids[node]
