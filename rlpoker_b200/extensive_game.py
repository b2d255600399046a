"""Host-side data model for two-player extensive-form games.

This is the Python surface the solvers consume; it mirrors the reference's
``rlpoker/extensive_game.py`` (ActionFloat :10-109, Strategy :112-166,
ExtensiveGameNode :196-250, ExtensiveGame :253-490) so that callers of the
reference can switch packages without touching their code.  Nothing in here is
on the hot path: games are flattened once (``rlpoker_b200.flatten``) into
structure-of-arrays buffers and all per-iteration work happens in CUDA.
"""
import collections
from collections.abc import Mapping

import numpy as np

TERMINAL, CHANCE = -1, 0


class ActionFloat(Mapping):
    """One float per action (a read-only mapping action -> float)."""

    def __init__(self, action_floats):
        self.action_floats = action_floats

    # -- constructors -------------------------------------------------------
    @staticmethod
    def initialise_zero(actions):
        return ActionFloat(dict.fromkeys(actions, 0.0))

    @staticmethod
    def initialise_uniform(actions):
        p = 1.0 / len(actions)
        return ActionFloat(dict.fromkeys(actions, p))

    # -- arithmetic (all return new objects) --------------------------------
    @staticmethod
    def sum(a, b):
        """Elementwise a + b over the union of actions (missing == 0.0)."""
        out = dict(a.items())
        for action, x in b.items():
            out[action] = out.get(action, 0.0) + x
        return ActionFloat(out)

    @staticmethod
    def scalar_multiply(a, scalar):
        return ActionFloat({action: x * scalar for action, x in a.items()})

    @staticmethod
    def normalise(a):
        """a / sum(a); uniform when the total is zero; ValueError on negatives.

        ``sum`` is the Python builtin on purpose: it is Neumaier-compensated on
        CPython >= 3.12, which the device kernels reproduce.
        """
        xs = list(a.values())
        if min(xs) < 0.0:
            raise ValueError("All action floats must be positive.")
        total = sum(xs)
        if total == 0.0:
            return ActionFloat.initialise_uniform(list(a))
        return ActionFloat({action: x / total for action, x in a.items()})

    # -- mapping protocol ---------------------------------------------------
    @property
    def action_list(self):
        return list(self.action_floats)

    def __getitem__(self, action):
        return self.action_floats[action]

    def __iter__(self):
        return iter(self.action_floats)

    def __len__(self):
        return len(self.action_floats)

    def __eq__(self, other):
        return isinstance(other, ActionFloat) and self.action_floats == other.action_floats

    __hash__ = None

    def __repr__(self):
        return "ActionFloat({})".format(self.action_floats)

    __str__ = __repr__

    def copy(self):
        return ActionFloat(dict(self.action_floats))


class Strategy:
    """Information set -> distribution over the actions available there."""

    def __init__(self, strategy):
        self.strategy = strategy

    @staticmethod
    def initialise():
        return Strategy({})

    def set_uniform_action_probs(self, info_set, available_actions):
        assert info_set not in self.strategy, "Info set {} already exists".format(info_set)
        self.strategy[info_set] = ActionFloat.initialise_uniform(available_actions)

    def get_action_probs(self, info_set):
        return self.strategy[info_set]

    def get_info_sets(self):
        return list(self.strategy)

    def keys(self):
        return self.strategy.keys()

    def items(self):
        return self.strategy.items()

    def __getitem__(self, info_set):
        return self.strategy[info_set]

    def __setitem__(self, info_set, probs):
        self.strategy[info_set] = probs

    def __contains__(self, info_set):
        return info_set in self.strategy

    def __iter__(self):
        return iter(self.strategy)

    def __len__(self):
        return len(self.strategy)

    def __eq__(self, other):
        return isinstance(other, Strategy) and self.strategy == other.strategy

    __hash__ = None

    def copy(self):
        return Strategy({k: v.copy() for k, v in self.strategy.items()})

    def __str__(self):
        body = "\n".join("Info set: {}, Action probs: {}".format(k, v) for k, v in self.strategy.items())
        return "Strategy({})".format(body)


def compute_weighted_strategy(strategies):
    """Per information set, the weight-averaged ActionFloat of [(weight, ActionFloat), ...]."""
    out = {}
    for info_set, weighted in strategies.items():
        weights = np.array([w for w, _ in weighted])
        per_action = collections.defaultdict(list)
        for _, probs in weighted:
            for action in probs.action_list:
                per_action[action].append(probs[action])
        out[info_set] = ActionFloat({
            action: np.sum(weights * np.array(ps)) / np.sum(weights) for action, ps in per_action.items()})
    return Strategy(out)


class ExtensiveGameNode:
    """A node of the game tree.

    player: -1 terminal, 0 chance, 1 / 2 the acting player.
    children: action -> node, in insertion order (that order is part of the
        tree layout contract: it fixes child slots and summation order).
    hidden_from: players that do not observe the action taken *at* this node.
    chance_probs: action -> probability (chance nodes); utility: {1: u1, 2: u2} (terminals).
    """

    def __init__(self, player, action_list=(), children=None, hidden_from=None, chance_probs=None, utility=None):
        self.player = player
        self.action_list = action_list
        self.children = {} if children is None else children
        self.hidden_from = set() if hidden_from is None else hidden_from
        self.chance_probs = {} if chance_probs is None else chance_probs
        self.utility = {} if utility is None else utility
        self.extra_info = {}

    @property
    def actions(self):
        return list(self.children)

    def __str__(self):
        return ("Node(\nPlayer: {}\nActions: {}\nHidden from: {}\nUtility: {}\nChance probs: {}\n"
                "Action list: {}\n)").format(self.player, self.actions, self.hidden_from, self.utility,
                                             self.chance_probs, self.action_list)


class ExtensiveGame:
    """A game given by its root node; also owns the node -> information-set-id map.

    An information-set id is the root->node action path in which every action
    taken at a node hidden from the acting player is replaced by ``-1``
    (reference: extensive_game.py:286-335).
    """

    def __init__(self, root):
        self.root = root
        self.info_set_ids = self.build_info_set_ids()

    # -- information sets ---------------------------------------------------
    def build_information_sets(self, player):
        """node -> visible action tuple from ``player``'s point of view, for every node."""
        seen = {}
        stack = [(self.root, ())]
        while stack:
            node, visible = stack.pop()
            seen[node] = visible
            masked = node.hidden_from is not None and player in node.hidden_from
            for action, child in node.children.items():
                stack.append((child, visible + ((-1,) if masked else (action,))))
        return seen

    def build_info_set_ids(self):
        ids = {}
        for player in (1, 2):
            for node, visible in self.build_information_sets(player).items():
                if node.player == player:
                    ids[node] = visible
        return ids

    def get_info_set_id(self, node):
        assert node.player > 0
        return self.info_set_ids[node]

    def get_node(self, actions):
        node = self.root
        for action in actions:
            node = node.children.get(action)
            if node is None:
                return None
        return node

    # -- strategies ---------------------------------------------------------
    def complete_strategy_uniformly(self, strategy):
        """Copy of ``strategy`` with uniform entries for every missing information set."""
        completed = strategy.copy()
        num_missing = 0
        for node, info_set in self.info_set_ids.items():
            if info_set not in completed:
                n = float(len(node.children))
                completed[info_set] = {a: 1.0 / n for a in node.children}
                num_missing += 1
        return completed, num_missing

    def is_strategy_complete(self, strategy):
        for node, info_set in self.info_set_ids.items():
            if info_set not in strategy:
                return False
            if set(strategy[info_set].action_list) != set(node.children):
                return False
        return True

    # -- evaluation ---------------------------------------------------------
    def expected_value(self, strategy_1, strategy_2, num_iters):
        """Monte-Carlo play-outs; returns the list of player-1 utilities (uses numpy's global RNG)."""
        by_player = {1: strategy_1, 2: strategy_2}
        results = []
        for _ in range(num_iters):
            node = self.root
            while node.player != TERMINAL:
                actions = list(node.children)
                probs = [1.0 / len(actions)] * len(actions)
                if node.player == CHANCE:
                    probs = [node.chance_probs[a] for a in actions]
                else:
                    info_set = self.info_set_ids[node]
                    if info_set in by_player[node.player]:
                        probs = [by_player[node.player][info_set][a] for a in actions]
                assert np.isclose(sum(probs), 1.0)
                idx = np.random.choice(len(actions), p=np.array(probs))
                node = node.children[actions[idx]]
            results.append(node.utility[1])
        return results

    def expected_value_exact(self, strategy1, strategy2):
        """Exact (u1, u2) under the two strategies: breadth-first reach propagation,
        terminal contributions added in breadth-first order (extensive_game.py:382-423)."""
        by_player = {1: strategy1, 2: strategy2}
        total = {1: 0.0, 2: 0.0}
        queue = collections.deque([(self.root, 1.0)])
        while queue:
            node, reach = queue.popleft()
            if node.player == TERMINAL:
                total[1] += reach * node.utility[1]
                total[2] += reach * node.utility[2]
                continue
            for action, child in node.children.items():
                if node.player == CHANCE:
                    p = node.chance_probs[action]
                else:
                    p = by_player[node.player][self.info_set_ids[node]][action]
                queue.append((child, reach * p))
        return total[1], total[2]

    # -- debugging ----------------------------------------------------------
    def print_tree(self, only_leaves=False):
        stack = [(self.root, ())]
        while stack:
            node, path = stack.pop()
            if only_leaves:
                if not node.children:
                    print(path, node.utility)
            else:
                print("Node for action list: {}".format(path))
                print(node)
                print("-----")
            for action, child in reversed(list(node.children.items())):
                stack.append((child, path + (action,)))
