"""ORACLE (test infrastructure): an independent breadth-first flattening of any object shaped
like the reference's ExtensiveGame (rlpoker/extensive_game.py:196-335) into the plain arrays
``cfr_oracle.c`` works on.  Deliberately naive Python loops -- it is the checker for
``rlpoker_b200.flatten`` and for the native Leduc generator, not a product path."""
import collections

import numpy as np


class OracleTree:
    pass


def flatten_game(game):
    root = game.root
    order, parent, label_of = [root], [-1], [None]
    queue = collections.deque([0])
    while queue:
        i = queue.popleft()
        for action, child in order[i].children.items():
            queue.append(len(order))
            order.append(child)
            parent.append(i)
            label_of.append(action)
    n = len(order)
    index = {id(nd): i for i, nd in enumerate(order)}
    t = OracleTree()
    t.num_nodes = n
    t.parent = np.array(parent, np.int32)
    t.player = np.array([nd.player for nd in order], np.int8)
    first_child = np.zeros(n + 1, np.int32)
    for i, nd in enumerate(order):
        kids = list(nd.children.values())
        first_child[i] = index[id(kids[0])] if kids else -1
    # children are contiguous: fill gaps (terminals) from the right
    nxt = n
    first_child[n] = n
    for i in range(n - 1, -1, -1):
        if first_child[i] < 0:
            first_child[i] = nxt
        nxt = first_child[i]
    for i, nd in enumerate(order):
        assert first_child[i + 1] - first_child[i] == len(nd.children)
    t.first_child = first_child
    labels, label_idx = [], {}
    t.label = np.zeros(n, np.int32)
    for i in range(1, n):
        a = label_of[i]
        if a not in label_idx:
            label_idx[a] = len(labels)
            labels.append(a)
        t.label[i] = label_idx[a]
    t.labels = labels
    t.hidden = np.zeros(n, np.uint8)
    t.cprob = np.ones(n, np.float64)
    t.u1 = np.zeros(n, np.float64)
    t.u2 = np.zeros(n, np.float64)
    t.infoset = np.full(n, -1, np.int32)
    keys, key_idx = [], {}
    for i, nd in enumerate(order):
        hf = nd.hidden_from or ()
        t.hidden[i] = (1 if 1 in hf else 0) + (2 if 2 in hf else 0)
        if nd.player == -1:
            t.u1[i], t.u2[i] = nd.utility[1], nd.utility[2]
        elif nd.player == 0:
            for action, child in nd.children.items():
                t.cprob[index[id(child)]] = nd.chance_probs[action]
        else:
            key = game.info_set_ids[nd]
            if key not in key_idx:
                key_idx[key] = len(keys)
                keys.append(key)
            t.infoset[i] = key_idx[key]
    t.keys = keys
    return finish(t)


def finish(t):
    """Derive depth / level / information-set tables from the primary arrays."""
    n = t.num_nodes
    depth = np.zeros(n, np.int32)
    for i in range(1, n):
        depth[i] = depth[t.parent[i]] + 1
    assert np.all(np.diff(depth) >= 0)
    t.depth = depth
    nlev = int(depth.max()) + 1
    t.level_off = np.searchsorted(depth, np.arange(nlev + 1)).astype(np.int64)
    ni = int(t.infoset.max()) + 1 if n else 0
    t.num_infosets = ni
    members = [[] for _ in range(ni)]
    for i in range(n):
        if t.infoset[i] >= 0:
            members[t.infoset[i]].append(i)
    t.iset_node_off = np.zeros(ni + 1, np.int64)
    t.iset_node_off[1:] = np.cumsum([len(m) for m in members])
    t.iset_nodes = np.array([x for m in members for x in m], np.int32)
    t.nact = np.array([t.first_child[m[0] + 1] - t.first_child[m[0]] for m in members], np.int32)
    t.act_off = np.zeros(ni + 1, np.int64)
    t.act_off[1:] = np.cumsum(t.nact)
    t.iset_player = np.array([t.player[m[0]] for m in members], np.int8)
    return t


def from_arrays(num_nodes, parent, first_child, player, infoset, label, hidden, cprob, u1, u2, labels=None,
                nact=None):
    """Wrap arrays produced elsewhere (e.g. the product's Leduc generator, once it has been
    checked against flatten_game on small decks) so the C oracle can run on large trees."""
    t = OracleTree()
    t.num_nodes = int(num_nodes)
    t.parent = np.ascontiguousarray(parent, np.int32)
    t.first_child = np.ascontiguousarray(first_child, np.int32)
    t.player = np.ascontiguousarray(player, np.int8)
    t.infoset = np.ascontiguousarray(infoset, np.int32)
    t.label = np.ascontiguousarray(label, np.int32)
    t.hidden = np.ascontiguousarray(hidden, np.uint8)
    t.cprob = np.ascontiguousarray(cprob, np.float64)
    t.u1 = np.ascontiguousarray(u1, np.float64)
    t.u2 = np.ascontiguousarray(u2, np.float64)
    t.labels = labels
    t.keys = None
    # vectorised version of finish() for big trees
    n = t.num_nodes
    depth = np.zeros(n, np.int32)
    lo, hi = 0, 1
    level_off = [0, 1]
    while hi < n:
        nhi = int(t.first_child[hi])
        depth[hi:nhi] = len(level_off) - 1
        level_off.append(nhi)
        lo, hi = hi, nhi
    t.depth = depth
    t.level_off = np.array(level_off, np.int64)
    dec = np.flatnonzero(t.infoset >= 0)
    ids = t.infoset[dec]
    ni = (int(ids.max()) + 1 if len(dec) else 0) if nact is None else len(nact)
    t.num_infosets = ni
    o = np.argsort(ids, kind='stable')
    t.iset_nodes = dec[o].astype(np.int32)
    t.iset_node_off = np.zeros(ni + 1, np.int64)
    t.iset_node_off[1:] = np.cumsum(np.bincount(ids, minlength=ni))
    if nact is None:
        rep = t.iset_nodes[t.iset_node_off[:-1]]
        t.nact = (t.first_child[rep + 1] - t.first_child[rep]).astype(np.int32)
        t.iset_player = t.player[rep].astype(np.int8)
    else:                       # a shard: information sets keep the full game's numbering, some are empty
        t.nact = np.ascontiguousarray(nact, np.int32)
        t.iset_player = np.zeros(ni, np.int8)
        t.iset_player[ids] = t.player[dec]
    t.act_off = np.zeros(ni + 1, np.int64)
    t.act_off[1:] = np.cumsum(t.nact)
    return t


def strategy_to_sigma(t, strategy):
    """dict-like strategy -> (sigma[Q] completed uniformly, present[I])."""
    q = int(t.act_off[-1])
    sigma = np.zeros(q)
    present = np.zeros(t.num_infosets, np.uint8)
    for i, key in enumerate(t.keys):
        lo, k = int(t.act_off[i]), int(t.nact[i])
        h = int(t.iset_nodes[t.iset_node_off[i]])
        acts = [t.labels[t.label[j]] for j in range(t.first_child[h], t.first_child[h + 1])]
        if key in strategy:
            sigma[lo:lo + k] = [strategy[key][a] for a in acts]
            present[i] = 1
        else:
            sigma[lo:lo + k] = 1.0 / float(k)
    return sigma, present


def postorder_infoset_order(t):
    """Information sets by first appearance in depth-first post-order of the decision nodes
    (the dict order of cfr_metrics.compute_immediate_regret's per-step sums)."""
    seen = np.zeros(t.num_infosets, bool)
    out = []
    stack = [(0, False)]
    while stack:
        node, done = stack.pop()
        if done:
            i = t.infoset[node]
            if i >= 0 and not seen[i]:
                seen[i] = True
                out.append(i)
            continue
        stack.append((node, True))
        for j in range(t.first_child[node + 1] - 1, t.first_child[node] - 1, -1):
            stack.append((j, False))
    return np.array(out, np.int32)
