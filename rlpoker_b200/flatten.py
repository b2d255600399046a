"""One-time flattening of an ExtensiveGame into depth-ordered structure-of-arrays buffers.

Replaces the tree walking of the reference (rlpoker/extensive_game.py:196-335 and the
recursions in rlpoker/cfr/cfr.py:158-255, rlpoker/best_response.py:9-100) by arrays a
level sweep can stream through.  Layout contract:

* node ids are breadth-first from the root, children in ``node.children`` insertion
  order, so the children of a node are contiguous and every depth is a contiguous range;
* information sets are numbered by first appearance in that order;
* per information set the nodes are listed in ascending id, which equals the order the
  reference's recursion visits them in, so ordered accumulation reproduces its sums.

Everything here is host-side numpy; ``rlpoker_b200.solver`` uploads it through the C ABI.
"""
import collections

import numpy as np


class FlatTree:
    """Breadth-first structure-of-arrays form of a game tree (all numpy, host)."""

    __slots__ = (
        'num_nodes', 'parent', 'slot', 'depth', 'player', 'first_child', 'level_off',
        'infoset', 'num_infosets', 'nact', 'act_off', 'iset_player', 'iset_depth',
        'cprob', 'u1', 'u2', 'zero_sum', 'hidden', 'label', 'labels',
        'iset_node_off', 'iset_nodes', '_keys', '_key_fn',
    )

    # ---- sizes -------------------------------------------------------------
    @property
    def num_levels(self):
        return len(self.level_off) - 1

    @property
    def num_pairs(self):
        """Q: number of (information set, action) pairs."""
        return int(self.act_off[-1])

    @property
    def num_terminals(self):
        return int(np.count_nonzero(self.player == -1))

    @property
    def num_decisions(self):
        return int(np.count_nonzero(self.player > 0))

    # ---- information-set keys (host only; lazily built for generated trees) --
    @property
    def info_set_keys(self):
        if self._keys is None:
            self._keys = self._key_fn()
        return self._keys

    def info_set_key_of(self, iset):
        """Visible action tuple of information set ``iset`` rebuilt from the arrays."""
        node = int(self.iset_nodes[self.iset_node_off[iset]])
        who = int(self.player[node])
        path = []
        while node > 0:
            up = int(self.parent[node])
            path.append(-1 if (int(self.hidden[up]) >> (who - 1)) & 1 else self.labels[int(self.label[node])])
            node = up
        return tuple(reversed(path))

    def actions_of(self, iset):
        """Action labels of an information set in child-slot order."""
        node = int(self.iset_nodes[self.iset_node_off[iset]])
        first = int(self.first_child[node])
        return [self.labels[int(self.label[first + a])] for a in range(int(self.nact[iset]))]

    def finalize(self):
        """Fill every array that is derived from (parent, player, infoset): call after the
        primary arrays are set."""
        n = self.num_nodes
        self.depth = np.zeros(n, np.int32)
        # children are contiguous and parents ascend, so depth is non-decreasing in id
        order_ok = np.all(np.diff(self.parent[1:]) >= 0) if n > 2 else True
        assert order_ok, "nodes are not in breadth-first order"
        counts = np.bincount(self.parent[1:], minlength=n).astype(np.int64) if n > 1 else np.zeros(n, np.int64)
        self.first_child = np.zeros(n + 1, np.int32)
        self.first_child[0] = 1
        np.cumsum(counts, out=counts)
        self.first_child[1:] = 1 + counts
        self.slot = (np.arange(n, dtype=np.int64) - self.first_child[np.maximum(self.parent, 0)]).astype(np.int32)
        self.slot[0] = 0
        # depth by levels: level d+1 = children of level d (contiguous)
        level_off = [0, 1]
        while level_off[-1] < n:
            lo, hi = level_off[-2], level_off[-1]
            nxt = int(self.first_child[hi])
            assert nxt > hi, "unreachable nodes"
            self.depth[hi:nxt] = len(level_off) - 1
            level_off.append(nxt)
        self.level_off = np.asarray(level_off, np.int64)
        # information-set tables
        dec = np.flatnonzero(self.player > 0)
        isets = self.infoset[dec]
        ni = int(isets.max()) + 1 if len(dec) else 0
        self.num_infosets = ni
        order = np.argsort(isets, kind='stable')
        self.iset_nodes = dec[order].astype(np.int32)
        self.iset_node_off = np.zeros(ni + 1, np.int64)
        np.cumsum(np.bincount(isets, minlength=ni), out=self.iset_node_off[1:])
        rep = self.iset_nodes[self.iset_node_off[:-1]] if ni else np.zeros(0, np.int32)
        self.nact = (self.first_child[rep + 1] - self.first_child[rep]).astype(np.int32)
        self.act_off = np.zeros(ni + 1, np.int64)
        np.cumsum(self.nact, out=self.act_off[1:])
        self.iset_player = self.player[rep].astype(np.int8)
        self.iset_depth = self.depth[rep].astype(np.int32)
        nchild = self.first_child[1:] - self.first_child[:-1]
        assert np.all(nchild[dec] == self.nact[isets]), "nodes of one information set disagree on #actions"
        assert np.all(self.player[dec] == self.iset_player[isets]), "information set spans players"
        assert np.all(self.depth[dec] == self.iset_depth[isets]), "information set spans depths"
        term = self.player == -1
        self.zero_sum = bool(np.all(self.u2[term] == -self.u1[term]))
        return self


def flatten(game):
    """ExtensiveGame (ours or any object with the same node attributes) -> FlatTree."""
    cached = getattr(game, '_flat_tree', None)
    if cached is not None:
        return cached
    info_ids = game.info_set_ids
    nodes = [game.root]
    parent = [-1]
    label = [0]
    labels, label_index = [], {}
    queue = collections.deque([(game.root, 0)])
    while queue:
        node, idx = queue.popleft()
        for action, child in node.children.items():
            li = label_index.get(action)
            if li is None:
                li = label_index[action] = len(labels)
                labels.append(action)
            queue.append((child, len(nodes)))
            nodes.append(child)
            parent.append(idx)
            label.append(li)
    n = len(nodes)
    ft = FlatTree()
    ft.num_nodes = n
    ft.parent = np.asarray(parent, np.int32)
    ft.label = np.asarray(label, np.int32)
    ft.labels = labels
    ft.player = np.fromiter((nd.player for nd in nodes), np.int8, n)
    ft.hidden = np.zeros(n, np.uint8)
    ft.cprob = np.ones(n, np.float64)
    ft.u1 = np.zeros(n, np.float64)
    ft.u2 = np.zeros(n, np.float64)
    ft.infoset = np.full(n, -1, np.int32)
    keys, key_index = [], {}
    pos = 1
    for idx, nd in enumerate(nodes):
        hf = nd.hidden_from
        if hf:
            ft.hidden[idx] = (1 if 1 in hf else 0) | (2 if 2 in hf else 0)
        if nd.player == -1:
            ft.u1[idx] = nd.utility[1]
            ft.u2[idx] = nd.utility[2]
        elif nd.player == 0:
            for action in nd.children:
                ft.cprob[pos] = nd.chance_probs[action]
                pos += 1
            continue
        else:
            key = info_ids[nd]
            ki = key_index.get(key)
            if ki is None:
                ki = key_index[key] = len(keys)
                keys.append(key)
            ft.infoset[idx] = ki
        pos += len(nd.children)
    ft._keys = keys
    ft._key_fn = None
    ft.finalize()
    try:
        game._flat_tree = ft
    except AttributeError:
        pass
    return ft


def strategy_to_array(ft, strategy, fill_uniform=True):
    """Strategy (dict-like: key -> {action: p}) -> (sigma[Q], present[I]).  Missing information
    sets are uniform (extensive_game.py:425-440)."""
    sigma = np.empty(ft.num_pairs, np.float64)
    present = np.zeros(ft.num_infosets, bool)
    keys = ft.info_set_keys
    for i in range(ft.num_infosets):
        lo, k = int(ft.act_off[i]), int(ft.nact[i])
        key = keys[i]
        if key in strategy:
            probs = strategy[key]
            sigma[lo:lo + k] = [probs[a] for a in ft.actions_of(i)]
            present[i] = True
        else:
            sigma[lo:lo + k] = 1.0 / float(k) if fill_uniform else 0.0
    return sigma, present


def array_to_strategy_dict(ft, values, present=None, order=None, wrap=None):
    """values[Q] -> {key: {action: value}} for the information sets flagged in ``present``
    (all when None), emitted in ``order`` (canonical when None)."""
    out = {}
    keys = ft.info_set_keys
    for i in (range(ft.num_infosets) if order is None else order):
        if present is not None and not present[i]:
            continue
        lo = int(ft.act_off[i])
        entry = {a: float(values[lo + j]) for j, a in enumerate(ft.actions_of(i))}
        out[keys[i]] = wrap(entry) if wrap else entry
    return out


def postorder_infoset_order(ft):
    """Information sets ordered by first appearance in depth-first POST-order of the decision nodes: the dict
    order in which cfr_metrics.compute_immediate_regret (cfr_metrics.py:40-52,74) sums its per-set terms.
    Vectorised: pre-order rank from subtree sizes, post rank = pre - depth + size - 1."""
    n = ft.num_nodes
    fc = ft.first_child.astype(np.int64)
    size = np.ones(n, np.int64)
    levels = ft.level_off
    for d in range(ft.num_levels - 2, -1, -1):
        lo, hi, nxt = int(levels[d]), int(levels[d + 1]), int(levels[d + 2])
        csum = np.concatenate([[0], np.cumsum(size[hi:nxt])])
        size[lo:hi] += csum[fc[lo + 1:hi + 1] - hi] - csum[fc[lo:hi] - hi]
    pre = np.zeros(n, np.int64)
    for d in range(ft.num_levels - 1):
        lo, hi, nxt = int(levels[d]), int(levels[d + 1]), int(levels[d + 2])
        if nxt == hi:
            break
        csum = np.concatenate([[0], np.cumsum(size[hi:nxt])])
        par = ft.parent[hi:nxt].astype(np.int64)
        pre[hi:nxt] = pre[par] + 1 + csum[:-1] - csum[fc[par] - hi]
    post = pre - ft.depth + size - 1
    dec = np.flatnonzero(ft.player > 0)
    by_post = dec[np.argsort(post[dec], kind='stable')]
    isets = ft.infoset[by_post]
    _, first = np.unique(isets, return_index=True)
    return isets[np.sort(first)].astype(np.int32)
