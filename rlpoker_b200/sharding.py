"""Deal sharding of a flattened game across ranks.

Below the root chance deals the tree is a forest of subtrees that share only the strategy (read) and the
per-information-set accumulators (write).  Rank r keeps a contiguous, size-balanced range of the depth-``cut``
subtrees plus their ancestors; information-set numbering stays GLOBAL so the per-pair deltas of all ranks
line up and one sum-allreduce per iteration (2*Q doubles) completes the update.  Reach probabilities of the
kept nodes are unchanged because chance probabilities are kept as they are.
"""
import os

import numpy as np

from rlpoker_b200.flatten import FlatTree


class _Sharded(FlatTree):
    """Rank-local tree; ``global_nodes[i]`` is the id of local node i in the full tree."""
    __slots__ = ('global_nodes',)


def choose_cut(ft, world):
    """Shallowest depth d such that every node above d is a chance node and depth d has >= 8*world nodes
    (falls back to the deepest all-chance prefix)."""
    best = None
    for d in range(1, ft.num_levels):
        above = ft.player[:ft.level_off[d]]
        if np.any(above != 0):
            break
        best = d
        if ft.level_off[d + 1] - ft.level_off[d] >= 8 * world:
            break
    if best is None:
        raise ValueError("this game does not shard: the root is not a chance node (run replicas instead)")
    return best


def subtree_sizes(ft):
    size = np.ones(ft.num_nodes, np.int64)
    for d in range(ft.num_levels - 1, 0, -1):
        lo, hi = int(ft.level_off[d]), int(ft.level_off[d + 1])
        np.add.at(size, ft.parent[lo:hi], size[lo:hi])
    return size


def shard_ranges(ft, world, cut=None):
    """[(lo, hi)) node-id ranges of depth-``cut`` subtree roots per rank, balanced by subtree size."""
    cut = choose_cut(ft, world) if cut is None else cut
    lo, hi = int(ft.level_off[cut]), int(ft.level_off[cut + 1])
    sizes = subtree_sizes(ft)[lo:hi]
    csum = np.concatenate([[0], np.cumsum(sizes)])
    bounds = [lo + int(np.searchsorted(csum, csum[-1] * r / world, side='left')) for r in range(world)] + [hi]
    bounds[0] = lo
    return cut, [(bounds[r], bounds[r + 1]) for r in range(world)]


def shard_flat_tree(ft, rank, world, cut=None):
    """The rank-local FlatTree (global information-set ids; possibly empty information sets)."""
    if world == 1:
        return ft
    cut, ranges = shard_ranges(ft, world, cut)
    lo, hi = ranges[rank]
    n = ft.num_nodes
    keep = np.zeros(n, bool)
    keep[lo:hi] = True
    for d in range(cut + 1, ft.num_levels):                   # descendants
        a, b = int(ft.level_off[d]), int(ft.level_off[d + 1])
        keep[a:b] = keep[ft.parent[a:b]]
    for d in range(cut, 0, -1):                               # ancestors
        a, b = int(ft.level_off[d]), int(ft.level_off[d + 1])
        keep[ft.parent[a:b][keep[a:b]]] = True
    new_id = np.cumsum(keep) - 1
    idx = np.flatnonzero(keep)
    out = _Sharded()
    out.num_nodes = len(idx)
    out.parent = np.where(ft.parent[idx] >= 0, new_id[np.maximum(ft.parent[idx], 0)], -1).astype(np.int32)
    for name in ('player', 'infoset', 'cprob', 'u1', 'u2', 'label', 'hidden'):
        setattr(out, name, getattr(ft, name)[idx].copy())
    out.labels = ft.labels
    out._keys, out._key_fn = None, (lambda: ft.info_set_keys)
    _finalize_sharded(out, ft)
    out.global_nodes = idx
    return out


def _finalize_sharded(out, full):
    """Like FlatTree.finalize but with the GLOBAL information-set tables (empty sets allowed)."""
    n = out.num_nodes
    counts = np.bincount(out.parent[1:], minlength=n).astype(np.int64)
    out.first_child = np.zeros(n + 1, np.int32)
    out.first_child[0] = 1
    out.first_child[1:] = 1 + np.cumsum(counts)
    out.slot = (np.arange(n, dtype=np.int64) - out.first_child[np.maximum(out.parent, 0)]).astype(np.int32)
    out.slot[0] = 0
    out.depth = np.zeros(n, np.int32)
    level_off = [0, 1]
    while level_off[-1] < n:
        hi = level_off[-1]
        nxt = int(out.first_child[hi])
        out.depth[hi:nxt] = len(level_off) - 1
        level_off.append(nxt)
    out.level_off = np.asarray(level_off, np.int64)
    out.num_infosets = full.num_infosets
    out.nact, out.act_off = full.nact, full.act_off
    out.iset_player, out.iset_depth = full.iset_player, full.iset_depth
    dec = np.flatnonzero(out.player > 0)
    isets = out.infoset[dec]
    order = np.argsort(isets, kind='stable')
    out.iset_nodes = dec[order].astype(np.int32)
    out.iset_node_off = np.zeros(full.num_infosets + 1, np.int64)
    np.cumsum(np.bincount(isets, minlength=full.num_infosets), out=out.iset_node_off[1:])
    out.zero_sum = full.zero_sum




# ---------------------------------------------------------------------------------------------------
# the per-iteration exchange


def allreduce_delta(delta, group=None):
    """Sum the [2*Q] (regret || strategy-sum) delta buffer over all ranks, in place (NCCL on CUDA
    tensors, gloo on CPU tensors)."""
    import torch.distributed as dist
    dist.all_reduce(delta, op=dist.ReduceOp.SUM, group=group)
    return delta


class _DeviceArray:
    """CUDA-array-interface view of library-owned device memory (so torch can all-reduce it in place)."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {'shape': (count,), 'typestr': '<f8', 'data': (int(ptr), False), 'version': 2}


class ShardedCFR:
    """Vanilla CFR with the deal subtrees sharded over the ranks of a torch.distributed group.

    Every rank holds the full [Q] regret / strategy-sum / strategy arrays (replicated, kept identical by
    construction) and sweeps only its own subtrees; one all-reduce of the deltas per iteration.  Summation
    order across ranks differs from the single-GPU order, so parity with the reference is by tolerance
    (1e-9 on strategies) rather than bit-exact."""

    def __init__(self, flat_tree, rank, world, group=None, stream=None):
        import torch
        from rlpoker_b200.solver import DeviceCFR, DeviceTree
        self.rank, self.world, self.group = rank, world, group
        self.full = flat_tree
        self.local = shard_flat_tree(flat_tree, rank, world)
        self.tree = DeviceTree(self.local, stream=stream)
        self.solver = DeviceCFR(self.tree, stream=stream)
        self.delta = torch.as_tensor(_DeviceArray(self.solver.delta_ptr(), 2 * flat_tree.num_pairs), device='cuda')
        self.nodes_touched = 0
        self._cuda_graph = None
        self._last_linear = None
        self._native = False
        if os.environ.get('RLP_SHARD_NATIVE', '1') == '1':
            self._init_native_comm()

    def _init_native_comm(self):
        """NCCL communicator owned by the library: the whole loop then runs in C (rlp_cfr_sharded_iters)."""
        import ctypes as C
        import torch
        import torch.distributed as dist
        from rlpoker_b200 import _lib
        lib = _lib.load()
        ident = torch.zeros(128, dtype=torch.uint8)
        ok = torch.ones(1, dtype=torch.int32)
        if self.world > 1:
            if not dist.is_initialized():
                return
            if self.rank == 0:
                buf = (C.c_uint8 * 128)()
                if lib.rlp_comm_unique_id(buf) != 0:
                    ok[0] = 0
                ident = torch.tensor(list(buf), dtype=torch.uint8)
            dev = torch.device('cuda', torch.cuda.current_device())
            ident_d, ok_d = ident.to(dev), ok.to(dev)
            dist.broadcast(ident_d, src=0, group=self.group)
            dist.broadcast(ok_d, src=0, group=self.group)
            if int(ok_d.item()) == 0:
                return
            ident = ident_d.cpu()
        else:
            buf = (C.c_uint8 * 128)()
            if lib.rlp_comm_unique_id(buf) != 0:
                return
            ident = torch.tensor(list(buf), dtype=torch.uint8)
        arr = (C.c_uint8 * 128)(*ident.tolist())
        rc = lib.rlp_comm_init(self.solver.handle, self.rank, self.world, arr)
        ok_all = rc == 0
        if self.world > 1:
            # every rank must take the same path: agree on the outcome before anyone enters a collective
            flag = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32, device=torch.device('cuda', torch.cuda.current_device()))
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            ok_all = int(flag.item()) == 1
            if not ok_all:
                lib.rlp_comm_destroy(self.solver.handle)
                raise _lib.RlpError("rlp_comm_init failed on at least one rank (this rank: status %d, %s)"
                                    % (rc, lib.rlp_last_error().decode()))
        self._native = ok_all

    def close(self):
        """Free the solver, its communicator and the rank-local device tree."""
        self.delta = None
        self._cuda_graph = None
        solver, self.solver = self.solver, None
        if solver is not None:
            solver.close()
        tree, self.tree = self.tree, None
        if tree is not None:
            tree.close()

    GRAPH_ITERS = 10

    def _one(self, linear_weight, keep_counter):
        self.solver.local_sweep(linear_weight=linear_weight, keep_counter=keep_counter)
        if self.world > 1:
            allreduce_delta(self.delta, self.group)
        self.solver.apply_delta()

    def _graph(self):
        """CUDA graph of GRAPH_ITERS iterations (local sweep, NCCL all-reduce, apply) -- without it the loop is
        bound by host launch latency, not by the GPUs.  The graph continues from the device-side iteration counter
        and weighting flag, which every eager iteration sets."""
        import torch
        if self._cuda_graph is None:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            t_before = self.solver.t
            with torch.cuda.graph(g):
                for _ in range(self.GRAPH_ITERS):
                    self._one(False, keep_counter=True)
            self.solver.t = t_before                              # capture records, it does not run
            self._cuda_graph = g
        return self._cuda_graph

    def iterate(self, n_iters, linear_weight=False, use_graph=None):
        # Graph replay is validated single-process only (tests/test_gpu_sharded.py); a 2-rank run with the NCCL
        # all-reduce captured inside torch.cuda.graph hung in round 1, so multi-rank runs stay eager unless
        # RLP_SHARD_GRAPH=1 asks for it.
        n = int(n_iters)
        if self._native and use_graph is None:
            from rlpoker_b200 import _lib
            from rlpoker_b200.solver import _stream_ptr
            _lib.check(_lib.load().rlp_cfr_sharded_iters(self.solver.handle, self.solver.t, n, int(bool(linear_weight)),
                                                         _stream_ptr(None)))
            self.solver.t += n
            self._last_linear = None
            self.nodes_touched += 2 * self.full.num_nodes * n
            return
        if use_graph is None:
            use_graph = self.world == 1 or os.environ.get('RLP_SHARD_GRAPH') == '1' 
        done = 0
        if use_graph and n >= self.GRAPH_ITERS:
            if self._last_linear != bool(linear_weight):          # (re)sets counter + flag on the device; warms NCCL up
                self._one(linear_weight, keep_counter=False)
                self._last_linear = bool(linear_weight)
                done += 1
            g = self._graph()
            while n - done >= self.GRAPH_ITERS:
                g.replay()
                self.solver.t += self.GRAPH_ITERS
                done += self.GRAPH_ITERS
        for _ in range(n - done):
            self._one(linear_weight, keep_counter=False)
            self._last_linear = bool(linear_weight)
        self.nodes_touched += 2 * self.full.num_nodes * n


# ---------------------------------------------------------------------------------------------------
# sharding by public state (the group-fused kernel's multi-GPU mode)


class PublicStateShardedCFR:
    """Vanilla CFR over the ranks of a torch.distributed group, sharded by PUBLIC STATE instead of by deal.

    Every rank holds the full tree; the micro subtrees fall into components that share no information set (Leduc: one
    per board card and round-1 betting line) and rank r sweeps a contiguous range of them.  The only per-iteration
    exchange is the value of every micro subtree, stored straight into the peers' buffers over NVLink by the micro
    phase of the persistent kernel (csrc/fused.cu); there is no all-reduce and no host involvement inside a call.
    Regrets / action counts / strategy of a component live on its rank only: ``merged()`` assembles the full arrays
    (every information set is owned by exactly one rank, so the merge is a sum with zeros: exact).  All sums keep the
    single-GPU order, hence results are bit-identical to one GPU and to the reference."""

    def __init__(self, flat_tree, rank, world, group=None, stream=None):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from rlpoker_b200 import _lib
        from rlpoker_b200.solver import DeviceCFR, DeviceTree
        if world > 8:
            raise ValueError("sharding by public state works inside one NVLink domain (<= 8 ranks)")
        self.rank, self.world, self.group = rank, world, group
        self.full = flat_tree
        self.tree = DeviceTree(flat_tree, stream=stream)
        self.solver = DeviceCFR(self.tree, stream=stream)
        if self.tree.layout['fused'] != 1:
            raise _lib.RlpError("this tree does not run the group-fused kernel: use ShardedCFR (deal sharding)")
        lib = _lib.load()
        mine = (C.c_uint8 * 64)()
        _lib.check(lib.rlp_cfr_p2p_export(self.solver.handle, mine))
        dev = torch.device('cuda', torch.cuda.current_device())
        handles = torch.zeros(world, 64, dtype=torch.uint8, device=dev)
        me = torch.tensor(list(mine), dtype=torch.uint8, device=dev)
        if world > 1:
            dist.all_gather_into_tensor(handles.view(-1), me, group=group)
        else:
            handles[0] = me
        flat = handles.cpu().numpy().reshape(-1)
        arr = (C.c_uint8 * (64 * world))(*flat.tolist())
        rc = lib.rlp_cfr_p2p_connect(self.solver.handle, rank, world, arr)
        ok = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32, device=dev)
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)       # every rank takes the same path
        if int(ok.item()) != 1:
            raise _lib.RlpError("rlp_cfr_p2p_connect failed on at least one rank (this rank: status %d, %s)"
                                % (rc, lib.rlp_last_error().decode()))
        owned = np.zeros(flat_tree.num_infosets, np.uint8)
        _lib.check(lib.rlp_cfr_p2p_owned(self.solver.handle, _lib.ptr(owned, _lib._u8p)))
        self.owned_isets = owned.astype(bool)
        self.owned_pairs = np.repeat(self.owned_isets, flat_tree.nact)
        self.nodes_touched = 0

    def iterate(self, n_iters, linear_weight=False):
        """The same call on every rank; returns without synchronising (the ranks meet inside the kernel)."""
        from rlpoker_b200 import _lib
        self.solver.vanilla(int(n_iters), linear_weight=linear_weight)
        self.nodes_touched += 2 * self.full.num_nodes * int(n_iters)

    def check(self):
        """Raises if a peer was lost during the last call (bounded wait inside the kernel)."""
        import torch
        from rlpoker_b200 import _lib
        torch.cuda.synchronize()
        if _lib.load().rlp_cfr_p2p_failed(self.solver.handle):
            raise _lib.RlpError("a peer GPU did not reach the exchange barrier in time; the solver state is undefined")

    def merged(self, which):
        """The full [Q] array (canonical order) assembled from the owners, on every rank."""
        import torch
        import torch.distributed as dist
        local = self.solver.read(which)
        local[~self.owned_pairs] = 0.0
        if self.world == 1:
            return local
        t = torch.from_numpy(local).cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)      # x + 0 + ... + 0: exact
        return t.cpu().numpy()

    def gather_state(self):
        """Make this rank's solver hold the complete regrets / action counts / strategy (e.g. before an evaluation or
        a checkpoint); the solve can simply continue afterwards."""
        from rlpoker_b200._lib import RLP_REGRET, RLP_SIGMA, RLP_STRAT_SUM
        for which in (RLP_REGRET, RLP_STRAT_SUM, RLP_SIGMA):
            self.solver.write(which, self.merged(which))

    def close(self):
        solver, self.solver = self.solver, None
        if solver is not None:
            solver.close()
        tree, self.tree = self.tree, None
        if tree is not None:
            tree.close()


# ---------------------------------------------------------------------------------------------------
# sampled variants: lanes sharded over ranks


class _DeviceBytes:
    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {'shape': (count,), 'typestr': '|u1', 'data': (int(ptr), False), 'version': 2}


class ShardedSampledCFR:
    """Chance-sampled / external-sampling CFR with the traversal LANES of every iteration split over the ranks
    (SURVEY 8e).  Every rank holds the whole tree and identical regrets / sums / strategy; lane ``l`` draws from the
    Philox stream keyed by its GLOBAL id, so the result does not depend on the number of ranks (up to the order in
    which the floating-point atomics add).  One exchange per iteration: a sum-all-reduce of the [2Q] deltas and a
    max-all-reduce of the dict-membership flags (NCCL)."""

    def __init__(self, flat_tree, rank, world, group=None, stream=None):
        import torch
        from rlpoker_b200 import _lib
        from rlpoker_b200.solver import DeviceCFR, DeviceTree
        self.rank, self.world, self.group = rank, world, group
        self.full = flat_tree
        self.tree = DeviceTree(flat_tree, stream=stream)
        self.solver = DeviceCFR(self.tree, stream=stream)
        lib = _lib.load()
        self.delta = torch.as_tensor(_DeviceArray(lib.rlp_cfr_delta_ptr(self.solver.handle), 2 * flat_tree.num_pairs),
                                     device='cuda')
        self.flags = torch.as_tensor(_DeviceBytes(lib.rlp_cfr_flags_ptr(self.solver.handle), 3 * flat_tree.num_infosets),
                                     device='cuda')
        self._graph = None
        self._graph_key = None

    @staticmethod
    def lane_range(lanes, rank, world):
        """[lo, hi) of the ``lanes`` global lane ids that belong to ``rank``: contiguous, sizes differ by at most one."""
        return lanes * rank // world, lanes * (rank + 1) // world

    def _one(self, variant, lo, hi, seed, t, linear_weight):
        """One iteration on torch's current stream: local lanes, the two all-reduces, finish."""
        import torch.distributed as dist
        from rlpoker_b200 import _lib
        from rlpoker_b200.solver import _stream_ptr
        lib = _lib.load()
        _lib.check(lib.rlp_cfr_sampled_local(self.solver.handle, int(variant), lo, hi - lo, int(seed), int(t),
                                             int(bool(linear_weight)), _stream_ptr(None)))
        if self.world > 1:
            dist.all_reduce(self.delta, op=dist.ReduceOp.SUM, group=self.group)
            dist.all_reduce(self.flags, op=dist.ReduceOp.MAX, group=self.group)
        _lib.check(lib.rlp_cfr_sampled_finish(self.solver.handle, int(variant), _stream_ptr(None)))

    def iterate(self, variant, n_iters, lanes, seed=0, linear_weight=False, graph=None):
        """``n_iters`` iterations of ``lanes`` global lanes.  An iteration is ~70 launches (level-by-level lanes, the
        average-strategy walk, two all-reduces); split over 8 ranks the kernels are short and the launches dominate, so
        from the third iteration of a call on the iteration is replayed as ONE CUDA graph (``graph=False``: eager).
        The graph has no per-iteration argument: the device-side iteration counter keys the Philox stream."""
        import torch
        n_iters = int(n_iters)
        lo, hi = self.lane_range(int(lanes), self.rank, self.world)
        key = (int(variant), lo, hi, int(seed), bool(linear_weight))
        if graph is None:
            graph = self.world > 1
        done = 0
        # the first iterations run eagerly: they set the device-side counter and allocate whatever the lanes need
        while done < n_iters and (not graph or self._graph_key != key and done < 2 or done < 1):
            self._one(variant, lo, hi, seed, self.solver.t, linear_weight)
            self.solver.t += 1
            done += 1
        if done == n_iters:
            return
        if self._graph_key != key:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._one(variant, lo, hi, seed, -1, linear_weight)
            # the capture recorded the iteration without running it
            self._graph, self._graph_key = g, key
        for _ in range(n_iters - done):
            self._graph.replay()
            self.solver.t += 1

    def nodes_touched(self):
        """Sum over the ranks of the nodes their lanes visited."""
        import torch
        import torch.distributed as dist
        t = torch.tensor([self.solver.nodes_touched], dtype=torch.int64, device='cuda')
        if self.world > 1:
            dist.all_reduce(t, group=self.group)
        return int(t.item())

    def close(self):
        self._graph = self._graph_key = None
        self.delta = self.flags = None
        solver, self.solver = self.solver, None
        if solver is not None:
            solver.close()
        tree, self.tree = self.tree, None
        if tree is not None:
            tree.close()
