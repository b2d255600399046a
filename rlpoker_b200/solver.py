"""Device-side solver objects: thin, typed wrappers over the C ABI (include/rlpoker_b200.h).

``DeviceTree`` = a flattened game resident in HBM; ``DeviceCFR`` = regrets / action counts /
strategy + scratch for that tree.  The reference keeps all of this in Python dicts
(rlpoker/cfr/cfr.py:78-99); here the dicts exist only at the API edge.
"""
import ctypes as C

import numpy as np

from rlpoker_b200 import _lib
from rlpoker_b200._lib import (RLP_AVERAGE, RLP_CHANCE_SAMPLED, RLP_CUM_IMM_REGRET, RLP_EXTERNAL_SAMPLED, RLP_REGRET,
                               RLP_SIGMA, RLP_STRAT_SUM, check, ptr)


def _stream_ptr(stream):
    """cudaStream_t of a torch stream / int / None (None = torch's current stream when torch has
    initialised CUDA, else the legacy default stream)."""
    if stream is None:
        try:
            import torch
            if torch.cuda.is_initialized():
                return C.c_void_p(torch.cuda.current_stream().cuda_stream)
        except Exception:
            pass
        return C.c_void_p(0)
    if hasattr(stream, 'cuda_stream'):
        return C.c_void_p(stream.cuda_stream)
    return C.c_void_p(int(stream))


def _tree_desc(ft):
    keep = [np.ascontiguousarray(ft.first_child, np.int32), np.ascontiguousarray(ft.player, np.int8),
            np.ascontiguousarray(ft.infoset, np.int32), np.ascontiguousarray(ft.cprob, np.float64),
            np.ascontiguousarray(ft.u1, np.float64), np.ascontiguousarray(ft.u2, np.float64),
            np.ascontiguousarray(ft.act_off, np.int64), np.ascontiguousarray(ft.level_off, np.int64)]
    desc = _lib.TreeDesc(
        num_nodes=ft.num_nodes, num_infosets=ft.num_infosets, num_levels=ft.num_levels,
        first_child=ptr(keep[0], _lib._i32p), player=ptr(keep[1], _lib._i8p), infoset=ptr(keep[2], _lib._i32p),
        cprob=ptr(keep[3], _lib._dp), u1=ptr(keep[4], _lib._dp), u2=ptr(keep[5], _lib._dp),
        act_off=ptr(keep[6], _lib._i64p), level_off=ptr(keep[7], _lib._i64p))
    return desc, keep


def plan_tree(flat_tree):
    """Host-only dry run of the device layout (rlp_tree_plan): works without a GPU."""
    desc, keep = _tree_desc(flat_tree)
    out = np.zeros(16, np.int64)
    check(_lib.load().rlp_tree_plan(C.byref(desc), ptr(out, _lib._i64p)))
    names = ['tiled', 'upper_tiles', 'micro_subtrees', 'micro_shapes', 'micro_nodes', 'micro_decisions', 'micro_leaves',
             'infosets_micro_only', 'infosets_upper', 'tile_nodes', 'tile_nonterminals', 'tile_smem', 'kernels_compiled',
             'kernels_failed', 'largest_infoset', 'zero_sum']
    return dict(zip(names, (int(x) for x in out)))


def plan_lane_records(flat_tree, variant):
    """Host-only: per level (number of explored ancestors) the most records ONE lane of the level-by-level sampled
    kernels can hold (rlp_tree_plan_lane_records); the device buffers are ``lanes`` times this."""
    desc, keep = _tree_desc(flat_tree)
    out = np.zeros(33, np.int64)
    check(_lib.load().rlp_tree_plan_lane_records(C.byref(desc), int(variant), ptr(out, _lib._i64p)))
    return out


class DeviceTree:
    def __init__(self, flat_tree, stream=None):
        _lib.require_device()
        ft = self.flat = flat_tree
        self._keep = [np.ascontiguousarray(ft.first_child, np.int32), np.ascontiguousarray(ft.player, np.int8),
                      np.ascontiguousarray(ft.infoset, np.int32), np.ascontiguousarray(ft.cprob, np.float64),
                      np.ascontiguousarray(ft.u1, np.float64), np.ascontiguousarray(ft.u2, np.float64),
                      np.ascontiguousarray(ft.act_off, np.int64), np.ascontiguousarray(ft.level_off, np.int64)]
        k = self._keep
        desc = _lib.TreeDesc(
            num_nodes=ft.num_nodes, num_infosets=ft.num_infosets, num_levels=ft.num_levels,
            first_child=ptr(k[0], _lib._i32p), player=ptr(k[1], _lib._i8p), infoset=ptr(k[2], _lib._i32p),
            cprob=ptr(k[3], _lib._dp), u1=ptr(k[4], _lib._dp), u2=ptr(k[5], _lib._dp),
            act_off=ptr(k[6], _lib._i64p), level_off=ptr(k[7], _lib._i64p))
        self.handle = C.c_void_p()
        check(_lib.load().rlp_tree_upload(C.byref(desc), _stream_ptr(stream), C.byref(self.handle)))
        self._keep = None
        self.num_pairs = ft.num_pairs

    @property
    def device_bytes(self):
        return _lib.load().rlp_tree_bytes(self.handle)

    @property
    def layout(self):
        names = ['tiled', 'upper_tiles', 'micro_subtrees', 'micro_shapes', 'jit_shapes', 'zero_sum',
                 'largest_infoset', 'tile_block', 'tile_smem', 'upper_jit', 'fused', 'fused_batches',
                 'fused_upper_groups', 'fused_grid', 'components', 'fused_br']
        return {n: int(_lib.load().rlp_tree_stat(self.handle, i)) for i, n in enumerate(names)}

    def best_response(self, sigma, player, want_argmax=False, stream=None):
        """(value[, argmax[I]]) of ``player``'s best response to the complete strategy array sigma[Q]."""
        sigma = np.ascontiguousarray(sigma, np.float64)
        assert sigma.shape == (self.num_pairs,)
        value = C.c_double()
        arg = np.empty(self.flat.num_infosets, np.int32) if want_argmax else None
        check(_lib.load().rlp_best_response(self.handle, ptr(sigma, _lib._dp), int(player), C.byref(value),
                                            ptr(arg, _lib._i32p) if want_argmax else None, _stream_ptr(stream)))
        return (value.value, arg) if want_argmax else value.value

    def exploitability(self, sigma, stream=None):
        return self.best_response(sigma, 1, stream=stream) + self.best_response(sigma, 2, stream=stream)

    def expected_value(self, sigma, stream=None):
        sigma = np.ascontiguousarray(sigma, np.float64)
        out = np.empty(2)
        check(_lib.load().rlp_expected_value(self.handle, ptr(sigma, _lib._dp), ptr(out, _lib._dp), _stream_ptr(stream)))
        return float(out[0]), float(out[1])

    def close(self):
        """Free the device copy (solvers created on this tree must be gone first)."""
        h, self.handle = self.handle, None
        if h:
            _lib.load().rlp_tree_free(h)

    def __del__(self):
        # Safe: every DeviceCFR holds a reference to its tree, so all solvers are freed before the tree is.
        try:
            self.close()
        except Exception:
            pass


def device_tree(game_or_flat, stream=None):
    """Flatten (once) and upload (once per game object)."""
    from rlpoker_b200.flatten import FlatTree, flatten
    if isinstance(game_or_flat, DeviceTree):
        return game_or_flat
    if isinstance(game_or_flat, FlatTree):
        ft, holder = game_or_flat, game_or_flat
    else:
        ft, holder = flatten(game_or_flat), game_or_flat
    dt = getattr(holder, '_device_tree', None)
    if dt is None:
        dt = DeviceTree(ft, stream=stream)
        try:
            holder._device_tree = dt
        except AttributeError:
            pass
    return dt


class DeviceCFR:
    def __init__(self, tree, stream=None):
        self.tree = tree            # keeps the tree (and its device memory) alive as long as the solver
        self.handle = C.c_void_p()
        check(_lib.load().rlp_cfr_create(tree.handle, _stream_ptr(stream), C.byref(self.handle)))
        self.t = 0
        self._touched_base = 0      # nodes touched before a resume point (SolverCheckpoint.load)

    def close(self):
        h, self.handle = getattr(self, 'handle', None), None
        if h:
            _lib.load().rlp_cfr_free(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self, stream=None):
        check(_lib.load().rlp_cfr_reset(self.handle, _stream_ptr(stream)))
        self.t = 0
        self._touched_base = 0

    def vanilla(self, n_iters, linear_weight=False, stream=None):
        check(_lib.load().rlp_cfr_vanilla_iters(self.handle, self.t, int(n_iters), int(bool(linear_weight)),
                                                _stream_ptr(stream)))
        self.t += int(n_iters)

    def sampled(self, variant, n_iters, lanes=1, seed=0, linear_weight=False, stream=None):
        check(_lib.load().rlp_cfr_sampled_iters(self.handle, int(variant), int(lanes), int(seed), self.t, int(n_iters),
                                                int(bool(linear_weight)), _stream_ptr(stream)))
        self.t += int(n_iters)

    def read(self, which, stream=None):
        out = np.empty(self.tree.num_pairs, np.float64)
        check(_lib.load().rlp_cfr_read(self.handle, which, ptr(out, _lib._dp), _stream_ptr(stream)))
        return out

    def write(self, which, values, stream=None):
        values = np.ascontiguousarray(values, np.float64)
        assert values.shape == (self.tree.num_pairs,)
        check(_lib.load().rlp_cfr_write(self.handle, which, ptr(values, _lib._dp), _stream_ptr(stream)))

    def flags(self, stream=None):
        out = np.empty(self.tree.flat.num_infosets, np.uint8)
        check(_lib.load().rlp_cfr_read_flags(self.handle, ptr(out, _lib._u8p), _stream_ptr(stream)))
        return out

    def profile_stages(self, reps=20, stream=None):
        """Mean device microseconds per iteration of the sweep stages (see rlp_cfr_profile_stages)."""
        out = (C.c_float * 4)()
        check(_lib.load().rlp_cfr_profile_stages(self.handle, int(reps), out, _stream_ptr(stream)))
        self.t += reps + 2
        names = ['upper_phase0', 'micro', 'upper_phase1', 'accumulate_rm']
        return {n: 1e3 * out[i] / reps for i, n in enumerate(names)}

    def fused_timeline(self, iters=20, stream=None):
        """Per-iteration phase times (microseconds, as seen by CTA 0) of the group-fused persistent kernel:
        micro phase, first grid barrier, upper phase, second barrier + loop turn-around.  Runs `iters` iterations."""
        buf = (C.c_uint64 * (9 * iters))()
        check(_lib.load().rlp_cfr_fused_timeline(self.handle, self.t, int(iters), buf, _stream_ptr(stream)))
        self.t += iters
        t = np.array(list(buf), np.float64).reshape(iters, 9) / 1e3
        out = {'micro_phase': float(np.mean(t[:, 1] - t[:, 0])), 'barrier_1': float(np.mean(t[:, 2] - t[:, 1])),
               'upper_phase': float(np.mean(t[:, 8] - t[:, 2]))}
        if np.all(t[:, 3:8] > 0):          # the recording CTA owns an upper unit: split of that unit
            out['upper_split'] = {'issue_loads': float(np.mean(t[:, 3] - t[:, 2])), 'fan_in': float(np.mean(t[:, 4] - t[:, 3])),
                                  'tile': float(np.mean(t[:, 5] - t[:, 4])), 'accumulate_rm': float(np.mean(t[:, 6] - t[:, 5])),
                                  'reach': float(np.mean(t[:, 7] - t[:, 6]))}
        if iters > 1:
            out['barrier_2'] = float(np.mean(t[1:, 0] - t[:-1, 8]))
            out['iteration'] = float((t[-1, 0] - t[0, 0]) / (iters - 1))
        return out

    def timeline(self, iters=16, stream=None):
        """[(kind, start_us, end_us)] of block 0 of every sweep kernel over `iters` iterations, relative to the first."""
        cap = 8 * iters + 16
        buf = (C.c_uint64 * (3 * cap))()
        n = C.c_int()
        check(_lib.load().rlp_cfr_timeline(self.handle, int(iters), buf, cap, C.byref(n), _stream_ptr(stream)))
        self.t += iters
        rec = [(int(buf[3 * i]), buf[3 * i + 1], buf[3 * i + 2]) for i in range(n.value)]
        rec.sort(key=lambda r: r[1])
        t0 = rec[0][1] if rec else 0
        return [(k, (a - t0) / 1e3, (b - t0) / 1e3) for k, a, b in rec]

    def write_flags(self, mask, bit=None, stream=None):
        """Set the dict-membership flags: ``mask`` is a [I] bitmask, or a boolean array for one ``bit`` (1, 2, 4)
        merged into the current flags."""
        mask = np.asarray(mask)
        if bit is not None:
            cur = self.flags()
            mask = (cur & ~np.uint8(bit)) | np.where(mask.astype(bool), np.uint8(bit), np.uint8(0))
        mask = np.ascontiguousarray(mask, np.uint8)
        check(_lib.load().rlp_cfr_write_flags(self.handle, ptr(mask, _lib._u8p), _stream_ptr(stream)))

    regrets = property(lambda self: self.read(RLP_REGRET))
    strategy_sum = property(lambda self: self.read(RLP_STRAT_SUM))
    sigma = property(lambda self: self.read(RLP_SIGMA))
    average = property(lambda self: self.read(RLP_AVERAGE))

    @property
    def nodes_touched(self):
        return self._touched_base + int(_lib.load().rlp_nodes_touched(self.handle))

    def set_nodes_touched(self, total):
        """Make ``nodes_touched`` continue from ``total`` (resume from a checkpoint)."""
        self._touched_base = int(total) - int(_lib.load().rlp_nodes_touched(self.handle))

    def exploitability(self, which=RLP_AVERAGE, stream=None):
        """(total, BR1, BR2) of the average (default) or current strategy, computed on the device."""
        out = np.empty(3)
        check(_lib.load().rlp_cfr_exploitability(self.handle, which, ptr(out, _lib._dp), _stream_ptr(stream)))
        return float(out[0]), float(out[1]), float(out[2])

    def immediate_regret_step(self, step, order=None, stream=None, detail=False):
        """One step of cfr_metrics.compute_immediate_regret for the current strategy; with ``detail`` also the [I]
        per-information-set terms (canonical order)."""
        out = C.c_double()
        if order is not None:
            order = np.ascontiguousarray(order, np.int32)
        per_iset = np.empty(self.tree.flat.num_infosets, np.float64) if detail else None
        check(_lib.load().rlp_cfr_immediate_regret_detail(self.handle, int(step),
                                                          ptr(order, _lib._i32p) if order is not None else None,
                                                          C.byref(out), ptr(per_iset, _lib._dp) if detail else None,
                                                          _stream_ptr(stream)))
        return (out.value, per_iset) if detail else out.value

    # multi-GPU split
    def local_sweep(self, linear_weight=False, stream=None, keep_counter=False):
        check(_lib.load().rlp_cfr_local_sweep(self.handle, -1 if keep_counter else self.t, int(bool(linear_weight)),
                                              _stream_ptr(stream)))

    def delta_ptr(self):
        return _lib.load().rlp_cfr_delta_ptr(self.handle)

    def apply_delta(self, stream=None):
        check(_lib.load().rlp_cfr_apply_delta(self.handle, _stream_ptr(stream)))
        self.t += 1
