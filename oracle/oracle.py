"""ORACLE (test infrastructure): ctypes front-end of cfr_oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libcfr_oracle.so')
_lib = None

_dp = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)


class _Tree(C.Structure):
    _fields_ = [
        ('num_nodes', C.c_int64), ('num_infosets', C.c_int32), ('num_levels', C.c_int32),
        ('parent', _i32p), ('first_child', _i32p), ('player', C.POINTER(C.c_int8)), ('infoset', _i32p),
        ('label', _i32p), ('hidden', _u8p), ('cprob', _dp), ('u1', _dp), ('u2', _dp),
        ('act_off', _i64p), ('iset_node_off', _i64p), ('iset_nodes', _i32p), ('level_off', _i64p),
    ]


class _Rng(C.Structure):
    _fields_ = [
        ('kind', C.c_int), ('stream', _dp), ('stream_len', C.c_int64), ('stream_pos', C.c_int64),
        ('seed', C.c_uint64), ('iteration', C.c_uint64), ('lane', C.c_uint32), ('trav', C.c_uint32),
        ('draw', C.c_uint32), ('exhausted', C.c_int),
    ]


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ('cfr_oracle.c', 'leduc_oracle.c')]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(['make', '-s', '-C', _HERE, '-B'])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_best_response.restype = C.c_double
        _lib.orc_exploitability.restype = C.c_double
        _lib.orc_philox_uniform.restype = C.c_double
        _lib.orc_philox_uniform.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
        _lib.orc_immediate_regret_step.restype = C.c_double
        _lib.orc_immediate_regret_detail.restype = C.c_double
        _lib.orc_cfr_recursive_iters.restype = C.c_int64
        _lib.orc_vanilla_level_iters.restype = C.c_int64
        _lib.orc_external_iters.restype = C.c_int64
        _lib.orc_leduc_build.restype = C.c_int64
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


def rng_stream(uniforms):
    """RNG that replays a precomputed array of uniforms (e.g. numpy's legacy MT19937 stream)."""
    u = np.ascontiguousarray(uniforms, np.float64)
    g = _Rng(kind=0, stream=_p(u, _dp), stream_len=len(u), stream_pos=0)
    g._keep = u
    return g


def rng_philox(seed):
    return _Rng(kind=1, seed=seed)


class Oracle:
    """State of one solver run on an OracleTree (oracle.flat)."""

    def __init__(self, tree):
        self.t = t = tree
        self.c = _Tree(
            num_nodes=t.num_nodes, num_infosets=t.num_infosets, num_levels=len(t.level_off) - 1,
            parent=_p(t.parent, _i32p), first_child=_p(t.first_child, _i32p),
            player=_p(t.player, C.POINTER(C.c_int8)), infoset=_p(t.infoset, _i32p), label=_p(t.label, _i32p),
            hidden=_p(t.hidden, _u8p), cprob=_p(t.cprob, _dp), u1=_p(t.u1, _dp), u2=_p(t.u2, _dp),
            act_off=_p(t.act_off, _i64p), iset_node_off=_p(t.iset_node_off, _i64p),
            iset_nodes=_p(t.iset_nodes, _i32p), level_off=_p(t.level_off, _i64p))
        self.Q = int(t.act_off[-1])
        self.num_labels = int(t.label.max()) + 1
        self.reset()

    def reset(self):
        t = self.t
        self.R = np.zeros(self.Q)
        self.S = np.zeros(self.Q)          # action counts (vanilla / chance) or alpha (external)
        self.sigma = self.uniform()
        self.flags = np.zeros(3 * t.num_infosets, np.uint8)
        self.t_next = 0
        self.nodes_touched = 0
        self.cum_regret = np.zeros(self.Q)

    def uniform(self):
        t = self.t
        return np.repeat(1.0 / t.nact.astype(np.float64), t.nact)

    # flags views
    @property
    def in_R(self):
        return self.flags[:self.t.num_infosets]

    @property
    def in_S(self):
        return self.flags[self.t.num_infosets:2 * self.t.num_infosets]

    @property
    def in_sigma(self):
        return self.flags[2 * self.t.num_infosets:]

    # ---- solvers -------------------------------------------------------------
    def cfr_recursive(self, n_iters, linear_weight=False, sampling=False, lanes=1, rng=None):
        r = lib().orc_cfr_recursive_iters(
            C.byref(self.c), _p(self.R, _dp), _p(self.S, _dp), _p(self.sigma, _dp), _p(self.flags, _u8p),
            C.c_int64(self.t_next), C.c_int64(n_iters), int(linear_weight), int(sampling), C.c_int64(lanes),
            C.byref(rng) if rng is not None else None)
        if r < 0:
            raise RuntimeError('uniform stream exhausted')
        self.t_next += n_iters
        self.nodes_touched += r
        return r

    def vanilla_levels(self, n_iters, linear_weight=False, threads=1):
        r = lib().orc_vanilla_level_iters(
            C.byref(self.c), _p(self.R, _dp), _p(self.S, _dp), _p(self.sigma, _dp),
            C.c_int64(self.t_next), C.c_int64(n_iters), int(linear_weight), int(threads))
        self.t_next += n_iters
        self.nodes_touched += r
        self.flags[:] = 1
        return r

    def external(self, n_iters, rng, lanes=1):
        r = lib().orc_external_iters(
            C.byref(self.c), _p(self.R, _dp), _p(self.S, _dp), _p(self.sigma, _dp), _p(self.flags, _u8p),
            C.c_int64(self.t_next), C.c_int64(n_iters), C.c_int64(lanes), C.byref(rng))
        if r < 0:
            raise RuntimeError('uniform stream exhausted')
        self.t_next += n_iters
        self.nodes_touched += r
        return r

    # ---- read-outs -----------------------------------------------------------
    def average_strategy(self):
        """(avg[Q], present[I]) per cfr.py:32-41."""
        avg = np.zeros(self.Q)
        present = np.zeros(self.t.num_infosets, np.uint8)
        lib().orc_average_strategy(C.byref(self.c), _p(self.S, _dp), _p(avg, _dp), _p(present, _u8p))
        return avg, present

    def alpha_strategy(self):
        """AverageStrategy.compute_strategy() (cfr_util.py:95-106); absent sets come out uniform."""
        out = np.zeros(self.Q)
        if lib().orc_normalise_alpha(C.byref(self.c), _p(self.S, _dp), _p(out, _dp)) != 0:
            raise ValueError('All action floats must be positive.')
        return out

    def best_response(self, sigma, player, want_argmax=False):
        sigma = np.ascontiguousarray(sigma, np.float64)
        arg = np.zeros(self.t.num_infosets, np.int32) if want_argmax else None
        v = lib().orc_best_response(C.byref(self.c), _p(sigma, _dp), int(player), self.num_labels,
                                    _p(arg, _i32p) if want_argmax else None)
        return (v, arg) if want_argmax else v

    def exploitability(self, sigma):
        sigma = np.ascontiguousarray(sigma, np.float64)
        return lib().orc_exploitability(C.byref(self.c), _p(sigma, _dp), self.num_labels)

    def expected_value_exact(self, sigma):
        sigma = np.ascontiguousarray(sigma, np.float64)
        out = np.zeros(2)
        lib().orc_expected_value_exact(C.byref(self.c), _p(sigma, _dp), _p(out, _dp))
        return float(out[0]), float(out[1])

    def immediate_regret_step(self, sigma, step, order=None, detail=False):
        sigma = np.ascontiguousarray(sigma, np.float64)
        scratch = np.zeros(self.Q)
        per_iset = np.zeros(self.t.num_infosets) if detail else None
        v = lib().orc_immediate_regret_detail(
            C.byref(self.c), _p(sigma, _dp), _p(self.cum_regret, _dp), _p(scratch, _dp), C.c_int64(step),
            _p(order, _i32p) if order is not None else None, _p(per_iset, _dp) if detail else None)
        return (v, per_iset) if detail else v

    def node_values(self, sigma):
        """(v1[N], v2[N]) per cfr_metrics.py:159-254."""
        sigma = np.ascontiguousarray(sigma, np.float64)
        v1, v2 = np.zeros(self.t.num_nodes), np.zeros(self.t.num_nodes)
        lib().orc_node_values(C.byref(self.c), _p(sigma, _dp), _p(v1, _dp), _p(v2, _dp))
        return v1, v2


def regret_matching(r, eps=1e-7, highest_regret=False):
    r = np.ascontiguousarray(r, np.float64)
    out = np.zeros_like(r)
    lib().orc_regret_matching(_p(r, _dp), len(r), C.c_double(eps), int(highest_regret), _p(out, _dp))
    return out


def normalise_probs(p, eps=1e-7):
    p = np.ascontiguousarray(p, np.float64)
    out = np.zeros_like(p)
    lib().orc_normalise_probs(_p(p, _dp), len(p), C.c_double(eps), _p(out, _dp))
    return out


def draw(p, u):
    p = np.ascontiguousarray(p, np.float64)
    return lib().orc_draw(_p(p, _dp), len(p), C.c_double(u))


def philox_uniform(seed, iteration, lane, trav, draw_idx):
    return lib().orc_philox_uniform(seed, iteration, lane, trav, draw_idx)


def leduc_tree(num_values, num_suits=2, max_raises=4, raise_amount=2):
    """Leduc(get_deck(num_values, num_suits)) as an OracleTree, built by the oracle's own C generator
    (leduc_oracle.c) -- no product code involved."""
    from oracle import flat
    ni = C.c_int32(0)
    n = lib().orc_leduc_build(int(num_values), int(num_suits), int(max_raises), int(raise_amount), C.byref(ni))
    if n < 0:
        raise ValueError('cannot build Leduc(%d, %d)' % (num_values, num_suits))
    parent, infoset, label = np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.int32)
    first_child = np.empty(n + 1, np.int32)
    player, hidden = np.empty(n, np.int8), np.empty(n, np.uint8)
    cprob, u1, u2 = np.empty(n), np.empty(n), np.empty(n)
    rc = lib().orc_leduc_fetch(_p(parent, _i32p), _p(first_child, _i32p), _p(player, C.POINTER(C.c_int8)),
                               _p(infoset, _i32p), _p(label, _i32p), _p(hidden, _u8p), _p(cprob, _dp), _p(u1, _dp),
                               _p(u2, _dp))
    assert rc == 0
    t = flat.from_arrays(n, parent, first_child, player, infoset, label, hidden, cprob, u1, u2)
    assert t.num_infosets == ni.value
    return t
