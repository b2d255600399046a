"""ctypes binding of librlpoker_b200.so (C ABI: include/rlpoker_b200.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is usable, every
solver entry point raises.  Build the library with ``python -c 'import __graft_entry__ as g; g.build()'``
or ``make -C rlpoker_b200/csrc``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'librlpoker_b200.so')

RLP_REGRET, RLP_STRAT_SUM, RLP_SIGMA, RLP_AVERAGE, RLP_CUM_IMM_REGRET = range(5)
RLP_CHANCE_SAMPLED, RLP_EXTERNAL_SAMPLED = 0, 1

_dp = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_i8p = C.POINTER(C.c_int8)
_u8p = C.POINTER(C.c_uint8)
_vp = C.c_void_p


class TreeDesc(C.Structure):
    _fields_ = [
        ('num_nodes', C.c_int64), ('num_infosets', C.c_int32), ('num_levels', C.c_int32),
        ('first_child', _i32p), ('player', _i8p), ('infoset', _i32p), ('cprob', _dp), ('u1', _dp), ('u2', _dp),
        ('act_off', _i64p), ('level_off', _i64p),
    ]


class RlpError(RuntimeError):
    pass


# every exported symbol of include/rlpoker_b200.h: name -> (restype, argtypes)
SYMBOLS = {
    'rlp_last_error': (C.c_char_p, []),
    'rlp_version': (C.c_int, []),
    'rlp_device_count': (C.c_int, []),
    'rlp_kernel_launches': (C.c_uint64, []),
    'rlp_tree_upload': (C.c_int, [C.POINTER(TreeDesc), _vp, C.POINTER(_vp)]),
    'rlp_tree_free': (None, [_vp]),
    'rlp_tree_bytes': (C.c_int64, [_vp]),
    'rlp_tree_stat': (C.c_int64, [_vp, C.c_int]),
    'rlp_tree_plan': (C.c_int, [C.POINTER(TreeDesc), _i64p]),
    'rlp_tree_plan_fused_source': (C.c_int, [C.POINTER(TreeDesc), _i64p, C.c_char_p, C.c_int64, _i64p]),
    'rlp_tree_plan_lane_records': (C.c_int, [C.POINTER(TreeDesc), C.c_int, _i64p]),
    'rlp_cfr_fused_timeline': (C.c_int, [_vp, C.c_int64, C.c_int, C.POINTER(C.c_uint64), _vp]),
    'rlp_leduc_generate': (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, _i64p, _i32p, _i32p, _i8p, _i32p,
                                     _dp, _dp, _dp, _i32p, _u8p]),
    'rlp_cfr_create': (C.c_int, [_vp, _vp, C.POINTER(_vp)]),
    'rlp_cfr_free': (None, [_vp]),
    'rlp_cfr_reset': (C.c_int, [_vp, _vp]),
    'rlp_cfr_vanilla_iters': (C.c_int, [_vp, C.c_int64, C.c_int64, C.c_int, _vp]),
    'rlp_cfr_sampled_iters': (C.c_int, [_vp, C.c_int, C.c_int64, C.c_uint64, C.c_int64, C.c_int64, C.c_int, _vp]),
    'rlp_cfr_sampled_local': (C.c_int, [_vp, C.c_int, C.c_int64, C.c_int64, C.c_uint64, C.c_int64, C.c_int, _vp]),
    'rlp_cfr_sampled_finish': (C.c_int, [_vp, C.c_int, _vp]),
    'rlp_cfr_flags_ptr': (_vp, [_vp]),
    'rlp_cfr_update_average': (C.c_int, [_vp, C.c_double, _vp]),
    'rlp_cfr_read': (C.c_int, [_vp, C.c_int, _dp, _vp]),
    'rlp_cfr_write': (C.c_int, [_vp, C.c_int, _dp, _vp]),
    'rlp_cfr_read_flags': (C.c_int, [_vp, _u8p, _vp]),
    'rlp_cfr_profile_stages': (C.c_int, [_vp, C.c_int, C.POINTER(C.c_float), _vp]),
    'rlp_cfr_timeline': (C.c_int, [_vp, C.c_int, C.POINTER(C.c_uint64), C.c_int, C.POINTER(C.c_int), _vp]),
    'rlp_cfr_write_flags': (C.c_int, [_vp, _u8p, _vp]),
    'rlp_nodes_touched': (C.c_uint64, [_vp]),
    'rlp_cfr_exploitability': (C.c_int, [_vp, C.c_int, _dp, _vp]),
    'rlp_cfr_immediate_regret_step': (C.c_int, [_vp, C.c_int64, _i32p, _dp, _vp]),
    'rlp_cfr_immediate_regret_detail': (C.c_int, [_vp, C.c_int64, _i32p, _dp, _dp, _vp]),
    'rlp_node_values': (C.c_int, [_vp, _dp, _dp, _dp, _vp]),
    'rlp_best_response': (C.c_int, [_vp, _dp, C.c_int, _dp, _i32p, _vp]),
    'rlp_expected_value': (C.c_int, [_vp, _dp, _dp, _vp]),
    'rlp_cfr_p2p_export': (C.c_int, [_vp, _u8p]),
    'rlp_cfr_p2p_connect': (C.c_int, [_vp, C.c_int, C.c_int, _u8p]),
    'rlp_cfr_p2p_owned': (C.c_int, [_vp, _u8p]),
    'rlp_cfr_p2p_failed': (C.c_int, [_vp]),
    'rlp_cfr_local_sweep': (C.c_int, [_vp, C.c_int64, C.c_int, _vp]),
    'rlp_cfr_delta_ptr': (_vp, [_vp]),
    'rlp_comm_unique_id': (C.c_int, [_u8p]),
    'rlp_comm_init': (C.c_int, [_vp, C.c_int, C.c_int, _u8p]),
    'rlp_comm_destroy': (None, [_vp]),
    'rlp_cfr_sharded_iters': (C.c_int, [_vp, C.c_int64, C.c_int64, C.c_int, _vp]),
    'rlp_cfr_apply_delta': (C.c_int, [_vp, _vp]),
}

_lib = None


def load():
    """The loaded library (raises RlpError when it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RlpError("%s is missing: build it with __graft_entry__.build() (there is no CPU fallback)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise RlpError("rlpoker_b200: %s (status %d)" % (load().rlp_last_error().decode(), rc))


def require_device():
    n = load().rlp_device_count()
    if n <= 0:
        raise RlpError("rlpoker_b200 needs a CUDA device (found none: %s); there is no CPU fallback"
                       % load().rlp_last_error().decode())
    return n


def ptr(a, t):
    return a.ctypes.data_as(t)
