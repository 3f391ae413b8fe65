"""ctypes binding of libhfl.so (the C ABI declared in include/hfl.h).

There is no fallback: if the shared object is missing or a call fails, an exception is raised.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('HFL_LIB') or os.path.join(HERE, 'libhfl.so')   # HFL_LIB: an A/B build of the same ABI

HFL_OK = 0
FORCING_SINE = 0
FORCING_SAMPLES = 1
COARSE_ASSEMBLED_PCR = 0
COARSE_FLUX_SCAN = 1
COARSE_ASSEMBLED_EXACT = 2
MAX_M, MAX_N, MAX_F = 32, 256, 256


class HflError(RuntimeError):
    pass


_vp, _i32, _i64, _f64, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_size_t

# name -> (restype, argtypes); kept in the order of include/hfl.h (tests/test_abi.py cross-checks)
SIGNATURES = {
    'hfl_version': (C.c_char_p, []),
    'hfl_last_error': (C.c_char_p, []),
    'hfl_device_info': (_i32, [C.POINTER(_i32)] * 3),
    'hfl_plan_create': (_i32, [C.POINTER(_vp), _i32, _i32, _i32, _f64]),
    'hfl_plan_destroy': (_i32, [_vp]),
    'hfl_mesh_linspace': (_i32, [_f64, _f64, _i64, _i64, _i64, _vp, _vp]),
    'hfl_fem_p1_workspace_bytes': (_sz, [_i64]),
    'hfl_fem_p1_solve': (_i32, [_i64, _vp, _f64, _f64, _f64, _i32, _vp, _vp, _vp, _sz, _vp]),
    'hfl_fem_p1_multi_workspace_bytes': (_sz, [_i64, _i32]),
    'hfl_fem_p1_solve_multi': (_i32, [_i64, _vp, _i32, _vp, _f64, _f64, _i32, _vp, _vp, _sz, _vp]),
    'hfl_fem_p1_solve_general': (_i32, [_i64, _vp, _vp, _vp, _vp, _f64, _f64, _vp, _vp, _sz, _vp]),
    'hfl_spike_interface_solve': (_i32, [_i32, C.POINTER(_f64), _f64, _f64, C.POINTER(_f64)]),
    'hfl_spike_interface_solve_device': (_i32, [_i32, _vp, _f64, _f64, _i32, _vp, _vp]),
    'hfl_peer_buffer_bytes': (_sz, []),
    'hfl_peer_buffer_create': (_i32, [C.POINTER(_vp), C.POINTER(C.c_ubyte)]),
    'hfl_peer_buffer_open': (_i32, [C.POINTER(C.c_ubyte), C.POINTER(_vp)]),
    'hfl_peer_buffer_close': (_i32, [_vp]),
    'hfl_peer_buffer_destroy': (_i32, [_vp]),
    'hfl_peer_allgather': (_i32, [_i32, _i32, _i32, _vp, _vp, C.c_uint32, _i32, _vp, _vp, _vp]),
    'hfl_peer_spike_exchange': (_i32, [_i32, _i32, _vp, _vp, C.c_uint32, _i32, _f64, _f64, _vp, _vp, _vp, _vp]),
    'hfl_fem_apply_bc': (_i32, [_i64, _vp, _vp, _f64, _f64, _vp]),
    'hfl_lssvr_primal_batch': (_i32, [_vp, _i64, _vp, _vp, _i32, _f64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'hfl_lssvr_dual_batch': (_i32, [_vp, _i64, _vp, _vp, _i32, _f64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'hfl_lssvr_dual_multi': (_i32, [_vp, _i64, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'hfl_lssvr_general_batch': (_i32, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'hfl_evaluate_points': (_i32, [_i64, _vp, _i32, _vp, _i64, _vp, _vp, _vp]),
    'hfl_error_fine': (_i32, [_i64, _i32, _vp, _vp, _f64, _vp, _vp]),
    'hfl_error_nodal': (_i32, [_i64, _vp, _vp, _f64, _vp, _vp]),
    'hfl_set_option': (_i32, [C.c_char_p, _i32]),
    'hfl_get_option': (_i32, [C.c_char_p, C.POINTER(_i32)]),
    'hfl_launch_count': (_i64, []),
    'hfl_store_probe': (_i32, [_i32, _i32, _i64, _i32, _i32, _vp, _vp]),
    'hfl_fp64_probe': (_i32, [_i32, _i32, _vp, C.POINTER(_f64), _vp]),
}

_lib = None


def load():
    """Load libhfl.so once; raise HflError with build instructions when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HflError('%s not found: build it with `python -m hybrid_fem_lssvr_b200.build` '
                       '(there is no CPU fallback)' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != HFL_OK:
        msg = load().hfl_last_error().decode('utf-8', 'replace')
        raise HflError('%s failed (code %d): %s' % (what, rc, msg))


def launch_count():
    return int(load().hfl_launch_count())
