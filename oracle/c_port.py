"""ctypes loader of the C restatement (oracle/c/hfl_oracle.c).  TEST INFRASTRUCTURE / CPU baseline."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, 'c', 'libhfl_oracle.so')
_lib = None


def load(build=True):
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB) and build:
        subprocess.run(['make', '-C', os.path.join(HERE, 'c')], check=True, capture_output=True)
    lib = C.CDLL(LIB)
    dp = np.ctypeslib.ndpointer(dtype=np.float64, flags='C_CONTIGUOUS')
    lib.oracle_threads.restype = C.c_int
    lib.oracle_set_threads.argtypes = [C.c_int]
    lib.oracle_fem_p1.restype = C.c_int
    lib.oracle_fem_p1.argtypes = [C.c_long, dp, C.c_double, dp]
    lib.oracle_fem_p1_quad.restype = C.c_int
    lib.oracle_fem_p1_quad.argtypes = [C.c_long, dp, C.c_double, C.c_int, C.c_double, C.c_double, dp]
    lib.oracle_primal_batch.restype = C.c_double
    lib.oracle_primal_batch.argtypes = [C.c_long, dp, dp, C.c_int, C.c_double, C.c_int, C.c_double, C.c_int,
                                        C.c_void_p, C.c_void_p]
    _lib = lib
    return lib


def threads():
    return int(load().oracle_threads())


def set_threads(n):
    load().oracle_set_threads(int(n))


def fem_p1(nodes, k_freq=1.0):
    nodes = np.ascontiguousarray(nodes, dtype=np.float64)
    u = np.empty_like(nodes)
    assert load().oracle_fem_p1(nodes.size, nodes, float(k_freq), u) == 0
    return u


def fem_p1_quad(nodes, k_freq=1.0, exact_rowsum=False, u_left=0.0, u_right=0.0):
    """Nodal values of the reference's rounded coarse system (oracle/fem_p1.assemble_p1's entries, Dirichlet rows as
    identity rows) solved in IEEE binary128: its exact solution to double precision (P:117-145)."""
    nodes = np.ascontiguousarray(nodes, dtype=np.float64)
    u = np.empty_like(nodes)
    rc = load().oracle_fem_p1_quad(nodes.size, nodes, float(k_freq), int(bool(exact_rowsum)), float(u_left), float(u_right), u)
    assert rc == 0, rc
    return u


def primal_batch(nodes, u, M, gamma, N=12, k_freq=1.0, F=32, want_coef=True, want_fine=True, fine_out=None):
    """Returns (coef [E, M] | None, fine [E, F] | None, max |u - sin(k pi x)| on the fine grid)."""
    nodes = np.ascontiguousarray(nodes, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    E = nodes.size - 1
    coef = np.empty((E, M)) if want_coef else None
    fine = fine_out if fine_out is not None else (np.empty((E, F)) if want_fine else None)
    mx = load().oracle_primal_batch(E, nodes, u, M, float(gamma), N, float(k_freq), F,
                                    coef.ctypes.data if coef is not None else None,
                                    fine.ctypes.data if fine is not None else None)
    return coef, fine, float(mx)
