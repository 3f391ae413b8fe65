"""The C-ABI library loads on a CPU-only box and exports every symbol include/hfl.h declares.
No compute entry point is called here (they need a GPU); argument validation and the host-only
interface solve are exercised because they run without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from hybrid_fem_lssvr_b200 import _lib

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))


def _declared_functions():
    src = open(os.path.join(ROOT, 'include', 'hfl.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(hfl_[a-z0-9_]+)\s*\(', src)))


@pytest.fixture(scope='module')
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        from hybrid_fem_lssvr_b200 import build
        build.build()
    return _lib.load()


def test_header_and_binding_agree(lib):
    declared = _declared_functions()
    assert len(declared) >= 19
    assert sorted(_lib.SIGNATURES) == declared
    for name in declared:
        assert hasattr(lib, name), name


def test_version_and_error_paths(lib):
    assert b'sm_100a' in lib.hfl_version()
    h = C.c_void_p()
    assert lib.hfl_plan_create(C.byref(h), 2, 12, 32, 1e4) == 1           # M too small
    assert b'M=2' in lib.hfl_last_error()
    assert lib.hfl_plan_create(C.byref(h), 9, 12, 32, -1.0) == 1          # gamma <= 0
    assert lib.hfl_plan_create(C.byref(h), 9, 1000, 32, 1e4) == 1         # N too large
    assert lib.hfl_lssvr_primal_batch(None, 10, None, None, 0, 1.0, None, None, None, None, None, None, None) == 1
    assert lib.hfl_fem_p1_solve(1, None, 1.0, 0.0, 0.0, 0, None, None, None, 0, None) == 1
    assert lib.hfl_set_option(b'no_such_option', 1) == 1
    assert lib.hfl_set_option(b'primal_store', 2) == 0
    v = C.c_int(-1)
    assert lib.hfl_get_option(b'primal_store', C.byref(v)) == 0 and v.value == 2
    assert lib.hfl_set_option(b'primal_store', 0) == 0
    assert lib.hfl_fem_p1_workspace_bytes(10_000_001) >= (17 + 768) * 4883 * 8


def test_interface_solve_host(lib):
    """Interface values of 4 ranges against the coarse P1 system assembled with numpy."""
    rng = np.random.default_rng(0)
    xs = np.sort(np.concatenate([[-1.0, 1.0], rng.uniform(-1, 1, 3)]))
    G = 4
    rl, rr = rng.normal(size=G), rng.normal(size=G)
    gathered = np.empty(4 * G)
    for r in range(G):
        gathered[4 * r:4 * r + 4] = [xs[r], xs[r + 1], rl[r], rr[r]]
    out = (C.c_double * (G + 1))()
    arr = (C.c_double * (4 * G))(*gathered)
    assert lib.hfl_spike_interface_solve(G, arr, 0.25, -0.5, out) == 0
    L = np.diff(xs)
    A = np.zeros((G - 1, G - 1))
    b = np.zeros(G - 1)
    for r in range(1, G):
        A[r - 1, r - 1] = 1 / L[r - 1] + 1 / L[r]
        if r > 1:
            A[r - 1, r - 2] = -1 / L[r - 1]
        if r < G - 1:
            A[r - 1, r] = -1 / L[r]
        b[r - 1] = rr[r - 1] + rl[r]
    b[0] += 0.25 / L[0]
    b[-1] += -0.5 / L[-1]
    expect = np.concatenate([[0.25], np.linalg.solve(A, b), [-0.5]])
    assert np.max(np.abs(np.array(list(out)) - expect)) <= 1e-13


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', str(tmp_path / 'libhfl.so'))
    with pytest.raises(_lib.HflError):
        _lib.load()
