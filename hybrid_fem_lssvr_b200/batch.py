"""Device-level batched entry points: thin Python over the C ABI (include/hfl.h).

torch is used for device memory and streams only; every computation is a kernel of libhfl.so
launched on torch's current CUDA stream.  Nothing here has a CPU path.
"""
import ctypes as C
import math

import torch

from . import _lib

_PLANS = {}


class Plan:
    """Element-independent tables for (M, N, F, gamma); see hfl_plan_create."""

    def __init__(self, M, N, F, gamma):
        lib = _lib.load()
        h = C.c_void_p()
        _lib.check(lib.hfl_plan_create(C.byref(h), int(M), int(N), int(F), float(gamma)), 'hfl_plan_create')
        self.handle = h
        self.M, self.N, self.F, self.gamma = int(M), int(N), int(F), float(gamma)

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                _lib.load().hfl_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def get_plan(M, N, F, gamma, device=None):
    """Plan for (M, N, F, gamma) on `device` (default: the current device); its tables live on that device."""
    index = torch.cuda.current_device() if device is None else _index(device)
    key = (int(M), int(N), int(F), float(gamma), index)
    p = _PLANS.get(key)
    if p is None:
        with torch.cuda.device(index):
            p = _PLANS[key] = Plan(M, N, F, gamma)
    return p


def _index(device):
    device = torch.device(device)
    return torch.cuda.current_device() if device.index is None else device.index


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _on_device_of(arg=0):
    """Run the wrapped entry point with the device of its tensor argument current: plan tables, scratch space and the
    stream then belong to the device the data live on, whatever device the caller has selected."""
    import functools

    def deco(fn):
        @functools.wraps(fn)
        def wrapper(*args, **kwargs):
            t = args[arg]
            if isinstance(t, torch.Tensor) and t.is_cuda and t.device.index != torch.cuda.current_device():
                with torch.cuda.device(t.device):
                    return fn(*args, **kwargs)
            return fn(*args, **kwargs)
        return wrapper
    return deco


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _require_cuda_f64(t, name, numel=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.float64 or not t.is_contiguous():
        raise TypeError('%s must be a contiguous float64 CUDA tensor' % name)
    if numel is not None and t.numel() != numel:
        raise ValueError('%s has %d entries, expected %d' % (name, t.numel(), numel))


def mesh_linspace(a, b, n_global, i0=0, n_local=None, device='cuda'):
    """numpy.linspace(a, b, n_global)[i0:i0+n_local] generated on the device, bit for bit (P:120)."""
    n_local = n_global - i0 if n_local is None else n_local
    with torch.cuda.device(_index(device)):
        out = torch.empty(n_local, dtype=torch.float64, device='cuda')
        _lib.check(_lib.load().hfl_mesh_linspace(float(a), float(b), int(n_global), int(i0), int(n_local),
                                                 _ptr(out), _stream()), 'hfl_mesh_linspace')
    return out


_WORKSPACES = {}


def _workspace(nbytes, device):
    # one scratch buffer per (device, stream): calls on the same stream are ordered, calls on different
    # streams must not share scratch space
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream)
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = _WORKSPACES[key] = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
    return ws


# 'assembled': the reference's rounded tridiagonal system (partition + PCR); 'assembled_exact': the same kernels on the
# unrounded diagonal (zero row sums); 'flux': first-order form by prefix sums.  See include/hfl.h.
COARSE_MODES = {'assembled': _lib.COARSE_ASSEMBLED_PCR, 'pcr': _lib.COARSE_ASSEMBLED_PCR,
                'flux': _lib.COARSE_FLUX_SCAN, 'assembled_exact': _lib.COARSE_ASSEMBLED_EXACT}


@_on_device_of(0)
def fem_p1_solve(nodes, k_freq=1.0, u_left=0.0, u_right=0.0, coarse_solver='assembled', out=None,
                 want_reaction=False):
    """K1: nodal values of the coarse P1 FEM solve (P:117-145) for -u'' = (k pi)^2 sin(k pi x).

    Returns u (and the interface record {x_first, x_last, r_left, r_right} as a 4-vector tensor when
    want_reaction).
    """
    _require_cuda_f64(nodes, 'nodes')
    n = nodes.numel()
    lib = _lib.load()
    mode = COARSE_MODES[coarse_solver]
    u = out if out is not None else torch.empty(n, dtype=torch.float64, device=nodes.device)
    _require_cuda_f64(u, 'out', n)
    nbytes = int(lib.hfl_fem_p1_workspace_bytes(n))
    ws = _workspace(nbytes, nodes.device)
    react = torch.empty(4, dtype=torch.float64, device=nodes.device) if want_reaction else None
    _lib.check(lib.hfl_fem_p1_solve(n, _ptr(nodes), float(k_freq), float(u_left), float(u_right), mode,
                                    _ptr(u), _ptr(react), _ptr(ws), ws.numel(), _stream()), 'hfl_fem_p1_solve')
    return (u, react) if want_reaction else u


@_on_device_of(0)
def fem_p1_solve_multi(nodes, k_freqs, u_left=0.0, u_right=0.0, coarse_solver='assembled', out=None):
    """K1 for R forcing frequencies on one mesh in a single launch sequence: u [R, n], row r bit-identical to
    fem_p1_solve(nodes, k_freq=k_freqs[r]).  k_freqs: float64 device tensor [R]."""
    _require_cuda_f64(nodes, 'nodes')
    n = nodes.numel()
    R = k_freqs.numel()
    _require_cuda_f64(k_freqs, 'k_freqs', R)
    lib = _lib.load()
    mode = COARSE_MODES[coarse_solver]
    u = out if out is not None else torch.empty((R, n), dtype=torch.float64, device=nodes.device)
    _require_cuda_f64(u, 'out', R * n)
    ws = _workspace(int(lib.hfl_fem_p1_multi_workspace_bytes(n, R)), nodes.device)
    _lib.check(lib.hfl_fem_p1_solve_multi(n, _ptr(nodes), R, _ptr(k_freqs), float(u_left), float(u_right), mode,
                                          _ptr(u), _ptr(ws), ws.numel(), _stream()), 'hfl_fem_p1_solve_multi')
    return u


GAUSS_X = (0.5 - 0.5 / math.sqrt(3.0), 0.5 + 0.5 / math.sqrt(3.0))   # 2-point Gauss abscissae on [0, 1]


@_on_device_of(0)
def fem_p1_solve_general(nodes, aq, fq, cq=None, u_left=0.0, u_right=0.0, out=None):
    """Coarse P1 solve of -(a u')' + c u = f; aq, cq, fq are [2, E] samples at the Gauss points
    x_e + h_e * GAUSS_X[q] of every element."""
    _require_cuda_f64(nodes, 'nodes')
    n = nodes.numel()
    for name, t in (('aq', aq), ('fq', fq), ('cq', cq)):
        if t is not None:
            _require_cuda_f64(t, name, 2 * (n - 1))
    lib = _lib.load()
    u = out if out is not None else torch.empty(n, dtype=torch.float64, device=nodes.device)
    ws = _workspace(int(lib.hfl_fem_p1_workspace_bytes(n)), nodes.device)
    _lib.check(lib.hfl_fem_p1_solve_general(n, _ptr(nodes), _ptr(aq), _ptr(cq), _ptr(fq), float(u_left), float(u_right),
                                            _ptr(u), _ptr(ws), ws.numel(), _stream()), 'hfl_fem_p1_solve_general')
    return u


@_on_device_of(0)
def fem_apply_bc(nodes, u, bc_left, bc_right):
    _require_cuda_f64(nodes, 'nodes')
    _require_cuda_f64(u, 'u', nodes.numel())
    _lib.check(_lib.load().hfl_fem_apply_bc(nodes.numel(), _ptr(nodes), _ptr(u), float(bc_left), float(bc_right),
                                            _stream()), 'hfl_fem_apply_bc')
    return u


def spike_interface_solve(gathered, u_left=0.0, u_right=0.0):
    """Host: interface values of G contiguous ranges from their {x_first, x_last, r_left, r_right}."""
    G = len(gathered) // 4
    arr = (C.c_double * (4 * G))(*[float(v) for v in gathered])
    out = (C.c_double * (G + 1))()
    _lib.check(_lib.load().hfl_spike_interface_solve(G, arr, float(u_left), float(u_right), out),
               'hfl_spike_interface_solve')
    return list(out)


@_on_device_of(1)
def _element_batch(fn_name, nodes, u, M, gamma, N, F, forcing, k_freq, bc2, want_coef, want_fine, want_status,
                   err3, coef_out, fine_out):
    _require_cuda_f64(nodes, 'nodes')
    E = nodes.numel() - 1
    _require_cuda_f64(u, 'u', E + 1)
    lib = _lib.load()
    plan = get_plan(M, N, F if (want_fine or err3 is not None) else 0, gamma)
    dev = nodes.device
    if isinstance(forcing, torch.Tensor):
        _require_cuda_f64(forcing, 'forcing samples', N * E)
        kind, fs = _lib.FORCING_SAMPLES, forcing
    elif forcing == 'sine':
        kind, fs = _lib.FORCING_SINE, None
    else:
        raise ValueError("forcing must be 'sine' or a [N, E] CUDA tensor of samples")
    coef = None
    if want_coef:
        coef = coef_out if coef_out is not None else torch.empty((E, M), dtype=torch.float64, device=dev)
        _require_cuda_f64(coef, 'coef_out', E * M)
    fine = None
    if want_fine:
        fine = fine_out if fine_out is not None else torch.empty((E, F), dtype=torch.float64, device=dev)
        _require_cuda_f64(fine, 'fine_out', E * F)
    status = torch.empty(E, dtype=torch.int32, device=dev) if want_status else None
    if bc2 is not None:
        _require_cuda_f64(bc2, 'bc2', 2)
    if err3 is not None:
        _require_cuda_f64(err3, 'err3', 3)
    _lib.check(getattr(lib, fn_name)(plan.handle, E, _ptr(nodes), _ptr(u), kind, float(k_freq), _ptr(fs), _ptr(bc2),
                                     _ptr(coef), _ptr(fine), _ptr(status), _ptr(err3), _stream()), fn_name)
    return coef, fine, status


def lssvr_primal_batch(nodes, u, M, gamma, N=12, F=0, forcing='sine', k_freq=1.0, bc2=None, want_coef=True,
                       want_fine=False, want_status=False, err3=None, coef_out=None, fine_out=None):
    """K2 + K3 (+ fused K5): every element's primal LSSVR solve (P:147-176 over P:20-105).

    Returns (coef [E, M] | None, fine [E, F] | None, status [E] int32 | None).
    """
    return _element_batch('hfl_lssvr_primal_batch', nodes, u, M, gamma, N, F, forcing, k_freq, bc2, want_coef,
                          want_fine, want_status, err3, coef_out, fine_out)


def lssvr_dual_batch(nodes, u, M, gamma, N=12, F=0, forcing='sine', k_freq=1.0, bc2=None, want_coef=True,
                     want_fine=False, want_status=False, err3=None, coef_out=None, fine_out=None):
    """K4: the same solves through the (N+2) x (N+2) dual (kernel) system."""
    return _element_batch('hfl_lssvr_dual_batch', nodes, u, M, gamma, N, F, forcing, k_freq, bc2, want_coef,
                          want_fine, want_status, err3, coef_out, fine_out)


@_on_device_of(0)
def lssvr_dual_multi(nodes, u, k_freqs, M, gamma, N=12, F=0, forcing='sine', bc2=None, want_coef=True,
                     want_fine=False, want_status=False, err3=None):
    """K4 with R right-hand sides per element sharing one factorisation (BASELINE configs[4]).

    u [R, E+1], k_freqs [R] (CUDA tensor), forcing 'sine' or samples [R, N, E]; returns
    (coef [R, E, M] | None, fine [R, E, F] | None, status [E] | None); err3, if given, is [R, 3].
    """
    _require_cuda_f64(nodes, 'nodes')
    E = nodes.numel() - 1
    _require_cuda_f64(u, 'u')
    R = u.shape[0]
    if u.shape != (R, E + 1):
        raise ValueError('u must be [R, E + 1]')
    _require_cuda_f64(k_freqs, 'k_freqs', R)
    dev = nodes.device
    plan = get_plan(M, N, F if (want_fine or err3 is not None) else 0, gamma)
    if isinstance(forcing, torch.Tensor):
        _require_cuda_f64(forcing, 'forcing samples', R * N * E)
        kind, fs = _lib.FORCING_SAMPLES, forcing
    else:
        kind, fs = _lib.FORCING_SINE, None
    coef = torch.empty((R, E, M), dtype=torch.float64, device=dev) if want_coef else None
    fine = torch.empty((R, E, F), dtype=torch.float64, device=dev) if want_fine else None
    status = torch.empty(E, dtype=torch.int32, device=dev) if want_status else None
    if err3 is not None:
        _require_cuda_f64(err3, 'err3', 3 * R)
    _lib.check(_lib.load().hfl_lssvr_dual_multi(plan.handle, E, R, _ptr(nodes), _ptr(u), kind, _ptr(k_freqs), _ptr(fs),
                                                _ptr(bc2), _ptr(coef), _ptr(fine), _ptr(status), _ptr(err3), _stream()),
               'hfl_lssvr_dual_multi')
    return coef, fine, status


@_on_device_of(0)
def lssvr_general_batch(nodes, u, a, f, M, gamma, N=12, F=0, da=None, c=None, bc2=None, want_coef=True,
                        want_fine=False, want_status=False):
    """Per-element LSSVR for -(a u')' + c u = f; a, da (= a'), c, f are [N, E] CUDA tensors of samples at the
    collocation points (da / c may be None).  Returns (coef, fine, status) like lssvr_primal_batch."""
    _require_cuda_f64(nodes, 'nodes')
    E = nodes.numel() - 1
    _require_cuda_f64(u, 'u', E + 1)
    for name, t in (('a', a), ('f', f), ('da', da), ('c', c)):
        if t is not None:
            _require_cuda_f64(t, name, N * E)
    dev = nodes.device
    plan = get_plan(M, N, F if want_fine else 0, gamma)
    coef = torch.empty((E, M), dtype=torch.float64, device=dev) if want_coef else None
    fine = torch.empty((E, F), dtype=torch.float64, device=dev) if want_fine else None
    status = torch.empty(E, dtype=torch.int32, device=dev) if want_status else None
    _lib.check(_lib.load().hfl_lssvr_general_batch(plan.handle, E, _ptr(nodes), _ptr(u), _ptr(a), _ptr(da), _ptr(c), _ptr(f),
                                                   _ptr(bc2), _ptr(coef), _ptr(fine), _ptr(status), _stream()),
               'hfl_lssvr_general_batch')
    return coef, fine, status


@_on_device_of(0)
def evaluate_points(nodes, coef, x):
    """K3 unstructured: evaluate_solution's element search + legval on the device (P:184-211)."""
    _require_cuda_f64(nodes, 'nodes')
    E = nodes.numel() - 1
    _require_cuda_f64(coef, 'coef')
    M = coef.shape[1]
    if coef.shape[0] != E:
        raise ValueError('coef has %d rows, expected %d' % (coef.shape[0], E))
    _require_cuda_f64(x, 'x')
    out = torch.empty_like(x)
    _lib.check(_lib.load().hfl_evaluate_points(E, _ptr(nodes), M, _ptr(coef), x.numel(), _ptr(x), _ptr(out),
                                               _stream()), 'hfl_evaluate_points')
    return out


def new_error_accumulator(device='cuda'):
    return torch.zeros(3, dtype=torch.float64, device=device)


@_on_device_of(0)
def error_fine(nodes, fine, k_freq=1.0, err3=None):
    _require_cuda_f64(nodes, 'nodes')
    E = nodes.numel() - 1
    _require_cuda_f64(fine, 'fine')
    F = fine.shape[1]
    err3 = new_error_accumulator(nodes.device) if err3 is None else err3
    _lib.check(_lib.load().hfl_error_fine(E, F, _ptr(nodes), _ptr(fine), float(k_freq), _ptr(err3), _stream()),
               'hfl_error_fine')
    return err3


@_on_device_of(0)
def error_nodal(nodes, u, k_freq=1.0, err3=None):
    _require_cuda_f64(nodes, 'nodes')
    _require_cuda_f64(u, 'u', nodes.numel())
    err3 = new_error_accumulator(nodes.device) if err3 is None else err3
    _lib.check(_lib.load().hfl_error_nodal(nodes.numel(), _ptr(nodes), _ptr(u), float(k_freq), _ptr(err3), _stream()),
               'hfl_error_nodal')
    return err3


def finish_error(err3):
    """(L2, max) from an accumulator: sqrt of the weighted sum of squares, and the max."""
    v = err3.tolist()
    return math.sqrt(v[0]), v[1]


def set_option(key, value):
    _lib.check(_lib.load().hfl_set_option(key.encode(), int(value)), 'hfl_set_option')
