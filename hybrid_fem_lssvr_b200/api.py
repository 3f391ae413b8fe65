"""Drop-in surface of the reference script (P: = 1D-Possion/Hybrid-FEM-LSSVR.py).

Same names, argument meaning and return types as the reference:

* ``true_solution``, ``poisson_rhs``, ``main_boundary_condition_left/right``      P:8-18
* ``lssvr_primal(rhs_func, domain_range, u_xmin, u_xmax, M, gamma, ...)``         P:20-105
* ``FEMLSSVRPrimalSolver(num_fem_nodes, lssvr_M, lssvr_gamma, global_domain)``    P:107-211
  with ``solve_fem``, ``solve_lssvr_subproblems``, ``solve``, ``evaluate_solution`` and the public
  attributes ``fem_nodes``, ``fem_values``, ``lssvr_functions``.

The arithmetic runs in libhfl.so on the GPU (no CPU fallback): host buffers go in, host numpy
objects come out, as in the reference.  Device-resident callers use ``hybrid_fem_lssvr_b200.batch``.
"""
import numpy as np
import torch
from numpy.polynomial.legendre import Legendre

from . import batch

N_COLLOCATION = 12   # P:40


def true_solution(x):
    return np.sin(np.pi * x)


def poisson_rhs(x):
    return np.pi ** 2 * np.sin(np.pi * x)


def main_boundary_condition_left(x):
    return 0.0  # u(-1) = 0


def main_boundary_condition_right(x):
    return 0.0  # u(1) = 0


def _sample_rhs(rhs_func, pts):
    """rhs_func on an array of points; falls back to a scalar loop for non-vectorised callables."""
    try:
        vals = np.asarray(rhs_func(pts), dtype=np.float64)
        if vals.shape == pts.shape:
            return vals
        if vals.ndim == 0:
            return np.full(pts.shape, float(vals))
    except Exception:
        pass
    return np.array([float(rhs_func(float(p))) for p in pts.ravel()], dtype=np.float64).reshape(pts.shape)


def _collocation_points(nodes, N):
    """[N, E] array whose column e is np.linspace(x_e, x_{e+1}, N) (P:40)."""
    return np.linspace(nodes[:-1], nodes[1:], N, axis=0)


def lssvr_primal(rhs_func, domain_range, u_xmin, u_xmax, M, gamma,
                 is_left_boundary=False, is_right_boundary=False,
                 global_domain_range=(-1, 1), *, n_colloc=N_COLLOCATION):
    """LSSVR primal solve of one element; returns ``numpy.polynomial.Legendre`` (P:20-105).

    M >= 3 (at least one coefficient beyond the two that the boundary conditions fix; the reference accepts M = 2, where
    the QP has the linear interpolant as its only feasible point)."""
    if M < 3:
        raise ValueError('lssvr_primal: M = %d; this implementation needs M >= 3 (M = 2 leaves only the linear '
                         'interpolant of the two boundary values)' % M)
    xmin, xmax = domain_range
    global_xmin, global_xmax = global_domain_range
    # boundary-flag branches P:68-69 / P:75-76
    g_left = main_boundary_condition_left(global_xmin) if (is_left_boundary and xmin == global_xmin) else u_xmin
    g_right = main_boundary_condition_right(global_xmax) if (is_right_boundary and xmax == global_xmax) else u_xmax
    nodes_h = np.array([xmin, xmax], dtype=np.float64)
    f_h = _sample_rhs(rhs_func, _collocation_points(nodes_h, n_colloc))
    dev = torch.device('cuda', torch.cuda.current_device())
    nodes = torch.from_numpy(nodes_h).to(dev)
    u = torch.tensor([float(g_left), float(g_right)], dtype=torch.float64, device=dev)
    f = torch.from_numpy(np.ascontiguousarray(f_h)).to(dev)
    coef, _, status = batch.lssvr_primal_batch(nodes, u, M, gamma, N=n_colloc, forcing=f, want_status=True)
    if int(status[0].item()) != 0:   # mirrors the warning of P:93-95
        print('Warning: Optimization may not have converged: factorisation broke down; linear fallback used')
    return Legendre(coef[0].cpu().numpy(), domain_range)


class _LegendreList:
    """List-like view of the per-element solutions that builds Legendre objects on demand."""

    def __init__(self, coef, nodes):
        self._coef, self._nodes = coef, nodes

    def __len__(self):
        return self._coef.shape[0]

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        n = len(self)
        if i < 0:
            i += n
        if not 0 <= i < n:
            raise IndexError(i)
        return Legendre(self._coef[i], [self._nodes[i], self._nodes[i + 1]])

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class P1Basis:
    """Stand-in for the scikit-fem ``Basis`` returned by the reference's solve_fem (P:145):
    carries the mesh nodes and interpolates the P1 solution."""

    def __init__(self, nodes):
        self.nodes = nodes

    def interpolator(self, u):
        nodes = self.nodes
        return lambda x: np.interp(np.asarray(x).reshape(-1), nodes, u)


class FEMLSSVRPrimalSolver:
    """Reference class P:107-211 with the arithmetic on the GPU.

    Keyword-only additions (defaults reproduce the reference): ``rhs_func`` (default: the shipped
    ``poisson_rhs``, evaluated on the device; any other callable is sampled on the host - at the two Gauss points of
    every element for the coarse solve and at the collocation points for the element solves - so that BOTH stages solve
    -u'' = rhs_func), ``k_freq`` (forcing (k pi)^2 sin(k pi x) when rhs_func is None), ``n_colloc`` (P:40 hard-codes 12), ``coarse_solver`` ('assembled': the reference's rounded
    tridiagonal system, faithful at any size | 'assembled_exact' | 'flux': the same equations without the rounded
    diagonal, which is what to use beyond ~1e5 nodes - see include/hfl.h), ``form``
    ('primal' | 'dual').
    """

    def __init__(self, num_fem_nodes=5, lssvr_M=12, lssvr_gamma=1e6, global_domain=(-1, 1), *,
                 rhs_func=None, k_freq=1.0, n_colloc=N_COLLOCATION, coarse_solver='assembled', form='primal'):
        self.num_fem_nodes = num_fem_nodes
        self.lssvr_M = lssvr_M
        self.lssvr_gamma = lssvr_gamma
        self.global_domain = global_domain
        self.fem_nodes = None
        self.fem_values = None
        self.lssvr_functions = []
        self.rhs_func = rhs_func
        self.k_freq = k_freq
        self.n_colloc = n_colloc
        self.coarse_solver = coarse_solver
        self.form = form
        self.element_status = None
        self._d_nodes = self._d_u = self._d_coef = None

    def solve_fem(self):
        """Coarse P1 solve (P:117-145).  Returns (u_fem, basis) like the reference."""
        a, b = self.global_domain
        self._d_nodes = batch.mesh_linspace(a, b, self.num_fem_nodes)   # P:120
        self.fem_nodes = self._d_nodes.cpu().numpy()
        ul, ur = main_boundary_condition_left(a), main_boundary_condition_right(b)
        if self._sine_frequency() is not None:
            self._d_u = batch.fem_p1_solve(self._d_nodes, k_freq=self._sine_frequency(), u_left=ul, u_right=ur,
                                           coarse_solver=self.coarse_solver)
        else:
            # arbitrary rhs_func: the same P1 assembly (2-point Gauss load, P:129-136) from host samples of the callable
            from .host_api import stream_samples
            fq = stream_samples(self.rhs_func, self.fem_nodes[:-1], self.fem_nodes[1:], batch.GAUSS_X, self._d_nodes.device)
            self._d_u = batch.fem_p1_solve_general(self._d_nodes, torch.ones_like(fq), fq, u_left=ul, u_right=ur)
        self.fem_values = self._d_u.cpu().numpy()
        return self.fem_values.copy(), P1Basis(self.fem_nodes)

    def _sine_frequency(self):
        """k when the forcing is the device family (k pi)^2 sin(k pi x) (rhs_func None or the shipped poisson_rhs), else None."""
        if self.rhs_func is None:
            return self.k_freq
        if self.rhs_func is poisson_rhs:
            return 1.0
        return None

    def solve_lssvr_subproblems(self):
        """All element solves in one launch (P:147-176)."""
        if self._d_nodes is None:
            raise RuntimeError('solve_fem() must run first')
        # the first / last element use the global Dirichlet values (P:158-159, P:68-79)
        u = self._d_u.clone()
        u[0] = main_boundary_condition_left(self.global_domain[0])
        u[-1] = main_boundary_condition_right(self.global_domain[1])
        if self._sine_frequency() is not None:
            forcing = 'sine'
            k = self._sine_frequency()
        else:
            from .host_api import stream_samples          # pinned staging + async copies (P:45 evaluated on the host)
            forcing = stream_samples(self.rhs_func, self.fem_nodes[:-1], self.fem_nodes[1:],
                                     np.linspace(0.0, 1.0, self.n_colloc), self._d_nodes.device)
            k = self.k_freq
        fn = batch.lssvr_primal_batch if self.form == 'primal' else batch.lssvr_dual_batch
        coef, _, status = fn(self._d_nodes, u, self.lssvr_M, self.lssvr_gamma, N=self.n_colloc,
                             forcing=forcing, k_freq=k, want_status=True)
        self._d_coef = coef
        self.element_status = status.cpu().numpy()
        for i in np.nonzero(self.element_status)[0]:
            print(f'Error in element {i + 1}: factorisation broke down; linear interpolation used')   # P:172
        self.lssvr_functions = _LegendreList(coef.cpu().numpy(), self.fem_nodes)

    def solve(self):
        """Complete solution: FEM + LSSVR (P:178-181)."""
        self.solve_fem()
        self.solve_lssvr_subproblems()

    def evaluate_solution(self, x_points):
        """Hybrid solution at arbitrary points (P:184-211)."""
        x_points = np.asarray(x_points)
        xs = torch.from_numpy(np.ascontiguousarray(x_points, dtype=np.float64).reshape(-1)).to(self._d_nodes.device)
        vals = batch.evaluate_points(self._d_nodes, self._d_coef, xs).cpu().numpy()
        return vals.reshape(x_points.shape)      # vals is in C order of x_points, whatever its memory layout
