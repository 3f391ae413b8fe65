"""Closed-form restatement of the reference's per-element LSSVR problem.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Pinned against the reference's
own ``lssvr_primal`` through tests/golden/ (tests/test_oracle.py).

Reference (P: = /root/reference/1D-Possion/Hybrid-FEM-LSSVR.py):

* P:40       training points ``np.linspace(xmin, xmax, 12)``
* P:47-51    objective  1/2 |w|^2 + gamma/2 |e|^2
* P:43-45,62 PDE rows   -u''(x_j) - f(x_j) + e_j = 0,  u = Legendre(w, [xmin, xmax])
* P:68-79    BC rows    u(xmin) = u_xmin, u(xmax) = u_xmax (global Dirichlet
             functions, both 0.0, for the first / last element)
* P:84-91    solved with SLSQP from a random start (ftol 1e-12)
* P:98       result ``Legendre(res.x[:M], domain_range)``

The problem is a strictly convex equality-constrained QP, so its unique
solution is the solution of the KKT system

    [[I + gamma A^T A, B^T], [B, 0]] [w; lam] = [gamma A^T f; g]

with  A[j, k] = -(d^2/dx^2) P_k(off + scl x)|_{x_j}  (numpy ``legder(..., scl)``
semantics, numpy/polynomial/legendre.py:684-700),  B = [P_k(-1); P_k(+1)] and
g = (u_xmin, u_xmax).  SLSQP lands within ~1e-11 of it.
"""
import numpy as np
from numpy.polynomial import legendre as npleg

N_COLLOCATION_REFERENCE = 12  # P:40


def mapparms(xmin, xmax):
    """numpy.polynomial.polyutils.mapparms([xmin, xmax], [-1, 1]) (polyutils.py:284-288)."""
    oldlen = xmax - xmin
    off = (xmax * (-1.0) - xmin * 1.0) / oldlen
    scl = 2.0 / oldlen
    return off, scl


def element_matrices(xmin, xmax, M, N=N_COLLOCATION_REFERENCE):
    """A (N x M), B (2 x M) and the collocation abscissae of one element.

    Follows P:40 (points), P:45/P:59 (u.deriv(2) of Legendre(w, domain)) and
    P:69-78 (u(xmin), u(xmax)) with numpy's own basis routines, so the mapped
    abscissae carry exactly the rounding the reference sees.
    """
    off, scl = mapparms(xmin, xmax)
    x = np.linspace(xmin, xmax, N)
    eye = np.eye(M)
    d2 = npleg.legder(eye, 2, scl=scl, axis=0)          # column k: coefficients of (P_k o map)''
    A = -npleg.legval(off + scl * x, d2).T               # (N, M); columns 0 and 1 vanish
    B = np.stack([npleg.legval(off + scl * xmin, eye),
                  npleg.legval(off + scl * xmax, eye)])  # (2, M)
    return A, B, x


def solve_kkt(A, B, f, g, gamma):
    """Block solve of the KKT system: H = I + gamma A^T A is SPD, S = B H^-1 B^T is 2x2."""
    M = A.shape[1]
    H = np.eye(M) + gamma * (A.T @ A)
    r = gamma * (A.T @ f)
    c = np.linalg.cholesky(H)
    def hsolve(v):
        y = np.linalg.solve(c, v)
        return np.linalg.solve(c.T, y)
    z = hsolve(r)
    Y = hsolve(B.T)
    S = B @ Y
    lam = np.linalg.solve(S, B @ z - g)
    w = z - Y @ lam
    return w, lam


def lssvr_primal_kkt(rhs_func, domain_range, u_xmin, u_xmax, M, gamma,
                     is_left_boundary=False, is_right_boundary=False,
                     global_domain_range=(-1, 1), N=N_COLLOCATION_REFERENCE,
                     return_multipliers=False):
    """Same signature as the reference ``lssvr_primal`` (P:20-22); returns the coefficient vector.

    The boundary-flag branches (P:68-69, P:75-76) replace u_xmin / u_xmax by the
    global Dirichlet value 0.0 (P:14-18).
    """
    xmin, xmax = domain_range
    gxmin, gxmax = global_domain_range
    A, B, x = element_matrices(xmin, xmax, M, N)
    f = np.asarray(rhs_func(x), dtype=np.float64) * np.ones(N)
    gl = 0.0 if (is_left_boundary and xmin == gxmin) else u_xmin
    gr = 0.0 if (is_right_boundary and xmax == gxmax) else u_xmax
    w, lam = solve_kkt(A, B, f, np.array([gl, gr], dtype=np.float64), gamma)
    if return_multipliers:
        return w, lam
    return w


def reference_tables(M, N):
    """Element-independent tables on the ideal abscissae xi_j = -1 + 2j/(N-1).

    D[j, k] = P_k''(xi_j)  (so A = -scl^2 D),  Bm = [(-1)^k; 1].
    """
    xi = np.linspace(-1.0, 1.0, N)
    eye = np.eye(M)
    D = npleg.legval(xi, npleg.legder(eye, 2, axis=0)).T
    Bm = np.stack([(-1.0) ** np.arange(M), np.ones(M)])
    return xi, D, Bm


def lssvr_primal_kkt_batch(nodes, u, f_samples, M, gamma):
    """Vectorised KKT solve for every element of a mesh.

    nodes (E+1,), u (E+1,), f_samples (E, N) = f at linspace(x_e, x_{e+1}, N).
    Uses the shared ideal abscissae (the per-element mapped
    abscissae differ from them only in the last bit, which does not change the
    result at the 1e-13 level; tests/test_oracle.py checks that claim).
    Returns coefficients (E, M).
    """
    nodes = np.asarray(nodes, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    f_samples = np.asarray(f_samples, dtype=np.float64)
    E, N = f_samples.shape
    _, D, Bm = reference_tables(M, N)
    h = nodes[1:] - nodes[:-1]
    scl = 2.0 / h
    sig = scl * scl
    G = D.T @ D
    H = np.eye(M)[None] + (gamma * sig * sig)[:, None, None] * G[None]
    r = (-gamma * sig)[:, None] * (f_samples @ D)                       # gamma A^T f, A = -sig D
    rhs = np.concatenate([r[:, :, None], np.broadcast_to(Bm.T, (E, M, 2))], axis=2)
    X = np.linalg.solve(H, rhs)                                         # H^-1 [r, B^T]
    z, Y = X[:, :, 0], X[:, :, 1:]
    S = np.einsum('am,emb->eab', Bm, Y)
    g = np.stack([u[:-1], u[1:]], axis=1)
    lam = np.linalg.solve(S, (np.einsum('am,em->ea', Bm, z) - g)[:, :, None])[:, :, 0]
    return z - np.einsum('emb,eb->em', Y, lam)


def fine_points(nodes, F):
    """Structured fine grid: F points linspace(x_e, x_{e+1}, F) per element."""
    nodes = np.asarray(nodes, dtype=np.float64)
    t = np.linspace(0.0, 1.0, F)
    return nodes[:-1, None] + (nodes[1:] - nodes[:-1])[:, None] * t[None, :]


def evaluate_fine(coef, F):
    """u at the structured fine grid: sum_k w_k P_k(xi_i), xi_i = -1 + 2i/(F-1)  (P:193 via legval)."""
    xi = np.linspace(-1.0, 1.0, F)
    V = npleg.legvander(xi, coef.shape[1] - 1)      # (F, M)
    return coef @ V.T


def evaluate_solution(nodes, coefs, x_points):
    """Restatement of ``FEMLSSVRPrimalSolver.evaluate_solution`` (P:184-211).

    First element j with nodes[j] <= x <= nodes[j+1] wins (shared nodes go to the
    left element, P:190-197); points outside the mesh use the first / last
    element (P:199-209).  Value = Legendre(w_j, [x_j, x_{j+1}])(x) (P:193).
    """
    nodes = np.asarray(nodes, dtype=np.float64)
    x_points = np.asarray(x_points, dtype=np.float64)
    out = np.zeros_like(x_points)
    ne = len(nodes) - 1
    for i, xi in enumerate(x_points):
        j = int(np.searchsorted(nodes, xi, side='left')) - 1   # first j with nodes[j] <= xi <= nodes[j+1]
        if xi < nodes[0]:
            j = 0
        elif xi > nodes[-1]:
            j = ne - 1
        else:
            j = min(max(j, 0), ne - 1)
        off, scl = mapparms(nodes[j], nodes[j + 1])
        out[i] = npleg.legval(off + scl * xi, coefs[j])
    return out
