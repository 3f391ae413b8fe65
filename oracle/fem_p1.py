"""CPU restatement of the reference's coarse P1 FEM solve (P:117-145).

TEST INFRASTRUCTURE.  PARITY UNPINNED: the reference calls scikit-fem 11.0.0
(README.md:37), which is neither installed nor fetchable here, so this file
follows scikit-fem's documented behaviour instead of being checked against a run
of it.  What each step restates:

* P:120  ``MeshLine(np.linspace(a, b, n))``  -> nodes, elements (i, i+1)
* P:121-122 ``Basis(m, ElementLineP1())``     -> default integration order
           2*maxdeg = 2 -> 2-point Gauss-Legendre mapped to [0, 1]:
           X = 1/2 -+ 1/(2 sqrt 3), W = 1/2, 1/2; affine map x = h X + x_e,
           |detDF| = h, grad(phi) = -+ 1/h
* P:125-127, P:135  A = assemble(-grad u . grad v)
* P:129-132, P:136  b = assemble(-(pi^2) sin(pi x) v)
* P:137  ``enforce(A, b, D=basis.get_dofs())`` -> rows of the two boundary dofs
           zeroed, unit diagonal, b[D] = 0 (columns are left alone)
* P:138  ``solve(A, b)`` -> scipy.sparse.linalg.spsolve (SuperLU)
* P:141-143  interpolation at the nodes returns the dof values

Anchor used instead of a scikit-fem run: on a uniform mesh the discrete solution
is exactly c(h) sin(k pi x_i) with the closed form in ``c_factor`` (sin is an
eigenvector of the P1 stencil and of the 2-point-Gauss load); c(1/12) - 1 =
3.274e-6 is the signature of 2-point Gauss (exact load integration would give
~1e-16, 3-point Gauss ~1e-9).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

_GX = 0.5 * np.polynomial.legendre.leggauss(2)[0] + 0.5   # Gauss points on [0, 1]
_GW = 0.5 * np.polynomial.legendre.leggauss(2)[1]         # weights 1/2, 1/2


def forcing(x, k_freq=1.0):
    """(k pi)^2 sin(k pi x); k=1 is the reference's poisson_rhs (P:11-12)."""
    kp = k_freq * np.pi
    return kp ** 2 * np.sin(kp * x)


def assemble_p1(nodes, k_freq=1.0):
    """Tridiagonal stiffness (sub, diag, sup) and load of K u = b before Dirichlet rows.

    Sign: the reference assembles A = -K and b = -load (P:127, P:132); negation is
    exact in floating point, so K u = load has the same rounded entries.
    """
    nodes = np.asarray(nodes, dtype=np.float64)
    n = nodes.size
    h = nodes[1:] - nodes[:-1]
    invh = 1.0 / h
    gg = invh * invh
    kloc = gg * (h * _GW[0]) + gg * (h * _GW[1])     # sum_q (1/h)(1/h) * (h W_q)
    diag = np.zeros(n)
    diag[:-1] += kloc
    diag[1:] += kloc
    off = -kloc
    b = np.zeros(n)
    kp2 = (k_freq * np.pi) ** 2
    for q in range(2):
        xq = h * _GX[q] + nodes[:-1]
        fq = kp2 * np.sin(k_freq * np.pi * xq)
        dx = h * _GW[q]
        b[:-1] += (fq * (1.0 - _GX[q])) * dx
        b[1:] += (fq * _GX[q]) * dx
    return off, diag, b


def solve_fem_p1(nodes, k_freq=1.0, solver='spsolve'):
    """Nodal values of the reference's FEM stage.  solver: 'spsolve' (SuperLU, what
    skfem.solve uses) or 'banded' (LAPACK dgbsv, used to measure solver-to-solver spread)."""
    nodes = np.asarray(nodes, dtype=np.float64)
    n = nodes.size
    off, diag, b = assemble_p1(nodes, k_freq)
    lower = off.copy()
    upper = off.copy()
    diag = diag.copy()
    b = b.copy()
    # enforce(): zero the rows of dofs 0 and n-1, unit diagonal, zero rhs
    diag[0] = 1.0
    diag[-1] = 1.0
    upper[0] = 0.0        # row 0 -> column 1
    lower[-1] = 0.0       # row n-1 -> column n-2
    b[0] = 0.0
    b[-1] = 0.0
    if solver == 'spsolve':
        A = sp.diags([lower, diag, upper], [-1, 0, 1], format='csr')
        return spla.spsolve(A, b)
    if solver == 'banded':
        from scipy.linalg import solve_banded
        ab = np.zeros((3, n))
        ab[0, 1:] = upper
        ab[1] = diag
        ab[2, :-1] = lower
        return solve_banded((1, 1), ab, b)
    raise ValueError(solver)


def c_factor(h, k_freq=1.0):
    """Analytic ratio u_i / sin(k pi x_i) of the discrete solution on a uniform mesh."""
    t = k_freq * np.pi * h
    x1 = 0.5 - 0.5 / np.sqrt(3.0)
    x2 = 0.5 + 0.5 / np.sqrt(3.0)
    s = np.sin(0.5 * t)
    return t * t * (x2 * np.cos(t * x1) + x1 * np.cos(t * x2)) / (4.0 * s * s)   # 2 - 2 cos t without cancellation


def solve_fem_p1_general(nodes, a_func, c_func, f_func, u_left=0.0, u_right=0.0):
    """P1 FEM for -(a u')' + c u = f with the same 2-point Gauss rule (stiffness of a, mass matrix of c, load of f),
    Dirichlet rows as identity rows, SuperLU.  Restatement for the general-operator row;
    the reference has no such code, so this is its own anchor (checked on a manufactured solution)."""
    nodes = np.asarray(nodes, dtype=np.float64)
    n = nodes.size
    h = nodes[1:] - nodes[:-1]
    b = np.zeros(n)
    kd = np.zeros(n); ko = np.zeros(n - 1)
    for q in range(2):
        xq = h * _GX[q] + nodes[:-1]
        w = h * _GW[q]
        aq, cq, fq = a_func(xq), c_func(xq), f_func(xq)
        pl, pr = 1.0 - _GX[q], _GX[q]
        ka = aq / h / h * w
        kd[:-1] += ka + cq * pl * pl * w
        kd[1:] += ka + cq * pr * pr * w
        ko += -ka + cq * pl * pr * w
        b[:-1] += fq * pl * w
        b[1:] += fq * pr * w
    lower, upper, diag = ko.copy(), ko.copy(), kd.copy()
    diag[0] = diag[-1] = 1.0
    upper[0] = 0.0
    lower[-1] = 0.0
    b[0], b[-1] = u_left, u_right
    A = sp.diags([lower, diag, upper], [-1, 0, 1], format='csr')
    return spla.spsolve(A, b)
