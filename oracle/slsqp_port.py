"""The reference's own solver formulation, restated: SLSQP on the element QP.

TEST INFRASTRUCTURE / CPU baseline.  Follows P:47-91 - variables [w (M), e (N)], objective
1/2 |w|^2 + gamma/2 |e|^2 (P:47-51), equality constraints (P:53-81) handed to
``scipy.optimize.minimize(method='SLSQP', options={'maxiter': 1000, 'ftol': 1e-12})`` from a small
random start (P:84-91), no analytic Jacobians.  Unlike the reference it evaluates the constraints with
the precomputed matrices of oracle/kkt.py instead of building ``Legendre`` objects on every call, so it
times the algorithm, not numpy's polynomial class.
"""
import numpy as np
from scipy.optimize import minimize

from . import kkt


def lssvr_primal_slsqp(rhs_func, domain_range, u_xmin, u_xmax, M, gamma, N=kkt.N_COLLOCATION_REFERENCE, rng=None):
    xmin, xmax = domain_range
    A, B, x = kkt.element_matrices(xmin, xmax, M, N)
    f = np.asarray(rhs_func(x), dtype=np.float64) * np.ones(N)
    g = np.array([u_xmin, u_xmax], dtype=np.float64)

    def objective(v):
        return 0.5 * np.dot(v[:M], v[:M]) + 0.5 * gamma * np.dot(v[M:], v[M:])

    def constraints(v):
        return np.concatenate([A @ v[:M] - f + v[M:], B @ v[:M] - g])    # -u'' - f + e,  u(x_L) - u_L, u(x_R) - u_R

    rng = rng or np.random.default_rng()
    start = np.concatenate([rng.random(M) * 0.01, np.zeros(N)])
    res = minimize(objective, x0=start, constraints={'type': 'eq', 'fun': constraints}, method='SLSQP',
                   options={'maxiter': 1000, 'ftol': 1e-12})
    return res.x[:M]
