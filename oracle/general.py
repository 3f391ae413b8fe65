"""Per-element LSSVR for a general 1-D elliptic operator  L u = -(a u')' + c u = f  (SURVEY.md section 8f-2).

TEST INFRASTRUCTURE.  The reference implements only the Poisson case (P:43-45: residual = -u'' - f); its README
(README.md:3) advertises elliptic problems in general.  This is the same QP as oracle/kkt.py with the PDE rows
generalised: A[j, k] = (L phi_k)(x_j), phi_k(x) = P_k(off + scl x), i.e.

    A[j, k] = -a_j scl^2 P_k''(xi_j) - a'_j scl P_k'(xi_j) + c_j P_k(xi_j),

a, a', c, f given as samples at the N equispaced collocation points of each element.  With a = 1, a' = c = 0 it is
oracle/kkt.py exactly (tests/test_oracle.py checks that).  A now depends on the element, so the Gram matrix
A^T A is formed per element.
"""
import numpy as np
from numpy.polynomial import legendre as npleg


def basis_tables(M, N):
    """P, P', P'' of degree < M at xi_j = -1 + 2j/(N-1): three (N, M) arrays."""
    xi = np.linspace(-1.0, 1.0, N)
    eye = np.eye(M)
    P0 = npleg.legval(xi, eye).T
    P1 = npleg.legval(xi, npleg.legder(eye, 1, axis=0)).T
    P2 = npleg.legval(xi, npleg.legder(eye, 2, axis=0)).T
    return P0, P1, P2


def lssvr_general_kkt_batch(nodes, u, a_s, da_s, c_s, f_s, M, gamma):
    """nodes (E+1,), u (E+1,), samples (E, N) each.  Returns coefficients (E, M).

    Solved as the full KKT system of the problem scaled by 1/sigma^2 (sigma = scl^2):
        min tau/2 |w|^2 + 1/2 |f/sigma - Ah w|^2  s.t. B w = g,   Ah = A / sigma, tau = 1 / (gamma sigma^2),
        [[tau I + Ah^T Ah, B^T], [B, 0]] [w; mu] = [Ah^T f / sigma; g]
    by LU with partial pivoting.  Unlike the Poisson case the first two columns of A do not vanish (c, a' terms),
    so I + gamma A^T A is no longer block diagonal and its condition number (~gamma sigma^2) would be felt; the
    scaled KKT matrix is well conditioned because B is invertible on the near-null space of Ah^T Ah.
    tests/test_oracle.py checks this against an 80-digit solve.
    """
    nodes = np.asarray(nodes, dtype=np.float64)
    E, N = f_s.shape
    P0, P1, P2 = basis_tables(M, N)
    h = nodes[1:] - nodes[:-1]
    hh, isig = 0.5 * h, 0.25 * h * h
    Ah = (-a_s[:, :, None] * P2[None]
          - (da_s * hh[:, None])[:, :, None] * P1[None]
          + (c_s * isig[:, None])[:, :, None] * P0[None])                # (E, N, M) = A / sigma
    fh = f_s * isig[:, None]
    tau = isig * isig / gamma
    Bm = np.stack([(-1.0) ** np.arange(M), np.ones(M)])
    K = np.zeros((E, M + 2, M + 2))
    K[:, :M, :M] = np.einsum('enk,enm->ekm', Ah, Ah) + tau[:, None, None] * np.eye(M)[None]
    K[:, :M, M:] = Bm.T[None]
    K[:, M:, :M] = Bm[None]
    rhs = np.zeros((E, M + 2))
    rhs[:, :M] = np.einsum('enk,en->ek', Ah, fh)
    rhs[:, M] = u[:-1]
    rhs[:, M + 1] = u[1:]
    return np.linalg.solve(K, rhs[:, :, None])[:, :M, 0]


def lssvr_general_mp(xmin, xmax, u_l, u_r, a_s, da_s, c_s, f_s, M, gamma, dps=80):
    """The same element problem in mpmath (unscaled KKT system, LU at `dps` digits); returns w as floats."""
    import mpmath as mp
    from .kkt_mp import _legendre_012
    mp.mp.dps = dps
    N = len(f_s)
    h = mp.mpf(float(xmax)) - mp.mpf(float(xmin))
    scl = 2 / h
    A = mp.zeros(N, M)
    for j in range(N):
        P, d1, d2 = _legendre_012(M, -1 + mp.mpf(2 * j) / (N - 1))
        for k in range(M):
            A[j, k] = (-mp.mpf(float(a_s[j])) * scl * scl * d2[k] - mp.mpf(float(da_s[j])) * scl * d1[k]
                       + mp.mpf(float(c_s[j])) * P[k])
    K = mp.zeros(M + 2, M + 2)
    H = mp.eye(M) + mp.mpf(float(gamma)) * (A.T * A)
    r = mp.mpf(float(gamma)) * (A.T * mp.matrix([mp.mpf(float(v)) for v in f_s]))
    rhs = mp.zeros(M + 2, 1)
    for i in range(M):
        for k in range(M):
            K[i, k] = H[i, k]
        K[i, M] = (-1) ** i
        K[M, i] = (-1) ** i
        K[i, M + 1] = 1
        K[M + 1, i] = 1
        rhs[i] = r[i]
    rhs[M] = mp.mpf(float(u_l))
    rhs[M + 1] = mp.mpf(float(u_r))
    sol = mp.lu_solve(K, rhs)
    return np.array([float(sol[i]) for i in range(M)])
