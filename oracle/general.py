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
    """nodes (E+1,), u (E+1,), samples (E, N) each.  Returns coefficients (E, M)."""
    nodes = np.asarray(nodes, dtype=np.float64)
    E, N = f_s.shape
    P0, P1, P2 = basis_tables(M, N)
    h = nodes[1:] - nodes[:-1]
    scl = 2.0 / h
    A = (-(a_s * (scl * scl)[:, None])[:, :, None] * P2[None]
         - (da_s * scl[:, None])[:, :, None] * P1[None]
         + c_s[:, :, None] * P0[None])                                   # (E, N, M)
    Bm = np.stack([(-1.0) ** np.arange(M), np.ones(M)])
    # scale rows by 1 / max|A| per element (the KKT solution is invariant; keeps H = I s^2 + gamma A^T A in range)
    H = np.eye(M)[None] + gamma * np.einsum('enk,enm->ekm', A, A)
    r = gamma * np.einsum('enk,en->ek', A, f_s)
    rhs = np.concatenate([r[:, :, None], np.broadcast_to(Bm.T, (E, M, 2))], axis=2)
    X = np.linalg.solve(H, rhs)
    z, Y = X[:, :, 0], X[:, :, 1:]
    S = np.einsum('am,emb->eab', Bm, Y)
    g = np.stack([u[:-1], u[1:]], axis=1)
    lam = np.linalg.solve(S, (np.einsum('am,em->ea', Bm, z) - g)[:, :, None])[:, :, 0]
    return z - np.einsum('emb,eb->em', Y, lam)
