"""Dual form of the reference's per-element LSSVR problem.

TEST INFRASTRUCTURE.  The reference contains no dual implementation: D: (Hybrid-FEM-LSSVR-Dual.py) is the
primal script with comments removed (D:20-98 is P:20-105).  The dual is derived
from the same QP (P:47-81).  With A, B, f, g as in oracle/kkt.py, stationarity of the Lagrangian gives
w = A^T alpha + B^T beta, e = alpha / gamma, and the multipliers solve

    [[A A^T + I/gamma, A B^T], [B A^T, B B^T]] [alpha; beta] = [f; g]

(the LSSVR "kernel" system with K(x, y) = sum_k (P_k o map)''(x) (P_k o map)''(y)).  By strong duality the
resulting w is the primal minimiser, which is this oracle's own check (tests/test_oracle.py).

Numerics: the matrix is SPD in exact arithmetic but has numerical rank <= M (fact 8); it is solved after
the symmetric scaling diag(1/sigma, 1) (sigma = (2/h)^2), which makes every entry O(1)..O(|D|^2), by LU
with partial pivoting.
"""
import numpy as np

from . import kkt


def dual_system(h, M, N, gamma):
    """Scaled dual matrix K0 + tau J and Ct = [-D; B] (so that w = Ct^T z) for an element of width h."""
    _, D, Bm = kkt.reference_tables(M, N)
    sig = (2.0 / h) ** 2
    tau = 1.0 / (gamma * sig * sig)
    Ct = np.vstack([-D, Bm])
    K = Ct @ Ct.T
    K[:N, :N] += tau * np.eye(N)
    return K, Ct, sig


def lssvr_dual(f_vals, xmin, xmax, u_xmin, u_xmax, M, gamma):
    """Coefficients w from the dual system; f_vals = f at the N equispaced collocation points."""
    f_vals = np.asarray(f_vals, dtype=np.float64)
    N = f_vals.shape[0]
    K, Ct, sig = dual_system(xmax - xmin, M, N, gamma)
    rhs = np.concatenate([f_vals / sig, [u_xmin, u_xmax]])
    z = np.linalg.solve(K, rhs)
    return Ct.T @ z


def lssvr_dual_batch(nodes, u, f_samples, M, gamma):
    """All elements of a mesh; f_samples (E, N).  Returns (E, M)."""
    nodes = np.asarray(nodes, dtype=np.float64)
    out = np.empty((len(nodes) - 1, M))
    for e in range(len(nodes) - 1):
        out[e] = lssvr_dual(f_samples[e], nodes[e], nodes[e + 1], u[e], u[e + 1], M, gamma)
    return out
