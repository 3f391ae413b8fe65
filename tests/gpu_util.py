"""Helpers shared by the GPU parity tests (the checker side uses the oracle; the tested side calls the C ABI)."""
import numpy as np
import torch

from oracle import fem_p1, kkt


def jittered_mesh(E, seed=0, a=-1.0, b=1.0):
    """Widths (2/E)(1 + 0.5 U(-1, 1)), renormalised (SURVEY.md section 8d non-uniform variant)."""
    w = 1.0 + 0.5 * np.random.default_rng(seed).uniform(-1, 1, E)
    x = np.concatenate([[0.0], np.cumsum(w)])
    return a + (b - a) * x / x[-1]


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).cuda()


def sine_samples(nodes, N, k):
    """[N, E] samples of (k pi)^2 sin(k pi x) at the reference's collocation points (P:40)."""
    pts = np.linspace(nodes[:-1], nodes[1:], N, axis=0)
    return fem_p1.forcing(pts, k)


def oracle_coef(nodes, u, M, gamma, N, k=None, f_samples=None):
    f = sine_samples(nodes, N, k) if f_samples is None else f_samples
    return kkt.lssvr_primal_kkt_batch(nodes, u, f.T.copy(), M, gamma)


def rel(a, b):
    return np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b)))
