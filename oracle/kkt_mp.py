"""High-precision (mpmath) solve of the reference's per-element QP.

TEST INFRASTRUCTURE.  Same problem as oracle/kkt.py (P:40-91), every quantity
built in ``mp.dps``-digit arithmetic from the FP64 inputs, full KKT matrix solved
by LU.  Used to bound the round-off of the FP64 restatement and of the CUDA path
independently of each other.
"""
import mpmath as mp


def _legendre_012(M, xi):
    """P_k, P_k', P_k'' at xi for k < M by the three-term recurrence, in mp arithmetic."""
    P = [mp.mpf(1), xi]
    d1 = [mp.mpf(0), mp.mpf(1)]
    d2 = [mp.mpf(0), mp.mpf(0)]
    for k in range(1, M - 1):
        P.append(((2 * k + 1) * xi * P[k] - k * P[k - 1]) / (k + 1))
        d1.append(((2 * k + 1) * (P[k] + xi * d1[k]) - k * d1[k - 1]) / (k + 1))
        d2.append(((2 * k + 1) * (2 * d1[k] + xi * d2[k]) - k * d2[k - 1]) / (k + 1))
    return P[:M], d1[:M], d2[:M]


def lssvr_primal_mp(f_vals, xmin, xmax, u_xmin, u_xmax, M, gamma, dps=80):
    """Exact (to dps digits) minimiser; f_vals = f at the N equispaced collocation points.

    Returns (w, lam) as lists of mpf.
    """
    mp.mp.dps = dps
    N = len(f_vals)
    xmin, xmax = mp.mpf(float(xmin)), mp.mpf(float(xmax))
    h = xmax - xmin
    scl = 2 / h
    A = mp.zeros(N, M)
    for j in range(N):
        xi = -1 + mp.mpf(2 * j) / (N - 1)
        _, _, d2 = _legendre_012(M, xi)
        for k in range(M):
            A[j, k] = -scl * scl * d2[k]
    B = mp.zeros(2, M)
    for k in range(M):
        B[0, k] = (-1) ** k
        B[1, k] = 1
    f = mp.matrix([mp.mpf(float(v)) for v in f_vals])
    g = mp.matrix([mp.mpf(float(u_xmin)), mp.mpf(float(u_xmax))])
    gam = mp.mpf(float(gamma))
    K = mp.zeros(M + 2, M + 2)
    H = mp.eye(M) + gam * (A.T * A)
    for i in range(M):
        for k in range(M):
            K[i, k] = H[i, k]
        for c in range(2):
            K[i, M + c] = B[c, i]
            K[M + c, i] = B[c, i]
    rhs = mp.zeros(M + 2, 1)
    r = gam * (A.T * f)
    for i in range(M):
        rhs[i] = r[i]
    rhs[M] = g[0]
    rhs[M + 1] = g[1]
    sol = mp.lu_solve(K, rhs)
    return [sol[i] for i in range(M)], [sol[M], sol[M + 1]]


def evaluate_mp(w, xi_points, dps=80):
    """sum_k w_k P_k(xi) in mp arithmetic."""
    mp.mp.dps = dps
    out = []
    for xi in xi_points:
        P, _, _ = _legendre_012(len(w), mp.mpf(xi) if not isinstance(xi, mp.mpf) else xi)
        out.append(mp.fsum(wk * pk for wk, pk in zip(w, P)))
    return out
