"""numpy model of the K1 partition scheme of csrc/hfl_fem.cu (chunk sweeps -> cyclic reduction over the chunk
heads of a tile -> tile records -> top-level system -> back-substitution), used to pin the index algebra of the CUDA
kernels before they run on a GPU: every formula below has the same name and sign convention as the device code.

Rows are (l, sigma, r, b) with sigma the row sum, d = sigma - l - r.  python scripts/k1_cr_prototype.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from oracle import fem_p1   # noqa: E402


def rows_assembled(nodes, k_freq, exact=False, uL=0.0, uR=0.0):
    """(l, sigma, r, b) per node, the reference's rounded entries; sigma = TwoSum residue of fl(kl + kr)."""
    off, diag, b = fem_p1.assemble_p1(nodes, k_freq)
    k = -off
    n = nodes.size
    l = np.zeros(n); r = np.zeros(n); sg = np.zeros(n)
    l[1:] = -k; r[:-1] = -k
    kl, kr = k[:-1], k[1:]
    d = kl + kr
    t = d - kl
    err = (kl - (d - t)) + (kr - t)
    sg[1:-1] = 0.0 if exact else -err
    l[0] = r[0] = l[-1] = r[-1] = 0.0
    sg[0] = sg[-1] = 1.0
    b = b.copy(); b[0] = uL; b[-1] = uR
    return l, sg, r, b


def chunk_reduce(l, sg, r, b):
    """Interior rows [1:S) of each chunk (arrays [T, S]).  Returns y1, v1, w1, e1, ys, vs, ws, es, each [T]."""
    S = l.shape[1]
    tp, vp, rp, bp = sg[:, 1].copy(), l[:, 1].copy(), r[:, 1].copy(), b[:, 1].copy()
    dp = tp - vp - rp
    for i in range(2, S):
        m = l[:, i] / dp
        tp = sg[:, i] - m * tp
        vp = -m * vp
        bp = b[:, i] - m * bp
        rp = r[:, i]
        dp = tp - vp - rp
    ys, vs, ws, es = bp / dp, vp / dp, rp / dp, tp / dp
    tp, bp, wp, lp = sg[:, S - 1].copy(), b[:, S - 1].copy(), r[:, S - 1].copy(), l[:, S - 1].copy()
    dp = tp - lp - wp
    for i in range(S - 2, 0, -1):
        m = r[:, i] / dp
        tp = sg[:, i] - m * tp
        wp = -m * wp
        bp = b[:, i] - m * bp
        lp = l[:, i]
        dp = tp - lp - wp
    return bp / dp, lp / dp, wp / dp, tp / dp, ys, vs, ws, es


def tile_reduce(l, sg, r, b, T, S):
    """One tile of T * S rows (padded with identity rows).  Returns the tile record (12 numbers) and the per-head
    reduced rows (Ld, Rd, Bd) [T] for the back-substitution pass."""
    l, sg, r, b = (x.reshape(T, S) for x in (l, sg, r, b))
    y1, v1, w1, e1, ys, vs, ws, es = chunk_reduce(l, sg, r, b)
    lp, sp, rp, bp = l[:, 0], sg[:, 0], r[:, 0], b[:, 0]
    L = np.zeros(T + 1); Sg = np.ones(T + 1); R = np.zeros(T + 1); B = np.zeros(T + 1)
    t = np.arange(1, T)
    L[t] = -lp[t] * vs[t - 1]
    R[t] = -rp[t] * w1[t]
    Sg[t] = sp[t] - lp[t] * es[t - 1] - rp[t] * e1[t]
    B[t] = bp[t] - lp[t] * ys[t - 1] - rp[t] * y1[t]
    # cyclic reduction over heads 1 .. T-1; index 0 = tile head, index T = next tile's head (kept as unknown columns)
    delta = 1
    while 2 * delta < T:
        i = np.arange(2 * delta, T, 2 * delta)
        im, ip = i - delta, i + delta
        dm = Sg[im] - L[im] - R[im]
        dq = Sg[ip] - L[ip] - R[ip]
        al = -L[i] / dm
        be = -R[i] / dq
        Ln, Rn = al * L[im], be * R[ip]
        Sn = Sg[i] + al * Sg[im] + be * Sg[ip]
        Bn = B[i] + al * B[im] + be * B[ip]
        L[i], R[i], Sg[i], B[i] = Ln, Rn, Sn, Bn
        delta *= 2
    d = Sg - L - R
    Ld, Rd, Bd, Sd = L / d, R / d, B / d, Sg / d
    # head 1 and head T-1 as functions of the two tile heads: u = Y - V uP - W uQ, E = 1 + V + W
    half = T // 2
    Y, V, W, E = Bd[half], Ld[half], Rd[half], Sd[half]
    YL, VL, WL, EL = Y, V, W, E
    delta = half // 2
    while delta >= 1:          # left path: node delta, right neighbour 2 delta
        i = delta
        YL, VL, WL, EL = Bd[i] - Rd[i] * YL, Ld[i] - Rd[i] * VL, -Rd[i] * WL, Sd[i] - Rd[i] * EL
        delta //= 2
    YR, VR, WR, ER = Y, V, W, E
    delta = half // 2
    while delta >= 1:          # right path: node T - delta, left neighbour T - 2 delta
        i = T - delta
        YR, VR, WR, ER = Bd[i] - Ld[i] * YR, -Ld[i] * VR, Rd[i] - Ld[i] * WR, Sd[i] - Ld[i] * ER
        delta //= 2
    rec = np.array([lp[0], sp[0], rp[0], bp[0],
                    y1[0] - w1[0] * YL, v1[0] - w1[0] * VL, -w1[0] * WL, e1[0] - w1[0] * EL,
                    ys[T - 1] - vs[T - 1] * YR, -vs[T - 1] * VR, ws[T - 1] - vs[T - 1] * WR, es[T - 1] - vs[T - 1] * ER])
    return rec, (Ld[:T].copy(), Rd[:T].copy(), Bd[:T].copy())


def top_rows(recs):
    """Tile-head system (l, sigma, r, b) from the records (same algebra as fem_top_kernel)."""
    nt = recs.shape[0]
    l = np.zeros(nt); sg = np.zeros(nt); r = np.zeros(nt); b = np.zeros(nt)
    for c in range(nt):
        rc = recs[c]
        lc, s, bb = 0.0, rc[1], rc[3]
        if c > 0:
            rp = recs[c - 1]
            lc = -rc[0] * rp[9]
            s = s - rc[0] * rp[11]
            bb = bb - rc[0] * rp[8]
        else:
            s -= rc[0]
        s = s - rc[2] * rc[7]
        rr = -rc[2] * rc[6]
        bb = bb - rc[2] * rc[4]
        l[c], sg[c], r[c], b[c] = lc, s, rr, bb
    return l, sg, r, b


def thomas_rowsum(l, sg, r, b):
    n = l.size
    d = sg - l - r
    c = np.zeros(n); g = np.zeros(n)
    c[0] = r[0] / d[0]; g[0] = b[0] / d[0]
    for i in range(1, n):
        den = d[i] - l[i] * c[i - 1]
        c[i] = r[i] / den
        g[i] = (b[i] - l[i] * g[i - 1]) / den
    x = np.zeros(n)
    x[-1] = g[-1]
    for i in range(n - 2, -1, -1):
        x[i] = g[i] - c[i] * x[i + 1]
    return x


def tile_backsub(l, sg, r, b, T, S, heads, uP, uQ):
    Ld, Rd, Bd = heads
    u = np.zeros(T + 1)
    u[0], u[T] = uP, uQ
    delta = T // 2
    while delta >= 1:
        i = np.arange(delta, T, 2 * delta)
        u[i] = Bd[i] - Ld[i] * u[i - delta] - Rd[i] * u[i + delta]
        delta //= 2
    l, sg, r, b = (x.reshape(T, S) for x in (l, sg, r, b))
    out = np.zeros((T, S))
    out[:, 0] = u[:T]
    for t in range(T):
        ll, ss, rr, bb = l[t, 1:].copy(), sg[t, 1:].copy(), r[t, 1:].copy(), b[t, 1:].copy()
        bb[0] -= ll[0] * u[t]; ss[0] -= ll[0]; ll[0] = 0.0
        bb[-1] -= rr[-1] * u[t + 1]; ss[-1] -= rr[-1]; rr[-1] = 0.0
        out[t, 1:] = thomas_rowsum(ll, ss, rr, bb)
    return out.reshape(-1)


def solve(nodes, k_freq=1.0, T=16, S=8, exact=False, uL=0.0, uR=0.0):
    n = nodes.size
    l, sg, r, b = rows_assembled(nodes, k_freq, exact, uL, uR)
    TS = T * S
    nt = (n + TS - 1) // TS
    pad = nt * TS - n
    l, r = np.concatenate([l, np.zeros(pad)]), np.concatenate([r, np.zeros(pad)])
    sg, b = np.concatenate([sg, np.ones(pad)]), np.concatenate([b, np.zeros(pad)])
    recs, heads = [], []
    for c in range(nt):
        sl = slice(c * TS, (c + 1) * TS)
        rec, hd = tile_reduce(l[sl], sg[sl], r[sl], b[sl], T, S)
        recs.append(rec); heads.append(hd)
    utop = thomas_rowsum(*top_rows(np.array(recs)))
    u = np.zeros(nt * TS)
    for c in range(nt):
        sl = slice(c * TS, (c + 1) * TS)
        uQ = utop[c + 1] if c + 1 < nt else 0.0
        u[sl] = tile_backsub(l[sl], sg[sl], r[sl], b[sl], T, S, heads[c], utop[c], uQ)
    return u[:n]


if __name__ == '__main__':
    rng = np.random.default_rng(0)
    worst = 0.0
    for n in (2, 3, 9, 25, 127, 128, 129, 130, 257, 1000, 4097, 10001):
        for T in (4, 16, 256):
            if T == 256 and n < 1000:
                continue
            nodes = np.linspace(-1, 1, n) if n % 2 else np.sort(np.concatenate([[-1.0, 1.0], rng.uniform(-1, 1, n - 2)]))
            k = 1.0 if n < 100 else 3.0
            ref = fem_p1.solve_fem_p1(nodes, k)
            u = solve(nodes, k, T=T)
            err = np.max(np.abs(u - ref)) / max(1.0, np.max(np.abs(ref)))
            worst = max(worst, err)
            print('n=%6d T=%3d  max|u - spsolve| = %.2e' % (n, T, err))
            assert err < 1e-10, (n, T, err)
    nodes = np.linspace(-1, 1, 5001)
    ref = fem_p1.solve_fem_p1(nodes, 2.0) + (0.3 * (1 - nodes) + -0.8 * (nodes + 1)) / 2
    u = solve(nodes, 2.0, uL=0.3, uR=-0.8)
    print('dirichlet data: %.2e' % np.max(np.abs(u - ref)))
    assert np.max(np.abs(u - ref)) < 1e-10
    n = 200001
    nodes = np.linspace(-1, 1, n)
    u = solve(nodes, 1.0, T=256, exact=True)
    ex = fem_p1.c_factor(2.0 / (n - 1)) * np.sin(np.pi * nodes)
    print('exact row sums, n=%d: max|u - analytic| = %.2e' % (n, np.max(np.abs(u - ex))))
    assert np.max(np.abs(u - ex)) < 1e-12
    print('ok, worst %.2e' % worst)
