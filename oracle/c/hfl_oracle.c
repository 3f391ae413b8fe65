/* CPU restatement in plain C (OpenMP over elements) of the hot path.  TEST INFRASTRUCTURE / CPU BASELINE ONLY
 * (see oracle/__init__.py): nothing under hybrid_fem_lssvr_b200/ links or loads this.
 *
 * Same mathematics as oracle/kkt.py and oracle/fem_p1.py, written the straightforward way (no parity split, no
 * shared-operator shortcuts): per element the full M x M matrix H = I + gamma A^T A is formed and factorised
 * (Cholesky), the two boundary rows are imposed through the 2 x 2 Schur complement (P:47-81 -> KKT system), and
 * the reconstruction is evaluated on the structured fine grid (P:184-211 -> legval).  The coarse solve restates
 * P:117-145 (2-point Gauss load, Dirichlet rows) with the Thomas algorithm.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXM 32

static void legendre012(int M, double x, double* P, double* d1, double* d2) {
    P[0] = 1.0; d1[0] = 0.0; d2[0] = 0.0;
    if (M > 1) { P[1] = x; d1[1] = 1.0; d2[1] = 0.0; }
    for (int k = 1; k + 1 < M; ++k) {
        double a = 2 * k + 1, b = k, c = k + 1;
        P[k + 1] = (a * x * P[k] - b * P[k - 1]) / c;
        d1[k + 1] = (a * (P[k] + x * d1[k]) - b * d1[k - 1]) / c;
        d2[k + 1] = (a * (2.0 * d1[k] + x * d2[k]) - b * d2[k - 1]) / c;
    }
}

int oracle_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the baseline must use all host threads it can. */
void oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* Coarse P1 solve: nodes[n] -> u[n], forcing (k pi)^2 sin(k pi x), u = 0 at both ends. */
int oracle_fem_p1(long n, const double* nodes, double kf, double* u) {
    const double pi = 3.14159265358979323846, kpi = kf * pi, kp2 = kpi * kpi;
    const double gx0 = 0.5 * (-0.5773502691896257) + 0.5, gx1 = 0.5 * (0.5773502691896257) + 0.5;
    double* kk = (double*)malloc(sizeof(double) * (size_t)n);
    double* b = (double*)calloc((size_t)n, sizeof(double));
    double* cp = (double*)malloc(sizeof(double) * (size_t)n);
    if (!kk || !b || !cp) return 1;
    for (long e = 0; e + 1 < n; ++e) {
        double h = nodes[e + 1] - nodes[e], invh = 1.0 / h, hw = 0.5 * h;
        double kq = (invh * invh) * hw;
        kk[e] = kq + kq;
        double f0 = kp2 * sin(kpi * (h * gx0 + nodes[e])), f1 = kp2 * sin(kpi * (h * gx1 + nodes[e]));
        b[e] += (f0 * (1.0 - gx0)) * hw + (f1 * (1.0 - gx1)) * hw;
        b[e + 1] += (f0 * gx0) * hw + (f1 * gx1) * hw;
    }
    /* Thomas on rows 1 .. n-2 (u_0 = u_{n-1} = 0) */
    u[0] = 0.0; u[n - 1] = 0.0;
    if (n > 2) {
        double d = kk[0] + kk[1];
        cp[1] = -kk[1] / d; u[1] = b[1] / d;
        for (long i = 2; i + 1 < n; ++i) {
            double l = -kk[i - 1];
            d = (kk[i - 1] + kk[i]) - l * cp[i - 1];
            cp[i] = -kk[i] / d;
            u[i] = (b[i] - l * u[i - 1]) / d;
        }
        for (long i = n - 3; i >= 1; --i) u[i] -= cp[i] * u[i + 1];
    }
    free(kk); free(b); free(cp);
    return 0;
}

/* The same coarse system as oracle/fem_p1.py assembles - every matrix entry rounded to double exactly as the scikit-fem
 * restatement rounds it (k_e = fl(fl(1/h)^2 * (h/2)) twice, d_i = fl(k_{i-1} + k_i), no fused contraction) - solved by
 * the Thomas algorithm in IEEE binary128.  cond(K) ~ n^2 <= 1e15 leaves ~1e-19 of error: this is the exact solution of
 * the reference's rounded system to double precision, the yardstick for the GPU coarse solve at sizes where double
 * precision direct solvers (SuperLU, LAPACK) are themselves 1e-8 apart.  exact_rowsum != 0: unrounded diagonal
 * k_{i-1} + k_i (HFL_COARSE_ASSEMBLED_EXACT). */
__attribute__((optimize("fp-contract=off")))
int oracle_fem_p1_quad(long n, const double* nodes, double kf, int exact_rowsum, double u_left, double u_right, double* u) {
    const double pi = 3.14159265358979323846, kpi = kf * pi, kp2 = kpi * kpi;
    const double gx0 = 0.5 * (-0.5773502691896257) + 0.5, gx1 = 0.5 * (0.5773502691896257) + 0.5;
    if (n < 2) return 2;
    double* kk = (double*)malloc(sizeof(double) * (size_t)n);
    double* b = (double*)calloc((size_t)n, sizeof(double));
    __float128* cp = (__float128*)malloc(sizeof(__float128) * (size_t)n);
    __float128* g = (__float128*)malloc(sizeof(__float128) * (size_t)n);
    if (!kk || !b || !cp || !g) return 1;
    for (long e = 0; e + 1 < n; ++e) {
        const double h = nodes[e + 1] - nodes[e], invh = 1.0 / h, gg = invh * invh, hw = h * 0.5;
        const double kq = gg * hw;
        kk[e] = kq + kq;
        const double x0 = h * gx0 + nodes[e], x1 = h * gx1 + nodes[e];
        const double f0 = kp2 * sin(kpi * x0), f1 = kp2 * sin(kpi * x1);
        b[e] += (f0 * (1.0 - gx0)) * hw;
        b[e + 1] += (f0 * gx0) * hw;
        b[e] += (f1 * (1.0 - gx1)) * hw;
        b[e + 1] += (f1 * gx1) * hw;
    }
    u[0] = u_left; u[n - 1] = u_right;
    if (n > 2) {
        __float128 prev_c = 0, prev_g = u_left;      /* row 0 is the identity row u_0 = u_left */
        cp[0] = 0; g[0] = u_left;
        for (long i = 1; i + 1 < n; ++i) {
            const __float128 l = -(__float128)kk[i - 1], r = -(__float128)kk[i];
            const __float128 dd = exact_rowsum ? (__float128)kk[i - 1] + (__float128)kk[i] : (__float128)(kk[i - 1] + kk[i]);
            const __float128 den = dd - l * prev_c;
            prev_c = r / den;
            prev_g = ((__float128)b[i] - l * prev_g) / den;
            cp[i] = prev_c; g[i] = prev_g;
        }
        __float128 x = u_right;
        for (long i = n - 2; i >= 1; --i) {
            x = g[i] - cp[i] * x;
            u[i] = (double)x;
        }
    }
    free(kk); free(b); free(cp); free(g);
    return 0;
}

/* All element solves: coef[E][M] and (optionally) fine[E][F]; returns max |u - sin(k pi x)| over the fine grid. */
double oracle_primal_batch(long E, const double* nodes, const double* u, int M, double gamma, int N, double kf, int F,
                           double* coef, double* fine) {
    const double pi = 3.14159265358979323846, kpi = kf * pi, kp2 = kpi * kpi;
    double* D = (double*)malloc(sizeof(double) * (size_t)N * M);   /* P_k''(xi_j) */
    double* V = (double*)malloc(sizeof(double) * (size_t)(F > 0 ? F : 1) * M);
    double G[MAXM * MAXM];
    double P[MAXM], d1[MAXM], d2[MAXM];
    for (int j = 0; j < N; ++j) {
        legendre012(M, -1.0 + 2.0 * j / (N - 1), P, d1, d2);
        for (int k = 0; k < M; ++k) D[j * M + k] = d2[k];
    }
    for (int i = 0; i < F; ++i) {
        legendre012(M, -1.0 + 2.0 * i / (F - 1), P, d1, d2);
        for (int k = 0; k < M; ++k) V[i * M + k] = P[k];
    }
    for (int a = 0; a < M; ++a)
        for (int c = 0; c < M; ++c) {
            double s = 0.0;
            for (int j = 0; j < N; ++j) s += D[j * M + a] * D[j * M + c];
            G[a * M + c] = s;
        }
    double maxerr = 0.0;
#pragma omp parallel for schedule(static) reduction(max : maxerr)
    for (long e = 0; e < E; ++e) {
        double H[MAXM * MAXM], r[MAXM], y0[MAXM], y1[MAXM], z[MAXM], f[256];
        const double xl = nodes[e], xr = nodes[e + 1], h = xr - xl;
        const double scl = 2.0 / h, sig = scl * scl, gs2 = gamma * sig * sig;
        for (int j = 0; j < N; ++j) f[j] = kp2 * sin(kpi * (xl + h * j / (N - 1)));
        for (int a = 0; a < M; ++a) {
            double s = 0.0;
            for (int j = 0; j < N; ++j) s += D[j * M + a] * f[j];
            r[a] = -gamma * sig * s;                       /* gamma A^T f, A = -sigma D */
            for (int c = 0; c < M; ++c) H[a * M + c] = gs2 * G[a * M + c] + (a == c ? 1.0 : 0.0);
        }
        /* Cholesky H = L L^T (lower, in place) */
        for (int j = 0; j < M; ++j) {
            double d = H[j * M + j];
            for (int k = 0; k < j; ++k) d -= H[j * M + k] * H[j * M + k];
            d = sqrt(d);
            H[j * M + j] = d;
            for (int i = j + 1; i < M; ++i) {
                double s = H[i * M + j];
                for (int k = 0; k < j; ++k) s -= H[i * M + k] * H[j * M + k];
                H[i * M + j] = s / d;
            }
        }
#define HSOLVE(vec)                                                                  \
        for (int i = 0; i < M; ++i) {                                                \
            double s = vec[i];                                                       \
            for (int k = 0; k < i; ++k) s -= H[i * M + k] * vec[k];                  \
            vec[i] = s / H[i * M + i];                                               \
        }                                                                            \
        for (int i = M - 1; i >= 0; --i) {                                           \
            double s = vec[i];                                                       \
            for (int k = i + 1; k < M; ++k) s -= H[k * M + i] * vec[k];              \
            vec[i] = s / H[i * M + i];                                               \
        }
        for (int k = 0; k < M; ++k) { z[k] = r[k]; y0[k] = (k & 1) ? -1.0 : 1.0; y1[k] = 1.0; }
        HSOLVE(z) HSOLVE(y0) HSOLVE(y1)
        double s00 = 0, s01 = 0, s11 = 0, b0 = 0, b1 = 0;
        for (int k = 0; k < M; ++k) {
            double sg = (k & 1) ? -1.0 : 1.0;
            s00 += sg * y0[k]; s01 += sg * y1[k]; s11 += y1[k];
            b0 += sg * z[k]; b1 += z[k];
        }
        b0 -= u[e]; b1 -= u[e + 1];
        const double det = s00 * s11 - s01 * s01;
        const double l0 = (s11 * b0 - s01 * b1) / det, l1 = (s00 * b1 - s01 * b0) / det;
        double w[MAXM];
        for (int k = 0; k < M; ++k) { w[k] = z[k] - y0[k] * l0 - y1[k] * l1; if (coef) coef[e * M + k] = w[k]; }
        for (int i = 0; i < F; ++i) {
            double s = 0.0;
            for (int k = 0; k < M; ++k) s += w[k] * V[i * M + k];
            if (fine) fine[e * (long)F + i] = s;
            double err = fabs(s - sin(kpi * (xl + h * i / (F - 1))));
            if (err > maxerr) maxerr = err;
        }
    }
    free(D); free(V);
    return maxerr;
}
