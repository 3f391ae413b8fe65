// K4: batched per-element DUAL LSSVR solve.
//
// The reference ships no dual code (D: is a copy of P:).  The dual of the QP of
// P:47-81 (Lagrangian stationarity w = A^T alpha + B^T beta, e = alpha / gamma) is the (N+2) x (N+2) system
//     [[A A^T + I/gamma, A B^T], [B A^T, B B^T]] [alpha; beta] = [f; g],      A = -sigma D, sigma = (2/h)^2.
// Scaled by diag(1/sigma, 1) on both sides it reads
//     (K0 + tau J) z = [f / sigma; g],   K0 = Ct Ct^T, Ct = [-D; B], J = diag(I_N, 0), tau = h^4 / (16 gamma),
//     w = Ct^T z
// so the element enters through tau, 1/sigma and the data only.  K0 has rank <= M < N + 2: once
// tau < eps |K0| the matrix is numerically singular (plain Cholesky returns NaN).  It is
// factorised PER ELEMENT by a diagonally pivoted (rank-revealing) Cholesky that stops when the largest
// remaining diagonal entry falls below 2^-10 eps max_diag; the basic solution (zeros outside the pivot
// set) gives the same w as the primal to ~1e-13 wherever the dual formula itself is well conditioned.
//
// One TEAM per element: a warp when N + 2 <= 32 (4 elements per CTA), a 256-thread CTA otherwise.  The
// matrix lives in shared memory and is addressed through a permutation (no physical row swaps); the
// factorisation is right-looking with the trailing update spread over the team.  Right-hand sides (R
// forcing frequencies per element, BASELINE configs[4]) share the factorisation: thread r solves RHS r
// with broadcast reads of L, then the whole team evaluates the R x F fine values.
#include "hfl_dual.cuh"

namespace hfl {

template <int TS>
__device__ __forceinline__ void team_sync() {
    if (TS == 32) __syncwarp(); else __syncthreads();
}

// (value, index) arg-max over the team; result valid in every thread.  red: 2 * (TS / 32) doubles of smem.
template <int TS>
__device__ __forceinline__ void team_argmax(double& v, int& idx, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    if (TS > 32) {
        const int w = threadIdx.x >> 5, nw = TS / 32;
        __syncthreads();
        if ((threadIdx.x & 31) == 0) { red[w] = v; red[nw + w] = (double)idx; }
        __syncthreads();
        v = red[0]; idx = (int)red[nw];
        for (int q = 1; q < nw; ++q) {
            const double ov = red[q];
            const int oi = (int)red[nw + q];
            if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
        }
    }
}


template <int TS>
__global__ void __launch_bounds__(TS == 32 ? 128 : TS) dual_kernel(const DualArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int TEAMS = (TS == 32) ? 4 : 1;
    const int n = a.n, ld = a.ld, M = a.M, N = a.N, F = a.F, R = a.R;
    const int team = (TS == 32) ? (threadIdx.x >> 5) : 0;
    const int t = (TS == 32) ? (threadIdx.x & 31) : threadIdx.x;
    // per-team shared memory: A [n][ld] | invl [n] | wbuf [R][M] | eacc [2R] | red [16] | perm [n] (int) | misc
    const size_t team_doubles = (size_t)n * ld + n + (size_t)R * M + 2 * (size_t)R + 16;
    const size_t team_bytes = ((team_doubles * 8 + (size_t)n * 4 + 16) + 15) / 16 * 16;
    unsigned char* base = smem_raw + team * team_bytes;
    double* A = reinterpret_cast<double*>(base);
    double* invl = A + (size_t)n * ld;
    double* wbuf = invl + n;
    double* eacc = wbuf + (size_t)R * M;      // per right-hand side: sum of weighted squares, max (as bits)
    double* red = eacc + 2 * (size_t)R;
    int* perm = reinterpret_cast<int*>(red + 16);
    int* misc = perm + n;            // [0] = rank

    double bcl = 0.0, bcr = 0.0, x_first = 0.0, x_last = 0.0, invL = 0.0;
    if (a.bc2 != nullptr) {
        bcl = a.bc2[0]; bcr = a.bc2[1];
        x_first = a.nodes[0]; x_last = a.nodes[a.E];
        invL = 1.0 / (x_last - x_first);
    }
    const double eps_tol = 2.220446049250313e-16 * (1.0 / 1024.0);
    for (int i = t; i < 2 * R; i += TS) eacc[i] = 0.0;
    int nfail = 0;

    for (long long e = (long long)blockIdx.x * TEAMS + team; e < a.E; e += (long long)gridDim.x * TEAMS) {
        const double xl = a.nodes[e], xr = a.nodes[e + 1];
        const double h = xr - xl, h2 = h * h;
        const double isig = 0.25 * h2, tau = (h2 * h2) * a.c_tau;

        // ---- A = K0 + tau J
        for (int idx = t; idx < n * n; idx += TS) {
            const int i = idx / n, j = idx - i * n;
            double v = a.K0[idx];
            if (i == j && i < N) v += tau;
            A[i * ld + j] = v;
        }
        for (int i = t; i < n; i += TS) perm[i] = i;
        team_sync<TS>();

        // ---- diagonally pivoted Cholesky with truncation (right-looking, permutation-addressed)
        double dmax0 = 0.0;
        int rank = 0;
        for (int k = 0; k < n; ++k) {
            double v = -1.0;
            int pos = 0x7fffffff;
            for (int i = k + t; i < n; i += TS) {
                const int pi = perm[i];
                const double d = A[pi * ld + pi];
                if (d > v) { v = d; pos = i; }
            }
            team_argmax<TS>(v, pos, red);
            if (k == 0) dmax0 = v;
            if (!(v > eps_tol * dmax0)) break;
            if (t == 0) {
                const int tmp = perm[k]; perm[k] = perm[pos]; perm[pos] = tmp;
            }
            team_sync<TS>();
            const int pk = perm[k];
            const double il = rsqrt(v);
            if (t == 0) invl[k] = il;
            for (int i = k + 1 + t; i < n; i += TS) A[perm[i] * ld + pk] *= il;     // L_ik
            team_sync<TS>();
            const int m = n - k - 1;
            for (int idx = t; idx < m * m; idx += TS) {
                const int ii = idx / m, jj = idx - ii * m;
                const int pi = perm[k + 1 + ii], pj = perm[k + 1 + jj];
                A[pi * ld + pj] = fma(-A[pi * ld + pk], A[pj * ld + pk], A[pi * ld + pj]);
            }
            team_sync<TS>();
            rank = k + 1;
        }
        if (t == 0 && a.status != nullptr) a.status[e] = (rank >= 2) ? 0 : 1;

        // ---- solves: thread r handles right-hand side r
        for (int r0 = 0; r0 < R; r0 += TS) {
            const int r = r0 + t;
            if (r < R) {
                const double kf = a.kf ? a.kf[r] : a.k_scalar;
                const double kk = (kf * 3.14159265358979323846) * (kf * 3.14159265358979323846);
                double ul = a.u[(long long)r * (a.E + 1) + e], ur = a.u[(long long)r * (a.E + 1) + e + 1];
                if (a.bc2 != nullptr) {
                    ul += (bcl * (x_last - xl) + bcr * (xl - x_first)) * invL;
                    ur += (bcl * (x_last - xr) + bcr * (xr - x_first)) * invL;
                }
                double y[DUAL_NMAX];
                const double hstep = h / (double)(N - 1);
                for (int k = 0; k < rank; ++k) {
                    const int pk = perm[k];
                    double b;
                    if (pk < N) {
                        const double fv = (a.forcing == HFL_FORCING_SINE)
                                              ? kk * sinpi(kf * fma((double)pk, hstep, xl))
                                              : a.f[((long long)r * N + pk) * a.E + e];
                        b = fv * isig;
                    } else {
                        b = (pk == N) ? ul : ur;
                    }
                    for (int j = 0; j < k; ++j) b = fma(-A[pk * ld + perm[j]], y[j], b);
                    y[k] = b * invl[k];
                }
                for (int k = rank - 1; k >= 0; --k) {
                    double b = y[k];
                    const int pk = perm[k];
                    for (int j = k + 1; j < rank; ++j) b = fma(-A[perm[j] * ld + pk], y[j], b);
                    y[k] = b * invl[k];
                }
                double* w = wbuf + (size_t)r * M;
                if (rank >= 2) {
                    for (int mm = 0; mm < M; ++mm) {
                        double s = 0.0;
                        for (int k = 0; k < rank; ++k) s = fma(a.Ct[perm[k] * M + mm], y[k], s);
                        w[mm] = s;
                    }
                } else {   // P:171-176 fallback: linear interpolant of the nodal values
                    for (int mm = 0; mm < M; ++mm) w[mm] = 0.0;
                    w[0] = 0.5 * (ul + ur);
                    w[1] = 0.5 * (ur - ul);
                }
                if (a.coef != nullptr) {
                    double* cp = a.coef + ((long long)r * a.E + e) * M;
                    for (int mm = 0; mm < M; ++mm) cp[mm] = w[mm];
                }
            }
            team_sync<TS>();
            // ---- fine grid (and error norms) for the right-hand sides of this batch
            if (F > 0 && (a.fine != nullptr || a.want_err)) {
                const int rb = min(TS, R - r0);
                const double xc = 0.5 * (xl + xr);
                for (int idx = t; idx < rb * F; idx += TS) {
                    const int rr = idx / F, i = idx - rr * F;
                    const double* w = wbuf + (size_t)(r0 + rr) * M;
                    double s = 0.0;
                    for (int mm = M - 1; mm >= 0; --mm) s = fma(w[mm], a.V[i * M + mm], s);
                    if (a.fine != nullptr) a.fine[((long long)(r0 + rr) * a.E + e) * F + i] = s;
                    if (a.want_err) {
                        const double kf = a.kf ? a.kf[r0 + rr] : a.k_scalar;
                        const double xi = (double)(2 * i - (F - 1)) / (double)(F - 1);
                        const double d = s - sinpi(kf * fma(0.5 * h, xi, xc));
                        const double wq = ((i == 0 || i == F - 1) ? 0.5 : 1.0) * h / (double)(F - 1);
                        atomicAdd(eacc + 2 * (r0 + rr), wq * d * d);
                        atomic_max_nonneg(eacc + 2 * (r0 + rr) + 1, fabs(d));
                    }
                }
            }
            team_sync<TS>();
        }
        if (rank < 2) ++nfail;
    }
    if (a.err3 != nullptr) {
        team_sync<TS>();
        for (int r = t; r < R; r += TS) {
            if (a.want_err) {
                atomicAdd(a.err3 + 3 * r, eacc[2 * r]);
                atomic_max_nonneg(a.err3 + 3 * r + 1, eacc[2 * r + 1]);
            }
            if (nfail) atomicAdd(a.err3 + 3 * r + 2, (double)nfail);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Parity-split team kernel (even N): the (N+2) system decouples into an even and an odd block of nh = N/2 + 1
// unknowns (derivation in hfl_dual_small.cu).  One CTA of 256 threads per element = two teams of 128, team 0
// factorises the even block, team 1 the odd one, concurrently (named barriers); the R right-hand sides are solved
// one per thread inside each team; the whole CTA then evaluates the R x F fine values.  Compared with the
// full-system kernel above: 4x fewer flops per pivot step, half the pivot steps per team, half the shared memory.

__device__ __forceinline__ void half_sync(int team) {
    asm volatile("bar.sync %0, 128;" ::"r"(team + 1) : "memory");
}

__device__ __forceinline__ void half_argmax(double& v, int& idx, double* red, int team, int t) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    half_sync(team);
    if ((t & 31) == 0) { red[t >> 5] = v; red[4 + (t >> 5)] = (double)idx; }
    half_sync(team);
    v = red[0]; idx = (int)red[4];
#pragma unroll
    for (int q = 1; q < 4; ++q) {
        const double ov = red[q];
        const int oi = (int)red[4 + q];
        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
}

__global__ void __launch_bounds__(256) dual_parity_kernel(const DualParityArgs pa) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const DualArgs& a = pa.d;
    const int nh = pa.nh, ld = pa.ldh, M = a.M, N = a.N, NHc = N / 2, F = a.F, R = a.R;
    const int team = threadIdx.x >> 7, t = threadIdx.x & 127;
    // shared: per team {A [nh][ld], invl [nh], red [8], perm [nh] int, rank int}; then wbuf [R][M], eacc [2R]
    const size_t team_doubles = (size_t)nh * ld + nh + 8;
    const size_t team_bytes = ((team_doubles * 8 + (size_t)(nh + 2) * 4) + 15) / 16 * 16;
    unsigned char* base = smem_raw + team * team_bytes;
    double* A = reinterpret_cast<double*>(base);
    double* invl = A + (size_t)nh * ld;
    double* red = invl + nh;
    int* perm = reinterpret_cast<int*>(red + 8);
    int* rank_s = perm + nh;
    double* wbuf = reinterpret_cast<double*>(smem_raw + 2 * team_bytes);
    double* eacc = wbuf + (size_t)R * M;
    const int* rank_other = reinterpret_cast<int*>(reinterpret_cast<double*>(smem_raw + (1 - team) * team_bytes) +
                                                   (size_t)nh * ld + nh + 8) + nh;

    double bcl = 0.0, bcr = 0.0, x_first = 0.0, x_last = 0.0, invL = 0.0;
    if (a.bc2 != nullptr) {
        bcl = a.bc2[0]; bcr = a.bc2[1];
        x_first = a.nodes[0]; x_last = a.nodes[a.E];
        invL = 1.0 / (x_last - x_first);
    }
    const double eps_tol = 2.220446049250313e-16 * (1.0 / 1024.0);
    for (int i = threadIdx.x; i < 2 * R; i += 256) eacc[i] = 0.0;
    int nfail = 0;
    const double* Kp = pa.Kp[team];
    const double* Cp = pa.Cp[team];
    const int MA = pa.MA[team];

    for (long long e = blockIdx.x; e < a.E; e += gridDim.x) {
        const double xl = a.nodes[e], xr = a.nodes[e + 1];
        const double h = xr - xl, h2 = h * h;
        const double isig = 0.25 * h2, th = 0.5 * (h2 * h2) * a.c_tau;
        for (int idx = t; idx < nh * nh; idx += 128) {
            const int i = idx / nh, j = idx - i * nh;
            double v = Kp[idx];
            if (i == j && i < NHc) v += th;
            A[i * ld + j] = v;
        }
        for (int i = t; i < nh; i += 128) perm[i] = i;
        half_sync(team);
        double dmax0 = 0.0;
        int rank = 0;
        for (int k = 0; k < nh; ++k) {
            double v = -1.0;
            int pos = 0x7fffffff;
            for (int i = k + t; i < nh; i += 128) {
                const int pi = perm[i];
                const double d = A[pi * ld + pi];
                if (d > v) { v = d; pos = i; }
            }
            half_argmax(v, pos, red, team, t);
            if (k == 0) dmax0 = v;
            if (!(v > eps_tol * dmax0)) break;
            if (t == 0) { const int tmp = perm[k]; perm[k] = perm[pos]; perm[pos] = tmp; }
            half_sync(team);
            const int pk = perm[k];
            const double il = rsqrt(v);
            if (t == 0) invl[k] = il;
            for (int i = k + 1 + t; i < nh; i += 128) A[perm[i] * ld + pk] *= il;
            half_sync(team);
            const int m = nh - k - 1;
            for (int idx = t; idx < m * m; idx += 128) {
                const int ii = idx / m, jj = idx - ii * m;
                const int pi = perm[k + 1 + ii], pj = perm[k + 1 + jj];
                A[pi * ld + pj] = fma(-A[pi * ld + pk], A[pj * ld + pk], A[pi * ld + pj]);
            }
            half_sync(team);
            rank = k + 1;
        }
        if (t == 0) *rank_s = rank;
        __syncthreads();
        const bool ok = rank >= 1 && *rank_other >= 1;
        if (threadIdx.x == 0 && a.status != nullptr) a.status[e] = ok ? 0 : 1;
        if (!ok) ++nfail;

        for (int r0 = 0; r0 < R; r0 += 128) {
            const int r = r0 + t;
            if (r < R) {
                const double kf = a.kf ? a.kf[r] : a.k_scalar;
                const double kk = (kf * 3.14159265358979323846) * (kf * 3.14159265358979323846);
                double ul = a.u[(long long)r * (a.E + 1) + e], ur = a.u[(long long)r * (a.E + 1) + e + 1];
                if (a.bc2 != nullptr) {
                    ul += (bcl * (x_last - xl) + bcr * (xl - x_first)) * invL;
                    ur += (bcl * (x_last - xr) + bcr * (xr - x_first)) * invL;
                }
                const double gpar = team == 0 ? 0.5 * (ul + ur) : 0.5 * (ur - ul);
                double S = 0.0, C = 0.0;
                if (a.forcing == HFL_FORCING_SINE) sincospi(kf * (0.5 * (xl + xr)), &S, &C);
                const double amp = isig * kk * (team == 0 ? S : C);
                const double tb = kf * h * (0.5 / (double)(N - 1));       // base angle / pi
                double y[DUAL_NMAX / 2 + 2];
                for (int k = 0; k < rank; ++k) {
                    const int pk = perm[k];
                    double b;
                    if (pk < NHc) {
                        if (a.forcing == HFL_FORCING_SINE) {
                            double sj, cj;
                            sincospi(tb * (double)(2 * pk + 1), &sj, &cj);
                            b = amp * (team == 0 ? cj : sj);
                        } else {
                            const double fp = a.f[((long long)r * N + NHc + pk) * a.E + e];
                            const double fm = a.f[((long long)r * N + NHc - 1 - pk) * a.E + e];
                            b = isig * (team == 0 ? 0.5 * (fp + fm) : 0.5 * (fp - fm));
                        }
                    } else {
                        b = gpar;
                    }
                    for (int j = 0; j < k; ++j) b = fma(-A[pk * ld + perm[j]], y[j], b);
                    y[k] = b * invl[k];
                }
                for (int k = rank - 1; k >= 0; --k) {
                    double b = y[k];
                    const int pk = perm[k];
                    for (int j = k + 1; j < rank; ++j) b = fma(-A[perm[j] * ld + pk], y[j], b);
                    y[k] = b * invl[k];
                }
                double* w = wbuf + (size_t)r * M;
                for (int q = 0; q < MA; ++q) {
                    double s = 0.0;
                    for (int k = 0; k < rank; ++k) s = fma(Cp[perm[k] * MA + q], y[k], s);
                    w[2 * q + team] = ok ? s : (q == 0 ? gpar : 0.0);   // P:171-176 fallback: linear interpolant
                }
            }
            __syncthreads();
            const int rb = min(128, R - r0);
            if (a.coef != nullptr)
                for (int idx = threadIdx.x; idx < rb * M; idx += 256)
                    a.coef[((long long)(r0 + idx / M) * a.E + e) * M + idx % M] = wbuf[(size_t)(r0 + idx / M) * M + idx % M];
            if (F > 0 && (a.fine != nullptr || a.want_err)) {
                const double xc = 0.5 * (xl + xr);
                for (int idx = threadIdx.x; idx < rb * F; idx += 256) {
                    const int rr = idx / F, i = idx - rr * F;
                    const double* w = wbuf + (size_t)(r0 + rr) * M;
                    double s = 0.0;
                    for (int mm = M - 1; mm >= 0; --mm) s = fma(w[mm], a.V[i * M + mm], s);
                    if (a.fine != nullptr) a.fine[((long long)(r0 + rr) * a.E + e) * F + i] = s;
                    if (a.want_err) {
                        const double kf = a.kf ? a.kf[r0 + rr] : a.k_scalar;
                        const double xi = (double)(2 * i - (F - 1)) / (double)(F - 1);
                        const double d = s - sinpi(kf * fma(0.5 * h, xi, xc));
                        const double wq = ((i == 0 || i == F - 1) ? 0.5 : 1.0) * h / (double)(F - 1);
                        atomicAdd(eacc + 2 * (r0 + rr), wq * d * d);
                        atomic_max_nonneg(eacc + 2 * (r0 + rr) + 1, fabs(d));
                    }
                }
            }
            __syncthreads();
        }
    }
    if (a.err3 != nullptr) {
        __syncthreads();
        for (int r = threadIdx.x; r < R; r += 256) {
            if (a.want_err) {
                atomicAdd(a.err3 + 3 * r, eacc[2 * r]);
                atomic_max_nonneg(a.err3 + 3 * r + 1, eacc[2 * r + 1]);
            }
            if (nfail) atomicAdd(a.err3 + 3 * r + 2, (double)nfail);
        }
    }
}

int launch_dual_small(const hfl_plan* plan, long long E, const double* d_nodes, const double* d_u, int forcing_kind,
                      double k_freq, const double* d_f, const double* d_bc2, double* d_coef, double* d_fine,
                      int* d_status, double* d_err3, cudaStream_t s);   // hfl_dual_small.cu

static int launch_dual(const hfl_plan* plan, long long E, int R, const double* d_nodes, const double* d_u,
                       int forcing_kind, const double* d_kf, double k_scalar, const double* d_f, const double* d_bc2,
                       double* d_coef, double* d_fine, int* d_status, double* d_err3, cudaStream_t s) {
    DualArgs a;
    a.E = E; a.R = R; a.nodes = d_nodes; a.u = d_u; a.f = d_f; a.kf = d_kf; a.k_scalar = k_scalar; a.bc2 = d_bc2;
    a.coef = d_coef; a.fine = d_fine; a.status = d_status; a.err3 = d_err3;
    a.K0 = plan->d_tables + plan->off_K0; a.Ct = plan->d_tables + plan->off_Ct; a.V = plan->d_tables + plan->off_V;
    a.M = plan->M; a.N = plan->N; a.F = plan->F; a.n = plan->N + 2; a.ld = a.n | 1;
    a.forcing = forcing_kind; a.c_tau = 1.0 / (16.0 * plan->gamma);
    a.want_err = (d_err3 != nullptr) && plan->F >= 2;
    const int n = a.n;
    const size_t team_doubles = (size_t)n * a.ld + n + (size_t)R * a.M + 2 * (size_t)R + 16;
    const size_t team_bytes = ((team_doubles * 8 + (size_t)n * 4 + 16) + 15) / 16 * 16;
    int dev = 0, max_smem = 0;
    HFL_CUDA_CHECK(cudaGetDevice(&dev));
    HFL_CUDA_CHECK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (n <= 32 && 4 * team_bytes <= (size_t)max_smem) {
        const size_t smem = 4 * team_bytes;
        HFL_CUDA_CHECK(cudaFuncSetAttribute(dual_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        HFL_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dual_kernel<32>, 128, smem));
        if (per_sm < 1) per_sm = 1;
        long long grid = (E + 3) / 4;
        const long long cap = (long long)sm_count() * per_sm;
        if (grid > cap) grid = cap;
        dual_kernel<32><<<(unsigned)grid, 128, smem, s>>>(a);
    } else if (plan->N % 2 == 0 && get_option_dual_team() != 2) {
        DualParityArgs pa;
        pa.d = a;
        pa.nh = plan->N / 2 + 1; pa.ldh = pa.nh | 1;
        pa.Kp[0] = plan->d_tables + plan->off_Kpe; pa.Kp[1] = plan->d_tables + plan->off_Kpo;
        pa.Cp[0] = plan->d_tables + plan->off_Cpe; pa.Cp[1] = plan->d_tables + plan->off_Cpo;
        pa.MA[0] = n_even(plan->M) + 1; pa.MA[1] = n_odd(plan->M) + 1;
        pa.Vt = plan->d_tables + plan->off_Vt;
        if (get_option_dual_team() != 3) {                  // left-looking kernel (nh <= 96)
            const int rc = launch_dual_parity_left(pa, max_smem, plan, s);
            if (rc == HFL_OK) {
                count_launch();
                HFL_CUDA_CHECK(cudaGetLastError());
                return HFL_OK;
            }
            if (rc != HFL_ERR_UNSUPPORTED) return rc;       // a CUDA failure is reported, not papered over by the fallback
        }
        const size_t td = (size_t)pa.nh * pa.ldh + pa.nh + 8;
        const size_t tb = ((td * 8 + (size_t)(pa.nh + 2) * 4) + 15) / 16 * 16;
        const size_t smem = 2 * tb + ((size_t)R * a.M + 2 * (size_t)R) * 8;
        if (smem > (size_t)max_smem) {
            set_error("hfl_lssvr_dual: N=%d, R=%d, M=%d need %zu bytes of shared memory per element (limit %d)",
                      plan->N, R, plan->M, smem, max_smem);
            return HFL_ERR_UNSUPPORTED;
        }
        HFL_CUDA_CHECK(cudaFuncSetAttribute(dual_parity_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        HFL_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dual_parity_kernel, 256, smem));
        if (per_sm < 1) per_sm = 1;
        long long grid = E;
        const long long cap = (long long)sm_count() * per_sm;
        if (grid > cap) grid = cap;
        dual_parity_kernel<<<(unsigned)grid, 256, smem, s>>>(pa);
    } else {
        if (team_bytes > (size_t)max_smem) {
            set_error("hfl_lssvr_dual: N=%d, R=%d, M=%d need %zu bytes of shared memory per element (limit %d)",
                      plan->N, R, plan->M, team_bytes, max_smem);
            return HFL_ERR_UNSUPPORTED;
        }
        HFL_CUDA_CHECK(cudaFuncSetAttribute(dual_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)team_bytes));
        int per_sm = 0;
        HFL_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dual_kernel<256>, 256, team_bytes));
        if (per_sm < 1) per_sm = 1;
        long long grid = E;
        const long long cap = (long long)sm_count() * per_sm;
        if (grid > cap) grid = cap;
        dual_kernel<256><<<(unsigned)grid, 256, team_bytes, s>>>(a);
    }
    count_launch();
    HFL_CUDA_CHECK(cudaGetLastError());
    return HFL_OK;
}

}  // namespace hfl

using namespace hfl;

static int dual_check(const hfl_plan_t* plan, int64_t E, int R, const double* d_nodes, const double* d_u,
                      int forcing_kind, const double* d_f, double* d_fine) {
    HFL_REQUIRE(plan != nullptr, "hfl_lssvr_dual: plan is NULL");
    { const int drc = plan_on_current_device(plan, "hfl_lssvr_dual"); if (drc != HFL_OK) return drc; }
    HFL_REQUIRE(E >= 0 && R >= 1, "hfl_lssvr_dual: E < 0 or R < 1");
    HFL_REQUIRE(E == 0 || (d_nodes != nullptr && d_u != nullptr), "hfl_lssvr_dual: d_nodes / d_u is NULL");
    HFL_REQUIRE(forcing_kind == HFL_FORCING_SINE || forcing_kind == HFL_FORCING_SAMPLES,
                "hfl_lssvr_dual: unknown forcing_kind %d", forcing_kind);
    HFL_REQUIRE(forcing_kind != HFL_FORCING_SAMPLES || d_f != nullptr, "hfl_lssvr_dual: HFL_FORCING_SAMPLES needs d_f_samples");
    HFL_REQUIRE(d_fine == nullptr || plan->F >= 2, "hfl_lssvr_dual: d_fine given but the plan has F = 0");
    HFL_REQUIRE(plan->N + 2 <= DUAL_NMAX, "hfl_lssvr_dual: N=%d exceeds the dual limit of %d collocation points", plan->N,
                DUAL_NMAX - 2);
    return HFL_OK;
}

extern "C" int hfl_lssvr_dual_batch(const hfl_plan_t* plan, int64_t E, const double* d_nodes, const double* d_u,
                                    int forcing_kind, double k_freq, const double* d_f_samples, const double* d_bc2,
                                    double* d_coef, double* d_fine, int32_t* d_status, double* d_err3, void* stream) {
    int rc = dual_check(plan, E, 1, d_nodes, d_u, forcing_kind, d_f_samples, d_fine);
    if (rc != HFL_OK || E == 0) return rc;
    if (get_option_dual_team() == 0) {   // parity-split register kernel when the shape is covered
        rc = launch_dual_small(plan, E, d_nodes, d_u, forcing_kind, k_freq, d_f_samples, d_bc2, d_coef, d_fine, d_status,
                               d_err3, (cudaStream_t)stream);
        if (rc >= 0) return rc;
    }
    return launch_dual(plan, E, 1, d_nodes, d_u, forcing_kind, nullptr, k_freq, d_f_samples, d_bc2, d_coef, d_fine,
                       d_status, d_err3, (cudaStream_t)stream);
}

extern "C" int hfl_lssvr_dual_multi(const hfl_plan_t* plan, int64_t E, int R, const double* d_nodes, const double* d_u,
                                    int forcing_kind, const double* d_k_freq, const double* d_f_samples,
                                    const double* d_bc2, double* d_coef, double* d_fine, int32_t* d_status,
                                    double* d_err3, void* stream) {
    int rc = dual_check(plan, E, R, d_nodes, d_u, forcing_kind, d_f_samples, d_fine);
    if (rc != HFL_OK || E == 0) return rc;
    HFL_REQUIRE(forcing_kind != HFL_FORCING_SINE || d_k_freq != nullptr, "hfl_lssvr_dual_multi: d_k_freq is NULL");
    return launch_dual(plan, E, R, d_nodes, d_u, forcing_kind, d_k_freq, 1.0, d_f_samples, d_bc2, d_coef, d_fine,
                       d_status, d_err3, (cudaStream_t)stream);
}
