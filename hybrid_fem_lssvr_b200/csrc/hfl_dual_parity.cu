// K4, parity-split dual solve, LEFT-LOOKING (blocks of nh = N/2 + 1 <= 96 unknowns: every even N the dual path
// supports; BASELINE configs[4] has N = 128, nh = 65).
//
// Same mathematics as dual_parity_kernel (hfl_dual.cu): per element and parity the block K_par + tau J is factorised
// by a diagonally pivoted Cholesky that stops at the numerical rank r, then R right-hand sides are solved one per
// thread and the CTA evaluates the R x F fine values.  K_par = C C^T has rank <= M/2 + 1, so on any mesh fine enough
// for tau to drop below eps |K| the factorisation stops after r ~ M/2 + 2 of the nh possible steps.  A right-looking
// update pays nh^2 flops for each of them on a matrix that must live somewhere; the left-looking form needs only the
// columns it actually pivots on:
//     step k:  p = argmax_i d_i            (running diagonal d_i = A_ii - sum_j L_ij^2, one 64-bit key per row)
//              a_i = A_ip - sum_{j<k} L_ij L_pj,  l_i = a_i / sqrt(a_p),  d_i -= l_i^2,   L[:, k] = l
// with one thread per row (team of 96 threads per parity, CTA = 2 teams), A_ip read from the constant table of the
// plan (L1-resident; tau on the diagonal), L stored column-wise in shared memory (rank x nh).  Per step: one
// redux-based arg-max, 2 + 2k shared loads and k + 1 FMAs per thread, two team barriers.  Flops r^2 nh / 2 instead of
// r nh^2; measured 10x fewer executed instructions per element than the right-looking shared-memory kernel.
// The solve keeps the coefficient accumulation w = C^T z inside the backward sweep (w in registers), and the fine
// evaluation reads the transposed basis table so that a warp covers one row of 32 fine points with coalesced loads.
#include "hfl_dual.cuh"

namespace hfl {

constexpr int LT = 96;          // threads (rows) per team
constexpr int LMAXMA = 17;      // coefficients per parity: M <= 32

__device__ __forceinline__ void lt_sync(int team) {
    asm volatile("bar.sync %0, 96;" ::"r"(team + 1) : "memory");
}

// Largest key over the team (every thread gets it).  red: 3 x 64-bit slots, free on entry (see the call site).
__device__ __forceinline__ unsigned long long lt_max_key(unsigned long long key, unsigned long long* red, int team, int t) {
    const unsigned int hi = (unsigned int)(key >> 32);
    const unsigned int mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned int ml = __reduce_max_sync(0xffffffffu, hi == mh ? (unsigned int)key : 0u);
    if ((t & 31) == 0) red[t >> 5] = ((unsigned long long)mh << 32) | ml;
    lt_sync(team);
    const unsigned long long k0 = red[0], k1 = red[1], k2 = red[2];
    const unsigned long long m = k0 > k1 ? k0 : k1;
    return m > k2 ? m : k2;
}

__device__ __forceinline__ double lt_back_dot(const double* lk, const int* perm, const double* y, int k, int rank) {
    double d0 = 0.0, d1 = 0.0;
    int j = k + 1;
    for (; j + 1 < rank; j += 2) {
        d0 = fma(lk[perm[j]], y[j], d0);
        d1 = fma(lk[perm[j + 1]], y[j + 1], d1);
    }
    if (j < rank) d0 = fma(lk[perm[j]], y[j], d0);
    return d0 + d1;
}

// Shared memory per team: Lc [kc][ldl] (doubles; the first kc = min(nh, LT_KC) columns of L), invl [nh], red [4]
// (64-bit); perm [nh] + rank (int).  Columns kc.. (only reached on coarse meshes, where tau keeps the block at full
// rank) go to a per-CTA slice of a global scratch buffer: shared memory per CTA drops from 83 KB to 40 KB at N = 128
// (5 resident CTAs per SM instead of 2), which is worth 1.6x on this latency-bound kernel.
constexpr int LT_KC = 24;
__host__ __device__ inline size_t lt_team_bytes(int nh, int ldl, int kc) {
    return ((((size_t)kc * ldl + nh + 4) * 8 + (size_t)(nh + 2) * 4) + 15) / 16 * 16;
}

#ifndef HFL_DUAL_LEFT_MINB
#define HFL_DUAL_LEFT_MINB 4      // <= 85 registers: 4 CTAs per SM (measured 2x faster than the uncapped build, same as 5)
#endif
__global__ void __launch_bounds__(2 * LT, HFL_DUAL_LEFT_MINB) dual_parity_left_kernel(const DualParityArgs pa) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const DualArgs& a = pa.d;
    const int nh = pa.nh, ldl = pa.ldh, M = a.M, N = a.N, NHc = N / 2, F = a.F, R = a.R;
    const int team = threadIdx.x / LT, row = threadIdx.x - team * LT;
    const int kc = pa.kc;
    const size_t team_bytes = lt_team_bytes(nh, ldl, kc);
    unsigned char* base = smem_raw + team * team_bytes;
    double* Lc = reinterpret_cast<double*>(base);
    double* invl = Lc + (size_t)kc * ldl;
    double* Lg = pa.spill + ((size_t)blockIdx.x * 2 + team) * (size_t)(nh - kc) * ldl;   // columns kc.. (unused when kc = nh)
    unsigned long long* red = reinterpret_cast<unsigned long long*>(invl + nh);
    int* perm = reinterpret_cast<int*>(red + 4);
    int* rank_s = perm + nh;
    double* wbuf = reinterpret_cast<double*>(smem_raw + 2 * team_bytes);
    double* eacc = wbuf + (size_t)R * M;
    const int* rank_other = reinterpret_cast<int*>(reinterpret_cast<double*>(smem_raw + (1 - team) * team_bytes) +
                                                   (size_t)kc * ldl + nh + 4) + nh;

    double bcl = 0.0, bcr = 0.0, x_first = 0.0, x_last = 0.0, invL = 0.0;
    if (a.bc2 != nullptr) {
        bcl = a.bc2[0]; bcr = a.bc2[1];
        x_first = a.nodes[0]; x_last = a.nodes[a.E];
        invL = 1.0 / (x_last - x_first);
    }
    const double eps_tol = 2.220446049250313e-16 * (1.0 / 1024.0);
    for (int i = threadIdx.x; i < 2 * R; i += 2 * LT) eacc[i] = 0.0;
    int nfail = 0;
    const double* Kp = pa.Kp[team];
    const double* Cp = pa.Cp[team];
    const int MA = pa.MA[team];
    const bool live_row = row < nh;
    const double kdiag = live_row ? __ldg(Kp + (size_t)row * nh + row) : 0.0;

    for (long long e = blockIdx.x; e < a.E; e += gridDim.x) {
        const double xl = a.nodes[e], xr = a.nodes[e + 1];
        const double h = xr - xl, h2 = h * h;
        const double isig = 0.25 * h2, th = 0.5 * (h2 * h2) * a.c_tau;
        double dii = live_row ? kdiag + (row < NHc ? th : 0.0) : 0.0;    // running diagonal entry of this row
        bool alive = live_row;
        double dmax0 = 0.0;
        int rank = 0;
        for (int k = 0; k < nh; ++k) {
            // key = bits of the (positive) diagonal entry with the low 7 mantissa bits replaced by 127 - row: one
            // integer maximum picks the largest entry (to 2^-45 relative) and the smallest row among ties.
            // red is free here: its previous readers have passed the barrier that ends their pivot step (or the CTA
            // barrier that ends an element).
            unsigned long long key = 0ull;
            if (alive && dii > 0.0) key = ((unsigned long long)__double_as_longlong(dii) & ~127ull) | (unsigned long long)(127 - row);
            key = lt_max_key(key, red, team, row);
            const double vkey = __longlong_as_double((long long)(key & ~127ull));
            const int p = 127 - (int)(key & 127ull);
            if (k == 0) dmax0 = vkey;
            if (!(vkey > eps_tol * dmax0)) break;
            // column p of the current Schur complement (this row's entry) and its diagonal entry v (every thread, bitwise
            // the same): two independent accumulation chains each
            double a0 = 0.0, a1 = 0.0, v0 = 0.0, v1 = 0.0;
            if (live_row) a0 = __ldg(Kp + (size_t)p * nh + row) + ((row == p && row < NHc) ? th : 0.0);   // symmetric table
            v0 = __ldg(Kp + (size_t)p * nh + p) + (p < NHc ? th : 0.0);
            const int rs = live_row ? row : 0;
            const double* cp = Lc + p;             // column walkers: entry p / this row's entry of column j
            const double* cr = Lc + rs;
            const int ks = min(k, kc);             // columns held in shared memory
            int j = 0;
            for (; j + 1 < ks; j += 2, cp += 2 * ldl, cr += 2 * ldl) {
                const double lp0 = cp[0], lp1 = cp[ldl];
                a0 = fma(-cr[0], lp0, a0);
                a1 = fma(-cr[ldl], lp1, a1);
                v0 = fma(-lp0, lp0, v0);
                v1 = fma(-lp1, lp1, v1);
            }
            if (j < ks) {
                const double lp0 = cp[0];
                a0 = fma(-cr[0], lp0, a0);
                v0 = fma(-lp0, lp0, v0);
            }
            for (j = kc; j < k; ++j) {             // spilled columns
                const double* g = Lg + (size_t)(j - kc) * ldl;
                const double lp0 = g[p];
                a1 = fma(-g[rs], lp0, a1);
                v1 = fma(-lp0, lp0, v1);
            }
            const double v = v0 + v1;
            if (!(v > 0.0)) break;                  // team-uniform: the tracked diagonal overestimated a vanishing pivot
            const double il = rsqrt(v);
            const double l = alive ? (a0 + a1) * il : 0.0;
            if (live_row) {
                if (k < kc) Lc[(size_t)k * ldl + row] = l;
                else Lg[(size_t)(k - kc) * ldl + row] = l;
            }
            dii = fma(-l, l, dii);
            if (row == p) { alive = false; perm[k] = p; invl[k] = il; }
            rank = k + 1;
            lt_sync(team);
        }
        // No CTA barrier here: a team's solves need only its own factor (perm / invl / L were published by the team
        // barrier that ends every pivot step), so the team that finishes its factorisation first starts solving while
        // the other is still factorising.  The other team's rank (element status) is read after the CTA barrier that
        // follows the solves.
        if (row == 0) *rank_s = rank;

        for (int r0 = 0; r0 < R; r0 += LT) {
            const int r = r0 + row;
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
                double y[LT];
                for (int k = 0; k < rank; ++k) {                           // L y = b in pivot order
                    const int pk = perm[k];
                    double b;
                    if (pk < NHc) {
                        if (a.forcing == HFL_FORCING_SINE) {
                            double sj, cj;
                            sincospi_base(tb * (double)(2 * pk + 1), &sj, &cj);     // Taylor below 2^-7 (any fine mesh)
                            b = amp * (team == 0 ? cj : sj);
                        } else {
                            const double fp = a.f[((long long)r * N + NHc + pk) * a.E + e];
                            const double fm = a.f[((long long)r * N + NHc - 1 - pk) * a.E + e];
                            b = isig * (team == 0 ? 0.5 * (fp + fm) : 0.5 * (fp - fm));
                        }
                    } else {
                        b = gpar;
                    }
                    double b1 = 0.0;
                    const double* cp = Lc + pk;                                  // L[pk][j] = column j, entry pk
                    const int ks = min(k, kc);
                    int j = 0;
                    for (; j + 1 < ks; j += 2, cp += 2 * ldl) {
                        b = fma(-cp[0], y[j], b);
                        b1 = fma(-cp[ldl], y[j + 1], b1);
                    }
                    if (j < ks) b = fma(-cp[0], y[j], b);
                    for (j = kc; j < k; ++j) b1 = fma(-Lg[(size_t)(j - kc) * ldl + pk], y[j], b1);
                    y[k] = (b + b1) * invl[k];
                }
                double wq[LMAXMA];
#pragma unroll
                for (int q = 0; q < LMAXMA; ++q) wq[q] = 0.0;
                for (int k = rank - 1; k >= 0; --k) {                      // L^T z = y, and w += C[perm[k]][:] z_k on the way
                    // sum_j L[perm[j]][k] z_j over the later pivots; the two call sites keep the address space static
                    const double dot = k < kc ? lt_back_dot(Lc + (size_t)k * ldl, perm, y, k, rank)
                                              : lt_back_dot(Lg + (size_t)(k - kc) * ldl, perm, y, k, rank);
                    const double zk = (y[k] - dot) * invl[k];
                    y[k] = zk;
                    const double* crow = Cp + (size_t)perm[k] * MA;
#pragma unroll
                    for (int q = 0; q < LMAXMA; ++q)
                        if (q < MA) wq[q] = fma(__ldg(crow + q), zk, wq[q]);
                }
                double* w = wbuf + (size_t)r * M;
#pragma unroll
                for (int q = 0; q < LMAXMA; ++q)
                    if (q < MA) w[2 * q + team] = (rank >= 1) ? wq[q] : (q == 0 ? gpar : 0.0);
            }
            __syncthreads();
            const bool ok = rank >= 1 && *rank_other >= 1;
            if (!ok) {      // P:171-176 fallback: linear interpolant of the nodal values, both parities (CTA-uniform branch)
                for (int r = r0 + row; r < min(R, r0 + LT); r += LT) {
                    double ul = a.u[(long long)r * (a.E + 1) + e], ur = a.u[(long long)r * (a.E + 1) + e + 1];
                    if (a.bc2 != nullptr) {
                        ul += (bcl * (x_last - xl) + bcr * (xl - x_first)) * invL;
                        ur += (bcl * (x_last - xr) + bcr * (xr - x_first)) * invL;
                    }
                    double* w = wbuf + (size_t)r * M;
                    for (int q = 0; q < MA; ++q) w[2 * q + team] = (q == 0) ? (team == 0 ? 0.5 * (ul + ur) : 0.5 * (ur - ul)) : 0.0;
                }
                __syncthreads();
            }
            if (r0 == 0) {
                if (threadIdx.x == 0 && a.status != nullptr) a.status[e] = ok ? 0 : 1;
                if (!ok) ++nfail;
            }
            const int rb = min(LT, R - r0);
            if (a.coef != nullptr)
                for (int idx = threadIdx.x; idx < rb * M; idx += 2 * LT)
                    a.coef[((long long)(r0 + idx / M) * a.E + e) * M + idx % M] = wbuf[(size_t)(r0 + idx / M) * M + idx % M];
            if (F > 0 && (a.fine != nullptr || a.want_err)) {
                const double xc = 0.5 * (xl + xr);
                for (int idx = threadIdx.x; idx < rb * F; idx += 2 * LT) {
                    const int rr = idx / F, i = idx - rr * F;
                    const double* w = wbuf + (size_t)(r0 + rr) * M;
                    const double* vt = pa.Vt + i + (size_t)(M - 1) * F;
                    double s = 0.0;
                    for (int mm = M - 1; mm >= 0; --mm, vt -= F) s = fma(w[mm], __ldg(vt), s);
                    if (a.fine != nullptr) a.fine[((long long)(r0 + rr) * a.E + e) * F + i] = s;
                    if (a.want_err) {
                        const double kf = a.kf ? a.kf[r0 + rr] : a.k_scalar;
                        const double xi = (double)(2 * i - (F - 1)) / (double)(F - 1);
                        const double d = s - sinpi(kf * fma(0.5 * h, xi, xc));
                        const double wgt = ((i == 0 || i == F - 1) ? 0.5 : 1.0) * h / (double)(F - 1);
                        atomicAdd(eacc + 2 * (r0 + rr), wgt * d * d);
                        atomic_max_nonneg(eacc + 2 * (r0 + rr) + 1, fabs(d));
                    }
                }
            }
            __syncthreads();
        }
    }
    if (a.err3 != nullptr) {
        __syncthreads();
        for (int r = threadIdx.x; r < R; r += 2 * LT) {
            if (a.want_err) {
                atomicAdd(a.err3 + 3 * r, eacc[2 * r]);
                atomic_max_nonneg(a.err3 + 3 * r + 1, eacc[2 * r + 1]);
            }
            if (nfail) atomicAdd(a.err3 + 3 * r + 2, (double)nfail);
        }
    }
}

int dual_parity_left_kc(int nh) { return nh < LT_KC ? nh : LT_KC; }

// Doubles of spill scratch the launch needs: grid x 2 teams x (nh - kc) columns.  0 when every column fits.
size_t dual_parity_left_spill_doubles(int nh, int ldh, long long max_grid) {
    return (size_t)max_grid * 2 * (size_t)(nh - dual_parity_left_kc(nh)) * ldh;
}

// Returns HFL_OK (launched), HFL_ERR_UNSUPPORTED (shape not covered by this kernel: the caller falls back to the
// shared-memory parity kernel) or HFL_ERR_CUDA (a CUDA call failed; the error string is set and nothing was launched).
int launch_dual_parity_left(DualParityArgs pa, int max_smem, const hfl_plan* plan, cudaStream_t s) {
    const DualArgs& a = pa.d;
    if (pa.nh > LT || pa.MA[0] > LMAXMA || pa.MA[1] > LMAXMA) return HFL_ERR_UNSUPPORTED;
    pa.kc = dual_parity_left_kc(pa.nh);
    const size_t smem = 2 * lt_team_bytes(pa.nh, pa.ldh, pa.kc) + ((size_t)a.R * a.M + 2 * (size_t)a.R) * 8;
    if (smem > (size_t)max_smem) return HFL_ERR_UNSUPPORTED;
    HFL_CUDA_CHECK(cudaFuncSetAttribute(dual_parity_left_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    HFL_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dual_parity_left_kernel, 2 * LT, smem));
    if (per_sm < 1) per_sm = 1;
    long long grid = a.E;
    const long long cap = (long long)sm_count() * per_sm;
    if (grid > cap) grid = cap;
    pa.spill = nullptr;
    if (pa.kc < pa.nh) {
        pa.spill = plan_scratch(plan, s, dual_parity_left_spill_doubles(pa.nh, pa.ldh, grid) * sizeof(double));
        if (pa.spill == nullptr) return HFL_ERR_CUDA;       // plan_scratch has set the error string
    }
    dual_parity_left_kernel<<<(unsigned)grid, 2 * LT, smem, s>>>(pa);
    return HFL_OK;
}

}  // namespace hfl
