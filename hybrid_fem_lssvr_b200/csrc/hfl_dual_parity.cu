// K4, parity-split dual solve, LEFT-LOOKING (blocks of nh = N/2 + 1 unknowns; BASELINE configs[4] has N = 128,
// nh = 65).
//
// Same mathematics as dual_parity_kernel (hfl_dual.cu): per element and parity the block A = K_par + tau J is factorised
// by a diagonally pivoted Cholesky that stops at the numerical rank r; R right-hand sides share the factor.
// K_par = C C^T has rank <= M/2 + 1, so on any mesh fine enough for tau to drop below eps |K| the factorisation stops
// after r ~ M/2 + 2 of the nh possible steps.  The left-looking form touches only the columns it pivots on:
//     step k:  p = argmax_i d_i            (running diagonal d_i = A_ii - sum_j L_ij^2, one 64-bit key per row)
//              a_i = A_ip - sum_{j<k} L_ij L_pj,  l_i = a_i / sqrt(a_p),  d_i -= l_i^2,   L[:, k] = l
// with one thread per row (team of 96 threads per parity, CTA = 2 teams), A_ip read from the constant table of the
// plan (L1-resident; tau on the diagonal), L stored column-wise in shared memory.
//
// The R right-hand sides never see a triangular solve.  The coefficients are w = C_P^T A_PP^-1 b (P = the pivoted
// rows), a linear map G = C_P^T L^-T L^-1 of MA x r numbers per element and parity:
//   * the MA rows of C^T ride along the factorisation as EXTRA rows of L (threads nh.. of the team, idle otherwise):
//     the left-looking update applied to the "matrix entries" C[p][q] yields Y^T = C_P^T L^-T column by column;
//   * the same MA threads then back-substitute their row against L (G L = Y^T, r^2 / 2 FMAs, in place);
//   * every right-hand side is the matrix-vector product w = G b with b built on the fly (no y[] array, no dependent
//     chain, 128-bit broadcast loads of G): r MA FMAs instead of 2 r^2 + r MA;
//   * the fine grid uses the parity split once more: the even team evaluates E(xi_i) = sum_q w_2q P_2q(xi_i) and the odd
//     team O(xi_i) for the first half of the points only, straight from the coefficient registers, and the CTA writes
//     u(xi_i) = E + O, u(-xi_i) = E - O: half the FMAs of the full table product and no coefficient round trip.
// One team barrier per pivot step: the step that publishes column k of L also publishes every row's updated diagonal
// and the warp maxima of the pivot keys (double-buffered), so the next step starts from three shared loads.
// Round-1 form (two barriers per step, forward/backward solve per right-hand side, full-table fine grid): 58.7 k warp
// instructions per element at M = 25, R = 64; see profiles/ for the line-level breakdown that led here.
#include "hfl_dual.cuh"
#include <type_traits>

namespace hfl {

constexpr int LT = 96;          // threads (rows) per team
constexpr int LDL = LT;         // leading dimension of L (compile-time: column offsets become immediates)
constexpr int LMAXMA = 17;      // coefficients per parity: M <= 32
constexpr int LT_KC_MAX = 24;   // most columns of L held in shared memory
constexpr int LT_KC_MIN = 16;

__device__ __forceinline__ void lt_sync(int team) {
    asm volatile("bar.sync %0, 96;" ::"r"(team + 1) : "memory");
}

// Shared memory per team (doubles unless noted): Lc [kc][LDL] (the first kc columns of L; rows 0..nh-1 the block, rows
// goff..goff+MAPT-1 the extra rows that end up holding G), invl [LT], pc [LT] (per pivot: 2 p + 1 for a collocation
// row, -1 for the constraint row), dval [2][LT] (running diagonals, double-buffered), red [2][4] (64-bit warp maxima of
// the pivot keys, double-buffered), perm [LT] (int), rank + pad (4 int).  Columns kc.. (only reached on coarse meshes,
// where tau keeps the block at full rank) go to a per-CTA slice of a global scratch buffer, which keeps 4 CTAs
// resident per SM at N = 128.
__host__ __device__ inline size_t lt_team_bytes(int kc) {
    return ((size_t)kc * LDL + 4 * LT + 8) * 8 + (size_t)(LT + 4) * 4;
}
__host__ __device__ inline int lt_goff(int nh) { return (nh + 1) & ~1; }
__host__ __device__ inline int lt_nhalf_padded(int F) { return (((F + 1) / 2) + 7) & ~7; }

// w += G[:, k] b for one pivot: g points at the MAPT (even) entries of column k, 16-byte aligned.
template <int MAPT>
__device__ __forceinline__ void lt_axpy(double (&wq)[MAPT], const double* g, double b) {
#pragma unroll
    for (int q = 0; q < MAPT; q += 2) {
        const double2 gg = *reinterpret_cast<const double2*>(g + q);
        wq[q] = fma(gg.x, b, wq[q]);
        wq[q + 1] = fma(gg.y, b, wq[q + 1]);
    }
}

// Warp maximum of the pivot keys into red[warp] (lane 0 writes).
__device__ __forceinline__ void lt_publish_key(unsigned long long key, unsigned long long* red, int t) {
    const unsigned int hi = (unsigned int)(key >> 32);
    const unsigned int mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned int ml = __reduce_max_sync(0xffffffffu, hi == mh ? (unsigned int)key : 0u);
    if ((t & 31) == 0) red[t >> 5] = ((unsigned long long)mh << 32) | ml;
}

// key = bits of the (positive) diagonal entry with the low 7 mantissa bits replaced by 127 - row: one integer maximum
// picks the largest entry (to 2^-45 relative) and the smallest row among ties.
__device__ __forceinline__ unsigned long long lt_key(bool alive, double dii, int row) {
    return (alive && dii > 0.0) ? (((unsigned long long)__double_as_longlong(dii) & ~127ull) | (unsigned long long)(127 - row)) : 0ull;
}

// Right-hand side r of element e, parity par: w = G b.  b is built on the fly per pivot k: collocation rows carry the
// (anti)symmetrised forcing (pc[k] = 2 p + 1 > 0), the constraint row (pc[k] < 0) the nodal value gpar.
template <int MAPT>
__device__ __forceinline__ double lt_rhs_weights(const DualArgs& a, int par, int r, long long e, double xl, double xr, double h,
                                                 const double* bcv, const double* Gs, const double* Gg, int kc, int rank,
                                                 const double* pc, const int* perm, double (&wq)[MAPT]) {
    const int N = a.N, NHc = N / 2;
    const double kf = a.kf ? a.kf[r] : a.k_scalar;
    const double kk = (kf * 3.14159265358979323846) * (kf * 3.14159265358979323846);
    double ul = a.u[(long long)r * (a.E + 1) + e], ur = a.u[(long long)r * (a.E + 1) + e + 1];
    if (a.bc2 != nullptr) {     // bcv = {bc_left, bc_right, x_first, x_last, 1 / (x_last - x_first)}
        ul += (bcv[0] * (bcv[3] - xl) + bcv[1] * (xl - bcv[2])) * bcv[4];
        ur += (bcv[0] * (bcv[3] - xr) + bcv[1] * (xr - bcv[2])) * bcv[4];
    }
    const double gpar = par == 0 ? 0.5 * (ul + ur) : 0.5 * (ur - ul);
    double S = 0.0, C = 0.0;
    const bool sine = a.forcing == HFL_FORCING_SINE;
    if (sine) sincospi(kf * (0.5 * (xl + xr)), &S, &C);
    const double isig = 0.25 * (h * h);
    const double amp = isig * kk * (par == 0 ? S : C);
    const double tb = kf * h * (0.5 / (double)(N - 1));       // base angle / pi
    const bool tiny = fabs(tb * (double)(N - 1)) < 0.0078125;  // every angle below 2^-7: Taylor (any fine mesh)
    const double xb = 3.14159265358979323846 * tb;
    auto rhs_entry = [&](int k) -> double {
        const double c = pc[k];
        if (c < 0.0) return gpar;
        if (sine) {
            if (tiny) {
                const double x = xb * c, z = x * x;
                if (par == 0)
                    return amp * fma(z, fma(z, fma(z, fma(z, 2.48015873015873e-05, -1.388888888888889e-03), 4.1666666666666664e-02), -0.5), 1.0);
                return amp * (x * fma(z, fma(z, fma(z, -1.984126984126984e-04, 8.333333333333333e-03), -1.6666666666666666e-01), 1.0));
            }
            double sj, cj;
            sincospi(tb * c, &sj, &cj);
            return amp * (par == 0 ? cj : sj);
        }
        const int pk = perm[k];
        const double fp = a.f[((long long)r * N + NHc + pk) * a.E + e];
        const double fm = a.f[((long long)r * N + NHc - 1 - pk) * a.E + e];
        return isig * (par == 0 ? 0.5 * (fp + fm) : 0.5 * (fp - fm));
    };
#pragma unroll
    for (int q = 0; q < MAPT; ++q) wq[q] = 0.0;
    const int ks = min(rank, kc);
    for (int k = 0; k < ks; ++k) lt_axpy<MAPT>(wq, Gs + k * LDL, rhs_entry(k));
    for (int k = kc; k < rank; ++k) lt_axpy<MAPT>(wq, Gg + (size_t)(k - kc) * LDL, rhs_entry(k));
    return gpar;
}

// This parity's share of u at 8 of the first half of the fine points (entries q >= MA of wq and of the table are zero).
template <int MAPT>
__device__ __forceinline__ void lt_half_points(const double (&wq)[MAPT], const double* vt, int nhp, double (&acc)[8]) {
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) acc[jj] = 0.0;
#pragma unroll
    for (int q = 0; q < MAPT; ++q) {
        const double2* vv = reinterpret_cast<const double2*>(vt + q * nhp);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const double2 t2 = vv[jj];
            acc[2 * jj] = fma(wq[q], t2.x, acc[2 * jj]);
            acc[2 * jj + 1] = fma(wq[q], t2.y, acc[2 * jj + 1]);
        }
    }
}

// General right-hand side (sampled forcing, or a sine the element does not resolve): w = G b pivot by pivot, then this
// parity's half points from the Legendre table into hp [i * hp_stride] (if asked for); the coefficients go to global
// memory if asked for.  Kept out of line: its MAPT + 8 live doubles would otherwise set the register budget of the
// moment path below.
template <int MAPT>
__device__ __noinline__ double lt_general_task(const DualArgs& a, int par, int r, long long e, double xl, double xr, const double* bcv,
                                               const double* Gs, const double* Gg, int kc, int rank, const double* pc, const int* perm,
                                               const double* vt, int nhp, int MA, double* hp, int hp_stride) {
    double wq[MAPT];
    const double gpar = lt_rhs_weights<MAPT>(a, par, r, e, xl, xr, xr - xl, bcv, Gs, Gg, kc, rank, pc, perm, wq);
    if (a.coef != nullptr) {
        double* w = a.coef + ((long long)r * a.E + e) * a.M;
#pragma unroll
        for (int q = 0; q < MAPT; ++q)
            if (q < MA) w[2 * q + par] = wq[q];
    }
    if (hp != nullptr)
        for (int i0 = 0; i0 < nhp; i0 += 8) {
            double acc[8];
            lt_half_points<MAPT>(wq, vt + i0, nhp, acc);
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) hp[(i0 + jj) * hp_stride] = acc[jj];
        }
    return gpar;
}

// ---- Taylor-moment form of a resolved sine right-hand side ------------------------------------------------------------
// With every collocation angle below 2^-7 (tiny; any mesh fine enough to matter for throughput) the right-hand side of
// pivot k is amp cos(x_b c_k) (even parity) or amp sin(x_b c_k) (odd), c_k = 2 p_k + 1, a degree-8 / degree-9 Taylor
// polynomial in x_b c_k, and the constraint pivot carries gpar.  Everything downstream is linear, so with y = x_b^2
//     w   = fac sum_j y^j  m_j  + gpar m_c,     m_j = G (a_j c^(2j) or a_j c^(2j+1)) over the collocation pivots,
//     u/2 = fac sum_j y^j  t_j  + gpar t_c,     t_j = V_half m_j              (fac = amp, or amp x_b for the odd parity)
// with m_j (6 x MA) and t_j (6 x half the fine points) formed once per factorisation: a right-hand side costs 6 FMAs per
// half point instead of a Taylor polynomial per pivot, r MA for w = G b and MA per half point.
constexpr int LT_NMOM = 6;      // five Taylor moments + the constraint column
__device__ __forceinline__ double lt_taylor_coef(int par, int j) {
    // cos: 1, -1/2, 1/24, -1/720, 1/40320;  sin / x: 1, -1/6, 1/120, -1/5040, 1/362880
    const double ce[5] = {1.0, -0.5, 4.1666666666666664e-02, -1.388888888888889e-03, 2.48015873015873e-05};
    const double co[5] = {1.0, -1.6666666666666666e-01, 8.333333333333333e-03, -1.984126984126984e-04, 2.7557319223985893e-06};
    return par == 0 ? ce[j] : co[j];
}

// Per right-hand side and element: nodal constraint value, amplitude and base angle of the sine forcing.
struct LtTask {
    double gpar, fac, y;
    bool fast;          // sine forcing with every collocation angle below 2^-7: the moment form applies
};
__device__ __forceinline__ LtTask lt_task_setup(const DualArgs& a, int par, int r, long long e, double xl, double xr, const double* bcv) {
    LtTask t;
    const int N = a.N;
    const double h = xr - xl;
    const double kf = a.kf ? a.kf[r] : a.k_scalar;
    const double kk = (kf * 3.14159265358979323846) * (kf * 3.14159265358979323846);
    const long long ec = e < a.E ? e : a.E - 1;        // (groups of the STREAM pass may reach past the mesh: clamped, never stored)
    double ul = a.u[(long long)r * (a.E + 1) + ec], ur = a.u[(long long)r * (a.E + 1) + ec + 1];
    if (a.bc2 != nullptr) {     // bcv = {bc_left, bc_right, x_first, x_last, 1 / (x_last - x_first)}
        ul += (bcv[0] * (bcv[3] - xl) + bcv[1] * (xl - bcv[2])) * bcv[4];
        ur += (bcv[0] * (bcv[3] - xr) + bcv[1] * (xr - bcv[2])) * bcv[4];
    }
    t.gpar = par == 0 ? 0.5 * (ul + ur) : 0.5 * (ur - ul);
    const double tb = kf * h * (0.5 / (double)(N - 1));       // base angle / pi
    t.fast = a.forcing == HFL_FORCING_SINE && fabs(tb * (double)(N - 1)) < 0.0078125;
    double S = 0.0, C = 0.0;
    if (t.fast) sincospi(kf * (0.5 * (xl + xr)), &S, &C);
    const double xb = 3.14159265358979323846 * tb;
    const double amp = (0.25 * (h * h)) * kk * (par == 0 ? S : C);
    t.fac = par == 0 ? amp : amp * xb;
    t.y = xb * xb;
    return t;
}

// Both parities of a streamable task (sine forcing, resolved) from one set of loads and one sincospi.
__device__ __forceinline__ void lt_task_setup2(const DualArgs& a, int r, long long e, double xl, double xr, const double* bcv,
                                               LtTask& te, LtTask& to) {
    const int N = a.N;
    const double h = xr - xl;
    const double kf = a.kf ? a.kf[r] : a.k_scalar;
    const double kk = (kf * 3.14159265358979323846) * (kf * 3.14159265358979323846);
    double ul = a.u[(long long)r * (a.E + 1) + e], ur = a.u[(long long)r * (a.E + 1) + e + 1];
    if (a.bc2 != nullptr) {
        ul += (bcv[0] * (bcv[3] - xl) + bcv[1] * (xl - bcv[2])) * bcv[4];
        ur += (bcv[0] * (bcv[3] - xr) + bcv[1] * (xr - bcv[2])) * bcv[4];
    }
    te.gpar = 0.5 * (ul + ur);
    to.gpar = 0.5 * (ur - ul);
    double S, C;
    sincospi(kf * (0.5 * (xl + xr)), &S, &C);
    const double xb = 3.14159265358979323846 * (kf * h * (0.5 / (double)(N - 1)));
    const double ak = (0.25 * (h * h)) * kk;
    te.fac = ak * S;
    to.fac = (ak * C) * xb;
    te.y = to.y = xb * xb;
    te.fast = to.fast = true;
}

// 8 half points from the moment table tm [LT_NMOM][nhp] (pointer already offset to the first of the 8).
__device__ __forceinline__ void lt_half_points_mom(const LtTask& t, const double* tm, int nhp, double (&acc)[8]) {
#pragma unroll
    for (int jj = 0; jj < 8; jj += 2) {
        double2 p = *reinterpret_cast<const double2*>(tm + 4 * nhp + jj);
#pragma unroll
        for (int j = 3; j >= 0; --j) {
            const double2 c = *reinterpret_cast<const double2*>(tm + j * nhp + jj);
            p.x = fma(p.x, t.y, c.x);
            p.y = fma(p.y, t.y, c.y);
        }
        const double2 tc = *reinterpret_cast<const double2*>(tm + 5 * nhp + jj);
        acc[jj] = fma(t.gpar, tc.x, t.fac * p.x);
        acc[jj + 1] = fma(t.gpar, tc.y, t.fac * p.y);
    }
}

// The coefficients of this parity from the moment table mm [LT_NMOM][MAPT].
template <int MAPT>
__device__ __forceinline__ void lt_weights_mom(const LtTask& t, const double* mm, double (&wq)[MAPT]) {
#pragma unroll
    for (int q = 0; q < MAPT; ++q) {
        double p = mm[4 * MAPT + q];
#pragma unroll
        for (int j = 3; j >= 0; --j) p = fma(p, t.y, mm[j * MAPT + q]);
        wq[q] = fma(t.gpar, mm[5 * MAPT + q], t.fac * p);
    }
}

#ifndef HFL_DUAL_LEFT_MINB
#define HFL_DUAL_LEFT_MINB 4      // <= 85 registers: 4 CTAs per SM
#endif
// MAPT: compile-time even bound on the coefficients per parity (extra rows of the factorisation, width of G).
//
// Two passes per CTA over its elements (a contiguous chunk of E / gridDim.x):
//   STREAM (pa.reuse, sine forcing, no fused error norms): tau enters A only as fl(K_ii + tau) on the collocation rows;
//     while tau stays below half an ulp of the smallest K_ii every such element sees the SAME floating-point matrix as
//     tau = 0.  The CTA factorises that matrix once, forms the moment tables, and then streams the (element,
//     right-hand side) pairs of all elements that also resolve every frequency (k_max h / 2 < 2^-7) without a single
//     barrier: a lane pair (l, l + 16) per right-hand side (even / odd parity), 6 FMAs per half point from the moment
//     table, one shuffle to combine E +- O, 16-byte stores.  Bit for bit what a factorisation per element produces
//     (tests/test_gpu_dual.py::test_factor_reuse_is_bitwise).
//   TEAM: every other element (coarse meshes, sampled forcing, unresolved frequencies, fused error norms, or everything
//     with pa.reuse = 0): factorise per element, 96 threads per parity, right-hand sides one per thread (moment form where
//     it applies, pivot by pivot otherwise), halves combined through shared memory.  The factor of the previous element
//     is still kept while consecutive elements share the matrix.
template <int MAPT>
__global__ void __launch_bounds__(2 * LT, HFL_DUAL_LEFT_MINB) dual_parity_left_kernel(const DualParityArgs pa) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const DualArgs& a = pa.d;
    const int nh = pa.nh, M = a.M, N = a.N, NHc = N / 2, F = a.F, R = a.R;
    const int team = threadIdx.x >= LT ? 1 : 0, row = threadIdx.x - team * LT;
    const int kc = pa.kc;
    const int MA = pa.MA[team];
    const int goff = lt_goff(nh);
    const int nhalf = (F + 1) / 2, nhp = lt_nhalf_padded(F);
    const int RB = R < LT ? R : LT, RP = RB | 1;
    const size_t team_bytes = lt_team_bytes(kc);
    double* Lc = reinterpret_cast<double*>(smem_raw + team * team_bytes);
    double* invl = Lc + (size_t)kc * LDL;
    double* pc = invl + LT;
    double* dval = pc + LT;
    unsigned long long* red = reinterpret_cast<unsigned long long*>(dval + 2 * LT);
    int* perm = reinterpret_cast<int*>(red + 8);
    int* rank_s = perm + LT;
    const size_t spill_team = (size_t)(nh - kc) * LDL;
    double* Lg = pa.spill + ((size_t)blockIdx.x * 2 + team) * spill_team;   // columns kc.. (unused when kc = nh)
    double* eo = reinterpret_cast<double*>(smem_raw + 2 * team_bytes);     // [2][nhp][RP]: E / O of the first half of the fine points
    double* vh = eo + max(2 * nhp * RP, 2 * LT * LT_NMOM);                   // [2][MAPT][nhp]: P_{2q+team}(xi_i), i < nhalf
    double* eacc = vh + (size_t)2 * MAPT * nhp;
    double* bcv = eacc + 2 * R;                                              // {bc_left, bc_right, x_first, x_last, 1 / length}
    double* mom = bcv + 6;                                                   // [2][LT_NMOM][MAPT]: moments of G
    double* tmom = mom + 2 * LT_NMOM * MAPT;                                 // [2][LT_NMOM][nhp]: their half-point values
    const int* rank_other = reinterpret_cast<const int*>(smem_raw + (1 - team) * team_bytes + ((size_t)kc * LDL + 4 * LT + 8) * 8) + LT;
    const bool want_fine = F > 0 && (a.fine != nullptr || a.want_err);

    if (threadIdx.x == 0) {
        double bl = 0.0, br = 0.0, xf = 0.0, xe = 1.0;
        if (a.bc2 != nullptr) { bl = a.bc2[0]; br = a.bc2[1]; xf = a.nodes[0]; xe = a.nodes[a.E]; }
        bcv[0] = bl; bcv[1] = br; bcv[2] = xf; bcv[3] = xe; bcv[4] = 1.0 / (xe - xf);
    }
    const double eps_tol = 2.220446049250313e-16 * (1.0 / 1024.0);
    for (int i = threadIdx.x; i < 2 * R; i += 2 * LT) eacc[i] = 0.0;
    for (int idx = threadIdx.x; idx < 2 * MAPT * nhp; idx += 2 * LT) {
        const int t = idx / (MAPT * nhp), rem = idx - t * (MAPT * nhp), q = rem / nhp, i = rem - q * nhp;
        vh[idx] = (q < pa.MA[t] && i < nhalf) ? __ldg(pa.Vt + (size_t)(2 * q + t) * F + i) : 0.0;
    }
    int nfail = 0;
    const double* Kp = pa.Kp[team];
    const double* Cp = pa.Cp[team];
    const bool live_row = row < nh;
    const int q_ext = row - goff;
    const bool ext = q_ext >= 0 && q_ext < MAPT;       // extra rows: C^T (zero rows pad MA to MAPT)
    const bool ext_live = ext && q_ext < MA;
    const bool stores = live_row || ext;
    const double kdiag = live_row ? __ldg(Kp + row * nh + row) : 0.0;
    const double* vht = vh + team * MAPT * nhp;
    const double* crow = Lc + (stores ? row : 0);      // this row's entries of the columns of L
    // this row's entry of column p of A (the table is symmetric) or, for an extra row, of C^T: abase[p * astride]
    const double* abase = live_row ? Kp + row : (ext_live ? Cp + q_ext : Kp);
    const int astride = live_row ? nh : (ext_live ? MA : 0);
    const bool active = live_row || ext_live;
    // thr_same = min K_ii 2^-54 over the collocation rows of both parities: fl(K_ii + tau) = K_ii below it
    dval[row] = (live_row && row < NHc) ? kdiag : 1.7976931348623157e308;
    __syncthreads();
    double kmin = 1.7976931348623157e308;
    {
        const double* d0 = reinterpret_cast<const double*>(smem_raw) + (size_t)kc * LDL + 2 * LT;
        const double* d1 = reinterpret_cast<const double*>(smem_raw + team_bytes) + (size_t)kc * LDL + 2 * LT;
        for (int i = 0; i < NHc; ++i) kmin = fmin(kmin, fmin(d0[i], d1[i]));
    }
    const double thr_same = pa.reuse ? kmin * 5.551115123125783e-17 : -1.0;     // 2^-54
    bool cached = false;
    int rank = 0;
    __syncthreads();

    // An element goes through the STREAM pass when its matrix is the tau = 0 matrix AND every right-hand side is a sine
    // it resolves (all collocation angles below 2^-7, the moment form): decided from h alone, CTA-uniform.
    double kfmax = fabs(a.k_scalar);
    if (a.kf != nullptr) {
        kfmax = 0.0;
        for (int rr = 0; rr < R; ++rr) kfmax = fmax(kfmax, fabs(a.kf[rr]));
    }
    const bool sine_forcing = a.forcing == HFL_FORCING_SINE;
    auto streamable = [&](double hh) -> bool {
        const double hh2 = hh * hh;
        return sine_forcing && 0.5 * (hh2 * hh2) * a.c_tau < thr_same && fabs(kfmax * hh * 0.5) < 0.0078125;
    };
    bool stream_pass = pa.reuse && !a.want_err && sine_forcing;       // CTA-uniform
    bool streamed = false;                            // elements below thr_same were handled by the STREAM pass
    // this CTA's elements: a contiguous chunk (rows (r, e), (r, e + 1), ... of the fine grid are adjacent in memory)
    const long long per_cta = (a.E + gridDim.x - 1) / gridDim.x;
    const long long e_first = (long long)blockIdx.x * per_cta;
    const long long e_end = e_first + per_cta < a.E ? e_first + per_cta : a.E;
    long long e = e_first;
    while ((stream_pass && e_first < e_end) || e < e_end) {
        double xl = 0.0, xr = 1.0;
        if (!stream_pass) { xl = a.nodes[e]; xr = a.nodes[e + 1]; }
        const double h = xr - xl, h2 = h * h;
        const double th = stream_pass ? 0.0 : 0.5 * (h2 * h2) * a.c_tau;
        const bool same = th < thr_same;          // CTA-uniform (every thread holds the same th and threshold)
        if (!stream_pass && streamed && streamable(h)) { ++e; continue; }
        const int n_mom = LT_NMOM * (MAPT + nhp);       // per parity: mom [LT_NMOM][MAPT] then tmom [LT_NMOM][nhp]
        if (stream_pass && pa.dual0_mode == 1) {
            // the tau = 0 tables of this plan were computed by an earlier launch: {mom, tmom} of both parities, two ranks
            for (int idx = threadIdx.x; idx < 2 * LT_NMOM * MAPT; idx += 2 * LT) mom[idx] = pa.dual0[idx];
            for (int idx = threadIdx.x; idx < 2 * LT_NMOM * nhp; idx += 2 * LT) tmom[idx] = pa.dual0[2 * LT_NMOM * MAPT + idx];
            rank = (int)pa.dual0[2 * n_mom + team];
            if (row == 0) *rank_s = rank;
        } else if (!(cached && same)) {
            cached = same;
            double dii = live_row ? kdiag + (row < NHc ? th : 0.0) : 0.0;    // running diagonal entry of this row
            bool alive = live_row;
            double dmax0 = 0.0;
            rank = 0;
            // red / dval buffer 0 are free: their last readers passed the CTA barrier that ended the previous element
            lt_publish_key(lt_key(alive, dii, row), red, row);
            dval[row] = dii;
            lt_sync(team);
            for (int k = 0; k < nh; ++k) {
                const int buf = k & 1;
                const unsigned long long* rk = red + 4 * buf;
                const unsigned long long k0 = rk[0], k1 = rk[1], k2 = rk[2];
                unsigned long long key = k0 > k1 ? k0 : k1;
                key = key > k2 ? key : k2;
                const double vkey = __longlong_as_double((long long)(key & ~127ull));
                const int p = 127 - (int)(key & 127ull);
                if (k == 0) dmax0 = vkey;
                if (!(vkey > eps_tol * dmax0)) break;
                const double v = dval[buf * LT + p];       // the pivot: running diagonal of row p (>= vkey > 0)
                // this row's entry of column p of the current Schur complement (for an extra row: of C^T), two chains
                // (the pivot row itself needs no entry: its Schur complement entry IS the running diagonal v)
                double a0 = __ldg(abase + p * astride), a1 = 0.0;
                const double* cp = Lc + p;
                const int ks = min(k, kc);             // columns held in shared memory
                int j = 0;
#pragma unroll 2
                for (; j + 1 < ks; j += 2) {
                    a0 = fma(-crow[j * LDL], cp[j * LDL], a0);
                    a1 = fma(-crow[(j + 1) * LDL], cp[(j + 1) * LDL], a1);
                }
                if (j < ks) a0 = fma(-crow[j * LDL], cp[j * LDL], a0);
                for (j = kc; j < k; ++j) {             // spilled columns
                    const double* g = Lg + (size_t)(j - kc) * LDL;
                    a1 = fma(-g[stores ? row : 0], g[p], a1);
                }
                const double il = rsqrt(v);
                const double l = (row == p) ? v * il : (((alive || ext_live) && active) ? (a0 + a1) * il : 0.0);
                if (stores) {
                    if (k < kc) Lc[k * LDL + row] = l;
                    else Lg[(size_t)(k - kc) * LDL + row] = l;
                }
                dii = fma(-l, l, dii);
                if (row == p) {
                    alive = false; perm[k] = p; invl[k] = il;
                    pc[k] = p < NHc ? (double)(2 * p + 1) : -1.0;
                }
                rank = k + 1;
                lt_publish_key(lt_key(alive, dii, row), red + 4 * (buf ^ 1), row);
                dval[(buf ^ 1) * LT + row] = dii;
                lt_sync(team);
            }
            if (row == 0) *rank_s = rank;
            // G L = Y^T, one extra row per thread, in place (only the thread's own entries are written; L, perm and invl
            // were published by the team barrier that ends every pivot step)
            if (ext_live) {
                for (int k = rank - 1; k >= 0; --k) {
                    double acc, acc1 = 0.0;
                    if (k < kc) {
                        const double* lk = Lc + k * LDL;
                        acc = lk[row];
                        const int js = min(rank, kc);
                        int j = k + 1;
                        for (; j + 1 < js; j += 2) {
                            acc = fma(-lk[perm[j]], crow[j * LDL], acc);
                            acc1 = fma(-lk[perm[j + 1]], crow[(j + 1) * LDL], acc1);
                        }
                        if (j < js) { acc = fma(-lk[perm[j]], crow[j * LDL], acc); ++j; }
                        for (j = max(j, kc); j < rank; ++j) acc1 = fma(-lk[perm[j]], Lg[(size_t)(j - kc) * LDL + row], acc1);
                        Lc[k * LDL + row] = (acc + acc1) * invl[k];
                    } else {
                        double* lk = Lg + (size_t)(k - kc) * LDL;
                        acc = lk[row];
                        for (int j = k + 1; j < rank; ++j) acc = fma(-lk[perm[j]], Lg[(size_t)(j - kc) * LDL + row], acc);
                        lk[row] = acc * invl[k];
                    }
                }
            }
            // moments of G over the pivots (own entries only: no barrier needed yet), then their half-point values
            double* Mm = mom + team * LT_NMOM * MAPT;
            if (ext) {
                double mj[5] = {0.0, 0.0, 0.0, 0.0, 0.0}, mcn = 0.0;
                if (ext_live)
                    for (int k = 0; k < rank; ++k) {
                        const double g = k < kc ? Lc[k * LDL + row] : Lg[(size_t)(k - kc) * LDL + row];
                        const double c = pc[k];
                        if (c < 0.0) {
                            mcn += g;
                        } else {
                            const double c2 = c * c;
                            double pw = team == 0 ? 1.0 : c;
#pragma unroll
                            for (int j = 0; j < 5; ++j) {
                                mj[j] = fma(g, lt_taylor_coef(team, j) * pw, mj[j]);
                                pw *= c2;
                            }
                        }
                    }
#pragma unroll
                for (int j = 0; j < 5; ++j) Mm[j * MAPT + q_ext] = mj[j];
                Mm[5 * MAPT + q_ext] = mcn;
            }
            lt_sync(team);          // G and its moments
            for (int idx = row; idx < LT_NMOM * nhp; idx += LT) {
                const int j = idx / nhp, i = idx - j * nhp;
                double sum = 0.0;
                for (int q = 0; q < MAPT; ++q) sum = fma(vht[q * nhp + i], Mm[j * MAPT + q], sum);
                tmom[team * LT_NMOM * nhp + idx] = sum;
            }
            lt_sync(team);
        }

        if (stream_pass) {
            // ---- STREAM: every element of this CTA whose matrix is the tau = 0 matrix and that resolves every frequency.
            // Each lane sets up one (element, right-hand side) task - nodal values, amplitudes, y = x_b^2 for both
            // parities from one sincospi.  The warp then walks its 32 tasks two at a time: 16 lanes per task, lane c
            // owning half point c with its 2 x 6 moment-table entries in registers (no shared-memory read in the loop),
            // 12 FMAs for E and O, and two coalesced stores per task: u(xi_c) = E + O forwards, u(-xi_c) = E - O backwards.
            stream_pass = false;
            __syncthreads();                                   // both parities' moment tables and ranks
            const int rank_q = rank_other[0];
            if (pa.dual0_mode == 2) {                          // this launch only fills the plan's tables
                for (int idx = threadIdx.x; idx < 2 * LT_NMOM * MAPT; idx += 2 * LT) pa.dual0[idx] = mom[idx];
                for (int idx = threadIdx.x; idx < 2 * LT_NMOM * nhp; idx += 2 * LT) pa.dual0[2 * LT_NMOM * MAPT + idx] = tmom[idx];
                if (row == 0) pa.dual0[2 * n_mom + team] = (double)rank;
                return;
            }
            if (rank >= 1 && rank_q >= 1) {                    // CTA-uniform (rank / rank_q are the two teams' ranks)
                streamed = true;
                // task t = r * ne + j: consecutive lanes take consecutive elements of one right-hand side (coalesced
                // loads of the nodal values, adjacent rows of the fine grid)
                const int lane = threadIdx.x & 31, half = lane >> 4, hl = lane & 15;
                const int ne = (int)(e_end - e_first);
                // does the TEAM pass have anything left to do in this chunk?  (one coalesced sweep over the nodes instead of
                // a serial walk over every element just to skip it)
                bool leftover = false;
                for (long long ee = e_first + threadIdx.x; ee < e_end; ee += 2 * LT)
                    leftover = leftover || !streamable(a.nodes[ee + 1] - a.nodes[ee]);
                if (__syncthreads_count(leftover) == 0) e = e_end;
                const long long ntask = (long long)ne * R;
                int r = (int)threadIdx.x / ne, j = (int)threadIdx.x % ne;
                const int dr = 2 * LT / ne, dj = 2 * LT % ne;
                for (long long t0 = 0; t0 < ntask; t0 += 2 * LT) {
                    const long long es = e_first + j;
                    bool work = t0 + threadIdx.x < ntask;
                    double sxl = 0.0, sxr = 1.0;
                    if (work) { sxl = a.nodes[es]; sxr = a.nodes[es + 1]; }
                    work = work && streamable(sxr - sxl);
                    LtTask te_k, to_k;                         // this lane's task, even and odd parity
                    te_k.gpar = to_k.gpar = te_k.fac = to_k.fac = te_k.y = 0.0;
                    long long off = 0;
                    if (work) {
                        lt_task_setup2(a, r, es, sxl, sxr, bcv, te_k, to_k);
                        off = ((long long)r * a.E + es) * F;
                        if (a.coef != nullptr) {
                            double wq[MAPT];
                            double* w = a.coef + ((long long)r * a.E + es) * M;
                            lt_weights_mom<MAPT>(te_k, mom, wq);
#pragma unroll
                            for (int q = 0; q < MAPT; ++q)
                                if (q < pa.MA[0]) w[2 * q] = wq[q];
                            lt_weights_mom<MAPT>(to_k, mom + LT_NMOM * MAPT, wq);
#pragma unroll
                            for (int q = 0; q < MAPT; ++q)
                                if (q < pa.MA[1]) w[2 * q + 1] = wq[q];
                        }
                        if (r == 0 && a.status != nullptr) a.status[es] = 0;
                    }
                    const unsigned vmask = __ballot_sync(0xffffffffu, work);
                    if (a.fine != nullptr && vmask != 0u) {
                        // the six numbers of every task go through a per-warp staging area (the E / O buffer of the TEAM
                        // pass, idle here): three broadcast 128-bit loads per step instead of twelve shuffles
                        double2* stage = reinterpret_cast<double2*>(eo) + (threadIdx.x >> 5) * 96;
                        __syncwarp();                              // the previous iteration's readers are done
                        stage[3 * lane] = make_double2(te_k.y, te_k.fac);
                        stage[3 * lane + 1] = make_double2(to_k.fac, te_k.gpar);
                        stage[3 * lane + 2] = make_double2(to_k.gpar, __longlong_as_double(off));
                        __syncwarp();
                        for (int ib = 0; ib < nhalf; ib += 16) {
                            const int ih = ib + hl;
                            const bool inb = ih < nhalf;
                            double tbe[LT_NMOM], tbo[LT_NMOM];
#pragma unroll
                            for (int q = 0; q < LT_NMOM; ++q) {
                                tbe[q] = tmom[q * nhp + (inb ? ih : 0)];
                                tbo[q] = tmom[(LT_NMOM + q) * nhp + (inb ? ih : 0)];
                            }
                            // odd F: the middle point is written once
                            const bool w_lo = inb, w_hi = inb && F - 1 - ih != ih;
                            double* out_lo = a.fine + ih;
                            double* out_hi = a.fine + (F - 1 - ih);
                            // (ALL: every lane of the warp has a task - the common case - so the per-step validity tests go)
                            auto steps = [&](auto all_tag) {
                                constexpr bool ALL = decltype(all_tag)::value;
#pragma unroll
                                for (int sstep = 0; sstep < 16; ++sstep) {
                                    if (!ALL && ((vmask >> (2 * sstep)) & 3u) == 0u) continue;          // warp-uniform
                                    const int src = 2 * sstep + half;
                                    const double2 s0 = stage[3 * src], s1 = stage[3 * src + 1], s2 = stage[3 * src + 2];
                                    const double y = s0.x, fe = s0.y, fo = s1.x, ge = s1.y, go = s2.x;
                                    const long long o2 = __double_as_longlong(s2.y);
                                    double pe = tbe[4], po = tbo[4];
#pragma unroll
                                    for (int q = 3; q >= 0; --q) {
                                        pe = fma(pe, y, tbe[q]);
                                        po = fma(po, y, tbo[q]);
                                    }
                                    const double ev = fma(ge, tbe[5], fe * pe), od = fma(go, tbo[5], fo * po);
                                    if (ALL || ((vmask >> src) & 1u)) {
                                        if (w_lo) out_lo[o2] = ev + od;
                                        if (w_hi) out_hi[o2] = ev - od;
                                    }
                                }
                            };
                            if (vmask == 0xffffffffu) steps(std::true_type{});
                            else steps(std::false_type{});
                        }
                    }
                    r += dr; j += dj;
                    if (j >= ne) { j -= ne; ++r; }
                }
            }
            cached = false;       // the TEAM pass starts from its own factor (the tau = 0 one is not an element's)
            __syncthreads();
            continue;
        }

        // ---- TEAM: this element's right-hand sides, one per thread and parity.  No CTA barrier before them: a team's
        // right-hand sides need only its own G, so the team that finishes first starts while the other is still
        // factorising.  The other team's rank (element status) is read after the CTA barrier that follows.
        for (int r0 = 0; r0 < R; r0 += LT) {
            const int r = r0 + row;
            double gpar = 0.0;
            if (r < R) {
                const LtTask tk = lt_task_setup(a, team, r, e, xl, xr, bcv);
                gpar = tk.gpar;
                if (!tk.fast) {
                    lt_general_task<MAPT>(a, team, r, e, xl, xr, bcv, Lc + goff, Lg + goff, kc, rank, pc, perm, vht, nhp, MA,
                                          want_fine ? eo + team * nhp * RP + row : nullptr, RP);
                } else if (a.coef != nullptr) {
                    double wq[MAPT];
                    lt_weights_mom<MAPT>(tk, mom + team * LT_NMOM * MAPT, wq);
                    double* w = a.coef + ((long long)r * a.E + e) * M;
#pragma unroll
                    for (int q = 0; q < MAPT; ++q)
                        if (q < MA) w[2 * q + team] = wq[q];
                }
                if (want_fine && tk.fast) {
                    double* eor = eo + team * nhp * RP + row;        // rows nhalf..nhp-1 are padding (zeros from the table)
                    for (int i0 = 0; i0 < nhp; i0 += 8, eor += 8 * RP) {
                        double acc[8];
                        lt_half_points_mom(tk, tmom + team * LT_NMOM * nhp + i0, nhp, acc);
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) eor[jj * RP] = acc[jj];
                    }
                }
            }
            __syncthreads();
            const bool ok = rank >= 1 && *rank_other >= 1;
            if (!ok) {      // P:171-176 fallback: linear interpolant of the nodal values, both parities (CTA-uniform branch)
                if (r < R) {
                    // team 0: the mean times P_0 = 1; team 1: the half difference times P_1 = xi
                    if (a.coef != nullptr) {
                        double* w = a.coef + ((long long)r * a.E + e) * M;
                        for (int q = 0; q < MA; ++q) w[2 * q + team] = (q == 0) ? gpar : 0.0;
                    }
                    if (want_fine)
                        for (int i = 0; i < nhalf; ++i) eo[((size_t)team * nhp + i) * RP + row] = gpar * vht[i];
                }
                __syncthreads();
            }
            if (r0 == 0) {
                if (threadIdx.x == 0 && a.status != nullptr) a.status[e] = ok ? 0 : 1;
                if (!ok) ++nfail;
            }
            const int rb = min(LT, R - r0);
            if (want_fine) {
                // one warp per right-hand side, lanes along the fine points: coalesced rows, no index division
                const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
                const long long row_step = (long long)(2 * LT / 32) * a.E * F;
                for (int i0 = 0; i0 < F; i0 += 32) {
                    const int i = i0 + lane;
                    const bool valid = i < F;
                    const int ih = valid ? min(i, F - 1 - i) : 0;
                    const bool mirrored = i != ih;
                    const double* pe = eo + ih * RP;
                    const double* po = pe + nhp * RP;
                    if (a.fine != nullptr && valid) {
                        double* out = a.fine + ((long long)(r0 + warp) * a.E + e) * F + i;
                        for (int rr = warp; rr < rb; rr += 2 * LT / 32, out += row_step) {
                            const double od = po[rr];
                            *out = pe[rr] + (mirrored ? -od : od);
                        }
                    }
                    if (a.want_err) {
                        const double xc = 0.5 * (xl + xr);
                        const double xi = (double)(2 * i - (F - 1)) / (double)(F - 1);
                        const double wgt = valid ? ((i == 0 || i == F - 1) ? 0.5 : 1.0) * h / (double)(F - 1) : 0.0;
                        for (int rr = warp; rr < rb; rr += 2 * LT / 32) {
                            const double od = po[rr];
                            const double sv = pe[rr] + (mirrored ? -od : od);
                            const double kf = a.kf ? a.kf[r0 + rr] : a.k_scalar;
                            const double d = valid ? sv - sinpi(kf * fma(0.5 * h, xi, xc)) : 0.0;
                            const double esum = warp_sum(wgt * d * d);      // one shared-memory update per warp and right-hand side
                            const double emax = warp_max(fabs(d));
                            if (lane == 0) {
                                atomicAdd(eacc + 2 * (r0 + rr), esum);
                                atomic_max_nonneg(eacc + 2 * (r0 + rr) + 1, emax);
                            }
                        }
                    }
                }
            }
            __syncthreads();
        }
        ++e;
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

template <int MAPT>
static int launch_left(DualParityArgs pa, int max_smem, const hfl_plan* plan, cudaStream_t s) {
    const DualArgs& a = pa.d;
    if (lt_goff(pa.nh) + MAPT > LT) return HFL_ERR_UNSUPPORTED;            // block rows + extra rows: one thread each
    const int RB = a.R < LT ? a.R : LT, nhp = lt_nhalf_padded(a.F);
    size_t eo_doubles = (size_t)2 * nhp * (RB | 1);
    if (eo_doubles < (size_t)2 * LT * LT_NMOM) eo_doubles = (size_t)2 * LT * LT_NMOM;      // staging area of the STREAM pass
    const size_t common = (eo_doubles + (size_t)2 * MAPT * nhp + 2 * (size_t)a.R + 6 +
                           (size_t)2 * LT_NMOM * (MAPT + nhp)) * 8;
    // as many columns of L in shared memory as keep 4 CTAs on an SM (56 KB each), between LT_KC_MIN and LT_KC_MAX
    int kc = pa.nh < LT_KC_MAX ? pa.nh : LT_KC_MAX;
    while (kc > LT_KC_MIN && 2 * lt_team_bytes(kc) + common > 56 * 1024) --kc;
    if (kc > pa.nh) kc = pa.nh;
    pa.kc = kc;
    pa.ldh = LDL;
    pa.reuse = get_option_dual_reuse();
    const size_t smem = 2 * lt_team_bytes(kc) + common;
    if (smem > (size_t)max_smem) return HFL_ERR_UNSUPPORTED;
    HFL_CUDA_CHECK(cudaFuncSetAttribute(dual_parity_left_kernel<MAPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    HFL_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dual_parity_left_kernel<MAPT>, 2 * LT, smem));
    if (per_sm < 1) per_sm = 1;
    long long grid = a.E;
    const long long cap = (long long)sm_count() * per_sm;
    if (grid > cap) grid = cap;
    // Plan-level tables of the tau = 0 factorisation (STREAM pass): filled once per plan by a one-CTA launch of this
    // kernel, synchronised before the flag is set so that launches on other streams may read them.  Not under stream
    // capture (the synchronisation is illegal there): such launches factorise in the kernel as before.
    pa.dual0 = nullptr;
    pa.dual0_mode = 0;
    const bool stream_pass_possible = pa.reuse && !a.want_err && a.forcing == HFL_FORCING_SINE;
    if (stream_pass_possible) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        (void)cudaStreamIsCapturing(s, &cap);
        std::lock_guard<std::mutex> guard(plan->scratch_mu);
        const size_t n0 = (size_t)2 * LT_NMOM * (MAPT + nhp) + 2;
        if (!plan->dual0_ready && cap == cudaStreamCaptureStatusNone) {
            if (plan->d_dual0 == nullptr) HFL_CUDA_CHECK(cudaMalloc(&plan->d_dual0, n0 * sizeof(double)));
            DualParityArgs pp = pa;
            pp.dual0 = plan->d_dual0;
            pp.dual0_mode = 2;
            pp.spill = nullptr;
            if (pp.kc < pp.nh) {
                HFL_CUDA_CHECK(cudaMallocAsync(&pp.spill, (size_t)2 * (pp.nh - pp.kc) * LDL * sizeof(double), s));
            }
            dual_parity_left_kernel<MAPT><<<1, 2 * LT, smem, s>>>(pp);
            if (pp.spill != nullptr) HFL_CUDA_CHECK(cudaFreeAsync(pp.spill, s));
            HFL_CUDA_CHECK(cudaStreamSynchronize(s));
            HFL_CUDA_CHECK(cudaGetLastError());
            plan->dual0_ready = true;
            count_launch();
        }
        if (plan->dual0_ready) { pa.dual0 = plan->d_dual0; pa.dual0_mode = 1; }
    }
    pa.spill = nullptr;
    if (pa.kc < pa.nh) {
        // grid x 2 teams x (nh - kc) columns of LDL doubles
        pa.spill = plan_scratch(plan, s, (size_t)grid * 2 * (size_t)(pa.nh - pa.kc) * LDL * sizeof(double));
        if (pa.spill == nullptr) return HFL_ERR_CUDA;       // plan_scratch has set the error string
    }
    dual_parity_left_kernel<MAPT><<<(unsigned)grid, 2 * LT, smem, s>>>(pa);
    return HFL_OK;
}

// Returns HFL_OK (launched), HFL_ERR_UNSUPPORTED (shape not covered by this kernel: the caller falls back to the
// shared-memory parity kernel) or HFL_ERR_CUDA (a CUDA call failed; the error string is set and nothing was launched).
int launch_dual_parity_left(DualParityArgs pa, int max_smem, const hfl_plan* plan, cudaStream_t s) {
    const int ma = pa.MA[0] > pa.MA[1] ? pa.MA[0] : pa.MA[1];
    if (ma > LMAXMA) return HFL_ERR_UNSUPPORTED;
    if (ma <= 4) return launch_left<4>(pa, max_smem, plan, s);
    if (ma <= 8) return launch_left<8>(pa, max_smem, plan, s);
    if (ma <= 14) return launch_left<14>(pa, max_smem, plan, s);
    return launch_left<18>(pa, max_smem, plan, s);
}

}  // namespace hfl
