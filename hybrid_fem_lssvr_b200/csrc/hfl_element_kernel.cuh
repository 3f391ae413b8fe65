// Element kernel template shared by the primal instantiations (hfl_primal.cu) and the parity-split dual
// instantiations (hfl_dual_small.cu).
#pragma once
#include <cmath>
#include <cstring>
#include "hfl_device.cuh"

namespace hfl {

struct PrimalArgs {
    long long E;
    const double* nodes;
    const double* u;
    const double* f;       // samples [N][E] or NULL
    const double* bc2;     // optional {bc_left, bc_right}
    double* coef;          // optional [E][M]
    double* fine;          // optional [E][F]
    int* status;           // optional [E]
    double* err3;          // optional accumulators
    const double* De;      // [NH][ME]
    const double* Do;      // [NH][MO]
    int N, NH, F;
    int forcing;
    int debug;             // 0 normal; 1 = skip the TMA issue (compute only); 2 = skip the solve (stores only).  Profiling aid.
    double k_freq;
    double kk;             // (k pi)^2
    double c_tau;          // 1 / (16 gamma)
    double cN;             // 0.5 / (N - 1)
    double cF;             // 0.5 / (F - 1)
    double hpk;            // k pi / 2: half-width angle of an element per unit h
};

// Tables of the parity-split DUAL form (hfl_dual.cu header comment): per parity block the (NHD + 1) x (NHD + 1)
// system (C C^T + tau/2 J) z = [f_par / sigma; g_par], w_par = C^T z, already permuted into the plan's static
// pivot order.  NHD = 0 is the primal kernel (empty tables).
template <int M, int NHD>
struct DualSmallTables {
    static constexpr int NE = NHD + 1, MEA = n_even(M) + 1, MOA = n_odd(M) + 1;
    double Ke[NE * (NE + 1) / 2], Ko[NE * (NE + 1) / 2];   // packed lower, pivot order
    double je[NE], jo[NE];                                 // 1 if the pivot is a collocation row (gets tau/2)
    double Ce[NE][MEA], Co[NE][MOA];                       // rows of [-D+; 1] in pivot order
    int perm_e[NE], perm_o[NE];                            // pivot k = natural row perm[k]
    // Moment form (DMOM kernels): tau enters only as fl(K_jj + tau/2) on the collocation pivots, so below thr_same
    // (2^-54 of the smallest of those K_jj) every element has the tau = 0 matrix bit for bit and its solution is ONE linear
    // map of the right-hand side, G = C_P^T (C_P C_P^T)^-1 over the numerical-rank pivots.  For a sine forcing the element
    // resolves (collocation angles below 2^-7) the right-hand side is amp cos / sin(x_b (2 j + 1)), a Taylor polynomial in
    // y = x_b^2, hence  w_par = fac sum_m y^m mom[m] + g_par mom[5]  with mom[m] = G (a_m c^(2m) or a_m c^(2m+1)) formed
    // on the host in long double (hfl_dual_small.cu): 6 FMAs per coefficient instead of the factorisation.
    double mom_e[6][MEA], mom_o[6][MOA];
    double thr_same;
};
template <int M>
struct DualSmallTables<M, 0> { int unused; };

template <int M, int FH>
struct PrimalTables {
    static constexpr int ME = n_even(M), MO = n_odd(M);
    double Ge[ME * (ME + 1) / 2];
    double Go[MO > 0 ? MO * (MO + 1) / 2 : 1];
    double fineE[FH > 0 ? FH : 1][ME];
    double fineO[FH > 0 ? FH : 1][MO + 1];
    // Horner form of the fine-grid evaluation: the fine points xi+_i, z_i = xi_i^2 and the monomial coefficients of the
    // basis, P_{2+2k}(xi) = sum_j TE[j][k] z^j, P_{3+2k}(xi) = xi sum_j TO[j][k] z^j (j <= k + 1): two constants per
    // point pair instead of ME + MO + 1
    double xi[FH > 0 ? FH : 1], z[FH > 0 ? FH : 1];
    double TE[ME + 1][ME];
    double TO[MO + 1][MO > 0 ? MO : 1];
    // Moments of the collocation tables against the Taylor series of the sine forcing about the element centre:
    // sum_j De[j][i] cos(th xi_j) = sum_m FE[m][i] th^(2m), sum_j Do[j][i] sin(th xi_j) = th sum_m FO[m][i] th^(2m)
    double FE[3][ME];
    double FO[3][MO > 0 ? MO : 1];
};

enum { STORE_DIRECT = 1, STORE_SMEM = 2, STORE_TMA = 3, STORE_TMA_ROWS = 4, STORE_COOP = 5 };
// STORE_TMA: F/16 boxes [kThreads][16] of the [E][F] view (128-byte lines at stride 8F bytes).
// STORE_TMA_ROWS: one box [kThreads * F/16][16] of the [E * F/16][16] view: a contiguous 8F * kThreads byte write.
// STORE_COOP (F = 32): the solving thread stages its coefficients (14 doubles) in shared memory; then the warp
//   evaluates cooperatively, lane = (element parity, fine-point pair): basis values are per-lane registers and
//   every STG.64 of the warp writes two complete 128-byte lines.  No TMA, no constant-bank operands.

constexpr int kThreads = 128;
#ifndef HFL_TMA_PER_WARP
#define HFL_TMA_PER_WARP 0
#endif
constexpr bool kTmaPerWarp = HFL_TMA_PER_WARP != 0;   // STORE_TMA: one bulk store per warp (32-row boxes) instead of per CTA
constexpr int kWarps = kThreads / 32;

constexpr int kCoopPitch = 14;   // doubles per staged element row: w[0..M), h, S, C (M <= 11); 14 keeps STS.128 conflict-free

template <int STORE>
__host__ __device__ constexpr int tile_bytes(int F) {
    return (STORE == STORE_TMA || STORE == STORE_TMA_ROWS) ? (F / 16) * 4096
           : (STORE == STORE_SMEM ? 32 * (F + 2) * 8 : (STORE == STORE_COOP ? 32 * kCoopPitch * 8 : 0));
}

// CTAs per SM the register budget is sized for: small systems fit 128 registers (4 CTAs = 16 warps)
// (fine grid only: 128 registers, 4 CTAs; with the fused error norms: 168 registers, 3 CTAs - at 2 CTAs / 184 registers it
// runs 0.62 instead of 0.55 ms; fine grid + coefficient output: 2 CTAs, 0.60 instead of 0.63 ms at 3)
#ifndef HFL_ERR_MINB
#define HFL_ERR_MINB 3
#endif
#ifndef HFL_PLAIN_MINB
#define HFL_PLAIN_MINB 4
#endif
__host__ __device__ constexpr int min_ctas(int M, bool err, bool coef) {
    return (M <= 10 && !err && !coef) ? HFL_PLAIN_MINB : (err ? HFL_ERR_MINB : (coef ? 2 : 3));
}

// LDL^T of a packed-lower PSD matrix in a FIXED (plan-supplied) pivot order with a skip rule: a pivot that
// is not above 2^-10 eps * (first pivot) is dropped (its unknown is set to zero), which gives the basic solution
// of the numerically rank-deficient dual system.  b is overwritten by z.  Returns the number of kept pivots.
template <int n>
__device__ __forceinline__ int ldl_solve_skip(double (&A)[n * (n + 1) / 2], double (&b)[n]) {
    double dinv[n];
    int kept = 0;
    const double tol = A[0] * (2.220446049250313e-16 / 1024.0);
#pragma unroll
    for (int j = 0; j < n; ++j) {
        const double piv = A[j * (j + 1) / 2 + j];
        const bool keep = piv > tol;
        const double r = keep ? fast_rcp(keep ? piv : 1.0) : 0.0;
        kept += keep ? 1 : 0;
        dinv[j] = r;
        double l[n];
#pragma unroll
        for (int i = j + 1; i < n; ++i) l[i] = A[i * (i + 1) / 2 + j] * r;
#pragma unroll
        for (int i = j + 1; i < n; ++i) {
#pragma unroll
            for (int k = j + 1; k <= i; ++k)
                A[i * (i + 1) / 2 + k] = fma(-l[i], A[k * (k + 1) / 2 + j], A[i * (i + 1) / 2 + k]);
        }
#pragma unroll
        for (int i = j + 1; i < n; ++i) A[i * (i + 1) / 2 + j] = l[i];
    }
#pragma unroll
    for (int i = 1; i < n; ++i) {
#pragma unroll
        for (int j = 0; j < i; ++j) b[i] = fma(-A[i * (i + 1) / 2 + j], b[j], b[i]);
    }
#pragma unroll
    for (int i = 0; i < n; ++i) b[i] *= dinv[i];
#pragma unroll
    for (int j = n - 2; j >= 0; --j) {
        double acc = b[j];
#pragma unroll
        for (int i = j + 1; i < n; ++i) acc = fma(-A[i * (i + 1) / 2 + j], b[i], acc);
        b[j] = (dinv[j] != 0.0) ? acc : 0.0;
    }
    return kept;
}

// COEF: the coefficient-output path is compiled in (a separate instantiation, so that the fine-grid-only kernel
// keeps its register budget).
// Dual form, everything that is not the moment form: right-hand side, gather into the pivot order, two LDL^T solves with
// the skip rule, w = C^T z.  The DMOM kernels keep it out of line (its 2 x 28 matrix entries would otherwise set their
// register budget) and get the result through the thread's shared-memory slots: slot[q * kThreads], q = 0.. MEA - 1 the
// even coefficients {w0, re[]}, then the odd ones {w1, ro[]}.  Returns the number-of-pivots test (false: fall back).
template <int M, int NHD>
__device__ __noinline__ bool dual_small_solve_slots(const PrimalArgs& a, const DualSmallTables<M, NHD>& dt, long long e, double xl,
                                                    double xr, double abar, double bbar, double* slot);

template <int M, int FH, bool ERR, int STORE, int NHD = 0, bool COEF = true, bool DMOM = false>
#ifndef HFL_DUAL_SMALL_MINB
#define HFL_DUAL_SMALL_MINB 3      // 168 registers, no spills since the Horner / Taylor epilogue (round 2): 0.645 against 0.72 ms per 1e7 elements
                                   // at 2 CTAs (254 registers); 4 CTAs (128 registers) spill and run at 0.89 ms.  Round 1's epilogue spilled at 3.
#endif
__global__ void __launch_bounds__(kThreads, NHD > 0 ? (DMOM ? (COEF ? 2 : 3) : HFL_DUAL_SMALL_MINB) : min_ctas(M, ERR, COEF))
// (DMOM at 4 CTAs / 128 registers spills 200 bytes and runs at 0.59 ms per 1e7 elements against 0.50 ms at 3 CTAs)
lssvr_element_kernel(const __grid_constant__ PrimalArgs a, const __grid_constant__ PrimalTables<M, FH> t,
              const __grid_constant__ CUtensorMap tmap, const __grid_constant__ DualSmallTables<M, NHD> dt) {
    constexpr int ME = n_even(M), MO = n_odd(M);
    constexpr int F = 2 * FH;
    constexpr int TILE = tile_bytes<STORE>(F);
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double* sDe = reinterpret_cast<double*>(smem_raw + kWarps * TILE);
    double* sDo = sDe + a.NH * ME;
    if (NHD == 0) {
        for (int i = threadIdx.x; i < a.NH * ME; i += kThreads) sDe[i] = a.De[i];
        for (int i = threadIdx.x; i < a.NH * MO; i += kThreads) sDo[i] = a.Do[i];
        __syncthreads();
    }
    double* sPerm = sDe;   // dual form: [2 (NHD + 1)][kThreads] per-thread slots for the pivot-order gather
    // coefficient staging: [kWarps][32][M] so that the warp writes its 32 x M block of d_coef with coalesced stores
    double* sCoef = sDe + (NHD > 0 ? 2 * (NHD + 1) * kThreads : a.NH * (ME + MO));

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    unsigned char* tile_ptr = smem_raw + warp * TILE;
    const uint32_t tile_s = smem_u32(tile_ptr);
    const bool do_fine = (a.fine != nullptr) || ERR;
    // STORE_COOP: basis values of this lane's fine-point pair (lane & 15) live in registers for the whole kernel
    double ltE[ME], ltO[MO + 1];
    if (STORE == STORE_COOP) {
#pragma unroll
        for (int k = 0; k < ME; ++k) ltE[k] = t.fineE[(FH > 0 ? lane & (FH - 1) : 0)][k];
#pragma unroll
        for (int k = 0; k < MO + 1; ++k) ltO[k] = t.fineO[(FH > 0 ? lane & (FH - 1) : 0)][k];
    }

    double bcl = 0.0, bcr = 0.0, x_first = 0.0, x_last = 0.0, invL = 0.0;
    if (a.bc2 != nullptr) {
        bcl = a.bc2[0]; bcr = a.bc2[1];
        x_first = a.nodes[0]; x_last = a.nodes[a.E];
        invL = 1.0 / (x_last - x_first);
    }
    double acc_sq = 0.0, acc_mx = 0.0;
    int nfail = 0;
    bool store_pending = false;

    // CTA tile = kThreads consecutive elements (one per thread); warp w owns rows 32 w .. 32 w + 31 of it.
    // The loop bounds depend on blockIdx only, so the CTA-wide barriers of the TMA path are uniform.
    const long long nct = (a.E + kThreads - 1) / kThreads;
    long long ct = blockIdx.x;
    // nodal data of the next tile are fetched while the current one computes (hides the DRAM latency
    // that 3-4 warps per scheduler cannot)
    double nxl = 0.0, nxr = 0.0, nul = 0.0, nur = 0.0;
    if (ct < nct) {
        const long long e0 = min(ct * kThreads + threadIdx.x, a.E - 1);
        nxl = __ldg(a.nodes + e0); nxr = __ldg(a.nodes + e0 + 1);
        nul = __ldg(a.u + e0); nur = __ldg(a.u + e0 + 1);
    }
    for (; ct < nct; ct += gridDim.x) {
        const long long wtile_e0 = ct * kThreads + warp * 32;      // first element of this warp's 32 rows
        const long long e_raw = wtile_e0 + lane;
        const bool valid = e_raw < a.E;
        const long long e = valid ? e_raw : a.E - 1;
        const double xl = nxl, xr = nxr;
        double ul = nul, ur = nur;
        if (ct + gridDim.x < nct) {
            const long long en = min((ct + gridDim.x) * kThreads + threadIdx.x, a.E - 1);
            nxl = __ldg(a.nodes + en); nxr = __ldg(a.nodes + en + 1);
            nul = __ldg(a.u + en); nur = __ldg(a.u + en + 1);
        }
        if (a.bc2 != nullptr) {
            ul += (bcl * (x_last - xl) + bcr * (xl - x_first)) * invL;
            ur += (bcl * (x_last - xr) + bcr * (xr - x_first)) * invL;
        }
        const double h = xr - xl;
        const double h2 = h * h;
        const double isig = 0.25 * h2;            // 1 / sigma, sigma = (2/h)^2
        const double tau = (h2 * h2) * a.c_tau;   // 1 / (gamma sigma^2)
        const double abar = 0.5 * (ul + ur), bbar = 0.5 * (ur - ul);

        double re[ME], ro[MO > 0 ? MO : 1];
        double S = 0.0, C = 0.0;   // sin / cos of k pi x_c
        bool ok = true;
        double w0, w1;
        if constexpr (NHD > 0 && DMOM) {
            // ---- DUAL form through the moment tables (see DualSmallTables); anything else out of line
            constexpr int MEA = ME + 1;
            if (a.forcing == HFL_FORCING_SINE && 0.5 * tau < dt.thr_same && fabs(0.5 * a.k_freq * h) < 0.0078125) {
                sincospi(a.k_freq * (0.5 * (xl + xr)), &S, &C);
                const double xb = 3.14159265358979323846 * (a.k_freq * h * a.cN);
                const double y = xb * xb, ak = isig * a.kk;
                const double fe = ak * S, fo = (ak * C) * xb;
                auto mom_e = [&](int q) {
                    return fma(abar, dt.mom_e[5][q], fe * fma(fma(fma(fma(dt.mom_e[4][q], y, dt.mom_e[3][q]), y, dt.mom_e[2][q]), y, dt.mom_e[1][q]), y, dt.mom_e[0][q]));
                };
                auto mom_o = [&](int q) {
                    return fma(bbar, dt.mom_o[5][q], fo * fma(fma(fma(fma(dt.mom_o[4][q], y, dt.mom_o[3][q]), y, dt.mom_o[2][q]), y, dt.mom_o[1][q]), y, dt.mom_o[0][q]));
                };
                w0 = mom_e(0);
                w1 = mom_o(0);
#pragma unroll
                for (int i = 0; i < ME; ++i) re[i] = mom_e(1 + i);
#pragma unroll
                for (int i = 0; i < MO; ++i) ro[i] = mom_o(1 + i);
            } else {
                if (ERR) sincospi(a.k_freq * (0.5 * (xl + xr)), &S, &C);
                double* slot = sPerm + threadIdx.x;
                ok = dual_small_solve_slots<M, NHD>(a, dt, e, xl, xr, abar, bbar, slot);
                w0 = slot[0];
                w1 = slot[MEA * kThreads];
#pragma unroll
                for (int i = 0; i < ME; ++i) re[i] = slot[(1 + i) * kThreads];
#pragma unroll
                for (int i = 0; i < MO; ++i) ro[i] = slot[(MEA + 1 + i) * kThreads];
                if (!ok) {
#pragma unroll
                    for (int i = 0; i < ME; ++i) re[i] = 0.0;
#pragma unroll
                    for (int i = 0; i < MO; ++i) ro[i] = 0.0;
                    w0 = abar; w1 = bbar;
                    if (valid) ++nfail;
                }
            }
        } else if constexpr (NHD > 0) {
            // ---- DUAL form, parity-split: two (NHD + 1)-unknown systems in the plan's static pivot order
            constexpr int NE = NHD + 1;
            double be[NE], bo[NE];
            if (a.forcing == HFL_FORCING_SINE || ERR) sincospi(a.k_freq * (0.5 * (xl + xr)), &S, &C);
            if (a.forcing == HFL_FORCING_SINE) {
                double sb, cb;
                sincospi_base(a.k_freq * h * a.cN, &sb, &cb);
                const double s2 = 2.0 * sb * cb, c2 = fma(-2.0 * sb, sb, 1.0);
                double s = sb, c = cb;                      // N even
                const double fE = isig * a.kk * S, fO = isig * a.kk * C;
#pragma unroll
                for (int j = 0; j < NHD; ++j) {
                    be[j] = fE * c;
                    bo[j] = fO * s;
                    rotate(s, c, s2, c2);
                }
            } else {
#pragma unroll
                for (int j = 0; j < NHD; ++j) {
                    const double fp = __ldg(a.f + (long long)(NHD + j) * a.E + e);
                    const double fm = __ldg(a.f + (long long)(NHD - 1 - j) * a.E + e);
                    be[j] = isig * (0.5 * (fp + fm));
                    bo[j] = isig * (0.5 * (fp - fm));
                }
            }
            be[NHD] = abar;
            bo[NHD] = bbar;
            double* slot = sPerm + threadIdx.x;
#pragma unroll
            for (int j = 0; j < NE; ++j) { slot[j * kThreads] = be[j]; slot[(NE + j) * kThreads] = bo[j]; }
#pragma unroll
            for (int k = 0; k < NE; ++k) { be[k] = slot[dt.perm_e[k] * kThreads]; bo[k] = slot[(NE + dt.perm_o[k]) * kThreads]; }
            const double th = 0.5 * tau;
            int kept;
            {
                double A[NE * (NE + 1) / 2];
#pragma unroll
                for (int i = 0; i < NE; ++i)
#pragma unroll
                    for (int j = 0; j <= i; ++j)
                        A[i * (i + 1) / 2 + j] = (i == j) ? fma(dt.je[i], th, dt.Ke[i * (i + 1) / 2 + j]) : dt.Ke[i * (i + 1) / 2 + j];
                kept = ldl_solve_skip<NE>(A, be);
                ok = kept >= 1;
            }
            {
                double A[NE * (NE + 1) / 2];
#pragma unroll
                for (int i = 0; i < NE; ++i)
#pragma unroll
                    for (int j = 0; j <= i; ++j)
                        A[i * (i + 1) / 2 + j] = (i == j) ? fma(dt.jo[i], th, dt.Ko[i * (i + 1) / 2 + j]) : dt.Ko[i * (i + 1) / 2 + j];
                kept = ldl_solve_skip<NE>(A, bo);
                ok = ok && kept >= 1;
            }
            // w = C^T z
            w0 = 0.0; w1 = 0.0;
#pragma unroll
            for (int i = 0; i < ME; ++i) re[i] = 0.0;
#pragma unroll
            for (int i = 0; i < MO; ++i) ro[i] = 0.0;
#pragma unroll
            for (int k = 0; k < NE; ++k) {
                w0 = fma(dt.Ce[k][0], be[k], w0);
                w1 = fma(dt.Co[k][0], bo[k], w1);
#pragma unroll
                for (int i = 0; i < ME; ++i) re[i] = fma(dt.Ce[k][1 + i], be[k], re[i]);
#pragma unroll
                for (int i = 0; i < MO; ++i) ro[i] = fma(dt.Co[k][1 + i], bo[k], ro[i]);
            }
            if (!ok) {
#pragma unroll
                for (int i = 0; i < ME; ++i) re[i] = 0.0;
#pragma unroll
                for (int i = 0; i < MO; ++i) ro[i] = 0.0;
                w0 = abar; w1 = bbar;
                if (valid) ++nfail;
            }
        } else {
        // ---- right-hand sides: re = tau abar 1 - isig De^T f_even, ro = tau bbar 1 - isig Do^T f_odd
#pragma unroll
        for (int i = 0; i < ME; ++i) re[i] = 0.0;
#pragma unroll
        for (int i = 0; i < MO; ++i) ro[i] = 0.0;
        if ((a.forcing == HFL_FORCING_SINE || ERR) && a.debug != 2) sincospi(a.k_freq * (0.5 * (xl + xr)), &S, &C);
        double sclE = 0.0, sclO = 0.0;
        if (a.debug == 2) {
        } else if (a.forcing == HFL_FORCING_SINE) {
            // f(x_c + (h/2) xi) = (k pi)^2 [S cos(th xi) + C sin(th xi)], th = k pi h / 2: the even part projects on De, the
            // odd part on Do.  |th| <= 2^-10 (any mesh worth timing): the projections from the moment tables (three Taylor
            // terms, truncation < 1e-20); otherwise the angle-addition rotation over the collocation pairs.
            const double th = a.hpk * h;
            if (fabs(th) <= 0.0009765625) {
                const double z = th * th;
#pragma unroll
                for (int i = 0; i < ME; ++i) re[i] = fma(fma(t.FE[2][i], z, t.FE[1][i]), z, t.FE[0][i]);
#pragma unroll
                for (int i = 0; i < MO; ++i) ro[i] = th * fma(fma(t.FO[2][i], z, t.FO[1][i]), z, t.FO[0][i]);
            } else {
                double sb, cb;
                sincospi_base(a.k_freq * h * a.cN, &sb, &cb);       // base angle k pi (h/2) / (N-1)
                const double s2 = 2.0 * sb * cb, c2 = fma(-2.0 * sb, sb, 1.0);
                double s = (a.N & 1) ? 0.0 : sb, c = (a.N & 1) ? 1.0 : cb;
#pragma unroll 1
                for (int j = 0; j < a.NH; ++j) {
#pragma unroll
                    for (int i = 0; i < ME; ++i) re[i] = fma(sDe[j * ME + i], c, re[i]);
#pragma unroll
                    for (int i = 0; i < MO; ++i) ro[i] = fma(sDo[j * MO + i], s, ro[i]);
                    rotate(s, c, s2, c2);
                }
            }
            sclE = -isig * a.kk * S;
            sclO = -isig * a.kk * C;
        } else {
            const int jp0 = a.N >> 1, jm0 = (a.N - 1) >> 1;   // first right / left sample of pair 0
            for (int j = 0; j < a.NH; ++j) {
                const double fp = __ldg(a.f + (long long)(jp0 + j) * a.E + e);
                const double fm = __ldg(a.f + (long long)(jm0 - j) * a.E + e);
                const double fe = 0.5 * (fp + fm), fo = 0.5 * (fp - fm);
#pragma unroll
                for (int i = 0; i < ME; ++i) re[i] = fma(sDe[j * ME + i], fe, re[i]);
#pragma unroll
                for (int i = 0; i < MO; ++i) ro[i] = fma(sDo[j * MO + i], fo, ro[i]);
            }
            sclE = -isig;
            sclO = -isig;
        }
        const double ta = tau * abar, tb = tau * bbar;
#pragma unroll
        for (int i = 0; i < ME; ++i) re[i] = fma(sclE, re[i], ta);
#pragma unroll
        for (int i = 0; i < MO; ++i) ro[i] = fma(sclO, ro[i], tb);

        // ---- per-element matrices tau (I + 1 1^T) + G and their LDL^T solves
        double Ae[ME * (ME + 1) / 2], Ao[MO > 0 ? MO * (MO + 1) / 2 : 1];
        const double tau2 = tau + tau;
#pragma unroll
        for (int i = 0; i < ME; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j)
                Ae[i * (i + 1) / 2 + j] = t.Ge[i * (i + 1) / 2 + j] + (i == j ? tau2 : tau);
#pragma unroll
        for (int i = 0; i < MO; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j)
                Ao[i * (i + 1) / 2 + j] = t.Go[i * (i + 1) / 2 + j] + (i == j ? tau2 : tau);
        if (a.debug != 2) {
            ok = ldl_solve<ME>(Ae, re);
            if (MO > 0) ok = ldl_solve<MO>(Ao, ro) && ok;
        }
        if (!ok) {   // P:171-176: fall back to the linear interpolant of the nodal values
#pragma unroll
            for (int i = 0; i < ME; ++i) re[i] = 0.0;
#pragma unroll
            for (int i = 0; i < MO; ++i) ro[i] = 0.0;
            if (valid) ++nfail;
        }
        w0 = abar; w1 = bbar;
#pragma unroll
        for (int i = 0; i < ME; ++i) w0 -= re[i];
#pragma unroll
        for (int i = 0; i < MO; ++i) w1 -= ro[i];

        }

        if (valid && a.status != nullptr) a.status[e] = ok ? 0 : 1;
        if (COEF && a.coef != nullptr) {
            // stage the thread's M coefficients (row pitch M; odd or even, 64-bit accesses stay conflict-light),
            // then the warp writes its contiguous 32 x M block of [E][M] with consecutive lanes on consecutive doubles
            double* cw = sCoef + warp * (32 * M);
            double* cp = cw + lane * M;
            cp[0] = w0;
            cp[1] = w1;
#pragma unroll
            for (int i = 0; i < ME; ++i) cp[2 + 2 * i] = re[i];
#pragma unroll
            for (int i = 0; i < MO; ++i) cp[3 + 2 * i] = ro[i];
            __syncwarp();
            const long long rows_here = min((long long)32, a.E - wtile_e0);
            double* g = a.coef + wtile_e0 * M;
#pragma unroll
            for (int q = 0; q < M; ++q) {
                const int idx = q * 32 + lane;
                if (idx < rows_here * M) g[idx] = cw[idx];
            }
            __syncwarp();
        }

        if (STORE == STORE_COOP && FH == 16 && do_fine) {
            // ---- stage {w, h, S, C}; then lane (hf, i) evaluates points FH + i and FH - 1 - i of element 2 s + hf
            double* crow = reinterpret_cast<double*>(tile_ptr) + lane * kCoopPitch;
            crow[0] = w0; crow[1] = w1;
#pragma unroll
            for (int i = 0; i < ME; ++i) crow[2 + 2 * i] = re[i];
#pragma unroll
            for (int i = 0; i < MO; ++i) crow[3 + 2 * i] = ro[i];
            crow[M] = h; crow[M + 1] = S; crow[M + 2] = C;
            __syncwarp();
            const int hf = lane >> 4, pi = lane & 15;
            const bool st = (a.fine != nullptr);
            const double wgt = (pi == FH - 1) ? 0.5 : 1.0;
            const double xi = ltO[0];
#pragma unroll 4
            for (int sidx = 0; sidx < 16; ++sidx) {
                const double* r = reinterpret_cast<const double*>(tile_ptr) + (2 * sidx + hf) * kCoopPitch;
                double Ee = r[0], Oo = r[1] * ltO[0];
#pragma unroll
                for (int k = 0; k < ME; ++k) Ee = fma(r[2 + 2 * k], ltE[k], Ee);
#pragma unroll
                for (int k = 0; k < MO; ++k) Oo = fma(r[3 + 2 * k], ltO[1 + k], Oo);
                const double up = Ee + Oo, um = Ee - Oo;
                const long long eo = wtile_e0 + 2 * sidx + hf;
                const bool ev = eo < a.E;
                if (st && ev) {
                    double* g = a.fine + eo * F;
                    if (a.debug == 3) {           // streaming (evict-first) stores, profiling aid
                        __stcs(g + FH + pi, up);
                        __stcs(g + FH - 1 - pi, um);
                    } else {
                        g[FH + pi] = up;
                        g[FH - 1 - pi] = um;
                    }
                }
                if (ERR) {
                    // exact = S cos(phi) +- C sin(phi), phi = k pi (h/2) xi: Taylor for |phi| <= 1/4 (|err| < 1e-14)
                    const double he = r[M], Se = r[M + 1], Ce = r[M + 2];
                    const double ph = (1.5707963267948966 * a.k_freq) * he * xi;
                    double sp, cp;
                    if (fabs(ph) <= 0.25) {
                        const double z = ph * ph;
                        sp = ph * fma(z, fma(z, fma(z, fma(z, 2.7557319223985893e-06, -1.984126984126984e-04),
                                                  8.333333333333333e-03), -1.6666666666666666e-01), 1.0);
                        cp = fma(z, fma(z, fma(z, fma(z, fma(z, -2.755731922398589e-07, 2.48015873015873e-05),
                                                     -1.388888888888889e-03), 4.1666666666666664e-02), -0.5), 1.0);
                    } else {
                        sincos(ph, &sp, &cp);
                    }
                    const double dE = fma(-Se, cp, Ee), dO = fma(-Ce, sp, Oo);
                    if (ev) {
                        acc_sq = fma(2.0 * wgt * he * (2.0 * a.cF), fma(dE, dE, dO * dO), acc_sq);
                        acc_mx = fmax(acc_mx, fabs(dE) + fabs(dO));
                    }
                }
            }
            __syncwarp();
        } else
        // ---- fine grid: u(+-xi_i) = Ee +- Oo; rows go to shared memory (or straight to global)
        if (FH > 0 && do_fine) {
            // even / odd parts as polynomials in z = xi^2: Ee = sum_j ae[j] z^j, Oo = xi sum_j ao[j] z^j (Horner)
            double ae[ME + 1], ao[MO + 1];
            ae[0] = w0; ao[0] = w1;
#pragma unroll
            for (int j = 1; j <= ME; ++j) ae[j] = 0.0;
#pragma unroll
            for (int j = 1; j <= MO; ++j) ao[j] = 0.0;
#pragma unroll
            for (int k = 0; k < ME; ++k)
#pragma unroll
                for (int j = 0; j <= k + 1; ++j) ae[j] = fma(t.TE[j][k], re[k], ae[j]);
#pragma unroll
            for (int k = 0; k < MO; ++k)
#pragma unroll
                for (int j = 0; j <= k + 1; ++j) ao[j] = fma(t.TO[j][k], ro[k], ao[j]);
            // exact solution at +-xi: S cos(th xi) +- C sin(th xi), th = k pi h / 2.  |th| <= 2^-10 (any mesh worth timing):
            // Taylor in z, truncation < 1e-17; otherwise a sincospi per point pair after the loop.
            double ce1 = 0.0, ce2 = 0.0, co0 = 0.0, co1 = 0.0, sq = 0.0, sq_last = 0.0;
            const double acc_mx0 = acc_mx;
            bool taylor = false;
            if (ERR) {
                const double th = a.hpk * h;
                taylor = fabs(th) <= 0.0009765625;
                const double q = -th * th;
                ce1 = 0.5 * (S * q); ce2 = (S * q) * (q * 4.1666666666666664e-02);
                co0 = C * th; co1 = (C * th) * (q * 1.6666666666666666e-01);
            }
            const bool st = (a.fine != nullptr);
            if ((STORE == STORE_TMA || STORE == STORE_TMA_ROWS) && st && store_pending) {
                if (kTmaPerWarp && STORE == STORE_TMA) {
                    // the warp's buffer is free once its issuing lane has seen its last bulk stores read it
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    __syncwarp();
                } else {
                    // the CTA buffer is free once the issuing thread has seen its last bulk store read it
                    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    __syncthreads();
                }
                store_pending = false;
            }
            // TMA layout: F/16 boxes of [rows][128 B], row = thread (CTA boxes of kThreads rows) or lane (warp boxes of 32
            // rows, kTmaPerWarp), 16-byte chunks XOR-swizzled by row & 7
            constexpr uint32_t kBoxBytes = (kTmaPerWarp && STORE == STORE_TMA) ? 32 * 128 : kThreads * 128;
            const uint32_t row_tma = (kTmaPerWarp && STORE == STORE_TMA) ? tile_s + lane * 128 : smem_u32(smem_raw) + threadIdx.x * 128;
            const uint32_t sw = (uint32_t)(lane & 7) << 4;
            const uint32_t row_sm = tile_s + lane * ((F + 2) * 8);
            double2* row_g = reinterpret_cast<double2*>(a.fine + e * F);
#pragma unroll
            for (int i = 0; i < FH; i += 2) {
                double up[2], um[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const double z = t.z[i + q], x = t.xi[i + q];
                    double Ee = ae[ME], Oo = ao[MO];
#pragma unroll
                    for (int j = ME - 1; j >= 0; --j) Ee = fma(Ee, z, ae[j]);
#pragma unroll
                    for (int j = MO - 1; j >= 0; --j) Oo = fma(Oo, z, ao[j]);
                    Oo *= x;
                    up[q] = Ee + Oo;
                    um[q] = Ee - Oo;
                    if (ERR) {
                        // errors at +-xi are dE +- dO (dE, dO = even / odd part of u - exact), so
                        // err+^2 + err-^2 = 2 (dE^2 + dO^2) and max(|err+|, |err-|) = |dE| + |dO|
                        const double dE = Ee - fma(fma(ce2, z, ce1), z, S), dO = fma(-x, fma(co1, z, co0), Oo);
                        const double q2 = fma(dE, dE, dO * dO);
                        if (i + q == FH - 1) sq_last = q2;     // the two end points carry half the trapezoid weight
                        sq += q2;
                        acc_mx = fmax(acc_mx, fabs(dE) + fabs(dO));
                    }
                }
                if (st) {
                    const int pp = (FH + i) >> 1;        // chunk holding points FH+i, FH+i+1
                    const int pm = (FH - 2 - i) >> 1;    // chunk holding points FH-2-i, FH-1-i
                    if (STORE == STORE_DIRECT) {
                        if (valid) {
                            row_g[pp] = make_double2(up[0], up[1]);
                            row_g[pm] = make_double2(um[1], um[0]);
                        }
                    } else if (STORE == STORE_SMEM) {
                        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(row_sm + pp * 16), "d"(up[0]), "d"(up[1]) : "memory");
                        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(row_sm + pm * 16), "d"(um[1]), "d"(um[0]) : "memory");
                    } else if (STORE == STORE_TMA_ROWS) {
                        // row of the [E * F/16][16] view = thread * (F/16) + (chunk >> 3); swizzle by that row & 7
                        constexpr int RPE = F / 16 > 0 ? F / 16 : 1;
                        const uint32_t base = smem_u32(smem_raw);
                        const uint32_t rp = threadIdx.x * RPE + (pp >> 3), rm = threadIdx.x * RPE + (pm >> 3);
                        const uint32_t ap = base + rp * 128 + ((((uint32_t)pp & 7) ^ (rp & 7)) << 4);
                        const uint32_t am = base + rm * 128 + ((((uint32_t)pm & 7) ^ (rm & 7)) << 4);
                        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(ap), "d"(up[0]), "d"(up[1]) : "memory");
                        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(am), "d"(um[1]), "d"(um[0]) : "memory");
                    } else {
                        const uint32_t ap = row_tma + (pp >> 3) * kBoxBytes + ((((uint32_t)pp & 7) << 4) ^ sw);
                        const uint32_t am = row_tma + (pm >> 3) * kBoxBytes + ((((uint32_t)pm & 7) << 4) ^ sw);
                        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(ap), "d"(up[0]), "d"(up[1]) : "memory");
                        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(am), "d"(um[1]), "d"(um[0]) : "memory");
                    }
                }
            }
            if (ERR && !taylor) {      // coarse element: exact values by sincospi, one point pair at a time
                sq = 0.0;
                acc_mx = acc_mx0;
#pragma unroll 1
                for (int i = 0; i < FH; ++i) {
                    const double z = t.z[i], x = t.xi[i];
                    double Ee = ae[ME], Oo = ao[MO];
#pragma unroll
                    for (int j = ME - 1; j >= 0; --j) Ee = fma(Ee, z, ae[j]);
#pragma unroll
                    for (int j = MO - 1; j >= 0; --j) Oo = fma(Oo, z, ao[j]);
                    Oo *= x;
                    double sp, cp;
                    sincospi((0.5 * a.k_freq) * h * x, &sp, &cp);
                    const double dE = fma(-S, cp, Ee), dO = fma(-C, sp, Oo);
                    const double q2 = fma(dE, dE, dO * dO);
                    if (i == FH - 1) sq_last = q2;
                    sq += q2;
                    acc_mx = fmax(acc_mx, fabs(dE) + fabs(dO));
                }
            }
            if (ERR && valid) acc_sq = fma(fma(2.0, sq, -sq_last), h * (2.0 * a.cF), acc_sq);
            if (st && STORE == STORE_SMEM) {
                __syncwarp();
                double2* g = reinterpret_cast<double2*>(a.fine + wtile_e0 * F);
#pragma unroll 4
                for (int it = 0; it < FH; ++it) {
                    const int idx = it * 32 + lane;       // 16-byte unit inside the 32 x F tile
                    constexpr int FHD = FH > 0 ? FH : 1;
                    const int r = idx / FHD, p = idx - r * FHD;
                    if (wtile_e0 + r < a.E) {
                        double2 v;
                        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(tile_s + r * ((F + 2) * 8) + p * 16));
                        g[idx] = v;
                    }
                }
                __syncwarp();
            }
            if (st && STORE == STORE_TMA_ROWS) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncthreads();
                if (threadIdx.x == 0 && a.debug != 1) {
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                                     reinterpret_cast<uint64_t>(&tmap)),
                                 "r"(0), "r"((int)(ct * kThreads * (F / 16))), "r"(smem_u32(smem_raw))
                                 : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                store_pending = true;
            }
            if (st && STORE == STORE_TMA && kTmaPerWarp) {
                // one bulk tensor store per 32-row box, issued by lane 0 of the warp that filled it: no CTA barrier, the
                // warps of a CTA drift apart freely.  warp_u is the warp index as a warp-uniform value (shuffle from lane
                // 0), which keeps the operands of the bulk store in uniform registers.
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0 && a.debug != 1) {
                    const int row0 = (int)(ct * kThreads) + warp_u * 32;
                    const uint32_t buf = smem_u32(smem_raw) + (uint32_t)warp_u * TILE;
#pragma unroll
                    for (int b = 0; b < F / 16; ++b) {
                        asm volatile(
                            "cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                                reinterpret_cast<uint64_t>(&tmap)),
                            "r"(b * 16), "r"(row0), "r"(buf + b * (32 * 128))
                            : "memory");
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                store_pending = true;
            }
            if (st && STORE == STORE_TMA && !kTmaPerWarp) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncthreads();
                if (threadIdx.x == 0 && a.debug != 1) {   // one bulk tensor store per box for the whole CTA tile; operands are CTA-uniform
                    const int row0 = (int)(ct * kThreads);
                    const uint32_t buf = smem_u32(smem_raw);
#pragma unroll
                    for (int b = 0; b < F / 16; ++b) {
                        asm volatile(
                            "cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                                reinterpret_cast<uint64_t>(&tmap)),
                            "r"(b * 16), "r"(row0), "r"(buf + b * (kThreads * 128))
                            : "memory");
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                store_pending = true;
            }
        }
    }
    if ((STORE == STORE_TMA || STORE == STORE_TMA_ROWS) && store_pending) {
        if (kTmaPerWarp && STORE == STORE_TMA) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            __syncwarp();
        } else {
            if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            __syncthreads();
        }
    }
    if (a.err3 != nullptr) {
        const double wsq = warp_sum(acc_sq), wmx = warp_max(acc_mx);
        const double wf = warp_sum((double)nfail);
        if (lane == 0) {
            if (ERR) {
                atomicAdd(a.err3 + 0, wsq);
                atomic_max_nonneg(a.err3 + 1, wmx);
            }
            if (wf != 0.0) atomicAdd(a.err3 + 2, wf);
        }
    }
}


template <int M, int NHD>
__device__ __noinline__ bool dual_small_solve_slots(const PrimalArgs& a, const DualSmallTables<M, NHD>& dt, long long e, double xl,
                                                    double xr, double abar, double bbar, double* slot) {
    constexpr int NE = NHD + 1, ME = n_even(M), MO = n_odd(M), MEA = ME + 1;
    const double h = xr - xl, h2 = h * h;
    const double isig = 0.25 * h2, tau = (h2 * h2) * a.c_tau;
    double be[NE], bo[NE];
    if (a.forcing == HFL_FORCING_SINE) {
        double S, C, sb, cb;
        sincospi(a.k_freq * (0.5 * (xl + xr)), &S, &C);
        sincospi_base(a.k_freq * h * a.cN, &sb, &cb);
        const double s2 = 2.0 * sb * cb, c2 = fma(-2.0 * sb, sb, 1.0);
        double s = sb, c = cb;                      // N even
        const double fE = isig * a.kk * S, fO = isig * a.kk * C;
#pragma unroll
        for (int j = 0; j < NHD; ++j) {
            be[j] = fE * c;
            bo[j] = fO * s;
            rotate(s, c, s2, c2);
        }
    } else {
#pragma unroll
        for (int j = 0; j < NHD; ++j) {
            const double fp = __ldg(a.f + (long long)(NHD + j) * a.E + e);
            const double fm = __ldg(a.f + (long long)(NHD - 1 - j) * a.E + e);
            be[j] = isig * (0.5 * (fp + fm));
            bo[j] = isig * (0.5 * (fp - fm));
        }
    }
    be[NHD] = abar;
    bo[NHD] = bbar;
#pragma unroll
    for (int j = 0; j < NE; ++j) { slot[j * kThreads] = be[j]; slot[(NE + j) * kThreads] = bo[j]; }
#pragma unroll
    for (int k = 0; k < NE; ++k) { be[k] = slot[dt.perm_e[k] * kThreads]; bo[k] = slot[(NE + dt.perm_o[k]) * kThreads]; }
    const double th = 0.5 * tau;
    bool ok;
    {
        double A[NE * (NE + 1) / 2];
#pragma unroll
        for (int i = 0; i < NE; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j)
                A[i * (i + 1) / 2 + j] = (i == j) ? fma(dt.je[i], th, dt.Ke[i * (i + 1) / 2 + j]) : dt.Ke[i * (i + 1) / 2 + j];
        ok = ldl_solve_skip<NE>(A, be) >= 1;
    }
    {
        double A[NE * (NE + 1) / 2];
#pragma unroll
        for (int i = 0; i < NE; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j)
                A[i * (i + 1) / 2 + j] = (i == j) ? fma(dt.jo[i], th, dt.Ko[i * (i + 1) / 2 + j]) : dt.Ko[i * (i + 1) / 2 + j];
        ok = (ldl_solve_skip<NE>(A, bo) >= 1) && ok;
    }
    // w = C^T z, into the slots
    double w0 = 0.0, w1 = 0.0, re[ME], ro[MO > 0 ? MO : 1];
#pragma unroll
    for (int i = 0; i < ME; ++i) re[i] = 0.0;
#pragma unroll
    for (int i = 0; i < MO; ++i) ro[i] = 0.0;
#pragma unroll
    for (int k = 0; k < NE; ++k) {
        w0 = fma(dt.Ce[k][0], be[k], w0);
        w1 = fma(dt.Co[k][0], bo[k], w1);
#pragma unroll
        for (int i = 0; i < ME; ++i) re[i] = fma(dt.Ce[k][1 + i], be[k], re[i]);
#pragma unroll
        for (int i = 0; i < MO; ++i) ro[i] = fma(dt.Co[k][1 + i], bo[k], ro[i]);
    }
    slot[0] = w0;
    slot[MEA * kThreads] = w1;
#pragma unroll
    for (int i = 0; i < ME; ++i) slot[(1 + i) * kThreads] = re[i];
#pragma unroll
    for (int i = 0; i < MO; ++i) slot[(MEA + 1 + i) * kThreads] = ro[i];
    return ok;
}

// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = (EncodeTiledFn)p;
    return fn;
}

// [E][F] doubles, box = kThreads rows x 16 doubles (128 B inner extent, 128-byte swizzle)
static int make_fine_tensor_map(CUtensorMap* m, double* d_fine, long long E, int F, bool rows_view = false, int box_rows = kThreads) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return HFL_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)F, (cuuint64_t)E};
    cuuint64_t strides[1] = {(cuuint64_t)F * sizeof(double)};
    cuuint32_t box[2] = {16, (cuuint32_t)box_rows};
    if (rows_view) {   // [E * F/16][16]: every row is one 128-byte line, the box is a contiguous block
        dims[0] = 16; dims[1] = (cuuint64_t)E * (F / 16);
        strides[0] = 128;
        box[1] = (cuuint32_t)(kThreads * (F / 16));
    }
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d_fine, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return HFL_ERR_CUDA; }
    return HFL_OK;
}

// Horner tables of PrimalTables: monomial coefficients of P_n (long double recurrence, the entries are dyadic rationals:
// exact) by parity, and the fine points xi+_i with their squares.
template <int M, int FH>
static void fill_horner_tables(PrimalTables<M, FH>& t) {
    constexpr int ME = n_even(M), MO = n_odd(M);
    if (FH <= 0) return;
    long double mono[HFL_MAX_M + 2][HFL_MAX_M + 2];
    for (int n = 0; n < M; ++n)
        for (int d = 0; d <= M; ++d) mono[n][d] = 0.0L;
    mono[0][0] = 1.0L;
    if (M > 1) mono[1][1] = 1.0L;
    for (int n = 1; n + 1 < M; ++n)
        for (int d = 0; d <= n + 1; ++d)
            mono[n + 1][d] = ((2 * n + 1) * (d > 0 ? mono[n][d - 1] : 0.0L) - n * mono[n - 1][d]) / (long double)(n + 1);
    for (int k = 0; k < ME; ++k)
        for (int j = 0; j <= k + 1; ++j) t.TE[j][k] = (double)mono[2 + 2 * k][2 * j];
    for (int k = 0; k < MO; ++k)
        for (int j = 0; j <= k + 1; ++j) t.TO[j][k] = (double)mono[3 + 2 * k][2 * j + 1];
    for (int i = 0; i < FH; ++i) {
        const long double x = (long double)(2 * i + 1) / (long double)(2 * FH - 1);      // F = 2 FH points, xi+ ascending
        t.xi[i] = (double)x;
        t.z[i] = (double)(x * x);
    }
}

template <int M, int FH, bool ERR, int STORE, int NHD = 0, bool COEF = true, bool DMOM = false>
static int launch_fast(const hfl_plan* plan, const PrimalArgs& a, cudaStream_t stream,
                       const DualSmallTables<M, NHD>* dtp = nullptr) {
    constexpr int ME = n_even(M), MO = n_odd(M), F = 2 * FH;
    PrimalTables<M, FH> t;
    memset(&t, 0, sizeof(t));
    for (int i = 0; i < ME * (ME + 1) / 2; ++i) t.Ge[i] = plan->Ge[i];
    for (int i = 0; i < MO * (MO + 1) / 2; ++i) t.Go[i] = plan->Go[i];
    for (int i = 0; i < FH; ++i) {
        for (int k = 0; k < ME; ++k) t.fineE[i][k] = plan->fineE[(size_t)i * ME + k];
        for (int k = 0; k < MO + 1; ++k) t.fineO[i][k] = plan->fineO[(size_t)i * (MO + 1) + k];
    }
    {
        // Taylor moments of the collocation tables (see PrimalTables::FE): xi_j = the non-negative half of the N points
        const int N = plan->N, NH = plan->NH;
        for (int m = 0; m < 3; ++m) {
            long double fe = 1.0L, fo = 1.0L;
            for (int q = 1; q <= 2 * m; ++q) fe *= q;
            for (int q = 1; q <= 2 * m + 1; ++q) fo *= q;
            const long double sgn = (m & 1) ? -1.0L : 1.0L;
            for (int i = 0; i < ME; ++i) {
                long double acc = 0.0L;
                for (int j = 0; j < NH; ++j) {
                    const long double x = (long double)((N & 1) ? 2 * j : 2 * j + 1) / (long double)(N - 1);
                    acc += (long double)plan->De[(size_t)j * ME + i] * powl(x, 2 * m);
                }
                t.FE[m][i] = (double)(sgn * acc / fe);
            }
            for (int i = 0; i < MO; ++i) {
                long double acc = 0.0L;
                for (int j = 0; j < NH; ++j) {
                    const long double x = (long double)((N & 1) ? 2 * j : 2 * j + 1) / (long double)(N - 1);
                    acc += (long double)plan->Do[(size_t)j * MO + i] * powl(x, 2 * m + 1);
                }
                t.FO[m][i] = (double)(sgn * acc / fo);
            }
        }
    }
    fill_horner_tables<M, FH>(t);
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    if ((STORE == STORE_TMA || STORE == STORE_TMA_ROWS) && a.fine != nullptr) {
        int rc = make_fine_tensor_map(&tmap, a.fine, a.E, F, STORE == STORE_TMA_ROWS,
                                      (kTmaPerWarp && STORE == STORE_TMA) ? 32 : kThreads);
        if (rc != HFL_OK) return rc;
    }
    auto kern = lssvr_element_kernel<M, FH, ERR, STORE, NHD, COEF, DMOM>;
    DualSmallTables<M, NHD> dt;
    if (dtp) dt = *dtp; else memset(&dt, 0, sizeof(dt));
    const size_t smem = (size_t)kWarps * tile_bytes<STORE>(F) +
                        (NHD > 0 ? (size_t)2 * (NHD + 1) * kThreads : (size_t)a.NH * (ME + MO)) * sizeof(double) +
                        (a.coef != nullptr ? (size_t)kWarps * 32 * M * sizeof(double) : 0);
    HFL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    HFL_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem));
    if (per_sm < 1) per_sm = 1;
    long long grid = (a.E + kThreads - 1) / kThreads;
    const long long cap = (long long)sm_count() * per_sm;
    if (grid > cap) grid = cap;
    kern<<<(unsigned)grid, kThreads, smem, stream>>>(a, t, tmap, dt);
    count_launch();
    HFL_CUDA_CHECK(cudaGetLastError());
    return HFL_OK;
}


}  // namespace hfl
