// K4 fast path: the dual LSSVR system split by parity (see the header comment of hfl_dual.cu).
//
// With symmetric collocation points the (N+2) x (N+2) dual system decouples, in the variables
// s_j = alpha(+xi_j) + alpha(-xi_j), a_j = alpha(+xi_j) - alpha(-xi_j), beta_L +- beta_R, into an even and an odd
// block of N/2 + 1 unknowns each:
//     (C_par C_par^T + tau/2 J) z_par = [f_par / sigma; g_par],      w_par = C_par^T z_par,
//     C_even = [-P''_{0,2,4..}(xi+_j); 1 ... 1],  C_odd = [-P''_{1,3,5..}(xi+_j); 1 ... 1].
// For N <= 14 (7 x 7 blocks at the reference's N = 12) each block lives in registers, one element per thread, in
// the element kernel of hfl_element_kernel.cuh (template parameter NHD = N/2).  The rank-revealing pivot order
// is computed once per plan from C C^T (it is element independent as soon as tau is below eps |C C^T|, and any
// order is stable when tau is large because the matrix is then well conditioned); pivots below
// 2^-10 eps * (first pivot) are skipped, which yields the basic solution of the rank-deficient system.
#include "hfl_element_kernel.cuh"
#include <algorithm>
#include <cmath>

namespace hfl {

// Pivot order of diagonally pivoted Cholesky on the n x n PSD matrix K (long double); indices whose remaining
// diagonal is negligible are appended in natural order.
// Returns the number of pivots taken (the numerical rank at 1e-13 of the largest diagonal entry).
static int pivot_order(int n, std::vector<long double> K, int* perm) {
    std::vector<int> rem(n);
    for (int i = 0; i < n; ++i) rem[i] = i;
    long double dmax0 = 0.0L;
    for (int i = 0; i < n; ++i) dmax0 = std::max(dmax0, K[(size_t)i * n + i]);
    int cnt = 0;
    while (!rem.empty()) {
        int best = 0;
        for (size_t q = 1; q < rem.size(); ++q)
            if (K[(size_t)rem[q] * n + rem[q]] > K[(size_t)rem[best] * n + rem[best]]) best = (int)q;
        const int j = rem[best];
        const long double d = K[(size_t)j * n + j];
        if (!(d > 1e-13L * dmax0)) break;
        perm[cnt++] = j;
        rem.erase(rem.begin() + best);
        std::vector<long double> l(n);
        for (int i = 0; i < n; ++i) l[i] = K[(size_t)i * n + j] / sqrtl(d);
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < n; ++k) K[(size_t)i * n + k] -= l[i] * l[k];
    }
    const int rank = cnt;
    for (int r : rem) perm[cnt++] = r;
    return rank;
}

// Moment tables of one parity block (see DualSmallTables): G = C_P^T (C_P C_P^T)^-1 over the first `rank` pivots
// (Cholesky in long double), then mom[m][q] = sum_k G[q][k] a_m c_k^(2m) (even) or a_m c_k^(2m+1) (odd) over the
// collocation pivots, c_k = 2 j_k + 1, and mom[5][q] = G[q][constraint pivot].  K, C in pivot order.
static void moment_tables(int NE, int NHD, int ma, int rank, const std::vector<long double>& Kp, const std::vector<long double>& Cp,
                          const int* perm, bool odd, double* mom /* [6][ma] */) {
    const int r = rank;
    std::vector<long double> L((size_t)r * r, 0.0L);
    for (int j = 0; j < r; ++j) {
        long double d = Kp[(size_t)j * NE + j];
        for (int k = 0; k < j; ++k) d -= L[(size_t)j * r + k] * L[(size_t)j * r + k];
        const long double ljj = sqrtl(d);
        L[(size_t)j * r + j] = ljj;
        for (int i = j + 1; i < r; ++i) {
            long double v = Kp[(size_t)i * NE + j];
            for (int k = 0; k < j; ++k) v -= L[(size_t)i * r + k] * L[(size_t)j * r + k];
            L[(size_t)i * r + j] = v / ljj;
        }
    }
    // G[q][k] = sum_i C[i][q] Z[i][k], Z = A_PP^-1 (column k = solve for the unit vector e_k)
    std::vector<long double> G((size_t)ma * r, 0.0L), z(r);
    for (int k = 0; k < r; ++k) {
        for (int i = 0; i < r; ++i) {
            long double v = (i == k) ? 1.0L : 0.0L;
            for (int j = 0; j < i; ++j) v -= L[(size_t)i * r + j] * z[j];
            z[i] = v / L[(size_t)i * r + i];
        }
        for (int i = r - 1; i >= 0; --i) {
            long double v = z[i];
            for (int j = i + 1; j < r; ++j) v -= L[(size_t)j * r + i] * z[j];
            z[i] = v / L[(size_t)i * r + i];
        }
        for (int q = 0; q < ma; ++q) {
            long double acc = 0.0L;
            for (int i = 0; i < r; ++i) acc += Cp[(size_t)i * ma + q] * z[i];
            G[(size_t)q * r + k] = acc;
        }
    }
    // Taylor coefficients: cos x = sum a_m x^(2m), sin x = sum a_m x^(2m+1)
    const long double ae[5] = {1.0L, -1.0L / 2, 1.0L / 24, -1.0L / 720, 1.0L / 40320};
    const long double ao[5] = {1.0L, -1.0L / 6, 1.0L / 120, -1.0L / 5040, 1.0L / 362880};
    for (int q = 0; q < ma; ++q) {
        for (int m = 0; m < 6; ++m) mom[(size_t)m * ma + q] = 0.0;
        long double acc[5] = {0.0L, 0.0L, 0.0L, 0.0L, 0.0L}, con = 0.0L;
        for (int k = 0; k < r; ++k) {
            const int j = perm[k];
            if (j >= NHD) { con += G[(size_t)q * r + k]; continue; }
            const long double c = (long double)(2 * j + 1);
            long double pw = odd ? c : 1.0L;
            for (int m = 0; m < 5; ++m) {
                acc[m] += G[(size_t)q * r + k] * (odd ? ao[m] : ae[m]) * pw;
                pw *= c * c;
            }
        }
        for (int m = 0; m < 5; ++m) mom[(size_t)m * ma + q] = (double)acc[m];
        mom[(size_t)5 * ma + q] = (double)con;
    }
}

template <int M, int NHD>
static void build_tables(const hfl_plan* plan, DualSmallTables<M, NHD>& dt) {
    constexpr int NE = NHD + 1, MEA = n_even(M) + 1, MOA = n_odd(M) + 1;
    const int N = plan->N;
    // natural-order C matrices from the full table D2[N][M] = P_k''(xi_j): positive half rows N/2 + j
    std::vector<long double> Ce((size_t)NE * MEA), Co((size_t)NE * MOA);
    for (int j = 0; j < NHD; ++j) {
        for (int a = 0; a < MEA; ++a) Ce[(size_t)j * MEA + a] = -(long double)plan->D2[(size_t)(N / 2 + j) * M + 2 * a];
        for (int b = 0; b < MOA; ++b) Co[(size_t)j * MOA + b] = -(long double)plan->D2[(size_t)(N / 2 + j) * M + 2 * b + 1];
    }
    for (int a = 0; a < MEA; ++a) Ce[(size_t)NHD * MEA + a] = 1.0L;
    for (int b = 0; b < MOA; ++b) Co[(size_t)NHD * MOA + b] = 1.0L;
    auto gram = [&](const std::vector<long double>& C, int m) {
        std::vector<long double> K((size_t)NE * NE, 0.0L);
        for (int i = 0; i < NE; ++i)
            for (int j = 0; j < NE; ++j)
                for (int k = 0; k < m; ++k) K[(size_t)i * NE + j] += C[(size_t)i * m + k] * C[(size_t)j * m + k];
        return K;
    };
    std::vector<long double> Ke = gram(Ce, MEA), Ko = gram(Co, MOA);
    const int rank_e = pivot_order(NE, Ke, dt.perm_e);
    const int rank_o = pivot_order(NE, Ko, dt.perm_o);
    for (int i = 0; i < NE; ++i) {
        dt.je[i] = dt.perm_e[i] < NHD ? 1.0 : 0.0;
        dt.jo[i] = dt.perm_o[i] < NHD ? 1.0 : 0.0;
        for (int j = 0; j <= i; ++j) {
            dt.Ke[i * (i + 1) / 2 + j] = (double)Ke[(size_t)dt.perm_e[i] * NE + dt.perm_e[j]];
            dt.Ko[i * (i + 1) / 2 + j] = (double)Ko[(size_t)dt.perm_o[i] * NE + dt.perm_o[j]];
        }
        for (int a = 0; a < MEA; ++a) dt.Ce[i][a] = (double)Ce[(size_t)dt.perm_e[i] * MEA + a];
        for (int b = 0; b < MOA; ++b) dt.Co[i][b] = (double)Co[(size_t)dt.perm_o[i] * MOA + b];
    }
    // moment form of the tau = 0 solve (DMOM kernels), from the double-rounded tables the kernels use
    {
        std::vector<long double> Kpe((size_t)NE * NE), Kpo((size_t)NE * NE), Cpe((size_t)NE * MEA), Cpo((size_t)NE * MOA);
        long double kmin = 1e300L;
        for (int i = 0; i < NE; ++i) {
            for (int j = 0; j < NE; ++j) {
                const int lo = i > j ? j : i, hi = i > j ? i : j;
                Kpe[(size_t)i * NE + j] = (long double)dt.Ke[hi * (hi + 1) / 2 + lo];
                Kpo[(size_t)i * NE + j] = (long double)dt.Ko[hi * (hi + 1) / 2 + lo];
            }
            for (int a = 0; a < MEA; ++a) Cpe[(size_t)i * MEA + a] = (long double)dt.Ce[i][a];
            for (int b = 0; b < MOA; ++b) Cpo[(size_t)i * MOA + b] = (long double)dt.Co[i][b];
            if (dt.je[i] != 0.0) kmin = std::min(kmin, (long double)dt.Ke[i * (i + 1) / 2 + i]);
            if (dt.jo[i] != 0.0) kmin = std::min(kmin, (long double)dt.Ko[i * (i + 1) / 2 + i]);
        }
        moment_tables(NE, NHD, MEA, rank_e, Kpe, Cpe, dt.perm_e, false, &dt.mom_e[0][0]);
        moment_tables(NE, NHD, MOA, rank_o, Kpo, Cpo, dt.perm_o, true, &dt.mom_o[0][0]);
        dt.thr_same = (double)kmin * 5.551115123125783e-17;     // 2^-54
    }
}

template <int M>
static int launch_m(const hfl_plan* plan, const PrimalArgs& a, bool err, cudaStream_t s) {
    DualSmallTables<M, 6> dt;
    {
        // the tables depend on the plan alone: built once (long double, a few hundred microseconds), then copied
        std::lock_guard<std::mutex> guard(plan->scratch_mu);
        if (plan->dual_small_tables.size() != sizeof(dt)) {
            memset(&dt, 0, sizeof(dt));
            build_tables<M, 6>(plan, dt);
            plan->dual_small_tables.assign(reinterpret_cast<unsigned char*>(&dt), reinterpret_cast<unsigned char*>(&dt) + sizeof(dt));
        }
        memcpy(&dt, plan->dual_small_tables.data(), sizeof(dt));
    }
    if (!get_option_dual_reuse()) dt.thr_same = -1.0;
    // sine forcing: the kernels that take the moment form wherever an element allows it (everything else out of line);
    // sampled forcing: the kernels with the factorisation in line
    if (a.forcing == HFL_FORCING_SINE && get_option_dual_reuse()) {
        if (a.coef != nullptr)
            return err ? launch_fast<M, 16, true, STORE_TMA, 6, true, true>(plan, a, s, &dt)
                       : launch_fast<M, 16, false, STORE_TMA, 6, true, true>(plan, a, s, &dt);
#ifdef HFL_DMOM_PLAIN_KERNEL
        return err ? launch_fast<M, 16, true, STORE_TMA, 6, false, true>(plan, a, s, &dt)
                   : launch_fast<M, 16, false, STORE_TMA, 6, false, true>(plan, a, s, &dt);
#else
        // one instantiation serves both: without an accumulator the fused-norm kernel only skips its final reduction, and it
        // is the faster of the two (no spill around the out-of-line call: 0.50 against 0.53 ms per 1e7 elements)
        return launch_fast<M, 16, true, STORE_TMA, 6, false, true>(plan, a, s, &dt);
#endif
    }
    return err ? launch_fast<M, 16, true, STORE_TMA, 6>(plan, a, s, &dt)
               : launch_fast<M, 16, false, STORE_TMA, 6>(plan, a, s, &dt);
}

// Returns -1 when the shape is not covered (the caller then uses the team kernel of hfl_dual.cu).
int launch_dual_small(const hfl_plan* plan, long long E, const double* d_nodes, const double* d_u, int forcing_kind,
                      double k_freq, const double* d_f, const double* d_bc2, double* d_coef, double* d_fine,
                      int* d_status, double* d_err3, cudaStream_t s) {
    if (plan->N != 12 || plan->F != 32 || plan->M < 3 || plan->M > 14) return -1;
    if ((reinterpret_cast<uintptr_t>(d_fine) & 15) != 0) return -1;
    const double pi = 3.14159265358979323846;
    PrimalArgs a;
    memset(&a, 0, sizeof(a));
    a.E = E; a.nodes = d_nodes; a.u = d_u; a.f = d_f; a.bc2 = d_bc2;
    a.coef = d_coef; a.fine = d_fine; a.status = d_status; a.err3 = d_err3;
    a.De = nullptr; a.Do = nullptr;
    a.N = plan->N; a.NH = plan->NH; a.F = plan->F; a.forcing = forcing_kind; a.debug = 0;
    a.k_freq = k_freq; a.kk = (k_freq * pi) * (k_freq * pi); a.hpk = 0.5 * pi * k_freq;
    a.c_tau = 1.0 / (16.0 * plan->gamma);
    a.cN = 0.5 / (double)(plan->N - 1);
    a.cF = 0.5 / (double)(plan->F - 1);
    const bool err = d_err3 != nullptr;
    switch (plan->M) {
#define HFL_CASE(m) case m: return launch_m<m>(plan, a, err, s);
        HFL_CASE(3) HFL_CASE(4) HFL_CASE(5) HFL_CASE(6) HFL_CASE(7) HFL_CASE(8) HFL_CASE(9) HFL_CASE(10)
        HFL_CASE(11) HFL_CASE(12) HFL_CASE(13) HFL_CASE(14)
#undef HFL_CASE
        default: return -1;
    }
}

}  // namespace hfl
