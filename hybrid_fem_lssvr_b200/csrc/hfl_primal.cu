// K2 + K3 (+ fused K5): batched per-element primal LSSVR solve with fine-grid reconstruction.
//
// Replaces the serial loop P:151-170 over lssvr_primal (P:20-105) and the structured part of
// evaluate_solution (P:184-211).  The reference hands the QP
//     min 1/2 |w|^2 + gamma/2 |e|^2   s.t.  -u''(x_j) - f(x_j) + e_j = 0,  u(x_L) = u_L, u(x_R) = u_R
// to SLSQP; its unique minimiser is computed here in closed form, one element per thread:
//
//  * xi_j and the fine points are symmetric about the element centre and P_k has parity (-1)^k, so
//    with the two boundary rows eliminated (w_0 = (u_L+u_R)/2 - sum_even w_k, w_1 = (u_R-u_L)/2 -
//    sum_odd w_k) the normal equations split into an even and an odd SPD block
//        (tau (I + 1 1^T) + G_par) w_par = tau g_par 1 - (h^2/4) D_par^T f_par,   tau = h^4 / (16 gamma)
//    of sizes floor((M-1)/2) and floor((M-2)/2) (4 and 3 at M = 9).  G_par, D_par are
//    element-independent tables; the element enters through tau, h^2 and the data.
//  * each block is factorised per element (LDL^T in registers, reciprocal pivots), the factorisation is
//    NOT shared between elements even on a uniform mesh.
//  * forcing family (k pi)^2 sin(k pi x): f at the collocation points by the angle-addition rotation
//    from one sincospi of the centre and one of the base angle (even part ~ cos, odd part ~ sin).
//  * fine grid: u(+-xi) = E(xi) +- O(xi) with the basis values as immediate constant-bank operands
//    (kernel parameter block), rows staged in swizzled shared memory and written by TMA tensor stores.
#include "hfl_primal_dispatch.cuh"

namespace hfl {

// ---------------------------------------------------------------------------------------------
// Generic kernel: any M <= HFL_MAX_M, any N, any F (odd counts included).  Same algorithm with
// run-time loop bounds (per-thread work arrays live in local memory); used for shapes the
// specialised kernel is not instantiated for.  Stores go straight to global memory.
struct GenericTables {
    const double* De; const double* Do; const double* Ge; const double* Go;
    const double* fineE; const double* fineO;
    int M, ME, MO, FH;
};

__device__ inline bool ldl_solve_rt(int n, double* A, double* b, double* dinv) {
    bool ok = true;
    for (int j = 0; j < n; ++j) {
        const double piv = A[j * (j + 1) / 2 + j];
        ok = ok && (piv > 0.0);
        const double r = 1.0 / piv;
        dinv[j] = r;
        for (int i = j + 1; i < n; ++i) {
            const double l = A[i * (i + 1) / 2 + j] * r;
            for (int k = j + 1; k <= i; ++k) A[i * (i + 1) / 2 + k] -= l * A[k * (k + 1) / 2 + j];
        }
        for (int i = j + 1; i < n; ++i) A[i * (i + 1) / 2 + j] *= r;
    }
    for (int i = 1; i < n; ++i)
        for (int j = 0; j < i; ++j) b[i] -= A[i * (i + 1) / 2 + j] * b[j];
    for (int i = 0; i < n; ++i) b[i] *= dinv[i];
    for (int j = n - 2; j >= 0; --j)
        for (int i = j + 1; i < n; ++i) b[j] -= A[i * (i + 1) / 2 + j] * b[i];
    return ok;
}

__global__ void __launch_bounds__(128)
primal_generic_kernel(const PrimalArgs a, const GenericTables t, const bool want_err) {
    constexpr int MX = HFL_MAX_M / 2;          // max block size (15 at M = 32)
    const int ME = t.ME, MO = t.MO, M = t.M, F = a.F, FH = t.FH, N = a.N;
    double acc_sq = 0.0, acc_mx = 0.0;
    int nfail = 0;
    double bcl = 0.0, bcr = 0.0, x_first = 0.0, x_last = 0.0, invL = 0.0;
    if (a.bc2 != nullptr) {
        bcl = a.bc2[0]; bcr = a.bc2[1];
        x_first = a.nodes[0]; x_last = a.nodes[a.E];
        invL = 1.0 / (x_last - x_first);
    }
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < a.E;
         e += (long long)gridDim.x * blockDim.x) {
        const double xl = a.nodes[e], xr = a.nodes[e + 1];
        double ul = a.u[e], ur = a.u[e + 1];
        if (a.bc2 != nullptr) {
            ul += (bcl * (x_last - xl) + bcr * (xl - x_first)) * invL;
            ur += (bcl * (x_last - xr) + bcr * (xr - x_first)) * invL;
        }
        const double h = xr - xl, h2 = h * h, isig = 0.25 * h2, tau = (h2 * h2) * a.c_tau;
        const double abar = 0.5 * (ul + ur), bbar = 0.5 * (ur - ul);
        double re[MX], ro[MX], dinv[MX];
        double Ae[MX * (MX + 1) / 2], Ao[MX * (MX + 1) / 2];
        for (int i = 0; i < ME; ++i) re[i] = 0.0;
        for (int i = 0; i < MO; ++i) ro[i] = 0.0;
        double S = 0.0, C = 0.0;
        if (a.forcing == HFL_FORCING_SINE || want_err) sincospi(a.k_freq * (0.5 * (xl + xr)), &S, &C);
        double sclE, sclO;
        if (a.forcing == HFL_FORCING_SINE) {
            double sb, cb;
            sincospi(a.k_freq * h * a.cN, &sb, &cb);
            const double s2 = 2.0 * sb * cb, c2 = fma(-2.0 * sb, sb, 1.0);
            double s = (N & 1) ? 0.0 : sb, c = (N & 1) ? 1.0 : cb;
            for (int j = 0; j < a.NH; ++j) {
                for (int i = 0; i < ME; ++i) re[i] = fma(t.De[j * ME + i], c, re[i]);
                for (int i = 0; i < MO; ++i) ro[i] = fma(t.Do[j * MO + i], s, ro[i]);
                rotate(s, c, s2, c2);
            }
            sclE = -isig * a.kk * S;
            sclO = -isig * a.kk * C;
        } else {
            const int jp0 = N >> 1, jm0 = (N - 1) >> 1;
            for (int j = 0; j < a.NH; ++j) {
                const double fp = a.f[(long long)(jp0 + j) * a.E + e];
                const double fm = a.f[(long long)(jm0 - j) * a.E + e];
                const double fe = 0.5 * (fp + fm), fo = 0.5 * (fp - fm);
                for (int i = 0; i < ME; ++i) re[i] = fma(t.De[j * ME + i], fe, re[i]);
                for (int i = 0; i < MO; ++i) ro[i] = fma(t.Do[j * MO + i], fo, ro[i]);
            }
            sclE = -isig;
            sclO = -isig;
        }
        for (int i = 0; i < ME; ++i) re[i] = fma(sclE, re[i], tau * abar);
        for (int i = 0; i < MO; ++i) ro[i] = fma(sclO, ro[i], tau * bbar);
        for (int i = 0; i < ME; ++i)
            for (int j = 0; j <= i; ++j) Ae[i * (i + 1) / 2 + j] = t.Ge[i * (i + 1) / 2 + j] + (i == j ? 2.0 * tau : tau);
        for (int i = 0; i < MO; ++i)
            for (int j = 0; j <= i; ++j) Ao[i * (i + 1) / 2 + j] = t.Go[i * (i + 1) / 2 + j] + (i == j ? 2.0 * tau : tau);
        bool ok = ldl_solve_rt(ME, Ae, re, dinv);
        if (MO > 0) ok = ldl_solve_rt(MO, Ao, ro, dinv) && ok;
        if (!ok) {
            for (int i = 0; i < ME; ++i) re[i] = 0.0;
            for (int i = 0; i < MO; ++i) ro[i] = 0.0;
            ++nfail;
        }
        double w0 = abar, w1 = bbar;
        for (int i = 0; i < ME; ++i) w0 -= re[i];
        for (int i = 0; i < MO; ++i) w1 -= ro[i];
        if (a.status != nullptr) a.status[e] = ok ? 0 : 1;
        if (a.coef != nullptr) {
            double* cp = a.coef + e * M;
            cp[0] = w0; cp[1] = w1;
            for (int i = 0; i < ME; ++i) cp[2 + 2 * i] = re[i];
            for (int i = 0; i < MO; ++i) cp[3 + 2 * i] = ro[i];
        }
        if (F > 0 && (a.fine != nullptr || want_err)) {
            double sf = 0.0, cf = 1.0, s2f = 0.0, c2f = 1.0, sq = 0.0;
            if (want_err) {
                double sb, cb;
                sincospi(a.k_freq * h * a.cF, &sb, &cb);
                s2f = 2.0 * sb * cb; c2f = fma(-2.0 * sb, sb, 1.0);
                if (!(F & 1)) { sf = sb; cf = cb; }
            }
            const int ip0 = F >> 1, im0 = (F - 1) >> 1;
            for (int i = 0; i < FH; ++i) {
                double Ee = w0, Oo = w1 * t.fineO[i * (MO + 1)];
                for (int k = 0; k < ME; ++k) Ee = fma(re[k], t.fineE[i * ME + k], Ee);
                for (int k = 0; k < MO; ++k) Oo = fma(ro[k], t.fineO[i * (MO + 1) + 1 + k], Oo);
                const double up = Ee + Oo, um = Ee - Oo;
                if (a.fine != nullptr) {
                    a.fine[e * F + ip0 + i] = up;
                    a.fine[e * F + im0 - i] = um;
                }
                if (want_err) {
                    const double xe = S * cf, xo = C * sf;
                    const double ep = up - (xe + xo), em = um - (xe - xo);
                    const bool self = (F & 1) && i == 0;
                    const double wgt = (i == FH - 1) ? 0.5 : 1.0;
                    sq += self ? wgt * ep * ep : wgt * (ep * ep + em * em);
                    acc_mx = fmax(acc_mx, fmax(fabs(ep), fabs(em)));
                    rotate(sf, cf, s2f, c2f);
                }
            }
            if (want_err) acc_sq = fma(sq, h * (2.0 * a.cF), acc_sq);
        }
    }
    if (a.err3 != nullptr) {
        const double wsq = warp_sum(acc_sq), wmx = warp_max(acc_mx), wf = warp_sum((double)nfail);
        if ((threadIdx.x & 31) == 0) {
            if (want_err) {
                atomicAdd(a.err3 + 0, wsq);
                atomic_max_nonneg(a.err3 + 1, wmx);
            }
            if (wf != 0.0) atomicAdd(a.err3 + 2, wf);
        }
    }
}

int primal_dispatch_fh8(const hfl_plan* plan, const PrimalArgs& a, bool err, int store, cudaStream_t s);    // hfl_primal_f16.cu
int primal_dispatch_fh32(const hfl_plan* plan, const PrimalArgs& a, bool err, int store, cudaStream_t s);   // hfl_primal_f64.cu

}  // namespace hfl

using namespace hfl;

extern "C" int hfl_lssvr_primal_batch(const hfl_plan_t* plan, int64_t E, const double* d_nodes, const double* d_u,
                                      int forcing_kind, double k_freq, const double* d_f_samples,
                                      const double* d_bc2, double* d_coef, double* d_fine, int32_t* d_status,
                                      double* d_err3, void* stream) {
    HFL_REQUIRE(plan != nullptr, "hfl_lssvr_primal_batch: plan is NULL");
    { const int drc = plan_on_current_device(plan, "hfl_lssvr_primal_batch"); if (drc != HFL_OK) return drc; }
    HFL_REQUIRE(E >= 0, "hfl_lssvr_primal_batch: E < 0");
    if (E == 0) return HFL_OK;
    HFL_REQUIRE(d_nodes != nullptr && d_u != nullptr, "hfl_lssvr_primal_batch: d_nodes / d_u is NULL");
    HFL_REQUIRE(forcing_kind == HFL_FORCING_SINE || forcing_kind == HFL_FORCING_SAMPLES,
                "hfl_lssvr_primal_batch: unknown forcing_kind %d", forcing_kind);
    HFL_REQUIRE(forcing_kind != HFL_FORCING_SAMPLES || d_f_samples != nullptr,
                "hfl_lssvr_primal_batch: HFL_FORCING_SAMPLES needs d_f_samples");
    HFL_REQUIRE(d_fine == nullptr || plan->F >= 2, "hfl_lssvr_primal_batch: d_fine given but the plan has F = 0");
    HFL_REQUIRE(E < (1LL << 31) - 64, "hfl_lssvr_primal_batch: E too large for one call");
    const bool want_err = (d_err3 != nullptr) && plan->F >= 2;
    const double pi = 3.14159265358979323846;
    PrimalArgs a;
    a.E = E; a.nodes = d_nodes; a.u = d_u; a.f = d_f_samples; a.bc2 = d_bc2;
    a.coef = d_coef; a.fine = d_fine; a.status = d_status; a.err3 = d_err3;
    a.De = plan->d_tables + plan->off_De; a.Do = plan->d_tables + plan->off_Do;
    a.N = plan->N; a.NH = plan->NH; a.F = plan->F; a.forcing = forcing_kind; a.debug = get_option_debug();
    a.k_freq = k_freq; a.kk = (k_freq * pi) * (k_freq * pi); a.hpk = 0.5 * pi * k_freq;
    a.c_tau = 1.0 / (16.0 * plan->gamma);
    a.cN = 0.5 / (double)(plan->N - 1);
    a.cF = plan->F >= 2 ? 0.5 / (double)(plan->F - 1) : 0.0;
    cudaStream_t s = (cudaStream_t)stream;

    int store = get_option_store();
    const bool aligned16 = (reinterpret_cast<uintptr_t>(d_fine) & 15) == 0;
    int rc = -1;
    if (plan->F == 32 && aligned16) {
        if (store == 0) store = STORE_TMA;
        rc = dispatch_M<16>(plan, a, want_err, store, s);
    } else if (plan->F == 16 && aligned16) {      // TMA-store instantiations only
        rc = primal_dispatch_fh8(plan, a, want_err, STORE_TMA, s);
    } else if (plan->F == 64 && aligned16) {
        rc = primal_dispatch_fh32(plan, a, want_err, STORE_TMA, s);
    } else if (plan->F == 0) {
        rc = dispatch_M<0>(plan, a, false, STORE_DIRECT, s);
    }
    if (rc >= 0) return rc;

    GenericTables t;
    t.De = plan->d_tables + plan->off_De; t.Do = plan->d_tables + plan->off_Do;
    t.Ge = plan->d_tables + plan->off_Ge; t.Go = plan->d_tables + plan->off_Go;
    t.fineE = plan->d_tables + plan->off_fineE; t.fineO = plan->d_tables + plan->off_fineO;
    t.M = plan->M; t.ME = plan->me; t.MO = plan->mo; t.FH = plan->FH;
    long long blocks = (E + 127) / 128;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    primal_generic_kernel<<<(unsigned)blocks, 128, 0, s>>>(a, t, want_err);
    count_launch();
    HFL_CUDA_CHECK(cudaGetLastError());
    return HFL_OK;
}
