// Per-element LSSVR for a general 1-D elliptic operator  L u = -(a u')' + c u = f  (SURVEY.md section 8f-2).
//
// The reference codes the Poisson residual only (P:43-45) while its README advertises elliptic problems in
// general (README.md:3).  Same QP as P:47-81 with PDE rows A[j,k] = (L phi_k)(x_j):
//     A[j,k] = -a_j scl^2 P_k''(xi_j) - a'_j scl P_k'(xi_j) + c_j P_k(xi_j).
// A depends on the element, so nothing is tabulated beyond the basis values: the Gram matrix is formed per element
// by N rank-1 updates.  Scaled by 1/sigma (sigma = scl^2) and with the two boundary rows eliminated
// (w_0 = (u_L+u_R)/2 - sum_even w_k, w_1 = (u_R-u_L)/2 - sum_odd w_k) the normal equations for v = w[2:] are
//     [tau (I + s_e s_e^T + s_o s_o^T) + At^T At] v = tau (abar s_e + bbar s_o) + At^T (f/sigma - A0 abar - A1 bbar),
//     At = Ah[:, 2:] - Ah[:, 0] s_e^T - Ah[:, 1] s_o^T,   Ah = A / sigma,   tau = h^4 / (16 gamma)
// (no parity split: a', c and a non-constant a break the symmetry).  One element per thread, the (M-2) x (M-2) SPD
// matrix packed in registers, LDL^T with reciprocal pivots; fine-grid rows staged through shared memory and written
// with coalesced 16-byte stores.  Coefficient / forcing samples are read as [N][E] (coalesced across elements).
#include "hfl_device.cuh"

namespace hfl {

struct GeneralArgs {
    long long E;
    const double* nodes; const double* u;
    const double* a; const double* da; const double* c; const double* f;   // [N][E]; da, c may be NULL (zero)
    const double* bc2;
    double* coef; double* fine; int* status;
    const double* P0; const double* P1; const double* P2;   // [N][M] basis values / derivatives at the collocation points
    const double* V;                                        // [F][M] basis values at the fine points
    int N, F;
    double c_tau;
};

constexpr int GT = 128;

template <int M>
__global__ void __launch_bounds__(GT, 3) general_kernel(const GeneralArgs g) {
    constexpr int m = M - 2;
    extern __shared__ __align__(16) double gsm[];
    const int N = g.N, F = g.F;
    double* sP0 = gsm; double* sP1 = sP0 + N * M; double* sP2 = sP1 + N * M; double* sV = sP2 + N * M;
    double* tile = sV + (F > 0 ? F * M : 0);            // [4 warps][32][F + 2]
    for (int i = threadIdx.x; i < N * M; i += GT) { sP0[i] = g.P0[i]; sP1[i] = g.P1[i]; sP2[i] = g.P2[i]; }
    for (int i = threadIdx.x; i < F * M; i += GT) sV[i] = g.V[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pitch = F + 2;
    double* wt = tile + (size_t)warp * 32 * pitch;
    double bcl = 0.0, bcr = 0.0, x_first = 0.0, x_last = 0.0, invL = 0.0;
    if (g.bc2 != nullptr) {
        bcl = g.bc2[0]; bcr = g.bc2[1];
        x_first = g.nodes[0]; x_last = g.nodes[g.E];
        invL = 1.0 / (x_last - x_first);
    }
    const long long nct = (g.E + GT - 1) / GT;
    for (long long ct = blockIdx.x; ct < nct; ct += gridDim.x) {
        const long long w_e0 = ct * GT + warp * 32;
        const long long e_raw = w_e0 + lane;
        const bool valid = e_raw < g.E;
        const long long e = valid ? e_raw : g.E - 1;
        const double xl = g.nodes[e], xr = g.nodes[e + 1];
        double ul = g.u[e], ur = g.u[e + 1];
        if (g.bc2 != nullptr) {
            ul += (bcl * (x_last - xl) + bcr * (xl - x_first)) * invL;
            ur += (bcl * (x_last - xr) + bcr * (xr - x_first)) * invL;
        }
        const double h = xr - xl, hh = 0.5 * h, isig = 0.25 * h * h, tau = (isig * isig) * (16.0 * g.c_tau);
        const double abar = 0.5 * (ul + ur), bbar = 0.5 * (ur - ul);
        double H[m * (m + 1) / 2], rhs[m];
#pragma unroll
        for (int i = 0; i < m; ++i) {
            rhs[i] = tau * ((i & 1) ? bbar : abar);          // unknown i is coefficient k = i + 2: even k <-> even i
#pragma unroll
            for (int j = 0; j <= i; ++j)
                H[i * (i + 1) / 2 + j] = (((i ^ j) & 1) == 0) ? (i == j ? 2.0 * tau : tau) : 0.0;
        }
        // samples of the next collocation point are fetched while the current rank-1 update runs
        double na = __ldg(g.a + e), nd = g.da ? __ldg(g.da + e) : 0.0, nc = g.c ? __ldg(g.c + e) : 0.0, nf = __ldg(g.f + e);
        for (int j = 0; j < N; ++j) {
            const double aj = na;
            const double dj = nd * hh;        // a' h/2
            const double cj = nc * isig;      // c h^2/4
            const double fj = nf * isig;      // f / sigma
            if (j + 1 < N) {
                const long long o = (long long)(j + 1) * g.E + e;
                na = __ldg(g.a + o); nf = __ldg(g.f + o);
                if (g.da) nd = __ldg(g.da + o);
                if (g.c) nc = __ldg(g.c + o);
            }
            const double* p0 = sP0 + j * M; const double* p1 = sP1 + j * M; const double* p2 = sP2 + j * M;
            const double A0 = cj;                                   // k = 0: P = 1, P' = P'' = 0
            const double A1 = fma(cj, p0[1], -dj);                  // k = 1: P = xi, P' = 1
            const double res = fj - A0 * abar - A1 * bbar;
            double row[m];
#pragma unroll
            for (int i = 0; i < m; ++i) {
                const double Ak = fma(cj, p0[i + 2], fma(-dj, p1[i + 2], -aj * p2[i + 2]));
                row[i] = Ak - ((i & 1) ? A1 : A0);
            }
#pragma unroll
            for (int i = 0; i < m; ++i) {
                rhs[i] = fma(row[i], res, rhs[i]);
#pragma unroll
                for (int k = 0; k <= i; ++k) H[i * (i + 1) / 2 + k] = fma(row[i], row[k], H[i * (i + 1) / 2 + k]);
            }
        }
        bool ok = ldl_solve<m>(H, rhs);
        double w[M];
        w[0] = abar; w[1] = bbar;
#pragma unroll
        for (int i = 0; i < m; ++i) {
            const double v = ok ? rhs[i] : 0.0;           // P:171-176 fallback: linear interpolant
            w[i + 2] = v;
            if (i & 1) w[1] -= v; else w[0] -= v;
        }
        if (valid && g.status != nullptr) g.status[e] = ok ? 0 : 1;
        if (valid && g.coef != nullptr) {
#pragma unroll
            for (int k = 0; k < M; ++k) g.coef[e * M + k] = w[k];
        }
        if (g.fine != nullptr) {
            double* rowp = wt + lane * pitch;
            for (int i = 0; i < F; ++i) {
                double s = 0.0;
#pragma unroll
                for (int k = M - 1; k >= 0; --k) s = fma(w[k], sV[i * M + k], s);
                rowp[i] = s;
            }
            __syncwarp();
            const long long rows_here = min((long long)32, g.E - w_e0);
            double* gout = g.fine + w_e0 * F;
            for (int idx = lane; idx < rows_here * F; idx += 32) {
                const int r = idx / F, i = idx - r * F;
                gout[idx] = wt[r * pitch + i];
            }
            __syncwarp();
        }
    }
}

template <int M>
static int launch_general(const GeneralArgs& g, cudaStream_t s) {
    const size_t smem = ((size_t)3 * g.N * M + (size_t)g.F * M + (size_t)4 * 32 * (g.F + 2)) * sizeof(double);
    HFL_CUDA_CHECK(cudaFuncSetAttribute(general_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    HFL_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, general_kernel<M>, GT, smem));
    if (per_sm < 1) per_sm = 1;
    long long grid = (g.E + GT - 1) / GT;
    const long long cap = (long long)sm_count() * per_sm;
    if (grid > cap) grid = cap;
    general_kernel<M><<<(unsigned)grid, GT, smem, s>>>(g);
    count_launch();
    HFL_CUDA_CHECK(cudaGetLastError());
    return HFL_OK;
}

}  // namespace hfl

using namespace hfl;

extern "C" int hfl_lssvr_general_batch(const hfl_plan_t* plan, int64_t E, const double* d_nodes, const double* d_u,
                                       const double* d_a, const double* d_da, const double* d_c, const double* d_f,
                                       const double* d_bc2, double* d_coef, double* d_fine, int32_t* d_status,
                                       void* stream) {
    HFL_REQUIRE(plan != nullptr, "hfl_lssvr_general_batch: plan is NULL");
    { const int drc = plan_on_current_device(plan, "hfl_lssvr_general_batch"); if (drc != HFL_OK) return drc; }
    HFL_REQUIRE(E >= 0, "hfl_lssvr_general_batch: E < 0");
    if (E == 0) return HFL_OK;
    HFL_REQUIRE(d_nodes && d_u && d_a && d_f, "hfl_lssvr_general_batch: d_nodes / d_u / d_a / d_f is NULL");
    HFL_REQUIRE(d_fine == nullptr || plan->F >= 2, "hfl_lssvr_general_batch: d_fine given but the plan has F = 0");
    GeneralArgs g;
    g.E = E; g.nodes = d_nodes; g.u = d_u; g.a = d_a; g.da = d_da; g.c = d_c; g.f = d_f; g.bc2 = d_bc2;
    g.coef = d_coef; g.fine = d_fine; g.status = d_status;
    g.P0 = plan->d_tables + plan->off_D0; g.P1 = plan->d_tables + plan->off_D1; g.P2 = plan->d_tables + plan->off_D2;
    g.V = plan->d_tables + plan->off_V;
    g.N = plan->N; g.F = d_fine ? plan->F : 0;
    g.c_tau = 1.0 / (16.0 * plan->gamma);
    cudaStream_t s = (cudaStream_t)stream;
    switch (plan->M) {
#define HFL_CASE(mm) case mm: return launch_general<mm>(g, s);
        HFL_CASE(3) HFL_CASE(4) HFL_CASE(5) HFL_CASE(6) HFL_CASE(7) HFL_CASE(8) HFL_CASE(9) HFL_CASE(10)
        HFL_CASE(11) HFL_CASE(12)
#undef HFL_CASE
        default:
            set_error("hfl_lssvr_general_batch: M=%d outside the instantiated range 3..12", plan->M);
            return HFL_ERR_UNSUPPORTED;
    }
}
