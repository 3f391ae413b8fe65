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
#include "hfl_element_kernel.cuh"      // TMA tensor map of the [E][F] fine grid (make_fine_tensor_map), kThreads

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
static_assert(GT == kThreads, "the fine-grid tensor map boxes are kThreads rows");

// Horner form of the fine-grid evaluation, as in the Poisson element kernel (hfl_element_kernel.cuh): fine points
// xi+_i, z_i = xi_i^2, and the monomial coefficients of the Legendre basis by parity,
// P_{2k}(xi) = sum_j TE[j][k] z^j (k = 0 .. ceil(M/2) - 1), P_{2k+1}(xi) = xi sum_j TO[j][k] z^j.
template <int M, int FH>
struct GeneralFineTables {
    static constexpr int KE = (M + 1) / 2, KO = M / 2;
    double xi[FH], z[FH];
    double TE[KE][KE];
    double TO[KO > 0 ? KO : 1][KO > 0 ? KO : 1];
};

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


// F = 2 FH compile-time: fine rows evaluated in Horner form, staged in 128-byte-swizzled shared memory and written by one
// TMA bulk tensor store per 16-column box and CTA (the store path of the Poisson kernel).  Gram formation as above.
// CTAs per SM the register budget is sized for: the packed (M-2) x (M-2) Gram matrix lives in registers.  M = 9: 126
// registers without spills at 4 CTAs (1.92 ms per 1e7 elements; 2.27 ms at 3 CTAs / 160 registers, 3.5 ms at 5 with spills).
__host__ __device__ constexpr int general_minb(int M) { return M <= 9 ? 4 : (M <= 10 ? 3 : 2); }
// The coefficient samples a, a', c, f of collocation point j ([N][E] arrays: one 8-byte load per thread, array and point)
// arrive through a ring of GD asynchronous copy groups (cp.async, 8 bytes, each thread copying and reading only its own
// words: no barrier), GD points ahead of their use and running on into the next tile: ncu had 60 % of the stall samples
// of the one-point-ahead register prefetch waiting on these loads (DRAM latency ~1500 cycles against ~300 of work per point).
#ifndef HFL_GENERAL_GD
#define HFL_GENERAL_GD 5
#endif
constexpr int GD = HFL_GENERAL_GD;
__device__ __forceinline__ void cp_async8(uint32_t dst, const double* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
template <int M, int FH>
__global__ void __launch_bounds__(GT, general_minb(M)) general_fast_kernel(const GeneralArgs g, const __grid_constant__ GeneralFineTables<M, FH> ft,
                                                              const __grid_constant__ CUtensorMap tmap) {
    constexpr int m = M - 2, F = 2 * FH, KE = (M + 1) / 2, KO = M / 2;
    extern __shared__ __align__(1024) unsigned char gsm_raw[];
    constexpr int TILE = (F / 16) * GT * 128;            // F/16 boxes of [GT rows][128 B]
    double* sP0 = reinterpret_cast<double*>(gsm_raw + TILE);
    const int N = g.N;
    double* sP1 = sP0 + N * M; double* sP2 = sP1 + N * M;
    double* ring = sP2 + N * M;                          // [GD][4][GT]: a, a', c, f of GD points in flight
    for (int i = threadIdx.x; i < N * M; i += GT) { sP0[i] = g.P0[i]; sP1[i] = g.P1[i]; sP2[i] = g.P2[i]; }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t ring_s = smem_u32(ring) + threadIdx.x * 8;
    double bcl = 0.0, bcr = 0.0, x_first = 0.0, x_last = 0.0, invL = 0.0;
    if (g.bc2 != nullptr) {
        bcl = g.bc2[0]; bcr = g.bc2[1];
        x_first = g.nodes[0]; x_last = g.nodes[g.E];
        invL = 1.0 / (x_last - x_first);
    }
    bool store_pending = false;
    const long long nct = (g.E + GT - 1) / GT;
    // producer side of the ring: (tile, point) of the next group to issue; an empty group once the tiles are exhausted
    long long p_ct = blockIdx.x;
    int p_j = 0, p_slot = 0;
    auto issue_next = [&]() {
        if (p_ct < nct) {
            const long long pe = min(p_ct * GT + threadIdx.x, g.E - 1);
            const long long o = (long long)p_j * g.E + pe;
            const uint32_t dst = ring_s + p_slot * (4 * GT * 8);
            cp_async8(dst, g.a + o);
            if (g.da) cp_async8(dst + GT * 8, g.da + o);
            if (g.c) cp_async8(dst + 2 * GT * 8, g.c + o);
            cp_async8(dst + 3 * GT * 8, g.f + o);
            if (++p_j == N) { p_j = 0; p_ct += gridDim.x; }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        p_slot = p_slot + 1 == GD ? 0 : p_slot + 1;
    };
#pragma unroll
    for (int d = 0; d < GD; ++d) issue_next();
    int c_slot = 0;
    for (long long ct = blockIdx.x; ct < nct; ct += gridDim.x) {
        const long long e_raw = ct * GT + threadIdx.x;
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
            rhs[i] = tau * ((i & 1) ? bbar : abar);
#pragma unroll
            for (int j = 0; j <= i; ++j)
                H[i * (i + 1) / 2 + j] = (((i ^ j) & 1) == 0) ? (i == j ? 2.0 * tau : tau) : 0.0;
        }
#pragma unroll 1
        for (int j = 0; j < N; ++j) {
            asm volatile("cp.async.wait_group %0;" ::"n"(GD - 1) : "memory");      // the oldest group in flight has landed
            const double* slot = ring + c_slot * (4 * GT) + threadIdx.x;
            const double aj = slot[0];
            const double dj = (g.da ? slot[GT] : 0.0) * hh;          // a' h/2
            const double cj = (g.c ? slot[2 * GT] : 0.0) * isig;     // c h^2/4
            const double fj = slot[3 * GT] * isig;                   // f / sigma
            c_slot = c_slot + 1 == GD ? 0 : c_slot + 1;
            const double* p0 = sP0 + j * M; const double* p1 = sP1 + j * M; const double* p2 = sP2 + j * M;
            const double A0 = cj;
            const double A1 = fma(cj, p0[1], -dj);
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
            issue_next();       // into the slot just consumed (its values are in registers by now)
        }
        const bool ok = ldl_solve<m>(H, rhs);
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
            double ae[KE], ao[KO > 0 ? KO : 1];
#pragma unroll
            for (int j = 0; j < KE; ++j) ae[j] = 0.0;
#pragma unroll
            for (int j = 0; j < KO; ++j) ao[j] = 0.0;
#pragma unroll
            for (int k = 0; k < KE; ++k)
#pragma unroll
                for (int j = 0; j <= k; ++j) ae[j] = fma(ft.TE[j][k], w[2 * k], ae[j]);
#pragma unroll
            for (int k = 0; k < KO; ++k)
#pragma unroll
                for (int j = 0; j <= k; ++j) ao[j] = fma(ft.TO[j][k], w[2 * k + 1], ao[j]);
            if (store_pending) {     // the CTA buffer is free once the issuing thread has seen its last bulk store read it
                if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncthreads();
                store_pending = false;
            }
            const uint32_t row_tma = smem_u32(gsm_raw) + threadIdx.x * 128, sw = (uint32_t)(lane & 7) << 4;
#pragma unroll
            for (int i = 0; i < FH; i += 2) {
                double up[2], um[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const double z = ft.z[i + q], x = ft.xi[i + q];
                    double Ee = ae[KE - 1], Oo = KO > 0 ? ao[KO > 0 ? KO - 1 : 0] : 0.0;
#pragma unroll
                    for (int j = KE - 2; j >= 0; --j) Ee = fma(Ee, z, ae[j]);
#pragma unroll
                    for (int j = KO - 2; j >= 0; --j) Oo = fma(Oo, z, ao[j]);
                    Oo *= x;
                    up[q] = Ee + Oo;
                    um[q] = Ee - Oo;
                }
                const int pp = (FH + i) >> 1, pm = (FH - 2 - i) >> 1;
                const uint32_t ap = row_tma + (pp >> 3) * (GT * 128) + ((((uint32_t)pp & 7) << 4) ^ sw);
                const uint32_t am = row_tma + (pm >> 3) * (GT * 128) + ((((uint32_t)pm & 7) << 4) ^ sw);
                asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(ap), "d"(up[0]), "d"(up[1]) : "memory");
                asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(am), "d"(um[1]), "d"(um[0]) : "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (threadIdx.x == 0) {
                const int row0 = (int)(ct * GT);
                const uint32_t buf = smem_u32(gsm_raw);
#pragma unroll
                for (int b = 0; b < F / 16; ++b) {
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                                     reinterpret_cast<uint64_t>(&tmap)),
                                 "r"(b * 16), "r"(row0), "r"(buf + b * (GT * 128))
                                 : "memory");
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            store_pending = true;
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (store_pending) {
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        __syncthreads();
    }
}

template <int M, int FH>
static int launch_general_fast(const GeneralArgs& g, cudaStream_t s) {
    constexpr int F = 2 * FH, KE = (M + 1) / 2, KO = M / 2;
    GeneralFineTables<M, FH> ft;
    memset(&ft, 0, sizeof(ft));
    long double mono[HFL_MAX_M + 2][HFL_MAX_M + 2];
    for (int n = 0; n < M; ++n)
        for (int d = 0; d <= M; ++d) mono[n][d] = 0.0L;
    mono[0][0] = 1.0L;
    if (M > 1) mono[1][1] = 1.0L;
    for (int n = 1; n + 1 < M; ++n)
        for (int d = 0; d <= n + 1; ++d)
            mono[n + 1][d] = ((2 * n + 1) * (d > 0 ? mono[n][d - 1] : 0.0L) - n * mono[n - 1][d]) / (long double)(n + 1);
    for (int k = 0; k < KE; ++k)
        for (int j = 0; j <= k; ++j) ft.TE[j][k] = (double)mono[2 * k][2 * j];
    for (int k = 0; k < KO; ++k)
        for (int j = 0; j <= k; ++j) ft.TO[j][k] = (double)mono[2 * k + 1][2 * j + 1];
    for (int i = 0; i < FH; ++i) {
        const long double x = (long double)(2 * i + 1) / (long double)(F - 1);
        ft.xi[i] = (double)x;
        ft.z[i] = (double)(x * x);
    }
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    int rc = make_fine_tensor_map(&tmap, g.fine, g.E, F);
    if (rc != HFL_OK) return rc;
    auto kern = general_fast_kernel<M, FH>;
    const size_t smem = (size_t)(F / 16) * GT * 128 + ((size_t)3 * g.N * M + (size_t)GD * 4 * GT) * sizeof(double);
    HFL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    HFL_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, GT, smem));
    if (per_sm < 1) per_sm = 1;
    long long grid = (g.E + GT - 1) / GT;
    const long long cap = (long long)sm_count() * per_sm;
    if (grid > cap) grid = cap;
    kern<<<(unsigned)grid, GT, smem, s>>>(g, ft, tmap);
    count_launch();
    HFL_CUDA_CHECK(cudaGetLastError());
    return HFL_OK;
}

// Any M <= HFL_MAX_M (and any N, F): the same algorithm with run-time loop bounds, work arrays in local memory.
__global__ void __launch_bounds__(GT) general_generic_kernel(const GeneralArgs g, int M) {
    constexpr int MX = HFL_MAX_M - 2;
    const int m = M - 2, N = g.N, F = g.F;
    double bcl = 0.0, bcr = 0.0, x_first = 0.0, x_last = 0.0, invL = 0.0;
    if (g.bc2 != nullptr) {
        bcl = g.bc2[0]; bcr = g.bc2[1];
        x_first = g.nodes[0]; x_last = g.nodes[g.E];
        invL = 1.0 / (x_last - x_first);
    }
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < g.E; e += (long long)gridDim.x * blockDim.x) {
        const double xl = g.nodes[e], xr = g.nodes[e + 1];
        double ul = g.u[e], ur = g.u[e + 1];
        if (g.bc2 != nullptr) {
            ul += (bcl * (x_last - xl) + bcr * (xl - x_first)) * invL;
            ur += (bcl * (x_last - xr) + bcr * (xr - x_first)) * invL;
        }
        const double h = xr - xl, hh = 0.5 * h, isig = 0.25 * h * h, tau = (isig * isig) * (16.0 * g.c_tau);
        const double abar = 0.5 * (ul + ur), bbar = 0.5 * (ur - ul);
        double H[MX * (MX + 1) / 2], rhs[MX], row[MX], dinv[MX];
        for (int i = 0; i < m; ++i) {
            rhs[i] = tau * ((i & 1) ? bbar : abar);
            for (int j = 0; j <= i; ++j)
                H[i * (i + 1) / 2 + j] = (((i ^ j) & 1) == 0) ? (i == j ? 2.0 * tau : tau) : 0.0;
        }
        for (int j = 0; j < N; ++j) {
            const long long o = (long long)j * g.E + e;
            const double aj = g.a[o], dj = (g.da ? g.da[o] : 0.0) * hh, cj = (g.c ? g.c[o] : 0.0) * isig, fj = g.f[o] * isig;
            const double* p0 = g.P0 + j * M; const double* p1 = g.P1 + j * M; const double* p2 = g.P2 + j * M;
            const double A0 = cj, A1 = fma(cj, p0[1], -dj);
            const double res = fj - A0 * abar - A1 * bbar;
            for (int i = 0; i < m; ++i) {
                const double Ak = fma(cj, p0[i + 2], fma(-dj, p1[i + 2], -aj * p2[i + 2]));
                row[i] = Ak - ((i & 1) ? A1 : A0);
            }
            for (int i = 0; i < m; ++i) {
                rhs[i] = fma(row[i], res, rhs[i]);
                for (int k = 0; k <= i; ++k) H[i * (i + 1) / 2 + k] = fma(row[i], row[k], H[i * (i + 1) / 2 + k]);
            }
        }
        bool ok = true;
        for (int j = 0; j < m; ++j) {
            const double piv = H[j * (j + 1) / 2 + j];
            ok = ok && (piv > 0.0);
            const double r = 1.0 / piv;
            dinv[j] = r;
            for (int i = j + 1; i < m; ++i) {
                const double l = H[i * (i + 1) / 2 + j] * r;
                for (int k = j + 1; k <= i; ++k) H[i * (i + 1) / 2 + k] -= l * H[k * (k + 1) / 2 + j];
            }
            for (int i = j + 1; i < m; ++i) H[i * (i + 1) / 2 + j] *= r;
        }
        for (int i = 1; i < m; ++i)
            for (int j = 0; j < i; ++j) rhs[i] -= H[i * (i + 1) / 2 + j] * rhs[j];
        for (int i = 0; i < m; ++i) rhs[i] *= dinv[i];
        for (int j = m - 2; j >= 0; --j)
            for (int i = j + 1; i < m; ++i) rhs[j] -= H[i * (i + 1) / 2 + j] * rhs[i];
        double w0 = abar, w1 = bbar;
        for (int i = 0; i < m; ++i) {
            if (!ok) rhs[i] = 0.0;
            if (i & 1) w1 -= rhs[i]; else w0 -= rhs[i];
        }
        if (g.status != nullptr) g.status[e] = ok ? 0 : 1;
        if (g.coef != nullptr) {
            g.coef[e * M] = w0; g.coef[e * M + 1] = w1;
            for (int i = 0; i < m; ++i) g.coef[e * M + 2 + i] = rhs[i];
        }
        if (g.fine != nullptr) {
            for (int i = 0; i < F; ++i) {
                const double* v = g.V + (size_t)i * M;
                double sacc = 0.0;
                for (int k = m - 1; k >= 0; --k) sacc = fma(rhs[k], v[k + 2], sacc);
                g.fine[e * F + i] = fma(w1, v[1], sacc) + w0;
            }
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
    const bool fast32 = d_fine != nullptr && plan->F == 32 && (reinterpret_cast<uintptr_t>(d_fine) & 15) == 0;
    switch (plan->M) {
#define HFL_CASE(mm) case mm: return fast32 ? launch_general_fast<mm, 16>(g, s) : launch_general<mm>(g, s);
        HFL_CASE(3) HFL_CASE(4) HFL_CASE(5) HFL_CASE(6) HFL_CASE(7) HFL_CASE(8) HFL_CASE(9) HFL_CASE(10)
        HFL_CASE(11) HFL_CASE(12)
#undef HFL_CASE
        default: {     // M = 13 .. HFL_MAX_M: run-time loop bounds, work arrays in local memory
            long long blocks = (E + GT - 1) / GT;
            const long long cap = (long long)sm_count() * 8;
            if (blocks > cap) blocks = cap;
            general_generic_kernel<<<(unsigned)blocks, GT, 0, s>>>(g, plan->M);
            count_launch();
            HFL_CUDA_CHECK(cudaGetLastError());
            return HFL_OK;
        }
    }
}
