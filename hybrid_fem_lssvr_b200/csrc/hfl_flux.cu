// K1, better-conditioned variant (HFL_COARSE_FLUX_SCAN): the same P1 equations (same rounded
// stiffness entries k_e and loads b_i as hfl_fem.cu) solved in first-order form.  With the element
// flux q_e = k_e (u_{e+1} - u_e) the rows  q_{i-1} - q_i = b_i  give
//     q_e = q_0 - B_e,  B_e = sum_{j=1..e} b_j,      u_i = u_L + q_0 C_i - D_i,
//     C_i = sum_{e<i} 1/k_e,  D_i = sum_{e<i} B_e / k_e,   q_0 = (u_R - u_L + D_{n-1}) / C_{n-1}
// i.e. one prefix scan of the triple (b, c, d) under the associative operator
//     (b1,c1,d1) o (b2,c2,d2) = (b1+b2, c1+c2, d1+d2+b1*c2).
// No 1/h^2 amplification: the round-off stays at the 1e-15 level where the assembled tridiagonal
// solve (any solver, CPU or GPU) loses cond(K) ~ n^2.
// Kernels: flux_tile_kernel (tile aggregates) -> flux_top_kernel (scan of aggregates, q_0)
// -> flux_apply_kernel (in-tile scan + write u).
#include "hfl_fem.cuh"

namespace hfl {

struct Tri { double b, c, d; };
__device__ __forceinline__ Tri tri_op(const Tri& x, const Tri& y) {   // x before y
    return Tri{x.b + y.b, x.c + y.c, x.d + y.d + x.b * y.c};
}

constexpr int FLUX_SCAN = 2 * EL_LEN;          // 3 * FT doubles of scan scratch after the two element arrays
constexpr int FLUX_SMEM = 2 * EL_LEN + 3 * FT;

// Tile = elements [P, P + FTS).  Element e contributes b_e (load of node e; 0 for e = 0), c = 1/k_e,
// d = B_e / k_e with B inclusive.  Thread t owns elements P + t*FS .. P + t*FS + FS - 1.  The element arrays are
// the ones of hfl_fem.cu (local element q <-> global element P - 1 + q, node load of local node m = q - 1).
__device__ __forceinline__ void tile_local(const FemArgs& a, long long P, double* sm, Tri (&inc)[FS], Tri& agg) {
    load_tile_elements(a, P, sm);
    agg = Tri{0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < FS; ++i) {
        const int q = threadIdx.x * FS + i + 1;          // local index of element P + t*FS + i
        const long long ge = P - 1 + q;
        Tri x{0.0, 0.0, 0.0};
        if (ge <= a.n - 2) {
            const double b = (ge >= 1) ? sm[SM_B + padi(q - 1)] : 0.0;      // load of node ge
            const double c = 1.0 / sm[SM_K + padi(q)];
            x = Tri{b, c, b * c};
        }
        inc[i] = x;
        agg = tri_op(agg, x);
    }
}

// Exclusive scan of one Tri per thread over the CTA (Hillis-Steele in shared memory).
template <int T>
__device__ __forceinline__ Tri cta_exclusive_scan(Tri v, double* sm, Tri& total) {
    const int t = threadIdx.x;
    double* sb = sm; double* sc = sm + T; double* sd = sm + 2 * T;
    __syncthreads();
    sb[t] = v.b; sc[t] = v.c; sd[t] = v.d;
    __syncthreads();
    Tri cur = v;
    for (int off = 1; off < T; off <<= 1) {
        Tri left{0.0, 0.0, 0.0};
        const bool has = t >= off;
        if (has) left = Tri{sb[t - off], sc[t - off], sd[t - off]};
        __syncthreads();
        if (has) cur = tri_op(left, cur);
        sb[t] = cur.b; sc[t] = cur.c; sd[t] = cur.d;
        __syncthreads();
    }
    total = Tri{sb[T - 1], sc[T - 1], sd[T - 1]};
    Tri ex{0.0, 0.0, 0.0};
    if (t > 0) ex = Tri{sb[t - 1], sc[t - 1], sd[t - 1]};
    return ex;
}

__global__ void __launch_bounds__(FT, 4) flux_tile_kernel(const FemArgs a_in, double* __restrict__ agg3) {
    extern __shared__ double sm[];
    const FemArgs a = select_rhs(a_in);
    agg3 += (size_t)blockIdx.y * a.ws_stride;
    Tri inc[FS], agg, total;
    tile_local(a, (long long)blockIdx.x * FTS, sm, inc, agg);
    cta_exclusive_scan<FT>(agg, sm + FLUX_SCAN, total);
    if (threadIdx.x == 0) {
        agg3[3 * (size_t)blockIdx.x + 0] = total.b;
        agg3[3 * (size_t)blockIdx.x + 1] = total.c;
        agg3[3 * (size_t)blockIdx.x + 2] = total.d;
    }
}

// One CTA: exclusive scan of the tile aggregates; out: prefix per tile [3 * nt], then q_0 and the sum of all loads.
__global__ void __launch_bounds__(TOPT) flux_top_kernel(const double* __restrict__ agg3, int nt, int S, double uL,
                                                        double uR, double* __restrict__ prefix, long long ws_stride) {
    extern __shared__ double sm[];
    const int t = threadIdx.x;
    agg3 += (size_t)blockIdx.y * ws_stride; prefix += (size_t)blockIdx.y * ws_stride;
    Tri acc{0.0, 0.0, 0.0};
    for (int i = 0; i < S; ++i) {
        const int c = t * S + i;
        if (c < nt) acc = tri_op(acc, Tri{agg3[3 * (size_t)c], agg3[3 * (size_t)c + 1], agg3[3 * (size_t)c + 2]});
    }
    Tri total;
    Tri run = cta_exclusive_scan<TOPT>(acc, sm, total);
    for (int i = 0; i < S; ++i) {
        const int c = t * S + i;
        if (c < nt) {
            prefix[3 * (size_t)c] = run.b; prefix[3 * (size_t)c + 1] = run.c; prefix[3 * (size_t)c + 2] = run.d;
            run = tri_op(run, Tri{agg3[3 * (size_t)c], agg3[3 * (size_t)c + 1], agg3[3 * (size_t)c + 2]});
        }
    }
    if (t == 0) {
        prefix[3 * (size_t)nt] = (uR - uL + total.d) / total.c;      // q_0
        prefix[3 * (size_t)nt + 1] = total.b;                         // B_{n-2}: q_{n-2} = q_0 - B_{n-2} (end-node reactions)
    }
}

__global__ void __launch_bounds__(FT, 4) flux_apply_kernel(const FemArgs a_in, const double* __restrict__ prefix, int nt,
                                                        double* __restrict__ u) {
    extern __shared__ double sm[];
    const int t = threadIdx.x;
    const FemArgs a = select_rhs(a_in);
    prefix += (size_t)blockIdx.y * a.ws_stride; u += (size_t)blockIdx.y * a.n;
    const long long P = (long long)blockIdx.x * FTS;
    Tri inc[FS], agg, total;
    tile_local(a, P, sm, inc, agg);
    Tri run = cta_exclusive_scan<FT>(agg, sm + FLUX_SCAN, total);
    const Tri base{prefix[3 * (size_t)blockIdx.x], prefix[3 * (size_t)blockIdx.x + 1], prefix[3 * (size_t)blockIdx.x + 2]};
    const double q0 = prefix[3 * (size_t)nt];
    run = tri_op(base, run);
    __syncthreads();
    double* stage = sm;   // element arrays are dead
#pragma unroll
    for (int i = 0; i < FS; ++i) {
        run = tri_op(run, inc[i]);                       // inclusive through element P + t*FS + i
        stage[padi(t * FS + i)] = a.uL + (q0 * run.c - run.d);   // = u at node (that element) + 1
    }
    __syncthreads();
    for (int m = t; m < FTS; m += FT) {
        const long long node = P + m + 1;
        if (node < a.n - 1) u[node] = stage[padi(m)];
    }
    if (blockIdx.x == 0 && t == 0) { u[0] = a.uL; u[a.n - 1] = a.uR; }
}

}  // namespace hfl

using namespace hfl;

int hfl_fem_flux_scan(const FemArgs& a, int R, double* d_u, void* d_ws, size_t ws_bytes, cudaStream_t s) {
    const long long nt = (a.n - 1 + FTS - 1) / FTS;   // tiles of elements
    if (nt > (long long)TOPT * TOP_MAX_CHUNK) {
        set_error("hfl_fem_p1_solve: %lld nodes exceed the single-call limit; split the mesh across GPUs", a.n);
        return HFL_ERR_UNSUPPORTED;
    }
    double* agg3 = reinterpret_cast<double*>(d_ws);
    double* prefix = agg3 + 3 * (size_t)nt;
    if ((size_t)(6 * nt + 2) * sizeof(double) > ws_bytes) {
        set_error("hfl_fem_p1_solve: workspace too small for the flux scan");
        return HFL_ERR_ARG;
    }
    const size_t smem = (size_t)FLUX_SMEM * sizeof(double);
    {
        HFL_CUDA_CHECK(cudaFuncSetAttribute(flux_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        HFL_CUDA_CHECK(cudaFuncSetAttribute(flux_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const int S = (int)((nt + TOPT - 1) / TOPT);
    const dim3 grid((unsigned)nt, (unsigned)R);
    flux_tile_kernel<<<grid, FT, smem, s>>>(a, agg3);
    flux_top_kernel<<<dim3(1, R), TOPT, 3 * TOPT * sizeof(double), s>>>(agg3, (int)nt, S, a.uL, a.uR, prefix, a.ws_stride);
    flux_apply_kernel<<<grid, FT, smem, s>>>(a, prefix, (int)nt, d_u);
    count_launch(3);
    HFL_CUDA_CHECK(cudaGetLastError());
    return HFL_OK;
}
