// K3 (unstructured) and K5: point evaluation of the piecewise Legendre reconstruction and error norms.
#include "hfl_device.cuh"

namespace hfl {

// numpy.polynomial.legendre.legval as of numpy 2.3 (the reference pins no numpy version, README.md:34-36;
// 2.3.5 is what generated tests/golden): Clenshaw with the ratios (nd-1)/nd and (2nd-1)/nd formed first,
// operation for operation and without fused multiply-adds, so the value matches Legendre.__call__ of the
// reference (P:193) bit for bit given the same coefficients.
__device__ __forceinline__ double legval_numpy(double x, const double* __restrict__ c, int M) {
    if (M == 1) return __dadd_rn(c[0], __dmul_rn(0.0, x));
    if (M == 2) return __dadd_rn(c[0], __dmul_rn(c[1], x));
    int nd = M;
    double c0 = c[M - 2], c1 = c[M - 1];
    for (int i = 3; i <= M; ++i) {
        const double tmp = c0;
        nd = nd - 1;
        c0 = __dsub_rn(c[M - i], __dmul_rn(c1, __ddiv_rn((double)(nd - 1), (double)nd)));
        c1 = __dadd_rn(tmp, __dmul_rn(__dmul_rn(c1, x), __ddiv_rn((double)(2 * nd - 1), (double)nd)));
    }
    return __dadd_rn(c0, __dmul_rn(c1, x));
}

// evaluate_solution (P:184-211): first element j with nodes[j] <= x <= nodes[j+1] (a shared node goes
// to the left element, P:190-197), first / last element outside the mesh (P:199-209); argument
// mapping off + scl * x with numpy's mapparms rounding (polyutils.py:284-288).
__global__ void evaluate_points_kernel(long long E, const double* __restrict__ nodes, int M,
                                       const double* __restrict__ coef, long long P, const double* __restrict__ xq,
                                       double* __restrict__ out) {
    const double x_first = nodes[0];
    const double inv_len = 1.0 / (nodes[E] - x_first);
    // Is the mesh quasi-uniform?  Warp 0 compares 32 nodes spread over the mesh with the index a uniform mesh would give
    // them; the guess is used only when all of them land within 3 elements (otherwise the four probing loads around a
    // wrong guess are wasted: +50 % on a random-walk mesh).
    __shared__ int guess_ok;
    if (threadIdx.x < 32) {
        const long long is = (E * (long long)threadIdx.x) / 32 + (threadIdx.x & 1);
        const double t = (nodes[is] - x_first) * inv_len * (double)E;
        const bool near = fabs(t - (double)is) <= 3.0;
        const unsigned all = __ballot_sync(0xffffffffu, near);
        if (threadIdx.x == 0) guess_ok = (all == 0xffffffffu);
    }
    __syncthreads();
    const bool use_guess = guess_ok != 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (long long)gridDim.x * blockDim.x) {
        const double x = xq[i];
        // lower bound: smallest idx with nodes[idx] >= x  (idx in [0, E+1]).  The bracket [lo, hi] starts from the
        // position a uniform mesh would give (the reference's meshes are linspace, P:120) and gallops outwards for at
        // most three steps (offsets 1, 2, 4); if that does not bracket x the plain search over the whole array runs (its upper
        // levels are shared by all threads and stay in cache, which a one-sided search from the guess would lose).
        // Quasi-uniform mesh: 3-5 dependent loads instead of log2(E) = 24 (1e8 random points on 1e7 elements 10.2 ->
        // 8.2 ms, sorted points 5.0 -> 2.4 ms); any monotone mesh stays correct and costs at most 4 extra loads.
        long long lo = 0, hi = E + 1;
#ifndef HFL_EVAL_PLAIN_SEARCH
        if (use_guess) {
            const double t = (x - x_first) * inv_len;                 // NaN / out of range handled by the clamps
            const long long g = (t > 0.0) ? (long long)fmin(t * (double)E, (double)E) : 0;
            if (nodes[g] < x) {                                       // answer is to the right of g
                long long l = g, r = g + 1;
                for (int step = 1; step <= 4 && r <= E && nodes[r] < x; step <<= 1) { l = r; r += step; }
                if (!(r <= E && nodes[r] < x)) { lo = l + 1; hi = (r <= E) ? r : E + 1; }   // bracketed; else full search
            } else {                                                  // nodes[g] >= x: answer is g or to its left
                long long r = g, l = g - 1;
                for (int step = 1; step <= 4 && l >= 0 && !(nodes[l] < x); step <<= 1) { r = l; l -= step; }
                if (!(l >= 0 && !(nodes[l] < x))) { hi = r; lo = (l >= 0) ? l + 1 : 0; }    // bracketed; else full search
            }
        }
#endif
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if (nodes[mid] < x) lo = mid + 1; else hi = mid;
        }
        long long j = lo - 1;
        if (j < 0) j = 0;
        if (j > E - 1) j = E - 1;
        const double xmin = nodes[j], xmax = nodes[j + 1];
        const double oldlen = __dsub_rn(xmax, xmin);
        const double off = __ddiv_rn(__dsub_rn(__dmul_rn(xmax, -1.0), __dmul_rn(xmin, 1.0)), oldlen);
        const double scl = __ddiv_rn(2.0, oldlen);
        const double arg = __dadd_rn(off, __dmul_rn(scl, x));
        out[i] = legval_numpy(arg, coef + j * M, M);
    }
}

// Fine-grid error vs sin(k pi x): one warp per element row, lanes over the F points.
__global__ void error_fine_kernel(long long E, int F, const double* __restrict__ nodes,
                                  const double* __restrict__ fine, double k_freq, double* __restrict__ err3) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarp = ((long long)gridDim.x * blockDim.x) >> 5;
    double acc_sq = 0.0, acc_mx = 0.0;
    const double invF = 1.0 / (double)(F - 1);
    for (long long e = warp; e < E; e += nwarp) {
        const double xl = nodes[e], xr = nodes[e + 1];
        const double h = xr - xl, xc = 0.5 * (xl + xr);
        double sq = 0.0;
        for (int i = lane; i < F; i += 32) {
            const double xi = (double)(2 * i - (F - 1)) * invF;           // -1 .. 1
            const double ex = sinpi(k_freq * fma(0.5 * h, xi, xc));
            const double d = fine[e * F + i] - ex;
            const double w = (i == 0 || i == F - 1) ? 0.5 : 1.0;
            sq = fma(w * d, d, sq);
            acc_mx = fmax(acc_mx, fabs(d));
        }
        acc_sq = fma(sq, h * invF, acc_sq);
    }
    acc_sq = warp_sum(acc_sq);
    acc_mx = warp_max(acc_mx);
    if (lane == 0) {
        atomicAdd(err3 + 0, acc_sq);
        atomic_max_nonneg(err3 + 1, acc_mx);
    }
}

// Same norms, one element per thread (even F): the exact values at the F points come from one sincospi of the
// element centre, one of the base angle and the angle-addition rotation (as in the fused path of the element
// kernel) instead of F sinpi calls; the row is read with 16-byte loads and stays in L1 while it is consumed.
__global__ void __launch_bounds__(128) error_fine_rows_kernel(long long E, int F, const double* __restrict__ nodes,
                                                              const double* __restrict__ fine, double k_freq,
                                                              double* __restrict__ err3) {
    double acc_sq = 0.0, acc_mx = 0.0;
    const int FH = F >> 1;
    const double cF = 0.5 / (double)(F - 1);
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (long long)gridDim.x * blockDim.x) {
        const double xl = nodes[e], xr = nodes[e + 1], h = xr - xl;
        double S, C, sb, cb;
        sincospi(k_freq * (0.5 * (xl + xr)), &S, &C);
        sincospi_base(k_freq * h * cF, &sb, &cb);
        const double s2 = 2.0 * sb * cb, c2 = fma(-2.0 * sb, sb, 1.0);
        double sf = sb, cf = cb, sq = 0.0;
        const double2* row = reinterpret_cast<const double2*>(fine + e * F);
        for (int i = 0; i < FH; i += 2) {            // pairs i, i+1: points FH+i, FH+i+1 and FH-1-i, FH-2-i
            const double2 up = row[(FH + i) >> 1], um = row[(FH - 2 - i) >> 1];
            const double upv[2] = {up.x, up.y}, umv[2] = {um.y, um.x};
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const double xe = S * cf, xo = C * sf;
                const double ep = upv[q] - (xe + xo), em = umv[q] - (xe - xo);
                const double w = (i + q == FH - 1) ? 0.5 : 1.0;
                sq = fma(w, fma(ep, ep, em * em), sq);
                acc_mx = fmax(acc_mx, fmax(fabs(ep), fabs(em)));
                rotate(sf, cf, s2, c2);
            }
        }
        acc_sq = fma(sq, h * (2.0 * cF), acc_sq);
    }
    acc_sq = warp_sum(acc_sq);
    acc_mx = warp_max(acc_mx);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(err3 + 0, acc_sq);
        atomic_max_nonneg(err3 + 1, acc_mx);
    }
}

__global__ void error_nodal_kernel(long long n, const double* __restrict__ nodes, const double* __restrict__ u,
                                   double k_freq, double* __restrict__ err3) {
    double acc_sq = 0.0, acc_mx = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double x = nodes[i];
        const double xm = (i > 0) ? nodes[i - 1] : x, xp = (i < n - 1) ? nodes[i + 1] : x;
        const double d = u[i] - sinpi(k_freq * x);
        acc_sq = fma(0.5 * (xp - xm) * d, d, acc_sq);
        acc_mx = fmax(acc_mx, fabs(d));
    }
    acc_sq = warp_sum(acc_sq);
    acc_mx = warp_max(acc_mx);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(err3 + 0, acc_sq);
        atomic_max_nonneg(err3 + 1, acc_mx);
    }
}

}  // namespace hfl

using namespace hfl;

extern "C" int hfl_evaluate_points(int64_t E, const double* d_nodes, int M, const double* d_coef, int64_t P,
                                   const double* d_x, double* d_out, void* stream) {
    HFL_REQUIRE(E >= 1 && M >= 1 && M <= HFL_MAX_M, "hfl_evaluate_points: bad E or M");
    HFL_REQUIRE(P >= 0, "hfl_evaluate_points: P < 0");
    if (P == 0) return HFL_OK;
    HFL_REQUIRE(d_nodes && d_coef && d_x && d_out, "hfl_evaluate_points: NULL pointer");
    long long blocks = (P + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    evaluate_points_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(E, d_nodes, M, d_coef, P, d_x, d_out);
    count_launch();
    HFL_CUDA_CHECK(cudaGetLastError());
    return HFL_OK;
}

extern "C" int hfl_error_fine(int64_t E, int F, const double* d_nodes, const double* d_fine, double k_freq,
                              double* d_err3, void* stream) {
    HFL_REQUIRE(E >= 0 && F >= 2, "hfl_error_fine: bad E or F");
    if (E == 0) return HFL_OK;
    HFL_REQUIRE(d_nodes && d_fine && d_err3, "hfl_error_fine: NULL pointer");
    if (F % 4 == 0 && (reinterpret_cast<uintptr_t>(d_fine) & 15) == 0) {
        long long blocks = (E + 127) / 128;
        const long long cap = (long long)sm_count() * 16;
        if (blocks > cap) blocks = cap;
        error_fine_rows_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(E, F, d_nodes, d_fine, k_freq, d_err3);
    } else {
        long long blocks = (E * 32 + 255) / 256;
        const long long cap = (long long)sm_count() * 8;
        if (blocks > cap) blocks = cap;
        error_fine_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(E, F, d_nodes, d_fine, k_freq, d_err3);
    }
    count_launch();
    HFL_CUDA_CHECK(cudaGetLastError());
    return HFL_OK;
}

extern "C" int hfl_error_nodal(int64_t n, const double* d_nodes, const double* d_u, double k_freq, double* d_err3,
                               void* stream) {
    HFL_REQUIRE(n >= 2, "hfl_error_nodal: need at least 2 nodes");
    HFL_REQUIRE(d_nodes && d_u && d_err3, "hfl_error_nodal: NULL pointer");
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    error_nodal_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(n, d_nodes, d_u, k_freq, d_err3);
    count_launch();
    HFL_CUDA_CHECK(cudaGetLastError());
    return HFL_OK;
}
