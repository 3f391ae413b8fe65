// Device helpers shared by the element kernels.
#pragma once
#include "hfl_common.cuh"

namespace hfl {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// 1/d to ~1 ulp: hardware seed (MUFU.RCP64H, ~2^-20) + two Newton steps on the FP64 pipe.
// Used for pivots, which are well inside the normal range (no denormal / inf handling needed).
__device__ __forceinline__ double fast_rcp(double d) {
#ifdef HFL_EXACT_RCP
    return __drcp_rn(d);
#endif
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// In-place LDL^T of a packed-lower SPD matrix (A[i(i+1)/2 + j], j <= i) followed by the solve
// of A x = b (b overwritten by x).  Returns false when a pivot is not positive (or NaN).
template <int n>
__device__ __forceinline__ bool ldl_solve(double (&A)[n * (n + 1) / 2 > 0 ? n * (n + 1) / 2 : 1],
                                          double (&b)[n > 0 ? n : 1]) {
    bool ok = true;
    double dinv[n > 0 ? n : 1];
#pragma unroll
    for (int j = 0; j < n; ++j) {
        const double piv = A[j * (j + 1) / 2 + j];
        ok = ok && (piv > 0.0);
        const double r = fast_rcp(piv);
        dinv[j] = r;
        double l[n > 0 ? n : 1];
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
    // forward: L y = b
#pragma unroll
    for (int i = 1; i < n; ++i) {
#pragma unroll
        for (int j = 0; j < i; ++j) b[i] = fma(-A[i * (i + 1) / 2 + j], b[j], b[i]);
    }
#pragma unroll
    for (int i = 0; i < n; ++i) b[i] *= dinv[i];
    // backward: L^T x = y
#pragma unroll
    for (int j = n - 2; j >= 0; --j) {
#pragma unroll
        for (int i = j + 1; i < n; ++i) b[j] = fma(-A[i * (i + 1) / 2 + j], b[i], b[j]);
    }
    return ok;
}

// sin(pi t), cos(pi t) for the per-element BASE angles (k h / (2 (N-1)) and the like), which are tiny on any
// mesh fine enough to matter for throughput: below 2^-7 a degree-7 / degree-8 Taylor polynomial is exact to
// rounding (truncation < 1e-19 relative); otherwise the library sincospi.
__device__ __forceinline__ void sincospi_base(double t, double* s, double* c) {
    if (fabs(t) < 0.0078125) {
        const double x = 3.14159265358979323846 * t, z = x * x;
        *s = x * fma(z, fma(z, fma(z, -1.984126984126984e-04, 8.333333333333333e-03), -1.6666666666666666e-01), 1.0);
        *c = fma(z, fma(z, fma(z, fma(z, 2.48015873015873e-05, -1.388888888888889e-03), 4.1666666666666664e-02), -0.5), 1.0);
    } else {
        sincospi(t, s, c);
    }
}

// Rotation of (s, c) = (sin t, cos t) by the angle whose sine / cosine are (s2, c2).
__device__ __forceinline__ void rotate(double& s, double& c, double s2, double c2) {
    const double sn = fma(s, c2, c * s2);
    const double cn = fma(c, c2, -(s * s2));
    s = sn;
    c = cn;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// max for non-negative doubles through the integer order of their bit patterns
__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
    atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

}  // namespace hfl
