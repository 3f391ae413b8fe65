// Shared pieces of the coarse-solve kernels (hfl_fem.cu, hfl_flux.cu).
#pragma once
#include "hfl_device.cuh"

namespace hfl {

constexpr int FT = 256;           // threads per CTA
constexpr int FS = 8;             // nodes per thread chunk (head + 7 interior)
constexpr int FTS = FT * FS;      // nodes per tile
constexpr int TOPT = 1024;        // threads of the top-level CTA
constexpr int TOP_MAX_CHUNK = 64; // tile heads per top-level thread
constexpr int REC = 10;           // doubles per tile record

__host__ __device__ inline int padi(int i) { return i + (i >> 3); }

struct FemArgs {
    long long n;
    const double* nodes;
    double k;        // forcing frequency k
    double kpi;      // k pi
    double kp2;      // (k pi)^2
    double uL, uR;
    double gx0, gx1; // Gauss points on [0, 1]
};

// Rows of the level-0 system from shared memory: k[q] = stiffness entry of local element q (global element
// P - 1 + q), b[m] = load of local node m (global node P + m).  SPECIAL = the tile contains a Dirichlet node
// or padding past the mesh (first / last tile only); interior tiles take the branch-free path.
template <bool SPECIAL>
struct MeshRows {
    const double* k; const double* b;
    long long P, n; double uL, uR;
    __device__ __forceinline__ void get(int m, double& l, double& d, double& r, double& bo) const {
        if (SPECIAL) {
            const long long g = P + m;
            if (g >= n) { l = 0.0; d = 1.0; r = 0.0; bo = 0.0; return; }
            if (g == 0) { l = 0.0; d = 1.0; r = 0.0; bo = uL; return; }
            if (g == n - 1) { l = 0.0; d = 1.0; r = 0.0; bo = uR; return; }
        }
        const double kl = k[padi(m)], kr = k[padi(m + 1)];
        l = -kl; r = -kr; d = kl + kr;
        bo = b[padi(m)];
    }
};

// Rows of the top-level system, stored as four arrays in global memory.
struct ArrayRows {
    const double* l; const double* d; const double* r; const double* b; int count;
    __device__ __forceinline__ void get(int m, double& lo, double& di, double& ro, double& bo) const {
        if (m >= count) { lo = 0.0; di = 1.0; ro = 0.0; bo = 0.0; return; }
        lo = l[m]; di = d[m]; ro = r[m]; bo = b[m];
    }
};

// Stiffness entry and load contributions of one element, with the reference's rounding
// (no fused multiply-adds): P:125-136 through scikit-fem's quadrature loop.
__device__ __forceinline__ void element_terms(const FemArgs& a, double x0, double x1, double& k, double& Ls, double& Rs) {
    const double h = x1 - x0;
    const double invh = __ddiv_rn(1.0, h);
    const double gg = __dmul_rn(invh, invh);
    const double hw = 0.5 * h;                       // |detDF| * W_q
    const double kq = __dmul_rn(gg, hw);
    k = __dadd_rn(kq, kq);
    const double xq0 = __dadd_rn(__dmul_rn(h, a.gx0), x0), xq1 = __dadd_rn(__dmul_rn(h, a.gx1), x0);
    const double f0 = __dmul_rn(a.kp2, sinpi(__dmul_rn(a.k, xq0)));   // sin(k pi x) without the argument-reduction slow path
    const double f1 = __dmul_rn(a.kp2, sinpi(__dmul_rn(a.k, xq1)));
    Ls = __dadd_rn(__dmul_rn(__dmul_rn(f0, 1.0 - a.gx0), hw), __dmul_rn(__dmul_rn(f1, 1.0 - a.gx1), hw));
    Rs = __dadd_rn(__dmul_rn(__dmul_rn(f0, a.gx0), hw), __dmul_rn(__dmul_rn(f1, a.gx1), hw));
}

}  // namespace hfl
