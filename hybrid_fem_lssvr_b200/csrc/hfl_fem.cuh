// Shared pieces of the coarse-solve kernels (hfl_fem.cu, hfl_flux.cu).
#pragma once
#include "hfl_device.cuh"

namespace hfl {

#ifndef HFL_FEM_FT
#define HFL_FEM_FT 256
#endif
constexpr int FT = HFL_FEM_FT;    // threads per CTA
constexpr int FS = 8;             // nodes per thread chunk (head + 7 interior)
constexpr int FTS = FT * FS;      // nodes per tile
constexpr int TOPT = 1024;        // threads of the top-level CTA
constexpr int TOP_MAX_CHUNK = 64; // tile heads per top-level thread
constexpr int REC = 12;           // doubles per tile record

__host__ __device__ inline int padi(int i) { return i + (i >> 3); }

struct FemArgs {
    long long n;
    const double* nodes;
    double k;        // forcing frequency k
    double kpi;      // k pi
    double kp2;      // (k pi)^2
    double uL, uR;
    double gx0, gx1; // Gauss points on [0, 1]
    // general operator -(a u')' + c u = f: coefficient / forcing samples at the two Gauss points of every element,
    // [2][n-1] each (NULL = the reference's Poisson problem with the sine forcing family)
    const double* aq; const double* cq; const double* fq;
    // several right-hand sides of the same mesh in one launch (hfl_fem_p1_solve_multi): grid.y = R, right-hand side
    // r = blockIdx.y uses forcing frequency kfreqs[r], workspace slice r * ws_stride (doubles) and output row r * n
    const double* kfreqs; long long ws_stride;
    int exact_rowsum;   // HFL_COARSE_ASSEMBLED_EXACT: interior rows have row sum 0 (unrounded diagonal)
    // Taylor coefficients of the forcing about a reference point: (k pi)^2 sin(k pi (x_ref + dx)) = sum_j ta[j] dx^j T_j,
    // T_j = sin(k pi x_ref) for even j, cos(k pi x_ref) for odd j; ta[j] = (k pi)^(2 + j) * (+, +, -, -, ...) / j!
    double ta[10];
};

__host__ __device__ inline void fem_taylor_table(FemArgs& a) {
    double p = a.kp2, f = 1.0;
    for (int j = 0; j < 10; ++j) {
        if (j > 0) { p *= a.kpi; f *= (double)j; }
        a.ta[j] = ((j & 2) ? -p : p) / f;
    }
}

// The launch arguments specialised to this CTA's right-hand side (no-op for a single solve).
__device__ __forceinline__ FemArgs select_rhs(FemArgs a) {
    if (a.kfreqs != nullptr) {
        a.k = __ldg(a.kfreqs + blockIdx.y);
        a.kpi = __dmul_rn(a.k, 3.14159265358979323846);      // the same two roundings as the host does for one solve
        a.kp2 = __dmul_rn(a.kpi, a.kpi);
        fem_taylor_table(a);
    }
    return a;
}

// Row representation used by every elimination below: (l, sigma, r, b) with sigma = l + d + r the ROW SUM, the
// diagonal being recovered as d = sigma - l - r.  For this M-matrix l, r <= 0 and sigma is tiny (exactly the
// rounding residue of d = fl(k_left + k_right), obtained by an error-free TwoSum), so every diagonal that the
// eliminations form is a sum of same-signed terms: no cancellation (the GTH idea for M-matrices).  Plain (l, d, r)
// elimination loses ~cond * eps on this matrix (5e-12 at 1e4 nodes, 1e-10 at 1e5, like LAPACK or SuperLU); this form
// stays at ~1e-14 of the exact solution of the same rounded system.
//
// Rows of the level-0 system from shared memory: k[q] = stiffness entry of local element q (global element
// P - 1 + q), b[m] = load of local node m (global node P + m).  SPECIAL = the tile contains a Dirichlet node
// or padding past the mesh (first / last tile only); interior tiles take the branch-free path.
template <bool SPECIAL, bool GENERAL = false, bool EXACT = false>
struct MeshRows {
    const double* k; const double* b; const double* s;    // s: row sums per node (general operator only)
    long long P, n; double uL, uR;
    __device__ __forceinline__ void get(int m, double& l, double& sg, double& r, double& bo) const {
        if (SPECIAL) {
            const long long g = P + m;
            if (g >= n) { l = 0.0; sg = 1.0; r = 0.0; bo = 0.0; return; }
            if (g == 0) { l = 0.0; sg = 1.0; r = 0.0; bo = uL; return; }
            if (g == n - 1) { l = 0.0; sg = 1.0; r = 0.0; bo = uR; return; }
        }
        const double kl = k[padi(m)], kr = k[padi(m + 1)];
        l = -kl; r = -kr;
        bo = b[padi(m)];
        if (GENERAL) { sg = s[padi(m)]; return; }          // mass-matrix row sum, assembled without cancellation
        if (EXACT) { sg = 0.0; return; }                   // HFL_COARSE_ASSEMBLED_EXACT: unrounded diagonal kl + kr
        // d = fl(kl + kr) is the reference's assembled diagonal; kl + kr = d + err exactly (TwoSum), so sigma = -err
        const double d = __dadd_rn(kl, kr);
        const double t = __dsub_rn(d, kl);
        sg = -__dadd_rn(__dsub_rn(kl, __dsub_rn(d, t)), __dsub_rn(kr, t));
    }
};

// Rows of the top-level system, stored as four arrays (l, sigma, r, b) in global memory.
struct ArrayRows {
    const double* l; const double* sg; const double* r; const double* b; int count;
    __device__ __forceinline__ void get(int m, double& lo, double& so, double& ro, double& bo) const {
        if (m >= count) { lo = 0.0; so = 1.0; ro = 0.0; bo = 0.0; return; }
        lo = l[m]; so = sg[m]; ro = r[m]; bo = b[m];
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

// Taylor polynomial of the forcing about a reference point (the chunk's head node in hfl_fem.cu, the tile's first node in hfl_flux.cu): (k pi)^2 sin(k pi (x_ref + dx)) = sum_j tc[j] dx^j,
// tc[j] = ta[j] * (sin | cos)(k pi x_ref).  TIER 1: degree 9, |k pi dx| <= 2^-4 (truncation < 3e-19 of the amplitude);
// TIER 2: degree 4, |k pi dx| <= 2^-10 (truncation < 1e-17); TIER 3: degree 2, |k pi dx| <= 2^-17.5 (truncation < 4e-17).
template <int TIER>
struct ForcingPoly {
    static constexpr int DEG = (TIER == 3) ? 2 : ((TIER == 2) ? 4 : 9);
    double tc[DEG + 1];
    __device__ __forceinline__ void init(const FemArgs& a, double S, double C) {
#pragma unroll
        for (int j = 0; j <= DEG; ++j) tc[j] = a.ta[j] * ((j & 1) ? C : S);
    }
    __device__ __forceinline__ double eval(double dx) const {
        double p = tc[DEG];
#pragma unroll
        for (int j = DEG - 1; j >= 0; --j) p = fma(p, dx, tc[j]);
        return p;
    }
};
template <>
struct ForcingPoly<0> {
    __device__ __forceinline__ void init(const FemArgs&, double, double) {}
    __device__ __forceinline__ double eval(double) const { return 0.0; }
};

// element_terms with the forcing from a Taylor polynomial about xref (same k, load to rounding).
__device__ __forceinline__ void element_terms_taylor(const FemArgs& a, double x0, double x1, double xref, const ForcingPoly<1>& fp,
                                                     double& k, double& Ls, double& Rs) {
    const double h = x1 - x0;
    const double invh = __drcp_rn(h);
    const double gg = __dmul_rn(invh, invh);
    const double hw = 0.5 * h;
    const double kq = __dmul_rn(gg, hw);
    k = __dadd_rn(kq, kq);
    const double d0 = x0 - xref;
    const double f0 = fp.eval(fma(h, a.gx0, d0)), f1 = fp.eval(fma(h, a.gx1, d0));
    Ls = hw * fma(f0, 1.0 - a.gx0, f1 * (1.0 - a.gx1));
    Rs = hw * fma(f0, a.gx0, f1 * a.gx1);
}

// General operator: P1 stiffness of a (2-point Gauss), mass matrix of c, load of f, all from samples at the two
// Gauss points.  k = -(off-diagonal entry) = k_a - m_LR; sL, sR = the element's share of the row sums of its
// left / right node (the stiffness part cancels analytically: only the mass row sums c phi_i remain).
__device__ __forceinline__ void element_terms_general(const FemArgs& a, long long ge, double x0, double x1, double& k,
                                                      double& sL, double& sR, double& Ls, double& Rs) {
    const long long E = a.n - 1;
    const double h = x1 - x0, hw = 0.5 * h, invh = 1.0 / h;
    const double a0 = __ldg(a.aq + ge), a1 = __ldg(a.aq + E + ge);
    const double c0 = a.cq ? __ldg(a.cq + ge) : 0.0, c1 = a.cq ? __ldg(a.cq + E + ge) : 0.0;
    const double f0 = __ldg(a.fq + ge), f1 = __ldg(a.fq + E + ge);
    const double pL0 = 1.0 - a.gx0, pL1 = 1.0 - a.gx1, pR0 = a.gx0, pR1 = a.gx1;
    k = (a0 + a1) * (invh * invh) * hw - (c0 * pL0 * pR0 + c1 * pL1 * pR1) * hw;
    sL = (c0 * pL0 + c1 * pL1) * hw;
    sR = (c0 * pR0 + c1 * pR1) * hw;
    Ls = (f0 * pL0 + f1 * pL1) * hw;
    Rs = (f0 * pR0 + f1 * pR1) * hw;
}

// Shared-memory layout of the level-0 kernels (doubles): two padded arrays (element stiffness, node load);
// the exchange and PCR buffers of the reduce pass alias them once the chunk sweeps are done.
constexpr int EL_LEN = FTS + 1 + (FTS + 1) / 8 + 8;   // padded array of FTS + 1 entries
constexpr int SM_K = 0, SM_B = EL_LEN;
constexpr int SM_EX = 0;                               // 8 * FT exchange (aliases SM_K, after a barrier)
constexpr int SM_PCR = 0;                              // 2 * 7 * FT PCR buffers (alias the same space, later)
constexpr int SM_S = 2 * EL_LEN;                       // general operator only: row sums per node
__host__ __device__ constexpr int sm_uh(bool general) { return (general ? 3 : 2) * EL_LEN; }   // FT + 1 chunk-head values
__host__ __device__ constexpr int sm_total(bool general) { return sm_uh(general) + FT + 8; }
static_assert(2 * 7 * FT <= 2 * EL_LEN && 8 * FT <= 2 * EL_LEN, "exchange / PCR buffers must fit in the element arrays they alias");

// Element terms of local elements q = t, t + FT, ...; the nodes of the next element are fetched while the
// current one is being computed.  Node loads are accumulated as (0 + L_i) + R_{i-1}: every element first
// writes its left-node share, then (after a barrier) adds its right-node share.
template <bool GENERAL = false>
__device__ __forceinline__ void load_tile_elements(const FemArgs& a, long long P, double* sm) {
    auto fetch = [&](int q, double& x0, double& x1) {
        const long long ge = P - 1 + q;
        const bool ok = (q <= FTS) && ge >= 0 && ge <= a.n - 2;
        x0 = ok ? __ldg(a.nodes + ge) : 0.0;
        x1 = ok ? __ldg(a.nodes + ge + 1) : 1.0;
    };
    constexpr int NQ = (FTS + FT) / FT;     // elements per thread (the last one only for thread 0)
    double rs[NQ], ss[NQ];
    double nx0, nx1;
    fetch(threadIdx.x, nx0, nx1);
    // forcing from a Taylor polynomial about the tile's first node when the tile spans less than 1/16 rad (any mesh of
    // more than ~2e5 k nodes): one sincospi per thread instead of two sinpi per element
    bool taylor = false;
    double xref = 0.0;
    ForcingPoly<1> fp;
    if (!GENERAL) {
        const long long i0 = P > 0 ? P - 1 : 0, i1 = (P + FTS < a.n) ? P + FTS : a.n - 1;
        xref = __ldg(a.nodes + i0);
        taylor = fabs(a.kpi * (__ldg(a.nodes + i1) - xref)) <= 0.0625;
        if (taylor) {
            double S, C;
            sincospi(__dmul_rn(a.k, xref), &S, &C);
            fp.init(a, S, C);
        }
    }
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
        const int q = threadIdx.x + j * FT;
        rs[j] = 0.0; ss[j] = 0.0;
        if (q <= FTS) {
            const double x0 = nx0, x1 = nx1;
            fetch(q + FT, nx0, nx1);
            const long long ge = P - 1 + q;
            double k = 0.0, Ls = 0.0, Rs = 0.0, sL = 0.0, sR = 0.0;
            if (ge >= 0 && ge <= a.n - 2) {
                if (GENERAL) element_terms_general(a, ge, x0, x1, k, sL, sR, Ls, Rs);
                else if (taylor) element_terms_taylor(a, x0, x1, xref, fp, k, Ls, Rs);
                else element_terms(a, x0, x1, k, Ls, Rs);
            }
            sm[SM_K + padi(q)] = k;
            if (q >= 1) sm[SM_B + padi(q - 1)] = Ls;       // left node of local element q is local node q - 1
            if (GENERAL && q >= 1) sm[SM_S + padi(q - 1)] = sL;
            rs[j] = Rs;
            ss[j] = sR;
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
        const int q = threadIdx.x + j * FT;
        if (q < FTS) {
            sm[SM_B + padi(q)] += rs[j];                   // right node of local element q is local node q
            if (GENERAL) sm[SM_S + padi(q)] += ss[j];
        }
    }
    __syncthreads();
}

// Interface (SPIKE) system of G contiguous ranges on one thread: gathered[4 r + {0,1,2,3}] = {x_first, x_last, r_left,
// r_right}; writes bc2 = {U_rank, U_rank+1}.  Used by spike_iface_kernel (hfl_fem.cu) and by the fused exchange kernel
// (hfl_peer.cu).
__device__ __forceinline__ void spike_iface_solve(int G, const double* g, double uL, double uR, int rank, double* bc2) {
    double dl[64], dd[64], du[64], rb[64], U[66];
    U[0] = uL; U[G] = uR;
    const int m = G - 1;
    for (int r = 1; r < G; ++r) {
        const double Ll = g[4 * (r - 1) + 1] - g[4 * (r - 1) + 0];
        const double Lr = g[4 * r + 1] - g[4 * r + 0];
        dl[r - 1] = -1.0 / Ll; du[r - 1] = -1.0 / Lr; dd[r - 1] = 1.0 / Ll + 1.0 / Lr;
        rb[r - 1] = g[4 * (r - 1) + 3] + g[4 * r + 2];
    }
    if (m >= 1) {
        rb[0] -= dl[0] * uL;
        rb[m - 1] -= du[m - 1] * uR;
        for (int i = 1; i < m; ++i) {
            const double w = dl[i] / dd[i - 1];
            dd[i] -= w * du[i - 1];
            rb[i] -= w * rb[i - 1];
        }
        U[m] = rb[m - 1] / dd[m - 1];
        for (int i = m - 2; i >= 0; --i) U[i + 1] = (rb[i] - du[i] * U[i + 2]) / dd[i];
    }
    bc2[0] = U[rank];
    bc2[1] = U[rank + 1];
}

}  // namespace hfl
