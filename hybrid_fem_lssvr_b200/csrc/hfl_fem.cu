// K1: coarse P1 FEM nodal solve, replaces FEMLSSVRPrimalSolver.solve_fem (P:117-145).
//
// The reference assembles the P1 stiffness and a 2-point-Gauss load with scikit-fem, turns the two
// boundary rows into identity rows and calls a sparse direct solver.  Here the same rounded matrix entries
// (k_e = fl(1/h)^2 * (h/2) summed over the two Gauss points, d_i = fl(k_{i-1} + k_i)) and the same two-point
// load are formed on the fly from the node array and the tridiagonal system is solved by a two-level
// partition method:
//
//   level 0  chunks of FS = 8 consecutive nodes (head + 7 interior), one per thread, FT = 256 threads per CTA; the
//            thread loads its nine nodes with vector loads, forms its eight elements itself and keeps the rows in
//            REGISTERS.  fem_chunk_reduce_kernel: two interleaved division-free sweeps eliminate the chunk interior
//            (first / last entries of T^-1 b, T^-1 l e_1, T^-1 r e_s and of the "leak" 1 + v + w) and give the reduced
//            row of every chunk head among the heads.  fem_chunk_backsub_kernel: heads known -> chunk interiors by
//            Thomas in registers -> u with 16-byte stores.  Both are streaming kernels without a serial phase.
//   level 1  the same once more on the heads (every 8th node): chunks of 8 heads per thread, rows from the workspace,
//            2048 heads = 16384 nodes per CTA (a tile).  fem_heads_reduce_kernel: chunk sweeps, then the 255 super heads
//            of the tile are reduced by CYCLIC REDUCTION in shared memory (7 levels, active rows compacted onto the first
//            warps, single-warp levels without CTA barriers: 247 row updates per tile), every super head keeps its final
//            row (L/d, R/d, B/d), and two threads walk the two root-to-leaf paths of the reduction tree to express the
//            tile's first and last interior head through the two tile heads: the 12-number tile record.
//            fem_heads_backsub_kernel: tile heads known -> super heads by the back-substitution tree -> heads by Thomas.
//   top      fem_top_kernel, one CTA: the tile heads (every 16384th node; 611 at 1e7 nodes), chunk per thread + cyclic
//            reduction.
//
// The load uses sin(k pi x_q) = S cos(theta) + C sin(theta), (S, C) = sincospi of the chunk's head node and
// theta = k pi (x_q - x_head) by a short Taylor polynomial when the chunk spans less than 1/16 rad (any mesh of
// more than ~800 k nodes), the library sinpi otherwise: ~12 FP64 instructions per Gauss point instead of ~45.
// Every elimination runs in row-sum form (l, sigma, r), d = sigma - l - r (hfl_fem.cuh): no cancellation,
// ~1e-14 from the exact solution of the rounded system where LU-type solvers lose cond * eps.
// Node traffic: the node array is read twice and u written once (24 B/node) plus ~10 B/node of head rows and head values
// (mostly L2 hits); the element terms are recomputed in the second pass instead of stored (16 B/node each way would
// cost more).
#include <type_traits>
#include "hfl_fem.cuh"

namespace hfl {

// ---------------------------------------------------------------------------------------------------------
// Element terms.

// Stiffness entry and load shares of one element of the reference's Poisson problem.  k carries the reference's
// rounding (P:125-136 through scikit-fem's quadrature loop: fl(fl(1/h)^2 * (h W_q)) twice) because the row-sum
// residues of the assembled diagonal depend on its last bit; the load (2-point Gauss, P:129-136) is formed with
// fused multiply-adds.  TIER 0: library sinpi at the Gauss points; TIER 1, 2: the chunk's Taylor polynomial.
template <int TIER>
__device__ __forceinline__ void element_terms_fast(const FemArgs& a, double x0, double x1, double xref,
                                                   const ForcingPoly<TIER>& fp, double& k, double& Ls, double& Rs) {
    const double h = x1 - x0;
    const double invh = __drcp_rn(h);                // = fl(1 / h), the same bits as __ddiv_rn(1.0, h)
    const double gg = __dmul_rn(invh, invh);
    const double hw = 0.5 * h;                       // |detDF| * W_q
    const double kq = __dmul_rn(gg, hw);
    k = __dadd_rn(kq, kq);
    double f0, f1;
    if (TIER == 0) {
        const double xq0 = __dadd_rn(__dmul_rn(h, a.gx0), x0), xq1 = __dadd_rn(__dmul_rn(h, a.gx1), x0);
        f0 = a.kp2 * sinpi(__dmul_rn(a.k, xq0));     // sin(k pi x) without the argument-reduction slow path
        f1 = a.kp2 * sinpi(__dmul_rn(a.k, xq1));
    } else {
        const double d0 = x0 - xref;
        f0 = fp.eval(fma(h, a.gx0, d0));
        f1 = fp.eval(fma(h, a.gx1, d0));
    }
    Ls = hw * fma(f0, 1.0 - a.gx0, f1 * (1.0 - a.gx1));
    Rs = hw * fma(f0, a.gx0, f1 * a.gx1);
}

// ---------------------------------------------------------------------------------------------------------
// Per-thread chunk: local nodes m0 .. m0 + FS - 1 of the tile, global nodes g0 .. g0 + FS - 1.
template <bool GENERAL>
struct Chunk {
    double k[FS + 1];             // k[j] = stiffness of the element LEFT of node g0 + j (global element g0 - 1 + j)
    double b[FS];                 // load of node g0 + i
    double s[GENERAL ? FS : 1];   // general operator: mass-matrix row sum of node g0 + i
};



// Node loads of the level-0 kernels: streaming (evict-first) so that the 80 MB of nodes do not push the head rows and
// head values (40 + 10 MB, written by one kernel and read by the next two) out of L2.
#ifndef HFL_NODE_LOAD
#define HFL_NODE_LOAD __ldcs
#endif

// First half: loads the chunk's nodes, forms elements j = 1 .. FS (left of nodes g0 + 1 .. g0 + FS), publishes the last
// one for the next thread; thread 0 forms the element left of its head itself.  Call __syncthreads(), then
// chunk_build_finish.  SPECIAL = the tile holds a Dirichlet node or padding past the mesh.
template <bool SPECIAL, bool GENERAL, bool EXCH = true>
__device__ __forceinline__ void chunk_build_start(const FemArgs& a, long long g0, double* __restrict__ ex, Chunk<GENERAL>& c,
                                                  double& k0, double& r0, double& s0) {
    const int t = threadIdx.x;
    double x[FS + 1];
    if (!SPECIAL) {
        const double* p = a.nodes + g0;
        if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
            for (int i = 0; i < FS; i += 2) {
                const double2 v = HFL_NODE_LOAD(reinterpret_cast<const double2*>(p + i));
                x[i] = v.x; x[i + 1] = v.y;
            }
        } else {
#pragma unroll
            for (int i = 0; i < FS; ++i) x[i] = HFL_NODE_LOAD(p + i);
        }
        x[FS] = HFL_NODE_LOAD(p + FS);
    } else {
#pragma unroll
        for (int i = 0; i <= FS; ++i) x[i] = __ldg(a.nodes + min(g0 + i, a.n - 1));
    }
    int tier = 0;
    double S = 0.0, C = 0.0;
    if (!SPECIAL && !GENERAL) {
        const double span = fabs(a.kpi * (x[FS] - x[0]));
        tier = (span <= 5.3e-6) ? 3 : ((span <= 0.0009765625) ? 2 : ((span <= 0.0625) ? 1 : 0));
#ifdef HFL_FEM_NO_TAYLOR
        tier = 0;
#endif
        if (tier != 0) sincospi(__dmul_rn(a.k, x[0]), &S, &C);
    }
    double Ls[FS + 1], Rs[FS + 1], sL[GENERAL ? FS + 1 : 1], sR[GENERAL ? FS + 1 : 1];
    auto elements = [&](auto tier_tag) {
        constexpr int TIER = decltype(tier_tag)::value;
        ForcingPoly<TIER> fp;
        fp.init(a, S, C);
#pragma unroll
        for (int j = 1; j <= FS; ++j) {
            const bool valid = !SPECIAL || (g0 + j <= a.n - 1);
            double k = 0.0, l = 0.0, r = 0.0, sl = 0.0, sr = 0.0;
            if (valid) {
                if (GENERAL) element_terms_general(a, g0 - 1 + j, x[j - 1], x[j], k, sl, sr, l, r);
                else element_terms_fast<TIER>(a, x[j - 1], x[j], x[0], fp, k, l, r);
            }
            c.k[j] = k; Ls[j] = l; Rs[j] = r;
            if (GENERAL) { sL[j] = sl; sR[j] = sr; }
        }
    };
    if (tier == 3) elements(std::integral_constant<int, 3>{});
    else if (tier == 2) elements(std::integral_constant<int, 2>{});
    else if (tier == 1) elements(std::integral_constant<int, 1>{});
    else elements(std::integral_constant<int, 0>{});
#pragma unroll
    for (int i = 0; i < FS; ++i) {
        c.b[i] = Ls[i + 1] + ((i >= 1) ? Rs[i] : 0.0);
        if (GENERAL) c.s[i] = sL[i + 1] + ((i >= 1) ? sR[i] : 0.0);
    }
    k0 = 0.0; r0 = 0.0; s0 = 0.0;
    if (!EXCH) return;        // the caller only needs rows 1 .. FS-1
    ex[0 * FT + t] = c.k[FS];
    ex[1 * FT + t] = Rs[FS];
    if (GENERAL) ex[2 * FT + t] = sR[FS];
    if (t == 0 && g0 >= 1 && g0 <= a.n - 1) {      // element left of the tile head: nobody in this CTA owns it
        const double xm = __ldg(a.nodes + g0 - 1);
        double l, sl = 0.0;
        if (GENERAL) element_terms_general(a, g0 - 1, xm, x[0], k0, sl, s0, l, r0);
        else element_terms_fast<0>(a, xm, x[0], 0.0, ForcingPoly<0>{}, k0, l, r0);
    }
}

template <bool GENERAL>
__device__ __forceinline__ void chunk_build_finish(const double* __restrict__ ex, Chunk<GENERAL>& c, double k0, double r0,
                                                   double s0) {
    const int t = threadIdx.x;
    if (t >= 1) {
        k0 = ex[0 * FT + t - 1];
        r0 = ex[1 * FT + t - 1];
        if (GENERAL) s0 = ex[2 * FT + t - 1];
    }
    c.k[0] = k0;
    c.b[0] += r0;
    if (GENERAL) c.s[0] += s0;
}

// Row i of the chunk in (l, sigma, r, b) form (hfl_fem.cuh explains the row-sum form).
template <bool SPECIAL, bool GENERAL, bool EXACT>
__device__ __forceinline__ void chunk_row(const Chunk<GENERAL>& c, int i, long long g0, const FemArgs& a, double& l,
                                          double& sg, double& r, double& b) {
    if (SPECIAL) {
        const long long g = g0 + i;
        if (g >= a.n) { l = 0.0; sg = 1.0; r = 0.0; b = 0.0; return; }
        if (g == 0) { l = 0.0; sg = 1.0; r = 0.0; b = a.uL; return; }
        if (g == a.n - 1) { l = 0.0; sg = 1.0; r = 0.0; b = a.uR; return; }
    }
    const double kl = c.k[i], kr = c.k[i + 1];
    l = -kl; r = -kr;
    b = c.b[i];
    if (GENERAL) { sg = c.s[i]; return; }              // mass-matrix row sum, assembled without cancellation
    if (EXACT) { sg = 0.0; return; }                   // HFL_COARSE_ASSEMBLED_EXACT: unrounded diagonal kl + kr
    // d = fl(kl + kr) is the reference's assembled diagonal; kl + kr = d + err exactly (TwoSum), so sigma = -err
    const double d = __dadd_rn(kl, kr);
    const double tt = __dsub_rn(d, kl);
    sg = -__dadd_rn(__dsub_rn(kl, __dsub_rn(d, tt)), __dsub_rn(kr, tt));
}

// Diagonal of a row from its row sum and its two off-diagonal entries, d = sigma - (p + q): a sum of same-signed terms,
// and bitwise symmetric in (p, q).  Every elimination below forms its left / right quantities by mirror-image
// operations with explicitly rounded products (no fused contraction chosen by the compiler): on a (nearly) uniform mesh
// the roundings of the two directions then agree instead of differing by a systematic fraction of an ulp - the
// reduction tree doubles such a left / right bias at every level (2^17 from the chunk to the top of a 1e7-node mesh),
// which showed up as 1e-11 instead of 1e-14.
__device__ __forceinline__ double diag_of(double sg, double p, double q) { return __dsub_rn(sg, __dadd_rn(p, q)); }

// Interior of a chunk of FS rows (rows 1 .. FS-1; row(i, l, sg, r, b) with compile-time i after unrolling): first / last
// entries of the three partial solutions x = y - u_head v - u_next w and the "leaks" e = 1 + v + w, all formed without
// cancellation; the forward and the backward sweep are independent chains and run interleaved.
// out = {y1, v1, w1, e1, ys, vs, ws, es}.
template <class RowFn>
__device__ __forceinline__ void chunk_sweeps(RowFn&& row, double (&out)[8]) {
    double l, sg, r, b;
    // forward: transformed row i has entries (head: v, i: d, i+1: r); tp = v + d + r is its row sum
    row(1, l, sg, r, b);
    double tpf = sg, vpf = l, rpf = r, bpf = b, dpf = diag_of(tpf, vpf, rpf);
    // backward: entries (i-1: l, i: d, next head: w)
    row(FS - 1, l, sg, r, b);
    double tpb = sg, bpb = b, wpb = r, lpb = l, dpb = diag_of(tpb, wpb, lpb);
    // division-free elimination: row_i <- d_prev row_i - l_i row_prev (every product keeps its sign, so the row sums stay
    // sums of same-signed terms); the rows grow by ~d per step, 7 steps stay far inside the double range, and the only
    // reciprocals are the two at the end.  The two sweeps are exact mirror images of each other, operation by operation
    // and rounding by rounding (no compiler contraction): see diag_of.
#pragma unroll
    for (int s = 2; s < FS; ++s) {
        row(s, l, sg, r, b);
        tpf = __fma_rn(dpf, sg, __dmul_rn(-l, tpf));
        vpf = __dmul_rn(-l, vpf);
        bpf = fma(dpf, b, -l * bpf);
        rpf = __dmul_rn(dpf, r);
        dpf = diag_of(tpf, vpf, rpf);
        row(FS - s, l, sg, r, b);
        tpb = __fma_rn(dpb, sg, __dmul_rn(-r, tpb));
        wpb = __dmul_rn(-r, wpb);
        bpb = fma(dpb, b, -r * bpb);
        lpb = __dmul_rn(dpb, l);
        dpb = diag_of(tpb, wpb, lpb);
    }
    const double invf = fast_rcp(dpf), invb = fast_rcp(dpb);
    out[4] = bpf * invf; out[5] = __dmul_rn(vpf, invf); out[6] = __dmul_rn(rpf, invf); out[7] = __dmul_rn(tpf, invf);
    out[0] = bpb * invb; out[1] = __dmul_rn(lpb, invb); out[2] = __dmul_rn(wpb, invb); out[3] = __dmul_rn(tpb, invb);
}

// Thomas on a chunk interior (rows 1 .. FS-1) between the two known heads ua (row 0) and ub (the next chunk's head),
// division-free forward sweep in row-sum form: transformed row i is dN x_i + rN x_{i+1} = bN with sP = dN + rN formed as
// a sum of same-signed terms (row_i <- d_prev row_i - l_i row_prev); the reciprocals of the 7 pivots are independent of
// each other.  IDENT: identity rows (Dirichlet nodes, padding) reproduce their value bit for bit.
template <bool IDENT, class RowFn>
__device__ __forceinline__ void chunk_thomas(RowFn&& row, double ua, double ub, double (&xs)[FS]) {
    double cq[FS], bq[FS];
    double lo, so, ro, bo, dP = 1.0, sP = 1.0, bP = ua;     // "previous row" = the known head: x = ua
#pragma unroll
    for (int i = 1; i < FS; ++i) {
        row(i, lo, so, ro, bo);
        if (i == FS - 1) bo = fma(-ro, ub, bo);
        double sN = fma(dP, so, -lo * sP);
        double rN = dP * ro;
        double bN = fma(dP, bo, -lo * bP);
        if (IDENT && lo == 0.0 && ro == 0.0) { sN = so; rN = 0.0; bN = bo; }
        const double dN = sN - rN;
        const double inv = fast_rcp(dN);
        cq[i] = rN * inv; bq[i] = bN * inv;
        dP = dN; sP = sN; bP = bN;
    }
    xs[FS - 1] = bq[FS - 1];
#pragma unroll
    for (int i = FS - 2; i >= 1; --i) xs[i] = fma(-cq[i], xs[i + 1], bq[i]);
    xs[0] = ua;
}

// Chunk interior of the top level: rows through a getter with run-time indices (the rows live in shared or global
// memory), chunk length S up to TOP_MAX_CHUNK.  Same partial solutions as chunk_reduce_reg, with a reciprocal per row
// (a division-free sweep would overflow on long chunks); the two sweeps run interleaved.
template <class Rows>
__device__ __forceinline__ void chunk_reduce(const Rows& rows, int m0, int S, double (&out)[8]) {
    if (S < 2) {   // no interior: neighbouring heads couple directly (x_first = u_next, x_last = u_head)
        out[0] = 0.0; out[1] = 0.0; out[2] = -1.0; out[3] = 0.0;
        out[4] = 0.0; out[5] = -1.0; out[6] = 0.0; out[7] = 0.0;
        return;
    }
    double l, sg, r, b;
    rows.get(m0 + 1, l, sg, r, b);
    double tpf = sg, vpf = l, rpf = r, bpf = b, dpf = diag_of(tpf, vpf, rpf);
    rows.get(m0 + S - 1, l, sg, r, b);
    double tpb = sg, bpb = b, wpb = r, lpb = l, dpb = diag_of(tpb, wpb, lpb);
    for (int i = 2; i < S; ++i) {          // the two sweeps mirror each other rounding by rounding (see diag_of)
        rows.get(m0 + i, l, sg, r, b);
        const double mf = __dmul_rn(l, fast_rcp(dpf));
        tpf = __fma_rn(-mf, tpf, sg);
        vpf = __dmul_rn(-mf, vpf);
        bpf = fma(-mf, bpf, b);
        rpf = r;
        dpf = diag_of(tpf, vpf, rpf);
        rows.get(m0 + S - i, l, sg, r, b);
        const double mb = __dmul_rn(r, fast_rcp(dpb));
        tpb = __fma_rn(-mb, tpb, sg);
        wpb = __dmul_rn(-mb, wpb);
        bpb = fma(-mb, bpb, b);
        lpb = l;
        dpb = diag_of(tpb, wpb, lpb);
    }
    const double invf = fast_rcp(dpf), invb = fast_rcp(dpb);
    out[4] = bpf * invf; out[5] = __dmul_rn(vpf, invf); out[6] = __dmul_rn(rpf, invf); out[7] = __dmul_rn(tpf, invf);
    out[0] = bpb * invb; out[1] = __dmul_rn(lpb, invb); out[2] = __dmul_rn(wpb, invb); out[3] = __dmul_rn(tpb, invb);
}

// Reduced equation of a chunk head from its own row (lp, sp, rp, bp), the previous chunk's {ys, vs, ws, es} and
// its own {y1, v1, w1, e1}: couplings L, R to the neighbouring heads, row sum S, right-hand side B.
__device__ __forceinline__ void head_equation(double lp, double sp, double rp, double bp, double ys_prev,
                                              double vs_prev, double es_prev, const double (&e8)[8], double& L,
                                              double& S, double& R, double& B) {
    L = __dmul_rn(-lp, vs_prev);
    R = __dmul_rn(-rp, e8[2]);
    S = __dadd_rn(sp, __dadd_rn(__dmul_rn(-lp, es_prev), __dmul_rn(-rp, e8[3])));      // mirror-symmetric (see diag_of)
    B = fma(-rp, e8[0], fma(-lp, ys_prev, bp));
}

// Shared-memory index of chunk head i (0 .. FT; index FT = the next tile's head): one pad per 16 doubles keeps the
// strided accesses of the cyclic reduction (stride 2, 4, ..., 128 heads) free of bank conflicts.
__host__ __device__ constexpr int cp(int i) { return i + (i >> 4); }

// Level 1: FT1 threads per CTA, one chunk of FS heads each: a tile is FTS1 = FT1 * FS heads = 8192 nodes.  Smaller tiles
// than the level-0 CTA (FT = 256): the level-1 kernels are chains of dependent steps (sweeps, reduction tree, paths), so
// what counts is how many of them are resident at once and how deep the tree is (6 levels over 127 super heads).
#ifndef HFL_FEM_FT1
#define HFL_FEM_FT1 128
#endif
constexpr int FT1 = HFL_FEM_FT1;
constexpr int FTS1 = FT1 * FS;
constexpr int CRLEN1 = cp(FT1) + 1;
static_assert(FT1 >= 32 && (FT1 & (FT1 - 1)) == 0 && FT1 % (FT / FS) == 0, "level-1 CTA size");

// Forward cyclic reduction over rows 1 .. N-1 (rows as (l, sigma, r, b) at padded index cp(i); index 0 and index N are
// unknown columns without rows of their own).  Level delta updates the rows at multiples of 2 delta from their
// neighbours at distance delta; the active rows are compacted onto the first threads, and once a level fits one warp
// the remaining levels run on warp 0 alone behind __syncwarp (no CTA barrier, nobody else executes the loop).
// Every thread of the CTA must call this; ends with a CTA barrier.  Afterwards the row of index i is final for its level
// (it couples to i -+ lowbit(i)).
template <int N>
__device__ __forceinline__ void cr_forward(double* sL, double* sS, double* sR, double* sB, int t) {
    auto update = [&](int lv) {
        const int delta = 1 << lv, i = (t + 1) << (lv + 1);
        const int im = cp(i - delta), ip = cp(i + delta), ii = cp(i);
        const double lm = sL[im], sm_ = sS[im], rm = sR[im], bm = sB[im];
        const double lq = sL[ip], sq = sS[ip], rq = sR[ip], bq = sB[ip];
        const double al = __dmul_rn(-sL[ii], fast_rcp(diag_of(sm_, lm, rm)));      // >= 0
        const double be = __dmul_rn(-sR[ii], fast_rcp(diag_of(sq, lq, rq)));
        sL[ii] = __dmul_rn(al, lm);
        sR[ii] = __dmul_rn(be, rq);
        sS[ii] = __dadd_rn(sS[ii], __dadd_rn(__dmul_rn(al, sm_), __dmul_rn(be, sq)));  // mirror-symmetric (see diag_of)
        sB[ii] = fma(be, bq, fma(al, bm, sB[ii]));
    };
    int lv = 0;
#pragma unroll 1
    for (; (N >> (lv + 1)) - 1 > 32; ++lv) {
        if (t < (N >> (lv + 1)) - 1) update(lv);
        __syncthreads();
    }
    if (t < 32) {
#pragma unroll 1
        for (; (2 << lv) < N; ++lv) {
            if (t < (N >> (lv + 1)) - 1) update(lv);
            __syncwarp();
        }
    }
    __syncthreads();
}


// ---------------------------------------------------------------------------------------------------------
// Three levels.  Level 0: chunks of FS = 8 nodes, one per thread (FT = 256 threads = 2048 nodes per CTA), rows formed
// from the node array and kept in registers.  Level 1: the chunk heads (every 8th node) form a tridiagonal system of
// their own; chunks of 8 heads, one per thread, FT threads = 2048 heads = 16384 nodes per CTA (a "tile"), rows read
// from the workspace; the 256 "super heads" of a tile are reduced by cyclic reduction in shared memory.  Top: the
// tile heads (every 16384th node), one CTA.  The two level-0 kernels are streaming kernels without a serial phase; the
// reduction trees, which are latency-bound, run in the two light level-1 kernels (1/8 of the unknowns, 4 doubles each).
//
// Workspace per right-hand side (doubles): rec [12][ntile] | utop [ntile] | top rows [6 ntile] | sheads [ntile][3][FT] |
// hrow [4][NH] | edge [8 nchunkcta] | uh [NH + 8], NH = FT * nchunkcta heads.

// Level 0, pass 1: head equations.  hrow[f * NH + h], f = 0..3: reduced row {L, sigma, R, B} of head h among the heads
// (h - 1, h, h + 1).  The head of a CTA's first chunk needs the previous CTA's last chunk, so its row is finished by the
// level-1 kernel from edge[cta][0..4] = {l, sigma, -r e1, -r w1, b - r y1} of that head and edge[cta - 1][5..7] =
// {ys, vs, es} of the chunk before it.
template <bool SPECIAL, bool GENERAL, bool EXACT>
__device__ __forceinline__ void fem_chunk_reduce_body(const FemArgs& a, double* __restrict__ hrow, double* __restrict__ edge,
                                                      double* s_ex) {
    const int t = threadIdx.x;
    const long long g0 = (long long)blockIdx.x * FTS + (long long)t * FS;
    const long long NH = (long long)gridDim.x * FT, h = (long long)blockIdx.x * FT + t;
    Chunk<GENERAL> c;
    double k0, r0, s0;
    chunk_build_start<SPECIAL, GENERAL>(a, g0, s_ex, c, k0, r0, s0);
    double e8[8];
    chunk_sweeps([&](int i, double& l, double& sg, double& r, double& b) { chunk_row<SPECIAL, GENERAL, EXACT>(c, i, g0, a, l, sg, r, b); },
                 e8);      // rows 1 .. FS-1 do not involve the element left of the head
    s_ex[3 * FT + t] = e8[4]; s_ex[4 * FT + t] = e8[5]; s_ex[5 * FT + t] = e8[7];     // ys, vs, es of this chunk
    __syncthreads();
    chunk_build_finish<GENERAL>(s_ex, c, k0, r0, s0);
    double lp, sp, rp, bp;
    chunk_row<SPECIAL, GENERAL, EXACT>(c, 0, g0, a, lp, sp, rp, bp);
    if (t >= 1) {
        double L, S, R, B;
        head_equation(lp, sp, rp, bp, s_ex[3 * FT + t - 1], s_ex[4 * FT + t - 1], s_ex[5 * FT + t - 1], e8, L, S, R, B);
        hrow[h] = L; hrow[NH + h] = S; hrow[2 * NH + h] = R; hrow[3 * NH + h] = B;
    } else {
        double* eo = edge + (size_t)blockIdx.x * 8;
        eo[0] = lp; eo[1] = sp; eo[2] = __dmul_rn(-rp, e8[3]); eo[3] = __dmul_rn(-rp, e8[2]); eo[4] = fma(-rp, e8[0], bp);
        hrow[h] = 0.0; hrow[NH + h] = 1.0; hrow[2 * NH + h] = 0.0; hrow[3 * NH + h] = 0.0;      // finished at level 1
    }
    if (t == FT - 1) {
        double* eo = edge + (size_t)blockIdx.x * 8;
        eo[5] = e8[4]; eo[6] = e8[5]; eo[7] = e8[7];
    }
}

#ifndef HFL_FEM_MINB
#define HFL_FEM_MINB 3
#endif
template <bool GENERAL, bool EXACT = false>
__global__ void __launch_bounds__(FT, HFL_FEM_MINB) fem_chunk_reduce_kernel(const FemArgs a_in, double* __restrict__ hrow,
                                                                            double* __restrict__ edge) {
    __shared__ double s_ex[6 * FT];
    const FemArgs a = select_rhs(a_in);
    hrow += (size_t)blockIdx.y * a.ws_stride; edge += (size_t)blockIdx.y * a.ws_stride;
    const long long P = (long long)blockIdx.x * FTS;
    if (P == 0 || P + FTS >= a.n - 1) fem_chunk_reduce_body<true, GENERAL, EXACT>(a, hrow, edge, s_ex);
    else fem_chunk_reduce_body<false, GENERAL, EXACT>(a, hrow, edge, s_ex);
}

// The 8 head rows of a level-1 chunk (heads H0 .. H0 + 7), heads past NH padded with identity rows.
struct HeadRows {
    double l[FS], s[FS], r[FS], b[FS];
    __device__ __forceinline__ void load(const double* hrow, long long NH, long long H0) {
        if (H0 + FS <= NH) {
#pragma unroll
            for (int i = 0; i < FS; i += 2) {
                const double2 vl = __ldcg(reinterpret_cast<const double2*>(hrow + H0 + i));
                const double2 vs = __ldcg(reinterpret_cast<const double2*>(hrow + NH + H0 + i));
                const double2 vr = __ldcg(reinterpret_cast<const double2*>(hrow + 2 * NH + H0 + i));
                const double2 vb = __ldcg(reinterpret_cast<const double2*>(hrow + 3 * NH + H0 + i));
                l[i] = vl.x; l[i + 1] = vl.y; s[i] = vs.x; s[i + 1] = vs.y;
                r[i] = vr.x; r[i + 1] = vr.y; b[i] = vb.x; b[i + 1] = vb.y;
            }
        } else {
#pragma unroll
            for (int i = 0; i < FS; ++i) { l[i] = 0.0; s[i] = 1.0; r[i] = 0.0; b[i] = 0.0; }
        }
    }
    __device__ __forceinline__ void get(int i, double& lo, double& so, double& ro, double& bo) const {
        lo = l[i]; so = s[i]; ro = r[i]; bo = b[i];
    }
};

// Level 1, pass 1.  Tile record (SoA, rec[f * ntile + tile]): f = 0..3 {l, sigma, r, b} of the tile head's row,
// 4..7 {y, v, w, e} of the tile's first interior head, 8..11 of its last one (x = y - v u_P - w u_Q, e = 1 + v + w).
// sheads[tile][3][FT1]: final cyclic-reduction row {L/d, R/d, B/d} of every super head.
__device__ __forceinline__ void heads_reduce_tile(long long tile, long long ntile, const double* hrow, const double* edge,
                                                  long long NH, double* __restrict__ rec, double* __restrict__ sheads,
                                                  double* s_ex2, double* s_cr) {
    const int t = threadIdx.x;
    const long long H0 = tile * FTS1 + (long long)t * FS;
    HeadRows hr;
    hr.load(hrow, NH, H0);
    if ((t & (FT / FS - 1)) == 0 && H0 < NH) {            // first head of a level-0 CTA: finish its row from the edge records
        const long long cta = H0 / FT;
        const double* eo = edge + (size_t)cta * 8;
        const double lp = __ldcg(eo + 0), sp = __ldcg(eo + 1);
        double ysp = 0.0, vsp = 0.0, esp = 0.0;
        if (cta > 0) { ysp = __ldcg(eo + 5 - 8); vsp = __ldcg(eo + 6 - 8); esp = __ldcg(eo + 7 - 8); }
        hr.l[0] = __dmul_rn(-lp, vsp);
        hr.r[0] = __ldcg(eo + 3);
        hr.s[0] = __dadd_rn(sp, __dadd_rn(__dmul_rn(-lp, esp), __ldcg(eo + 2)));          // the sum head_equation forms
        hr.b[0] = fma(-lp, ysp, __ldcg(eo + 4));
    }
    double e8[8];
    chunk_sweeps([&](int i, double& l, double& sg, double& r, double& b) { hr.get(i, l, sg, r, b); }, e8);
    const double lp = hr.l[0], sp = hr.s[0], rp = hr.r[0], bp = hr.b[0];
    s_ex2[0 * FT1 + t] = e8[4]; s_ex2[1 * FT1 + t] = e8[5]; s_ex2[2 * FT1 + t] = e8[7];     // ys, vs, es of this chunk
    __syncthreads();
    double* sL = s_cr; double* sS = s_cr + CRLEN1; double* sR = s_cr + 2 * CRLEN1; double* sB = s_cr + 3 * CRLEN1;
    if (t >= 1) {
        double L, S, R, B;
        head_equation(lp, sp, rp, bp, s_ex2[0 * FT1 + t - 1], s_ex2[1 * FT1 + t - 1], s_ex2[2 * FT1 + t - 1], e8, L, S, R, B);
        sL[cp(t)] = L; sS[cp(t)] = S; sR[cp(t)] = R; sB[cp(t)] = B;
    }
    __syncthreads();
    // cyclic reduction over super heads 1 .. FT1-1; heads 0 (this tile's) and FT1 (the next tile's) stay as unknown columns
    cr_forward<FT1>(sL, sS, sR, sB, t);
    if (t >= 1) {        // every row is final: scale by its diagonal, keep it for the back-substitution pass
        const int ii = cp(t);
        const double L = sL[ii], S = sS[ii], R = sR[ii], B = sB[ii];
        const double inv = fast_rcp(diag_of(S, L, R));
        const double Ld = __dmul_rn(L, inv), Rd = __dmul_rn(R, inv), Bd = B * inv;
        sL[ii] = Ld; sR[ii] = Rd; sB[ii] = Bd; sS[ii] = __dmul_rn(S, inv);
        double* o = sheads + (size_t)tile * 3 * FT1;
        o[t] = Ld; o[FT1 + t] = Rd; o[2 * FT1 + t] = Bd;
    } else {             // slot 0 is the tile head itself (solved at the top level)
        double* o = sheads + (size_t)tile * 3 * FT1;
        o[0] = 0.0; o[FT1] = 0.0; o[2 * FT1] = 0.0;
    }
    __syncthreads();
    const long long nt = ntile;
    double* out = rec + tile;
    if (t == 0 || t == FT1 - 1) {
        // super head FT1/2 through the two tile heads, then down the tree to super head 1 (t = 0) or FT1-1 (t = FT1-1):
        // u = Y - V u_P - W u_Q, E = 1 + V + W
        int i = cp(FT1 / 2);
        double Y = sB[i], V = sL[i], W = sR[i], E = sS[i];
        if (t == 0) {
#pragma unroll
            for (int delta = FT1 / 4; delta >= 1; delta >>= 1) {        // node delta: left neighbour 0, right neighbour 2 delta
                i = cp(delta);
                const double Rd = sR[i];
                Y = fma(-Rd, Y, sB[i]); V = __fma_rn(-Rd, V, sL[i]); W = __dmul_rn(-Rd, W); E = __fma_rn(-Rd, E, sS[i]);
            }
            out[0 * nt] = lp; out[1 * nt] = sp; out[2 * nt] = rp; out[3 * nt] = bp;
            out[4 * nt] = fma(-e8[2], Y, e8[0]);
            out[5 * nt] = __fma_rn(-e8[2], V, e8[1]);
            out[6 * nt] = __dmul_rn(-e8[2], W);
            out[7 * nt] = __fma_rn(-e8[2], E, e8[3]);
        } else {
#pragma unroll
            for (int delta = FT1 / 4; delta >= 1; delta >>= 1) {        // node FT1 - delta: left neighbour FT1 - 2 delta, right FT1
                i = cp(FT1 - delta);
                const double Ld = sL[i];
                Y = fma(-Ld, Y, sB[i]); V = __dmul_rn(-Ld, V); W = __fma_rn(-Ld, W, sR[i]); E = __fma_rn(-Ld, E, sS[i]);
            }
            out[8 * nt] = fma(-e8[5], Y, e8[4]);
            out[9 * nt] = __dmul_rn(-e8[5], V);
            out[10 * nt] = __fma_rn(-e8[5], W, e8[6]);
            out[11 * nt] = __fma_rn(-e8[5], E, e8[7]);
        }
    }
}

// Level 1, pass 1: one CTA per tile.
__global__ void __launch_bounds__(FT1) fem_heads_reduce_kernel(const double* hrow, const double* edge, long long NH,
                                                              double* __restrict__ rec, double* __restrict__ sheads,
                                                              long long ws_stride) {
    __shared__ double s_ex2[3 * FT1], s_cr[4 * CRLEN1];
    hrow += (size_t)blockIdx.y * ws_stride; edge += (size_t)blockIdx.y * ws_stride;
    rec += (size_t)blockIdx.y * ws_stride; sheads += (size_t)blockIdx.y * ws_stride;
    heads_reduce_tile(blockIdx.x, gridDim.x, hrow, edge, NH, rec, sheads, s_ex2, s_cr);
}

// Thomas elimination of a chunk interior between two known head values, in (l, sigma, r) form: s = row sum over
// the not-yet-eliminated columns, d = s - r.  One forward step; q carries s / d of the previous row.
__device__ __forceinline__ void thomas_step(double l, double sg, double r, double b, double& q, double& cprev,
                                            double& bprev) {
    const double s = fma(-l, q, sg);      // sg - l on the first row (q = 1), sg - l * (s'/d')_{i-1} afterwards
    const double inv = fast_rcp(s - r);
    bprev = fma(-l, bprev, b) * inv;
    cprev = r * inv;
    q = s * inv;
}

// Top level: solve the system of tile heads.  One CTA of TOPT threads, chunk of S heads per thread: chunk interiors
// by two interleaved sweeps, the TOPT chunk heads by cyclic reduction (forward with compacted levels, then the
// back-substitution tree), chunk interiors by Thomas.  Head 0 is global node 0, always a Dirichlet (identity) row, so it
// enters the reduction as a known column.  SMEM_ROWS: the rows (l, sigma, r, b) [4][cnt] live in shared memory behind
// the reduction arrays; otherwise in the workspace (wsrows).  The tile records are SoA: the row build reads them
// coalesced.
constexpr int CRT = cp(TOPT) + 1;
constexpr int TOP_SMEM_DOUBLES = 4 * CRT + 3 * TOPT;     // reduction arrays + head exchange

template <bool SMEM_ROWS>
__global__ void __launch_bounds__(TOPT) fem_top_kernel(const double* __restrict__ rec, int cnt, int S,
                                                       double* __restrict__ wsrows, double* __restrict__ utop,
                                                       long long ws_stride) {
    extern __shared__ double sm[];
    const int t = threadIdx.x;
    rec += (size_t)blockIdx.y * ws_stride; wsrows += (size_t)blockIdx.y * ws_stride; utop += (size_t)blockIdx.y * ws_stride;
    double* sL = sm; double* sS = sm + CRT; double* sR = sm + 2 * CRT; double* sB = sm + 3 * CRT;
    double* ex = sm + 4 * CRT;
    double* rowbase = SMEM_ROWS ? sm + TOP_SMEM_DOUBLES : wsrows;
    double* rl = rowbase; double* rs = rowbase + cnt; double* rr = rowbase + 2 * (size_t)cnt; double* rb = rowbase + 3 * (size_t)cnt;
    const size_t nt = (size_t)cnt;
    for (int c = t; c < cnt; c += TOPT) {
        const double* rc = rec + c;
        const double hl = rc[0], hr = rc[2 * nt];
        double l = 0.0, sg = rc[1 * nt], b = rc[3 * nt];
        double leak_l = -hl;      // no tile to the left: the coupling stays in the row sum (hl is 0 for the Dirichlet head anyway)
        if (c > 0) {
            l = __dmul_rn(-hl, rc[9 * nt - 1]);
            leak_l = __dmul_rn(-hl, rc[11 * nt - 1]);
            b = fma(-hl, rc[8 * nt - 1], b);
        }
        sg = __dadd_rn(sg, __dadd_rn(leak_l, __dmul_rn(-hr, rc[7 * nt])));      // mirror-symmetric (see diag_of)
        const double r = __dmul_rn(-hr, rc[6 * nt]);
        b = fma(-hr, rc[4 * nt], b);
        rl[c] = l; rs[c] = sg; rr[c] = r; rb[c] = b;
    }
    __syncthreads();
    ArrayRows rows{rl, rs, rr, rb, cnt};
    double e8[8];
    chunk_reduce(rows, t * S, S, e8);
    double lp, sp, rp, bp;
    rows.get(t * S, lp, sp, rp, bp);
    ex[0 * TOPT + t] = e8[4]; ex[1 * TOPT + t] = e8[5]; ex[2 * TOPT + t] = e8[7];
    __syncthreads();
    {
        double L, Sg, R, B;
        if (t >= 1) {
            head_equation(lp, sp, rp, bp, ex[0 * TOPT + t - 1], ex[1 * TOPT + t - 1], ex[2 * TOPT + t - 1], e8, L, Sg, R, B);
            sL[cp(t)] = L; sS[cp(t)] = Sg; sR[cp(t)] = R; sB[cp(t)] = B;
        } else {
            // head 0 = global node 0 (identity row, no coupling to its right): its value is its right-hand side
            head_equation(0.0, sp - lp, rp, bp, 0.0, 0.0, 0.0, e8, L, Sg, R, B);
            sB[cp(0)] = B * fast_rcp(diag_of(Sg, L, R));
        }
    }
    __syncthreads();
    cr_forward<TOPT>(sL, sS, sR, sB, t);
    if (t >= 1) {
        const int ii = cp(t);
        const double L = sL[ii], R = sR[ii];
        const double inv = fast_rcp(diag_of(sS[ii], L, R));
        sL[ii] = __dmul_rn(L, inv); sR[ii] = __dmul_rn(R, inv); sB[ii] = sB[ii] * inv;
    }
    __syncthreads();
    // back-substitution tree; the head values overwrite the (now unused) row sums
    double* su = sS;
    if (t == 0) { su[cp(0)] = sB[cp(0)]; su[cp(TOPT)] = 0.0; }
    auto level = [&](int q) {
        const int delta = (TOPT / 2) >> q;
        if (t < (1 << q)) {
            const int i = delta * (2 * t + 1), ii = cp(i);
            su[ii] = fma(-sR[ii], su[cp(i + delta)], fma(-sL[ii], su[cp(i - delta)], sB[ii]));
        }
    };
    int q = 0;
    if (t < 32) {
        __syncwarp();
#pragma unroll 1
        for (; q <= 5; ++q) { level(q); __syncwarp(); }
    }
    q = 6;
    __syncthreads();
#pragma unroll 1
    for (; (1 << q) < TOPT; ++q) { level(q); __syncthreads(); }
    const int m0 = t * S;
    const double ua = su[cp(t)];
    if (m0 < cnt) utop[m0] = ua;
    if (S >= 2 && m0 + 1 < cnt) {
        const double ub = su[cp(t + 1)];
        // Thomas on rows m0+1 .. m0+S-1 with known neighbours; c', b' stay thread-private
        double tc[TOP_MAX_CHUNK], tb[TOP_MAX_CHUNK];
        double lo, so, ro, bo, qq = 1.0, cpv = 0.0, bpv = ua;     // "previous row" = the known head: x = ua
        for (int i = 1; i < S; ++i) {
            rows.get(m0 + i, lo, so, ro, bo);
            if (i == S - 1) bo = fma(-ro, ub, bo);
            thomas_step(lo, so, ro, bo, qq, cpv, bpv);
            tc[i] = cpv; tb[i] = bpv;
        }
        double xv = bpv;
        if (m0 + S - 1 < cnt) utop[m0 + S - 1] = xv;
        for (int i = S - 2; i >= 1; --i) {
            if (m0 + i < cnt) {
                xv = tb[i] - tc[i] * xv;
                utop[m0 + i] = xv;
            } else {
                xv = 0.0;   // padded identity rows
            }
        }
    }
}

// Level 1, pass 2: tile head values known -> super heads by the back-substitution of the cyclic reduction,
// u_i = B_i - L_i u_{i-d} - R_i u_{i+d} with d = lowbit(i): levels d >= 4 (the 63 super heads at multiples of 4) on warp
// 0, then every thread evaluates the two rows (d = 2, d = 1) its own u_t, u_{t+1} need -> the 7 interior heads of every
// level-1 chunk by Thomas in registers -> uh[h] for every head h.
__global__ void __launch_bounds__(FT1) fem_heads_backsub_kernel(const double* __restrict__ hrow, long long NH,
                                                               const double* __restrict__ utop, int ntile,
                                                               const double* __restrict__ sheads, double* __restrict__ uh,
                                                               long long ws_stride) {
    __shared__ double s_u[CRLEN1];
    hrow += (size_t)blockIdx.y * ws_stride; utop += (size_t)blockIdx.y * ws_stride;
    sheads += (size_t)blockIdx.y * ws_stride; uh += (size_t)blockIdx.y * ws_stride;
    const int t = threadIdx.x;
    const long long H0 = (long long)blockIdx.x * FTS1 + (long long)t * FS;
    const double* o = sheads + (size_t)blockIdx.x * 3 * FT1;
    const int od = t | 1;                                  // the odd one of {t, t + 1}
    const int m2 = (od & 2) ? (od - 1) : (od + 1);         // its neighbour that is 2 mod 4 (the other one is 0 mod 4)
    const double oL = o[od], oR = o[FT1 + od], oB = o[2 * FT1 + od];
    const double mL = o[m2], mR = o[FT1 + m2], mB = o[2 * FT1 + m2];
    HeadRows hr;
    hr.load(hrow, NH, H0);                                 // rows 1 .. 7 are never the first head of a level-0 CTA
    if (t < 32) {
        constexpr int NLV = (FT1 == 256 ? 6 : (FT1 == 128 ? 5 : (FT1 == 64 ? 4 : 3)));   // d = FT1/2 .. 4
        static_assert((4 << NLV) == FT1 && FT1 / 8 <= 32, "level count of the warp-0 tree");
        double rl[NLV], rr[NLV], rb[NLV];
#pragma unroll
        for (int q = 0; q < NLV; ++q) {
            const int delta = (FT1 / 2) >> q, cnt = 1 << q;
            const int i = (t < cnt) ? delta * (2 * t + 1) : delta;
            rl[q] = o[i]; rr[q] = o[FT1 + i]; rb[q] = o[2 * FT1 + i];
        }
        if (t == 0) {
            s_u[cp(0)] = utop[blockIdx.x];
            s_u[cp(FT1)] = ((int)blockIdx.x + 1 < ntile) ? utop[blockIdx.x + 1] : 0.0;
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < NLV; ++q) {
            const int delta = (FT1 / 2) >> q, cnt = 1 << q;
            if (t < cnt) {
                const int i = delta * (2 * t + 1);
                s_u[cp(i)] = fma(-rr[q], s_u[cp(i + delta)], fma(-rl[q], s_u[cp(i - delta)], rb[q]));
            }
            __syncwarp();
        }
    }
    __syncthreads();
    double ua, ub;
    {
        const double um2 = fma(-mR, s_u[cp(m2 + 2)], fma(-mL, s_u[cp(m2 - 2)], mB));      // level d = 2
        const int z4 = 2 * od - m2;                                                         // the neighbour that is 0 mod 4
        const double uz4 = s_u[cp(z4)];
        const double ulo = (m2 < od) ? um2 : uz4, uhi = (m2 < od) ? uz4 : um2;
        const double uod = fma(-oR, uhi, fma(-oL, ulo, oB));                                 // level d = 1
        ua = (t & 1) ? uod : ulo;
        ub = (t & 1) ? uhi : uod;
    }
    double xs[FS];
    chunk_thomas<true>([&](int i, double& l, double& sg, double& r, double& b) { hr.get(i, l, sg, r, b); }, ua, ub, xs);
    if (H0 + FS <= NH) {
#pragma unroll
        for (int i = 0; i < FS; i += 2) *reinterpret_cast<double2*>(uh + H0 + i) = make_double2(xs[i], xs[i + 1]);
    }
}

// Level 0, pass 2: head values known -> chunk interiors by Thomas in registers -> u with 16-byte stores.  No shared
// memory, no barrier.  iface4 (optional): the end-node residuals of the multi-GPU interface system (hfl.h), written by the
// two threads that own nodes 0 and n - 2.
template <bool SPECIAL, bool GENERAL, bool EXACT>
__device__ __forceinline__ void fem_chunk_backsub_body(const FemArgs& a, const double* __restrict__ uh, double* __restrict__ u,
                                                       double* __restrict__ iface4) {
    const int t = threadIdx.x;
    const long long g0 = (long long)blockIdx.x * FTS + (long long)t * FS;
    const long long NH = (long long)gridDim.x * FT, h = (long long)blockIdx.x * FT + t;
    const double ua = uh[h], ub = (h + 1 < NH) ? uh[h + 1] : 0.0;
    Chunk<GENERAL> c;
    double k0, r0, s0;
    chunk_build_start<SPECIAL, GENERAL, false>(a, g0, nullptr, c, k0, r0, s0);      // rows 1 .. FS-1 only: no exchange
    double xs[FS];
    chunk_thomas<SPECIAL>([&](int i, double& l, double& sg, double& r, double& b) { chunk_row<SPECIAL, GENERAL, EXACT>(c, i, g0, a, l, sg, r, b); },
                          ua, ub, xs);
    double* p = u + g0;
    if (!SPECIAL && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
        for (int i = 0; i < FS; i += 2) *reinterpret_cast<double2*>(p + i) = make_double2(xs[i], xs[i + 1]);
    } else {
#pragma unroll
        for (int i = 0; i < FS; ++i)
            if (g0 + i < a.n) p[i] = xs[i];
    }
    if (SPECIAL && !GENERAL && iface4 != nullptr) {
        if (g0 == 0) {                                   // node 0: r_left = load_0 + k_0 (u_1 - u_0)
            const double x0 = __ldg(a.nodes), x1 = __ldg(a.nodes + 1);
            double k, Ls, Rs;
            element_terms_fast<0>(a, x0, x1, 0.0, ForcingPoly<0>{}, k, Ls, Rs);
            iface4[0] = x0;
            iface4[2] = Ls + k * (xs[1] - xs[0]);
        }
        const long long gm = a.n - 2;                    // node n - 2 and the element right of it
        if (gm >= g0 && gm < g0 + FS) {
            const int i = (int)(gm - g0);
            double um = xs[0], up = ub;
#pragma unroll
            for (int q = 0; q < FS; ++q) {
                if (q == i) um = xs[q];
                if (q == i + 1) up = xs[q];
            }
            const double x0 = __ldg(a.nodes + gm), x1 = __ldg(a.nodes + gm + 1);
            double k, Ls, Rs;
            element_terms_fast<0>(a, x0, x1, 0.0, ForcingPoly<0>{}, k, Ls, Rs);
            iface4[1] = x1;
            iface4[3] = Rs + k * (um - up);
        }
    }
}

template <bool GENERAL, bool EXACT = false>
__global__ void __launch_bounds__(FT, HFL_FEM_MINB) fem_chunk_backsub_kernel(const FemArgs a_in, const double* __restrict__ uh,
                                                                             double* __restrict__ u, double* __restrict__ iface4) {
    const FemArgs a = select_rhs(a_in);
    uh += (size_t)blockIdx.y * a.ws_stride; u += (size_t)blockIdx.y * a.n;
    const long long P = (long long)blockIdx.x * FTS;
    if (P == 0 || P + FTS >= a.n - 1) fem_chunk_backsub_body<true, GENERAL, EXACT>(a, uh, u, iface4);
    else fem_chunk_backsub_body<false, GENERAL, EXACT>(a, uh, u, iface4);
}

// End-node residuals for the multi-GPU interface system (see hfl.h).  flux2 = {q_0, B_{n-2}} from the flux scan
// (NULL for the assembled solvers): the end fluxes q_0 and q_{n-2} = q_0 - B_{n-2} are then used as they are instead of
// k (u_1 - u_0), which amplifies the rounding of u by k ~ 1/h (3e-9 at h = 1e-7).
__global__ void fem_reaction_kernel(const FemArgs a, const double* __restrict__ u, const double* __restrict__ flux2,
                                    double* __restrict__ out4) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double k, Ls, Rs;
        out4[0] = a.nodes[0];
        out4[1] = a.nodes[a.n - 1];
        element_terms(a, a.nodes[0], a.nodes[1], k, Ls, Rs);
        out4[2] = Ls + (flux2 ? flux2[0] : k * (u[1] - u[0]));
        element_terms(a, a.nodes[a.n - 2], a.nodes[a.n - 1], k, Ls, Rs);
        out4[3] = Rs + (flux2 ? flux2[1] - flux2[0] : k * (u[a.n - 2] - u[a.n - 1]));
    }
}

__global__ void fem_apply_bc_kernel(long long n, const double* __restrict__ nodes, double* __restrict__ u,
                                    double bl, double br) {
    const double x0 = nodes[0], x1 = nodes[n - 1];
    const double invL = 1.0 / (x1 - x0);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double x = nodes[i];
        u[i] += (bl * (x1 - x) + br * (x - x0)) * invL;
    }
}

// Device version of the interface solve (G <= 64): keeps the multi-GPU step stream-ordered (body in hfl_fem.cuh).
__global__ void spike_iface_kernel(int G, const double* __restrict__ g, double uL, double uR, int rank,
                                   double* __restrict__ bc2) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    spike_iface_solve(G, g, uL, uR, rank, bc2);
}

}  // namespace hfl

using namespace hfl;

static inline long long fem_nchunkcta(long long n) { return (n + FTS - 1) / FTS; }                  // level-0 CTAs (2048 nodes each)
static inline long long fem_ntile(long long n) { return (fem_nchunkcta(n) * FT + FTS1 - 1) / FTS1; }  // level-1 CTAs (FTS1 heads each)
// rec | utop | top rows | sheads | hrow | edge | uh (see the layout comment above fem_chunk_reduce_body); every block
// starts on a 16-byte boundary
static size_t fem_ws_doubles(long long n) {
    const size_t nc = (size_t)fem_nchunkcta(n), nt = (size_t)fem_ntile(n), NH = nc * FT;
    size_t top = (REC + 1 + 6 + 3 * FT1) * nt;
    top += top & 1;
    return top + 4 * NH + 8 * nc + NH + 8;
}

extern "C" size_t hfl_fem_p1_workspace_bytes(int64_t n_nodes) {
    if (n_nodes < 2) return 256;
    return fem_ws_doubles(n_nodes) * sizeof(double) + 256;
}

// per right-hand side: the single-solve layout rounded up to 256 bytes
static size_t fem_multi_stride_doubles(int64_t n_nodes) { return ((hfl_fem_p1_workspace_bytes(n_nodes) + 255) / 256) * 32; }

extern "C" size_t hfl_fem_p1_multi_workspace_bytes(int64_t n_nodes, int R) {
    if (R < 1) R = 1;
    return fem_multi_stride_doubles(n_nodes) * sizeof(double) * (size_t)R;
}

int hfl_fem_flux_scan(const FemArgs& a, int R, double* d_u, void* d_ws, size_t ws_bytes, cudaStream_t s);   // hfl_flux.cu

// R right-hand sides (grid.y): a.kfreqs / a.ws_stride set by the caller for R > 1, NULL / 0 for a single solve.
static int fem_solve_impl(FemArgs a, int R, int coarse_solver, double* d_u, double* d_iface4, void* d_ws, size_t ws_bytes,
                          cudaStream_t s) {
    const long long n = a.n;
    a.gx0 = 0.5 * (-0.5773502691896257) + 0.5;   // 0.5 * leggauss(2) + 0.5
    a.gx1 = 0.5 * (0.5773502691896257) + 0.5;
    a.exact_rowsum = (coarse_solver == HFL_COARSE_ASSEMBLED_EXACT) ? 1 : 0;
    fem_taylor_table(a);
    if (coarse_solver == HFL_COARSE_FLUX_SCAN) {
        int rc = hfl_fem_flux_scan(a, R, d_u, d_ws, ws_bytes, s);
        if (rc != HFL_OK) return rc;
    } else {
        const long long nc = fem_nchunkcta(n), nt = fem_ntile(n), NH = nc * FT;
        if (nt > (long long)TOPT * TOP_MAX_CHUNK) {
            set_error("hfl_fem_p1_solve: %lld nodes exceed the single-call limit of %lld; split the mesh across GPUs",
                      (long long)n, (long long)TOPT * TOP_MAX_CHUNK * FTS1 * FS);
            return HFL_ERR_UNSUPPORTED;
        }
        double* rec = reinterpret_cast<double*>(d_ws);
        double* utop = rec + (size_t)REC * nt;
        double* wsrows = utop + nt;
        double* sheads = wsrows + 6 * (size_t)nt;
        size_t top = (size_t)(REC + 1 + 6 + 3 * FT1) * nt;
        top += top & 1;
        double* hrow = rec + top;
        double* edge = hrow + 4 * (size_t)NH;
        double* uh = edge + 8 * (size_t)nc;
        const int S = (int)((nt + TOPT - 1) / TOPT);
        const dim3 grid0((unsigned)nc, (unsigned)R), grid1((unsigned)nt, (unsigned)R);
        // top level: rows in shared memory while they fit (fem_top_smem_kb option), else in the workspace
        const size_t top_small = (size_t)TOP_SMEM_DOUBLES * sizeof(double), top_rows = top_small + 4 * (size_t)nt * sizeof(double);
        const bool top_in_smem = top_rows <= (size_t)get_option_top_smem_kb() * 1024;
        const size_t top_smem = top_in_smem ? top_rows : top_small;
        if (top_in_smem)
            HFL_CUDA_CHECK(cudaFuncSetAttribute(fem_top_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)top_smem));
        else
            HFL_CUDA_CHECK(cudaFuncSetAttribute(fem_top_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)top_smem));
        if (a.aq != nullptr) fem_chunk_reduce_kernel<true><<<grid0, FT, 0, s>>>(a, hrow, edge);
        else if (a.exact_rowsum) fem_chunk_reduce_kernel<false, true><<<grid0, FT, 0, s>>>(a, hrow, edge);
        else fem_chunk_reduce_kernel<false><<<grid0, FT, 0, s>>>(a, hrow, edge);
        fem_heads_reduce_kernel<<<grid1, FT1, 0, s>>>(hrow, edge, NH, rec, sheads, a.ws_stride);
        if (top_in_smem) fem_top_kernel<true><<<dim3(1, R), TOPT, top_smem, s>>>(rec, (int)nt, S, wsrows, utop, a.ws_stride);
        else fem_top_kernel<false><<<dim3(1, R), TOPT, top_smem, s>>>(rec, (int)nt, S, wsrows, utop, a.ws_stride);
        fem_heads_backsub_kernel<<<grid1, FT1, 0, s>>>(hrow, NH, utop, (int)nt, sheads, uh, a.ws_stride);
        if (a.aq != nullptr) fem_chunk_backsub_kernel<true><<<grid0, FT, 0, s>>>(a, uh, d_u, nullptr);
        else if (a.exact_rowsum) fem_chunk_backsub_kernel<false, true><<<grid0, FT, 0, s>>>(a, uh, d_u, d_iface4);
        else fem_chunk_backsub_kernel<false><<<grid0, FT, 0, s>>>(a, uh, d_u, d_iface4);
        count_launch(5);
        HFL_CUDA_CHECK(cudaGetLastError());
    }
    if (d_iface4 != nullptr && coarse_solver == HFL_COARSE_FLUX_SCAN) {   // the assembled solvers write it in pass 2
        const double* flux2 = reinterpret_cast<const double*>(d_ws) + 6 * (size_t)((n - 1 + FTS - 1) / FTS);   // after the tile prefixes (hfl_flux.cu)
        fem_reaction_kernel<<<1, 32, 0, s>>>(a, d_u, flux2, d_iface4);
        count_launch();
        HFL_CUDA_CHECK(cudaGetLastError());
    }
    return HFL_OK;
}

extern "C" int hfl_fem_p1_solve(int64_t n, const double* d_nodes, double k_freq, double u_left, double u_right,
                                int coarse_solver, double* d_u, double* d_iface4, void* d_ws, size_t ws_bytes,
                                void* stream) {
    HFL_REQUIRE(n >= 2, "hfl_fem_p1_solve: need at least 2 nodes (got %lld)", (long long)n);
    HFL_REQUIRE(d_nodes != nullptr && d_u != nullptr, "hfl_fem_p1_solve: d_nodes / d_u is NULL");
    HFL_REQUIRE(coarse_solver == HFL_COARSE_ASSEMBLED_PCR || coarse_solver == HFL_COARSE_FLUX_SCAN ||
                    coarse_solver == HFL_COARSE_ASSEMBLED_EXACT,
                "hfl_fem_p1_solve: unknown coarse_solver %d", coarse_solver);
    HFL_REQUIRE(d_ws != nullptr && ws_bytes >= hfl_fem_p1_workspace_bytes(n),
                "hfl_fem_p1_solve: workspace too small (%zu < %zu)", ws_bytes, hfl_fem_p1_workspace_bytes(n));
    HFL_REQUIRE((reinterpret_cast<uintptr_t>(d_ws) & 255) == 0, "hfl_fem_p1_solve: workspace must be 256-byte aligned");
    const double pi = 3.14159265358979323846;
    FemArgs a;
    a.n = n; a.nodes = d_nodes; a.k = k_freq; a.kpi = k_freq * pi; a.kp2 = a.kpi * a.kpi; a.uL = u_left; a.uR = u_right;
    a.aq = nullptr; a.cq = nullptr; a.fq = nullptr; a.kfreqs = nullptr; a.ws_stride = 0;
    return fem_solve_impl(a, 1, coarse_solver, d_u, d_iface4, d_ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int hfl_fem_p1_solve_multi(int64_t n, const double* d_nodes, int R, const double* d_k_freq, double u_left,
                                      double u_right, int coarse_solver, double* d_u, void* d_ws, size_t ws_bytes,
                                      void* stream) {
    HFL_REQUIRE(n >= 2, "hfl_fem_p1_solve_multi: need at least 2 nodes (got %lld)", (long long)n);
    HFL_REQUIRE(R >= 1 && R <= 65535, "hfl_fem_p1_solve_multi: R = %d outside 1..65535", R);
    HFL_REQUIRE(d_nodes != nullptr && d_u != nullptr && d_k_freq != nullptr, "hfl_fem_p1_solve_multi: d_nodes / d_u / d_k_freq is NULL");
    HFL_REQUIRE(coarse_solver == HFL_COARSE_ASSEMBLED_PCR || coarse_solver == HFL_COARSE_FLUX_SCAN ||
                    coarse_solver == HFL_COARSE_ASSEMBLED_EXACT,
                "hfl_fem_p1_solve_multi: unknown coarse_solver %d", coarse_solver);
    HFL_REQUIRE(d_ws != nullptr && ws_bytes >= hfl_fem_p1_multi_workspace_bytes(n, R),
                "hfl_fem_p1_solve_multi: workspace too small (%zu < %zu)", ws_bytes, hfl_fem_p1_multi_workspace_bytes(n, R));
    HFL_REQUIRE((reinterpret_cast<uintptr_t>(d_ws) & 255) == 0, "hfl_fem_p1_solve_multi: workspace must be 256-byte aligned");
    FemArgs a;
    a.n = n; a.nodes = d_nodes; a.k = 0.0; a.kpi = 0.0; a.kp2 = 0.0; a.uL = u_left; a.uR = u_right;
    a.aq = nullptr; a.cq = nullptr; a.fq = nullptr;
    a.kfreqs = d_k_freq; a.ws_stride = (long long)fem_multi_stride_doubles(n);
    return fem_solve_impl(a, R, coarse_solver, d_u, nullptr, d_ws, a.ws_stride * sizeof(double), (cudaStream_t)stream);
}

extern "C" int hfl_fem_p1_solve_general(int64_t n, const double* d_nodes, const double* d_aq, const double* d_cq,
                                        const double* d_fq, double u_left, double u_right, double* d_u, void* d_ws,
                                        size_t ws_bytes, void* stream) {
    HFL_REQUIRE(n >= 2, "hfl_fem_p1_solve_general: need at least 2 nodes (got %lld)", (long long)n);
    HFL_REQUIRE(d_nodes && d_u && d_aq && d_fq, "hfl_fem_p1_solve_general: d_nodes / d_u / d_aq / d_fq is NULL");
    HFL_REQUIRE(d_ws != nullptr && ws_bytes >= hfl_fem_p1_workspace_bytes(n),
                "hfl_fem_p1_solve_general: workspace too small (%zu < %zu)", ws_bytes, hfl_fem_p1_workspace_bytes(n));
    HFL_REQUIRE((reinterpret_cast<uintptr_t>(d_ws) & 255) == 0, "hfl_fem_p1_solve_general: workspace must be 256-byte aligned");
    FemArgs a;
    a.n = n; a.nodes = d_nodes; a.k = 0.0; a.kpi = 0.0; a.kp2 = 0.0; a.uL = u_left; a.uR = u_right;
    a.aq = d_aq; a.cq = d_cq; a.fq = d_fq; a.kfreqs = nullptr; a.ws_stride = 0;
    return fem_solve_impl(a, 1, HFL_COARSE_ASSEMBLED_PCR, d_u, nullptr, d_ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int hfl_fem_apply_bc(int64_t n, const double* d_nodes, double* d_u, double bl, double br, void* stream) {
    HFL_REQUIRE(n >= 2 && d_nodes != nullptr && d_u != nullptr, "hfl_fem_apply_bc: bad arguments");
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    fem_apply_bc_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(n, d_nodes, d_u, bl, br);
    count_launch();
    HFL_CUDA_CHECK(cudaGetLastError());
    return HFL_OK;
}

// Interface system of G contiguous ranges: coarse P1 stiffness on the interface mesh,
//   (1/L_{r-1} + 1/L_r) U_r - U_{r-1}/L_{r-1} - U_{r+1}/L_r = r_right(r-1) + r_left(r),  r = 1..G-1
// with U_0 = u_left, U_G = u_right.  Solved by the Thomas algorithm (G <= a few dozen).
extern "C" int hfl_spike_interface_solve(int G, const double* g, double u_left, double u_right, double* iface) {
    HFL_REQUIRE(G >= 1 && g != nullptr && iface != nullptr, "hfl_spike_interface_solve: bad arguments");
    iface[0] = u_left;
    iface[G] = u_right;
    if (G == 1) return HFL_OK;
    const int m = G - 1;
    std::vector<double> dl(m), dd(m), du(m), rb(m);
    for (int r = 1; r < G; ++r) {
        const double Ll = g[4 * (r - 1) + 1] - g[4 * (r - 1) + 0];
        const double Lr = g[4 * r + 1] - g[4 * r + 0];
        HFL_REQUIRE(Ll > 0.0 && Lr > 0.0, "hfl_spike_interface_solve: rank %d has a non-positive length", r);
        dl[r - 1] = -1.0 / Ll; du[r - 1] = -1.0 / Lr; dd[r - 1] = 1.0 / Ll + 1.0 / Lr;
        rb[r - 1] = g[4 * (r - 1) + 3] + g[4 * r + 2];
    }
    rb[0] -= dl[0] * u_left;
    rb[m - 1] -= du[m - 1] * u_right;
    for (int i = 1; i < m; ++i) {
        const double w = dl[i] / dd[i - 1];
        dd[i] -= w * du[i - 1];
        rb[i] -= w * rb[i - 1];
    }
    iface[m] = rb[m - 1] / dd[m - 1];
    for (int i = m - 2; i >= 0; --i) iface[i + 1] = (rb[i] - du[i] * iface[i + 2]) / dd[i];
    return HFL_OK;
}

extern "C" int hfl_spike_interface_solve_device(int G, const double* d_gathered, double u_left, double u_right,
                                                int rank, double* d_bc2, void* stream) {
    HFL_REQUIRE(G >= 1 && G <= 64, "hfl_spike_interface_solve_device: G=%d outside [1, 64]", G);
    HFL_REQUIRE(rank >= 0 && rank < G, "hfl_spike_interface_solve_device: rank outside [0, G)");
    HFL_REQUIRE(d_gathered != nullptr && d_bc2 != nullptr, "hfl_spike_interface_solve_device: NULL pointer");
    spike_iface_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(G, d_gathered, u_left, u_right, rank, d_bc2);
    count_launch();
    HFL_CUDA_CHECK(cudaGetLastError());
    return HFL_OK;
}
