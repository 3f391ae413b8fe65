// K1: coarse P1 FEM nodal solve, replaces FEMLSSVRPrimalSolver.solve_fem (P:117-145).
//
// The reference assembles the P1 stiffness and a 2-point-Gauss load with scikit-fem, turns the two
// boundary rows into identity rows and calls a sparse direct solver.  Here the same rounded entries
// (k_e = fl(1/h)^2 * (h/2) summed over the two Gauss points, d_i = fl(k_{i-1} + k_i), load by the same
// two-point rule) are formed on the fly from the node array and the tridiagonal system is solved by a
// two-level partition method whose reduced systems are solved by parallel cyclic reduction (PCR):
//
//   level 0  tiles of T*S = 2048 nodes, one CTA each.  Every thread eliminates the S-1 interior nodes
//            of its chunk (two sweeps give the first/last entries of T^-1 b, T^-1 l e_1, T^-1 r e_s and of
//            the "leak" 1 + v + w), the T-1 chunk heads form a tridiagonal system solved by PCR in shared
//            memory (4 right-hand sides in the reduce pass; the back-substitution pass reuses the stored
//            head solutions).
//   top      one CTA solves the system of tile heads the same way (chunk per thread + PCR).
//
// Every elimination runs in row-sum form (l, sigma, r), d = sigma - l - r (hfl_fem.cuh): no cancellation,
// ~1e-14 from the exact solution of the rounded system where LU-type solvers lose cond * eps.
// Kernels: fem_reduce_kernel -> fem_top_kernel -> fem_backsub_kernel.  Node traffic: the node array
// is read twice and u written once (24 B/node) plus 3 B/node of head solutions; the load is recomputed
// in the second pass instead of stored (re-reading it measured slower than two sinpi per element).
#include "hfl_fem.cuh"

namespace hfl {

// Interior of a chunk (rows m0+1 .. m0+S-1): first / last entries of the three partial solutions
// x = y - u_head v - u_next w and the "leaks" e = 1 + v + w, all formed without cancellation.
// out = {y1, v1, w1, e1, ys, vs, ws, es}.
template <class Rows>
__device__ __forceinline__ void chunk_reduce(const Rows& rows, int m0, int S, double (&out)[8]) {
    if (S < 2) {   // no interior: neighbouring heads couple directly (x_first = u_next, x_last = u_head)
        out[0] = 0.0; out[1] = 0.0; out[2] = -1.0; out[3] = 0.0;
        out[4] = 0.0; out[5] = -1.0; out[6] = 0.0; out[7] = 0.0;
        return;
    }
    double l, sg, r, b;
    // forward: transformed row i has entries (head: v, i: d, i+1: r); tp = v + d + r is its row sum
    rows.get(m0 + 1, l, sg, r, b);
    double tp = sg, vp = l, rp = r, bp = b;
    double dp = tp - vp - rp;
    for (int i = 2; i < S; ++i) {
        rows.get(m0 + i, l, sg, r, b);
        const double m = l * fast_rcp(dp);        // <= 0
        tp = fma(-m, tp, sg);
        vp = -m * vp;
        bp = fma(-m, bp, b);
        rp = r;
        dp = tp - vp - rp;
    }
    double inv = fast_rcp(dp);
    out[4] = bp * inv; out[5] = vp * inv; out[6] = rp * inv; out[7] = tp * inv;
    // backward: entries (i-1: l, i: d, next head: w)
    rows.get(m0 + S - 1, l, sg, r, b);
    tp = sg; bp = b;
    double wp = r, lp = l;
    dp = tp - lp - wp;
    for (int i = S - 2; i >= 1; --i) {
        rows.get(m0 + i, l, sg, r, b);
        const double m = r * fast_rcp(dp);
        tp = fma(-m, tp, sg);
        wp = -m * wp;
        bp = fma(-m, bp, b);
        lp = l;
        dp = tp - lp - wp;
    }
    inv = fast_rcp(dp);
    out[0] = bp * inv; out[1] = lp * inv; out[2] = wp * inv; out[3] = tp * inv;
}

// Parallel cyclic reduction over equations first..last (one per thread, index = thread id), NR right-hand
// sides, rows as (l, sigma, r).  sm holds 2 * (3 + NR) * T doubles.  Every thread of the CTA must call this.
template <int NR, int T>
__device__ __forceinline__ void pcr_solve(double* sm, int t, int first, int last, double l, double sg, double r,
                                          double (&rhs)[NR], double (&x)[NR]) {
    constexpr int W = 3 + NR;
    int cur = 0;
    const bool active = (t >= first && t <= last);
    {
        double* bufw = sm + cur * W * T;
        bufw[0 * T + t] = l; bufw[1 * T + t] = sg; bufw[2 * T + t] = r;
#pragma unroll
        for (int q = 0; q < NR; ++q) bufw[(3 + q) * T + t] = rhs[q];
    }
    __syncthreads();
    const int count = last - first + 1;
    for (int delta = 1; delta < count; delta <<= 1) {
        const double* bufr = sm + cur * W * T;
        double* bufw = sm + (cur ^ 1) * W * T;
        if (active) {
            const int im = t - delta, ip = t + delta;
            double sn = sg, ln = 0.0, rn = 0.0;
            if (im >= first) {
                const double lm = bufr[0 * T + im], sm_ = bufr[1 * T + im], rm = bufr[2 * T + im];
                const double al = -l * fast_rcp(sm_ - lm - rm);      // >= 0
                sn = fma(al, sm_, sn);
                ln = al * lm;
#pragma unroll
                for (int q = 0; q < NR; ++q) rhs[q] = fma(al, bufr[(3 + q) * T + im], rhs[q]);
            } else {
                sn -= l;          // no such neighbour: l is 0 here by construction
            }
            if (ip <= last) {
                const double lq = bufr[0 * T + ip], sq = bufr[1 * T + ip], rq = bufr[2 * T + ip];
                const double be = -r * fast_rcp(sq - lq - rq);
                sn = fma(be, sq, sn);
                rn = be * rq;
#pragma unroll
                for (int q = 0; q < NR; ++q) rhs[q] = fma(be, bufr[(3 + q) * T + ip], rhs[q]);
            } else {
                sn -= r;
            }
            l = ln; sg = sn; r = rn;
        }
        bufw[0 * T + t] = l; bufw[1 * T + t] = sg; bufw[2 * T + t] = r;
#pragma unroll
        for (int q = 0; q < NR; ++q) bufw[(3 + q) * T + t] = rhs[q];
        __syncthreads();
        cur ^= 1;
    }
    const double inv = fast_rcp(sg - l - r);
#pragma unroll
    for (int q = 0; q < NR; ++q) x[q] = rhs[q] * inv;
}

// Reduced equation of a chunk head from its own row (lp, sp, rp, bp), the previous chunk's {ys, vs, ws, es} and
// its own {y1, v1, w1, e1}: couplings L, R to the neighbouring heads, row sum S, right-hand side B.
__device__ __forceinline__ void head_equation(double lp, double sp, double rp, double bp, double ys_prev,
                                              double vs_prev, double es_prev, const double (&e8)[8], double& L,
                                              double& S, double& R, double& B) {
    L = -lp * vs_prev;
    R = -rp * e8[2];
    S = fma(-rp, e8[3], fma(-lp, es_prev, sp));
    B = fma(-rp, e8[0], fma(-lp, ys_prev, bp));
}

// Level 0, pass 1: one record per tile = {l, sigma, r, b of the tile head, y1, v1, w1, e1, ys, vs, ws, es of the interior}.
template <bool SPECIAL, bool GENERAL, bool EXACT>
__device__ __forceinline__ void fem_reduce_body(const FemArgs& a, double* __restrict__ rec, double* __restrict__ yvw,
                                                double* sm) {
    const int t = threadIdx.x;
    const long long P = (long long)blockIdx.x * FTS;
    MeshRows<SPECIAL, GENERAL, EXACT> rows{sm + SM_K, sm + SM_B, sm + SM_S, P, a.n, a.uL, a.uR};
    double e8[8];
    chunk_reduce(rows, t * FS, FS, e8);
    double lp, sp, rp, bp;
    rows.get(t * FS, lp, sp, rp, bp);
    __syncthreads();                 // the element arrays are dead from here: exchange / PCR buffers alias them
    double* ex = sm + SM_EX;
    ex[0 * FT + t] = e8[4]; ex[1 * FT + t] = e8[5]; ex[2 * FT + t] = e8[7];     // ys, vs, es of this chunk
    __syncthreads();
    double l = 0.0, sg = 1.0, r = 0.0, rhs[4] = {0.0, 0.0, 0.0, 0.0}, x[4];
    if (t >= 1) {
        double B;
        head_equation(lp, sp, rp, bp, ex[0 * FT + t - 1], ex[1 * FT + t - 1], ex[2 * FT + t - 1], e8, l, sg, r, B);
        rhs[0] = B;
        rhs[3] = sg;                                       // T (1 + V + W) = full row sums  ->  E = 1 + V + W
        if (t == 1) { rhs[1] = l; sg -= l; l = 0.0; }      // coupling to the tile head moves to the right-hand side
        if (t == FT - 1) { rhs[2] = r; sg -= r; r = 0.0; } // ... and the one to the next tile's head
    }
    __syncthreads();                 // ex is consumed; the PCR buffers alias it
    pcr_solve<4, FT>(sm + SM_PCR, t, 1, FT - 1, l, sg, r, rhs, x);
    {   // chunk-head partial solutions, read back by the back-substitution pass
        double* o = yvw + (size_t)blockIdx.x * 3 * FT;
        o[t] = x[0]; o[FT + t] = x[1]; o[2 * FT + t] = x[2];
    }
    // x = {Y, V, W, E} of head t.  The tile's first interior node belongs to chunk 0 (it needs head 1's
    // solution), its last interior node to chunk T-1.
    __syncthreads();
    double* xb = sm;                 // 4 doubles: head 1's solution for thread 0
    if (t == 1) { xb[0] = x[0]; xb[1] = x[1]; xb[2] = x[2]; xb[3] = x[3]; }
    __syncthreads();
    double* out = rec + (long long)blockIdx.x * REC;
    if (t == 0) {
        const double Y1 = xb[0], V1 = xb[1], W1 = xb[2], E1 = xb[3];
        out[0] = lp; out[1] = sp; out[2] = rp; out[3] = bp;
        out[4] = fma(-e8[2], Y1, e8[0]);
        out[5] = fma(-e8[2], V1, e8[1]);
        out[6] = -e8[2] * W1;
        out[7] = fma(-e8[2], E1, e8[3]);
    }
    if (t == FT - 1) {
        out[8] = fma(-e8[5], x[0], e8[4]);
        out[9] = -e8[5] * x[1];
        out[10] = fma(-e8[5], x[2], e8[6]);
        out[11] = fma(-e8[5], x[3], e8[7]);
    }
}

template <bool GENERAL, bool EXACT = false>
__global__ void __launch_bounds__(FT, GENERAL ? 3 : 4) fem_reduce_kernel(const FemArgs a_in, double* __restrict__ rec,
                                                                         double* __restrict__ yvw) {
    extern __shared__ double sm[];
    const FemArgs a = select_rhs(a_in);
    rec += (size_t)blockIdx.y * a.ws_stride; yvw += (size_t)blockIdx.y * a.ws_stride;
    const long long P = (long long)blockIdx.x * FTS;
    load_tile_elements<GENERAL>(a, P, sm);
    if (P == 0 || P + FTS >= a.n - 1) fem_reduce_body<true, GENERAL, EXACT>(a, rec, yvw, sm);
    else fem_reduce_body<false, GENERAL, EXACT>(a, rec, yvw, sm);
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

// Top level: solve the system of tile heads.  One CTA of TOPT threads, chunk of S heads per thread.
// SMEM_ROWS: the rows (l, sigma, r, b) [4][cnt] live in shared memory behind the 8 * TOPT doubles of PCR buffers
// (small meshes, see the launch); otherwise in the workspace (wsrows).  The chunk sweeps walk their rows with
// dependent loads, so the shared-memory variant removes ~4 S L2 round trips from this serial kernel.
template <bool SMEM_ROWS>
__global__ void __launch_bounds__(TOPT) fem_top_kernel(const double* __restrict__ rec, int cnt, int S,
                                                       double* __restrict__ wsrows, double* __restrict__ utop,
                                                       long long ws_stride) {
    extern __shared__ double sm[];
    const int t = threadIdx.x;
    rec += (size_t)blockIdx.y * ws_stride; wsrows += (size_t)blockIdx.y * ws_stride; utop += (size_t)blockIdx.y * ws_stride;
    double* rowbase = SMEM_ROWS ? sm + 8 * TOPT : wsrows;
    double* rl = rowbase; double* rs = rowbase + cnt; double* rr = rowbase + 2 * (size_t)cnt; double* rb = rowbase + 3 * (size_t)cnt;
    for (int c = t; c < cnt; c += TOPT) {
        const double* rc = rec + (size_t)c * REC;
        double l = 0.0, sg = rc[1], b = rc[3];
        if (c > 0) {
            const double* rp = rec + (size_t)(c - 1) * REC;
            l = -rc[0] * rp[9];
            sg = fma(-rc[0], rp[11], sg);
            b = fma(-rc[0], rp[8], b);
        } else {
            sg -= rc[0];          // no tile to the left (rc[0] is 0 for the Dirichlet head anyway)
        }
        sg = fma(-rc[2], rc[7], sg);
        const double r = -rc[2] * rc[6];
        b = fma(-rc[2], rc[4], b);
        rl[c] = l; rs[c] = sg; rr[c] = r; rb[c] = b;
    }
    __syncthreads();
    ArrayRows rows{rl, rs, rr, rb, cnt};
    double e8[8];
    chunk_reduce(rows, t * S, S, e8);
    double lp, sp, rp, bp;
    rows.get(t * S, lp, sp, rp, bp);
    double* pcr = sm;                // 2 * 4 * TOPT
    double* ex = sm + 4 * TOPT;      // 3 * TOPT, inside the second PCR buffer: consumed before PCR first writes there
    ex[0 * TOPT + t] = e8[4]; ex[1 * TOPT + t] = e8[5]; ex[2 * TOPT + t] = e8[7];
    __syncthreads();
    double l = 0.0, sg = 1.0, r = 0.0, rhs[1] = {0.0}, x[1];
    if (t >= 1) {
        head_equation(lp, sp, rp, bp, ex[0 * TOPT + t - 1], ex[1 * TOPT + t - 1], ex[2 * TOPT + t - 1], e8, l, sg, r, rhs[0]);
    } else {
        // head 0 is global node 0 (identity row); its equation still carries the coupling to chunk 0's interior
        head_equation(0.0, sp - lp, rp, bp, 0.0, 0.0, 0.0, e8, l, sg, r, rhs[0]);
    }
    pcr_solve<1, TOPT>(pcr, t, 0, TOPT - 1, l, sg, r, rhs, x);
    double* uh = sm;   // head values; the PCR buffers are dead
    __syncthreads();
    uh[t] = x[0];
    __syncthreads();
    const int m0 = t * S;
    if (m0 < cnt) utop[m0] = x[0];
    if (S >= 2 && m0 + 1 < cnt) {
        const double ua = x[0];
        const double ub = (t + 1 < TOPT) ? uh[t + 1] : 0.0;
        // Thomas on rows m0+1 .. m0+S-1 with known neighbours; c', b' stay thread-private
        double tc[TOP_MAX_CHUNK], tb[TOP_MAX_CHUNK];
        double lo, so, ro, bo, q = 1.0, cp = 0.0, bpv = ua;     // "previous row" = the known head: x = ua
        for (int i = 1; i < S; ++i) {
            rows.get(m0 + i, lo, so, ro, bo);
            if (i == S - 1) bo = fma(-ro, ub, bo);
            thomas_step(lo, so, ro, bo, q, cp, bpv);
            tc[i] = cp; tb[i] = bpv;
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

// Level 0, pass 2: tile head values known -> chunk heads from the stored partial solutions
// (U_t = Y_t - u_P V_t - u_Q W_t) -> chunk interiors by Thomas -> u.
template <bool SPECIAL, bool GENERAL, bool EXACT>
__device__ __forceinline__ void fem_backsub_body(const FemArgs& a, const double* __restrict__ utop, int ntile,
                                                 const double* __restrict__ yvw, double* __restrict__ u, double* sm) {
    const int t = threadIdx.x;
    const long long P = (long long)blockIdx.x * FTS;
    const double uP = utop[blockIdx.x];
    const double uQ = ((int)blockIdx.x + 1 < ntile) ? utop[blockIdx.x + 1] : 0.0;
    double* uh = sm + sm_uh(GENERAL);    // FT head values (+1 for the next tile's head)
    {
        const double* o = yvw + (size_t)blockIdx.x * 3 * FT;
        uh[t] = (t == 0) ? uP : (o[t] - uP * o[FT + t] - uQ * o[2 * FT + t]);
        if (t == 0) uh[FT] = uQ;
    }
    __syncthreads();
    MeshRows<SPECIAL, GENERAL, EXACT> rows{sm + SM_K, sm + SM_B, sm + SM_S, P, a.n, a.uL, a.uR};
    const double ua = uh[t], ub = uh[t + 1];
    // Thomas on the chunk interior, compile-time length FS - 1 (row-sum form, see thomas_step)
    double cpv[FS], bpv[FS], xs[FS];
    {
        double lo, so, ro, bo, q = 1.0, cp = 0.0, bv = ua;      // "previous row" = the known head: x = ua
#pragma unroll
        for (int i = 1; i < FS; ++i) {
            rows.get(t * FS + i, lo, so, ro, bo);
            if (i == FS - 1) bo = fma(-ro, ub, bo);
            thomas_step(lo, so, ro, bo, q, cp, bv);
            cpv[i] = cp; bpv[i] = bv;
        }
        xs[FS - 1] = bpv[FS - 1];
#pragma unroll
        for (int i = FS - 2; i >= 1; --i) xs[i] = bpv[i] - cpv[i] * xs[i + 1];
        xs[0] = ua;
    }
    __syncthreads();               // everyone is done reading the element arrays
    double* stage = sm + SM_K;     // reuse as padded output staging
#pragma unroll
    for (int i = 0; i < FS; ++i) stage[padi(t * FS + i)] = xs[i];
    __syncthreads();
    for (int m = t; m < FTS; m += FT) {
        const long long g = P + m;
        if (g < a.n) u[g] = stage[padi(m)];
    }
}

template <bool GENERAL, bool EXACT = false>
__global__ void __launch_bounds__(FT, GENERAL ? 3 : 4) fem_backsub_kernel(const FemArgs a_in, const double* __restrict__ utop,
                                                                          int ntile, const double* __restrict__ yvw,
                                                                          double* __restrict__ u) {
    extern __shared__ double sm[];
    const FemArgs a = select_rhs(a_in);
    utop += (size_t)blockIdx.y * a.ws_stride; yvw += (size_t)blockIdx.y * a.ws_stride; u += (size_t)blockIdx.y * a.n;
    const long long P = (long long)blockIdx.x * FTS;
    load_tile_elements<GENERAL>(a, P, sm);   // recomputed: re-reading cached terms (16 B/node) measured slower than 2 sinpi
    if (P == 0 || P + FTS >= a.n - 1) fem_backsub_body<true, GENERAL, EXACT>(a, utop, ntile, yvw, u, sm);
    else fem_backsub_body<false, GENERAL, EXACT>(a, utop, ntile, yvw, u, sm);
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

static inline long long fem_ntile(long long n) { return (n + FTS - 1) / FTS; }
static size_t fem_ws_doubles(long long nt) { return (size_t)(REC + 1 + 6 + 3 * FT) * (size_t)nt; }

extern "C" size_t hfl_fem_p1_workspace_bytes(int64_t n_nodes) {
    if (n_nodes < 2) return 256;
    const long long nt = fem_ntile(n_nodes);
    return fem_ws_doubles(nt) * sizeof(double) + 256;
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
    if (coarse_solver == HFL_COARSE_FLUX_SCAN) {
        int rc = hfl_fem_flux_scan(a, R, d_u, d_ws, ws_bytes, s);
        if (rc != HFL_OK) return rc;
    } else {
        const long long nt = fem_ntile(n);
        if (nt > (long long)TOPT * TOP_MAX_CHUNK) {
            set_error("hfl_fem_p1_solve: %lld nodes exceed the single-call limit of %lld; split the mesh across GPUs",
                      (long long)n, (long long)TOPT * TOP_MAX_CHUNK * FTS);
            return HFL_ERR_UNSUPPORTED;
        }
        double* rec = reinterpret_cast<double*>(d_ws);
        double* utop = rec + (size_t)REC * nt;
        double* wsrows = utop + nt;
        double* yvw = wsrows + 6 * (size_t)nt;
        const int S = (int)((nt + TOPT - 1) / TOPT);
        const dim3 grid((unsigned)nt, (unsigned)R);
        // top level: rows in shared memory while the CTA stays within the shared-memory carve-out the level-0 kernels
        // already use (<= 96 KB, ~1000 tiles = 2e6 nodes).  Asking for more (220 KB would hold 1e7 nodes) costs more
        // in the carve-out switch between kernels than the saved L2 round trips: measured +17 us at 1e7 nodes.
        const size_t top_small = 8 * TOPT * sizeof(double), top_rows = top_small + 4 * (size_t)nt * sizeof(double);
        const bool top_in_smem = top_rows <= 96 * 1024;
        const size_t top_smem = top_in_smem ? top_rows : top_small;
        if (top_in_smem)
            HFL_CUDA_CHECK(cudaFuncSetAttribute(fem_top_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)top_smem));
        else
            HFL_CUDA_CHECK(cudaFuncSetAttribute(fem_top_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)top_smem));
        auto launch_top = [&]() {
            if (top_in_smem) fem_top_kernel<true><<<dim3(1, R), TOPT, top_smem, s>>>(rec, (int)nt, S, wsrows, utop, a.ws_stride);
            else fem_top_kernel<false><<<dim3(1, R), TOPT, top_smem, s>>>(rec, (int)nt, S, wsrows, utop, a.ws_stride);
        };
        if (a.aq != nullptr) {
            const size_t smem0 = (size_t)sm_total(true) * sizeof(double);
            HFL_CUDA_CHECK(cudaFuncSetAttribute(fem_reduce_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem0));
            HFL_CUDA_CHECK(cudaFuncSetAttribute(fem_backsub_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem0));
            fem_reduce_kernel<true><<<grid, FT, smem0, s>>>(a, rec, yvw);
            launch_top();
            fem_backsub_kernel<true><<<grid, FT, smem0, s>>>(a, utop, (int)nt, yvw, d_u);
        } else if (a.exact_rowsum) {
            const size_t smem0 = (size_t)sm_total(false) * sizeof(double);
            HFL_CUDA_CHECK(cudaFuncSetAttribute(fem_reduce_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem0));
            HFL_CUDA_CHECK(cudaFuncSetAttribute(fem_backsub_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem0));
            fem_reduce_kernel<false, true><<<grid, FT, smem0, s>>>(a, rec, yvw);
            launch_top();
            fem_backsub_kernel<false, true><<<grid, FT, smem0, s>>>(a, utop, (int)nt, yvw, d_u);
        } else {
            const size_t smem0 = (size_t)sm_total(false) * sizeof(double);
            HFL_CUDA_CHECK(cudaFuncSetAttribute(fem_reduce_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem0));
            HFL_CUDA_CHECK(cudaFuncSetAttribute(fem_backsub_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem0));
            fem_reduce_kernel<false><<<grid, FT, smem0, s>>>(a, rec, yvw);
            launch_top();
            fem_backsub_kernel<false><<<grid, FT, smem0, s>>>(a, utop, (int)nt, yvw, d_u);
        }
        count_launch(3);
        HFL_CUDA_CHECK(cudaGetLastError());
    }
    if (d_iface4 != nullptr) {
        const double* flux2 = nullptr;
        if (coarse_solver == HFL_COARSE_FLUX_SCAN)
            flux2 = reinterpret_cast<const double*>(d_ws) + 6 * (size_t)((n - 1 + FTS - 1) / FTS);   // after the tile prefixes (hfl_flux.cu)
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
