// K1: coarse P1 FEM nodal solve, replaces FEMLSSVRPrimalSolver.solve_fem (P:117-145).
//
// The reference assembles the P1 stiffness and a 2-point-Gauss load with scikit-fem, turns the two
// boundary rows into identity rows and calls a sparse direct solver.  Here the same rounded entries
// (k_e = fl(1/h)^2 * (h/2) summed over the two Gauss points, load by the same two-point rule) are
// formed on the fly from the node array and the tridiagonal system is solved by a two-level
// partition method whose reduced systems are solved by parallel cyclic reduction (PCR):
//
//   level 0  tiles of T*S = 2048 nodes, one CTA each.  Every thread eliminates the S-1 interior nodes
//            of its chunk (two sweeps give the first/last entries of T^-1 b, T^-1 l e_1, T^-1 r e_s),
//            the T-1 chunk heads form a tridiagonal system solved by PCR in shared memory
//            (3 right-hand sides in the reduce pass, 1 in the back-substitution pass).
//   top      one CTA solves the system of tile heads the same way (chunk per thread + PCR).
//
// Kernels: fem_reduce_kernel -> fem_top_kernel -> fem_backsub_kernel.  Node traffic: the node array
// is read twice and u written once (24 B/node); the load is recomputed instead of stored.
#include "hfl_fem.cuh"

namespace hfl {

// Interior of a chunk (rows m0+1 .. m0+S-1): first / last entries of the three partial solutions
// x = y - u_head v - u_next w.   out = {y1, v1, w1, ys, vs, ws}.
template <class Rows>
__device__ __forceinline__ void chunk_reduce(const Rows& rows, int m0, int S, double (&out)[6]) {
    if (S < 2) {   // no interior: neighbouring heads couple directly (x_first = u_next, x_last = u_head)
        out[0] = 0.0; out[1] = 0.0; out[2] = -1.0;
        out[3] = 0.0; out[4] = -1.0; out[5] = 0.0;
        return;
    }
    double l, d, r, b;
    rows.get(m0 + 1, l, d, r, b);
    double dp = d, bp = b, vp = l, rp = r;
    for (int i = 2; i < S; ++i) {
        rows.get(m0 + i, l, d, r, b);
        const double m = l * fast_rcp(dp);
        dp = d - m * rp;
        bp = b - m * bp;
        vp = -m * vp;
        rp = r;
    }
    double inv = fast_rcp(dp);
    out[3] = bp * inv; out[4] = vp * inv; out[5] = rp * inv;
    rows.get(m0 + S - 1, l, d, r, b);
    dp = d; bp = b;
    double wp = r, lp = l;
    for (int i = S - 2; i >= 1; --i) {
        rows.get(m0 + i, l, d, r, b);
        const double m = r * fast_rcp(dp);
        dp = d - m * lp;
        bp = b - m * bp;
        wp = -m * wp;
        lp = l;
    }
    inv = fast_rcp(dp);
    out[0] = bp * inv; out[1] = lp * inv; out[2] = wp * inv;
}

// Parallel cyclic reduction over equations first..last (one per thread, index = thread id), NR
// right-hand sides.  sm holds 2 * (3 + NR) * T doubles.  Every thread of the CTA must call this.
template <int NR, int T>
__device__ __forceinline__ void pcr_solve(double* sm, int t, int first, int last, double l, double d, double r,
                                          double (&rhs)[NR], double (&x)[NR]) {
    constexpr int W = 3 + NR;
    int cur = 0;
    const bool active = (t >= first && t <= last);
    {
        double* bufw = sm + cur * W * T;
        bufw[0 * T + t] = l; bufw[1 * T + t] = d; bufw[2 * T + t] = r;
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
            double dn = d, ln = 0.0, rn = 0.0;
            if (im >= first) {
                const double al = -l * fast_rcp(bufr[1 * T + im]);
                dn = fma(al, bufr[2 * T + im], dn);
                ln = al * bufr[0 * T + im];
#pragma unroll
                for (int q = 0; q < NR; ++q) rhs[q] = fma(al, bufr[(3 + q) * T + im], rhs[q]);
            }
            if (ip <= last) {
                const double be = -r * fast_rcp(bufr[1 * T + ip]);
                dn = fma(be, bufr[0 * T + ip], dn);
                rn = be * bufr[2 * T + ip];
#pragma unroll
                for (int q = 0; q < NR; ++q) rhs[q] = fma(be, bufr[(3 + q) * T + ip], rhs[q]);
            }
            l = ln; d = dn; r = rn;
        }
        bufw[0 * T + t] = l; bufw[1 * T + t] = d; bufw[2 * T + t] = r;
#pragma unroll
        for (int q = 0; q < NR; ++q) bufw[(3 + q) * T + t] = rhs[q];
        __syncthreads();
        cur ^= 1;
    }
    const double inv = fast_rcp(d);
#pragma unroll
    for (int q = 0; q < NR; ++q) x[q] = rhs[q] * inv;
}

// Reduced equation of chunk head t (1 <= t <= T-1) from its own row, the previous chunk's
// {ys, vs, ws} and its own {y1, v1, w1}.
__device__ __forceinline__ void head_equation(double lp, double dp, double rp, double bp, double ys_prev,
                                              double vs_prev, double ws_prev, const double (&six)[6], double& l,
                                              double& d, double& r, double& b) {
    l = -lp * vs_prev;
    d = dp - lp * ws_prev - rp * six[1];
    r = -rp * six[2];
    b = bp - lp * ys_prev - rp * six[0];
}

// Level 0, pass 1: one record per tile = {l, d, r, b of the tile head, y1, v1, w1, ys, vs, ws of the tile interior}.
template <bool SPECIAL>
__device__ __forceinline__ void fem_reduce_body(const FemArgs& a, double* __restrict__ rec, double* __restrict__ yvw,
                                                double* sm) {
    const int t = threadIdx.x;
    const long long P = (long long)blockIdx.x * FTS;
    MeshRows<SPECIAL> rows{sm + SM_K, sm + SM_B, P, a.n, a.uL, a.uR};
    double six[6];
    chunk_reduce(rows, t * FS, FS, six);
    double lp, dp, rp, bp;
    rows.get(t * FS, lp, dp, rp, bp);
    __syncthreads();                 // the element arrays are dead from here: exchange / PCR buffers alias them
    double* ex = sm + SM_EX;
#pragma unroll
    for (int i = 0; i < 6; ++i) ex[i * FT + t] = six[i];
    __syncthreads();
    double l = 0.0, d = 1.0, r = 0.0, rhs[3] = {0.0, 0.0, 0.0}, x[3];
    if (t >= 1) {
        double b;
        head_equation(lp, dp, rp, bp, ex[3 * FT + t - 1], ex[4 * FT + t - 1], ex[5 * FT + t - 1], six, l, d, r, b);
        rhs[0] = b;
        if (t == 1) { rhs[1] = l; l = 0.0; }
        if (t == FT - 1) { rhs[2] = r; r = 0.0; }
    }
    pcr_solve<3, FT>(sm + SM_PCR, t, 1, FT - 1, l, d, r, rhs, x);
    {   // chunk-head partial solutions, read back by the back-substitution pass
        double* o = yvw + (size_t)blockIdx.x * 3 * FT;
        o[t] = x[0]; o[FT + t] = x[1]; o[2 * FT + t] = x[2];
    }
    // x = {Y, V, W} of head t.  The tile's first interior node belongs to chunk 0 (it needs head 1's
    // solution), its last interior node to chunk T-1.  ex[0 .. 3 FT) is dead by now (pcr_solve synchronised).
    if (t == 1) { ex[0] = x[0]; ex[1] = x[1]; ex[2] = x[2]; }
    __syncthreads();
    double* out = rec + (long long)blockIdx.x * REC;
    if (t == 0) {
        const double Y1 = ex[0], V1 = ex[1], W1 = ex[2];
        out[0] = lp; out[1] = dp; out[2] = rp; out[3] = bp;
        out[4] = six[0] - six[2] * Y1;
        out[5] = six[1] - six[2] * V1;
        out[6] = -six[2] * W1;
    }
    if (t == FT - 1) {
        out[7] = six[3] - six[4] * x[0];
        out[8] = -six[4] * x[1];
        out[9] = six[5] - six[4] * x[2];
    }
}

__global__ void __launch_bounds__(FT, 4) fem_reduce_kernel(const FemArgs a, double* __restrict__ rec,
                                                           double* __restrict__ yvw) {
    extern __shared__ double sm[];
    const long long P = (long long)blockIdx.x * FTS;
    load_tile_elements(a, P, sm);
    if (P == 0 || P + FTS >= a.n - 1) fem_reduce_body<true>(a, rec, yvw, sm);
    else fem_reduce_body<false>(a, rec, yvw, sm);
}

// Top level: solve the system of tile heads.  One CTA of TOPT threads, chunk of S heads per thread.
// ws: rows l, d, r, b [4][cnt] followed by Thomas scratch c', b' [2][cnt].
__global__ void __launch_bounds__(TOPT) fem_top_kernel(const double* __restrict__ rec, int cnt, int S,
                                                       double* __restrict__ wsrows, double* __restrict__ utop) {
    extern __shared__ double sm[];
    const int t = threadIdx.x;
    double* rl = wsrows; double* rd = wsrows + cnt; double* rr = wsrows + 2 * (size_t)cnt; double* rb = wsrows + 3 * (size_t)cnt;
    double* tc = wsrows + 4 * (size_t)cnt; double* tb = wsrows + 5 * (size_t)cnt;
    for (int c = t; c < cnt; c += TOPT) {
        const double* rc = rec + (size_t)c * REC;
        double l = 0.0, d = rc[1], r, b = rc[3];
        if (c > 0) {
            const double* rp = rec + (size_t)(c - 1) * REC;
            l = -rc[0] * rp[8];
            d -= rc[0] * rp[9];
            b -= rc[0] * rp[7];
        }
        d -= rc[2] * rc[5];
        r = -rc[2] * rc[6];
        b -= rc[2] * rc[4];
        rl[c] = l; rd[c] = d; rr[c] = r; rb[c] = b;
    }
    __syncthreads();
    ArrayRows rows{rl, rd, rr, rb, cnt};
    double six[6];
    chunk_reduce(rows, t * S, S, six);
    double lp, dp, rp, bp;
    rows.get(t * S, lp, dp, rp, bp);
    double* ex = sm;                 // 6 * TOPT
    double* pcr = sm + 6 * TOPT;     // 2 * 4 * TOPT
#pragma unroll
    for (int i = 0; i < 6; ++i) ex[i * TOPT + t] = six[i];
    __syncthreads();
    double l = 0.0, d = 1.0, r = 0.0, rhs[1] = {0.0}, x[1];
    if (t >= 1) {
        head_equation(lp, dp, rp, bp, ex[3 * TOPT + t - 1], ex[4 * TOPT + t - 1], ex[5 * TOPT + t - 1], six, l, d, r, rhs[0]);
    } else {
        // head 0 is global node 0 (identity row); its equation still carries the coupling to chunk 0's interior
        l = 0.0;
        d = dp - rp * six[1];
        r = -rp * six[2];
        rhs[0] = bp - rp * six[0];
    }
    pcr_solve<1, TOPT>(pcr, t, 0, TOPT - 1, l, d, r, rhs, x);
    double* uh = ex;   // head values, reuse
    __syncthreads();
    uh[t] = x[0];
    __syncthreads();
    const int m0 = t * S;
    if (m0 < cnt) utop[m0] = x[0];
    if (S >= 2 && m0 + 1 < cnt) {
        const double ua = x[0];
        const double ub = (t + 1 < TOPT) ? uh[t + 1] : 0.0;
        // Thomas on rows m0+1 .. m0+S-1 with known neighbours
        double lo, di, ro, bo;
        rows.get(m0 + 1, lo, di, ro, bo);
        bo -= lo * ua;
        if (S == 2) bo -= ro * ub;
        double inv0 = fast_rcp(di);
        double cp = ro * inv0, bpv = bo * inv0;
        if (m0 + 1 < cnt) { tc[m0 + 1] = cp; tb[m0 + 1] = bpv; }
        for (int i = 2; i < S; ++i) {
            rows.get(m0 + i, lo, di, ro, bo);
            if (i == S - 1) bo -= ro * ub;
            const double den = fast_rcp(di - lo * cp);
            cp = ro * den;
            bpv = (bo - lo * bpv) * den;
            if (m0 + i < cnt) { tc[m0 + i] = cp; tb[m0 + i] = bpv; }
        }
        double xv = bpv;
        if (m0 + S - 1 < cnt) utop[m0 + S - 1] = xv;
        for (int i = S - 2; i >= 1; --i) {
            if (m0 + i < cnt) {
                xv = tb[m0 + i] - tc[m0 + i] * xv;
                utop[m0 + i] = xv;
            } else {
                xv = 0.0;   // padded identity rows
            }
        }
    }
}

// Level 0, pass 2: tile head values known -> chunk heads from the stored partial solutions
// (U_t = Y_t - u_P V_t - u_Q W_t) -> chunk interiors by Thomas -> u.
template <bool SPECIAL>
__device__ __forceinline__ void fem_backsub_body(const FemArgs& a, const double* __restrict__ utop, int ntile,
                                                 const double* __restrict__ yvw, double* __restrict__ u, double* sm) {
    const int t = threadIdx.x;
    const long long P = (long long)blockIdx.x * FTS;
    const double uP = utop[blockIdx.x];
    const double uQ = ((int)blockIdx.x + 1 < ntile) ? utop[blockIdx.x + 1] : 0.0;
    double* uh = sm + SM_UH;    // FT head values (+1 for the next tile's head)
    {
        const double* o = yvw + (size_t)blockIdx.x * 3 * FT;
        uh[t] = (t == 0) ? uP : (o[t] - uP * o[FT + t] - uQ * o[2 * FT + t]);
        if (t == 0) uh[FT] = uQ;
    }
    __syncthreads();
    MeshRows<SPECIAL> rows{sm + SM_K, sm + SM_B, P, a.n, a.uL, a.uR};
    const double ua = uh[t], ub = uh[t + 1];
    // Thomas on the chunk interior, compile-time length FS - 1
    double cpv[FS], bpv[FS], xs[FS];
    {
        double lo, di, ro, bo;
        rows.get(t * FS + 1, lo, di, ro, bo);
        bo -= lo * ua;
        double inv = fast_rcp(di);
        cpv[1] = ro * inv; bpv[1] = bo * inv;
#pragma unroll
        for (int i = 2; i < FS; ++i) {
            rows.get(t * FS + i, lo, di, ro, bo);
            if (i == FS - 1) bo -= ro * ub;
            inv = fast_rcp(di - lo * cpv[i - 1]);
            cpv[i] = ro * inv;
            bpv[i] = (bo - lo * bpv[i - 1]) * inv;
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

__global__ void __launch_bounds__(FT, 4) fem_backsub_kernel(const FemArgs a, const double* __restrict__ utop, int ntile,
                                                            const double* __restrict__ yvw, double* __restrict__ u) {
    extern __shared__ double sm[];
    const long long P = (long long)blockIdx.x * FTS;
    load_tile_elements(a, P, sm);   // recomputed: re-reading cached terms (16 B/node) measured slower than 2 sinpi
    if (P == 0 || P + FTS >= a.n - 1) fem_backsub_body<true>(a, utop, ntile, yvw, u, sm);
    else fem_backsub_body<false>(a, utop, ntile, yvw, u, sm);
}

// End-node residuals for the multi-GPU interface system (see hfl.h).
__global__ void fem_reaction_kernel(const FemArgs a, const double* __restrict__ u, double* __restrict__ out4) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double k, Ls, Rs;
        out4[0] = a.nodes[0];
        out4[1] = a.nodes[a.n - 1];
        element_terms(a, a.nodes[0], a.nodes[1], k, Ls, Rs);
        out4[2] = Ls + k * (u[1] - u[0]);
        element_terms(a, a.nodes[a.n - 2], a.nodes[a.n - 1], k, Ls, Rs);
        out4[3] = Rs + k * (u[a.n - 2] - u[a.n - 1]);
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

// Device version of the interface solve (G <= 64): keeps the multi-GPU step stream-ordered.
// gathered[4 r + {0,1,2,3}] = {x_first, x_last, r_left, r_right}; writes bc2 = {U_rank, U_rank+1}.
__global__ void spike_iface_kernel(int G, const double* __restrict__ g, double uL, double uR, int rank,
                                   double* __restrict__ bc2) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
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

using namespace hfl;

static inline long long fem_ntile(long long n) { return (n + FTS - 1) / FTS; }

extern "C" size_t hfl_fem_p1_workspace_bytes(int64_t n_nodes) {
    if (n_nodes < 2) return 256;
    const long long nt = fem_ntile(n_nodes);
    return (size_t)(REC + 1 + 6 + 3 * FT) * (size_t)nt * sizeof(double) + 256;
}

int hfl_fem_flux_scan(const FemArgs& a, double* d_u, void* d_ws, size_t ws_bytes, cudaStream_t s);   // hfl_flux.cu

extern "C" int hfl_fem_p1_solve(int64_t n, const double* d_nodes, double k_freq, double u_left, double u_right,
                                int coarse_solver, double* d_u, double* d_iface4, void* d_ws, size_t ws_bytes,
                                void* stream) {
    HFL_REQUIRE(n >= 2, "hfl_fem_p1_solve: need at least 2 nodes (got %lld)", (long long)n);
    HFL_REQUIRE(d_nodes != nullptr && d_u != nullptr, "hfl_fem_p1_solve: d_nodes / d_u is NULL");
    HFL_REQUIRE(coarse_solver == HFL_COARSE_ASSEMBLED_PCR || coarse_solver == HFL_COARSE_FLUX_SCAN,
                "hfl_fem_p1_solve: unknown coarse_solver %d", coarse_solver);
    HFL_REQUIRE(d_ws != nullptr && ws_bytes >= hfl_fem_p1_workspace_bytes(n),
                "hfl_fem_p1_solve: workspace too small (%zu < %zu)", ws_bytes, hfl_fem_p1_workspace_bytes(n));
    HFL_REQUIRE((reinterpret_cast<uintptr_t>(d_ws) & 255) == 0, "hfl_fem_p1_solve: workspace must be 256-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    const double pi = 3.14159265358979323846;
    FemArgs a;
    a.n = n; a.nodes = d_nodes; a.k = k_freq; a.kpi = k_freq * pi; a.kp2 = a.kpi * a.kpi; a.uL = u_left; a.uR = u_right;
    a.gx0 = 0.5 * (-0.5773502691896257) + 0.5;   // 0.5 * leggauss(2) + 0.5
    a.gx1 = 0.5 * (0.5773502691896257) + 0.5;
    if (coarse_solver == HFL_COARSE_FLUX_SCAN) {
        int rc = hfl_fem_flux_scan(a, d_u, d_ws, ws_bytes, s);
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
        const size_t smem0 = (size_t)SM_TOTAL * sizeof(double);
        {   // per device and cheap: set on every call (the function attributes do not carry over between devices)
            HFL_CUDA_CHECK(cudaFuncSetAttribute(fem_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem0));
            HFL_CUDA_CHECK(cudaFuncSetAttribute(fem_backsub_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem0));
            HFL_CUDA_CHECK(cudaFuncSetAttribute(fem_top_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)(14 * TOPT * sizeof(double))));
        }
        fem_reduce_kernel<<<(unsigned)nt, FT, smem0, s>>>(a, rec, yvw);
        const int S = (int)((nt + TOPT - 1) / TOPT);
        fem_top_kernel<<<1, TOPT, 14 * TOPT * sizeof(double), s>>>(rec, (int)nt, S, wsrows, utop);
        fem_backsub_kernel<<<(unsigned)nt, FT, smem0, s>>>(a, utop, (int)nt, yvw, d_u);
        count_launch(3);
        HFL_CUDA_CHECK(cudaGetLastError());
    }
    if (d_iface4 != nullptr) {
        fem_reaction_kernel<<<1, 32, 0, s>>>(a, d_u, d_iface4);
        count_launch();
        HFL_CUDA_CHECK(cudaGetLastError());
    }
    return HFL_OK;
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
