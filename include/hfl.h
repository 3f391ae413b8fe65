/*
 * hfl.h - C ABI of libhfl.so: the hybrid FEM + LSSVR hot path on NVIDIA B200 (sm_100a), FP64.
 *
 * The reference (maryambabaei/hybrid-FEM-LSSVR) has no FFI: its call surface is Python
 * (P: = 1D-Possion/Hybrid-FEM-LSSVR.py, D: = 1D-Possion/Hybrid-FEM-LSSVR-Dual.py).  Each entry
 * point below names the reference code it replaces.  All pointers named d_* are DEVICE pointers;
 * everything is stream-ordered on `stream` (a cudaStream_t passed as void*, NULL = legacy default
 * stream); no entry point synchronises the device or allocates device memory behind the caller's
 * back except hfl_plan_create (a few KB of tables), hfl_peer_buffer_create (16 KB) and the FIRST
 * hfl_lssvr_dual_* call of a (plan, stream) pair with N >= 48, which may grow a scratch buffer owned
 * by the plan (cudaMalloc; released by hfl_plan_destroy).  Every function returns HFL_OK or an error code
 * and records a message readable through hfl_last_error().  There is no CPU fallback anywhere.
 *
 * Conventions: E elements, n = E + 1 nodes, M Legendre coefficients (degree M - 1), N equispaced
 * collocation points per element (end points included; the reference hard-codes 12 at P:40),
 * F equispaced fine points per element (end points included).
 */
#ifndef HFL_H
#define HFL_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HFL_OK 0
#define HFL_ERR_ARG 1          /* bad argument (message says which) */
#define HFL_ERR_CUDA 2         /* a CUDA call failed (message carries cudaGetErrorString) */
#define HFL_ERR_UNSUPPORTED 3  /* valid request outside what this build covers */

#define HFL_MAX_M 32           /* largest number of Legendre coefficients */
#define HFL_MAX_N 256          /* largest number of collocation points */
#define HFL_MAX_F 256          /* largest number of fine points per element */

/* forcing_kind */
#define HFL_FORCING_SINE 0     /* f(x) = (k pi)^2 sin(k pi x) evaluated on the device; k = 1 is poisson_rhs, P:11-12 */
#define HFL_FORCING_SAMPLES 1  /* f given as samples d_f[j * E + e] = f(x_e + j h_e / (N - 1)), j < N (any rhs_func, P:20/P:45) */

/* coarse_solver */
#define HFL_COARSE_ASSEMBLED_PCR 0  /* the reference's assembled tridiagonal system: three-level partition (register-resident
                                       chunks of 8 nodes, cyclic reduction over the chunk heads, one CTA for the tile heads) */
#define HFL_COARSE_FLUX_SCAN 1      /* same equations in first-order (flux) form by two prefix sums; better conditioned */
#define HFL_COARSE_ASSEMBLED_EXACT 2 /* the assembled system with the UNROUNDED diagonal k_l + k_r (zero row sums), same
                                        partition kernels.  The reference's rounded diagonal fl(k_l + k_r) acts as a
                                        spurious reaction term eps k_i u_i: invisible at the reference's sizes (both
                                        modes are within 1e-10 of it up to ~1e4 nodes), 3e-6 at 1e6 nodes and 1e-3 at 1e7
                                        nodes on unlucky meshes.  Mode 0 reproduces that system faithfully; mode 2 (and
                                        mode 1) solve what it was meant to be.  The multi-GPU split uses mode 2 (or 1):
                                        its interface system relies on discrete harmonic functions being linear. */

typedef struct hfl_plan hfl_plan_t;

const char* hfl_version(void);
const char* hfl_last_error(void);          /* thread-local message of the last failing call */
int hfl_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- plans: element-independent tables (Legendre second derivatives at the collocation points,
 * Gram blocks, fine-grid basis values), built on the host in extended precision and kept on the
 * device.  Replaces the per-call `Legendre(w, domain)` / `.deriv(2)` object churn of P:43-62. */
int hfl_plan_create(hfl_plan_t** plan, int M, int N, int F, double gamma);
int hfl_plan_destroy(hfl_plan_t* plan);

/* ---- mesh: d_out[i] = numpy.linspace(a, b, n_global)[i0 + i], i < n_local, bit for bit
 * (P:120 `np.linspace(global_domain[0], global_domain[1], num_fem_nodes)`). */
int hfl_mesh_linspace(double a, double b, int64_t n_global, int64_t i0, int64_t n_local,
                      double* d_out, void* stream);

/* ---- K1: coarse P1 FEM nodal solve, replaces FEMLSSVRPrimalSolver.solve_fem (P:117-145).
 * Stiffness and the 2-point-Gauss load of -u'' = (k pi)^2 sin(k pi x) are formed on the fly from
 * d_nodes[n]; u(d_nodes[0]) = u_left and u(d_nodes[n-1]) = u_right (both 0 in the reference, P:137).
 * d_iface4 (optional, 4 doubles): {x_first, x_last, r_left, r_right}, the record the multi-GPU
 * interface system needs from this range: r_left = load_0 + k_0 (u_1 - u_0), r_right = load_{n-1} +
 * k_{n-2} (u_{n-2} - u_{n-1}) are the end-node residuals.
 * Workspace: hfl_fem_p1_workspace_bytes(n) bytes, 256-byte aligned. */
size_t hfl_fem_p1_workspace_bytes(int64_t n_nodes);
int hfl_fem_p1_solve(int64_t n_nodes, const double* d_nodes, double k_freq,
                     double u_left, double u_right, int coarse_solver,
                     double* d_u, double* d_iface4,
                     void* d_workspace, size_t workspace_bytes, void* stream);

/* R solves on the same mesh in one launch sequence, one forcing frequency each (BASELINE configs[4]: the nodal data
 * of hfl_lssvr_dual_multi; R calls of solve_fem P:117-145 with rhs (k pi)^2 sin(k pi x)): d_k_freq [R] (device),
 * d_u [R][n_nodes].  Each row is bit-identical to hfl_fem_p1_solve with that k_freq.  1 <= R <= 65535.
 * Workspace: hfl_fem_p1_multi_workspace_bytes(n, R) bytes, 256-byte aligned. */
size_t hfl_fem_p1_multi_workspace_bytes(int64_t n_nodes, int R);
int hfl_fem_p1_solve_multi(int64_t n_nodes, const double* d_nodes, int R, const double* d_k_freq,
                           double u_left, double u_right, int coarse_solver,
                           double* d_u, void* d_workspace, size_t workspace_bytes, void* stream);

/* Coarse P1 solve of the general operator -(a u')' + c u = f (stiffness of a, mass matrix of c, load of f, each by
 * the same 2-point Gauss rule): d_aq, d_cq, d_fq are samples [2][n-1] at the two Gauss points
 * x_e + h_e (1/2 -+ 1/(2 sqrt 3)) of every element (d_cq may be NULL = 0).  Assembled three-level partition solver in
 * row-sum form; same workspace as hfl_fem_p1_solve. */
int hfl_fem_p1_solve_general(int64_t n_nodes, const double* d_nodes,
                             const double* d_aq, const double* d_cq, const double* d_fq,
                             double u_left, double u_right, double* d_u,
                             void* d_workspace, size_t workspace_bytes, void* stream);

/* Interface (SPIKE) system of a mesh split into G contiguous ranges, one per GPU.  Host function.
 * gathered[4 * r + {0,1,2,3}] = {x_first, x_last, r_left, r_right} of rank r (its local solve had
 * zero Dirichlet data at both ends).  Writes the G + 1 interface values (iface[0] = u_left,
 * iface[G] = u_right); rank r then adds the linear correction with bc = {iface[r], iface[r+1]}. */
int hfl_spike_interface_solve(int G, const double* gathered, double u_left, double u_right,
                              double* iface);

/* Same solve on the device (G <= 64), stream-ordered: d_gathered[4 G] as above (e.g. straight out of
 * an NCCL all-gather); writes d_bc2 = {iface[rank], iface[rank + 1]} for hfl_fem_apply_bc / d_bc2. */
int hfl_spike_interface_solve_device(int G, const double* d_gathered, double u_left, double u_right,
                                     int rank, double* d_bc2, void* stream);

/* ---- exchange of a few doubles per rank over NVLink peer memory (one process per GPU, one node): the two
 * collectives of the partitioned path (4 doubles per rank for the interface system, 3 for the error norms) without
 * NCCL's per-call latency.  Every rank creates one buffer, the 64-byte IPC handles are swapped by the host (any
 * transport), every rank opens its peers' buffers and keeps the G device pointers in a device array d_bufs[G]
 * (its own buffer at [rank]).  hfl_peer_allgather: d_out[G][W] <- every rank's d_src[W] (W <= 4), stream-ordered;
 * all ranks must call it with the same (epoch, channel, W); epoch != 0 increases by one per call on a channel
 * (channel < 4 separates call sites), or every call passes HFL_PEER_EPOCH_DEVICE and the kernel keeps the counter in the
 * rank's own buffer (do not mix the two on one channel).  The receive spin is bounded (2^24 polls, ~8 s; hfl_set_option "peer_spin_log2"):
 * on expiry *d_status (optional) becomes 1 and the doubles that did not arrive are delivered as NaN (hfl_peer_spike_exchange
 * then also writes NaN interface values), so a late or dead peer poisons the results instead of passing off stale data.
 * Replaces the dist.all_gather_into_tensor calls a torch.distributed port of P:117-145 / K5 would make. */
#define HFL_PEER_EPOCH_DEVICE 0xFFFFFFFFu   /* epoch argument: use and advance the per-channel counter kept in the buffer (graph-replayable) */
size_t hfl_peer_buffer_bytes(void);
int hfl_peer_buffer_create(void** d_buf, unsigned char* ipc_handle64);
int hfl_peer_buffer_open(const unsigned char* ipc_handle64, void** d_peer);
int hfl_peer_buffer_close(void* d_peer);
int hfl_peer_buffer_destroy(void* d_buf);
int hfl_peer_allgather(int G, int rank, int W, const double* d_src, void* const* d_bufs, uint32_t epoch,
                       int channel, double* d_out, int32_t* d_status, void* stream);
/* The interface exchange of the partitioned coarse solve in one launch: all-gather of the 4-double records
 * (d_iface4 from hfl_fem_p1_solve) into d_gathered[4 G], then hfl_spike_interface_solve_device's solve -> d_bc2. */
int hfl_peer_spike_exchange(int G, int rank, const double* d_iface4, void* const* d_bufs, uint32_t epoch,
                            int channel, double u_left, double u_right, double* d_gathered, double* d_bc2,
                            int32_t* d_status, void* stream);

/* d_u[i] += bc_left * (x_last - x_i) / L + bc_right * (x_i - x_first) / L  (discrete-harmonic
 * correction of a local solve; the element kernels can apply it on the fly through d_bc2). */
int hfl_fem_apply_bc(int64_t n_nodes, const double* d_nodes, double* d_u,
                     double bc_left, double bc_right, void* stream);

/* ---- K2 + K3 (+ K5): batched per-element primal LSSVR solve, replaces the loop of
 * solve_lssvr_subproblems (P:147-176) over lssvr_primal (P:20-105) and, through d_fine, the
 * structured part of evaluate_solution (P:184-211).
 *   d_nodes[E+1], d_u[E+1]  nodal abscissae and FEM values (element e uses entries e, e+1)
 *   d_bc2        optional {bc_left, bc_right}: nodal values are corrected as in hfl_fem_apply_bc
 *   d_coef       optional out, [E][M] Legendre coefficients (P:98 `res.x[:M]`)
 *   d_fine       optional out, [E][F] u at linspace(x_e, x_{e+1}, F)
 *   d_status     optional out, [E]: 0 ok, 1 = factorisation broke down and the element fell back
 *                to the linear interpolant of its two nodal values (P:171-176)
 *   d_err3       optional in/out accumulators vs sin(k pi x) on the fine grid:
 *                [0] += sum_e h_e/(F-1) * trapezoid_i err^2, [1] = max(., max|err|), [2] += failed elements
 */
int hfl_lssvr_primal_batch(const hfl_plan_t* plan, int64_t E,
                           const double* d_nodes, const double* d_u,
                           int forcing_kind, double k_freq, const double* d_f_samples,
                           const double* d_bc2,
                           double* d_coef, double* d_fine, int32_t* d_status, double* d_err3,
                           void* stream);

/* ---- K4: batched per-element dual LSSVR solve ((N+2) x (N+2) kernel system
 * [[A A^T + I/gamma, A B^T], [B A^T, B B^T]] [alpha; beta] = [f; g], w = A^T alpha + B^T beta).
 * The reference ships no dual code (D: is a copy of P:); arguments as hfl_lssvr_primal_batch. */
int hfl_lssvr_dual_batch(const hfl_plan_t* plan, int64_t E,
                         const double* d_nodes, const double* d_u,
                         int forcing_kind, double k_freq, const double* d_f_samples,
                         const double* d_bc2,
                         double* d_coef, double* d_fine, int32_t* d_status, double* d_err3,
                         void* stream);

/* Same with R right-hand sides per element sharing one factorisation (BASELINE configs[4]: a batch of
 * forcing frequencies): d_u [R][E+1], d_k_freq [R] (device), d_f_samples [R][N][E], d_coef [R][E][M],
 * d_fine [R][E][F], d_status [E], d_err3 [R][3]. */
int hfl_lssvr_dual_multi(const hfl_plan_t* plan, int64_t E, int R,
                         const double* d_nodes, const double* d_u,
                         int forcing_kind, const double* d_k_freq, const double* d_f_samples,
                         const double* d_bc2,
                         double* d_coef, double* d_fine, int32_t* d_status, double* d_err3,
                         void* stream);

/* ---- general 1-D elliptic operator  -(a u')' + c u = f  (the reference codes the Poisson residual only, P:43-45; its
 * README advertises elliptic problems in general).  Same QP with PDE rows (L phi_k)(x_j); a, a', c, f are samples
 * [N][E] at the collocation points (d_da / d_c may be NULL = zero).  With a = 1, a' = c = 0 this is
 * hfl_lssvr_primal_batch with HFL_FORCING_SAMPLES.  M <= 12. */
int hfl_lssvr_general_batch(const hfl_plan_t* plan, int64_t E,
                            const double* d_nodes, const double* d_u,
                            const double* d_a, const double* d_da, const double* d_c, const double* d_f,
                            const double* d_bc2,
                            double* d_coef, double* d_fine, int32_t* d_status, void* stream);

/* ---- K3 unstructured: replaces FEMLSSVRPrimalSolver.evaluate_solution (P:184-211).
 * For each query x: first element j with nodes[j] <= x <= nodes[j+1] (a shared node goes to the
 * LEFT element), element 0 / E-1 outside the mesh; value = numpy legval(off + scl x, coef[j]). */
int hfl_evaluate_points(int64_t E, const double* d_nodes, int M, const double* d_coef,
                        int64_t P, const double* d_x, double* d_out, void* stream);

/* ---- K5: error norms against sin(k pi x).
 * hfl_error_fine: same accumulators as d_err3 above from a stored fine grid [E][F].
 * hfl_error_nodal: [0] += sum_i (u_i - sin)^2 * (x_{i+1} - x_{i-1}) / 2, [1] = max|err|. */
int hfl_error_fine(int64_t E, int F, const double* d_nodes, const double* d_fine, double k_freq,
                   double* d_err3, void* stream);
int hfl_error_nodal(int64_t n_nodes, const double* d_nodes, const double* d_u, double k_freq,
                    double* d_err3, void* stream);

/* ---- tuning / introspection (bench and tests) */
int hfl_set_option(const char* key, int value);   /* "peer_spin_log2": 4..40; "fem_top_smem_kb": 64..227; "primal_store": 0 auto, 1 direct, 2 smem, 3 tma, 4 tma rows, 5 warp-cooperative;
                                                      "dual_team": 1 generic warp kernel for N = 12, 2 full-system team kernel, 3 parity split in shared memory;
                                                      "dual_reuse_factor": 0 factorise every element in the dual kernels (default 1: elements whose tau is below half an ulp of the diagonal share the tau = 0 matrix bit for bit and take its solution map - plan tables - instead of a factorisation);
                                                      "primal_debug": profiling aid */
int hfl_get_option(const char* key, int* value);
int64_t hfl_launch_count(void);                    /* kernels launched by this library so far */
/* FP64 FMA throughput probe: launches `blocks` CTAs of 256 threads, 16 independent DFMA chains of
 * length `iters` each; *flops receives the flop count so the caller can time it with CUDA events. */
int hfl_fp64_probe(int blocks, int iters, double* d_out, double* flops, void* stream);
/* Store-path probe: n distinct doubles written with a given launch shape; pattern 0 = plain stream, 1 = contiguous
 * per-CTA tiles of `tile` doubles visited like the element kernel does. */
int hfl_store_probe(int blocks, int threads, int64_t n, int pattern, int tile, double* d_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HFL_H */
