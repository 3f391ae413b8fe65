// Shared declarations of libhfl (sm_100a, FP64).  See include/hfl.h for the ABI.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include "../../include/hfl.h"

namespace hfl {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int get_option_store();
int get_option_debug();
int get_option_dual_team();
int get_option_top_smem_kb();
int get_option_peer_spin_log2();
int get_option_dual_reuse();
int sm_count();

#define HFL_CUDA_CHECK(expr)                                                            \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            hfl::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                           __FILE__, __LINE__);                                         \
            return HFL_ERR_CUDA;                                                        \
        }                                                                               \
    } while (0)

#define HFL_REQUIRE(cond, ...)                                                          \
    do {                                                                                \
        if (!(cond)) {                                                                  \
            hfl::set_error(__VA_ARGS__);                                                \
            return HFL_ERR_ARG;                                                         \
        }                                                                               \
    } while (0)

// Number of even / odd Legendre indices k in [2, M-1] (the "bubble" unknowns per parity).
__host__ __device__ constexpr int n_even(int M) { return (M - 1) / 2; }
__host__ __device__ constexpr int n_odd(int M) { return (M - 2) / 2; }

}  // namespace hfl

// Plan: element-independent tables (host copies + device copies).
struct hfl_plan {
    int M, N, F;
    int device = 0;   // the device the tables live on; launches with another current device are refused
    double gamma;
    int me, mo;   // even / odd bubble counts
    int NH, FH;   // ceil(N/2) collocation pairs, ceil(F/2) fine pairs
    // Pair tables on the non-negative half xi+ (ascending).  Pair weight (2, or 1 for the
    // self-paired middle point of an odd count) is folded into De / Do.
    std::vector<double> De;     // [NH][me]  w_j P''_{2+2a}(xi+_j)
    std::vector<double> Do;     // [NH][mo]  w_j P''_{3+2b}(xi+_j)
    std::vector<double> Ge;     // packed lower [me(me+1)/2]  sum_j P''_{2+2a} P''_{2+2a'}
    std::vector<double> Go;     // packed lower [mo(mo+1)/2]
    std::vector<double> fineE;  // [FH][me]    P_{2+2a}(xi+_i)
    std::vector<double> fineO;  // [FH][mo+1]  {xi+_i, P_{3+2b}(xi+_i)}
    // Full tables (dual path, generic kernels)
    std::vector<double> D2;     // [N][M]  P_k''(xi_j)
    std::vector<double> D0, D1; // [N][M]  P_k(xi_j), P_k'(xi_j)  (general operators)
    std::vector<double> V;      // [F][M]  P_k(xi_i)
    std::vector<double> Vt;     // [M][F]  the same, transposed (coalesced reads with one thread per fine point)
    std::vector<double> Ct;     // [N+2][M] rows -P_k''(xi_j) (j < N), then (-1)^k, then 1: the scaled [A; B]
    std::vector<double> K0;     // [N+2][N+2] Ct Ct^T (dual kernel matrix without the tau I block)
    // Parity blocks of the dual system (even N): nh = N/2 + 1 rows {-P''_k(xi+_j), j < N/2; 1}, even / odd k
    std::vector<double> Cpe, Cpo;   // [nh][n_even(M)+1], [nh][n_odd(M)+1]
    std::vector<double> Kpe, Kpo;   // [nh][nh] Cp Cp^T
    // Device block holding all of the above back to back
    double* d_tables = nullptr;
    // Scratch buffers owned by the plan, one per stream that asked for one (kernels on one stream serialise, so a
    // buffer per stream is race-free); grown on demand, released by hfl_plan_destroy.
    // Left-looking dual kernel: moment tables and ranks of the tau = 0 factorisation (a function of the plan alone),
    // computed by the first launch that can use them and read by every later one (hfl_dual_parity.cu).
    // Register dual kernel (N <= 14): the host-built DualSmallTables of this plan (pivot order, permuted blocks, moment
    // tables), built by the first launch (hfl_dual_small.cu)
    mutable std::vector<unsigned char> dual_small_tables;
    mutable double* d_dual0 = nullptr;
    mutable bool dual0_ready = false;
    mutable std::mutex scratch_mu;
    mutable std::map<cudaStream_t, std::pair<void*, size_t>> scratch;
    size_t off_De, off_Do, off_Ge, off_Go, off_fineE, off_fineO, off_D2, off_V, off_Ct, off_K0, off_Cpe, off_Cpo, off_Kpe, off_Kpo, off_D0, off_D1, off_Vt, n_tables;
};

namespace hfl {
// At least `bytes` of device scratch tied to (plan, stream); NULL (and the error string set) when the allocation fails.
double* plan_scratch(const hfl_plan* plan, cudaStream_t s, size_t bytes);
// HFL_OK when the current device is the one the plan was created on, else HFL_ERR_ARG with the message set.
int plan_on_current_device(const hfl_plan* plan, const char* who);
}
