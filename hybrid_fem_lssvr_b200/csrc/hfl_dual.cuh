// Argument blocks shared by the dual (K4) kernels: hfl_dual.cu (team kernels), hfl_dual_parity.cu (left-looking parity kernel).
#pragma once
#include "hfl_device.cuh"

namespace hfl {

struct DualArgs {
    long long E;
    int R;                  // right-hand sides per element
    const double* nodes;    // [E+1]
    const double* u;        // [R][E+1]
    const double* f;        // samples [R][N][E] or NULL
    const double* kf;       // [R] forcing frequencies (device) or NULL -> k_scalar
    double k_scalar;
    const double* bc2;      // optional {bc_left, bc_right}
    double* coef;           // optional [R][E][M]
    double* fine;           // optional [R][E][F]
    int* status;            // optional [E]
    double* err3;           // optional [R][3]
    const double* K0;       // [n][n]
    const double* Ct;       // [n][M]
    const double* V;        // [F][M]
    int M, N, F, n, ld;
    int forcing;
    double c_tau;           // 1 / (16 gamma)
    bool want_err;
};

constexpr int DUAL_NMAX = 160 + 2;     // largest system this kernel holds in shared memory

// Parity split (even N): the (N+2) system decouples into an even and an odd block of nh = N/2 + 1 unknowns.
struct DualParityArgs {
    DualArgs d;
    const double* Kp[2];    // [nh][nh] per parity
    const double* Cp[2];    // [nh][MA] per parity
    int nh, ldh, MA[2];
    const double* Vt;       // [M][F]
    double* spill;          // left-looking kernel: global scratch for L columns kc.. (per CTA and team)
    int kc;                 // L columns held in shared memory
    int reuse;              // left-looking kernel: keep the factor while the element matrix is bitwise unchanged
    double* dual0;          // left-looking kernel: plan-level tables of the tau = 0 factorisation {mom, tmom, ranks} or NULL
    int dual0_mode;         // 0: factorise in the kernel; 1: read dual0; 2: this launch (one CTA) only fills dual0
};

// hfl_dual_parity.cu (left-looking parity kernel); returns false when the shape is not covered (nh > 96 or not
// enough shared memory) and the caller falls back to the shared-memory right-looking kernel
int launch_dual_parity_left(DualParityArgs pa, int max_smem, const hfl_plan* plan, cudaStream_t s);   // HFL_OK | HFL_ERR_UNSUPPORTED (fall back) | HFL_ERR_CUDA

}  // namespace hfl
