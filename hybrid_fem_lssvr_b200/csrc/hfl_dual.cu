// K4: batched per-element dual LSSVR solve (placeholder until the pivoted factorisation lands).
#include "hfl_device.cuh"

using namespace hfl;

extern "C" int hfl_lssvr_dual_batch(const hfl_plan_t* plan, int64_t E, const double* d_nodes, const double* d_u,
                                    int forcing_kind, double k_freq, const double* d_f_samples, const double* d_bc2,
                                    double* d_coef, double* d_fine, int32_t* d_status, double* d_err3, void* stream) {
    (void)plan; (void)E; (void)d_nodes; (void)d_u; (void)forcing_kind; (void)k_freq; (void)d_f_samples; (void)d_bc2;
    (void)d_coef; (void)d_fine; (void)d_status; (void)d_err3; (void)stream;
    set_error("hfl_lssvr_dual_batch: not implemented in this build");
    return HFL_ERR_UNSUPPORTED;
}
