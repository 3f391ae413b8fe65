// Launch dispatch of the specialised element kernel over (M, error fusion, store path) for one fine-grid size;
// every translation unit that includes this instantiates the kernels for the FH values it dispatches.
#pragma once
#include "hfl_element_kernel.cuh"

namespace hfl {

template <int M, int FH, bool ERR>
static int dispatch_store(const hfl_plan* plan, const PrimalArgs& a, int store, cudaStream_t s) {
    // The store-path option applies to the plain fine-grid kernel.
    // The fused error norms without coefficient output take the instantiation WITHOUT the coefficient path (130 registers
    // instead of 164 at M = 9; measured 0.500 against 0.513 ms - in round 1, before the Horner / Taylor rewrite of the
    // epilogue, it was the other way round).  Coefficient output always takes the TMA-store instantiation.
    if (ERR && a.coef == nullptr) return launch_fast<M, FH, ERR, STORE_TMA, 0, false>(plan, a, s);
    if (a.coef != nullptr || ERR) return launch_fast<M, FH, ERR, STORE_TMA, 0, true>(plan, a, s);
    if constexpr (!ERR && FH != 16) return launch_fast<M, FH, ERR, STORE_TMA, 0, false>(plan, a, s);   // other F: TMA only
    if constexpr (!ERR && FH == 16) {
    switch (store) {
        case STORE_DIRECT: return launch_fast<M, FH, ERR, STORE_DIRECT, 0, false>(plan, a, s);
        case STORE_SMEM: return launch_fast<M, FH, ERR, STORE_SMEM, 0, false>(plan, a, s);
        case STORE_TMA_ROWS: return launch_fast<M, FH, ERR, STORE_TMA_ROWS, 0, false>(plan, a, s);
        case STORE_COOP:
            if constexpr (FH == 16 && M + 3 <= kCoopPitch) return launch_fast<M, FH, ERR, STORE_COOP, 0, false>(plan, a, s);
            else return launch_fast<M, FH, ERR, STORE_TMA, 0, false>(plan, a, s);
        default: return launch_fast<M, FH, ERR, STORE_TMA, 0, false>(plan, a, s);
    }
    }
    return HFL_ERR_UNSUPPORTED;
}

template <int M, int FH>
static int dispatch_err(const hfl_plan* plan, const PrimalArgs& a, bool err, int store, cudaStream_t s) {
    return err ? dispatch_store<M, FH, true>(plan, a, store, s) : dispatch_store<M, FH, false>(plan, a, store, s);
}

template <int FH>
static int dispatch_M(const hfl_plan* plan, const PrimalArgs& a, bool err, int store, cudaStream_t s) {
    switch (plan->M) {
#define HFL_CASE(m)                                                                        \
    case m:                                                                                \
        if constexpr (FH == 0) return launch_fast<m, 0, false, STORE_DIRECT>(plan, a, s);  \
        else return dispatch_err<m, FH>(plan, a, err, store, s);
        HFL_CASE(3) HFL_CASE(4) HFL_CASE(5) HFL_CASE(6) HFL_CASE(7) HFL_CASE(8) HFL_CASE(9) HFL_CASE(10)
        HFL_CASE(11) HFL_CASE(12) HFL_CASE(13) HFL_CASE(14)
#undef HFL_CASE
        default: return -1;
    }
}

}  // namespace hfl
