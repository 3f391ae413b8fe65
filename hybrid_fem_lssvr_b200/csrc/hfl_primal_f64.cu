// Specialised element kernel for F = 64 fine points per element (TMA-store instantiations), split out of
// hfl_primal.cu so that the translation units compile in parallel.
#include "hfl_primal_dispatch.cuh"

namespace hfl {

int primal_dispatch_fh32(const hfl_plan* plan, const PrimalArgs& a, bool err, int store, cudaStream_t s) {
    return dispatch_M<32>(plan, a, err, store, s);
}

}  // namespace hfl
