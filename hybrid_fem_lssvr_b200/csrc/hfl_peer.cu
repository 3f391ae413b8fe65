// Multi-GPU exchange of a few doubles per rank over NVLink peer memory (one process per GPU, one node).
//
// The partitioned coarse solve exchanges 4 doubles per rank (SPIKE interface records) and the error norms 3
// (SURVEY.md section 8e): messages of 24-32 bytes, pure latency.  An NCCL all-gather of that size costs 15-25 us
// per call on NVSwitch; a store into every peer's buffer plus a spin on one's own costs a few.  Protocol (the
// "LL" idea): every 64-bit word on the wire carries 32 bits of payload and the 32-bit epoch of the call, written
// with one 8-byte store, so the receiver needs no fence - it spins (volatile loads, served by its own L2) until the
// word's epoch matches, and a torn double cannot be observed.  Slots are double-buffered by epoch parity: a rank can
// only be two epochs ahead of a peer after that peer has left the epoch in between, so a slot is never overwritten
// before it has been read.  The spin is bounded (2^peer_spin_log2 polls, default 2^24 ~ 8 s; hfl_set_option): on expiry
// the status word is set, the missing doubles are delivered as NaN (and so are the interface values computed from them),
// so a late or dead peer poisons every downstream result instead of passing off the stale payload of an older epoch,
// and the kernel returns instead of hanging the GPU.
//
// The epoch of a call is either given by the host or, with epoch = HFL_PEER_EPOCH_DEVICE (0xFFFFFFFF), kept on the device:
// a per-channel counter behind the words of the rank's own buffer, advanced by the kernel itself.  Every rank makes the
// same sequence of calls, so the counters agree without communication - and a launch no longer carries a number that
// changes from step to step, which is what lets a whole multi-GPU step be captured once and replayed as a CUDA graph.
//
// Buffers are plain cudaMalloc allocations exported with cudaIpcGetMemHandle; the host side (dist.PeerExchange)
// swaps the 64-byte handles through torch.distributed and opens them with cudaIpcOpenMemHandle.
#include <string.h>
#include "hfl_fem.cuh"

namespace hfl {

constexpr int PEER_MAX_RANKS = 64;
constexpr int PEER_MAX_DOUBLES = 4;
constexpr int PEER_CHANNELS = 4;
constexpr int PEER_WORDS = PEER_CHANNELS * 2 * PEER_MAX_RANKS * PEER_MAX_DOUBLES * 2;   // 64-bit payload words per buffer
constexpr unsigned int PEER_EPOCH_DEVICE = 0xFFFFFFFFu;                                 // epoch kept in the buffer's tail
constexpr size_t PEER_BYTES = (size_t)PEER_WORDS * sizeof(unsigned long long) + 64;     // + per-channel epoch counters

__device__ __forceinline__ unsigned long long* peer_slot(void* buf, int channel, unsigned int epoch, int src_rank) {
    return reinterpret_cast<unsigned long long*>(buf) +
           ((size_t)(channel * 2 + (int)(epoch & 1u)) * PEER_MAX_RANKS + src_rank) * (PEER_MAX_DOUBLES * 2);
}

// One CTA.  Thread t handles word j = t % (2 W) of peer p = t / (2 W) (looping over peers when G * 2 W > blockDim).
// bc2 != NULL (W = 4, channel of the interface records): thread 0 goes on to solve the interface system from the
// gathered records - the exchange and the (G-1)-unknown solve of the partitioned coarse solve in one launch.
__global__ void peer_allgather_kernel(int G, int rank, int W, const double* __restrict__ src, void* const* __restrict__ bufs,
                                      unsigned int epoch, int channel, double* __restrict__ out, int* __restrict__ status,
                                      double uL, double uR, double* __restrict__ bc2, long long max_spin) {
    __shared__ unsigned int halves[PEER_MAX_RANKS * PEER_MAX_DOUBLES * 2];
    __shared__ int s_expired;
    __shared__ unsigned int s_epoch;
    if (threadIdx.x == 0) {
        s_expired = 0;
        s_epoch = epoch;
        if (epoch == PEER_EPOCH_DEVICE) {       // this rank's counter of the channel: 1, 2, ... (0 is the cleared state)
            unsigned int* ctr = reinterpret_cast<unsigned int*>(reinterpret_cast<unsigned long long*>(bufs[rank]) + PEER_WORDS) + channel;
            unsigned int e = *ctr + 1u;
            if (e == 0u || e == PEER_EPOCH_DEVICE) e = (e & 1u) ? 1u : 2u;      // wrap-around keeps the parity alternating
            *ctr = e;
            s_epoch = e;
        }
    }
    __syncthreads();
    epoch = s_epoch;
    const int nw = 2 * W;
    for (int t = threadIdx.x; t < G * nw; t += blockDim.x) {
        const int p = t / nw, j = t - p * nw;
        const unsigned long long bits = (unsigned long long)__double_as_longlong(src[j >> 1]);
        const unsigned int payload = (j & 1) ? (unsigned int)(bits >> 32) : (unsigned int)bits;
        const unsigned long long word = ((unsigned long long)epoch << 32) | payload;
        unsigned long long* dst = peer_slot(bufs[p], channel, epoch, rank) + j;
        asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(dst), "l"(word) : "memory");
    }
    for (int t = threadIdx.x; t < G * nw; t += blockDim.x) {
        const int p = t / nw, j = t - p * nw;
        const unsigned long long* slot = peer_slot(bufs[rank], channel, epoch, p) + j;
        unsigned long long word = 0ull;
        bool got = false;
        for (long long spin = 0; spin < max_spin; ++spin) {
            asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(word) : "l"(slot) : "memory");
            if ((unsigned int)(word >> 32) == epoch) { got = true; break; }
            if (spin > 1024) __nanosleep(256);
        }
        if (!got) {
            s_expired = 1;
            if (status != nullptr) atomicExch(status, 1);
            word = (j & 1) ? 0x7ff80000ull : 0ull;        // the two halves of a quiet NaN
        }
        halves[t] = (unsigned int)word;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < G * W; t += blockDim.x) {
        const unsigned long long bits = ((unsigned long long)halves[2 * t + 1] << 32) | halves[2 * t];
        out[t] = __longlong_as_double((long long)bits);
    }
    if (bc2 != nullptr) {
        __syncthreads();           // out[] is complete (global writes of this CTA are visible to it after the barrier)
        if (threadIdx.x == 0) {
            spike_iface_solve(G, out, uL, uR, rank, bc2);
            if (s_expired) { bc2[0] = __longlong_as_double(0x7ff8000000000000ll); bc2[1] = bc2[0]; }
        }
    }
}

}  // namespace hfl

using namespace hfl;

extern "C" size_t hfl_peer_buffer_bytes(void) { return PEER_BYTES; }

extern "C" int hfl_peer_buffer_create(void** d_buf, unsigned char* ipc_handle64) {
    HFL_REQUIRE(d_buf != nullptr && ipc_handle64 != nullptr, "hfl_peer_buffer_create: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    HFL_CUDA_CHECK(cudaMalloc(&p, hfl_peer_buffer_bytes()));
    HFL_CUDA_CHECK(cudaMemset(p, 0, hfl_peer_buffer_bytes()));
    HFL_CUDA_CHECK(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
        return HFL_ERR_CUDA;
    }
    memcpy(ipc_handle64, &h, 64);
    *d_buf = p;
    return HFL_OK;
}

extern "C" int hfl_peer_buffer_open(const unsigned char* ipc_handle64, void** d_peer) {
    HFL_REQUIRE(ipc_handle64 != nullptr && d_peer != nullptr, "hfl_peer_buffer_open: NULL argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle64, 64);
    void* p = nullptr;
    HFL_CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *d_peer = p;
    return HFL_OK;
}

extern "C" int hfl_peer_buffer_close(void* d_peer) {
    if (d_peer != nullptr) HFL_CUDA_CHECK(cudaIpcCloseMemHandle(d_peer));
    return HFL_OK;
}

extern "C" int hfl_peer_buffer_destroy(void* d_buf) {
    if (d_buf != nullptr) HFL_CUDA_CHECK(cudaFree(d_buf));
    return HFL_OK;
}

extern "C" int hfl_peer_allgather(int G, int rank, int W, const double* d_src, void* const* d_bufs, uint32_t epoch,
                                  int channel, double* d_out, int32_t* d_status, void* stream) {
    HFL_REQUIRE(G >= 1 && G <= PEER_MAX_RANKS, "hfl_peer_allgather: G=%d outside [1, %d]", G, PEER_MAX_RANKS);
    HFL_REQUIRE(rank >= 0 && rank < G, "hfl_peer_allgather: rank outside [0, G)");
    HFL_REQUIRE(W >= 1 && W <= PEER_MAX_DOUBLES, "hfl_peer_allgather: W=%d outside [1, %d]", W, PEER_MAX_DOUBLES);
    HFL_REQUIRE(channel >= 0 && channel < PEER_CHANNELS, "hfl_peer_allgather: channel outside [0, %d)", PEER_CHANNELS);
    HFL_REQUIRE(epoch != 0, "hfl_peer_allgather: epoch 0 is the cleared state of the buffers");
    HFL_REQUIRE(d_src != nullptr && d_bufs != nullptr && d_out != nullptr, "hfl_peer_allgather: NULL pointer");
    peer_allgather_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(G, rank, W, d_src, d_bufs, epoch, channel, d_out, d_status,
                                                               0.0, 0.0, nullptr, 1ll << get_option_peer_spin_log2());
    count_launch();
    HFL_CUDA_CHECK(cudaGetLastError());
    return HFL_OK;
}

extern "C" int hfl_peer_spike_exchange(int G, int rank, const double* d_iface4, void* const* d_bufs, uint32_t epoch,
                                       int channel, double u_left, double u_right, double* d_gathered, double* d_bc2,
                                       int32_t* d_status, void* stream) {
    HFL_REQUIRE(G >= 1 && G <= PEER_MAX_RANKS, "hfl_peer_spike_exchange: G=%d outside [1, %d]", G, PEER_MAX_RANKS);
    HFL_REQUIRE(rank >= 0 && rank < G, "hfl_peer_spike_exchange: rank outside [0, G)");
    HFL_REQUIRE(channel >= 0 && channel < PEER_CHANNELS, "hfl_peer_spike_exchange: channel outside [0, %d)", PEER_CHANNELS);
    HFL_REQUIRE(epoch != 0, "hfl_peer_spike_exchange: epoch 0 is the cleared state of the buffers");
    HFL_REQUIRE(d_iface4 != nullptr && d_bufs != nullptr && d_gathered != nullptr && d_bc2 != nullptr,
                "hfl_peer_spike_exchange: NULL pointer");
    peer_allgather_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(G, rank, 4, d_iface4, d_bufs, epoch, channel, d_gathered,
                                                               d_status, u_left, u_right, d_bc2, 1ll << get_option_peer_spin_log2());
    count_launch();
    HFL_CUDA_CHECK(cudaGetLastError());
    return HFL_OK;
}
