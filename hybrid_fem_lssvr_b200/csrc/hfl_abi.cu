// libhfl: error reporting, options, plans (host-built tables) and the mesh generator.
#include "hfl_common.cuh"
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstring>

namespace hfl {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
static std::atomic<int> g_opt_store{0};
static std::atomic<int> g_opt_debug{0};
static std::atomic<int> g_opt_dual_team{0};
static std::atomic<int> g_opt_top_smem_kb{227};
static std::atomic<int> g_opt_peer_spin_log2{24};
static std::atomic<int> g_opt_dual_reuse{1};
static std::atomic<int> g_sm_count{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n); }
int get_option_store() { return g_opt_store.load(); }
int get_option_debug() { return g_opt_debug.load(); }
int get_option_dual_team() { return g_opt_dual_team.load(); }
int get_option_top_smem_kb() { return g_opt_top_smem_kb.load(); }
int get_option_peer_spin_log2() { return g_opt_peer_spin_log2.load(); }
int get_option_dual_reuse() { return g_opt_dual_reuse.load(); }

int sm_count() {
    int v = g_sm_count.load();
    if (v > 0) return v;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    g_sm_count.store(v);
    return v;
}

// P_k, P_k', P_k'' at x for k < M (three-term recurrence, extended precision on the host).
static void legendre012(int M, long double x, long double* P, long double* d1, long double* d2) {
    P[0] = 1.0L; d1[0] = 0.0L; d2[0] = 0.0L;
    if (M > 1) { P[1] = x; d1[1] = 1.0L; d2[1] = 0.0L; }
    for (int k = 1; k + 1 < M; ++k) {
        long double a = (long double)(2 * k + 1), b = (long double)k, c = (long double)(k + 1);
        P[k + 1] = (a * x * P[k] - b * P[k - 1]) / c;
        d1[k + 1] = (a * (P[k] + x * d1[k]) - b * d1[k - 1]) / c;
        d2[k + 1] = (a * (2.0L * d1[k] + x * d2[k]) - b * d2[k - 1]) / c;
    }
}

// i-th point (ascending) of the non-negative half of the symmetric grid -1 + 2 j / (n - 1).
static long double half_point(int n, int i) {
    if (n == 1) return 0.0L;
    int num = (n % 2 == 0) ? (2 * i + 1) : (2 * i);
    return (long double)num / (long double)(n - 1);
}

}  // namespace hfl

using namespace hfl;

extern "C" const char* hfl_version(void) { return "hfl 0.1 (sm_100a, fp64)"; }
extern "C" const char* hfl_last_error(void) { return g_err; }
extern "C" int64_t hfl_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" int hfl_device_info(int* sms, int* major, int* minor) {
    int dev = 0;
    HFL_CUDA_CHECK(cudaGetDevice(&dev));
    cudaDeviceProp p;
    HFL_CUDA_CHECK(cudaGetDeviceProperties(&p, dev));
    if (sms) *sms = p.multiProcessorCount;
    if (major) *major = p.major;
    if (minor) *minor = p.minor;
    return HFL_OK;
}

extern "C" int hfl_set_option(const char* key, int value) {
    HFL_REQUIRE(key != nullptr, "hfl_set_option: key is NULL");
    if (strcmp(key, "primal_store") == 0) {
        HFL_REQUIRE(value >= 0 && value <= 5, "primal_store must be 0..5");
        g_opt_store.store(value);
        return HFL_OK;
    }
    // 1 = skip the N = 12 register kernel; 2 = also skip the parity-split kernels; 3 = parity split in shared memory only
    if (strcmp(key, "dual_team") == 0) {
        HFL_REQUIRE(value >= 0 && value <= 3, "dual_team must be 0..3");
        g_opt_dual_team.store(value);
        return HFL_OK;
    }
    // K1 top level: tile-head rows live in shared memory while 64 KB + 32 B per tile fit in this many KB
    if (strcmp(key, "fem_top_smem_kb") == 0) {
        HFL_REQUIRE(value >= 64 && value <= 227, "fem_top_smem_kb must be 64..227");
        g_opt_top_smem_kb.store(value);
        return HFL_OK;
    }
    // receive spin of the peer-memory exchange: 2^value polls before the call gives up and poisons its output with NaN
    if (strcmp(key, "peer_spin_log2") == 0) {
        HFL_REQUIRE(value >= 4 && value <= 40, "peer_spin_log2 must be 4..40");
        g_opt_peer_spin_log2.store(value);
        return HFL_OK;
    }
    // dual kernels: 1 = elements whose K + tau J is the tau = 0 matrix bit for bit (tau below half an ulp of every diagonal
    // entry) take that matrix's solution map from plan tables (left-looking kernel: same bits as a factorisation per element;
    // register kernel: equal to rounding); 0 = factorise every element
    if (strcmp(key, "dual_reuse_factor") == 0) {
        HFL_REQUIRE(value == 0 || value == 1, "dual_reuse_factor must be 0 or 1");
        g_opt_dual_reuse.store(value);
        return HFL_OK;
    }
    if (strcmp(key, "primal_debug") == 0) {
        HFL_REQUIRE(value >= 0 && value <= 3, "primal_debug must be 0..3");
        g_opt_debug.store(value);
        return HFL_OK;
    }
    set_error("hfl_set_option: unknown key '%s'", key);
    return HFL_ERR_ARG;
}

extern "C" int hfl_get_option(const char* key, int* value) {
    HFL_REQUIRE(key != nullptr && value != nullptr, "hfl_get_option: NULL argument");
    if (strcmp(key, "primal_store") == 0) { *value = g_opt_store.load(); return HFL_OK; }
    set_error("hfl_get_option: unknown key '%s'", key);
    return HFL_ERR_ARG;
}

extern "C" int hfl_plan_create(hfl_plan_t** out, int M, int N, int F, double gamma) {
    HFL_REQUIRE(out != nullptr, "hfl_plan_create: plan pointer is NULL");
    HFL_REQUIRE(M >= 3 && M <= HFL_MAX_M, "hfl_plan_create: M=%d outside [3, %d]", M, HFL_MAX_M);
    HFL_REQUIRE(N >= 2 && N <= HFL_MAX_N, "hfl_plan_create: N=%d outside [2, %d]", N, HFL_MAX_N);
    HFL_REQUIRE(F >= 0 && F <= HFL_MAX_F && F != 1, "hfl_plan_create: F=%d outside {0, 2..%d}", F, HFL_MAX_F);
    HFL_REQUIRE(gamma > 0.0 && std::isfinite(gamma), "hfl_plan_create: gamma must be positive and finite");
    hfl_plan* p = new hfl_plan();
    p->M = M; p->N = N; p->F = F; p->gamma = gamma;
    p->me = n_even(M); p->mo = n_odd(M);
    p->NH = (N + 1) / 2; p->FH = (F + 1) / 2;
    const int me = p->me, mo = p->mo, NH = p->NH, FH = p->FH;
    std::vector<long double> P(M), d1(M), d2(M);

    p->De.assign((size_t)NH * me, 0.0);
    p->Do.assign((size_t)NH * (mo > 0 ? mo : 1), 0.0);
    std::vector<long double> GeL((size_t)me * me, 0.0L), GoL((size_t)(mo > 0 ? mo * mo : 1), 0.0L);
    for (int j = 0; j < NH; ++j) {
        long double x = half_point(N, j);
        legendre012(M, x, P.data(), d1.data(), d2.data());
        long double wj = (N % 2 == 1 && j == 0) ? 1.0L : 2.0L;   // self-paired middle point
        for (int a = 0; a < me; ++a) p->De[(size_t)j * me + a] = (double)(wj * d2[2 + 2 * a]);
        for (int b = 0; b < mo; ++b) p->Do[(size_t)j * mo + b] = (double)(wj * d2[3 + 2 * b]);
        for (int a = 0; a < me; ++a)
            for (int c = 0; c < me; ++c) GeL[(size_t)a * me + c] += wj * d2[2 + 2 * a] * d2[2 + 2 * c];
        for (int b = 0; b < mo; ++b)
            for (int c = 0; c < mo; ++c) GoL[(size_t)b * mo + c] += wj * d2[3 + 2 * b] * d2[3 + 2 * c];
    }
    p->Ge.assign((size_t)me * (me + 1) / 2, 0.0);
    p->Go.assign((size_t)(mo > 0 ? mo * (mo + 1) / 2 : 1), 0.0);
    for (int a = 0; a < me; ++a)
        for (int c = 0; c <= a; ++c) p->Ge[(size_t)a * (a + 1) / 2 + c] = (double)GeL[(size_t)a * me + c];
    for (int b = 0; b < mo; ++b)
        for (int c = 0; c <= b; ++c) p->Go[(size_t)b * (b + 1) / 2 + c] = (double)GoL[(size_t)b * mo + c];

    p->fineE.assign((size_t)(FH > 0 ? FH : 1) * me, 0.0);
    p->fineO.assign((size_t)(FH > 0 ? FH : 1) * (mo + 1), 0.0);
    for (int i = 0; i < FH; ++i) {
        long double x = half_point(F, i);
        legendre012(M, x, P.data(), d1.data(), d2.data());
        for (int a = 0; a < me; ++a) p->fineE[(size_t)i * me + a] = (double)P[2 + 2 * a];
        p->fineO[(size_t)i * (mo + 1)] = (double)x;
        for (int b = 0; b < mo; ++b) p->fineO[(size_t)i * (mo + 1) + 1 + b] = (double)P[3 + 2 * b];
    }
    p->D2.assign((size_t)N * M, 0.0);
    p->D0.assign((size_t)N * M, 0.0);
    p->D1.assign((size_t)N * M, 0.0);
    for (int j = 0; j < N; ++j) {
        long double x = -1.0L + 2.0L * (long double)j / (long double)(N - 1);
        legendre012(M, x, P.data(), d1.data(), d2.data());
        for (int k = 0; k < M; ++k) {
            p->D2[(size_t)j * M + k] = (double)d2[k];
            p->D1[(size_t)j * M + k] = (double)d1[k];
            p->D0[(size_t)j * M + k] = (double)P[k];
        }
    }
    p->V.assign((size_t)(F > 0 ? F : 1) * M, 0.0);
    for (int i = 0; i < F; ++i) {
        long double x = -1.0L + 2.0L * (long double)i / (long double)(F - 1);
        legendre012(M, x, P.data(), d1.data(), d2.data());
        for (int k = 0; k < M; ++k) p->V[(size_t)i * M + k] = (double)P[k];
    }

    {   // dual tables: Ct = [-D; B] and K0 = Ct Ct^T accumulated in extended precision
        const int n = N + 2;
        std::vector<long double> CtL((size_t)n * M);
        for (int j = 0; j < N; ++j) {
            long double x = -1.0L + 2.0L * (long double)j / (long double)(N - 1);
            legendre012(M, x, P.data(), d1.data(), d2.data());
            for (int k = 0; k < M; ++k) CtL[(size_t)j * M + k] = -d2[k];
        }
        for (int k = 0; k < M; ++k) {
            CtL[(size_t)N * M + k] = (k % 2 == 0) ? 1.0L : -1.0L;
            CtL[(size_t)(N + 1) * M + k] = 1.0L;
        }
        p->Ct.resize((size_t)n * M);
        for (size_t i = 0; i < p->Ct.size(); ++i) p->Ct[i] = (double)CtL[i];
        p->K0.resize((size_t)n * n);
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                long double s = 0.0L;
                for (int k = 0; k < M; ++k) s += CtL[(size_t)i * M + k] * CtL[(size_t)j * M + k];
                p->K0[(size_t)i * n + j] = (double)s;
            }
    }

    if (N % 2 == 0) {   // parity blocks of the dual system
        const int NHp = N / 2, nh = NHp + 1, MEA = n_even(M) + 1, MOA = n_odd(M) + 1;
        std::vector<long double> Ce((size_t)nh * MEA), Co((size_t)nh * MOA);
        for (int j = 0; j < NHp; ++j) {
            legendre012(M, half_point(N, j), P.data(), d1.data(), d2.data());
            for (int a = 0; a < MEA; ++a) Ce[(size_t)j * MEA + a] = -d2[2 * a];
            for (int b = 0; b < MOA; ++b) Co[(size_t)j * MOA + b] = -d2[2 * b + 1];
        }
        for (int a = 0; a < MEA; ++a) Ce[(size_t)NHp * MEA + a] = 1.0L;
        for (int b = 0; b < MOA; ++b) Co[(size_t)NHp * MOA + b] = 1.0L;
        auto gram = [&](const std::vector<long double>& C, int m, std::vector<double>& K) {
            K.resize((size_t)nh * nh);
            for (int i = 0; i < nh; ++i)
                for (int j = 0; j < nh; ++j) {
                    long double s = 0.0L;
                    for (int k = 0; k < m; ++k) s += C[(size_t)i * m + k] * C[(size_t)j * m + k];
                    K[(size_t)i * nh + j] = (double)s;
                }
        };
        gram(Ce, MEA, p->Kpe);
        gram(Co, MOA, p->Kpo);
        p->Cpe.resize(Ce.size());
        p->Cpo.resize(Co.size());
        for (size_t i = 0; i < Ce.size(); ++i) p->Cpe[i] = (double)Ce[i];
        for (size_t i = 0; i < Co.size(); ++i) p->Cpo[i] = (double)Co[i];
    } else {
        p->Cpe.assign(1, 0.0); p->Cpo.assign(1, 0.0); p->Kpe.assign(1, 0.0); p->Kpo.assign(1, 0.0);
    }

    // one device block, each table 16-double (128 B) aligned
    std::vector<double> blk;
    auto push = [&](const std::vector<double>& v) {
        while (blk.size() % 16) blk.push_back(0.0);
        size_t off = blk.size();
        blk.insert(blk.end(), v.begin(), v.end());
        return off;
    };
    p->off_De = push(p->De); p->off_Do = push(p->Do); p->off_Ge = push(p->Ge); p->off_Go = push(p->Go);
    p->off_fineE = push(p->fineE); p->off_fineO = push(p->fineO); p->off_D2 = push(p->D2); p->off_V = push(p->V);
    p->off_Ct = push(p->Ct); p->off_K0 = push(p->K0);
    p->off_D0 = push(p->D0); p->off_D1 = push(p->D1);
    p->Vt.assign(p->V.size(), 0.0);
    for (int i = 0; i < (F > 0 ? F : 0); ++i)
        for (int k = 0; k < M; ++k) p->Vt[(size_t)k * F + i] = p->V[(size_t)i * M + k];
    p->off_Vt = push(p->Vt);
    p->off_Cpe = push(p->Cpe); p->off_Cpo = push(p->Cpo); p->off_Kpe = push(p->Kpe); p->off_Kpo = push(p->Kpo);
    p->n_tables = blk.size();
    (void)cudaGetDevice(&p->device);
    cudaError_t e = cudaMalloc((void**)&p->d_tables, blk.size() * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpy(p->d_tables, blk.data(), blk.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        set_error("hfl_plan_create: device table upload failed: %s", cudaGetErrorString(e));
        if (p->d_tables) cudaFree(p->d_tables);
        delete p;
        return HFL_ERR_CUDA;
    }
    *out = p;
    return HFL_OK;
}

int hfl::plan_on_current_device(const hfl_plan* p, const char* who) {
    int dev = -1;
    HFL_CUDA_CHECK(cudaGetDevice(&dev));
    HFL_REQUIRE(dev == p->device, "%s: the plan was created on device %d but the current device is %d (its tables are not "
                "addressable from here); create one plan per device", who, p->device, dev);
    return HFL_OK;
}

double* hfl::plan_scratch(const hfl_plan* p, cudaStream_t s, size_t bytes) {
    std::lock_guard<std::mutex> guard(p->scratch_mu);
    auto& slot = p->scratch[s];
    if (slot.second < bytes) {
        // stream-ordered (no device synchronisation on the launch path, legal under stream capture): the old buffer is
        // released after the work already queued on it, the new one exists before the kernel that follows
        if (slot.first) cudaFreeAsync(slot.first, s);
        slot = {nullptr, 0};
        void* q = nullptr;
        cudaError_t e = cudaMallocAsync(&q, bytes, s);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();      // do not leave the failure latched for the next CUDA call
            set_error("plan scratch: cudaMallocAsync(%zu) failed: %s", bytes, cudaGetErrorString(e));
            return nullptr;
        }
        slot = {q, bytes};
    }
    return reinterpret_cast<double*>(slot.first);
}

extern "C" int hfl_plan_destroy(hfl_plan_t* p) {
    if (!p) return HFL_OK;
    if (p->d_tables) cudaFree(p->d_tables);
    if (p->d_dual0) cudaFree(p->d_dual0);
    for (auto& kv : p->scratch)
        if (kv.second.first) cudaFree(kv.second.first);
    delete p;
    return HFL_OK;
}

// ---- numpy.linspace, bit for bit: y_i = fl(fl(i * step) + a), last point = b (numpy/_core/function_base.py)
__global__ void linspace_kernel(double a, double b, double step, long long n_global, long long i0,
                                long long n_local, double* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n_local; i += stride) {
        long long g = i0 + i;
        double v = __dadd_rn(__dmul_rn((double)g, step), a);
        if (g == n_global - 1 && n_global > 1) v = b;
        out[i] = v;
    }
}

extern "C" int hfl_mesh_linspace(double a, double b, int64_t n_global, int64_t i0, int64_t n_local,
                                 double* d_out, void* stream) {
    HFL_REQUIRE(n_global >= 2, "hfl_mesh_linspace: n_global=%lld < 2", (long long)n_global);
    HFL_REQUIRE(i0 >= 0 && n_local >= 0 && i0 + n_local <= n_global, "hfl_mesh_linspace: range outside the mesh");
    HFL_REQUIRE(d_out != nullptr || n_local == 0, "hfl_mesh_linspace: d_out is NULL");
    if (n_local == 0) return HFL_OK;
    double step = (b - a) / (double)(n_global - 1);
    HFL_REQUIRE(step != 0.0, "hfl_mesh_linspace: zero step");
    int threads = 256;
    long long blocks = (n_local + threads - 1) / threads;
    long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    linspace_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(a, b, step, n_global, i0, n_local, d_out);
    count_launch();
    HFL_CUDA_CHECK(cudaGetLastError());
    return HFL_OK;
}

// ---- FP64 FMA throughput probe (roofline denominator for the compute-bound rows; MEASURED_PEAKS.json
// carries no FP64 figure).  Each thread runs 16 independent DFMA chains; flops = 2 * 16 * iters * threads.
__global__ void __launch_bounds__(256) fp64_probe_kernel(int iters, double seed, double* __restrict__ out) {
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + (double)(threadIdx.x + i);
    const double m = 1.0 - 1e-9, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], m, c);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 123.456) out[0] = s;   // never true; keeps the chains alive
}

extern "C" int hfl_fp64_probe(int blocks, int iters, double* d_out, double* flops, void* stream) {
    HFL_REQUIRE(blocks > 0 && iters > 0 && d_out != nullptr, "hfl_fp64_probe: bad arguments");
    fp64_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, 1.0, d_out);
    count_launch();
    HFL_CUDA_CHECK(cudaGetLastError());
    if (flops) *flops = 2.0 * 16.0 * (double)iters * 256.0 * (double)blocks;
    return HFL_OK;
}

// ---- store-path probe (profiling aid): writes n distinct doubles with a configurable launch shape.
// pattern 0: grid-stride, consecutive threads write consecutive doubles (a plain stream);
// pattern 1: like the element kernel - each CTA writes contiguous tiles of `tile` doubles (thread t writes a
//            16-byte chunk per step, warp = 512 contiguous bytes), tiles visited blockIdx + i * gridDim.
__global__ void store_probe_kernel(long long n, int pattern, int tile, double* __restrict__ out) {
    if (pattern == 0) {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
            out[i] = (double)i * 1e-9;
    } else {
        const long long ntile = n / tile;
        for (long long tb = blockIdx.x; tb < ntile; tb += gridDim.x) {
            double2* o = reinterpret_cast<double2*>(out + tb * tile);
            for (int j = threadIdx.x; j < tile / 2; j += blockDim.x) {
                const double v = (double)(tb * tile + 2 * j) * 1e-9;
                o[j] = make_double2(v, v + 1e-9);
            }
        }
    }
}

extern "C" int hfl_store_probe(int blocks, int threads, int64_t n, int pattern, int tile, double* d_out, void* stream) {
    HFL_REQUIRE(blocks > 0 && threads > 0 && n > 0 && d_out != nullptr && tile > 0 && tile % 2 == 0, "hfl_store_probe: bad arguments");
    store_probe_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(n, pattern, tile, d_out);
    count_launch();
    HFL_CUDA_CHECK(cudaGetLastError());
    return HFL_OK;
}
