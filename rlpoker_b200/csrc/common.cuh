// Shared helpers of the rlpoker_b200 CUDA library (sm_100a only; compiled with -fmad=false so that
// every floating-point operation rounds exactly as CPython's does in the reference).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/rlpoker_b200.h"

namespace rlp {

extern thread_local char g_error[512];
extern std::atomic<uint64_t> g_launches;
// Planning mode (rlp_tree_plan): derive the whole device layout on the host without touching a GPU.
extern thread_local bool g_dry_run;

int fail(int code, const char *fmt, ...);

#define RLP_CUDA(expr)                                                                                  \
    do {                                                                                                \
        cudaError_t e_ = (expr);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return rlp::fail(RLP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, \
                             __LINE__);                                                                 \
    } while (0)

#define RLP_LAUNCH_CHECK()                                                                              \
    do {                                                                                                \
        rlp::g_launches.fetch_add(1, std::memory_order_relaxed);                                        \
        cudaError_t e_ = cudaGetLastError();                                                            \
        if (e_ != cudaSuccess)                                                                          \
            return rlp::fail(RLP_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_),  \
                             __FILE__, __LINE__);                                                       \
    } while (0)

// Owning device buffer freed with the handle.
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    cudaError_t alloc(size_t count) {
        release();
        n = count;
        if (g_dry_run) return cudaSuccess;
        return cudaMalloc(&p, sizeof(T) * (count ? count : 1));
    }
    cudaError_t upload(const std::vector<T> &h, cudaStream_t s) {
        cudaError_t e = alloc(h.size());
        if (e != cudaSuccess || h.empty() || g_dry_run) return e;
        return cudaMemcpyAsync(p, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice, s);
    }
    size_t bytes() const { return sizeof(T) * n; }
};

// Programmatic dependent launch (sm_90+): a kernel launched with the attribute may start while its predecessor
// in the stream is still running; everything it does before pdl_wait() must touch only data the predecessor
// never writes (static topology).  pdl_wait() returns once the predecessor has completed and flushed.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
inline cudaError_t sync_stream(cudaStream_t s) { return g_dry_run ? cudaSuccess : cudaStreamSynchronize(s); }
bool pdl_enabled();
void pdl_disable();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

constexpr int kMaxActions = 64;        // widest information set the kernels accept
constexpr double kEpsilon = 1e-7;      // cfr_util.py:14 default floor

// CPython >= 3.12 builtin sum() (Neumaier-compensated), used by the reference's two normalisers
// (cfr_util.py:58, cfr.py:35, extensive_game.py:72).
struct Neumaier {
    double s = 0.0, c = 0.0;
    __host__ __device__ inline void add(double x) {
        double t = s + x;
        if (fabs(s) >= fabs(x)) c += (s - t) + x; else c += (x - t) + s;
        s = t;
    }
    __host__ __device__ inline double total() const { return s + c; }
};

// compute_regret_matching + normalise_probs (cfr_util.py:14-58): reads r[0..n), writes out[0..n).
// `r` and `out` may alias.
__device__ inline void regret_matching(const double *r, int n, double *out) {
    double mx = r[0];
    for (int a = 1; a < n; ++a) mx = r[a] > mx ? r[a] : mx;
    if (mx <= 0.0) {
        double u = 1.0 / (double)n;
        for (int a = 0; a < n; ++a) out[a] = u;
        return;
    }
    Neumaier acc;
    for (int a = 0; a < n; ++a) {
        double p = r[a] > 0.0 ? r[a] : 0.0;
        p = p > kEpsilon ? p : kEpsilon;
        acc.add(p);
    }
    double tot = acc.total();
    for (int a = 0; a < n; ++a) {
        double p = r[a] > 0.0 ? r[a] : 0.0;
        p = p > kEpsilon ? p : kEpsilon;
        out[a] = p / tot;
    }
}

// Philox-4x32-10, counter = (draw, traversing player, lane, iteration), key = seed; 53-bit uniform.
__host__ __device__ inline double philox_uniform(uint64_t seed, uint64_t iteration, uint32_t lane, uint32_t trav,
                                                 uint32_t draw) {
    uint32_t c0 = draw, c1 = trav, c2 = lane, c3 = (uint32_t)iteration;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32) ^ (uint32_t)(iteration >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    uint32_t a = c0 >> 5, b = c1 >> 6;
    return ((double)a * 67108864.0 + (double)b) * (1.0 / 9007199254740992.0);     // 2^-53: the same bits as the division
}

}  // namespace rlp
