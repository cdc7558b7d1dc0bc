// Shared internals of libfava_b200: context, error reporting, launch accounting, load helpers.
#pragma once

#include <cuda.h>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cufft.h>
#include <stdint.h>

#include <atomic>
#include <map>
#include <string>
#include <tuple>

#include "fava_b200.h"

namespace fava {

extern thread_local std::string g_last_error;
extern std::atomic<int64_t> g_launches;

int set_error(int code, const char* fmt, ...);

#define FAVA_CHECK_CUDA(expr)                                                                  \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return fava::set_error(FAVA_ECUDA, "%s failed: %s (%s:%d)", #expr,                 \
                                   cudaGetErrorString(e__), __FILE__, __LINE__);               \
    } while (0)

#define FAVA_CHECK_CUFFT(expr)                                                                 \
    do {                                                                                       \
        cufftResult r__ = (expr);                                                              \
        if (r__ != CUFFT_SUCCESS)                                                              \
            return fava::set_error(FAVA_ECUDA, "%s failed: cufftResult %d (%s:%d)", #expr,     \
                                   (int)r__, __FILE__, __LINE__);                              \
    } while (0)

#define FAVA_REQUIRE(cond, ...)                                                                \
    do {                                                                                       \
        if (!(cond)) return fava::set_error(FAVA_EINVAL, __VA_ARGS__);                         \
    } while (0)

// Count and check a kernel launch.
#define FAVA_LAUNCHED()                                                                        \
    do {                                                                                       \
        fava::g_launches.fetch_add(1, std::memory_order_relaxed);                              \
        FAVA_CHECK_CUDA(cudaGetLastError());                                                   \
    } while (0)

constexpr int kNumSMsB200 = 148;

enum WorkspaceSlot { WS_PARTIALS = 0, WS_TABLE = 1, WS_AUX = 2, WS_FFT0 = 3, WS_FFT1 = 4, WS_FFT2 = 5,
                     WS_YPART = 6 /* row partials of the three-axis moment pass */, WS_BINMAP = 7, WS_USER0 = FAVA_WS_USER0, WS_ITEMS0 = 16 /* +axis: cached block-list tables */,
                     WS_COUNT = FAVA_WS_NSLOTS };

struct Staging;  // pinned ring + reader threads (staging.cu)

}  // namespace fava

struct fava_ctx {
    int device = 0;
    int num_sms = fava::kNumSMsB200;
    void* ws[fava::WS_COUNT] = {};
    size_t ws_bytes[fava::WS_COUNT] = {};
    // cuFFT plans keyed by (kind, a, b, c)
    std::map<std::tuple<int, int64_t, int64_t, int64_t>, cufftHandle> plans;
    std::map<std::tuple<int, int64_t, int64_t, int64_t>, size_t> plan_work;  // work-area bytes per plan
    // host copies of the leaf tables whose device CSR tables are cached in WS_ITEMS0 + axis (block_moments.cu)
    std::string item_cache_key[3];
    uint64_t item_cache_uid[3] = {0, 0, 0};  // caller's table id of the cached tables (0 = none given)
    int64_t item_cache_nunits[3] = {0, 0, 0}, item_cache_nent[3] = {0, 0, 0};  // sizes of the cached tables
    std::string prolong_cache_key;  // same for the lattice table of fava_prolong (WS_TABLE)
    std::map<int64_t, void*> twiddles;  // twiddle tables of the hand-written FFT, per N (fft.cu)
    std::map<std::string, CUtensorMap> tensor_maps;  // TMA descriptors, keyed by (base, dtype, dims, strides, box)
    void* tile_counters = nullptr;  // 64 tile counters of the persistent column kernels, used in turn (fft.cu)
    unsigned tile_counter_next = 0;
    fava::Staging* staging = nullptr;
};

namespace fava {

// Grow-only device workspace owned by the context.
int ctx_workspace(fava_ctx* ctx, int slot, size_t bytes, void** out);

// Cached tiled TMA descriptor (cuTensorMapEncodeTiled through the runtime's driver entry point; no swizzle, no
// interleave, 128-byte L2 promotion).  dims / box in elements, strides in bytes (rank - 1 of them), innermost first.
// nan_fill: elements of a box that lie outside the tensor read as NaN instead of 0.
int ctx_tensor_map(fava_ctx* ctx, const void* base, CUtensorMapDataType dtype, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, CUtensorMap* out, bool nan_fill = false);

// hand-written FFT path for this grid size?  (power of two in [256, 2048]; other even N use cuFFT)
bool fft_native_supported(int64_t n);

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// ---- device helpers ---------------------------------------------------------------------------

// Streaming (evict-first) vector loads widened to fp64 in registers.  f32 -> f64 widening is exact,
// i.e. bit-identical to numpy's astype(float64) in the reference loader (_flash.py:333).
template <typename T, int V>
struct VecLoad;

template <>
struct VecLoad<double, 1> {
    static __device__ __forceinline__ void ld(const double* p, double (&v)[1]) { v[0] = __ldcs(p); }
};
template <>
struct VecLoad<double, 2> {
    static __device__ __forceinline__ void ld(const double* p, double (&v)[2]) {
        double2 t = __ldcs(reinterpret_cast<const double2*>(p));
        v[0] = t.x;
        v[1] = t.y;
    }
};
template <>
struct VecLoad<float, 1> {
    static __device__ __forceinline__ void ld(const float* p, double (&v)[1]) { v[0] = (double)__ldcs(p); }
};
template <>
struct VecLoad<float, 2> {
    static __device__ __forceinline__ void ld(const float* p, double (&v)[2]) {
        float2 t = __ldcs(reinterpret_cast<const float2*>(p));
        v[0] = (double)t.x;
        v[1] = (double)t.y;
    }
};
template <>
struct VecLoad<float, 4> {
    static __device__ __forceinline__ void ld(const float* p, double (&v)[4]) {
        float4 t = __ldcs(reinterpret_cast<const float4*>(p));
        v[0] = (double)t.x;
        v[1] = (double)t.y;
        v[2] = (double)t.z;
        v[3] = (double)t.w;
    }
};

__device__ __forceinline__ double warp_sum_fixed(double v) {
    // xor-butterfly: every lane ends with the same value, summation order fixed by lane ids
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- mbarrier / bulk-copy (TMA, cp.async.bulk) wrappers shared by the exchange and FFT kernels ---------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}

}  // namespace fava
