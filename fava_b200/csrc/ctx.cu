// Context lifetime, error reporting and small utilities of libfava_b200.
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace fava {

thread_local std::string g_last_error;
std::atomic<int64_t> g_launches{0};

int set_error(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

int ctx_workspace(fava_ctx* ctx, int slot, size_t bytes, void** out) {
    if (slot < 0 || slot >= WS_COUNT) return set_error(FAVA_EINVAL, "bad workspace slot %d", slot);
    if (ctx->ws_bytes[slot] < bytes) {
        if (ctx->ws[slot]) {
            // outstanding work may still read the old buffer
            FAVA_CHECK_CUDA(cudaDeviceSynchronize());
            FAVA_CHECK_CUDA(cudaFree(ctx->ws[slot]));
            ctx->ws[slot] = nullptr;
            ctx->ws_bytes[slot] = 0;
        }
        size_t want = (bytes + 255) & ~size_t(255);
        cudaError_t e = cudaMalloc(&ctx->ws[slot], want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return set_error(FAVA_ENOMEM, "workspace slot %d: cudaMalloc(%zu) failed: %s", slot, want,
                             cudaGetErrorString(e));
        }
        ctx->ws_bytes[slot] = want;
    }
    *out = ctx->ws[slot];
    return FAVA_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int ctx_tensor_map(fava_ctx* ctx, const void* base, CUtensorMapDataType dtype, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, CUtensorMap* out, bool nan_fill) {
    std::string key((const char*)&base, sizeof(base));
    key.push_back(nan_fill ? 'n' : 'z');
    key.append((const char*)&dtype, sizeof(dtype));
    key.append((const char*)dims, sizeof(uint64_t) * rank);
    key.append((const char*)strides_bytes, sizeof(uint64_t) * (rank - 1));
    key.append((const char*)box, sizeof(uint32_t) * rank);
    auto it = ctx->tensor_maps.find(key);
    if (it != ctx->tensor_maps.end()) {
        *out = it->second;
        return FAVA_OK;
    }
    static EncodeTiledFn enc = nullptr;
    if (!enc) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        FAVA_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (!fn || q != cudaDriverEntryPointSuccess)
            return set_error(FAVA_ECUDA, "cuTensorMapEncodeTiled is not available in this driver");
        enc = (EncodeTiledFn)fn;
    }
    cuuint64_t d[5], sb[4];
    cuuint32_t b[5], es[5];
    for (int i = 0; i < rank; ++i) d[i] = dims[i], b[i] = box[i], es[i] = 1;
    for (int i = 0; i + 1 < rank; ++i) sb[i] = strides_bytes[i];
    CUtensorMap m;
    const CUresult r = enc(&m, dtype, (cuuint32_t)rank, const_cast<void*>(base), d, sb, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           nan_fill ? CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA : CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error(FAVA_ECUDA, "cuTensorMapEncodeTiled(rank %d, inner %llu, box %u) failed: CUresult %d", rank,
                         (unsigned long long)dims[0], box[0], (int)r);
    if (ctx->tensor_maps.size() > 256) ctx->tensor_maps.clear();  // buffers are few and long-lived; bound the cache anyway
    ctx->tensor_maps[key] = m;
    *out = m;
    return FAVA_OK;
}

void staging_destroy(Staging* s);  // staging.cu

}  // namespace fava

using namespace fava;

extern "C" {

int fava_abi_version(void) { return FAVA_ABI_VERSION; }

const char* fava_last_error(void) { return g_last_error.c_str(); }

int64_t fava_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int fava_init(int device, fava_ctx** out) {
    if (!out) return set_error(FAVA_EINVAL, "fava_init: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return set_error(FAVA_ENODEV, "fava_init: no CUDA device (%s)",
                         e == cudaSuccess ? "count is 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= ndev)
        return set_error(FAVA_EINVAL, "fava_init: device %d out of range [0,%d)", device, ndev);
    cudaDeviceProp prop;
    FAVA_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return set_error(FAVA_ENODEV,
                         "fava_init: device %d is sm_%d%d; this library ships sm_100a code only",
                         device, prop.major, prop.minor);
    FAVA_CHECK_CUDA(cudaSetDevice(device));
    fava_ctx* ctx = new fava_ctx();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    *out = ctx;
    return FAVA_OK;
}

int fava_shutdown(fava_ctx* ctx) {
    if (!ctx) return FAVA_OK;
    DeviceGuard g(ctx->device);
    cudaDeviceSynchronize();
    if (ctx->staging) staging_destroy(ctx->staging);
    for (auto& kv : ctx->plans) cufftDestroy(kv.second);
    for (auto& kv : ctx->twiddles) cudaFree(kv.second);
    if (ctx->tile_counters) cudaFree(ctx->tile_counters);
    for (int i = 0; i < WS_COUNT; ++i)
        if (ctx->ws[i]) cudaFree(ctx->ws[i]);
    delete ctx;
    return FAVA_OK;
}

int fava_stream_sync(fava_ctx* ctx, void* stream) {
    if (!ctx) return set_error(FAVA_EINVAL, "fava_stream_sync: ctx is NULL");
    DeviceGuard g(ctx->device);
    FAVA_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return FAVA_OK;
}

int fava_ipc_export(void* d_ptr, unsigned char handle_out[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    FAVA_CHECK_CUDA(cudaIpcGetMemHandle(&h, d_ptr));
    memcpy(handle_out, &h, 64);
    return FAVA_OK;
}

int fava_ipc_open(const unsigned char handle[64], void** d_ptr_out) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    FAVA_CHECK_CUDA(cudaIpcOpenMemHandle(d_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return FAVA_OK;
}

int fava_ipc_close(void* d_ptr) {
    FAVA_CHECK_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return FAVA_OK;
}

}  // extern "C"
