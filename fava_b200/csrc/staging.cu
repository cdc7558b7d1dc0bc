// HDF5 block staging: file (or host array) -> pinned ring -> async H2D.
//
// Replaces the reference's field loader FLASH._read_variable_data (fava/mesh/FLASH/_flash.py:306-341):
// h5py read -> astype(float64) -> swapaxes -> ascontiguousarray -> copy into an MPI shared window
// (3-4 full copies, one of them a strided transpose, root rank only).  Here the dataset's raw bytes
// (contiguous layout; offset from the h5lite index) go straight to HBM in FILE order and FILE dtype:
// a pool of reader threads `pread`s slices of a chunk into a pinned ring buffer and every chunk is
// shipped with one cudaMemcpyAsync on the caller's stream, so disk/page-cache reads of chunk i+1
// overlap the PCIe copy of chunk i.  f32 -> f64 widening and the axis permutation happen inside the
// consuming kernels (registers / index arithmetic), never as a pass over memory.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace fava {

constexpr size_t kChunkBytes = size_t(16) << 20;  // one H2D copy
constexpr int kRingDepth = 4;
constexpr size_t kSliceBytes = size_t(2) << 20;  // one pread / memcpy task
constexpr int kMaxWorkers = 8;

class WorkerPool {
  public:
    explicit WorkerPool(int n) {
        for (int i = 0; i < n; ++i) threads_.emplace_back([this] { loop(); });
    }
    ~WorkerPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : threads_) t.join();
    }
    // Run fn(i) for i in [0,n) on the pool and wait.  Returns the first non-zero result.
    int parallel_for(int n, const std::function<int(int)>& fn) {
        if (n <= 0) return 0;
        std::unique_lock<std::mutex> lk(mu_);
        fn_ = &fn;
        next_ = 0;
        total_ = n;
        pending_ = n;
        result_ = 0;
        cv_.notify_all();
        done_cv_.wait(lk, [this] { return pending_ == 0; });
        fn_ = nullptr;
        return result_;
    }

  private:
    void loop() {
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            cv_.wait(lk, [this] { return stop_ || (fn_ && next_ < total_); });
            if (stop_) return;
            const int i = next_++;
            const std::function<int(int)>* fn = fn_;
            lk.unlock();
            const int r = (*fn)(i);
            lk.lock();
            if (r && !result_) result_ = r;
            if (--pending_ == 0) done_cv_.notify_all();
        }
    }
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    const std::function<int(int)>* fn_ = nullptr;
    int next_ = 0, total_ = 0, pending_ = 0, result_ = 0;
    bool stop_ = false;
};

struct Staging {
    void* ring[kRingDepth] = {};
    cudaEvent_t freed[kRingDepth] = {};
    bool in_flight[kRingDepth] = {};
    int head = 0;
    WorkerPool* pool = nullptr;
    std::mutex mu;  // one staging call at a time per context
};

static int staging_get(fava_ctx* ctx, Staging** out) {
    if (!ctx->staging) {
        Staging* s = new Staging();
        for (int i = 0; i < kRingDepth; ++i) {
            cudaError_t e = cudaHostAlloc(&s->ring[i], kChunkBytes, cudaHostAllocDefault);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->freed[i], cudaEventDisableTiming);
            if (e != cudaSuccess) {
                cudaGetLastError();
                for (int j = 0; j <= i; ++j) {
                    if (s->ring[j]) cudaFreeHost(s->ring[j]);
                    if (s->freed[j]) cudaEventDestroy(s->freed[j]);
                }
                delete s;
                return set_error(FAVA_ENOMEM, "staging: pinned ring allocation failed: %s", cudaGetErrorString(e));
            }
        }
        const unsigned hw = std::thread::hardware_concurrency();
        s->pool = new WorkerPool(std::max(1, std::min<int>(kMaxWorkers, hw ? (int)hw : 4)));
        ctx->staging = s;
    }
    *out = ctx->staging;
    return FAVA_OK;
}

void staging_destroy(Staging* s) {
    if (!s) return;
    delete s->pool;
    for (int i = 0; i < kRingDepth; ++i) {
        if (s->freed[i]) cudaEventDestroy(s->freed[i]);
        if (s->ring[i]) cudaFreeHost(s->ring[i]);
    }
    delete s;
}

// Ship `nbytes` produced by fill(dst, offset, len) -> 0/-errno through the ring to d_dst.
static int stage_through_ring(Staging* s, size_t nbytes, void* d_dst, cudaStream_t st,
                              const std::function<int(char*, size_t, size_t)>& fill) {
    std::lock_guard<std::mutex> lk(s->mu);
    size_t done = 0;
    while (done < nbytes) {
        const size_t len = std::min(kChunkBytes, nbytes - done);
        const int slot = s->head;
        s->head = (s->head + 1) % kRingDepth;
        if (s->in_flight[slot]) {
            FAVA_CHECK_CUDA(cudaEventSynchronize(s->freed[slot]));
            s->in_flight[slot] = false;
        }
        char* buf = (char*)s->ring[slot];
        const int nslices = (int)((len + kSliceBytes - 1) / kSliceBytes);
        const size_t base = done;
        const int rc = s->pool->parallel_for(nslices, [&](int i) -> int {
            const size_t o = (size_t)i * kSliceBytes;
            return fill(buf + o, base + o, std::min(kSliceBytes, len - o));
        });
        if (rc) return rc;
        FAVA_CHECK_CUDA(cudaMemcpyAsync((char*)d_dst + done, buf, len, cudaMemcpyHostToDevice, st));
        FAVA_CHECK_CUDA(cudaEventRecord(s->freed[slot], st));
        s->in_flight[slot] = true;
        done += len;
    }
    return FAVA_OK;
}

}  // namespace fava

using namespace fava;

extern "C" {

int fava_stage_h2d(fava_ctx* ctx, const char* path, int64_t file_offset, int64_t nbytes, void* d_dst,
                   void* stream) {
    FAVA_REQUIRE(ctx && path && (d_dst || nbytes == 0), "fava_stage_h2d: NULL argument");
    FAVA_REQUIRE(file_offset >= 0 && nbytes >= 0, "fava_stage_h2d: negative offset/size");
    if (nbytes == 0) return FAVA_OK;
    DeviceGuard g(ctx->device);
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return set_error(FAVA_EIO, "fava_stage_h2d: cannot open %s: %s", path, strerror(errno));
    struct stat sb;
    if (fstat(fd, &sb) != 0 || (int64_t)sb.st_size < file_offset + nbytes) {
        close(fd);
        return set_error(FAVA_EIO, "fava_stage_h2d: %s is shorter than offset %lld + %lld bytes", path,
                         (long long)file_offset, (long long)nbytes);
    }
    Staging* s;
    int rc = staging_get(ctx, &s);
    if (rc == FAVA_OK) {
        rc = stage_through_ring(s, (size_t)nbytes, d_dst, (cudaStream_t)stream,
                                [&](char* dst, size_t off, size_t len) -> int {
                                    size_t got = 0;
                                    while (got < len) {
                                        const ssize_t r = pread(fd, dst + got, len - got, file_offset + off + got);
                                        if (r <= 0) return FAVA_EIO;
                                        got += (size_t)r;
                                    }
                                    return 0;
                                });
        if (rc == FAVA_EIO) set_error(FAVA_EIO, "fava_stage_h2d: short read from %s", path);
    }
    close(fd);
    return rc;
}

int fava_stage_host_h2d(fava_ctx* ctx, const void* h_src, int64_t nbytes, void* d_dst, void* stream) {
    FAVA_REQUIRE(ctx && ((h_src && d_dst) || nbytes == 0), "fava_stage_host_h2d: NULL argument");
    FAVA_REQUIRE(nbytes >= 0, "fava_stage_host_h2d: negative size");
    if (nbytes == 0) return FAVA_OK;
    DeviceGuard g(ctx->device);
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, h_src);
    if (e != cudaSuccess) cudaGetLastError();
    if (e == cudaSuccess && attr.type == cudaMemoryTypeHost) {  // already pinned: one direct async copy
        FAVA_CHECK_CUDA(cudaMemcpyAsync(d_dst, h_src, (size_t)nbytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
        return FAVA_OK;
    }
    Staging* s;
    int rc = staging_get(ctx, &s);
    if (rc) return rc;
    const char* src = (const char*)h_src;
    return stage_through_ring(s, (size_t)nbytes, d_dst, (cudaStream_t)stream,
                              [&](char* dst, size_t off, size_t len) -> int {
                                  memcpy(dst, src + off, len);
                                  return 0;
                              });
}

}  // extern "C"
