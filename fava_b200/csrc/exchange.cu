// K5 — slab -> ky-pencil exchange of the distributed FFT, fused with the pack.
//
// The reference has no counterpart (its kinetic_energy_spectra is single-address-space NumPy,
// fava/mesh/FLASH/FlashUniform.py:229-304).  After the local 2-D transforms rank `me` holds complex
// [nz_local][n (ky)][nxh]; the z-transform needs every z of a (ky,kx) column on one rank.  Instead of
// pack -> all-to-all -> unpack (three passes over the data plus a staging buffer), ONE kernel reads each
// ky row once and stores it straight into the owning rank's receive buffer — a peer-mapped pointer
// (CUDA IPC over NVLink 5 / NVSwitch) or local memory when the owner is this rank — already in the
// [z][ky_local][kx] order the strided z-transform and the binning kernel consume.  16-byte coalesced
// loads and stores; the NVLink stores of one row overlap the loads of the next.
// Ownership is +-ky symmetric (fava_b200/spectrum.py:ky_ownership), padding rows (-1) are skipped.
#include <cstdlib>

#include "common.cuh"

namespace fava {

// Persistent form: a SMALL grid (default 48 CTAs, FAVA_A2A_CTAS) walks all rows, so the kernel occupies only a
// fraction of the SMs and — launched on a high-priority side stream — runs BESIDE the HBM-bound FFT / moment
// kernels instead of queueing behind them; NVLink (~0.6 TB/s) needs far fewer SMs than HBM does.
constexpr int kPackThreads = 512;  // upper bound; the launch may use fewer (FAVA_A2A_THREADS)
constexpr int kPackUnroll = 16;  // 16 x 16 B per lane = one 8 KB row (N = 1024) per warp in flight

__global__ void __launch_bounds__(kPackThreads)
    k_a2a_pack(const double2* __restrict__ in, double2* const* __restrict__ peer_recv,
               const int32_t* __restrict__ ky_of_dest, int me, int nranks, int nz_local, int n, int nyl, int nxh,
               int kmax2) {
    const int lane = threadIdx.x & 31;
    const int64_t total = (int64_t)nz_local * nranks * nyl;
    const int nwarps = blockDim.x >> 5;
    const int64_t stride = (int64_t)gridDim.x * nwarps;
    // one warp per ky row; consecutive row groups go to different peers (staggered by rank), so every NVLink
    // port of the switch is busy at any time
    for (int64_t it = (int64_t)blockIdx.x * nwarps + (threadIdx.x >> 5); it < total; it += stride) {
        const int jl = (int)(it % nyl);
        const int dest = (int)(((it / nyl) % nranks + me) % nranks);
        const int zl = (int)(it / ((int64_t)nyl * nranks));
        const int j = ky_of_dest[dest * nyl + jl];
        if (j < 0) continue;
        // columns with kx^2 + ky^2 beyond the last shell can never reach a bin, whatever kz: not sent
        // (21 % of the NVLink bytes; the receive buffers are zero there from their allocation)
        const int ky = j < n / 2 ? j : j - n;
        const int rem = kmax2 - ky * ky;
        if (rem < 0) continue;
        const int nsend = min(nxh, (int)sqrt((double)rem) + 2);
        const double2* src = in + ((int64_t)zl * n + j) * nxh;
        double2* dst = peer_recv[dest] + (((int64_t)me * nz_local + zl) * nyl + jl) * nxh;
        int x = lane;
        for (; x + (kPackUnroll - 1) * 32 < nsend; x += kPackUnroll * 32) {
            double2 v[kPackUnroll];
#pragma unroll
            for (int u = 0; u < kPackUnroll; ++u) v[u] = __ldcs(src + x + u * 32);
#pragma unroll
            for (int u = 0; u < kPackUnroll; ++u) dst[x + u * 32] = v[u];
        }
        for (; x < nsend; x += 32) dst[x] = __ldcs(src + x);
    }
}


// ---- TMA form: bulk asynchronous copies, one elected thread per CTA ---------------------------------------
// global -> shared (cp.async.bulk, completion on an mbarrier) -> peer global (cp.async.bulk store over NVLink).
// No registers or LSU slots are spent on the payload, so a handful of 32-thread CTAs keeps megabytes in
// flight while the HBM-bound kernels of the other stream own the SMs.  Ring of kSlots row buffers: row i is
// loaded kAhead iterations before it is stored; a slot is refilled once the store issued kSlots-kAhead
// iterations earlier has finished READING it (bulk-group completion is in order).
constexpr int kTmaSlots = 12;
constexpr int kTmaAhead = 8;  // loads in flight; kTmaSlots - kTmaAhead stores may still be reading their slot

struct PackRow {
    const double2* src;
    double2* dst;
    unsigned bytes;
};

__device__ __forceinline__ bool pack_row(int64_t it, const double2* in, double2* const* peer_recv,
                                         const int32_t* ky_of_dest, int me, int nranks, int nz_local, int n, int nyl,
                                         int nxh, int kmax2, PackRow* r) {
    const int jl = (int)(it % nyl);
    const int dest = (int)(((it / nyl) % nranks + me) % nranks);
    const int zl = (int)(it / ((int64_t)nyl * nranks));
    const int j = ky_of_dest[dest * nyl + jl];
    if (j < 0) return false;
    const int ky = j < n / 2 ? j : j - n;
    const int rem = kmax2 - ky * ky;
    if (rem < 0) return false;
    const int nsend = min(nxh, (int)sqrt((double)rem) + 2);
    r->src = in + ((int64_t)zl * n + j) * nxh;
    r->dst = peer_recv[dest] + (((int64_t)me * nz_local + zl) * nyl + jl) * nxh;
    r->bytes = (unsigned)nsend * 16u;
    return true;
}

__global__ void __launch_bounds__(32)
    k_a2a_pack_tma(const double2* __restrict__ in, double2* const* __restrict__ peer_recv,
                   const int32_t* __restrict__ ky_of_dest, int me, int nranks, int nz_local, int n, int nyl, int nxh,
                   int kmax2) {
    extern __shared__ __align__(128) unsigned char ring[];  // kTmaSlots x slot_bytes
    __shared__ __align__(8) uint64_t bars[kTmaSlots];
    if (threadIdx.x != 0) return;
    const unsigned slot_bytes = ((unsigned)nxh * 16u + 127u) & ~127u;
    for (int s = 0; s < kTmaSlots; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async;\n" ::: "memory");

    const int64_t total = (int64_t)nz_local * nranks * nyl;
    int64_t it_load = blockIdx.x;  // next work item to consider for loading
    // rows are numbered in issue order; row q lives in slot q % kTmaSlots and uses that slot's barrier with
    // parity (q / kTmaSlots) & 1
    int64_t q_load = 0, q_store = 0;
    PackRow rows[kTmaSlots];

    auto issue_load = [&]() -> bool {
        PackRow r;
        while (it_load < total) {
            const bool ok = pack_row(it_load, in, peer_recv, ky_of_dest, me, nranks, nz_local, n, nyl, nxh, kmax2, &r);
            it_load += gridDim.x;
            if (ok) {
                const int s = (int)(q_load % kTmaSlots);
                rows[s] = r;
                mbar_expect_tx(&bars[s], r.bytes);
                bulk_load(ring + (size_t)s * slot_bytes, r.src, r.bytes, &bars[s]);
                ++q_load;
                return true;
            }
        }
        return false;
    };

    for (int k = 0; k < kTmaAhead; ++k)
        if (!issue_load()) break;
    while (q_store < q_load) {
        const int s = (int)(q_store % kTmaSlots);
        mbar_wait(&bars[s], (unsigned)((q_store / kTmaSlots) & 1));
        bulk_store(rows[s].dst, ring + (size_t)s * slot_bytes, rows[s].bytes);
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        ++q_store;
        // the slot about to be refilled (row q_load) was last stored as row q_load - kTmaSlots: at most
        // kTmaSlots - kTmaAhead - 1 newer store groups may remain pending
        asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(kTmaSlots - kTmaAhead - 1) : "memory");
        issue_load();
    }
    asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
    __threadfence_system();
}

}  // namespace fava

using namespace fava;

extern "C" {

int fava_a2a_pack(fava_ctx* ctx, const double* d_in, double* const* d_peer_recv, const int32_t* d_ky_of_dest,
                  int my_rank, int nranks, int64_t nz_local, int64_t n, int64_t nyl, void* stream) {
    FAVA_REQUIRE(ctx && d_in && d_peer_recv && d_ky_of_dest, "fava_a2a_pack: NULL argument");
    FAVA_REQUIRE(nranks > 0 && my_rank >= 0 && my_rank < nranks, "fava_a2a_pack: bad rank %d of %d", my_rank, nranks);
    FAVA_REQUIRE(nz_local > 0 && n > 1 && (n & 1) == 0 && nyl > 0, "fava_a2a_pack: bad shape");
    DeviceGuard g(ctx->device);
    const int64_t total = nz_local * nranks * nyl;
    const char* e_ctas = getenv("FAVA_A2A_CTAS");  // tuning knobs, read per call
    const int env_ctas = e_ctas ? atoi(e_ctas) : 0;
    const char* e_thr = getenv("FAVA_A2A_THREADS");
    const int env_thr = e_thr ? std::max(32, std::min(kPackThreads, atoi(e_thr) / 32 * 32)) : 0;
    const char* e_mode = getenv("FAVA_A2A_MODE");  // "ldst" selects the load/store kernel; default: bulk-copy (TMA) kernel
    const int env_mode = (e_mode && e_mode[0] == 'l') ? 1 : 0;
    const int kmax2 = (int)(n * n / 4 - 3 * n / 2 + 2);
    const size_t slot_bytes = ((size_t)(n / 2 + 1) * 16 + 127) & ~size_t(127);
    const size_t ring_bytes = slot_bytes * kTmaSlots;
    if (env_mode == 0 && ring_bytes <= 200 * 1024) {
        const int want = env_ctas > 0 ? env_ctas : 64;
        const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>(total, want));
        FAVA_CHECK_CUDA(cudaFuncSetAttribute(k_a2a_pack_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_bytes));
        k_a2a_pack_tma<<<gx, 32, ring_bytes, (cudaStream_t)stream>>>((const double2*)d_in, (double2* const*)d_peer_recv,
                                                                     d_ky_of_dest, my_rank, nranks, (int)nz_local, (int)n,
                                                                     (int)nyl, (int)(n / 2 + 1), kmax2);
        FAVA_LAUNCHED();
        return FAVA_OK;
    }
    const int want = env_ctas > 0 ? env_ctas : 2 * ctx->num_sms;
    const int threads = env_thr > 0 ? env_thr : 128;
    const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>(total, want));
    k_a2a_pack<<<gx, threads, 0, (cudaStream_t)stream>>>((const double2*)d_in, (double2* const*)d_peer_recv,
                                                              d_ky_of_dest, my_rank, nranks, (int)nz_local, (int)n,
                                                              (int)nyl, (int)(n / 2 + 1), (int)(n * n / 4 - 3 * n / 2 + 2));
    FAVA_LAUNCHED();
    return FAVA_OK;
}

int fava_a2a_copy(fava_ctx* ctx, const double* d_in, double* const* h_peer_recv, const int32_t* h_ky_of_dest,
                  int my_rank, int nranks, int64_t nz_local, int64_t n, int64_t nyl, void* stream) {
    FAVA_REQUIRE(ctx && d_in && h_peer_recv && h_ky_of_dest, "fava_a2a_copy: NULL argument");
    FAVA_REQUIRE(nranks > 0 && my_rank >= 0 && my_rank < nranks, "fava_a2a_copy: bad rank %d of %d", my_rank, nranks);
    FAVA_REQUIRE(nz_local > 0 && n > 1 && (n & 1) == 0 && nyl > 0, "fava_a2a_copy: bad shape");
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nxh = n / 2 + 1;
    const size_t row = (size_t)nxh * 16;
    for (int k = 1; k <= nranks; ++k) {  // staggered: at any time the ranks target different peers
        const int dest = (my_rank + k) % nranks;
        const int32_t* own = h_ky_of_dest + (int64_t)dest * nyl;
        int64_t jl = 0;
        while (jl < nyl) {
            if (own[jl] < 0) {
                ++jl;
                continue;
            }
            int64_t len = 1;  // run of consecutive ky rows: contiguous in the source AND in the destination
            while (jl + len < nyl && own[jl + len] == own[jl] + len) ++len;
            const char* src = (const char*)d_in + (size_t)own[jl] * row;
            char* dst = (char*)h_peer_recv[dest] + ((size_t)my_rank * nz_local * nyl + jl) * row;
            FAVA_CHECK_CUDA(cudaMemcpy2DAsync(dst, (size_t)nyl * row, src, (size_t)n * row, (size_t)len * row,
                                              (size_t)nz_local, cudaMemcpyDefault, st));
            jl += len;
        }
    }
    return FAVA_OK;
}

int fava_workspace(fava_ctx* ctx, int slot, int64_t bytes, void** d_ptr_out) {
    FAVA_REQUIRE(ctx && d_ptr_out && bytes >= 0, "fava_workspace: bad argument");
    DeviceGuard g(ctx->device);
    const size_t before = (slot >= 0 && slot < WS_COUNT) ? ctx->ws_bytes[slot] : 0;
    int rc = ctx_workspace(ctx, slot, (size_t)bytes, d_ptr_out);
    if (rc) return rc;
    if (ctx->ws_bytes[slot] != before)  // fresh allocation: padding rows of exchange buffers must be finite
        FAVA_CHECK_CUDA(cudaMemset(*d_ptr_out, 0, ctx->ws_bytes[slot]));
    return FAVA_OK;
}

}  // extern "C"
