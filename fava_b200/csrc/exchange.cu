// K5 — slab -> ky-pencil exchange of the distributed FFT, fused with the pack.
//
// The reference has no counterpart (its kinetic_energy_spectra is single-address-space NumPy,
// fava/mesh/FLASH/FlashUniform.py:229-304).  After the local 2-D transforms rank `me` holds complex
// [nz_local][n (ky)][pitch]; the z-transform needs every z of a (ky,kx) column on one rank.  Instead of
// pack -> all-to-all -> unpack (three passes over the data plus a staging buffer), ONE kernel reads each
// ky row once and stores it straight into the owning rank's receive buffer - a peer-mapped pointer
// (CUDA IPC over NVLink 5 / NVSwitch) or local memory when the owner is this rank - already in the
// [z][ky_local][kx] order the strided z-transform and the binning kernel consume.
// Ownership is +-ky symmetric (fava_b200/spectrum.py:ky_ownership), padding rows (-1) are skipped.
#include <cstdlib>

#include "common.cuh"

namespace fava {

// ---- bulk asynchronous copies (TMA), one elected thread per CTA -------------------------------------------
// global -> shared (cp.async.bulk, completion on an mbarrier) -> peer global (cp.async.bulk store over NVLink).
// No registers or LSU slots are spent on the payload, so a handful of 32-thread CTAs keeps megabytes in
// flight while the HBM-bound kernels of the other stream own the SMs.  Ring of kSlots row buffers: row i is
// loaded kAhead iterations before it is stored; a slot is refilled once the store issued kSlots-kAhead
// iterations earlier has finished READING it (bulk-group completion is in order).
constexpr int kPackCtas = 64;
constexpr int kTmaSlots = 12;
constexpr int kTmaAhead = 8;  // loads in flight; kTmaSlots - kTmaAhead stores may still be reading their slot

struct PackRow {
    const double2* src;
    double2* dst;
    unsigned bytes;
};

__device__ __forceinline__ bool pack_row(int64_t it, const double2* in, double2* const* peer_recv,
                                         const int32_t* ky_of_dest, int me, int nranks, int nz_local, int n, int nyl,
                                         int pitch, int kmax2, PackRow* r) {
    const int jl = (int)(it % nyl);
    const int dest = (int)(((it / nyl) % nranks + me) % nranks);
    const int zl = (int)(it / ((int64_t)nyl * nranks));
    const int j = ky_of_dest[dest * nyl + jl];
    if (j < 0) return false;
    const int ky = j < n / 2 ? j : j - n;
    const int rem = kmax2 - ky * ky;
    if (rem < 0) return false;
    // columns with kx^2 + ky^2 beyond the last shell can never reach a bin, whatever kz: not sent (21 % of the
    // NVLink bytes); whole 8-column groups are sent so that the z pass reads what the y pass wrote
    const int nsend = min(pitch, (((int)sqrt((double)rem) + 1) | 7) + 1);
    r->src = in + ((int64_t)zl * n + j) * pitch;
    r->dst = peer_recv[dest] + (((int64_t)me * nz_local + zl) * nyl + jl) * pitch;
    r->bytes = (unsigned)nsend * 16u;
    return true;
}

__global__ void __launch_bounds__(32)
    k_a2a_pack_tma(const double2* __restrict__ in, double2* const* __restrict__ peer_recv,
                   const int32_t* __restrict__ ky_of_dest, int me, int nranks, int nz_local, int n, int nyl, int pitch,
                   int kmax2) {
    extern __shared__ __align__(128) unsigned char ring[];  // kTmaSlots x slot_bytes
    __shared__ __align__(8) uint64_t bars[kTmaSlots];
    if (threadIdx.x != 0) return;
    const unsigned slot_bytes = ((unsigned)pitch * 16u + 127u) & ~127u;
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
            const bool ok = pack_row(it_load, in, peer_recv, ky_of_dest, me, nranks, nz_local, n, nyl, pitch, kmax2, &r);
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
    const int kmax2 = (int)(n * n / 4 - 3 * n / 2 + 2);
    const int pitch = (int)fava_spectral_pitch(n);
    const size_t slot_bytes = ((size_t)pitch * 16 + 127) & ~size_t(127);
    const size_t ring_bytes = slot_bytes * kTmaSlots;
    FAVA_REQUIRE(ring_bytes <= 200 * 1024, "fava_a2a_pack: grid size %lld too large for the row ring", (long long)n);
    // 64 single-warp CTAs keep megabytes in flight; NVLink (~0.7 TB/s) needs far fewer SMs than HBM does
    const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>(total, kPackCtas));
    FAVA_CHECK_CUDA(cudaFuncSetAttribute(k_a2a_pack_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_bytes));
    k_a2a_pack_tma<<<gx, 32, ring_bytes, (cudaStream_t)stream>>>((const double2*)d_in, (double2* const*)d_peer_recv,
                                                                 d_ky_of_dest, my_rank, nranks, (int)nz_local, (int)n,
                                                                 (int)nyl, pitch, kmax2);
    FAVA_LAUNCHED();
    return FAVA_OK;
}

int fava_workspace(fava_ctx* ctx, int slot, int64_t bytes, void** d_ptr_out) {
    FAVA_REQUIRE(ctx && d_ptr_out && bytes >= 0, "fava_workspace: bad argument");
    DeviceGuard g(ctx->device);
    const size_t before = (slot >= 0 && slot < WS_COUNT) ? ctx->ws_bytes[slot] : 0;
    int rc = ctx_workspace(ctx, slot, (size_t)bytes, d_ptr_out);
    if (rc) return rc;
    if (ctx->ws_bytes[slot] != before)  // fresh allocation: start from finite values (synchronous: no stream may use it yet)
        FAVA_CHECK_CUDA(cudaMemset(*d_ptr_out, 0, ctx->ws_bytes[slot]));
    return FAVA_OK;
}

}  // extern "C"
