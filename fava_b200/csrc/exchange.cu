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
constexpr int kPackThreads = 512;
constexpr int kPackWarps = kPackThreads / 32;
constexpr int kPackUnroll = 16;  // 16 x 16 B per lane = one 8 KB row (N = 1024) per warp in flight

__global__ void __launch_bounds__(kPackThreads)
    k_a2a_pack(const double2* __restrict__ in, double2* const* __restrict__ peer_recv,
               const int32_t* __restrict__ ky_of_dest, int me, int nranks, int nz_local, int n, int nyl, int nxh,
               int kmax2) {
    const int lane = threadIdx.x & 31;
    const int64_t total = (int64_t)nz_local * nranks * nyl;
    const int64_t stride = (int64_t)gridDim.x * kPackWarps;
    // one warp per ky row; destination-major order staggered by rank: at any time the ranks write to different peers
    for (int64_t it = (int64_t)blockIdx.x * kPackWarps + (threadIdx.x >> 5); it < total; it += stride) {
        const int jl = (int)(it % nyl);
        const int zl = (int)((it / nyl) % nz_local);
        const int dest = (int)((it / ((int64_t)nyl * nz_local) + me) % nranks);
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
    static const int env_ctas = [] {
        const char* e = getenv("FAVA_A2A_CTAS");
        return e ? atoi(e) : 0;
    }();
    const int want = env_ctas > 0 ? env_ctas : 64;
    const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>(total, want));
    k_a2a_pack<<<gx, kPackThreads, 0, (cudaStream_t)stream>>>((const double2*)d_in, (double2* const*)d_peer_recv,
                                                              d_ky_of_dest, my_rank, nranks, (int)nz_local, (int)n,
                                                              (int)nyl, (int)(n / 2 + 1), (int)(n * n / 4 - 3 * n / 2 + 2));
    FAVA_LAUNCHED();
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
