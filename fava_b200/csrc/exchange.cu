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
#include "common.cuh"

namespace fava {

__global__ void __launch_bounds__(128)
    k_a2a_pack(const double2* __restrict__ in, double2* const* __restrict__ peer_recv,
               const int32_t* __restrict__ ky_of_dest, int me, int nranks, int nz_local, int n, int nyl, int nxh) {
    // blockIdx.x = ((zl * nranks) + dest) * nyl + jl
    const int64_t id = blockIdx.x + (int64_t)blockIdx.y * gridDim.x;
    const int64_t total = (int64_t)nz_local * nranks * nyl;
    if (id >= total) return;
    const int jl = (int)(id % nyl);
    const int dest = (int)((id / nyl) % nranks);
    const int zl = (int)(id / ((int64_t)nyl * nranks));
    const int j = ky_of_dest[dest * nyl + jl];
    if (j < 0) return;
    const double2* src = in + ((int64_t)zl * n + j) * nxh;
    double2* dst = peer_recv[dest] + (((int64_t)me * nz_local + zl) * nyl + jl) * nxh;
    for (int x = threadIdx.x; x < nxh; x += blockDim.x) dst[x] = __ldcs(src + x);
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
    const unsigned gx = (unsigned)std::min<int64_t>(total, 65535 * 16);
    const unsigned gy = (unsigned)((total + gx - 1) / gx);
    k_a2a_pack<<<dim3(gx, gy), 128, 0, (cudaStream_t)stream>>>((const double2*)d_in, (double2* const*)d_peer_recv,
                                                              d_ky_of_dest, my_rank, nranks, (int)nz_local, (int)n,
                                                              (int)nyl, (int)(n / 2 + 1));
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
