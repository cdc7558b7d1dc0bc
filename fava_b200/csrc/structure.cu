// Velocity structure functions of a uniform grid (reference FlashUniform.structure_functions,
// fava/mesh/FLASH/FlashUniform.py:306-445), the GPU part.
//
// The reference draws its point pairs from numpy's global RandomState (:361-395); that stream is the caller's
// contract (np.random.seed), so the pairs are generated on the host with the same calls and handed over as
// coordinates.  Everything that touches the fields runs here:
//   k_sf_gather  : cell index floor((p - lo) / cell) per axis (:400-406), velocities of both cells of every pair
//                  (:408-412) — random 4/8-byte gathers from the [z][y][x] field slabs a rank holds;
//   k_sf_moments : separation unit vector, longitudinal / transverse increments and the order-p sums over the
//                  points of one separation (:417-436), one CTA per separation, fixed-order tree reduction.
// Products and sums are kept un-fused (__dmul_rn / __dadd_rn) so the per-point values round like NumPy's.
#include "common.cuh"

namespace fava {
namespace {

template <typename T>
__global__ void k_sf_gather(const double* __restrict__ pts, int64_t npts, const T* __restrict__ ux,
                            const T* __restrict__ uy, const T* __restrict__ uz, int64_t nz, int64_t ny, int64_t nx,
                            int64_t zf0, int64_t zf1, double lox, double loy, double loz, double cx, double cy,
                            double cz, double* __restrict__ vel, int* __restrict__ err) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npts) return;
    const double px = pts[3 * i], py = pts[3 * i + 1], pz = pts[3 * i + 2];
    const double fx = floor(__ddiv_rn(__dsub_rn(px, lox), cx));
    const double fy = floor(__ddiv_rn(__dsub_rn(py, loy), cy));
    const double fz = floor(__ddiv_rn(__dsub_rn(pz, loz), cz));
    double v[3] = {0.0, 0.0, 0.0};
    if (!(fx >= 0.0 && fx < (double)nx && fy >= 0.0 && fy < (double)ny && fz >= 0.0 && fz < (double)nz)) {
        atomicExch(err, 1);  // the reference's fancy indexing raises IndexError here (a point on the upper face)
    } else {
        const int64_t ix = (int64_t)fx, iy = (int64_t)fy, iz = (int64_t)fz;
        if (iz >= zf0 && iz < zf1) {  // this rank holds the plane; the other ranks contribute exact zeros
            const int64_t o = ((iz - zf0) * ny + iy) * nx + ix;
            v[0] = (double)__ldg(ux + o);
            v[1] = (double)__ldg(uy + o);
            v[2] = (double)__ldg(uz + o);
        }
    }
    vel[3 * i] = v[0];
    vel[3 * i + 1] = v[1];
    vel[3 * i + 2] = v[2];
}

__device__ __forceinline__ double ipow(double x, int order) {
    // numpy: x ** 1 is a copy, x ** 2 a square, other scalar exponents go through pow()
    if (order == 1) return x;
    if (order == 2) return __dmul_rn(x, x);
    return pow(x, (double)order);
}

constexpr int kSfThreads = 512;

// p1, p2: [nsep][npts][3] coordinates; v1, v2: velocities gathered at them; out: [2][nsep] (longitudinal, transverse)
__global__ void __launch_bounds__(kSfThreads)
k_sf_moments(const double* __restrict__ p1, const double* __restrict__ p2, const double* __restrict__ v1,
             const double* __restrict__ v2, int64_t npts, int order, int anisotropic, double* __restrict__ out) {
    __shared__ double red[2][kSfThreads];
    const int64_t base = (int64_t)blockIdx.x * npts;
    double sl = 0.0, st = 0.0;
    for (int64_t j = threadIdx.x; j < npts; j += kSfThreads) {
        const int64_t o = 3 * (base + j);
        double r[3], dv[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            r[a] = __dsub_rn(p2[o + a], p1[o + a]);
            dv[a] = __dsub_rn(v2[o + a], v1[o + a]);
        }
        if (anisotropic) {
            r[0] = 1.0;
            r[1] = 0.0;
            r[2] = 0.0;
        } else {
            const double n2 = __dadd_rn(__dadd_rn(__dmul_rn(r[0], r[0]), __dmul_rn(r[1], r[1])), __dmul_rn(r[2], r[2]));
            const double n = sqrt(n2);
#pragma unroll
            for (int a = 0; a < 3; ++a) r[a] = __ddiv_rn(r[a], n);
        }
        const double lc =
            fabs(__dadd_rn(__dadd_rn(__dmul_rn(dv[0], r[0]), __dmul_rn(dv[1], r[1])), __dmul_rn(dv[2], r[2])));
        double t2 = 0.0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double d = __dsub_rn(dv[a], __dmul_rn(lc, r[a]));
            t2 = a == 0 ? __dmul_rn(d, d) : __dadd_rn(t2, __dmul_rn(d, d));
        }
        sl = __dadd_rn(sl, ipow(lc, order));
        st = __dadd_rn(st, ipow(sqrt(t2), order));
    }
    red[0][threadIdx.x] = sl;
    red[1][threadIdx.x] = st;
    __syncthreads();
    for (int s = kSfThreads / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            red[0][threadIdx.x] += red[0][threadIdx.x + s];
            red[1][threadIdx.x] += red[1][threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[blockIdx.x] = red[0][0] / (double)npts;
        out[gridDim.x + blockIdx.x] = red[1][0] / (double)npts;
    }
}

}  // namespace
}  // namespace fava

using namespace fava;

extern "C" {

int fava_sf_gather(fava_ctx* ctx, const double* d_points, int64_t npoints, const void* d_ux, const void* d_uy,
                   const void* d_uz, int dtype, int64_t nz, int64_t ny, int64_t nx, int64_t zf0, int64_t zf1,
                   const double* h_lo, const double* h_cell, double* d_vel, int* d_err, void* stream) {
    FAVA_REQUIRE(ctx && d_points && d_ux && d_uy && d_uz && h_lo && h_cell && d_vel && d_err,
                 "fava_sf_gather: NULL argument");
    FAVA_REQUIRE(dtype == FAVA_F32 || dtype == FAVA_F64, "fava_sf_gather: bad dtype %d", dtype);
    FAVA_REQUIRE(nz > 0 && ny > 0 && nx > 0, "fava_sf_gather: empty grid");
    FAVA_REQUIRE(0 <= zf0 && zf0 <= zf1 && zf1 <= nz, "fava_sf_gather: plane range [%lld, %lld) outside [0, %lld)",
                 (long long)zf0, (long long)zf1, (long long)nz);
    FAVA_REQUIRE(h_cell[0] > 0.0 && h_cell[1] > 0.0 && h_cell[2] > 0.0, "fava_sf_gather: cell sizes must be positive");
    if (npoints <= 0) return FAVA_OK;
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((npoints + 255) / 256);
    if (dtype == FAVA_F64)
        k_sf_gather<double><<<grid, 256, 0, st>>>(d_points, npoints, (const double*)d_ux, (const double*)d_uy,
                                                  (const double*)d_uz, nz, ny, nx, zf0, zf1, h_lo[0], h_lo[1], h_lo[2],
                                                  h_cell[0], h_cell[1], h_cell[2], d_vel, d_err);
    else
        k_sf_gather<float><<<grid, 256, 0, st>>>(d_points, npoints, (const float*)d_ux, (const float*)d_uy,
                                                 (const float*)d_uz, nz, ny, nx, zf0, zf1, h_lo[0], h_lo[1], h_lo[2],
                                                 h_cell[0], h_cell[1], h_cell[2], d_vel, d_err);
    FAVA_LAUNCHED();
    return FAVA_OK;
}

int fava_sf_moments(fava_ctx* ctx, const double* d_p1, const double* d_p2, const double* d_v1, const double* d_v2,
                    int64_t nsep, int64_t npoints, int order, int anisotropic, double* d_out, void* stream) {
    FAVA_REQUIRE(ctx && d_p1 && d_p2 && d_v1 && d_v2 && d_out, "fava_sf_moments: NULL argument");
    FAVA_REQUIRE(nsep > 0 && npoints > 0, "fava_sf_moments: nsep and npoints must be positive");
    FAVA_REQUIRE(order >= 1, "fava_sf_moments: order must be >= 1");
    DeviceGuard g(ctx->device);
    k_sf_moments<<<(unsigned)nsep, kSfThreads, 0, (cudaStream_t)stream>>>(d_p1, d_p2, d_v1, d_v2, npoints, order,
                                                                         anisotropic, d_out);
    FAVA_LAUNCHED();
    return FAVA_OK;
}

}  // extern "C"
