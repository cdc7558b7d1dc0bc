// Box-counting fractal dimension of an iso-contour (reference FlashUniform.fractal_dimension,
// fava/mesh/FLASH/FlashUniform.py:85-227), the GPU part: edge marking + filled-box counts per level.
//
// Reference algorithm: edata = (field == contour); every INTERIOR cell with val < contour looks at its six
// neighbours nb > contour and flags itself if int((contour - val) / (nb - val)) == 0, otherwise the neighbour
// (:133-177); then for box edges 2^level the boxes holding any flag are counted (:179-208).
//
// Here: flags are only ever set, so the scatter is restated as a per-cell gather (a cell is flagged iff it equals
// the contour, or it is an interior low cell with a "near" crossing, or one of its interior low neighbours has a
// "far" crossing towards it).  The quotient q = h / d with 0 < h <= d (h = contour - val, d = nb - val) truncates
// to 0 iff h < d: h <= pred(d) gives h / d <= 1 - 2^-53, which is representable, so the rounded quotient stays
// below 1; q == 1 exactly iff the rounded differences coincide.  The division is therefore replaced by a
// comparison of the two rounded differences — bit-identical decisions (tests/test_uniform_analysis_*.py check the
// oracle, which divides, against this on adversarial inputs).
//
// One CTA owns a 32^3 tile: warps = y rows, lanes = x, marching in z with the z-neighbours carried in registers,
// x-neighbours by shuffle, y-neighbours through L1.  Every row's flags are one ballot word; the tile's 32x32 words
// are folded level by level in shared memory (levels 0..5), the tile's occupancy goes to a coarse byte grid from
// which k_fractal_coarse counts the levels above.  HBM-bound: s bytes per cell read once (+ halo from L1/L2);
// counts are integers (atomic adds are exact and order-independent).
#include "common.cuh"

namespace fava {
namespace {

constexpr int kTile = 32;
constexpr int kTileLevels = 6;  // box edges 1..32 live inside one tile

// out bit i = in bit 2i | in bit 2i+1
__device__ __forceinline__ uint32_t fold_pairs(uint32_t m) {
    m = (m | (m >> 1)) & 0x55555555u;
    m = (m | (m >> 1)) & 0x33333333u;
    m = (m | (m >> 2)) & 0x0f0f0f0fu;
    m = (m | (m >> 4)) & 0x00ff00ffu;
    m = (m | (m >> 8)) & 0x0000ffffu;
    return m;
}

template <typename T>
__global__ void __launch_bounds__(1024, 1)
k_fractal_tiles(const T* __restrict__ f, int64_t nz, int64_t ny, int64_t nx, int64_t zf0, int64_t tz0, double c,
                unsigned long long* __restrict__ counts, uint8_t* __restrict__ coarse) {
    __shared__ uint32_t words[2][kTile * kTile];
    __shared__ int cnt[kTileLevels];
    const int lane = threadIdx.x & 31, row = threadIdx.x >> 5;
    if (threadIdx.x < kTileLevels) cnt[threadIdx.x] = 0;

    const int64_t x = (int64_t)blockIdx.x * kTile + lane;
    const int64_t y = (int64_t)blockIdx.y * kTile + row;
    const int64_t zt = ((int64_t)blockIdx.z + tz0) * kTile;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const bool in_xy = x < nx && y < ny;
    const bool ix = x >= 1 && x <= nx - 2, iy = y >= 1 && y <= ny - 2;
    const bool ixm = x - 1 >= 1 && x - 1 <= nx - 2, ixp = x + 1 >= 1 && x + 1 <= nx - 2;
    const bool iym = y - 1 >= 1 && y - 1 <= ny - 2, iyp = y + 1 >= 1 && y + 1 <= ny - 2;
    const bool has_up = in_xy && y + 1 < ny, has_dn = in_xy && y >= 1;
    const bool edge_l = lane == 0 && in_xy && x >= 1, edge_r = lane == 31 && in_xy && x + 1 < nx;

    const T* p = f + ((zt - zf0) * ny + (in_xy ? y : 0)) * nx + (in_xy ? x : 0);  // cell (x, y, zt)
    const int64_t plane = ny * nx;
    auto ld = [&](const T* q, bool ok) -> double { return ok ? (double)__ldg(q) : nan; };

    double vm = ld(p - plane, in_xy && zt >= 1);
    double v0 = ld(p, in_xy && zt < nz);
#pragma unroll 4
    for (int s = 0; s < kTile; ++s) {
        const int64_t z = zt + s;
        const bool zin = z < nz;
        const T* q = p + (int64_t)s * plane;
        const double vp = ld(q + plane, in_xy && z + 1 < nz);
        const double vu = ld(q + nx, has_up && zin);
        const double vd = ld(q - nx, has_dn && zin);
        double vl = __shfl_up_sync(0xffffffffu, v0, 1);
        double vr = __shfl_down_sync(0xffffffffu, v0, 1);
        if (lane == 0) vl = ld(q - 1, edge_l && zin);
        if (lane == 31) vr = ld(q + 1, edge_r && zin);

        const bool iz = z >= 1 && z <= nz - 2;
        const bool izm = z - 1 >= 1 && z - 1 <= nz - 2, izp = z + 1 >= 1 && z + 1 <= nz - 2;
        bool m = v0 == c;
        if (v0 < c) {
            if (ix && iy && iz) {  // this cell is visited by the reference loop: near crossings flag it
                const double h = c - v0;
                m = m || (vr > c && h < vr - v0) || (vl > c && h < vl - v0) || (vu > c && h < vu - v0) ||
                    (vd > c && h < vd - v0) || (vp > c && h < vp - v0) || (vm > c && h < vm - v0);
            }
        } else if (v0 > c) {  // a visited low neighbour n flags this cell when its crossing is not near n
            m = m || (ixm && iy && iz && vl < c && !(c - vl < v0 - vl)) ||
                (ixp && iy && iz && vr < c && !(c - vr < v0 - vr)) ||
                (ix && iym && iz && vd < c && !(c - vd < v0 - vd)) ||
                (ix && iyp && iz && vu < c && !(c - vu < v0 - vu)) ||
                (ix && iy && izm && vm < c && !(c - vm < v0 - vm)) ||
                (ix && iy && izp && vp < c && !(c - vp < v0 - vp));
        }
        const uint32_t w = __ballot_sync(0xffffffffu, m);
        if (lane == 0) words[0][s * kTile + row] = w;
        vm = v0;
        v0 = vp;
    }
    __syncthreads();

    // level 0: flagged cells; levels 1..5: fold 2x2x2 children (x pairs inside the word, y/z pairs across words)
    int edge = kTile;
    int src = 0;
    for (int level = 0; level < kTileLevels; ++level) {
        uint32_t w = 0;
        if (level == 0) {
            w = words[0][threadIdx.x];
        } else {
            const int half = edge >> 1;
            if ((int)threadIdx.x < half * half) {
                const int zz = threadIdx.x / half, yy = threadIdx.x % half;
                const uint32_t* a = &words[src][(2 * zz) * edge + 2 * yy];
                w = fold_pairs(a[0] | a[1] | a[edge] | a[edge + 1]);
                words[src ^ 1][zz * half + yy] = w;
            }
            edge = half;
            src ^= 1;
        }
        const int n = __reduce_add_sync(0xffffffffu, __popc(w));
        if (lane == 0 && n) atomicAdd(&cnt[level], n);
        __syncthreads();
    }
    if (threadIdx.x < kTileLevels && cnt[threadIdx.x])
        atomicAdd(&counts[threadIdx.x], (unsigned long long)cnt[threadIdx.x]);
    if (threadIdx.x == 0)
        coarse[(((int64_t)blockIdx.z + tz0) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = (uint8_t)(cnt[5] != 0);
}

// Levels >= 6 on the tile-occupancy grid [ctz][cty][ctx]: box edge 2^(level-5) tiles.
__global__ void k_fractal_coarse(const uint8_t* __restrict__ coarse, int64_t ctz, int64_t cty, int64_t ctx_, int nlevels,
                                 unsigned long long* __restrict__ counts) {
    for (int level = kTileLevels; level < nlevels; ++level) {
        const int64_t e = (int64_t)1 << (level - (kTileLevels - 1));
        const int64_t bx = (ctx_ + e - 1) / e, by = (cty + e - 1) / e, bz = (ctz + e - 1) / e;
        unsigned long long mine = 0;
        for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < bx * by * bz;
             b += (int64_t)gridDim.x * blockDim.x) {
            const int64_t x0 = (b % bx) * e, y0 = ((b / bx) % by) * e, z0 = (b / (bx * by)) * e;
            bool filled = false;
            for (int64_t z = z0; z < min(z0 + e, ctz) && !filled; ++z)
                for (int64_t y = y0; y < min(y0 + e, cty) && !filled; ++y)
                    for (int64_t x = x0; x < min(x0 + e, ctx_); ++x)
                        if (coarse[(z * cty + y) * ctx_ + x]) {
                            filled = true;
                            break;
                        }
            mine += filled;
        }
        if (mine) atomicAdd(&counts[level], mine);
    }
}

}  // namespace
}  // namespace fava

using namespace fava;

extern "C" {

int fava_fractal_tiles(fava_ctx* ctx, const void* d_field, int dtype, int64_t nz, int64_t ny, int64_t nx, int64_t zf0,
                       int64_t zf1, int64_t z0, int64_t z1, double contour, uint64_t* d_counts, uint8_t* d_coarse,
                       void* stream) {
    FAVA_REQUIRE(ctx && d_field && d_counts && d_coarse, "fava_fractal_tiles: NULL argument");
    FAVA_REQUIRE(dtype == FAVA_F32 || dtype == FAVA_F64, "fava_fractal_tiles: bad dtype %d", dtype);
    FAVA_REQUIRE(nz > 0 && ny > 0 && nx > 0, "fava_fractal_tiles: empty array");
    FAVA_REQUIRE(0 <= z0 && z0 < z1 && z1 <= nz, "fava_fractal_tiles: plane range [%lld, %lld) outside [0, %lld)",
                 (long long)z0, (long long)z1, (long long)nz);
    FAVA_REQUIRE(z0 % kTile == 0 && (z1 % kTile == 0 || z1 == nz),
                 "fava_fractal_tiles: the owned plane range must be aligned to %d-plane tiles", kTile);
    FAVA_REQUIRE(zf0 <= (z0 > 0 ? z0 - 1 : 0) && zf1 >= (z1 < nz ? z1 + 1 : nz),
                 "fava_fractal_tiles: the buffer [%lld, %lld) lacks the halo planes of [%lld, %lld)", (long long)zf0,
                 (long long)zf1, (long long)z0, (long long)z1);
    DeviceGuard g(ctx->device);
    const int64_t ctx_ = (nx + kTile - 1) / kTile, cty = (ny + kTile - 1) / kTile;
    const int64_t tz0 = z0 / kTile, tz1 = (z1 + kTile - 1) / kTile;
    FAVA_REQUIRE(cty <= 65535 && tz1 - tz0 <= 65535, "fava_fractal_tiles: grid too large");
    dim3 grid((unsigned)ctx_, (unsigned)cty, (unsigned)(tz1 - tz0));
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == FAVA_F64)
        k_fractal_tiles<double><<<grid, 1024, 0, st>>>((const double*)d_field, nz, ny, nx, zf0, tz0, contour,
                                                       (unsigned long long*)d_counts, d_coarse);
    else
        k_fractal_tiles<float><<<grid, 1024, 0, st>>>((const float*)d_field, nz, ny, nx, zf0, tz0, contour,
                                                      (unsigned long long*)d_counts, d_coarse);
    FAVA_LAUNCHED();
    return FAVA_OK;
}

int fava_fractal_coarse(fava_ctx* ctx, const uint8_t* d_coarse, int64_t nz, int64_t ny, int64_t nx, int nlevels,
                        uint64_t* d_counts, void* stream) {
    FAVA_REQUIRE(ctx && d_coarse && d_counts, "fava_fractal_coarse: NULL argument");
    FAVA_REQUIRE(nz > 0 && ny > 0 && nx > 0, "fava_fractal_coarse: empty array");
    FAVA_REQUIRE(nlevels >= 1 && nlevels <= FAVA_FRACTAL_MAXLEVELS, "fava_fractal_coarse: nlevels %d not in 1..%d",
                 nlevels, FAVA_FRACTAL_MAXLEVELS);
    if (nlevels <= kTileLevels) return FAVA_OK;
    DeviceGuard g(ctx->device);
    const int64_t ctx_ = (nx + kTile - 1) / kTile, cty = (ny + kTile - 1) / kTile, ctz = (nz + kTile - 1) / kTile;
    k_fractal_coarse<<<32, 256, 0, (cudaStream_t)stream>>>(d_coarse, ctz, cty, ctx_, nlevels,
                                                           (unsigned long long*)d_counts);
    FAVA_LAUNCHED();
    return FAVA_OK;
}

}  // extern "C"
