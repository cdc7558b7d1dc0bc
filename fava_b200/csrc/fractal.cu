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
// One CTA owns a 32^3 tile and works in four phases on 32-bit row words (one bit per cell of an x row):
//   A  every row of the tile and of its y/z halo (34 x 34 rows) is read ONCE, coalesced, and classified against the
//      contour: words lt (val < c), gt (val > c) and, for the tile's own rows, eq (val == c).  f32 data is compared
//      in f32 against the contour rounded up / down (val < c <=> val < ru(c), val > c <=> val > rd(c): exact);
//   B  one thread per row combines the words of the six neighbour rows (x neighbours by shifts, the two x-halo
//      cells of the row loaded by that thread) into CANDIDATES: low visited cells with a high neighbour, high cells
//      with a low visited neighbour.  Every other cell is decided: flagged iff eq;
//   C  candidate cells (the cells next to the surface) are evaluated exactly, one lane per cell, with the rule
//      above on the seven fp64 values (L1/L2 hits: the tile was just read);
//   D  the 32 x 32 flag words are folded level by level in shared memory (levels 0..5); the tile's occupancy goes
//      to a coarse byte grid from which k_fractal_coarse counts the levels above.
// HBM-bound: s bytes per cell read once (+ halo rows from L2); counts are integers (atomic adds are exact and
// order-independent).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace fava {
namespace {

constexpr int kTile = 32;
constexpr int kHalo = kTile + 2;  // rows per tile edge including the halo
constexpr int kTileLevels = 6;    // box edges 1..32 live inside one tile

// out bit i = in bit 2i | in bit 2i+1
__device__ __forceinline__ uint32_t fold_pairs(uint32_t m) {
    m = (m | (m >> 1)) & 0x55555555u;
    m = (m | (m >> 1)) & 0x33333333u;
    m = (m | (m >> 2)) & 0x0f0f0f0fu;
    m = (m | (m >> 4)) & 0x00ff00ffu;
    m = (m | (m >> 8)) & 0x0000ffffu;
    return m;
}

// Exact classification of a stored value against the fp64 contour, in the storage precision.
template <typename T>
struct Side;
template <>
struct Side<double> {
    double c;
    __device__ explicit Side(double contour) : c(contour) {}
    __device__ __forceinline__ bool lt(double v) const { return v < c; }
    __device__ __forceinline__ bool gt(double v) const { return v > c; }
};
template <>
struct Side<float> {
    float up, dn;  // smallest f32 >= c, largest f32 <= c
    __device__ explicit Side(double contour) {
#ifdef __CUDA_ARCH__
        up = __double2float_ru(contour);
        dn = __double2float_rd(contour);
#else
        up = dn = (float)contour;
#endif
    }
    __device__ __forceinline__ bool lt(float v) const { return v < up; }
    __device__ __forceinline__ bool gt(float v) const { return v > dn; }
};

template <typename T, int kThreads, int kMinCtas>
__global__ void __launch_bounds__(kThreads, kMinCtas)
k_fractal_tiles(const T* __restrict__ f, int64_t nz, int64_t ny, int64_t nx, int64_t zf0, int64_t tz0, double c,
                unsigned long long* __restrict__ counts, uint8_t* __restrict__ coarse) {
    __shared__ uint2 s_lg[kHalo * kHalo];  // (lt, gt) words of row [z + 1][y + 1], halo rows included
    __shared__ uint32_t s_flag[kTile * kTile], s_cand[kTile * kTile];  // [z][y]
    __shared__ uint16_t s_rows[kTile * kTile];  // rows holding candidates (any order)
    __shared__ int cnt[kTileLevels], s_nrows;
    constexpr int kWarps = kThreads / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < kTileLevels) cnt[threadIdx.x] = 0;
    if (threadIdx.x == kTileLevels) s_nrows = 0;

    const int64_t x0 = (int64_t)blockIdx.x * kTile, y0 = (int64_t)blockIdx.y * kTile;
    const int64_t zt = ((int64_t)blockIdx.z + tz0) * kTile;
    const int64_t plane = ny * nx;
    const T* base = f - zf0 * plane;  // base[(z * ny + y) * nx + x] for the planes z the buffer holds
    const Side<T> side(c);
    const double dnan = __longlong_as_double(0x7ff8000000000000LL);
    const T tnan = (T)dnan;

    // ---- A: classify the rows of the tile and of its y/z halo (the four corner lines are never needed) ---------
    // Loads are issued in register batches before any vote so that kBatch row requests per warp are in flight.
    const bool x_in = x0 + lane < nx;
    auto classify = [&](T v, int zzi, int yyi) {  // halo-inclusive row coordinates 0..33
        const bool lt = side.lt(v), gt = side.gt(v);
        const uint32_t wl = __ballot_sync(0xffffffffu, lt), wg = __ballot_sync(0xffffffffu, gt);
        uint32_t we = 0;
        if ((wl | wg) != 0xffffffffu) we = __ballot_sync(0xffffffffu, !lt && !gt && v == v);  // warp-uniform
        if (lane == 0) {
            s_lg[zzi * kHalo + yyi] = make_uint2(wl, wg);  // one 64-bit store
            if (zzi >= 1 && zzi <= kTile && yyi >= 1 && yyi <= kTile) s_flag[(zzi - 1) * kTile + (yyi - 1)] = we;
        }
    };
    constexpr int kBatch = sizeof(T) == 4 ? 12 : 9;
    // halo-inclusive plane indices zzi in [zz_lo, zz_hi) exist in the domain (plane zt - 1 + zzi)
    const int zz_lo = zt == 0 ? 1 : 0, zz_hi = (int)min((int64_t)kHalo, nz - zt + 1);
#pragma unroll 1
    for (int yyi = warp + 1; yyi <= kTile; yyi += kWarps) {  // z-march over the tile's own y rows
        const int64_t y = y0 + yyi - 1;
        const bool ok = x_in && y < ny;
        const T* q = base + ((zt - 1) * ny + y) * nx + x0 + lane;  // row (y, zt - 1); dereferenced only when valid
#pragma unroll 1
        for (int zb = 0; zb < kHalo; zb += kBatch) {
            T v[kBatch];
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                v[b] = (ok && zb + b >= zz_lo && zb + b < zz_hi) ? __ldg(q) : tnan;
                q += plane;
            }
#pragma unroll
            for (int b = 0; b < kBatch; ++b)
                if (zb + b < kHalo) classify(v[b], zb + b, yyi);
        }
    }
    {  // the two y-halo lines of the tile: 2 x 32 rows, kYh per warp
        constexpr int kYh = 2 * kTile / kWarps, kPer = kYh / 2;
        T v[kYh];
#pragma unroll
        for (int b = 0; b < kYh; ++b) {
            const int zzi = 1 + kPer * warp + (b % kPer), yyi = (b / kPer) * (kTile + 1);
            const int64_t z = zt + zzi - 1, y = y0 + yyi - 1;
            v[b] = (x_in && z < nz && y >= 0 && y < ny) ? __ldg(base + (z * ny + y) * nx + x0 + lane) : tnan;
        }
#pragma unroll
        for (int b = 0; b < kYh; ++b) classify(v[b], 1 + kPer * warp + (b % kPer), (b / kPer) * (kTile + 1));
    }
    __syncthreads();

    // ---- B: candidates, one thread per row -------------------------------------------------------------------
    uint32_t xvis = 0;  // bit i: cell x0 + i is visited by the reference loop along x (1 <= x <= nx - 2)
    {
        const int64_t lo = max((int64_t)1, x0), hi = min(nx - 2, x0 + kTile - 1);
        if (hi >= lo) xvis = (uint32_t)((((uint64_t)1 << (hi - lo + 1)) - 1) << (lo - x0));
    }
    const bool xvis_left = x0 - 1 >= 1 && x0 - 1 <= nx - 2, xvis_right = x0 + kTile <= nx - 2;
    for (int row = threadIdx.x; row < kTile * kTile; row += kThreads) {
        const int zz = row / kTile, yy = row % kTile;
        const int64_t z = zt + zz, y = y0 + yy;
        uint32_t cand = 0;
        if (z < nz && y < ny) {
            const int h = (zz + 1) * kHalo + (yy + 1);
            const uint32_t l = s_lg[h].x, g = s_lg[h].y;
            const T* rowp = base + (z * ny + y) * nx + x0;
            const T vl = x0 >= 1 ? __ldg(rowp - 1) : tnan;  // the row's two x-halo cells
            const T vr = x0 + kTile < nx ? __ldg(rowp + kTile) : tnan;
            const bool yv = y >= 1 && y <= ny - 2, zv = z >= 1 && z <= nz - 2;
            const bool yv_m = y - 1 >= 1, yv_p = y + 1 <= ny - 2, zv_m = z - 1 >= 1, zv_p = z + 1 <= nz - 2;
            if (yv && zv) {
                // low visited cells with a high neighbour
                const uint32_t g_any = (g >> 1) | (side.gt(vr) ? 0x80000000u : 0u) | (g << 1) | (side.gt(vl) ? 1u : 0u) |
                                       s_lg[h + 1].y | s_lg[h - 1].y | s_lg[h + kHalo].y | s_lg[h - kHalo].y;
                cand = l & xvis & g_any;
            }
            // high cells with a low visited neighbour
            uint32_t l_vis = 0;
            if (yv && zv) {
                const uint32_t lx = l & xvis;
                l_vis = (lx >> 1) | ((xvis_right && side.lt(vr)) ? 0x80000000u : 0u) | (lx << 1) |
                        ((xvis_left && side.lt(vl)) ? 1u : 0u);
            }
            if (yv_p && zv) l_vis |= s_lg[h + 1].x & xvis;
            if (yv_m && zv) l_vis |= s_lg[h - 1].x & xvis;
            if (yv && zv_p) l_vis |= s_lg[h + kHalo].x & xvis;
            if (yv && zv_m) l_vis |= s_lg[h - kHalo].x & xvis;
            cand |= g & l_vis;
        }
        s_cand[row] = cand;
        if (cand) s_rows[atomicAdd(&s_nrows, 1)] = (uint16_t)row;
    }
    __syncthreads();

    // ---- C: exact rule on the candidate cells, one lane per cell ----------------------------------------------
    {
        const int64_t x = x0 + lane;
        const bool xv = x >= 1 && x <= nx - 2, xv_m = x - 1 >= 1 && x - 1 <= nx - 2, xv_p = x + 1 >= 1 && x + 1 <= nx - 2;
        const int nrows = s_nrows;
        for (int i = warp; i < nrows; i += kWarps) {
            const int row = s_rows[i];
            const uint32_t cand = s_cand[row];
            bool m = false;
            if ((cand >> lane) & 1u) {
                const int64_t z = zt + row / kTile, y = y0 + row % kTile;
                const T* q = base + (z * ny + y) * nx + x;
                const double v0 = (double)__ldg(q);
                const double vl = x >= 1 ? (double)__ldg(q - 1) : dnan, vr = x + 1 < nx ? (double)__ldg(q + 1) : dnan;
                const double vd = y >= 1 ? (double)__ldg(q - nx) : dnan, vu = y + 1 < ny ? (double)__ldg(q + nx) : dnan;
                const double vm = z >= 1 ? (double)__ldg(q - plane) : dnan, vp = z + 1 < nz ? (double)__ldg(q + plane) : dnan;
                const bool yv = y >= 1 && y <= ny - 2, zv = z >= 1 && z <= nz - 2;
                const bool yv_m = y - 1 >= 1 && y - 1 <= ny - 2, yv_p = y + 1 >= 1 && y + 1 <= ny - 2;
                const bool zv_m = z - 1 >= 1 && z - 1 <= nz - 2, zv_p = z + 1 >= 1 && z + 1 <= nz - 2;
                if (v0 < c) {
                    if (xv && yv && zv) {  // visited by the reference loop: near crossings flag this cell
                        const double h = c - v0;
                        m = (vr > c && h < vr - v0) || (vl > c && h < vl - v0) || (vu > c && h < vu - v0) ||
                            (vd > c && h < vd - v0) || (vp > c && h < vp - v0) || (vm > c && h < vm - v0);
                    }
                } else if (v0 > c) {  // a visited low neighbour n flags this cell when its crossing is not near n
                    m = (xv_m && yv && zv && vl < c && !(c - vl < v0 - vl)) ||
                        (xv_p && yv && zv && vr < c && !(c - vr < v0 - vr)) ||
                        (xv && yv_m && zv && vd < c && !(c - vd < v0 - vd)) ||
                        (xv && yv_p && zv && vu < c && !(c - vu < v0 - vu)) ||
                        (xv && yv && zv_m && vm < c && !(c - vm < v0 - vm)) ||
                        (xv && yv && zv_p && vp < c && !(c - vp < v0 - vp));
                }
            }
            const uint32_t w = __ballot_sync(0xffffffffu, m);
            if (lane == 0 && w) s_flag[row] |= w;
        }
    }
    __syncthreads();

    // ---- D: level 0 = flagged cells; levels 1..5 fold 2x2x2 children (x pairs inside the word, y/z across words)
    uint32_t* src = s_flag;
    uint32_t* dst = s_cand;
    int edge = kTile;
    for (int level = 0; level < kTileLevels; ++level) {
        int n = 0;
        if (level == 0) {
            for (int i = threadIdx.x; i < kTile * kTile; i += kThreads) n += __popc(src[i]);
        } else {
            const int half = edge >> 1;
            for (int i = threadIdx.x; i < half * half; i += kThreads) {
                const int zz = i / half, yy = i % half;
                const uint32_t* a = &src[(2 * zz) * edge + 2 * yy];
                const uint32_t w = fold_pairs(a[0] | a[1] | a[edge] | a[edge + 1]);
                dst[i] = w;
                n += __popc(w);
            }
            edge = half;
            uint32_t* t = src;
            src = dst;
            dst = t;
        }
        n = __reduce_add_sync(0xffffffffu, n);
        if (lane == 0 && n) atomicAdd(&cnt[level], n);
        __syncthreads();
    }
    if (threadIdx.x < kTileLevels && cnt[threadIdx.x])
        atomicAdd(&counts[threadIdx.x], (unsigned long long)cnt[threadIdx.x]);
    if (threadIdx.x == 0)
        coarse[(((int64_t)blockIdx.z + tz0) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = (uint8_t)(cnt[5] != 0);
}

// Levels >= 6: repeated 2x2x2 OR-reduction of the tile-occupancy grid [sz][sy][sx] (one CTA; the reduced grids
// ping-pong between two workspace buffers), counting the occupied boxes of every level.
__global__ void __launch_bounds__(1024)
k_fractal_coarse(const uint8_t* __restrict__ coarse, int64_t sz, int64_t sy, int64_t sx, int nlevels, uint8_t* w0,
                 uint8_t* w1, unsigned long long* __restrict__ counts) {
    __shared__ unsigned long long total;
    const uint8_t* src = coarse;
    uint8_t* dst = w0;
    for (int level = kTileLevels; level < nlevels; ++level) {
        const int64_t dz = (sz + 1) / 2, dy = (sy + 1) / 2, dx = (sx + 1) / 2;
        if (threadIdx.x == 0) total = 0;
        __syncthreads();
        unsigned long long mine = 0;
        for (int64_t i = threadIdx.x; i < dz * dy * dx; i += blockDim.x) {
            const int64_t x = 2 * (i % dx), y = 2 * ((i / dx) % dy), z = 2 * (i / (dx * dy));
            unsigned o = 0;
            for (int64_t zz = z; zz < min(z + 2, sz); ++zz)
                for (int64_t yy = y; yy < min(y + 2, sy); ++yy)
                    for (int64_t xx = x; xx < min(x + 2, sx); ++xx) o |= src[(zz * sy + yy) * sx + xx];
            dst[i] = (uint8_t)(o != 0);
            mine += o != 0;
        }
        if (mine) atomicAdd(&total, mine);
        __syncthreads();  // the block's writes to dst are visible to the block after the barrier
        if (threadIdx.x == 0 && total) atomicAdd(&counts[level], total);
        src = dst;
        dst = dst == w0 ? w1 : w0;
        sz = dz;
        sy = dy;
        sx = dx;
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
    // CTA shape: the kernel is bound by its per-row instruction stream, not by occupancy - 512x2, 512x3, 256x4 and
    // 256x6 (threads x CTAs/SM) measured within 15 % of each other on B200 (profiles/r01_uniform_analysis_kernels.json);
    // 256x4 was the fastest and is the one that is built.
    if (dtype == FAVA_F64)
        k_fractal_tiles<double, 256, 4><<<grid, 256, 0, st>>>((const double*)d_field, nz, ny, nx, zf0, tz0, contour,
                                                               (unsigned long long*)d_counts, d_coarse);
    else
        k_fractal_tiles<float, 256, 4><<<grid, 256, 0, st>>>((const float*)d_field, nz, ny, nx, zf0, tz0, contour,
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
    const int64_t half = ((ctz + 1) / 2) * ((cty + 1) / 2) * ((ctx_ + 1) / 2);
    void* ws = nullptr;
    int rc = ctx_workspace(ctx, WS_AUX, (size_t)(2 * half), &ws);
    if (rc != FAVA_OK) return rc;
    k_fractal_coarse<<<1, 1024, 0, (cudaStream_t)stream>>>(d_coarse, ctz, cty, ctx_, nlevels, (uint8_t*)ws,
                                                           (uint8_t*)ws + half, (unsigned long long*)d_counts);
    FAVA_LAUNCHED();
    return FAVA_OK;
}

}  // extern "C"
