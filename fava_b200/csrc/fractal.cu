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
//   A  every row of the tile and of its y/z halo (34 x 34 rows) is read ONCE and classified against the contour:
//      words lt (val < c), gt (val > c) and, for the tile's own rows, eq (val == c).  f32 data is compared in f32
//      against the contour rounded up / down (val < c <=> val < ru(c), val > c <=> val > rd(c): exact).  Two front
//      ends: k_fractal_tiles_tma (rows of a multiple of 16 bytes: planes by TMA into per-warp slots, a lane
//      classifies a whole row out of shared memory) and k_fractal_tiles (any shape: coalesced loads + ballots);
//   B  one thread per row combines the words of the six neighbour rows (x neighbours by shifts, the two x-halo
//      cells of the row from phase A) into CANDIDATES: low visited cells with a high neighbour, high cells with a
//      low visited neighbour.  Every other cell is decided: flagged iff eq;
//   C  candidate cells (the cells next to the surface) are evaluated exactly with the rule above on the seven fp64
//      values, dealt to the lanes cell by cell (row by row in tiles crowded with candidates), branch-free;
//   D  the 32 x 32 flag words are folded level by level in shared memory (levels 0..5); the tile's occupancy goes
//      to a coarse byte grid from which k_fractal_coarse counts the levels above.
// HBM-bound: s bytes per cell read once (+ halo planes / lines, 6 %); counts are integers (atomic adds are exact
// and order-independent).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace fava {
namespace {

constexpr int kTile = 32;
constexpr int kHalo = kTile + 2;  // rows per tile edge including the halo
constexpr int kTileLevels = 6;    // box edges 1..32 live inside one tile
constexpr int kDenseCells = 4096; // candidate cells (of 32768) from which phase C goes row by row

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

// Phases B-D of a tile, shared by the two front ends.  In: s_lg (lt, gt) words of the 34 x 34 rows, s_flag = eq
// words of the tile's rows, s_xh[side][row] = (lt, gt) bits of each row's x-halo cells (bit 0 / bit 1; none for a
// cell outside the domain).  Out: counts of levels 0..5 and the tile's occupancy byte.
template <typename T, int kThreads>
__device__ __forceinline__ void fractal_finish(const T* __restrict__ base, int64_t nz, int64_t ny, int64_t nx, int64_t x0,
                                               int64_t y0, int64_t zt, int64_t plane, double c, const uint2* s_lg,
                                               uint32_t* s_flag, uint32_t* s_cand, uint16_t* s_rows,
                                               const uint8_t* s_xh, int* cnt, int* s_nrows_p,
                                               unsigned long long* __restrict__ counts, uint8_t* __restrict__ coarse,
                                               int64_t coarse_index) {
    constexpr int kWarps = kThreads / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double dnan = __longlong_as_double(0x7ff8000000000000LL);
    int& s_nrows = *s_nrows_p;
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
            const unsigned hl = s_xh[row], hr = s_xh[kTile * kTile + row];  // the row's two x-halo cells
            const bool vl_lt = hl & 1u, vl_gt = hl & 2u, vr_lt = hr & 1u, vr_gt = hr & 2u;
            const bool yv = y >= 1 && y <= ny - 2, zv = z >= 1 && z <= nz - 2;
            const bool yv_m = y - 1 >= 1, yv_p = y + 1 <= ny - 2, zv_m = z - 1 >= 1, zv_p = z + 1 <= nz - 2;
            if (yv && zv) {
                // low visited cells with a high neighbour
                const uint32_t g_any = (g >> 1) | (vr_gt ? 0x80000000u : 0u) | (g << 1) | (vl_gt ? 1u : 0u) |
                                       s_lg[h + 1].y | s_lg[h - 1].y | s_lg[h + kHalo].y | s_lg[h - kHalo].y;
                cand = l & xvis & g_any;
            }
            // high cells with a low visited neighbour
            uint32_t l_vis = 0;
            if (yv && zv) {
                const uint32_t lx = l & xvis;
                l_vis = (lx >> 1) | ((xvis_right && vr_lt) ? 0x80000000u : 0u) | (lx << 1) |
                        ((xvis_left && vl_lt) ? 1u : 0u);
            }
            if (yv_p && zv) l_vis |= s_lg[h + 1].x & xvis;
            if (yv_m && zv) l_vis |= s_lg[h - 1].x & xvis;
            if (yv && zv_p) l_vis |= s_lg[h + kHalo].x & xvis;
            if (yv && zv_m) l_vis |= s_lg[h - kHalo].x & xvis;
            cand |= g & l_vis;
        }
        s_cand[row] = cand;
        if (cand) {
            s_rows[atomicAdd(&s_nrows, 1)] = (uint16_t)row;
            atomicAdd(&s_nrows_p[1], __popc(cand));  // candidate cells of the tile
        }
    }
    __syncthreads();
    const bool crowded = s_nrows_p[1] >= kDenseCells;  // e.g. a contour through noise

    // ---- C: exact rule on the candidate cells ----------------------------------------------------------------
    // A smooth surface leaves two or three candidates in a row, so a lane per cell OF A ROW idles 29 lanes and
    // serialises a tile into ~128 dependent-load rounds per warp.  A warp therefore takes 32 candidate rows, counts
    // their candidates (prefix sum over the lanes) and deals the CELLS out to the lanes, 32 at a time: a lane finds
    // the row that holds its cell by a 5-step search over the prefix sums and the bit by __fns.
    // Branch-free on purpose: written with && / || the six-term rules compile to a chain of branches on which the
    // lanes of a warp part ways for good (profiled: 1-2 active threads per instruction through the whole loop, 15 of
    // the 16 G warp instructions of a noise field).  Comparisons with NaN (cells outside the domain) are false.
    auto exact = [&](int row, int bit) -> bool {
        const int64_t x = x0 + bit;
        const bool xv = (x >= 1) & (x <= nx - 2), xv_m = (x - 1 >= 1) & (x - 1 <= nx - 2), xv_p = (x + 1 >= 1) & (x + 1 <= nx - 2);
        const int64_t z = zt + row / kTile, y = y0 + row % kTile;
        const T* q = base + (z * ny + y) * nx + x;
        const double v0 = (double)__ldg(q);
        const double vl = x >= 1 ? (double)__ldg(q - 1) : dnan, vr = x + 1 < nx ? (double)__ldg(q + 1) : dnan;
        const double vd = y >= 1 ? (double)__ldg(q - nx) : dnan, vu = y + 1 < ny ? (double)__ldg(q + nx) : dnan;
        const double vm = z >= 1 ? (double)__ldg(q - plane) : dnan, vp = z + 1 < nz ? (double)__ldg(q + plane) : dnan;
        const bool yv = (y >= 1) & (y <= ny - 2), zv = (z >= 1) & (z <= nz - 2);
        const bool yv_m = (y - 1 >= 1) & (y - 1 <= ny - 2), yv_p = (y + 1 >= 1) & (y + 1 <= ny - 2);
        const bool zv_m = (z - 1 >= 1) & (z - 1 <= nz - 2), zv_p = (z + 1 >= 1) & (z + 1 <= nz - 2);
        // a low cell visited by the reference loop: a near crossing towards a high neighbour flags the cell itself
        const double h = c - v0;
        auto near = [&](double vn) { return (vn > c) & (h < vn - v0); };
        const bool m_low = (v0 < c) & xv & yv & zv & (near(vr) | near(vl) | near(vu) | near(vd) | near(vp) | near(vm));
        // a high cell: a visited low neighbour n flags it when n's crossing towards it is not near n
        auto far = [&](double vn, bool visited) { return visited & (vn < c) & !(c - vn < v0 - vn); };
        const bool m_high = (v0 > c) & (far(vl, xv_m & yv & zv) | far(vr, xv_p & yv & zv) | far(vd, xv & yv_m & zv) |
                                        far(vu, xv & yv_p & zv) | far(vm, xv & yv & zv_m) | far(vp, xv & yv & zv_p));
        return m_low | m_high;
    };
    if (crowded) {  // row by row, lane = x, the warps on neighbouring rows (they share their neighbour loads in L1)
        const int nrows = s_nrows;
        for (int i = warp; i < nrows; i += kWarps) {
            const int row = s_rows[i];
            const bool m = ((s_cand[row] >> lane) & 1u) ? exact(row, lane) : false;
            const uint32_t w = __ballot_sync(0xffffffffu, m);
            if (lane == 0 && w) s_flag[row] |= w;
        }
    } else {
        const int nrows = s_nrows;
        for (int r0 = warp * 32; r0 < nrows; r0 += kWarps * 32) {
            const int ri = r0 + lane;
            const int my_row = ri < nrows ? s_rows[ri] : 0;
            const uint32_t my_cand = ri < nrows ? s_cand[my_row] : 0u;
            const int n = __popc(my_cand);
            int incl = n;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += up;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31), excl = incl - n;
            for (int c0 = 0; c0 < total; c0 += 32) {
                const int ci = c0 + lane;  // this lane's cell among the chunk's candidates
                int pos = 0;               // the last lane whose exclusive prefix is <= ci holds the cell
#pragma unroll
                for (int step = 16; step >= 1; step >>= 1) {
                    const int e = __shfl_sync(0xffffffffu, excl, (pos + step) & 31);
                    if (e <= ci) pos += step;
                }
                const int row = __shfl_sync(0xffffffffu, my_row, pos);
                const uint32_t cand = __shfl_sync(0xffffffffu, my_cand, pos);
                const int before = __shfl_sync(0xffffffffu, excl, pos);
                if (ci < total) {
                    const int bit = (int)__fns(cand, 0, ci - before + 1);  // the (ci - before)-th candidate of the row
                    if (exact(row, bit)) atomicOr(&s_flag[row], 1u << bit);  // OR: the order does not matter
                }
            }
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
    if (threadIdx.x == 0) coarse[coarse_index] = (uint8_t)(cnt[5] != 0);
}

// ---- tensor-map front end (rows of a multiple of 16 bytes) ------------------------------------------------------
// The ballot front end below spends 42 warp instructions per 32-cell row word (address and range predicates, two
// votes, single-lane stores) and was bound by instruction issue at 0.35-0.42 of the HBM peak.  Here the planes of a
// tile arrive by TMA (`cp.async.bulk.tensor.3d`, box 32 x 32 x 1; cells outside the array read as NaN, which is
// neither below nor above the contour - the halo planes and ragged edges need no predicate) and a LANE classifies
// a whole row out of shared memory: 32 cells as 16-byte words, started at a lane-dependent word so that the 32 rows of
// a warp hit the banks evenly, compared and OR-ed into the lane's own (lt, gt) words - about 7 warp instructions
// per row word.  A warp owns planes w, w + 8, ... of the tile (with its own slots and mbarriers: no CTA-wide barrier
// while planes stream); the two y-halo lines are two more boxes (32 x 1 x 34), and the x-halo cells of a row are
// two scalar loads issued before the wait on the row's plane - the neighbouring tiles fetch the same sectors as part
// of their planes at about the same time, so they hit L2 (fetched after the tile they cost 0.5 x the field in DRAM
// reads).
constexpr int kTmaWarps = 8;
constexpr int kTmaThreads = kTmaWarps * 32;

__device__ __forceinline__ void tma_box_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}

// (lt, gt) words of one 32-cell row in shared memory.
template <typename T>
__device__ __forceinline__ void row_words(const T* __restrict__ row, int lane, const Side<T>& side, uint32_t& lt,
                                          uint32_t& gt) {
    constexpr int CPW = 16 / sizeof(T);  // cells per 16-byte word
    constexpr int NW = kTile / CPW;      // words per row
    lt = gt = 0;
#pragma unroll
    for (int j = 0; j < NW; ++j) {
        const int w = (j + lane) & (NW - 1);  // rotated start: the 32 rows of a warp hit the banks evenly
        T v[CPW];
        if (sizeof(T) == 4) {
            const float4 q = *reinterpret_cast<const float4*>(row + w * CPW);
            v[0] = (T)q.x, v[1] = (T)q.y, v[CPW - 2] = (T)q.z, v[CPW - 1] = (T)q.w;
        } else {
            const double2 q = *reinterpret_cast<const double2*>(row + w * CPW);
            v[0] = (T)q.x, v[CPW - 1] = (T)q.y;
        }
        uint32_t l = 0, g = 0;
#pragma unroll
        for (int k = 0; k < CPW; ++k) {
            l |= side.lt(v[k]) ? (1u << k) : 0u;
            g |= side.gt(v[k]) ? (1u << k) : 0u;
        }
        lt |= l << (w * CPW), gt |= g << (w * CPW);
    }
}

// eq word of a row: cells equal to the contour.  Only rows with a cell that is neither below nor above get here
// (the contour value itself, or NaN): ceq = the contour in the storage type if it is representable there, NaN
// otherwise, so that v == c exactly <=> v == ceq.
template <typename T>
__device__ __noinline__ uint32_t row_eq_word(const T* __restrict__ row, T ceq) {
    uint32_t eq = 0;
    for (int i = 0; i < kTile; ++i) eq |= row[i] == ceq ? (1u << i) : 0u;
    return eq;
}

template <typename T, int SLOTS>
__global__ void __launch_bounds__(kTmaThreads)
k_fractal_tiles_tma(const __grid_constant__ CUtensorMap tm_plane, const __grid_constant__ CUtensorMap tm_yline,
                    const T* __restrict__ f, int64_t nz, int64_t ny, int64_t nx, int64_t zf0, int64_t tz0, double c,
                    unsigned long long* __restrict__ counts, uint8_t* __restrict__ coarse) {
    extern __shared__ __align__(128) unsigned char tile_smem[];
    constexpr int kPlane = kTile * kTile;  // cells of a plane box
    T* slots = reinterpret_cast<T*>(tile_smem);                        // [warp][SLOTS][32][32]
    T* ylines = slots + kTmaWarps * SLOTS * kPlane;                    // [side][34][32]
    uint2* s_lg = reinterpret_cast<uint2*>(ylines + 2 * kHalo * kTile);  // [34][34]
    uint32_t* s_flag = reinterpret_cast<uint32_t*>(s_lg + kHalo * kHalo);
    uint32_t* s_cand = s_flag + kPlane;
    uint16_t* s_rows = reinterpret_cast<uint16_t*>(s_cand + kPlane);
    uint8_t* s_xh = reinterpret_cast<uint8_t*>(s_rows + kPlane);       // [side][32][32]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_xh + 2 * kPlane);   // [warp][SLOTS], then the y lines'
    int* cnt = reinterpret_cast<int*>(bars + kTmaWarps * SLOTS + 1);   // [6], then s_nrows, candidate cells
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kTmaWarps * SLOTS + 1; ++i) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async;\n" ::: "memory");
    }
    if (threadIdx.x < kTileLevels + 2) cnt[threadIdx.x] = 0;
    __syncthreads();

    const int64_t x0 = (int64_t)blockIdx.x * kTile, y0 = (int64_t)blockIdx.y * kTile;
    const int64_t zt = ((int64_t)blockIdx.z + tz0) * kTile;
    const int64_t plane = ny * nx;
    const T* base = f - zf0 * plane;
    const Side<T> side(c);
    const T tnan = (T)__longlong_as_double(0x7ff8000000000000LL);
    const T ceq = (double)(T)c == c ? (T)c : tnan;
    const int zc0 = (int)(zt - 1 - zf0);  // box coordinate of halo-inclusive plane 0 (may be -1: all NaN)

    // ---- A: planes warp, warp + 8, ... of the 34; row `lane` of a plane is this lane's ----
    uint64_t* my_bar = bars + warp * SLOTS;
    T* my_slot = slots + (size_t)warp * SLOTS * kPlane;
    auto request = [&](int k) {  // k-th plane of this warp into slot k % SLOTS (lane 0)
        const int sl = k % SLOTS;
        mbar_expect_tx(&my_bar[sl], (unsigned)(kPlane * sizeof(T)));
        tma_box_3d(my_slot + (size_t)sl * kPlane, &tm_plane, (int)x0, (int)y0, zc0 + warp + kTmaWarps * k, &my_bar[sl]);
    };
    const int nmine = (kHalo - warp + kTmaWarps - 1) / kTmaWarps;
    if (lane == 0) {
        for (int k = 0; k < SLOTS && k < nmine; ++k) request(k);
        if (warp == 0) {  // the two y-halo lines of all 34 planes
            uint64_t* yb = &bars[kTmaWarps * SLOTS];
            mbar_expect_tx(yb, (unsigned)(2 * kHalo * kTile * sizeof(T)));
            tma_box_3d(ylines, &tm_yline, (int)x0, (int)y0 - 1, zc0, yb);
            tma_box_3d(ylines + kHalo * kTile, &tm_yline, (int)x0, (int)y0 + kTile, zc0, yb);
        }
    }
    for (int k = 0; k < nmine; ++k) {
        const int zzi = warp + kTmaWarps * k, sl = k % SLOTS;
        const bool own = zzi >= 1 && zzi <= kTile;
        const int64_t z = zt - 1 + zzi, y = y0 + lane;
        T hl = tnan, hr = tnan;
        if (own && z < nz && y < ny) {
            const T* rowp = base + (z * ny + y) * nx + x0;
            if (x0 >= 1) hl = __ldg(rowp - 1);
            if (x0 + kTile < nx) hr = __ldg(rowp + kTile);
        }
        mbar_wait(&my_bar[sl], (unsigned)((k / SLOTS) & 1));
        uint32_t lt, gt;
        const T* my_row = my_slot + (size_t)sl * kPlane + lane * kTile;
        row_words<T>(my_row, lane, side, lt, gt);
        s_lg[zzi * kHalo + lane + 1] = make_uint2(lt, gt);
        if (own) {
            const int row = (zzi - 1) * kTile + lane;
            s_flag[row] = (lt | gt) == 0xffffffffu ? 0u : row_eq_word<T>(my_row, ceq);
            s_xh[row] = (uint8_t)((side.lt(hl) ? 1 : 0) | (side.gt(hl) ? 2 : 0));
            s_xh[kPlane + row] = (uint8_t)((side.lt(hr) ? 1 : 0) | (side.gt(hr) ? 2 : 0));
        }
        __syncwarp();  // every lane has read its row: the slot can take the warp's next plane
        if (lane == 0 && k + SLOTS < nmine) request(k + SLOTS);
    }
    if (warp >= kTmaWarps - 3) {  // y-halo row words: 2 sides x 34 planes, on the warps that own four planes
        const int r = (warp - (kTmaWarps - 3)) * 32 + lane;
        mbar_wait(&bars[kTmaWarps * SLOTS], 0);
        if (r < 2 * kHalo) {
            const int sd = r / kHalo, zzi = r - sd * kHalo;
            uint32_t lt, gt;
            row_words<T>(ylines + (size_t)r * kTile, lane, side, lt, gt);
            s_lg[zzi * kHalo + (sd ? kTile + 1 : 0)] = make_uint2(lt, gt);
        }
    }
    __syncthreads();
    fractal_finish<T, kTmaThreads>(base, nz, ny, nx, x0, y0, zt, plane, c, s_lg, s_flag, s_cand, s_rows, s_xh, cnt,
                                   &cnt[kTileLevels], counts, coarse,
                                   (((int64_t)blockIdx.z + tz0) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x);
}

template <typename T, int SLOTS>
constexpr size_t tma_tile_smem() {
    return sizeof(T) * (kTmaWarps * SLOTS * kTile * kTile + 2 * kHalo * kTile) + sizeof(uint2) * kHalo * kHalo +
           2 * sizeof(uint32_t) * kTile * kTile + sizeof(uint16_t) * kTile * kTile + 2 * kTile * kTile +
           sizeof(uint64_t) * (kTmaWarps * SLOTS + 1) + sizeof(int) * (kTileLevels + 2);
}

// Tile (bx, by, bz) of a grid of gx x gy tiles per plane of tiles; bz counts from the first tile plane of the call.
template <typename T, int kThreads>
__device__ __forceinline__ void ballot_tile(const T* __restrict__ f, int64_t nz, int64_t ny, int64_t nx, int64_t zf0,
                                            int64_t tz0, double c, unsigned long long* __restrict__ counts,
                                            uint8_t* __restrict__ coarse, int bx, int by, int bz, int gx, int gy) {
    __shared__ uint2 s_lg[kHalo * kHalo];  // (lt, gt) words of row [z + 1][y + 1], halo rows included
    __shared__ uint32_t s_flag[kTile * kTile], s_cand[kTile * kTile];  // [z][y]
    __shared__ uint16_t s_rows[kTile * kTile];  // rows holding candidates (any order)
    __shared__ uint8_t s_xh[2 * kTile * kTile];  // [side][z][y]: (lt, gt) bits of the rows' x-halo cells
    __shared__ int cnt[kTileLevels + 2];  // box counts of levels 0..5, candidate rows, candidate cells
    constexpr int kWarps = kThreads / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < kTileLevels + 2) cnt[threadIdx.x] = 0;

    const int64_t x0 = (int64_t)bx * kTile, y0 = (int64_t)by * kTile;
    const int64_t zt = ((int64_t)bz + tz0) * kTile;
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

    // x-halo cells of the tile's rows (late, scattered loads: this front end serves the shapes the tensor-map one
    // cannot address)
    for (int row = threadIdx.x; row < kTile * kTile; row += kThreads) {
        const int64_t z = zt + row / kTile, y = y0 + row % kTile;
        uint8_t hl = 0, hr = 0;
        if (z < nz && y < ny) {
            const T* rowp = base + (z * ny + y) * nx + x0;
            if (x0 >= 1) {
                const T v = __ldg(rowp - 1);
                hl = (uint8_t)((side.lt(v) ? 1 : 0) | (side.gt(v) ? 2 : 0));
            }
            if (x0 + kTile < nx) {
                const T v = __ldg(rowp + kTile);
                hr = (uint8_t)((side.lt(v) ? 1 : 0) | (side.gt(v) ? 2 : 0));
            }
        }
        s_xh[row] = hl, s_xh[kTile * kTile + row] = hr;
    }
    __syncthreads();
    fractal_finish<T, kThreads>(base, nz, ny, nx, x0, y0, zt, plane, c, s_lg, s_flag, s_cand, s_rows, s_xh, cnt,
                                &cnt[kTileLevels], counts, coarse, (((int64_t)bz + tz0) * gy + by) * gx + bx);
}

template <typename T, int kThreads, int kMinCtas>
__global__ void __launch_bounds__(kThreads, kMinCtas)
k_fractal_tiles(const T* __restrict__ f, int64_t nz, int64_t ny, int64_t nx, int64_t zf0, int64_t tz0, double c,
                unsigned long long* __restrict__ counts, uint8_t* __restrict__ coarse) {
    ballot_tile<T, kThreads>(f, nz, ny, nx, zf0, tz0, c, counts, coarse, (int)blockIdx.x, (int)blockIdx.y, (int)blockIdx.z,
                             (int)gridDim.x, (int)gridDim.y);
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
    const size_t esz = dtype == FAVA_F64 ? 8 : 4;
    if ((nx * esz) % 16 == 0 && (uintptr_t)d_field % 16 == 0 && nx < (int64_t(1) << 31) && ny < (int64_t(1) << 31)) {
        // tensor-map front end: planes of the buffer [zf0, zf1) as a 3-D tensor, NaN outside
        const uint64_t dims[3] = {(uint64_t)nx, (uint64_t)ny, (uint64_t)(zf1 - zf0)};
        const uint64_t strides[2] = {(uint64_t)nx * esz, (uint64_t)nx * ny * esz};
        const uint32_t box_plane[3] = {kTile, kTile, 1}, box_yline[3] = {kTile, 1, kHalo};
        const CUtensorMapDataType dt = dtype == FAVA_F64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
        CUtensorMap tm_plane, tm_yline;
        int rc = ctx_tensor_map(ctx, d_field, dt, 3, dims, strides, box_plane, &tm_plane, true);
        if (rc == FAVA_OK) rc = ctx_tensor_map(ctx, d_field, dt, 3, dims, strides, box_yline, &tm_yline, true);
        if (rc != FAVA_OK) return rc;
        if (dtype == FAVA_F64) {
            auto kern = k_fractal_tiles_tma<double, 1>;
            constexpr size_t dyn = tma_tile_smem<double, 1>();
            FAVA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
            kern<<<grid, kTmaThreads, dyn, st>>>(tm_plane, tm_yline, (const double*)d_field, nz, ny, nx, zf0, tz0, contour,
                                                 (unsigned long long*)d_counts, d_coarse);
        } else {
            auto kern = k_fractal_tiles_tma<float, 1>;
            constexpr size_t dyn = tma_tile_smem<float, 1>();
            FAVA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
            kern<<<grid, kTmaThreads, dyn, st>>>(tm_plane, tm_yline, (const float*)d_field, nz, ny, nx, zf0, tz0, contour,
                                                 (unsigned long long*)d_counts, d_coarse);
        }
    } else if (dtype == FAVA_F64)
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
