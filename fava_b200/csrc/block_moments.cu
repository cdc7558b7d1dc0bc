// K1 block-list front end — Reynolds / Favre plane moments of FLASH block datasets
// [nblocks][nzb][nyb][nxb] (AMR plt/chk files and multi-block uniform files).
//
// Replaces the per-leaf-block NumPy loops of FLASH.reynolds_stress (reference
// fava/mesh/FLASH/_flash.py:1564-1577 pass 1, :1584-1604 pass 2; ~1 ms of Python per block).
//
// Stage 1 (k_block_moments_*): one pass over the leaf blocks; every (leaf, block-plane) "item"
//   gets its 13 pivoted raw moments (same set as plane_moments.cu), taken about the pivot of the
//   FIRST fine bin the plane covers.  Each leaf's cells are read exactly once: 4*s bytes per cell.
// Stage 2 (k_combine_items): one CTA per fine bin walks the bin's item list (CSR built on the host
//   from the leaf table, leaf order preserved), re-expresses each item's moments about the bin's own
//   pivot (exact algebra, a no-op on single-level meshes), weights by vol_frac and sums in a fixed
//   order.  A coarse block plane (scale = 2^(lmax-level) > 1) feeds `scale` fine bins, exactly like
//   the reference's means[jlo:jhi] += ... / `for ii in range(mapping...)` (_flash.py:1576, :1596).
// No floating-point atomics: results are bitwise reproducible.
#include <algorithm>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "common.cuh"

namespace fava {

constexpr int kBT = 256;  // threads per CTA in stage 1
constexpr int kNMb = 13;

// One entry of a fine bin's item list (CSR built on the host): everything stage 2 needs to know about the item
// without touching the leaf table - the chain entry -> leaf descriptor -> pivot was what bound the combine pass.
struct alignas(16) BinEntry {
    int32_t item;  // leaf * nrb + plane
    int32_t pbin;  // the bin whose pivot the item's moments are taken about (first fine bin of the block plane)
    double w;      // vol_frac of the leaf
};
static_assert(sizeof(BinEntry) == 16, "BinEntry is loaded as one 16-byte word");

// Ring path: leaves that cover the same bins at the same weight - equal (ilo, scale, vol_frac) - are sorted next to
// each other on the host and cut into UNITS of a few dozen leaves; a unit's planes are summed in registers across
// its leaves and leave as ONE record per plane (13 sums + the leaf count), so stage 2 sees nleaf / unit-length
// items instead of nleaf.  With one record per (leaf, plane) an 8^3 block costs 104 B per 256 B of plane data to
// write and to read back, scattered (stage 2 alone: 0.16 ms for 2 M records against 0.33 ms for the fields).
struct alignas(16) SortedLeaf {
    int64_t block;  // source block of the leaf
    int64_t pad_;
};
struct alignas(16) LeafUnit {
    int32_t first, count;  // range in the sorted leaf list
    int32_t ilo, scale;    // bins of plane p: ilo + p * scale ... + scale - 1
};
constexpr int kUnitRec = kNMb + 1;  // words per unit record: 13 sums + leaf count

struct Acc13 {
    double m[kNMb];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < kNMb; ++i) m[i] = 0.0;
    }
    __device__ __forceinline__ void add(double r, double x, double y, double z, double c0, double c1, double c2) {
        const double dx = x - c0, dy = y - c1, dz = z - c2;
        const double rx = r * dx, ry = r * dy, rz = r * dz;
        m[0] += r;
        m[1] += dx;
        m[2] += dy;
        m[3] += dz;
        m[4] += rx;
        m[5] += ry;
        m[6] += rz;
        m[7] = fma(rx, dx, m[7]);
        m[8] = fma(rx, dy, m[8]);
        m[9] = fma(rx, dz, m[9]);
        m[10] = fma(ry, dy, m[10]);
        m[11] = fma(ry, dz, m[11]);
        m[12] = fma(rz, dz, m[12]);
    }
};

// ---- stage 1, fast path: power-of-two blocks with nxb*nyb <= 256 <= cells --------------------------
// One CTA per leaf.  Every thread owns ONE plane for the whole block (so it accumulates in registers)
// and walks the block with fully coalesced loads:
//   axis x: element e = t + BT r           -> x = t mod nxb                     (plane = t & (nxb-1))
//   axis y: same walk                      -> y = (t >> lx) mod nyb
//   axis z: plane = t / G, G = BT/nzb lanes share a plane and stride its contiguous cells.
// BT = 256 threads per leaf (leaves of at most 16 KB take k_block_moments_ring below instead).  The BT/nrb partial
// sums of a plane are then added in a fixed (rotated, bank-conflict-free) order.  VEC: 16-byte loads along z.
template <typename T, int AXIS, int BT, bool VEC>
__global__ void __launch_bounds__(BT, 1024 / BT)
    k_block_moments_pow2(const T* __restrict__ rho, const T* __restrict__ ux, const T* __restrict__ uy,
                         const T* __restrict__ uz, const fava_leaf_desc* __restrict__ leaves,
                         const double* __restrict__ piv, int64_t nbins, int lx, int ly, int lz,
                         double* __restrict__ partial) {
    const int t = threadIdx.x;
    const fava_leaf_desc leaf = leaves[blockIdx.x];
    const int cells = 1 << (lx + ly + lz);
    const int nrb = 1 << (AXIS == 0 ? lx : (AXIS == 1 ? ly : lz));
    const int G = BT / nrb;  // threads per plane
    const int64_t base = leaf.block * (int64_t)cells;

    int plane, e0, step, nit;
    if (AXIS == 0) {
        plane = t & (nrb - 1), e0 = t, step = BT, nit = cells / BT;
    } else if (AXIS == 1) {
        plane = (t >> lx) & (nrb - 1), e0 = t, step = BT, nit = cells / BT;
    } else {
        plane = t / G;
        e0 = (plane << (lx + ly)) + (t - plane * G), step = G, nit = (1 << (lx + ly)) / G;
    }
    const int64_t pbin = leaf.ilo + (int64_t)plane * leaf.scale;
    const double c0 = piv[pbin], c1 = piv[nbins + pbin], c2 = piv[2 * nbins + pbin];

    Acc13 acc;
    acc.clear();
    const T* pr = rho + base + e0;
    const T* px = ux + base + e0;
    const T* py = uy + base + e0;
    const T* pz = uz + base + e0;
    int j = 0;
    if (AXIS == 2 && VEC) {
        // along z the G threads of a plane stride its contiguous cells: 16-byte words per thread make a half-warp's
        // request one 256-byte run instead of 64 bytes per plane (16^3 blocks f32: 0.50 -> 0.39 ms, x / y take 0.41)
        constexpr int V = 16 / sizeof(T);
        typedef typename std::conditional<sizeof(T) == 4, float4, double2>::type Vec;
        const int first = (plane << (lx + ly)) + V * (t - plane * G);
        const int nv = nit / V;  // vector loads per field and thread
#pragma unroll 1
        for (int q = 0; q < nv; ++q) {
            const int64_t o = base + first + q * (V * G);
            const Vec a = __ldcs(reinterpret_cast<const Vec*>(rho + o)), b = __ldcs(reinterpret_cast<const Vec*>(ux + o));
            const Vec cc = __ldcs(reinterpret_cast<const Vec*>(uy + o)), d = __ldcs(reinterpret_cast<const Vec*>(uz + o));
            const T* ar = reinterpret_cast<const T*>(&a);
            const T* br = reinterpret_cast<const T*>(&b);
            const T* cr = reinterpret_cast<const T*>(&cc);
            const T* dr = reinterpret_cast<const T*>(&d);
#pragma unroll
            for (int k = 0; k < V; ++k) acc.add((double)ar[k], (double)br[k], (double)cr[k], (double)dr[k], c0, c1, c2);
        }
        j = nit;
    }
    for (; j + 4 <= nit; j += 4) {
        double vr[4], vx[4], vy[4], vz[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int o = (j + u) * step;
            vr[u] = (double)__ldcs(pr + o);
            vx[u] = (double)__ldcs(px + o);
            vy[u] = (double)__ldcs(py + o);
            vz[u] = (double)__ldcs(pz + o);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) acc.add(vr[u], vx[u], vy[u], vz[u], c0, c1, c2);
    }
    for (; j < nit; ++j) {
        const int o = j * step;
        acc.add((double)__ldcs(pr + o), (double)__ldcs(px + o), (double)__ldcs(py + o), (double)__ldcs(pz + o), c0,
                c1, c2);
    }

    __shared__ double sm[kNMb][BT];
#pragma unroll
    for (int m = 0; m < kNMb; ++m) sm[m][t] = acc.m[m];
    __syncthreads();
    for (int q = t; q < nrb * kNMb; q += BT) {
        const int m = q / nrb, p = q - m * nrb;
        double s = 0.0;
        for (int k = 0; k < G; ++k) {
            const int g = (k + p) & (G - 1);  // rotation: lanes of a warp hit distinct banks
            int tt;
            if (AXIS == 0) tt = p + g * nrb;
            else if (AXIS == 1) {
                const int nx = 1 << lx;
                tt = (g & (nx - 1)) + (p << lx) + ((g >> lx) << (lx + ly));
            } else tt = p * G + g;
            s += sm[m][tt];
        }
        partial[((int64_t)blockIdx.x * nrb + p) * kNMb + m] = s;
    }
}

// ---- stage 1, small power-of-two blocks (one leaf <= 16 KB over its four fields, e.g. 8^3) --------------------
// A CTA per leaf cannot keep enough bytes in flight when a leaf is 8 KB: 262144 CTAs of 64 threads, each exposing
// the latency of its own loads, reached 0.45-0.50 of the HBM peak (the same kernel on 16^3 blocks: 0.89).  Here the
// CTAs are persistent (one per SM, up to eight groups) and every 64-thread GROUP streams its leaves (fixed round-robin) through its own ring of
// `stages` shared-memory slots: one thread issues four bulk copies per leaf (`cp.async.bulk`, one per field - a
// block's cells are contiguous) `stages` leaves ahead, the group accumulates the current leaf out of shared memory
// with the plane ownership of the kernel above, and the leaf descriptors / pivots of the next leaves are fetched a
// leaf ahead.  The sums of a (leaf, plane) are formed in a fixed order (along z the walk over a plane's cells is
// rotated per plane against bank conflicts), so results are bitwise repeatable; a leaf's 13 x nrb sums leave as one
// contiguous record written by consecutive threads.
constexpr int kRingMaxGroups = 8;  // 64-thread groups per CTA (one CTA per SM)
constexpr int kRingBT = 64;
constexpr size_t kRingSmem = 220 * 1024;

__device__ __forceinline__ void group_bar(int grp) {  // barrier grp + 1 over the 64 threads of a group
    asm volatile("bar.sync %0, 64;\n" ::"r"(grp + 1) : "memory");
}

constexpr int kRingRow = kRingBT + 1;  // padded row of the per-group sums [13][65]: conflict-free both ways

// CUBE8 = true: 8 x 8 x 8 blocks with every index computation and trip count known at compile time (the leaf loop is
// bound by instruction issue, not by HBM, as soon as the shape is a run-time value).
template <typename T, int AXIS, bool CUBE8>
__global__ void __launch_bounds__(kRingMaxGroups* kRingBT, 1)
    k_block_moments_ring(const T* __restrict__ rho, const T* __restrict__ ux, const T* __restrict__ uy,
                         const T* __restrict__ uz, const SortedLeaf* __restrict__ sorted,
                         const LeafUnit* __restrict__ units, int64_t nunits, const double* __restrict__ piv,
                         int64_t nbins, int lx_, int ly_, int lz_, int stages, double* __restrict__ partial) {
    extern __shared__ __align__(128) unsigned char ring_smem[];
    constexpr int BT = kRingBT;
    const int lx = CUBE8 ? 3 : lx_, ly = CUBE8 ? 3 : ly_, lz = CUBE8 ? 3 : lz_;
    const int grp = threadIdx.x / BT, t = threadIdx.x % BT, ngroups = blockDim.x / BT;
    const int cells = 1 << (lx + ly + lz);
    const unsigned field_bytes = (unsigned)(cells * sizeof(T)), stage_bytes = 4u * field_bytes;
    unsigned char* ring = ring_smem + (size_t)grp * stages * stage_bytes;
    double* sm = reinterpret_cast<double*>(ring_smem + (size_t)ngroups * stages * stage_bytes) + grp * (kNMb * kRingRow);
    uint64_t* bar = reinterpret_cast<uint64_t*>(ring_smem + (size_t)ngroups * stages * stage_bytes +
                                                sizeof(double) * ngroups * kNMb * kRingRow) + grp * stages;
    if (t == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async;\n" ::: "memory");
    }
    __syncthreads();

    const int nrb = 1 << (AXIS == 0 ? lx : (AXIS == 1 ? ly : lz));
    const int G = BT / nrb;  // threads per plane
    int plane, e0, step, nit, rot = 0;
    if (AXIS == 0) {
        plane = t & (nrb - 1), e0 = t, step = BT, nit = cells / BT;
    } else if (AXIS == 1) {
        plane = (t >> lx) & (nrb - 1), e0 = t, step = BT, nit = cells / BT;
    } else {
        plane = t / G;
        e0 = (plane << (lx + ly)) + (t - plane * G), step = G, nit = (1 << (lx + ly)) / G, rot = plane;
    }

    // this group's units: first_unit + i * ustride; its leaves are the leaves of those units, one after the other
    const int64_t first_unit = (int64_t)blockIdx.x * ngroups + grp, ustride = (int64_t)gridDim.x * ngroups;

    // ---- producer side (thread 0 of the group): runs `stages` leaves ahead of the consumers ----
    int64_t p_unit = first_unit;     // unit of the next leaf to request
    int p_leaf = 0, p_count = 0;     // position inside it
    int64_t p_first = 0, p_block = 0;  // p_block: source block of that leaf, fetched one request early
    bool p_more = false;
    int p_slot = 0;
    auto p_fetch = [&]() {  // descriptor of the next leaf to request, if there is one
        p_more = p_unit < nunits;
        if (!p_more) return;
        if (p_leaf == 0) {
            const LeafUnit u = units[p_unit];
            p_first = u.first, p_count = u.count;
        }
        p_block = sorted[p_first + p_leaf].block;
    };
    auto p_issue = [&]() {  // request the leaf described by p_*, then step to the next one
        const int64_t base = p_block * (int64_t)cells;
        unsigned char* dst = ring + (size_t)p_slot * stage_bytes;
        mbar_expect_tx(&bar[p_slot], stage_bytes);
        bulk_load(dst, rho + base, field_bytes, &bar[p_slot]);
        bulk_load(dst + field_bytes, ux + base, field_bytes, &bar[p_slot]);
        bulk_load(dst + 2 * field_bytes, uy + base, field_bytes, &bar[p_slot]);
        bulk_load(dst + 3 * field_bytes, uz + base, field_bytes, &bar[p_slot]);
        if (++p_slot == stages) p_slot = 0;
        if (++p_leaf == p_count) p_leaf = 0, p_unit += ustride;
        p_fetch();
    };
    if (t == 0) {
        p_fetch();
        for (int i = 0; i < stages && p_more; ++i) p_issue();
    }

    // ---- consumers ----
    int s = 0;
    unsigned parity = 0;
    LeafUnit unit_next = first_unit < nunits ? units[first_unit] : LeafUnit{0, 0, 0, 0};
    for (int64_t u = first_unit; u < nunits; u += ustride) {
        const LeafUnit unit = unit_next;
        if (u + ustride < nunits) unit_next = units[u + ustride];  // needed a whole unit from now
        const int64_t pb = unit.ilo + (int64_t)plane * unit.scale;  // the bin whose pivot this plane's sums are about
        const double c0 = piv[pb], c1 = piv[nbins + pb], c2 = piv[2 * nbins + pb];
        Acc13 acc;
        acc.clear();
        for (int i = 0; i < unit.count; ++i) {
            mbar_wait(&bar[s], parity);
            const T* pr = reinterpret_cast<const T*>(ring + (size_t)s * stage_bytes);
            const T* px = pr + cells;
            const T* py = px + cells;
            const T* pz = py + cells;
            if (CUBE8) {
                constexpr int B = sizeof(T) == 8 ? 4 : 8;  // cells held in registers at a time
#pragma unroll
                for (int j0 = 0; j0 < 8; j0 += B) {
                    T vr[B], vx[B], vy[B], vz[B];
#pragma unroll
                    for (int j = 0; j < B; ++j) {
                        const int o = e0 + ((j0 + j + rot) & 7) * step;
                        vr[j] = pr[o], vx[j] = px[o], vy[j] = py[o], vz[j] = pz[o];
                    }
                    if (j0 + B == 8) {
                        group_bar(grp);  // every thread holds its last cells: the slot can be refilled
                        if (t == 0 && p_more) p_issue();
                    }
#pragma unroll
                    for (int j = 0; j < B; ++j) acc.add((double)vr[j], (double)vx[j], (double)vy[j], (double)vz[j], c0, c1, c2);
                }
            } else {
#pragma unroll 4
                for (int j = 0; j < nit; ++j) {
                    const int o = e0 + ((j + rot) & (nit - 1)) * step;
                    acc.add((double)pr[o], (double)px[o], (double)py[o], (double)pz[o], c0, c1, c2);
                }
                group_bar(grp);  // the slot is consumed
                if (t == 0 && p_more) p_issue();
            }
            if (++s == stages) s = 0, parity ^= 1u;
        }
        // the unit's record: [plane][13 sums + leaf count], consecutive threads write consecutive words
#pragma unroll
        for (int m = 0; m < kNMb; ++m) sm[m * kRingRow + t] = acc.m[m];
        group_bar(grp);
        double* out = partial + u * (int64_t)(nrb * kUnitRec);
#pragma unroll 2
        for (int q = t; q < nrb * kUnitRec; q += BT) {
            const int p = q / kUnitRec, m = q - p * kUnitRec;
            double sacc = 0.0;
            if (m == kNMb) sacc = (double)unit.count;
            else {
                const double* row = sm + m * kRingRow;
#pragma unroll 8
                for (int g = 0; g < G; ++g) {  // the G threads that own plane p, in thread order
                    int tt;
                    if (AXIS == 0) tt = p + g * nrb;
                    else if (AXIS == 1) tt = (g & ((1 << lx) - 1)) + (p << lx) + ((g >> lx) << (lx + ly));
                    else tt = p * G + g;
                    sacc += row[tt];
                }
            }
            out[q] = sacc;
        }
        group_bar(grp);  // the sums are read: the next unit may overwrite them
    }
}

// ---- stage 1, generic path: any block shape; one warp per (leaf, plane) ------------------------------
template <typename T>
__global__ void __launch_bounds__(kBT)
    k_block_moments_generic(const T* __restrict__ rho, const T* __restrict__ ux, const T* __restrict__ uy,
                            const T* __restrict__ uz, const fava_leaf_desc* __restrict__ leaves, int64_t nitems,
                            const double* __restrict__ piv, int64_t nbins, int nxb, int nyb, int nzb, int axis,
                            double* __restrict__ partial) {
    const int lane = threadIdx.x & 31;
    const int64_t item = (int64_t)blockIdx.x * (kBT / 32) + (threadIdx.x >> 5);
    if (item >= nitems) return;
    const int nrb = axis == 0 ? nxb : (axis == 1 ? nyb : nzb);
    const int64_t l = item / nrb;
    const int p = (int)(item - l * nrb);
    const fava_leaf_desc leaf = leaves[l];
    const int64_t base = leaf.block * ((int64_t)nxb * nyb * nzb);
    const int64_t pbin = leaf.ilo + (int64_t)p * leaf.scale;
    const double c0 = piv[pbin], c1 = piv[nbins + pbin], c2 = piv[2 * nbins + pbin];
    const int pc = axis == 0 ? nyb * nzb : (axis == 1 ? nxb * nzb : nxb * nyb);
    Acc13 acc;
    acc.clear();
    for (int c = lane; c < pc; c += 32) {
        int64_t e;
        if (axis == 0) e = (int64_t)c * nxb + p;                                  // c = z*nyb + y
        else if (axis == 1) e = ((int64_t)(c / nxb) * nyb + p) * nxb + (c % nxb);  // c = z*nxb + x
        else e = (int64_t)p * pc + c;
        e += base;
        acc.add((double)rho[e], (double)ux[e], (double)uy[e], (double)uz[e], c0, c1, c2);
    }
#pragma unroll
    for (int m = 0; m < kNMb; ++m) {
        const double s = warp_sum_fixed(acc.m[m]);
        if (lane == 0) partial[item * kNMb + m] = s;
    }
}

// ---- pivots: velocity at the first cell of the first block plane (in leaf order) that covers the bin -----------
// piv_src[b] = element index of that cell (resolved on the host with the other tables), -1 for an empty bin.
template <typename T>
__global__ void k_block_pivots(const T* __restrict__ ux, const T* __restrict__ uy, const T* __restrict__ uz,
                               const int64_t* __restrict__ piv_src, int64_t nbins, double* __restrict__ piv) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbins) return;
    double c[3] = {0.0, 0.0, 0.0};
    const int64_t e = piv_src[b];
    if (e >= 0) c[0] = (double)ux[e], c[1] = (double)uy[e], c[2] = (double)uz[e];
    piv[b] = c[0], piv[nbins + b] = c[1], piv[2 * nbins + b] = c[2];
}

// ---- stage 2: items -> fine-bin moments ---------------------------------------------------------------
constexpr int kCT = 128;
// One CTA per fine bin walks the bin's entry list, a thread per entry.  REC = 13: one record per (leaf, plane);
// REC = 14: one record per (unit, plane) whose last word is the number of leaves summed into it.
template <int REC>
__global__ void __launch_bounds__(kCT)
    k_combine_items(const double* __restrict__ partial, const int64_t* __restrict__ off, const BinEntry* __restrict__ ent,
                    const double* __restrict__ piv, int64_t nbins, double plane_cells, double* __restrict__ mom) {
    const int64_t b = blockIdx.x;
    const int t = threadIdx.x;
    const double cb[3] = {piv[b], piv[nbins + b], piv[2 * nbins + b]};
    double acc[FAVA_NMOM];
#pragma unroll
    for (int m = 0; m < FAVA_NMOM; ++m) acc[m] = 0.0;
    for (int64_t k = off[b] + t; k < off[b + 1]; k += kCT) {
        const BinEntry en = ent[k];
        const double* q = partial + (int64_t)en.item * REC;
        double e[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) e[i] = piv[i * nbins + en.pbin] - cb[i];  // c_item - c_bin
        const double s0 = q[0];
        double srd[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) srd[i] = q[4 + i];
        const double vf = en.w;
        const double ncell = REC == kUnitRec ? plane_cells * q[kNMb] : plane_cells;  // cells summed into the record
        acc[0] += vf * s0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            acc[1 + i] += vf * (q[1 + i] + ncell * e[i]);
            acc[4 + i] += vf * (srd[i] + e[i] * s0);
        }
        int kk = 7;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = i; j < 3; ++j, ++kk)
                acc[kk] += vf * (q[kk] + e[i] * srd[j] + e[j] * srd[i] + e[i] * e[j] * s0);
        acc[13] += vf * ncell;
    }
    __shared__ double sm[FAVA_NMOM][kCT / 32];
    const int lane = t & 31, warp = t >> 5;
#pragma unroll
    for (int m = 0; m < FAVA_NMOM; ++m) {
        const double sacc = warp_sum_fixed(acc[m]);
        if (lane == 0) sm[m][warp] = sacc;
    }
    __syncthreads();
    if (t < FAVA_NMOM) {
        double r = sm[t][0];
#pragma unroll
        for (int w = 1; w < kCT / 32; ++w) r += sm[t][w];
        mom[(int64_t)t * nbins + b] = r;
    }
}

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int ilog2_exact(int64_t v) {
    if (v <= 0 || (v & (v - 1))) return -1;
    int l = 0;
    while ((int64_t(1) << l) < v) ++l;
    return l;
}

struct ItemTables {
    fava_leaf_desc* leaves = nullptr;  // item layout: the caller's table
    SortedLeaf* sorted = nullptr;      // unit layout: leaves sorted by (ilo, scale, vol_frac), table order inside a key
    LeafUnit* units = nullptr;
    int64_t nunits = 0;
    int64_t* piv_src = nullptr;  // [nbins] element index of the bin's pivot cell, -1 = empty bin
    int64_t* off = nullptr;      // [nbins + 1] CSR bin -> entries
    BinEntry* ent = nullptr;
};

static inline size_t pad16(size_t b) { return (b + 15) / 16 * 16; }

// Host: CSR bin -> entries, uploaded with the leaf tables into a per-axis context buffer.
//   item layout (unit_len = 0): an entry per (leaf, plane) covering the bin, leaf order preserved inside a bin;
//   unit layout (unit_len > 0): leaves sorted by (ilo, scale, vol_frac) - stable, so table order inside a key - and cut
//     into units of <= unit_len leaves; an entry per (unit, plane) covering the bin, in unit order.
// The tables are cached: a time series over files of one mesh (or repeated calls on one file) passes the same leaf
// table again and again, and building the CSR on the host costs more than the kernels themselves (measured: 2-20 ms
// against 0.4-1.4 ms), so an identical table (memcmp with the kept host copy) re-uses the device tables as they are.
//
// Comparing costs too: 0.7 ms of memcmp at 262144 leaves (8 MB), more than the kernels take (0.4 ms), so a caller
// that keeps its tables immutable passes a non-zero `uid` per table object - equal uid, length and bin count mean
// the same table, and nothing is compared.
static int build_item_tables(fava_ctx* ctx, int axis, const fava_leaf_desc* h_leaves, int64_t nleaf, int nrb,
                             int64_t cells, int64_t plane_stride, int64_t nbins, int unit_len, uint64_t uid,
                             cudaStream_t st, ItemTables* out) {
    const int64_t nitems = nleaf * nrb;
    if (nitems > INT32_MAX) return set_error(FAVA_EINVAL, "block front end: too many block planes");
    // cache key = (nrb, nbins, layout, block cells, raw leaf bytes)
    const int64_t head[4] = {(int64_t)nrb, nbins, (int64_t)unit_len, cells};
    const size_t key_bytes = sizeof(head) + sizeof(fava_leaf_desc) * (size_t)nleaf;
    const int slot = WS_ITEMS0 + axis;
    auto carve = [&](char* tab, int64_t nunits, size_t nent) {  // the layout of the device buffer
        size_t o = 0;
        out->leaves = (fava_leaf_desc*)(tab + o), o += pad16(sizeof(fava_leaf_desc) * (size_t)(unit_len ? 0 : nleaf));
        out->sorted = (SortedLeaf*)(tab + o), o += sizeof(SortedLeaf) * (size_t)(unit_len ? nleaf : 0);
        out->units = (LeafUnit*)(tab + o), o += sizeof(LeafUnit) * (size_t)nunits;
        out->piv_src = (int64_t*)(tab + o), o += pad16(sizeof(int64_t) * (size_t)nbins);
        out->off = (int64_t*)(tab + o), o += pad16(sizeof(int64_t) * ((size_t)nbins + 1));
        out->ent = (BinEntry*)(tab + o), o += sizeof(BinEntry) * std::max<size_t>(nent, 1);
        out->nunits = nunits;
        return o;
    };
    const std::string& have = ctx->item_cache_key[axis];
    const bool same_shape = ctx->ws[slot] && have.size() == key_bytes && memcmp(have.data(), head, sizeof(head)) == 0;
    const bool same_uid = uid != 0 && uid == ctx->item_cache_uid[axis];
    if (same_shape && (same_uid || nleaf == 0 ||
                       memcmp(have.data() + sizeof(head), h_leaves, sizeof(fava_leaf_desc) * (size_t)nleaf) == 0)) {
        ctx->item_cache_uid[axis] = uid;
        carve((char*)ctx->ws[slot], ctx->item_cache_nunits[axis], (size_t)ctx->item_cache_nent[axis]);
        return FAVA_OK;
    }
    std::string key(key_bytes, '\0');
    memcpy(&key[0], head, sizeof(head));
    if (nleaf) memcpy(&key[sizeof(head)], h_leaves, sizeof(fava_leaf_desc) * (size_t)nleaf);
    ctx->item_cache_key[axis].clear();
    ctx->item_cache_uid[axis] = 0;

    std::vector<int64_t> piv_src((size_t)nbins, -1);
    for (int64_t l = 0; l < nleaf; ++l) {
        const fava_leaf_desc& d = h_leaves[l];
        if (d.scale < 1 || d.ilo < 0 || d.ilo + (int64_t)nrb * d.scale > nbins || d.block < 0)
            return set_error(FAVA_EINVAL, "block front end: leaf %lld (block %lld, ilo %lld, scale %d) "
                             "does not fit %lld bins", (long long)l, (long long)d.block, (long long)d.ilo, d.scale,
                             (long long)nbins);
        for (int p = 0; p < nrb; ++p)
            for (int sc = 0; sc < d.scale; ++sc) {
                int64_t& src = piv_src[(size_t)(d.ilo + (int64_t)p * d.scale + sc)];
                if (src < 0) src = d.block * cells + p * plane_stride;  // first covering plane in leaf order
            }
    }

    // what stage 2 sums: (leaf, plane) items or (unit, plane) items
    struct Source {
        int64_t ilo;
        int scale;
        double w;
    };
    std::vector<Source> src;
    std::vector<SortedLeaf> sorted;
    std::vector<LeafUnit> units;
    if (unit_len > 0) {
        std::vector<int32_t> order((size_t)nleaf);
        for (int64_t l = 0; l < nleaf; ++l) order[(size_t)l] = (int32_t)l;
        auto key_less = [&](int32_t a, int32_t b) {
            const fava_leaf_desc &x = h_leaves[a], &y = h_leaves[b];
            if (x.ilo != y.ilo) return x.ilo < y.ilo;
            if (x.scale != y.scale) return x.scale < y.scale;
            return memcmp(&x.vol_frac, &y.vol_frac, sizeof(double)) < 0;  // any total order on the bit pattern
        };
        std::stable_sort(order.begin(), order.end(), key_less);
        sorted.resize((size_t)nleaf);
        for (int64_t i = 0; i < nleaf; ++i) sorted[(size_t)i] = SortedLeaf{h_leaves[order[(size_t)i]].block, 0};
        for (int64_t i = 0; i < nleaf;) {
            int64_t j = i + 1;  // sorted, so "not less than the first" means "same key"
            while (j < nleaf && j - i < unit_len && !key_less(order[(size_t)i], order[(size_t)j])) ++j;
            const fava_leaf_desc& d = h_leaves[order[(size_t)i]];
            units.push_back(LeafUnit{(int32_t)i, (int32_t)(j - i), (int32_t)d.ilo, d.scale});
            src.push_back(Source{d.ilo, d.scale, d.vol_frac});
            i = j;
        }
    } else {
        src.resize((size_t)nleaf);
        for (int64_t l = 0; l < nleaf; ++l) src[(size_t)l] = Source{h_leaves[l].ilo, h_leaves[l].scale, h_leaves[l].vol_frac};
    }
    const int64_t nunits = (int64_t)units.size();
    std::vector<int64_t> off((size_t)nbins + 1, 0);
    for (const Source& d : src)
        for (int64_t b = d.ilo; b < d.ilo + (int64_t)nrb * d.scale; ++b) off[(size_t)b + 1]++;
    for (int64_t b = 0; b < nbins; ++b) off[(size_t)b + 1] += off[(size_t)b];
    const int64_t nent = off[(size_t)nbins];
    std::vector<BinEntry> ent((size_t)std::max<int64_t>(nent, 1));
    {
        std::vector<int64_t> cur(off.begin(), off.end() - 1);
        for (size_t i = 0; i < src.size(); ++i) {
            const Source& d = src[i];
            for (int p = 0; p < nrb; ++p)
                for (int sc = 0; sc < d.scale; ++sc)
                    ent[(size_t)cur[(size_t)(d.ilo + (int64_t)p * d.scale + sc)]++] =
                        BinEntry{(int32_t)((int64_t)i * nrb + p), (int32_t)(d.ilo + (int64_t)p * d.scale), d.w};
        }
    }
    void* tab;
    const size_t total = carve(nullptr, nunits, (size_t)nent);  // sizes only
    int rc = ctx_workspace(ctx, slot, total, &tab);
    if (rc) return rc;
    carve((char*)tab, nunits, (size_t)nent);
    // the vectors are pageable: cudaMemcpyAsync has staged them before it returns
    if (unit_len == 0 && nleaf)
        FAVA_CHECK_CUDA(cudaMemcpyAsync(out->leaves, h_leaves, sizeof(fava_leaf_desc) * nleaf, cudaMemcpyHostToDevice, st));
    if (unit_len > 0 && nleaf) {
        FAVA_CHECK_CUDA(cudaMemcpyAsync(out->sorted, sorted.data(), sizeof(SortedLeaf) * nleaf, cudaMemcpyHostToDevice, st));
        FAVA_CHECK_CUDA(cudaMemcpyAsync(out->units, units.data(), sizeof(LeafUnit) * nunits, cudaMemcpyHostToDevice, st));
    }
    FAVA_CHECK_CUDA(cudaMemcpyAsync(out->piv_src, piv_src.data(), sizeof(int64_t) * nbins, cudaMemcpyHostToDevice, st));
    FAVA_CHECK_CUDA(cudaMemcpyAsync(out->off, off.data(), sizeof(int64_t) * off.size(), cudaMemcpyHostToDevice, st));
    FAVA_CHECK_CUDA(cudaMemcpyAsync(out->ent, ent.data(), sizeof(BinEntry) * ent.size(), cudaMemcpyHostToDevice, st));
    ctx->item_cache_key[axis].swap(key);
    ctx->item_cache_uid[axis] = uid;
    ctx->item_cache_nunits[axis] = nunits;
    ctx->item_cache_nent[axis] = nent;
    return FAVA_OK;
}

template <typename T>
static int run_blocks(fava_ctx* ctx, const T* rho, const T* ux, const T* uy, const T* uz, int64_t nzb, int64_t nyb,
                      int64_t nxb, int axis, const fava_leaf_desc* h_leaves, int64_t nleaf, uint64_t uid, int64_t nbins,
                      double* mom, double* piv, cudaStream_t st) {
    const int nrb = (int)(axis == 0 ? nxb : (axis == 1 ? nyb : nzb));
    const int64_t nitems = nleaf * nrb;
    const int64_t cells = nxb * nyb * nzb;
    const int64_t plane_stride = axis == 0 ? 1 : (axis == 1 ? nxb : nxb * nyb);
    const int lx = ilog2_exact(nxb), ly = ilog2_exact(nyb), lz = ilog2_exact(nzb);
    auto fits = [&](int64_t bt) {
        return lx >= 0 && ly >= 0 && lz >= 0 && nxb * nyb <= bt && cells >= bt && nrb <= bt &&
               (axis != 2 || (nxb * nyb) >= bt / nzb);
    };
    // small blocks (a leaf <= 16 KB over its four fields) stream through the leaf rings, summed per unit of leaves
    const size_t stage_bytes = 4 * sizeof(T) * (size_t)cells;
    const bool aligned16 = ((uintptr_t)rho | (uintptr_t)ux | (uintptr_t)uy | (uintptr_t)uz) % 16 == 0;
    const bool ring = fits(kRingBT) && stage_bytes <= 16384 && stage_bytes >= 1024 && aligned16;
    const size_t sums_bytes = sizeof(double) * kNMb * kRingRow;
    // one CTA per SM: as many groups as fit with three slots each, then as many slots as fit (<= 8)
    const int groups = ring ? (int)std::min<size_t>(kRingMaxGroups, kRingSmem / (3 * stage_bytes + sums_bytes)) : 0;
    // unit length: >= 8 units per group when the mesh is large enough, at most 64 leaves (a fixed function of the
    // table and the device: the summation order, hence the result bits, do not change from call to call)
    const int unit_len = ring ? (int)std::max<int64_t>(4, std::min<int64_t>(64, nleaf / (8 * (int64_t)ctx->num_sms * groups))) : 0;

    ItemTables tb;
    int rc = build_item_tables(ctx, axis, h_leaves, nleaf, nrb, cells, plane_stride, nbins, unit_len, uid, st, &tb);
    if (rc) return rc;
    const int64_t nrec = ring ? tb.nunits * nrb * kUnitRec : nitems * kNMb;
    void* ws;
    rc = ctx_workspace(ctx, WS_PARTIALS, sizeof(double) * (size_t)std::max<int64_t>(nrec, 1), &ws);
    if (rc) return rc;
    double* partial = (double*)ws;

    k_block_pivots<T><<<(unsigned)cdiv(nbins, 128), 128, 0, st>>>(ux, uy, uz, tb.piv_src, nbins, piv);
    FAVA_LAUNCHED();
    if (nleaf) {
        const unsigned grid = (unsigned)nleaf;
        // z planes read as 16-byte words: aligned fields, and a plane's cells a multiple of what its threads take per load
        const bool vec2 = aligned16 && fits(kBT) && ((nxb * nyb) % ((16 / (int)sizeof(T)) * (kBT / nzb))) == 0;
#define FAVA_LAUNCH_POW2(AX, BTV, VECV) \
    k_block_moments_pow2<T, AX, BTV, VECV><<<grid, BTV, 0, st>>>(rho, ux, uy, uz, tb.leaves, piv, nbins, lx, ly, lz, partial)
        if (ring) {
            const int stages = (int)std::min<size_t>(8, (kRingSmem / groups - sums_bytes) / stage_bytes);
            const size_t dyn = (size_t)groups * (stages * stage_bytes + sums_bytes + sizeof(uint64_t) * stages);
            const unsigned rgrid = (unsigned)std::min<int64_t>(cdiv(tb.nunits, groups), (int64_t)ctx->num_sms);
            const bool cube8 = lx == 3 && ly == 3 && lz == 3;
#define FAVA_LAUNCH_RING(AX)                                                                                         \
    do {                                                                                                             \
        auto kern = cube8 ? k_block_moments_ring<T, AX, true> : k_block_moments_ring<T, AX, false>;                  \
        FAVA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));          \
        kern<<<rgrid, groups * kRingBT, dyn, st>>>(rho, ux, uy, uz, tb.sorted, tb.units, tb.nunits, piv, nbins, lx,  \
                                                   ly, lz, stages, partial);                                         \
    } while (0)
            if (axis == 0) FAVA_LAUNCH_RING(0);
            else if (axis == 1) FAVA_LAUNCH_RING(1);
            else FAVA_LAUNCH_RING(2);
#undef FAVA_LAUNCH_RING
        } else if (fits(kBT)) {
            if (axis == 0) FAVA_LAUNCH_POW2(0, kBT, false);
            else if (axis == 1) FAVA_LAUNCH_POW2(1, kBT, false);
            else if (vec2) FAVA_LAUNCH_POW2(2, kBT, true);
            else FAVA_LAUNCH_POW2(2, kBT, false);
#undef FAVA_LAUNCH_POW2
        } else {
            k_block_moments_generic<T><<<(unsigned)cdiv(nitems, kBT / 32), kBT, 0, st>>>(
                rho, ux, uy, uz, tb.leaves, nitems, piv, nbins, (int)nxb, (int)nyb, (int)nzb, axis, partial);
        }
        FAVA_LAUNCHED();
    }
    const double plane_cells = (double)(cells / nrb);
    if (ring)
        k_combine_items<kUnitRec><<<(unsigned)nbins, kCT, 0, st>>>(partial, tb.off, tb.ent, piv, nbins, plane_cells, mom);
    else
        k_combine_items<kNMb><<<(unsigned)nbins, kCT, 0, st>>>(partial, tb.off, tb.ent, piv, nbins, plane_cells, mom);
    FAVA_LAUNCHED();
    return FAVA_OK;
}


// ---- single-field plane integral of a block dataset (reference slice_integral, _flash.py:1451-1504) ---
template <typename T>
__global__ void __launch_bounds__(kBT)
    k_block_sum_generic(const T* __restrict__ f, const fava_leaf_desc* __restrict__ leaves, int64_t nitems, int nxb,
                        int nyb, int nzb, int axis, double* __restrict__ partial) {
    const int lane = threadIdx.x & 31;
    const int64_t item = (int64_t)blockIdx.x * (kBT / 32) + (threadIdx.x >> 5);
    if (item >= nitems) return;
    const int nrb = axis == 0 ? nxb : (axis == 1 ? nyb : nzb);
    const int64_t l = item / nrb;
    const int p = (int)(item - l * nrb);
    const int64_t base = leaves[l].block * ((int64_t)nxb * nyb * nzb);
    const int pc = axis == 0 ? nyb * nzb : (axis == 1 ? nxb * nzb : nxb * nyb);
    double acc = 0.0;
    for (int c = lane; c < pc; c += 32) {
        int64_t e;
        if (axis == 0) e = (int64_t)c * nxb + p;
        else if (axis == 1) e = ((int64_t)(c / nxb) * nyb + p) * nxb + (c % nxb);
        else e = (int64_t)p * pc + c;
        acc += (double)f[base + e];
    }
    acc = warp_sum_fixed(acc);
    if (lane == 0) partial[item] = acc;
}

__global__ void __launch_bounds__(kCT)
    k_combine_sums(const double* __restrict__ partial, const int64_t* __restrict__ off, const BinEntry* __restrict__ ent,
                   double* __restrict__ out) {
    const int64_t b = blockIdx.x;
    const int t = threadIdx.x;
    double acc = 0.0;
    for (int64_t k = off[b] + t; k < off[b + 1]; k += kCT) {
        const BinEntry e = ent[k];
        acc += e.w * partial[e.item];
    }
    __shared__ double sm[kCT / 32];
    const double s = warp_sum_fixed(acc);
    if ((t & 31) == 0) sm[t >> 5] = s;
    __syncthreads();
    if (t == 0) {
        double r = sm[0];
#pragma unroll
        for (int w = 1; w < kCT / 32; ++w) r += sm[w];
        out[b] = r;
    }
}

template <typename T>
static int run_block_sum(fava_ctx* ctx, const T* f, int64_t nzb, int64_t nyb, int64_t nxb, int axis,
                         const fava_leaf_desc* h_leaves, int64_t nleaf, uint64_t uid, int64_t nbins, double* out,
                         cudaStream_t st) {
    const int nrb = (int)(axis == 0 ? nxb : (axis == 1 ? nyb : nzb));
    const int64_t nitems = nleaf * nrb;
    ItemTables tb;
    // same key and same tables as a per-leaf moments call on this mesh (pivot sources included, unused here)
    const int64_t plane_stride = axis == 0 ? 1 : (axis == 1 ? nxb : nxb * nyb);
    int rc = build_item_tables(ctx, axis, h_leaves, nleaf, nrb, nxb * nyb * nzb, plane_stride, nbins, 0, uid, st, &tb);
    if (rc) return rc;
    void* ws;
    rc = ctx_workspace(ctx, WS_PARTIALS, sizeof(double) * (size_t)std::max<int64_t>(nitems, 1), &ws);
    if (rc) return rc;
    if (nitems) {
        k_block_sum_generic<T><<<(unsigned)cdiv(nitems, kBT / 32), kBT, 0, st>>>(f, tb.leaves, nitems, (int)nxb, (int)nyb,
                                                                               (int)nzb, axis, (double*)ws);
        FAVA_LAUNCHED();
    }
    k_combine_sums<<<(unsigned)nbins, kCT, 0, st>>>((const double*)ws, tb.off, tb.ent, out);
    FAVA_LAUNCHED();
    return FAVA_OK;
}

}  // namespace fava

using namespace fava;

extern "C" int fava_plane_moments_blocks_uid(fava_ctx* ctx, const void* d_rho, const void* d_ux, const void* d_uy,
                                             const void* d_uz, int dtype, int64_t nzb, int64_t nyb, int64_t nxb,
                                             int axis, const fava_leaf_desc* h_leaves, int64_t nleaf,
                                             uint64_t table_uid, int64_t nbins, double* d_moments, double* d_pivots,
                                             void* stream) {
    FAVA_REQUIRE(ctx && d_moments && d_pivots, "fava_plane_moments_blocks: NULL argument");
    FAVA_REQUIRE(nleaf >= 0 && (nleaf == 0 || (h_leaves && d_rho && d_ux && d_uy && d_uz)),
                 "fava_plane_moments_blocks: NULL field or leaf table");
    FAVA_REQUIRE(nzb > 0 && nyb > 0 && nxb > 0 && nzb * nyb * nxb < (int64_t(1) << 30),
                 "fava_plane_moments_blocks: bad block shape %lldx%lldx%lld", (long long)nzb, (long long)nyb,
                 (long long)nxb);
    FAVA_REQUIRE(axis >= 0 && axis <= 2, "fava_plane_moments_blocks: axis %d not in 0..2", axis);
    FAVA_REQUIRE(dtype == FAVA_F32 || dtype == FAVA_F64, "fava_plane_moments_blocks: bad dtype %d", dtype);
    FAVA_REQUIRE(nbins > 0 && nbins < (int64_t(1) << 31), "fava_plane_moments_blocks: bad bin count %lld", (long long)nbins);
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == FAVA_F64)
        return run_blocks<double>(ctx, (const double*)d_rho, (const double*)d_ux, (const double*)d_uy,
                                  (const double*)d_uz, nzb, nyb, nxb, axis, h_leaves, nleaf, table_uid, nbins, d_moments,
                                  d_pivots, st);
    return run_blocks<float>(ctx, (const float*)d_rho, (const float*)d_ux, (const float*)d_uy, (const float*)d_uz, nzb,
                             nyb, nxb, axis, h_leaves, nleaf, table_uid, nbins, d_moments, d_pivots, st);
}

extern "C" int fava_plane_moments_blocks(fava_ctx* ctx, const void* d_rho, const void* d_ux, const void* d_uy,
                                         const void* d_uz, int dtype, int64_t nzb, int64_t nyb, int64_t nxb,
                                         int axis, const fava_leaf_desc* h_leaves, int64_t nleaf, int64_t nbins,
                                         double* d_moments, double* d_pivots, void* stream) {
    return fava_plane_moments_blocks_uid(ctx, d_rho, d_ux, d_uy, d_uz, dtype, nzb, nyb, nxb, axis, h_leaves, nleaf, 0,
                                         nbins, d_moments, d_pivots, stream);
}

extern "C" int fava_plane_sum_blocks_uid(fava_ctx* ctx, const void* d_field, int dtype, int64_t nzb, int64_t nyb,
                                         int64_t nxb, int axis, const fava_leaf_desc* h_leaves, int64_t nleaf,
                                         uint64_t table_uid, int64_t nbins, double* d_out, void* stream) {
    FAVA_REQUIRE(ctx && d_out, "fava_plane_sum_blocks: NULL argument");
    FAVA_REQUIRE(nleaf >= 0 && (nleaf == 0 || (h_leaves && d_field)), "fava_plane_sum_blocks: NULL field or leaf table");
    FAVA_REQUIRE(nzb > 0 && nyb > 0 && nxb > 0 && nzb * nyb * nxb < (int64_t(1) << 30),
                 "fava_plane_sum_blocks: bad block shape");
    FAVA_REQUIRE(axis >= 0 && axis <= 2, "fava_plane_sum_blocks: axis %d not in 0..2", axis);
    FAVA_REQUIRE(dtype == FAVA_F32 || dtype == FAVA_F64, "fava_plane_sum_blocks: bad dtype %d", dtype);
    FAVA_REQUIRE(nbins > 0 && nbins < (int64_t(1) << 31), "fava_plane_sum_blocks: bad bin count");
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == FAVA_F64)
        return run_block_sum<double>(ctx, (const double*)d_field, nzb, nyb, nxb, axis, h_leaves, nleaf, table_uid, nbins, d_out, st);
    return run_block_sum<float>(ctx, (const float*)d_field, nzb, nyb, nxb, axis, h_leaves, nleaf, table_uid, nbins, d_out, st);
}

extern "C" int fava_plane_sum_blocks(fava_ctx* ctx, const void* d_field, int dtype, int64_t nzb, int64_t nyb,
                                     int64_t nxb, int axis, const fava_leaf_desc* h_leaves, int64_t nleaf,
                                     int64_t nbins, double* d_out, void* stream) {
    return fava_plane_sum_blocks_uid(ctx, d_field, dtype, nzb, nyb, nxb, axis, h_leaves, nleaf, 0, nbins, d_out, stream);
}
