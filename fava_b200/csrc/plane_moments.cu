// K1/K2 — Reynolds / Favre plane statistics for sm_100a.
//
// Replaces the two hot loops of FLASH.reynolds_stress (reference fava/mesh/FLASH/_flash.py:1564-1577
// plane means, :1584-1604 stresses).  One streaming pass reads rho,ux,uy,uz exactly once
// (32 B/cell fp64, 16 B/cell f32) and produces 13 pivoted raw moments per plane
//     S0 = sum rho, Sd_i = sum d_i, Srd_i = sum rho d_i, Srdd_ij = sum rho d_i d_j,  d_i = u_i - c_i
// from which means, <rho u'_i u'_j> (Reynolds, volume-averaged means as in the reference) and the
// Favre quantities follow algebraically.  The pivot c_i (first cell of the plane)
// keeps the single pass within ~1e-14 of the reference's two-pass arithmetic.
//
// Determinism: level 1 = per-thread sequential accumulation + fixed-pattern warp/block reduction into
// a per-CTA partial; level 2 = k_reduce_partials sums the partials of a bin in ascending chunk order.
// No floating-point atomics anywhere.
//
// HBM-bound (19 DFMA-class ops per 32 B): the design goal is bytes in flight — 128-bit streaming
// loads (ld.global.cs), U rows unrolled so each thread keeps 4*U independent 16 B requests
// outstanding, two 256-thread CTAs per SM, grid of many equal work items (>= 16 waves) so the tail
// wave is < 5 %.
#include "common.cuh"

namespace fava {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kNM = FAVA_NMOM - 1;  // 13 accumulated moments; row 13 (W) is analytic for dense input

// NM = 13: the pivoted moment set; NM = 1: a plain plane sum of the first field (slice_integral).
template <int NM>
struct Acc {
    double m[NM];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < NM; ++i) m[i] = 0.0;
    }
    __device__ __forceinline__ void add(double r, double x, double y, double z, double c0, double c1,
                                        double c2) {
        if constexpr (NM == 1) {
            m[0] += r;
            return;
        }
        const double dx = x - c0, dy = y - c1, dz = z - c2;
        const double rx = r * dx, ry = r * dy, rz = r * dz;
        if constexpr (NM > 1) {
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
    }
};

// loads the velocity operands only when they are used
template <typename T, int V, int NM>
__device__ __forceinline__ void load4(const T* rho, const T* ux, const T* uy, const T* uz, int64_t off, double (&vr)[V],
                                      double (&vx)[V], double (&vy)[V], double (&vz)[V]) {
    VecLoad<T, V>::ld(rho + off, vr);
    if constexpr (NM > 1) {
        VecLoad<T, V>::ld(ux + off, vx);
        VecLoad<T, V>::ld(uy + off, vy);
        VecLoad<T, V>::ld(uz + off, vz);
    } else {
#pragma unroll
        for (int v = 0; v < V; ++v) vx[v] = vy[v] = vz[v] = 0.0;
    }
}

// ---- axis 0 (x, the fastest index): column sums -----------------------------------------------
// A warp owns a 32*V-column strip, a CTA's 8 warps take 8 consecutive rows per step; every thread
// keeps 13*V accumulators for its V columns.  grid = (column strips, row chunks).
template <typename T, int V, int U, int NM>
__global__ void __launch_bounds__(kThreads, 2)
    k_moments_cols(const T* __restrict__ rho, const T* __restrict__ ux, const T* __restrict__ uy,
                   const T* __restrict__ uz, int64_t nrows, int64_t nx, const double* __restrict__ piv,
                   double* __restrict__ partial, int64_t rows_per_chunk) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t x0 = ((int64_t)blockIdx.x * 32 + lane) * V;
    const bool active = x0 < nx;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk;
    const int64_t r1 = min(nrows, r0 + rows_per_chunk);

    Acc<NM> acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v].clear();

    if (active) {
        double c[3][V];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int v = 0; v < V; ++v) c[i][v] = NM > 1 ? piv[i * nx + x0 + v] : 0.0;

        int64_t r = r0 + warp;
        for (; r + (int64_t)(U - 1) * kWarps < r1; r += (int64_t)U * kWarps) {
            double vr[U][V], vx[U][V], vy[U][V], vz[U][V];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t off = (r + (int64_t)u * kWarps) * nx + x0;
                load4<T, V, NM>(rho, ux, uy, uz, off, vr[u], vx[u], vy[u], vz[u]);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
#pragma unroll
                for (int v = 0; v < V; ++v)
                    acc[v].add(vr[u][v], vx[u][v], vy[u][v], vz[u][v], c[0][v], c[1][v], c[2][v]);
            }
        }
        for (; r < r1; r += kWarps) {
            double vr[V], vx[V], vy[V], vz[V];
            const int64_t off = r * nx + x0;
            load4<T, V, NM>(rho, ux, uy, uz, off, vr, vx, vy, vz);
#pragma unroll
            for (int v = 0; v < V; ++v) acc[v].add(vr[v], vx[v], vy[v], vz[v], c[0][v], c[1][v], c[2][v]);
        }
    }

    // fixed-order sum over the CTA's 8 warps, one moment at a time through 2 KB * V of smem
    __shared__ double sm[kWarps][32 * V];
    double* out = partial + (int64_t)blockIdx.y * NM * nx;
#pragma unroll
    for (int m = 0; m < NM; ++m) {
#pragma unroll
        for (int v = 0; v < V; ++v) sm[warp][lane * V + v] = acc[v].m[m];
        __syncthreads();
        if (warp == 0 && active) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
                double s = sm[0][lane * V + v];
#pragma unroll
                for (int w = 1; w < kWarps; ++w) s += sm[w][lane * V + v];
                out[(int64_t)m * nx + x0 + v] = s;
            }
        }
        __syncthreads();
    }
}

// ---- axis 1 / 2: every row of a bin goes to the same bin ---------------------------------------
// A bin is `rows_per_bin` rows of `row_len` contiguous elements: row j of bin b starts at
// b*bin_stride + j*row_stride (axis y: bin_stride=nx, row_stride=ny*nx; axis z: the plane is
// contiguous).  grid = (bins, row chunks); a CTA streams whole rows (4 KB per step for fp64) with
// 13 accumulators per thread, then block-reduces in a fixed pattern.
template <typename T, int V, int U, int NM>
__global__ void __launch_bounds__(kThreads, 2)
    k_moments_rows(const T* __restrict__ rho, const T* __restrict__ ux, const T* __restrict__ uy,
                   const T* __restrict__ uz, int64_t row_len, int64_t bin_stride, int64_t row_stride,
                   int64_t rows_per_bin, int64_t rows_per_chunk, const double* __restrict__ piv,
                   int64_t nbins, double* __restrict__ partial, int lanes_x) {
    const int64_t bin = blockIdx.x;
    const int tx = threadIdx.x % lanes_x, ty = threadIdx.x / lanes_x;
    const int rpi = kThreads / lanes_x;
    const int64_t j0 = (int64_t)blockIdx.y * rows_per_chunk;
    const int64_t j1 = min(rows_per_bin, j0 + rows_per_chunk);
    const double c0 = NM > 1 ? piv[bin] : 0.0, c1 = NM > 1 ? piv[nbins + bin] : 0.0,
                 c2 = NM > 1 ? piv[2 * nbins + bin] : 0.0;
    const int64_t base = bin * bin_stride;

    Acc<NM> acc;
    acc.clear();

    for (int64_t xv = (int64_t)tx * V; xv < row_len; xv += (int64_t)lanes_x * V) {
        int64_t j = j0 + ty;
        for (; j + (int64_t)(U - 1) * rpi < j1; j += (int64_t)U * rpi) {
            double vr[U][V], vx[U][V], vy[U][V], vz[U][V];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t off = base + (j + (int64_t)u * rpi) * row_stride + xv;
                load4<T, V, NM>(rho, ux, uy, uz, off, vr[u], vx[u], vy[u], vz[u]);
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int v = 0; v < V; ++v) acc.add(vr[u][v], vx[u][v], vy[u][v], vz[u][v], c0, c1, c2);
        }
        for (; j < j1; j += rpi) {
            double vr[V], vx[V], vy[V], vz[V];
            const int64_t off = base + j * row_stride + xv;
            load4<T, V, NM>(rho, ux, uy, uz, off, vr, vx, vy, vz);
#pragma unroll
            for (int v = 0; v < V; ++v) acc.add(vr[v], vx[v], vy[v], vz[v], c0, c1, c2);
        }
    }

    __shared__ double sm[NM][kWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int m = 0; m < NM; ++m) {
        const double s = warp_sum_fixed(acc.m[m]);
        if (lane == 0) sm[m][warp] = s;
    }
    __syncthreads();
    if (threadIdx.x < NM) {
        double s = sm[threadIdx.x][0];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) s += sm[threadIdx.x][w];
        partial[((int64_t)blockIdx.y * NM + threadIdx.x) * nbins + bin] = s;
    }
}

// ---- level 2: partials -> moments --------------------------------------------------------------
// nm = moments held per chunk; nrows = rows of `mom` (nm, or nm+1 when the analytic W row is appended)
__global__ void k_reduce_partials(const double* __restrict__ partial, int nchunk, int64_t nbins, int nm, int nrows,
                                  double* __restrict__ mom, int accumulate, double cells_per_bin) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)nrows * nbins) return;
    const int64_t m = idx / nbins, b = idx - m * nbins;
    double s;
    if (m < nm) {
        s = 0.0;
        for (int c = 0; c < nchunk; ++c) s += partial[((int64_t)c * nm + m) * nbins + b];
    } else {
        s = cells_per_bin;
    }
    mom[idx] = accumulate ? mom[idx] + s : s;
}

// Fused x+z pass: with one chunk per z-plane, the column partials of plane z (13 moments per x, about the
// x-pivots) also determine the z-bin moments: re-express each about the plane's own pivot (exact algebra, as
// k_repivot) and add over x in a fixed order.  One CTA per plane.
__global__ void __launch_bounds__(256)
    k_partials_to_planes(const double* __restrict__ partial, int64_t nx, int64_t nz, double cells_per_column,
                         const double* __restrict__ piv_x, const double* __restrict__ piv_z,
                         double* __restrict__ mom_z, double cells_per_plane) {
    const int64_t z = blockIdx.x;
    const double* P = partial + z * kNM * nx;
    const double cz[3] = {piv_z[z], piv_z[nz + z], piv_z[2 * nz + z]};
    double acc[kNM];
#pragma unroll
    for (int m = 0; m < kNM; ++m) acc[m] = 0.0;
    for (int64_t x = threadIdx.x; x < nx; x += blockDim.x) {
        double q[kNM];
#pragma unroll
        for (int m = 0; m < kNM; ++m) q[m] = P[(int64_t)m * nx + x];
        double e[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) e[i] = piv_x[i * nx + x] - cz[i];
        acc[0] += q[0];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            acc[1 + i] += q[1 + i] + cells_per_column * e[i];
            acc[4 + i] += q[4 + i] + e[i] * q[0];
        }
        int k = 7;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = i; j < 3; ++j, ++k) acc[k] += q[k] + e[i] * q[4 + j] + e[j] * q[4 + i] + e[i] * e[j] * q[0];
    }
    __shared__ double sm[kNM][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int m = 0; m < kNM; ++m) {
        const double v = warp_sum_fixed(acc[m]);
        if (lane == 0) sm[m][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < kNM) {
        double v = sm[threadIdx.x][0];
#pragma unroll
        for (int w = 1; w < 8; ++w) v += sm[threadIdx.x][w];
        mom_z[(int64_t)threadIdx.x * nz + z] = v;
    }
    if (threadIdx.x == kNM) mom_z[(int64_t)kNM * nz + z] = cells_per_plane;
}

template <typename T>
__global__ void k_plane_pivots(const T* __restrict__ ux, const T* __restrict__ uy, const T* __restrict__ uz,
                               int64_t stride, int64_t nbins, double* __restrict__ piv) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbins) return;
    piv[b] = (double)ux[b * stride];
    piv[nbins + b] = (double)uy[b * stride];
    piv[2 * nbins + b] = (double)uz[b * stride];
}

// Moments about c_old -> moments about c_new.  With e = c_old - c_new, d' = d + e:
//   Sd' = Sd + n e,  Srd' = Srd + e S0,  Srdd'_ij = Srdd_ij + e_i Srd_j + e_j Srd_i + e_i e_j S0
// where n = W (cell count, or weight sum for weighted moments).
__global__ void k_repivot(double* __restrict__ mom, const double* __restrict__ pold,
                          const double* __restrict__ pnew, int64_t nbins) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbins) return;
    double e[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) e[i] = pold[i * nbins + b] - pnew[i * nbins + b];
    const double s0 = mom[b], w = mom[13 * nbins + b];
    double srd[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) srd[i] = mom[(4 + i) * nbins + b];
    int k = 7;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = i; j < 3; ++j, ++k)
            mom[k * nbins + b] += e[i] * srd[j] + e[j] * srd[i] + e[i] * e[j] * s0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        mom[(1 + i) * nbins + b] += w * e[i];
        mom[(4 + i) * nbins + b] = srd[i] + e[i] * s0;
    }
}

// Moments -> profiles.  With all sums scaled by `weight`:
//   <q>      = sum(vf q)/LV                                          (_flash.py:1579-1582)
//   m_i      = <u_i> - c_i = (Sd_i + c_i (W - LV)) / LV
//   R_ij     = (Srdd_ij - m_i Srd_j - m_j Srd_i + m_i m_j S0) / LV   (_flash.py:1597-1609)
//   u~_i     = c_i + Srd_i/S0 ;  F_ij = (Srdd_ij - Srd_i Srd_j / S0) / LV
__global__ void k_finalize(const double* __restrict__ mom, const double* __restrict__ piv, int64_t nbins,
                           double weight, double lv, double* __restrict__ means, double* __restrict__ rey,
                           double* __restrict__ fmeans, double* __restrict__ favre) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbins) return;
    const double s0 = mom[b] * weight;
    const double w = mom[13 * nbins + b] * weight;
    double c[3], sd[3], srd[3], m[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        c[i] = piv[i * nbins + b];
        sd[i] = mom[(1 + i) * nbins + b] * weight;
        srd[i] = mom[(4 + i) * nbins + b] * weight;
        m[i] = (sd[i] + c[i] * (w - lv)) / lv;
    }
    if (means) {
        means[b] = s0 / lv;
#pragma unroll
        for (int i = 0; i < 3; ++i) means[(1 + i) * nbins + b] = c[i] + m[i];
    }
    if (fmeans) {
#pragma unroll
        for (int i = 0; i < 3; ++i) fmeans[i * nbins + b] = c[i] + srd[i] / s0;
    }
    int k = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = i; j < 3; ++j, ++k) {
            const double srdd = mom[(7 + k) * nbins + b] * weight;
            if (rey) rey[k * nbins + b] = (srdd - m[i] * srd[j] - m[j] * srd[i] + m[i] * m[j] * s0) / lv;
            if (favre) favre[k * nbins + b] = (srdd - srd[i] * srd[j] / s0) / lv;
        }
}


// ---- all three axes from ONE read ---------------------------------------------------------------------------
// The x/z pass above and the y pass read the snapshot once each (64 B/cell for the three profile sets).  This
// kernel reads it once (32 B/cell): a cell's 13 terms are accumulated twice out of shared memory - once by the
// thread that owns its COLUMN (x bins; the z bins follow from the per-plane column partials exactly as in the x/z
// pass) and once by the warp that owns its ROW (y bins).
//   work item = (z plane, strip of 256 columns); a persistent 512-thread CTA walks the item in tiles of 8 rows x 256
//   columns x 4 fields, which arrive through 2-D tensor maps (cp.async.bulk.tensor.2d, one box per field) into a
//   3-stage ring while earlier tiles are consumed; a 17th warp is the producer: it re-fills a stage as soon as the 16
//   consumer warps have released it (full / empty mbarriers, no CTA-wide barrier in the loop, so the two groups drift);
//   group A (256 threads): thread = column; 13 accumulators about the column's x pivot, kept in registers over the 128
//   tiles of the item, one store of [13][256] partials per item -> k_reduce_partials / k_partials_to_planes;
//   group B (8 warps): warp = row; lane l takes columns l, l+32, ... about the row's y pivot, then the 13 sums of the warp
//   are reduced by a halving exchange (16 shuffle pairs instead of 65 for 13 butterflies: at every step a lane keeps half
//   of its values and sends the other half) and lanes 0,2,..,24 store one 128-byte record per (z, strip, row)
//   -> k_reduce_rows.  Fixed orders throughout: bitwise reproducible.
// fp64 pipe ~55 %, shared-memory reads 64 B/cell, shuffles 4/cell: under the HBM time of 32 B/cell.
constexpr int kXyzCols = 256, kXyzRows = 8, kXyzStages = 3, kXyzConsumers = 512, kXyzThreads = kXyzConsumers + 32;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

// 13 (padded to 16) per-lane values -> lane l ends with the warp total of value (l >> 1) & 15
__device__ __forceinline__ double warp_halving_reduce16(double (&v)[16], int lane) {
#pragma unroll
    for (int half = 8, mask = 16; half >= 1; half >>= 1, mask >>= 1) {
        const bool up = (lane & mask) != 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < half) {
                const double send = up ? v[i] : v[i + half];
                const double keep = up ? v[i + half] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
            }
        }
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

template <typename T>
__global__ void __launch_bounds__(kXyzThreads, 1)
    k_moments_xyz(const __grid_constant__ CUtensorMap tm_rho, const __grid_constant__ CUtensorMap tm_ux,
                  const __grid_constant__ CUtensorMap tm_uy, const __grid_constant__ CUtensorMap tm_uz, int nz, int ny, int nx,
                  const double* __restrict__ piv_x, const double* __restrict__ piv_y, double* __restrict__ xpartial,
                  double* __restrict__ ypartial, unsigned long long* __restrict__ item_counter) {
    constexpr int TILE = kXyzRows * kXyzCols;  // cells of one field in a tile
    extern __shared__ __align__(1024) unsigned char xyz_smem[];
    T* ring = reinterpret_cast<T*>(xyz_smem);  // [stages][4][8][256]
    uint64_t* full = reinterpret_cast<uint64_t*>(xyz_smem + sizeof(T) * kXyzStages * 4 * TILE);
    uint64_t* empty = full + kXyzStages;
    long long* info = reinterpret_cast<long long*>(empty + kXyzStages);  // per stage: item * tpi + tile, -1 = no more work
    const int tid = threadIdx.x, lane = tid & 31;
    const bool colgroup = tid < kXyzCols;
    const int row = (tid - kXyzCols) >> 5;  // group B: the warp's row inside a tile
    const int nstrips = nx / kXyzCols, tpi = ny / kXyzRows;
    const int64_t nitems = (int64_t)nz * nstrips;

    if (tid == 0) {
        for (int s = 0; s < kXyzStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kXyzConsumers / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async;\n" ::: "memory");
    }
    __syncthreads();

    if (tid >= kXyzConsumers) {  // ---- producer warp: one lane feeds the ring ----
        // items come from a global counter: a CTA that starts late (another kernel still holds its SM) takes fewer
        if (lane == 0) {
            int64_t g = 0;
            for (;;) {
                const int64_t item = (int64_t)atomicAdd(item_counter, 1ull);
                if (item >= nitems) break;
                const int z = (int)(item / nstrips), strip = (int)(item - (int64_t)z * nstrips);
                for (int tile = 0; tile < tpi; ++tile, ++g) {
                    const int s = (int)(g % kXyzStages);
                    if (g >= kXyzStages) mbar_wait(&empty[s], (unsigned)(((g / kXyzStages) - 1) & 1));
                    info[s] = item * tpi + tile;
                    T* dst = ring + (size_t)s * 4 * TILE;
                    const int y = z * ny + tile * kXyzRows;
                    mbar_expect_tx(&full[s], (unsigned)(sizeof(T) * 4 * TILE));
                    tma_load_2d(dst, &tm_rho, strip * kXyzCols, y, &full[s]);
                    tma_load_2d(dst + TILE, &tm_ux, strip * kXyzCols, y, &full[s]);
                    tma_load_2d(dst + 2 * TILE, &tm_uy, strip * kXyzCols, y, &full[s]);
                    tma_load_2d(dst + 3 * TILE, &tm_uz, strip * kXyzCols, y, &full[s]);
                }
            }
            const int s = (int)(g % kXyzStages);  // end marker: completes the stage's phase without a copy
            if (g >= kXyzStages) mbar_wait(&empty[s], (unsigned)(((g / kXyzStages) - 1) & 1));
            info[s] = -1;
            mbar_arrive(&full[s]);
        }
        return;
    }

    Acc<kNM> ax;
    double cx0 = 0.0, cx1 = 0.0, cx2 = 0.0;
    for (int64_t g = 0;; ++g) {
        const int s = (int)(g % kXyzStages);
        mbar_wait(&full[s], (unsigned)((g / kXyzStages) & 1));
        const long long code = info[s];
        if (code < 0) break;
        const int64_t item = code / tpi;
        const int z = (int)(item / nstrips), strip = (int)(item - (int64_t)z * nstrips);
        const int tile = (int)(code - item * tpi), y0 = tile * kXyzRows;
        const T* st = ring + (size_t)s * 4 * TILE;
        if (colgroup && tile == 0) {
            ax.clear();
            const int64_t x = (int64_t)strip * kXyzCols + tid;
            cx0 = piv_x[x], cx1 = piv_x[nx + x], cx2 = piv_x[2 * (int64_t)nx + x];
        }
        double cy0 = 0.0, cy1 = 0.0, cy2 = 0.0;
        if (!colgroup) cy0 = piv_y[y0 + row], cy1 = piv_y[ny + y0 + row], cy2 = piv_y[2 * ny + y0 + row];
        if (colgroup) {
#pragma unroll
            for (int r = 0; r < kXyzRows; ++r) {
                const int o = r * kXyzCols + tid;
                ax.add((double)st[o], (double)st[TILE + o], (double)st[2 * TILE + o], (double)st[3 * TILE + o], cx0, cx1, cx2);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);  // this warp has read everything it needs from the stage
            if (tile == tpi - 1) {
                double* out = xpartial + (int64_t)z * kNM * nx + (int64_t)strip * kXyzCols + tid;
#pragma unroll
                for (int m = 0; m < kNM; ++m) out[(int64_t)m * nx] = ax.m[m];
            }
        } else {
            Acc<kNM> ay;
            ay.clear();
#pragma unroll
            for (int j = 0; j < kXyzCols / 32; ++j) {
                const int o = row * kXyzCols + lane + 32 * j;
                ay.add((double)st[o], (double)st[TILE + o], (double)st[2 * TILE + o], (double)st[3 * TILE + o], cy0, cy1, cy2);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            double v[16];
#pragma unroll
            for (int m = 0; m < 16; ++m) v[m] = m < kNM ? ay.m[m] : 0.0;
            const double tot = warp_halving_reduce16(v, lane);
            const int m = (lane >> 1) & 15;
            if ((lane & 1) == 0 && m < kNM)
                ypartial[((((int64_t)z * nstrips + strip) * ny) + y0 + row) * 16 + m] = tot;
        }
    }
}

// y bins = sum over (z, strip) of the 128-byte row records, in a fixed order; one CTA per y.
__global__ void __launch_bounds__(128)
    k_reduce_rows(const double* __restrict__ ypartial, int64_t nrec, int64_t ny, double cells_per_bin, double* __restrict__ mom) {
    const int64_t y = blockIdx.x;
    double acc[kNM];
#pragma unroll
    for (int m = 0; m < kNM; ++m) acc[m] = 0.0;
    for (int64_t q = threadIdx.x; q < nrec; q += blockDim.x) {
        const double2* rec = reinterpret_cast<const double2*>(ypartial + (q * ny + y) * 16);
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            const double2 p = rec[i];
            if (2 * i < kNM) acc[2 * i] += p.x;
            if (2 * i + 1 < kNM) acc[2 * i + 1] += p.y;
        }
    }
    __shared__ double sm[kNM][128];
#pragma unroll
    for (int m = 0; m < kNM; ++m) sm[m][threadIdx.x] = acc[m];
    __syncthreads();
    if (threadIdx.x < kNM) {
        double s = 0.0;
        for (int t = 0; t < 128; ++t) s += sm[threadIdx.x][t];
        mom[(int64_t)threadIdx.x * ny + y] = s;
    }
    if (threadIdx.x == kNM) mom[(int64_t)kNM * ny + y] = cells_per_bin;
}

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

static bool aligned_to(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

template <typename T, int V, int U, int NM>
static int launch_dense(fava_ctx* ctx, const T* rho, const T* ux, const T* uy, const T* uz, int64_t nz,
                        int64_t ny, int64_t nx, int axis, const double* piv, double* mom, int accumulate,
                        cudaStream_t st) {
    const int64_t nbins = axis == 0 ? nx : (axis == 1 ? ny : nz);
    const int64_t target_items = (int64_t)ctx->num_sms * 2 * 16;  // >= 16 waves of 2 CTAs/SM
    double* partial = nullptr;
    int nchunk = 1;
    if (axis == 0) {
        const int64_t nrows = nz * ny;
        const int64_t strips = ceil_div(nx, 32 * V);
        const int64_t step = (int64_t)kWarps * U;
        int64_t want = std::max<int64_t>(1, target_items / strips);
        int64_t rpc = round_up(ceil_div(nrows, want), step);
        nchunk = (int)ceil_div(nrows, rpc);
        void* ws;
        int rc = ctx_workspace(ctx, WS_PARTIALS, sizeof(double) * (size_t)nchunk * NM * nbins, &ws);
        if (rc) return rc;
        partial = (double*)ws;
        dim3 grid((unsigned)strips, (unsigned)nchunk);
        k_moments_cols<T, V, U, NM><<<grid, kThreads, 0, st>>>(rho, ux, uy, uz, nrows, nx, piv, partial, rpc);
        FAVA_LAUNCHED();
    } else {
        int64_t row_len, bin_stride, row_stride, rows_per_bin;
        if (axis == 1) {
            row_len = nx, bin_stride = nx, row_stride = ny * nx, rows_per_bin = nz;
        } else {
            // the plane is contiguous: re-tile it into rows of one CTA-wide vector step
            const int64_t plane = ny * nx, tile = (int64_t)kThreads * V;
            bin_stride = plane;
            if (plane % tile == 0) row_len = tile, rows_per_bin = plane / tile;
            else row_len = nx, rows_per_bin = ny;
            row_stride = row_len;
        }
        int lanes_x = kThreads;
        while (lanes_x > 1 && (int64_t)(lanes_x / 2) * V >= row_len) lanes_x /= 2;
        const int rpi = kThreads / lanes_x;
        const int64_t step = (int64_t)rpi * U;
        int64_t want = std::max<int64_t>(1, ceil_div(target_items, nbins));
        int64_t rpc = round_up(ceil_div(rows_per_bin, want), step);
        nchunk = (int)ceil_div(rows_per_bin, rpc);
        void* ws;
        int rc = ctx_workspace(ctx, WS_PARTIALS, sizeof(double) * (size_t)nchunk * NM * nbins, &ws);
        if (rc) return rc;
        partial = (double*)ws;
        dim3 grid((unsigned)nbins, (unsigned)nchunk);
        k_moments_rows<T, V, U, NM><<<grid, kThreads, 0, st>>>(rho, ux, uy, uz, row_len, bin_stride, row_stride,
                                                          rows_per_bin, rpc, piv, nbins, partial, lanes_x);
        FAVA_LAUNCHED();
    }
    const int nrows = NM > 1 ? FAVA_NMOM : 1;
    const int64_t n = (int64_t)nrows * nbins;
    const double cells = (double)(nz * ny * nx / nbins);
    k_reduce_partials<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(partial, nchunk, nbins, NM, nrows, mom, accumulate,
                                                                 cells);
    FAVA_LAUNCHED();
    return FAVA_OK;
}

template <typename T, int V, int U>
static int launch_xz(fava_ctx* ctx, const T* rho, const T* ux, const T* uy, const T* uz, int64_t nz, int64_t ny,
                     int64_t nx, const double* piv_x, const double* piv_z, double* mom_x, double* mom_z,
                     cudaStream_t st) {
    const int64_t strips = ceil_div(nx, 32 * V);
    void* ws;
    int rc = ctx_workspace(ctx, WS_PARTIALS, sizeof(double) * (size_t)nz * kNM * nx, &ws);
    if (rc) return rc;
    double* partial = (double*)ws;
    dim3 grid((unsigned)strips, (unsigned)nz);  // one chunk of ny rows per z-plane
    k_moments_cols<T, V, U, kNM><<<grid, kThreads, 0, st>>>(rho, ux, uy, uz, nz * ny, nx, piv_x, partial, ny);
    FAVA_LAUNCHED();
    const int64_t n = (int64_t)FAVA_NMOM * nx;
    k_reduce_partials<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(partial, (int)nz, nx, kNM, FAVA_NMOM, mom_x, 0,
                                                                 (double)(nz * ny));
    FAVA_LAUNCHED();
    k_partials_to_planes<<<(unsigned)nz, 256, 0, st>>>(partial, nx, nz, (double)ny, piv_x, piv_z, mom_z,
                                                      (double)(nx * ny));
    FAVA_LAUNCHED();
    return FAVA_OK;
}

template <typename T, int U>
static int dispatch_xz(fava_ctx* ctx, const void* rho, const void* ux, const void* uy, const void* uz, int64_t nz,
                       int64_t ny, int64_t nx, const double* piv_x, const double* piv_z, double* mom_x,
                       double* mom_z, cudaStream_t st) {
    const size_t va = 2 * sizeof(T);
    const bool vec = (nx % 2 == 0) && aligned_to(rho, va) && aligned_to(ux, va) && aligned_to(uy, va) &&
                     aligned_to(uz, va);
    if (vec)
        return launch_xz<T, 2, U>(ctx, (const T*)rho, (const T*)ux, (const T*)uy, (const T*)uz, nz, ny, nx, piv_x,
                                  piv_z, mom_x, mom_z, st);
    return launch_xz<T, 1, U>(ctx, (const T*)rho, (const T*)ux, (const T*)uy, (const T*)uz, nz, ny, nx, piv_x, piv_z,
                              mom_x, mom_z, st);
}

template <typename T, int U, int NM>
static int dispatch_dense(fava_ctx* ctx, const void* rho, const void* ux, const void* uy, const void* uz,
                          int64_t nz, int64_t ny, int64_t nx, int axis, const double* piv, double* mom,
                          int accumulate, cudaStream_t st) {
    const size_t va = 2 * sizeof(T);
    const bool vec = (nx % 2 == 0) && aligned_to(rho, va) && aligned_to(ux, va) && aligned_to(uy, va) &&
                     aligned_to(uz, va);
    if (vec)
        return launch_dense<T, 2, U, NM>(ctx, (const T*)rho, (const T*)ux, (const T*)uy, (const T*)uz, nz, ny, nx,
                                     axis, piv, mom, accumulate, st);
    return launch_dense<T, 1, U, NM>(ctx, (const T*)rho, (const T*)ux, (const T*)uy, (const T*)uz, nz, ny, nx, axis,
                                 piv, mom, accumulate, st);
}

template <typename T>
static int launch_xyz(fava_ctx* ctx, const T* rho, const T* ux, const T* uy, const T* uz, int64_t nz, int64_t ny, int64_t nx,
                      const double* piv_x, const double* piv_y, const double* piv_z, double* mom_x, double* mom_y,
                      double* mom_z, cudaStream_t st) {
    const int64_t nstrips = nx / kXyzCols;
    void *wx, *wy;
    int rc = ctx_workspace(ctx, WS_PARTIALS, sizeof(double) * (size_t)nz * kNM * nx, &wx);
    if (rc) return rc;
    rc = ctx_workspace(ctx, WS_YPART, sizeof(double) * 16 * (size_t)(nz * nstrips * ny), &wy);
    if (rc) return rc;
    const CUtensorMapDataType dt = sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    const uint64_t dims[2] = {(uint64_t)nx, (uint64_t)(nz * ny)};
    const uint64_t strides[1] = {(uint64_t)nx * sizeof(T)};
    const uint32_t box[2] = {(uint32_t)kXyzCols, (uint32_t)kXyzRows};
    CUtensorMap tm[4];
    const T* src[4] = {rho, ux, uy, uz};
    for (int f = 0; f < 4; ++f) {
        rc = ctx_tensor_map(ctx, src[f], dt, 2, dims, strides, box, &tm[f]);
        if (rc) return rc;
    }
    const size_t smem = sizeof(T) * kXyzStages * 4 * kXyzRows * kXyzCols + 24 * kXyzStages;
    auto kern = k_moments_xyz<T>;
    FAVA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)std::min<int64_t>(nz * nstrips, ctx->num_sms);
    if (!ctx->tile_counters) FAVA_CHECK_CUDA(cudaMalloc(&ctx->tile_counters, sizeof(unsigned long long) * 64));
    unsigned long long* counter = (unsigned long long*)ctx->tile_counters + (ctx->tile_counter_next++ & 63);
    FAVA_CHECK_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st));
    kern<<<grid, kXyzThreads, smem, st>>>(tm[0], tm[1], tm[2], tm[3], (int)nz, (int)ny, (int)nx, piv_x, piv_y, (double*)wx,
                                         (double*)wy, counter);
    FAVA_LAUNCHED();
    const int64_t n = (int64_t)FAVA_NMOM * nx;
    k_reduce_partials<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>((const double*)wx, (int)nz, nx, kNM, FAVA_NMOM, mom_x, 0,
                                                                 (double)(nz * ny));
    FAVA_LAUNCHED();
    k_partials_to_planes<<<(unsigned)nz, 256, 0, st>>>((const double*)wx, nx, nz, (double)ny, piv_x, piv_z, mom_z,
                                                      (double)(nx * ny));
    FAVA_LAUNCHED();
    k_reduce_rows<<<(unsigned)ny, 128, 0, st>>>((const double*)wy, nz * nstrips, ny, (double)(nz * nx), mom_y);
    FAVA_LAUNCHED();
    return FAVA_OK;
}

}  // namespace fava

using namespace fava;

extern "C" {

int fava_plane_pivots(fava_ctx* ctx, const void* d_ux, const void* d_uy, const void* d_uz, int dtype,
                      int64_t nz, int64_t ny, int64_t nx, int axis, double* d_pivots, void* stream) {
    FAVA_REQUIRE(ctx && d_ux && d_uy && d_uz && d_pivots, "fava_plane_pivots: NULL argument");
    FAVA_REQUIRE(nz > 0 && ny > 0 && nx > 0, "fava_plane_pivots: empty array %lldx%lldx%lld", (long long)nz,
                 (long long)ny, (long long)nx);
    FAVA_REQUIRE(axis >= 0 && axis <= 2, "fava_plane_pivots: axis %d not in 0..2", axis);
    FAVA_REQUIRE(dtype == FAVA_F32 || dtype == FAVA_F64, "fava_plane_pivots: bad dtype %d", dtype);
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nbins = axis == 0 ? nx : (axis == 1 ? ny : nz);
    const int64_t stride = axis == 0 ? 1 : (axis == 1 ? nx : ny * nx);
    const unsigned grid = (unsigned)ceil_div(nbins, 256);
    if (dtype == FAVA_F64)
        k_plane_pivots<double><<<grid, 256, 0, st>>>((const double*)d_ux, (const double*)d_uy,
                                                     (const double*)d_uz, stride, nbins, d_pivots);
    else
        k_plane_pivots<float><<<grid, 256, 0, st>>>((const float*)d_ux, (const float*)d_uy, (const float*)d_uz,
                                                    stride, nbins, d_pivots);
    FAVA_LAUNCHED();
    return FAVA_OK;
}

int fava_plane_moments(fava_ctx* ctx, const void* d_rho, const void* d_ux, const void* d_uy,
                       const void* d_uz, int dtype, int64_t nz, int64_t ny, int64_t nx, int axis,
                       const double* d_pivots, double* d_moments, int accumulate, void* stream) {
    FAVA_REQUIRE(ctx && d_rho && d_ux && d_uy && d_uz && d_pivots && d_moments,
                 "fava_plane_moments: NULL argument");
    FAVA_REQUIRE(nz > 0 && ny > 0 && nx > 0, "fava_plane_moments: empty array %lldx%lldx%lld", (long long)nz,
                 (long long)ny, (long long)nx);
    FAVA_REQUIRE(axis >= 0 && axis <= 2, "fava_plane_moments: axis %d not in 0..2", axis);
    FAVA_REQUIRE(dtype == FAVA_F32 || dtype == FAVA_F64, "fava_plane_moments: bad dtype %d", dtype);
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == FAVA_F64)
        return dispatch_dense<double, 4, kNM>(ctx, d_rho, d_ux, d_uy, d_uz, nz, ny, nx, axis, d_pivots, d_moments,
                                         accumulate, st);
    return dispatch_dense<float, 8, kNM>(ctx, d_rho, d_ux, d_uy, d_uz, nz, ny, nx, axis, d_pivots, d_moments,
                                    accumulate, st);
}

int fava_plane_moments_xz(fava_ctx* ctx, const void* d_rho, const void* d_ux, const void* d_uy, const void* d_uz,
                          int dtype, int64_t nz, int64_t ny, int64_t nx, const double* d_piv_x,
                          const double* d_piv_z, double* d_mom_x, double* d_mom_z, void* stream) {
    FAVA_REQUIRE(ctx && d_rho && d_ux && d_uy && d_uz && d_piv_x && d_piv_z && d_mom_x && d_mom_z,
                 "fava_plane_moments_xz: NULL argument");
    FAVA_REQUIRE(nz > 0 && ny > 0 && nx > 0 && nz <= 65535, "fava_plane_moments_xz: bad shape %lldx%lldx%lld",
                 (long long)nz, (long long)ny, (long long)nx);
    FAVA_REQUIRE(dtype == FAVA_F32 || dtype == FAVA_F64, "fava_plane_moments_xz: bad dtype %d", dtype);
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == FAVA_F64)
        return dispatch_xz<double, 4>(ctx, d_rho, d_ux, d_uy, d_uz, nz, ny, nx, d_piv_x, d_piv_z, d_mom_x, d_mom_z, st);
    return dispatch_xz<float, 8>(ctx, d_rho, d_ux, d_uy, d_uz, nz, ny, nx, d_piv_x, d_piv_z, d_mom_x, d_mom_z, st);
}

int fava_plane_moments_xyz_supported(int64_t nz, int64_t ny, int64_t nx) {
    return (nz > 0 && nz * ny < (int64_t(1) << 31) && nx % kXyzCols == 0 && ny % kXyzRows == 0) ? 1 : 0;
}

int fava_plane_moments_xyz(fava_ctx* ctx, const void* d_rho, const void* d_ux, const void* d_uy, const void* d_uz, int dtype,
                           int64_t nz, int64_t ny, int64_t nx, const double* d_piv_x, const double* d_piv_y,
                           const double* d_piv_z, double* d_mom_x, double* d_mom_y, double* d_mom_z, void* stream) {
    FAVA_REQUIRE(ctx && d_rho && d_ux && d_uy && d_uz && d_piv_x && d_piv_y && d_piv_z && d_mom_x && d_mom_y && d_mom_z,
                 "fava_plane_moments_xyz: NULL argument");
    FAVA_REQUIRE(dtype == FAVA_F32 || dtype == FAVA_F64, "fava_plane_moments_xyz: bad dtype %d", dtype);
    FAVA_REQUIRE(fava_plane_moments_xyz_supported(nz, ny, nx),
                 "fava_plane_moments_xyz: needs nx %% %d == 0 and ny %% %d == 0 (got %lldx%lldx%lld); use the per-axis passes",
                 kXyzCols, kXyzRows, (long long)nz, (long long)ny, (long long)nx);
    FAVA_REQUIRE(aligned_to(d_rho, 16) && aligned_to(d_ux, 16) && aligned_to(d_uy, 16) && aligned_to(d_uz, 16),
                 "fava_plane_moments_xyz: fields must be 16-byte aligned");
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == FAVA_F64)
        return launch_xyz<double>(ctx, (const double*)d_rho, (const double*)d_ux, (const double*)d_uy, (const double*)d_uz, nz,
                                  ny, nx, d_piv_x, d_piv_y, d_piv_z, d_mom_x, d_mom_y, d_mom_z, st);
    return launch_xyz<float>(ctx, (const float*)d_rho, (const float*)d_ux, (const float*)d_uy, (const float*)d_uz, nz, ny, nx,
                             d_piv_x, d_piv_y, d_piv_z, d_mom_x, d_mom_y, d_mom_z, st);
}

int fava_plane_sum(fava_ctx* ctx, const void* d_field, int dtype, int64_t nz, int64_t ny, int64_t nx, int axis,
                   double* d_out, void* stream) {
    FAVA_REQUIRE(ctx && d_field && d_out, "fava_plane_sum: NULL argument");
    FAVA_REQUIRE(nz > 0 && ny > 0 && nx > 0, "fava_plane_sum: empty array");
    FAVA_REQUIRE(axis >= 0 && axis <= 2, "fava_plane_sum: axis %d not in 0..2", axis);
    FAVA_REQUIRE(dtype == FAVA_F32 || dtype == FAVA_F64, "fava_plane_sum: bad dtype %d", dtype);
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    // same streaming kernels as the moment pass with one accumulator; only the first operand is read
    if (dtype == FAVA_F64)
        return dispatch_dense<double, 4, 1>(ctx, d_field, d_field, d_field, d_field, nz, ny, nx, axis, nullptr, d_out,
                                            0, st);
    return dispatch_dense<float, 8, 1>(ctx, d_field, d_field, d_field, d_field, nz, ny, nx, axis, nullptr, d_out, 0, st);
}

int fava_moments_repivot(fava_ctx* ctx, double* d_moments, const double* d_piv_old, const double* d_piv_new,
                         int64_t nbins, void* stream) {
    FAVA_REQUIRE(ctx && d_moments && d_piv_old && d_piv_new, "fava_moments_repivot: NULL argument");
    FAVA_REQUIRE(nbins > 0, "fava_moments_repivot: nbins must be positive");
    DeviceGuard g(ctx->device);
    k_repivot<<<(unsigned)ceil_div(nbins, 128), 128, 0, (cudaStream_t)stream>>>(d_moments, d_piv_old, d_piv_new,
                                                                                nbins);
    FAVA_LAUNCHED();
    return FAVA_OK;
}

int fava_moments_finalize(fava_ctx* ctx, const double* d_moments, const double* d_pivots, int64_t nbins,
                          double weight, double layer_volume, double* d_means, double* d_rey,
                          double* d_fmeans, double* d_favre, void* stream) {
    FAVA_REQUIRE(ctx && d_moments && d_pivots, "fava_moments_finalize: NULL argument");
    FAVA_REQUIRE(nbins > 0, "fava_moments_finalize: nbins must be positive");
    FAVA_REQUIRE(layer_volume > 0.0, "fava_moments_finalize: layer_volume must be positive");
    DeviceGuard g(ctx->device);
    k_finalize<<<(unsigned)ceil_div(nbins, 128), 128, 0, (cudaStream_t)stream>>>(
        d_moments, d_pivots, nbins, weight, layer_volume, d_means, d_rey, d_fmeans, d_favre);
    FAVA_LAUNCHED();
    return FAVA_OK;
}

}  // extern "C"
