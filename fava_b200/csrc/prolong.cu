// K3 — AMR -> uniform prolongation (piecewise-constant injection) through a block-index table.
//
// Replaces the gather of FLASH.from_amr (reference fava/mesh/FLASH/_flash.py:1262-1321): a Python dict
// with one entry per fine cell, mapping[(I,J,K)] = (leaf,i,j,k), built by triple loops (~8 us/cell)
// and replayed per field (~0.8 us/cell).  Here the host turns the selected-leaf list (built with the
// reference's integer arithmetic, _flash.py:1000-1022 / :1157-1199) into a lattice table
//   tile (X/nxb, Y/nyb, Z/nzb) of the fine grid  ->  (source block, corner, log2 scale) of the leaf that owns it
// (leaves are written in list order, so a later leaf overwrites an earlier one exactly like the dict
// does, and tiles nobody owns stay -1 => 0.0 as in_data[...] = 0.0, :1258).  The kernel is a pure
// gather over whole tile rows: source cell = (fine - corner) >> log2(scale), stores in FILE order [Z][Y][X].
// Bit-exact (f32 -> f64 widening is exact).
// Traffic: every selected source cell is read once from HBM (repeats hit L1/L2), 8 B written per cell.
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"

namespace fava {

struct ProlongGeom {
    int nxb, nyb, nzb;     // block shape
    int sx, sy, sz;        // lattice shift: tile t covers fine cells [t*nb + s - nb, ...)
    int tx, ty, tz;        // table dims
    int64_t NX, NY, NZ;    // output dims
};

// What a tile needs from its leaf, resolved on the host: ONE dependent load per tile instead of table -> leaf.
struct TileDesc {
    int64_t block;   // source block index, -1 = nobody owns the tile (zeros, in_data[...] = 0.0, _flash.py:1258)
    int32_t off[3];  // fine-cell corner of the leaf relative to the output
    int32_t shift;   // log2(scale)
};

// CTAs walk the lattice tiles (nxb x nyb x nzb fine cells, all owned by ONE leaf) with a grid stride.  Work inside a
// tile is assigned by ROWS: `tpr` threads share a row of the tile and each writes two adjacent cells with one 16-byte
// streaming store, the next row of a thread follows by increments - no division or modulo per cell (the first version
// spent ~100 integer instructions per cell on them and ran at 0.35 of the HBM peak, issue-bound).  Source cell =
// (fine - corner) >> log2(scale); repeats for scale > 1 hit L1.  Clipped or odd-aligned tiles take the scalar path.
template <typename T>
__global__ void __launch_bounds__(256)
    k_prolong(const T* __restrict__ blocks, const TileDesc* __restrict__ tiles, ProlongGeom g, int64_t ntile, int tpr,
              double* __restrict__ out) {
    const int lane_x = threadIdx.x % tpr, row0 = threadIdx.x / tpr, rows_per_iter = blockDim.x / tpr;
    const int64_t bcells = (int64_t)g.nzb * g.nyb * g.nxb;
    for (int64_t tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
        const int tx = (int)(tile % g.tx), ty = (int)((tile / g.tx) % g.ty), tz = (int)(tile / ((int64_t)g.tx * g.ty));
        // fine-cell box of the tile, clipped to the output
        const int x0 = tx * g.nxb + g.sx - g.nxb, y0 = ty * g.nyb + g.sy - g.nyb, z0 = tz * g.nzb + g.sz - g.nzb;
        const int xa = max(x0, 0), xb = (int)min((int64_t)x0 + g.nxb, g.NX);
        const int ya = max(y0, 0), yb = (int)min((int64_t)y0 + g.nyb, g.NY);
        const int za = max(z0, 0), zb = (int)min((int64_t)z0 + g.nzb, g.NZ);
        const int wx = xb - xa, wy = yb - ya, wz = zb - za;
        if (wx <= 0 || wy <= 0 || wz <= 0) continue;
        const TileDesc d = tiles[tile];
        const T* src = blocks + (d.block < 0 ? 0 : d.block) * bcells;
        const bool vec = ((g.NX | xa) & 1) == 0;  // 16-byte aligned pairs
        const int nrows = wy * wz;
        int iy = row0 % wy, iz = row0 / wy;  // once per tile; rows then advance by increments
        const int dy = rows_per_iter % wy, dz = rows_per_iter / wy;
        const bool whole = vec && wx == 2 * tpr;  // unclipped tile, one 16-byte pair per thread and row: the common case
        if (whole) {
            // four rows per step: the four source loads are in flight together, then four 16-byte streaming stores
            const int x = xa + 2 * lane_x;
            const int i0 = (x - d.off[0]) >> d.shift, i1 = (x + 1 - d.off[0]) >> d.shift;
            for (int r = row0; r < nrows; r += 4 * rows_per_iter) {
                double2 v[4];
                double* dst[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    dst[k] = nullptr;
                    if (r + k * rows_per_iter < nrows) {
                        const int Y = ya + iy, Z = za + iz;
                        dst[k] = out + ((int64_t)Z * g.NY + Y) * g.NX + x;
                        if (d.block < 0) {
                            v[k] = make_double2(0.0, 0.0);
                        } else {
                            const T* srow = src + (((Z - d.off[2]) >> d.shift) * g.nyb + ((Y - d.off[1]) >> d.shift)) * g.nxb;
                            v[k] = make_double2((double)srow[i0], (double)srow[i1]);
                        }
                        iy += dy, iz += dz;
                        if (iy >= wy) iy -= wy, ++iz;
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (dst[k]) __stcs(reinterpret_cast<double2*>(dst[k]), v[k]);
            }
            continue;
        }
        for (int r = row0; r < nrows; r += rows_per_iter) {
            const int Y = ya + iy, Z = za + iz;
            double* orow = out + ((int64_t)Z * g.NY + Y) * g.NX;
            const T* srow = src + (((Z - d.off[2]) >> d.shift) * g.nyb + ((Y - d.off[1]) >> d.shift)) * g.nxb;
            for (int x = xa + 2 * lane_x; x < xb; x += 2 * tpr) {
                const double v0 = d.block < 0 ? 0.0 : (double)srow[(x - d.off[0]) >> d.shift];
                if (x + 1 < xb) {
                    const double v1 = d.block < 0 ? 0.0 : (double)srow[(x + 1 - d.off[0]) >> d.shift];
                    if (vec) {
                        __stcs(reinterpret_cast<double2*>(orow + x), make_double2(v0, v1));
                    } else {
                        __stcs(orow + x, v0);
                        __stcs(orow + x + 1, v1);
                    }
                } else {
                    __stcs(orow + x, v0);
                }
            }
            iy += dy, iz += dz;
            if (iy >= wy) iy -= wy, ++iz;
        }
    }
}

// Whole output rows (tile lattice aligned with x = 0, even NX and nxb): a work item is a ROW OF TILES - every tile
// with the same (ty, tz) - and `tpr` threads write one output row together, two cells = one 16-byte store each, so the
// CTA's stores are runs of 16 * tpr bytes (a whole 4 KB row at NX = 512) instead of 128-byte pieces 4 KB apart.  A
// thread keeps the descriptor of the tile above its x position for the whole item; rows advance by increments, four
// rows in flight.  Used for blocks narrower than a 128-byte line (see run_prolong).
template <typename T>
__global__ void __launch_bounds__(256)
    k_prolong_rows(const T* __restrict__ blocks, const TileDesc* __restrict__ tiles, ProlongGeom g, int64_t nitems, int tpr,
                   double* __restrict__ out) {
    const int lane_x = threadIdx.x % tpr, row0 = threadIdx.x / tpr, rows_per_iter = blockDim.x / tpr;
    const int64_t bcells = (int64_t)g.nzb * g.nyb * g.nxb;
    for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int ty = (int)(item % g.ty), tz = (int)(item / g.ty);
        const int y0 = ty * g.nyb + g.sy - g.nyb, z0 = tz * g.nzb + g.sz - g.nzb;
        const int ya = max(y0, 0), yb = (int)min((int64_t)y0 + g.nyb, g.NY);
        const int za = max(z0, 0), zb = (int)min((int64_t)z0 + g.nzb, g.NZ);
        const int wy = yb - ya, wz = zb - za;
        if (wy <= 0 || wz <= 0) continue;
        const int nrows = wy * wz;
        const int dy = rows_per_iter % wy, dz = rows_per_iter / wy;
        for (int x = 2 * lane_x; x < g.NX; x += 2 * tpr) {  // one pass unless NX > 512
            const TileDesc d = tiles[((int64_t)tz * g.ty + ty) * g.tx + x / g.nxb + 1];  // sx = 0: tile 0 lies left of x = 0
            const T* src = blocks + (d.block < 0 ? 0 : d.block) * bcells;
            const int i0 = (x - d.off[0]) >> d.shift, i1 = (x + 1 - d.off[0]) >> d.shift;
            int iy = row0 % wy, iz = row0 / wy;
            for (int r = row0; r < nrows; r += 4 * rows_per_iter) {
                double2 v[4];
                double* dst[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    dst[k] = nullptr;
                    if (r + k * rows_per_iter < nrows) {
                        const int Y = ya + iy, Z = za + iz;
                        dst[k] = out + ((int64_t)Z * g.NY + Y) * g.NX + x;
                        if (d.block < 0) {
                            v[k] = make_double2(0.0, 0.0);
                        } else {
                            const T* srow = src + (((Z - d.off[2]) >> d.shift) * g.nyb + ((Y - d.off[1]) >> d.shift)) * g.nxb;
                            v[k] = make_double2((double)srow[i0], (double)srow[i1]);
                        }
                        iy += dy, iz += dz;
                        if (iy >= wy) iy -= wy, ++iz;
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (dst[k]) __stcs(reinterpret_cast<double2*>(dst[k]), v[k]);
            }
        }
    }
}

static inline int pmod(int64_t a, int64_t n) { return (int)(((a % n) + n) % n); }
static inline int64_t cdivp(int64_t a, int64_t b) { return (a + b - 1) / b; }

template <typename T>
static int run_prolong(fava_ctx* ctx, const T* blocks, int64_t nzb, int64_t nyb, int64_t nxb,
                       const fava_prolong_leaf* h_leaves, int64_t nleaf, int64_t NZ, int64_t NY, int64_t NX,
                       double* out, cudaStream_t st) {
    ProlongGeom g;
    g.nxb = (int)nxb, g.nyb = (int)nyb, g.nzb = (int)nzb;
    g.NX = NX, g.NY = NY, g.NZ = NZ;
    g.sx = g.sy = g.sz = 0;
    if (nleaf) {
        g.sx = pmod(h_leaves[0].off[0], nxb);
        g.sy = pmod(h_leaves[0].off[1], nyb);
        g.sz = pmod(h_leaves[0].off[2], nzb);
    }
    g.tx = (int)cdivp(NX - g.sx, nxb) + 1;
    g.ty = (int)cdivp(NY - g.sy, nyb) + 1;
    g.tz = (int)cdivp(NZ - g.sz, nzb) + 1;
    const int64_t ntile = (int64_t)g.tx * g.ty * g.tz;
    const size_t b_table = sizeof(TileDesc) * (size_t)ntile;
    // from_amr prolongs several fields with ONE leaf list: an identical request re-uses the device tables
    const int64_t head[7] = {nzb, nyb, nxb, NZ, NY, NX, nleaf};
    const size_t key_bytes = sizeof(head) + sizeof(fava_prolong_leaf) * (size_t)nleaf;
    const std::string& have = ctx->prolong_cache_key;
    TileDesc* d_table;
    if (ctx->ws[WS_TABLE] && have.size() == key_bytes && memcmp(have.data(), head, sizeof(head)) == 0 &&
        (nleaf == 0 || memcmp(have.data() + sizeof(head), h_leaves, sizeof(fava_prolong_leaf) * (size_t)nleaf) == 0)) {
        d_table = (TileDesc*)ctx->ws[WS_TABLE];
    } else {
        ctx->prolong_cache_key.clear();
        TileDesc none;
        none.block = -1, none.off[0] = none.off[1] = none.off[2] = 0, none.shift = 0;
        std::vector<TileDesc> table((size_t)ntile, none);
        for (int64_t l = 0; l < nleaf; ++l) {
            const fava_prolong_leaf& d = h_leaves[l];
            if (d.scale < 1 || (d.scale & (d.scale - 1)) || d.block < 0)
                return set_error(FAVA_EINVAL, "fava_prolong: leaf %lld has scale %d (must be a power of two) / block %lld",
                                 (long long)l, d.scale, (long long)d.block);
            if (pmod(d.off[0], nxb) != g.sx || pmod(d.off[1], nyb) != g.sy || pmod(d.off[2], nzb) != g.sz)
                return set_error(FAVA_EINVAL, "fava_prolong: leaf %lld corner (%d,%d,%d) is not on the block lattice",
                                 (long long)l, d.off[0], d.off[1], d.off[2]);
            // tiles covered by the leaf, clipped to the table
            int64_t lo[3], hi[3];
            const int dims[3] = {g.tx, g.ty, g.tz};
            bool empty = false;
            for (int a = 0; a < 3; ++a) {
                // off may be far negative: floor division
                const int64_t nb = a == 0 ? nxb : (a == 1 ? nyb : nzb);
                const int64_t num = (int64_t)d.off[a] - (a == 0 ? g.sx : (a == 1 ? g.sy : g.sz)) + nb;
                const int64_t first = num >= 0 ? num / nb : -((-num + nb - 1) / nb);
                lo[a] = std::max<int64_t>(first, 0);
                hi[a] = std::min<int64_t>(first + d.scale, dims[a]);
                if (hi[a] <= lo[a]) empty = true;
            }
            if (empty) continue;
            TileDesc td;
            td.block = d.block, td.off[0] = d.off[0], td.off[1] = d.off[1], td.off[2] = d.off[2];
            td.shift = 0;
            while ((1 << td.shift) < d.scale) ++td.shift;
            for (int64_t z = lo[2]; z < hi[2]; ++z)
                for (int64_t y = lo[1]; y < hi[1]; ++y)
                    for (int64_t x = lo[0]; x < hi[0]; ++x) table[(size_t)((z * g.ty + y) * g.tx + x)] = td;  // later leaves win
        }
        void* tab;
        int rc = ctx_workspace(ctx, WS_TABLE, b_table, &tab);
        if (rc) return rc;
        d_table = (TileDesc*)tab;
        // pageable source: the copy is staged before the call returns, so `table` may go out of scope
        FAVA_CHECK_CUDA(cudaMemcpyAsync(d_table, table.data(), b_table, cudaMemcpyHostToDevice, st));
        std::string key(key_bytes, '\0');
        memcpy(&key[0], head, sizeof(head));
        if (nleaf) memcpy(&key[sizeof(head)], h_leaves, sizeof(fava_prolong_leaf) * (size_t)nleaf);
        ctx->prolong_cache_key.swap(key);
    }

    // rows of tiles when a tile's own rows are shorter than a 128-byte line (8-cell blocks: tile-wise stores are
    // half lines, 0.10 ms for 256^3 against 0.039 ms row-wise); with 16-cell blocks a tile row is a whole line and the
    // tile-wise walk, whose source reads are contiguous inside one block, is the faster one (0.29 against 0.38 ms at 512^3)
    if (nxb * sizeof(double) < 128 && g.sx == 0 && NX % 2 == 0 && nxb % 2 == 0 && (uintptr_t)out % 16 == 0) {
        int tpr = 1;  // threads per output row: two cells each, a power of two <= 256
        while (tpr < 256 && 4 * tpr <= NX) tpr *= 2;
        const int64_t nitems = (int64_t)g.ty * g.tz;
        const unsigned grid = (unsigned)std::min<int64_t>(nitems, (int64_t)ctx->num_sms * 8);
        k_prolong_rows<T><<<grid, 256, 0, st>>>(blocks, d_table, g, nitems, tpr, out);
    } else {
        int tpr = 1;  // threads per row of a tile: two cells each, a power of two <= 32
        while (tpr < 32 && 2 * tpr < nxb) tpr *= 2;
        const unsigned grid = (unsigned)std::min<int64_t>(ntile, (int64_t)ctx->num_sms * 16);
        k_prolong<T><<<grid, 256, 0, st>>>(blocks, d_table, g, ntile, tpr, out);
    }
    FAVA_LAUNCHED();
    return FAVA_OK;
}

}  // namespace fava

using namespace fava;

extern "C" int fava_prolong(fava_ctx* ctx, const void* d_blocks, int dtype, int64_t nzb, int64_t nyb, int64_t nxb,
                            const fava_prolong_leaf* h_leaves, int64_t nleaf, int64_t NZ, int64_t NY, int64_t NX,
                            double* d_out, void* stream) {
    FAVA_REQUIRE(ctx && d_out, "fava_prolong: NULL argument");
    FAVA_REQUIRE(nleaf >= 0 && (nleaf == 0 || (h_leaves && d_blocks)), "fava_prolong: NULL block data or leaf table");
    FAVA_REQUIRE(nzb > 0 && nyb > 0 && nxb > 0, "fava_prolong: bad block shape");
    FAVA_REQUIRE(NZ > 0 && NY > 0 && NX > 0, "fava_prolong: empty output %lldx%lldx%lld", (long long)NZ,
                 (long long)NY, (long long)NX);
    FAVA_REQUIRE(dtype == FAVA_F32 || dtype == FAVA_F64, "fava_prolong: bad dtype %d", dtype);
    FAVA_REQUIRE(nleaf < INT32_MAX, "fava_prolong: too many leaves");
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == FAVA_F64)
        return run_prolong<double>(ctx, (const double*)d_blocks, nzb, nyb, nxb, h_leaves, nleaf, NZ, NY, NX, d_out, st);
    return run_prolong<float>(ctx, (const float*)d_blocks, nzb, nyb, nxb, h_leaves, nleaf, NZ, NY, NX, d_out, st);
}
