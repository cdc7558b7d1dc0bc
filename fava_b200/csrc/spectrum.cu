// Kinetic-energy spectrum on sm_100a — replaces FlashUniform.kinetic_energy_spectra
// (reference fava/mesh/FLASH/FlashUniform.py:229-304: np.fft.fftn of complex128 N^3 arrays on one
// thread, a 3*N^3 fp64 k-grid, fftshift copies and three scipy binned_statistic passes).
//
// Pipeline (per GPU; every array stays in FILE order [z][y][x], x fastest):
//   transform          power-of-two N (256..2048): the hand-written passes of csrc/fft.cu - weighting fused into the
//                      x pass, pruned y and z passes - in Hermitian storage complex [kz][ky][kx = 0..N/2-1]
//                      (row pitch N/2: the Nyquist column is never read by a bin and is not stored).
//                      Any other even N:  K4 k_ke_weight3 (w_n = sqrt(rho) u_n, one pass, row-padded) + cuFFT -
//                      the only library call on this path (BASELINE.json north_star) - batched 2-D D2Z over
//                      (y,x), then strided 1-D Z2Z along z, in place; row pitch N/2+1.
//   K6  k_spectrum_bin |u^|^2, the reference's longitudinal projection INCLUDING its `.T` quirk
//                      (FlashUniform.py:281: ffts[n].T reverses all axes, i.e. the operand is taken at
//                      the transposed wavevector (kz,ky,kx)), shell index floor(|k|+1/2) in exact
//                      integer arithmetic, per-shell sums with weight 2 for the kx>0 half.
//       k_spectrum_reduce / fava_spectrum_finalize: fixed-order merge, shell mean x 4 pi k^2.
//
// Why r2c is exact for this statistic: both `total` and the quirky `longitudinal` are invariant under
// k -> -k for a real input (u^(-k) = conj u^(k)), so the kx<0 half contributes the same values as its
// mirror; the Nyquist planes (|k_i| = N/2) lie beyond the last bin edge N/2-1.5 and never contribute.
// The transposed operand u^_n(kz,ky,kx) is read from the stored half directly when kz >= 0 and as
// conj(u^_n(-kz,-ky,-kx)) otherwise; tiles are transposed through shared memory so that both the
// direct and the transposed reads are coalesced.  Tiles entirely outside the sphere |k| <= N/2-1.5
// (~48 % of the half-cube) are skipped before any load.
//
// Determinism: lanes of a warp hold consecutive kx of one (ky,kz) row, so shell indices are
// non-decreasing along the warp; a segmented shuffle scan reduces them in a fixed order, segment
// tails add into warp-private tile bins (plain stores), the CTA merges its warps in a fixed order,
// CTAs own a fixed tile sequence, and k_spectrum_reduce sums CTA partials in index order.
#include <cmath>
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace fava {

// ------------------------------------------------------------------------------------------------
// K4: weighting
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
    k_ke_weight3(const T* __restrict__ rho, const T* __restrict__ ux, const T* __restrict__ uy,
                 const T* __restrict__ uz, int64_t nrows, int64_t nx, int64_t pitch, double* __restrict__ wx,
                 double* __restrict__ wy, double* __restrict__ wz) {
    // one thread = two consecutive x of one row (nx is even); grid-stride over pairs
    const int64_t half = nx >> 1;
    const int64_t npairs = nrows * half;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < npairs; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = q / half, xp = (q - row * half) * 2;
        const int64_t in = row * nx + xp, out = row * pitch + xp;
        double r[2], a[2], b[2], c[2];
        VecLoad<T, 2>::ld(rho + in, r);
        VecLoad<T, 2>::ld(ux + in, a);
        VecLoad<T, 2>::ld(uy + in, b);
        VecLoad<T, 2>::ld(uz + in, c);
        const double s0 = sqrt(r[0]), s1 = sqrt(r[1]);
        *reinterpret_cast<double2*>(wx + out) = make_double2(s0 * a[0], s1 * a[1]);
        *reinterpret_cast<double2*>(wy + out) = make_double2(s0 * b[0], s1 * b[1]);
        *reinterpret_cast<double2*>(wz + out) = make_double2(s0 * c[0], s1 * c[1]);
    }
}

// ------------------------------------------------------------------------------------------------
// K6: power, projection and shell binning
// ------------------------------------------------------------------------------------------------
constexpr int kTS = 32;        // tile edge in kx and kz
constexpr int kBinT = 256;     // threads per CTA (8 warps, 4 kz rows each)
constexpr int kBinWarps = kBinT / 32;
constexpr int kSlots = 48;     // shell span of one tile (<= 31*sqrt(2) + 2 = 45.8)

struct BinParams {
    int n, pitch, ny_local, nbins, kmax2, npairs;  // pitch = complex elements per kx row (N/2 or N/2+1)
    int64_t ngroups;
    int64_t zstride;  // ny_local * pitch (complex elements per kz plane)
    const int32_t* ky_of_local;  // NULL = identity (local row jl holds global ky index jl)
    const int32_t* local_of_ky;  // NULL = identity
    double norm2;
};

__device__ __forceinline__ int shell_of(int k2) {
    // m such that m^2 - m < k2 <= m^2 + m  <=>  m - 1/2 < sqrt(k2) < m + 1/2  (k2 integer: no ties)
    int m = (int)(sqrt((double)k2) + 0.5);
    while (m * m + m < k2) ++m;
    while (m > 0 && m * m - m >= k2) --m;
    return m;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
    // 16-byte global -> shared copy that bypasses registers (LDGSTS); src-size 0 zero-fills the destination
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}

// dynamic shared memory: [3][nbins] CTA bins | [3][32][33] transposed operand tiles | [3][8][64] warp bins
//
// Work unit = a GROUP of up to eight 32x32 tiles that touch the same memory: for a ky row j >= 0 (and its mirror
// -ky, row jm) and an unordered pair {a, b} of 32-wide index ranges, the tiles (kx in A, |kz| in B) and
// (kx in B, |kz| in A), for kz >= 0 and kz < 0, in planes j and jm.  The direct operand of one member is the
// transposed operand of the next, so a CTA that walks a group back to back reads every element from DRAM once
// and finds it in L2 the second time (ncu: 32.4 GB -> see profiles/).
struct BinTile {
    int jl, jml, ky;  // local ky row of the points, local row of -ky, wavenumber
    int a, b;         // kx tile, |kz| tile
    int neg;          // 1: kz = -(32 b + i), else kz = 32 b + i
};

__global__ void __launch_bounds__(kBinT, 3)
    k_spectrum_bin(const double2* __restrict__ fx, const double2* __restrict__ fy, const double2* __restrict__ fz,
                   BinParams p, double* __restrict__ partial) {
    extern __shared__ __align__(16) unsigned char dyn_raw[];
    double* dyn = reinterpret_cast<double*>(dyn_raw);
    const int nb_pad = (3 * p.nbins + 1) & ~1;  // keep the tiles 16-byte aligned
    double* cta_tot = dyn;
    double* cta_lon = dyn + p.nbins;
    double* cta_cnt = dyn + 2 * p.nbins;
    typedef double2 Tile[kTS][kTS + 1];
    Tile* S = reinterpret_cast<Tile*>(dyn + nb_pad);
    double* wb = reinterpret_cast<double*>(S + 3);  // [3][kBinWarps][kSlots]
    double(*wb_tot)[kSlots] = reinterpret_cast<double(*)[kSlots]>(wb);
    double(*wb_lon)[kSlots] = reinterpret_cast<double(*)[kSlots]>(wb + kBinWarps * kSlots);
    double(*wb_cnt)[kSlots] = reinterpret_cast<double(*)[kSlots]>(wb + 2 * kBinWarps * kSlots);

    __shared__ BinTile list[8];
    __shared__ int list_n;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int i = t; i < 3 * p.nbins; i += kBinT) dyn[i] = 0.0;
    for (int i = t; i < 3 * kBinWarps * kSlots; i += kBinT) wb[i] = 0.0;
    __syncthreads();

    const int n = p.n, nh = n >> 1;
    const double2* F[3] = {fx, fy, fz};

    auto process = [&](const BinTile& T) {
        const int a0 = T.a * kTS, b0 = T.b * kTS;
        const int kzlo = T.neg ? max(b0, 1) : b0;
        const int mlo = shell_of(a0 * a0 + T.ky * T.ky + kzlo * kzlo);
        // ---- phase A: all three transposed operand tiles in flight (no registers held) -------------------
        // S[c][ia][ib] = stored value whose (conjugate, if kz < 0) is u^_c at (x-wn = kz, y-wn = ky, z-wn = kx),
        // kx = a0 + ia, |kz| = b0 + ib; lanes run over ib (contiguous x index in memory)
        {
            const int q = b0 + lane;  // |kz| handled by this lane while loading
            const bool qok = q < nh && !(T.neg && q == 0);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int ia = warp + kBinWarps * i;
                const int kxa = a0 + ia;
                // only elements inside the sphere are fetched (zero-filled otherwise): tiles on the surface load no more than they use
                const bool ok = qok && kxa < nh && kxa * kxa + T.ky * T.ky + q * q <= p.kmax2;
                int64_t off = 0;
                if (ok) off = T.neg ? (int64_t)((n - kxa) % n) * p.zstride + (int64_t)T.jml * p.pitch + q
                                    : (int64_t)kxa * p.zstride + (int64_t)T.jl * p.pitch + q;
#pragma unroll
                for (int c = 0; c < 3; ++c) cp_async16(&S[c][ia][lane], F[c] + off, ok);
            }
            asm volatile("cp.async.commit_group;\n" ::);
        }
        // ---- phase B: this thread's own points: kx = a0 + lane, |kz| = b0 + warp + 8 i -------------------
        const int kx = a0 + lane;
        double tot[4];
        {
            double2 d[3][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int q = b0 + warp + kBinWarps * i;
                const bool ok = kx < nh && q < nh && !(T.neg && q == 0) && kx * kx + T.ky * T.ky + q * q <= p.kmax2;
                const int l = T.neg ? n - q : q;
                const int64_t off = ok ? (int64_t)l * p.zstride + (int64_t)T.jl * p.pitch + kx : 0;
#pragma unroll
                for (int c = 0; c < 3; ++c) d[c][i] = ok ? __ldcs(F[c] + off) : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                tot[i] = 0.0;
#pragma unroll
                for (int c = 0; c < 3; ++c) tot[i] += d[c][i].x * d[c][i].x + d[c][i].y * d[c][i].y;
            }
        }
        asm volatile("cp.async.wait_group 0;\n" ::);
        __syncthreads();

        // ---- phase C: projection, shell index, warp-level segmented reduction ---------------------------
        // two rows at a time: their scans are independent dependency chains (shuffle -> add, five rounds) and interleave
#pragma unroll
        for (int ip = 0; ip < 4; ip += 2) {
            int key[2];
            double vt[2], vl[2];
            unsigned seg[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = ip + h;
                const int ib = warp + kBinWarps * i;
                const int q = b0 + ib;
                key[h] = -1, vt[h] = 0.0, vl[h] = 0.0;
                if (kx < nh && q < nh && !(T.neg && q == 0)) {
                    const int kz = T.neg ? -q : q;
                    const int k2 = kx * kx + T.ky * T.ky + kz * kz;
                    if (k2 <= p.kmax2) {
                        const double sgn = T.neg ? -1.0 : 1.0;  // conjugate of the folded half
                        const double2 t0 = S[0][lane][ib], t1 = S[1][lane][ib], t2 = S[2][lane][ib];
                        const double lre = fma((double)kx, t0.x, fma((double)T.ky, t1.x, (double)kz * t2.x));
                        const double lim = sgn * fma((double)kx, t0.y, fma((double)T.ky, t1.y, (double)kz * t2.y));
                        key[h] = shell_of(k2);
                        const double w = kx == 0 ? 1.0 : 2.0;
                        vt[h] = w * 0.5 * tot[i] * p.norm2;
                        vl[h] = k2 > 0 ? w * (lre * lre + lim * lim) / (double)k2 * p.norm2 : 0.0;
                    }
                }
                // lanes with my shell form one contiguous run (|k| grows with kx): its lane mask replaces the key exchange
                // of the segmented scan, marks the tail, and gives the point count of the run without summing it
                seg[h] = __match_any_sync(0xffffffffu, key[h]);
            }
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const double to = __shfl_up_sync(0xffffffffu, vt[h], d);
                    const double lo = __shfl_up_sync(0xffffffffu, vl[h], d);
                    if (lane >= d && ((seg[h] >> (lane - d)) & 1u)) vt[h] += to, vl[h] += lo;
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (key[h] >= 0 && lane == 31 - __clz(seg[h])) {  // tail of the run
                    const int s = min(key[h] - mlo, kSlots - 1);
                    wb_tot[warp][s] += vt[h];
                    wb_lon[warp][s] += vl[h];
                    wb_cnt[warp][s] += 2.0 * __popc(seg[h]) - ((a0 == 0 && (seg[h] & 1u)) ? 1.0 : 0.0);  // weight 1 at kx = 0
                }
                __syncwarp();  // two rows of a warp may end runs in the same slot
            }
        }
        __syncthreads();  // tiles consumed, warp bins complete
        if (t < kSlots) {
            double st = 0.0, sl = 0.0, sc = 0.0;
#pragma unroll
            for (int w = 0; w < kBinWarps; ++w) {
                st += wb_tot[w][t], sl += wb_lon[w][t], sc += wb_cnt[w][t];
                wb_tot[w][t] = 0.0, wb_lon[w][t] = 0.0, wb_cnt[w][t] = 0.0;
            }
            const int m = mlo + t;
            if (m < p.nbins && sc != 0.0) cta_tot[m] += st, cta_lon[m] += sl, cta_cnt[m] += sc;
        }
        // the next tile's first __syncthreads (after its loads) orders these bin updates
    };

    for (int64_t g = blockIdx.x; g < p.ngroups; g += gridDim.x) {
        const int pr = (int)(g % p.npairs);
        const int jl = (int)(g / p.npairs);  // positive-ky rows are the first npos local rows
        // unordered pair {a, b}, a <= b, from the linear index pr = b (b + 1) / 2 + a
        int b = (int)((sqrt(8.0 * pr + 1.0) - 1.0) * 0.5);
        while ((b + 1) * (b + 2) / 2 <= pr) ++b;
        while (b * (b + 1) / 2 > pr) --b;
        const int a = pr - b * (b + 1) / 2;
        const int j = p.ky_of_local ? p.ky_of_local[jl] : jl;
        if (j < 0 || j >= nh) continue;
        const int ky = j;
        if (a * a * kTS * kTS + b * b * kTS * kTS + ky * ky > p.kmax2) continue;  // whole group outside the sphere
        const int jm = (n - j) % n;
        const int jml = p.local_of_ky ? p.local_of_ky[jm] : jm;
        const bool self = jm == j;  // ky = 0
        // the members of the group in the order that makes one member's direct operand the next one's transposed operand;
        // ONE copy of the tile code walks the list (inlined eight times it was 11 k instructions and the kernel stalled
        // on instruction fetch: ncu no_instruction 1.2 per issue)
        // (the list lives in shared memory, written by one thread: as a per-thread array it went to local memory and
        // cost 1.6 GB of DRAM writes per launch at 1024^3)
        if (t == 0) {
            int nt = 0;
            auto push = [&](int jl_, int jml_, int ky_, int a_, int b_, int neg_) {
                list[nt].jl = jl_, list[nt].jml = jml_, list[nt].ky = ky_, list[nt].a = a_, list[nt].b = b_, list[nt].neg = neg_;
                ++nt;
            };
            push(jl, jml, ky, a, b, 0);                         // M1: plane +ky, (A, +B)
            if (a != b) push(jl, jml, ky, b, a, 0);             // M2: plane +ky, (B, +A)
            push(jl, jml, ky, a, b, 1);                         // M3: plane +ky, (A, -B)
            if (!self) push(jml, jl, -ky, b, a, 1);             // M4: plane -ky, (B, -A): its operands are M3's, swapped
            if (a != b) {
                push(jl, jml, ky, b, a, 1);                     // M5: plane +ky, (B, -A)
                if (!self) push(jml, jl, -ky, a, b, 1);         // M6: plane -ky, (A, -B)
            }  // a == b: M4 already covered plane -ky, (A, -A)
            if (!self) {
                push(jml, jl, -ky, a, b, 0);                    // M7: plane -ky, (A, +B)
                if (a != b) push(jml, jl, -ky, b, a, 0);        // M8: plane -ky, (B, +A)
            }
            list_n = nt;
        }
        __syncthreads();  // also orders the previous group's last bin merge before this group's first tile
        const int nt = list_n;
#pragma unroll 1
        for (int i = 0; i < nt; ++i) process(list[i]);
        __syncthreads();  // the list is free for the next group
    }
    __syncthreads();
    double* out = partial + (int64_t)blockIdx.x * 3 * p.nbins;
    for (int i = t; i < 3 * p.nbins; i += kBinT) out[i] = dyn[i];
}

__global__ void k_spectrum_reduce(const double* __restrict__ partial, int ncta, int nvals, double* __restrict__ sums) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nvals) return;
    double s = 0.0;
    for (int c = 0; c < ncta; ++c) s += partial[(int64_t)c * nvals + i];
    sums[i] = s;
}

// ------------------------------------------------------------------------------------------------
// cuFFT plans (cached in the context; one shared work area)
// ------------------------------------------------------------------------------------------------
enum PlanKind { PLAN_XY = 1, PLAN_Z = 2 };

static int get_plan(fava_ctx* ctx, int kind, int64_t a, int64_t b, int64_t c, cufftHandle* out) {
    const auto key = std::make_tuple(kind, a, b, c);
    auto it = ctx->plans.find(key);
    if (it != ctx->plans.end()) {
        *out = it->second;
        return FAVA_OK;
    }
    cufftHandle h;
    FAVA_CHECK_CUFFT(cufftCreate(&h));
    FAVA_CHECK_CUFFT(cufftSetAutoAllocation(h, 0));
    size_t work = 0;
    cufftResult r;
    if (kind == PLAN_XY) {  // a = batch (planes), b = ny, c = nx : in-place 2-D D2Z over padded rows
        long long dims[2] = {(long long)b, (long long)c};
        long long nxh = c / 2 + 1;
        long long inembed[2] = {(long long)b, 2 * nxh};
        long long onembed[2] = {(long long)b, nxh};
        r = cufftMakePlanMany64(h, 2, dims, inembed, 1, (long long)b * 2 * nxh, onembed, 1, (long long)b * nxh,
                                CUFFT_D2Z, (long long)a, &work);
    } else {  // a = nz (transform length), b = rows (stride and batch) : in-place strided 1-D Z2Z
        long long dims[1] = {(long long)a};
        long long embed[1] = {(long long)a};
        r = cufftMakePlanMany64(h, 1, dims, embed, (long long)b, 1, embed, (long long)b, 1, CUFFT_Z2Z, (long long)b,
                                &work);
    }
    if (r != CUFFT_SUCCESS) {
        cufftDestroy(h);
        return set_error(FAVA_ECUDA, "cufftMakePlanMany64(kind %d, %lld, %lld, %lld) failed: cufftResult %d", kind,
                         (long long)a, (long long)b, (long long)c, (int)r);
    }
    ctx->plans[key] = h;
    ctx->plan_work[key] = work;
    *out = h;
    return FAVA_OK;
}

static int exec_plan(fava_ctx* ctx, int kind, int64_t a, int64_t b, int64_t c, double* data, cudaStream_t st) {
    cufftHandle h = 0;
    int rc = get_plan(ctx, kind, a, b, c, &h);
    if (rc) return rc;
    const size_t work = ctx->plan_work[std::make_tuple(kind, a, b, c)];
    void* ws = nullptr;
    if (work) {
        rc = ctx_workspace(ctx, WS_FFT0, work, &ws);
        if (rc) return rc;
    }
    FAVA_CHECK_CUFFT(cufftSetWorkArea(h, ws));
    FAVA_CHECK_CUFFT(cufftSetStream(h, st));
    if (kind == PLAN_XY) FAVA_CHECK_CUFFT(cufftExecD2Z(h, (cufftDoubleReal*)data, (cufftDoubleComplex*)data));
    else FAVA_CHECK_CUFFT(cufftExecZ2Z(h, (cufftDoubleComplex*)data, (cufftDoubleComplex*)data, CUFFT_FORWARD));
    g_launches.fetch_add(1, std::memory_order_relaxed);  // counted once per library call (cuFFT may launch several)
    return FAVA_OK;
}

static int check_cube(int64_t n, const char* who) {
    if (n < 4 || (n & 1)) return set_error(FAVA_EINVAL, "%s: grid size %lld must be even and >= 4 (the reference's "
                                           "k-grid is integer only for even N, FlashUniform.py:244-253)", who, (long long)n);
    if (n > 4096) return set_error(FAVA_EINVAL, "%s: grid size %lld too large", who, (long long)n);
    return FAVA_OK;
}

}  // namespace fava

using namespace fava;

extern "C" {

int fava_ke_weight3(fava_ctx* ctx, const void* d_rho, const void* d_ux, const void* d_uy, const void* d_uz, int dtype,
                    int64_t nrows, int64_t nx, int64_t pitch, double* d_wx, double* d_wy, double* d_wz, void* stream) {
    FAVA_REQUIRE(ctx && d_rho && d_ux && d_uy && d_uz && d_wx && d_wy && d_wz, "fava_ke_weight3: NULL argument");
    FAVA_REQUIRE(nrows > 0 && nx > 0 && (nx & 1) == 0, "fava_ke_weight3: need nrows > 0 and an even nx");
    FAVA_REQUIRE(pitch >= nx && (pitch & 1) == 0, "fava_ke_weight3: pitch must be even and >= nx");
    FAVA_REQUIRE(dtype == FAVA_F32 || dtype == FAVA_F64, "fava_ke_weight3: bad dtype %d", dtype);
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t npairs = nrows * (nx / 2);
    const unsigned grid = (unsigned)std::min<int64_t>((npairs + 255) / 256, (int64_t)ctx->num_sms * 32);
    if (dtype == FAVA_F64)
        k_ke_weight3<double><<<grid, 256, 0, st>>>((const double*)d_rho, (const double*)d_ux, (const double*)d_uy,
                                                   (const double*)d_uz, nrows, nx, pitch, d_wx, d_wy, d_wz);
    else
        k_ke_weight3<float><<<grid, 256, 0, st>>>((const float*)d_rho, (const float*)d_ux, (const float*)d_uy,
                                                  (const float*)d_uz, nrows, nx, pitch, d_wx, d_wy, d_wz);
    FAVA_LAUNCHED();
    return FAVA_OK;
}

int64_t fava_spectral_pitch(int64_t n) { return fft_native_supported(n) ? n / 2 : n / 2 + 1; }

int fava_ke_transform_x(fava_ctx* ctx, const void* d_rho, const void* d_ux, const void* d_uy, const void* d_uz, int dtype,
                        int64_t nz_local, int64_t n, double* d_wx, double* d_wy, double* d_wz, void* stream) {
    FAVA_REQUIRE(ctx && d_rho && d_ux && d_uy && d_uz && d_wx && d_wy && d_wz, "fava_ke_transform_x: NULL argument");
    FAVA_REQUIRE(nz_local > 0, "fava_ke_transform_x: empty slab");
    FAVA_REQUIRE(dtype == FAVA_F32 || dtype == FAVA_F64, "fava_ke_transform_x: bad dtype %d", dtype);
    int rc = check_cube(n, "fava_ke_transform_x");
    if (rc) return rc;
    if (fft_native_supported(n))
        return fava_fft_x_weight3(ctx, d_rho, d_ux, d_uy, d_uz, dtype, nz_local * n, n, n / 2, d_wx, d_wy, d_wz, stream);
    return fava_ke_weight3(ctx, d_rho, d_ux, d_uy, d_uz, dtype, nz_local * n, n, 2 * (n / 2 + 1), d_wx, d_wy, d_wz, stream);
}

int fava_ke_transform_y(fava_ctx* ctx, double* d_w, int64_t nz_local, int64_t n, void* stream) {
    FAVA_REQUIRE(ctx && d_w, "fava_ke_transform_y: NULL argument");
    FAVA_REQUIRE(nz_local > 0, "fava_ke_transform_y: empty slab");
    int rc = check_cube(n, "fava_ke_transform_y");
    if (rc) return rc;
    if (fft_native_supported(n)) return fava_fft_cols(ctx, d_w, n, n / 2, n / 2, n, nz_local, 1, 1, nullptr, stream);
    DeviceGuard g(ctx->device);
    return exec_plan(ctx, PLAN_XY, nz_local, n, n, d_w, (cudaStream_t)stream);
}

int fava_ke_transform_z(fava_ctx* ctx, double* d_w, int64_t n, int64_t ny_local, const int32_t* d_ky_of_local,
                        void* stream) {
    FAVA_REQUIRE(ctx && d_w, "fava_ke_transform_z: NULL argument");
    FAVA_REQUIRE(ny_local > 0 && ny_local <= n, "fava_ke_transform_z: ny_local %lld not in 1..%lld", (long long)ny_local,
                 (long long)n);
    FAVA_REQUIRE(d_ky_of_local || ny_local == n, "fava_ke_transform_z: a partial ky range needs the ky map");
    int rc = check_cube(n, "fava_ke_transform_z");
    if (rc) return rc;
    if (fft_native_supported(n)) return fava_fft_cols(ctx, d_w, n, n / 2, n / 2, ny_local, n, 2, 2, d_ky_of_local, stream);
    DeviceGuard g(ctx->device);
    return exec_plan(ctx, PLAN_Z, n, ny_local * (n / 2 + 1), 0, d_w, (cudaStream_t)stream);
}

int fava_spectrum_bin(fava_ctx* ctx, const double* d_fx, const double* d_fy, const double* d_fz, int64_t n,
                      int64_t ny_local, const int32_t* d_ky_of_local, const int32_t* d_local_of_ky, double norm,
                      double* d_sums, void* stream) {
    FAVA_REQUIRE(ctx && d_fx && d_fy && d_fz && d_sums, "fava_spectrum_bin: NULL argument");
    int rc = check_cube(n, "fava_spectrum_bin");
    if (rc) return rc;
    FAVA_REQUIRE(ny_local > 0 && ny_local <= n, "fava_spectrum_bin: ny_local %lld not in 1..%lld", (long long)ny_local,
                 (long long)n);
    FAVA_REQUIRE((d_ky_of_local == nullptr) == (d_local_of_ky == nullptr),
                 "fava_spectrum_bin: pass both ky maps or neither");
    FAVA_REQUIRE(d_ky_of_local || ny_local == n, "fava_spectrum_bin: a partial ky range needs the ky maps");
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    BinParams p;
    p.n = (int)n, p.pitch = (int)fava_spectral_pitch(n), p.ny_local = (int)ny_local;
    p.nbins = (int)(n / 2 - 1);
    p.kmax2 = (int)(n * n / 4 - 3 * n / 2 + 2);
    const int nt = (int)((n / 2 + kTS - 1) / kTS);  // 32-wide tiles of kx and of |kz|
    p.npairs = nt * (nt + 1) / 2;
    // rows with ky >= 0 come first: all n/2 of them on one GPU (identity map), the first half of a rank's
    // +-ky symmetric set otherwise (fava_b200/spectrum.py:ky_ownership)
    const int64_t npos = d_ky_of_local ? (ny_local + 1) / 2 : n / 2;
    p.ngroups = npos * p.npairs;
    p.zstride = ny_local * (int64_t)p.pitch;
    p.ky_of_local = d_ky_of_local, p.local_of_ky = d_local_of_ky;
    p.norm2 = norm * norm;
    const size_t nb_pad = (size_t)((3 * p.nbins + 1) & ~1);
    const size_t dyn = sizeof(double) * (nb_pad + 3 * kBinWarps * kSlots) + 3 * sizeof(double2) * kTS * (kTS + 1);
    // three CTAs per SM (80 registers, 72 KB of shared memory at n = 1024).  Measured at 1024^3: with the tile code inlined
    // once per group member (11 k instructions, instruction-fetch stalls) 2 CTAs/SM took 6.1 ms and 3 took 6.5; with ONE
    // copy walked over the member list 2 / 3 CTAs per SM take 6.4 / 5.4 ms.
    const int ncta = (int)std::min<int64_t>(p.ngroups, (int64_t)ctx->num_sms * 3);
    void* ws;
    rc = ctx_workspace(ctx, WS_PARTIALS, sizeof(double) * 3 * (size_t)p.nbins * ncta, &ws);
    if (rc) return rc;
    FAVA_CHECK_CUDA(cudaFuncSetAttribute(k_spectrum_bin, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    k_spectrum_bin<<<ncta, kBinT, dyn, st>>>((const double2*)d_fx, (const double2*)d_fy, (const double2*)d_fz, p,
                                            (double*)ws);
    FAVA_LAUNCHED();
    const int nvals = 3 * p.nbins;
    k_spectrum_reduce<<<(nvals + 127) / 128, 128, 0, st>>>((const double*)ws, ncta, nvals, d_sums);
    FAVA_LAUNCHED();
    return FAVA_OK;
}

int fava_spectrum_finalize(fava_ctx* ctx, const double* d_sums, int64_t n, double* h_k, double* h_total,
                           double* h_long, double* h_trans, void* stream) {
    FAVA_REQUIRE(ctx && d_sums && h_k && h_total && h_long && h_trans, "fava_spectrum_finalize: NULL argument");
    int rc = check_cube(n, "fava_spectrum_finalize");
    if (rc) return rc;
    DeviceGuard g(ctx->device);
    const int nb = (int)(n / 2 - 1);
    std::vector<double> h((size_t)3 * nb);
    FAVA_CHECK_CUDA(cudaMemcpyAsync(h.data(), d_sums, sizeof(double) * 3 * nb, cudaMemcpyDeviceToHost,
                                    (cudaStream_t)stream));
    FAVA_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    const double two_pi_dm1 = 2.0 * M_PI * 2.0;  // 2 pi (ndim - 1), ndim = 3 (FlashUniform.py:295-297)
    for (int m = 0; m < nb; ++m) {
        const double k = (double)m;  // bin_edges[:-1] + 0.5 (FlashUniform.py:291)
        const double cnt = h[2 * nb + m];
        const double factor = k * k * two_pi_dm1;
        const double mt = h[m] / cnt, ml = h[nb + m] / cnt;  // 0/0 -> NaN like an empty scipy bin
        h_k[m] = k;
        h_total[m] = mt * factor;
        h_long[m] = ml * factor;
        h_trans[m] = (mt - ml) * factor;
    }
    return FAVA_OK;
}

int fava_ke_spectrum(fava_ctx* ctx, const void* d_rho, const void* d_ux, const void* d_uy, const void* d_uz,
                     int dtype, int64_t n, double* h_k, double* h_total, double* h_long, double* h_trans,
                     void* stream) {
    FAVA_REQUIRE(ctx && d_rho && d_ux && d_uy && d_uz, "fava_ke_spectrum: NULL argument");
    int rc = check_cube(n, "fava_ke_spectrum");
    if (rc) return rc;
    DeviceGuard g(ctx->device);
    const size_t comp_bytes = sizeof(double) * 2 * (size_t)(n * n * fava_spectral_pitch(n));
    void* w[3];
    for (int c = 0; c < 3; ++c) {
        rc = ctx_workspace(ctx, WS_FFT1 + c, comp_bytes, &w[c]);
        if (rc) return rc;
    }
    void* sums;
    rc = ctx_workspace(ctx, WS_AUX, sizeof(double) * 3 * (size_t)(n / 2 - 1), &sums);
    if (rc) return rc;
    rc = fava_ke_transform_x(ctx, d_rho, d_ux, d_uy, d_uz, dtype, n, n, (double*)w[0], (double*)w[1], (double*)w[2], stream);
    if (rc) return rc;
    for (int c = 0; c < 3; ++c) {
        rc = fava_ke_transform_y(ctx, (double*)w[c], n, n, stream);
        if (rc) return rc;
    }
    for (int c = 0; c < 3; ++c) {
        rc = fava_ke_transform_z(ctx, (double*)w[c], n, n, nullptr, stream);
        if (rc) return rc;
    }
    const double norm = 1.0 / ((double)n * (double)n * (double)n);  // norm="forward" (FlashUniform.py:268)
    rc = fava_spectrum_bin(ctx, (const double*)w[0], (const double*)w[1], (const double*)w[2], n, n, nullptr, nullptr,
                           norm, (double*)sums, stream);
    if (rc) return rc;
    return fava_spectrum_finalize(ctx, (const double*)sums, n, h_k, h_total, h_long, h_trans, stream);
}

}  // extern "C"
