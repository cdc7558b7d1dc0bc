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
//   K6  k_spectrum_bin |u^|^2 and the reference's longitudinal projection INCLUDING its `.T` quirk, shell index
//                      floor(|k|+1/2) in exact integer arithmetic, per-shell sums with weight 2 for the kx>0 half.
//       k_spectrum_reduce / fava_spectrum_finalize: fixed-order merge, shell mean x 4 pi k^2.
//
// The `.T` quirk without a transposed operand.  The reference forms, at every wavevector k = (kx,ky,kz),
//     longitudinal(k) = | sum_n k_n  u^_n(rev k) |^2 / |k|^2,   rev k = (kz,ky,kx)
// (FlashUniform.py:281: `k[n] * ffts[n, ...].T`, .T reverses all three axes - cubic grids only), and then takes the
// MEAN of it over each shell of |k| (:286-293).  rev is a bijection of the (cubic, fftshift-ed) index cube onto itself
// and |rev k| = |k|, so substituting p = rev k inside a shell sum gives
//     sum_{k in shell} longitudinal(k) = sum_{p in shell} | p_z u^_x(p) + p_y u^_y(p) + p_x u^_z(p) |^2 / |p|^2 :
// the same terms, every one computed from the values stored AT p alone.  The binning is therefore a single streaming
// pass that reads each element inside the spectral sphere exactly once (round 1 read every element twice, as a point and
// as another point's transposed operand, through paired 32x32 tiles transposed in shared memory, and needed +-ky on the
// same rank).  Only the order of the floating-point additions differs from the reference's (it already did).
//
// Why r2c is exact for this statistic: both `total` and the quirky `longitudinal` are invariant under
// k -> -k for a real input (u^(-k) = conj u^(k)), so the kx<0 half contributes the same values as its
// mirror; the Nyquist planes (|k_i| = N/2) lie beyond the last bin edge N/2-1.5 and never contribute.
//
// Determinism: a warp owns whole (kz,ky) rows in a fixed round-robin; its lanes hold consecutive kx, so shell
// indices are non-decreasing along the warp; a segmented shuffle scan reduces each run in a fixed order, run tails
// add into warp-private bins, the CTA merges its warps in a fixed order and k_spectrum_reduce sums the CTA
// partials in index order.
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
constexpr int kBinWarps = 4;   // warps per CTA, each with private shell bins (3 x nbins doubles)
constexpr int kBinT = kBinWarps * 32;
constexpr int kBinU = 4;       // 32-point chunks of a row in flight per warp (12 independent 16-byte loads per lane)

struct BinParams {
    int n, pitch, ny_local, nbins, kmax2;  // pitch = complex elements per kx row (N/2 or N/2+1)
    int64_t nrows;                         // n (kz) * ny_local
    const int32_t* ky_of_local;            // NULL = identity (local row jl holds global ky index jl); -1 = padding row
    double norm2;
};

__device__ __forceinline__ int shell_of(int k2) {
    // m such that m^2 - m < k2 <= m^2 + m  <=>  m - 1/2 < sqrt(k2) < m + 1/2  (k2 integer: no ties)
    int m = (int)(sqrt((double)k2) + 0.5);
    while (m * m + m < k2) ++m;
    while (m > 0 && m * m - m >= k2) --m;
    return m;
}

// dynamic shared memory: [kBinWarps][3][nbins] doubles (total, longitudinal, count per shell, per warp)
__global__ void __launch_bounds__(kBinT)
    k_spectrum_bin(const double2* __restrict__ fx, const double2* __restrict__ fy, const double2* __restrict__ fz,
                   BinParams p, double* __restrict__ partial) {
    extern __shared__ __align__(16) unsigned char dyn_raw[];
    double* dyn = reinterpret_cast<double*>(dyn_raw);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int i = t; i < kBinWarps * 3 * p.nbins; i += kBinT) dyn[i] = 0.0;
    __syncthreads();
    double* w_tot = dyn + (size_t)warp * 3 * p.nbins;
    double* w_lon = w_tot + p.nbins;
    double* w_cnt = w_lon + p.nbins;
    const int n = p.n, nh = n >> 1;

    const int64_t wstride = (int64_t)gridDim.x * kBinWarps;
    for (int64_t r = (int64_t)blockIdx.x * kBinWarps + warp; r < p.nrows; r += wstride) {
        const int zi = (int)(r / p.ny_local), jl = (int)(r - (int64_t)zi * p.ny_local);
        const int j = p.ky_of_local ? p.ky_of_local[jl] : jl;
        if (j < 0) continue;
        const int ky = j < nh ? j : j - n, kz = zi < nh ? zi : zi - n;
        const int rem = p.kmax2 - ky * ky - kz * kz;
        if (rem < 0) continue;  // the whole row lies outside the sphere (also the Nyquist rows)
        int kxmax = (int)sqrt((double)rem);
        while (kxmax * kxmax > rem) --kxmax;
        while ((kxmax + 1) * (kxmax + 1) <= rem) ++kxmax;
        const int64_t row = r * p.pitch;
        const int base2 = ky * ky + kz * kz;
        for (int c0 = 0; c0 <= kxmax; c0 += 32 * kBinU) {
            double2 a[kBinU][3];
#pragma unroll
            for (int u = 0; u < kBinU; ++u) {
                const int kx = c0 + 32 * u + lane;
                const bool ok = kx <= kxmax;
                a[u][0] = ok ? __ldcs(fx + row + kx) : make_double2(0.0, 0.0);
                a[u][1] = ok ? __ldcs(fy + row + kx) : make_double2(0.0, 0.0);
                a[u][2] = ok ? __ldcs(fz + row + kx) : make_double2(0.0, 0.0);
            }
            // two chunks at a time: their scans are independent dependency chains (shuffle -> add, five rounds) and interleave
#pragma unroll
            for (int up = 0; up < kBinU; up += 2) {
                if (c0 + 32 * up > kxmax) break;  // warp-uniform
                int key[2];
                double vt[2], vl[2];
                unsigned seg[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int u = up + h;
                    const int kx = c0 + 32 * u + lane;
                    key[h] = -1, vt[h] = 0.0, vl[h] = 0.0;
                    if (kx <= kxmax) {
                        const int k2 = kx * kx + base2;
                        const double2 x = a[u][0], y = a[u][1], z = a[u][2];
                        // | p_z u^_x + p_y u^_y + p_x u^_z |^2 / |p|^2 at p = (kx, ky, kz): the reference's `.T` projection, re-indexed
                        const double lre = fma((double)kz, x.x, fma((double)ky, y.x, (double)kx * z.x));
                        const double lim = fma((double)kz, x.y, fma((double)ky, y.y, (double)kx * z.y));
                        const double w = kx == 0 ? 1.0 : 2.0;  // the kx < 0 half mirrors the kx > 0 half
                        key[h] = shell_of(k2);
                        vt[h] = w * 0.5 * (x.x * x.x + x.y * x.y + y.x * y.x + y.y * y.y + z.x * z.x + z.y * z.y) * p.norm2;
                        vl[h] = k2 > 0 ? w * (lre * lre + lim * lim) / (double)k2 * p.norm2 : 0.0;
                    }
                    // lanes with my shell form one contiguous run (|k| grows with kx): its lane mask drives the segmented
                    // scan, marks the tail, and gives the point count of the run without summing it
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
                    if (key[h] >= 0 && key[h] < p.nbins && lane == 31 - __clz(seg[h])) {  // tail of the run
                        w_tot[key[h]] += vt[h];
                        w_lon[key[h]] += vl[h];
                        w_cnt[key[h]] += 2.0 * __popc(seg[h]) - ((c0 + 32 * (up + h) == 0 && (seg[h] & 1u)) ? 1.0 : 0.0);  // weight 1 at kx = 0
                    }
                    __syncwarp();  // the next chunk's runs may end in the same shells
                }
            }
        }
    }
    __syncthreads();
    double* out = partial + (int64_t)blockIdx.x * 3 * p.nbins;
    for (int i = t; i < 3 * p.nbins; i += kBinT) {
        double sacc = dyn[i];
#pragma unroll
        for (int w = 1; w < kBinWarps; ++w) sacc += dyn[(size_t)w * 3 * p.nbins + i];
        out[i] = sacc;
    }
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
                      int64_t ny_local, const int32_t* d_ky_of_local, double norm, double* d_sums, void* stream) {
    FAVA_REQUIRE(ctx && d_fx && d_fy && d_fz && d_sums, "fava_spectrum_bin: NULL argument");
    int rc = check_cube(n, "fava_spectrum_bin");
    if (rc) return rc;
    FAVA_REQUIRE(ny_local > 0 && ny_local <= n, "fava_spectrum_bin: ny_local %lld not in 1..%lld", (long long)ny_local,
                 (long long)n);
    FAVA_REQUIRE(d_ky_of_local || ny_local == n, "fava_spectrum_bin: a partial ky range needs the ky map");
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    BinParams p;
    p.n = (int)n, p.pitch = (int)fava_spectral_pitch(n), p.ny_local = (int)ny_local;
    p.nbins = (int)(n / 2 - 1);
    p.kmax2 = (int)(n * n / 4 - 3 * n / 2 + 2);
    p.nrows = n * ny_local;
    p.ky_of_local = d_ky_of_local;
    p.norm2 = norm * norm;
    const size_t dyn = sizeof(double) * (size_t)kBinWarps * 3 * p.nbins;
    FAVA_REQUIRE(dyn <= 200 * 1024, "fava_spectrum_bin: grid size %lld too large for the shell bins", (long long)n);
    FAVA_CHECK_CUDA(cudaFuncSetAttribute(k_spectrum_bin, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    int per_sm = 1;
    FAVA_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_spectrum_bin, kBinT, dyn));
    per_sm = std::max(1, std::min(per_sm, 4));
    const int ncta = (int)std::min<int64_t>((p.nrows + kBinWarps - 1) / kBinWarps, (int64_t)ctx->num_sms * per_sm);
    void* ws;
    rc = ctx_workspace(ctx, WS_PARTIALS, sizeof(double) * 3 * (size_t)p.nbins * ncta, &ws);
    if (rc) return rc;
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
    rc = fava_spectrum_bin(ctx, (const double*)w[0], (const double*)w[1], (const double*)w[2], n, n, nullptr, norm,
                           (double*)sums, stream);
    if (rc) return rc;
    return fava_spectrum_finalize(ctx, (const double*)sums, n, h_k, h_total, h_long, h_trans, stream);
}

}  // extern "C"
