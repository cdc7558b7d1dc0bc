// Hand-written fp64 line FFTs for the kinetic-energy spectrum (power-of-two N, 64 <= N <= 4096).
//
// The reference calls np.fft.fftn on complex128 N^3 arrays (fava/mesh/FLASH/FlashUniform.py:266-270).
// cuFFT is the default engine (any even N); its three passes are 55 % of the single-GPU step and each moves
// its data at only 3.2-4.7 TB/s.  These kernels (FAVA_FFT=native) do the same separable transform with less
// traffic and are parity-tested against the cuFFT path; they become the default once they beat it:
//   k_fft_x_weight  fuses K4 (w = sqrt(rho) u) INTO the x pass: a CTA reads two rows of rho,ux,uy,uz once,
//                   forms z_c = w_c[row] + i w_c[row+1] for the three components, transforms the three
//                   complex lines in shared memory and splits each into the two rows' Hermitian halves
//                   (two-for-one real FFT).  The weighted real arrays are never written to HBM:
//                   32 B read + 24 B written per cell instead of 56 + 48.
//   k_fft_cols      in-place complex FFT along a strided axis (y, then z): a CTA stages a tile of C adjacent
//                   columns x N rows in shared memory (64 B contiguous per row for C = 4), transforms the C
//                   lines and writes them back.  For the z pass, tiles whose columns all lie outside the
//                   spectral disc kx^2 + ky^2 <= (N/2-1.5)^2 are skipped (21 % of the columns): no bin can
//                   ever read them.
// In shared memory the transform is an in-place decimation-in-frequency FFT with register-resident radix-16
// butterflies (16 = 4 x 4), i.e. three passes over shared memory for N = 1024 (16,16,4); the output of an
// in-place DIF is digit-reversed, which costs nothing here: rows are written back to their true frequency
// (column passes) or gathered by frequency (x pass, with a skewed layout against bank conflicts).
// Twiddles come from one exp(-2 pi i m / N) table per N, computed on the host in long double.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace fava {

// ---- complex helpers -----------------------------------------------------------------------------------
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ double2 mul_mi(double2 a) { return make_double2(a.y, -a.x); }  // a * (-i)
__device__ __forceinline__ double2 mul_pi(double2 a) { return make_double2(-a.y, a.x); }  // a * (+i)

// forward DFTs in registers, natural order in and out: X[q] = sum_m v[m] exp(-2 pi i m q / R)
__device__ __forceinline__ void dft2(double2& a, double2& b) {
    const double2 t = a;
    a = cadd(t, b);
    b = csub(t, b);
}
__device__ __forceinline__ void dft4(double2& v0, double2& v1, double2& v2, double2& v3) {
    const double2 a = cadd(v0, v2), b = csub(v0, v2), c = cadd(v1, v3), d = csub(v1, v3);
    v0 = cadd(a, c);
    v2 = csub(a, c);
    v1 = cadd(b, mul_mi(d));
    v3 = cadd(b, mul_pi(d));
}

template <int R>
__device__ __forceinline__ void dft(double2 (&v)[R]);

template <>
__device__ __forceinline__ void dft<2>(double2 (&v)[2]) {
    dft2(v[0], v[1]);
}
template <>
__device__ __forceinline__ void dft<4>(double2 (&v)[4]) {
    dft4(v[0], v[1], v[2], v[3]);
}
template <>
__device__ __forceinline__ void dft<8>(double2 (&v)[8]) {
    // m = 2a + b : 4-point DFTs over a for b = 0 (even m) and b = 1 (odd m), then X[c + 4d] = y0[c] + (-1)^d w8^c y1[c]
    dft4(v[0], v[2], v[4], v[6]);
    dft4(v[1], v[3], v[5], v[7]);
    const double h = 0.70710678118654752440;
    const double2 y0[4] = {v[0], v[2], v[4], v[6]};
    double2 y1[4] = {v[1], v[3], v[5], v[7]};
    y1[1] = make_double2(h * (y1[1].x + y1[1].y), h * (y1[1].y - y1[1].x));    // * (1 - i)/sqrt2
    y1[2] = mul_mi(y1[2]);                                                     // * (-i)
    y1[3] = make_double2(h * (y1[3].y - y1[3].x), -h * (y1[3].x + y1[3].y));   // * (-1 - i)/sqrt2
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        v[c] = cadd(y0[c], y1[c]);
        v[c + 4] = csub(y0[c], y1[c]);
    }
}
template <>
__device__ __forceinline__ void dft<16>(double2 (&v)[16]) {
    // m = 4a + b, q = c + 4d : y[b][c] = DFT4 over a of v[4a+b];  y[b][c] *= w16^(b c);  X[c+4d] = DFT4 over b
    double2 y[4][4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        y[b][0] = v[b], y[b][1] = v[4 + b], y[b][2] = v[8 + b], y[b][3] = v[12 + b];
        dft4(y[b][0], y[b][1], y[b][2], y[b][3]);
    }
    // w16^k = exp(-2 pi i k / 16)
    const double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173, h = 0.70710678118654752440;
    const double2 w1 = make_double2(c1, -s1), w2 = make_double2(h, -h), w3 = make_double2(s1, -c1);
    const double2 w6 = make_double2(-h, -h), w9 = make_double2(-c1, s1);
    y[1][1] = cmul(y[1][1], w1);
    y[1][2] = cmul(y[1][2], w2);
    y[1][3] = cmul(y[1][3], w3);
    y[2][1] = cmul(y[2][1], w2);
    y[2][2] = mul_mi(y[2][2]);  // w16^4 = -i
    y[2][3] = cmul(y[2][3], w6);
    y[3][1] = cmul(y[3][1], w3);
    y[3][2] = cmul(y[3][2], w6);
    y[3][3] = cmul(y[3][3], w9);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        dft4(y[0][c], y[1][c], y[2][c], y[3][c]);
#pragma unroll
        for (int d = 0; d < 4; ++d) v[c + 4 * d] = y[d][c];
    }
}

// ---- pass plan: radix 16 while four or more bits remain, then one pass of 8 / 4 / 2 ----------------------
template <int LOGN>
struct FftPlan {
    static constexpr int N = 1 << LOGN;
    static constexpr int n16 = LOGN / 4;
    static constexpr int last = 1 << (LOGN % 4);  // 1 = no extra pass
    static constexpr int npass = n16 + (last > 1 ? 1 : 0);
    __host__ __device__ static constexpr int radix(int s) { return s < n16 ? 16 : last; }
};

// position (digit-reversed storage) -> frequency, and back
template <int LOGN>
__device__ __forceinline__ int pos_to_freq(int p) {
    using P = FftPlan<LOGN>;
    int k = 0, mult = 1, len = P::N;
#pragma unroll
    for (int s = 0; s < P::npass; ++s) {
        const int r = P::radix(s), per = len / r;
        const int q = p / per;
        p -= q * per;
        k += q * mult;
        mult *= r;
        len = per;
    }
    return k;
}
template <int LOGN>
__device__ __forceinline__ int freq_to_pos(int k) {
    using P = FftPlan<LOGN>;
    int p = 0, len = P::N;
#pragma unroll
    for (int s = 0; s < P::npass; ++s) {
        const int r = P::radix(s), per = len / r;
        p += (k % r) * per;
        k /= r;
        len = per;
    }
    return p;
}

// Skewed position of element n in a line-major (ROWS) buffer: +1 every 8 elements, +5 every 64.  With 16-byte
// elements a 128-byte bank row holds 8 of them; this skew makes the 8 lanes of every access phase hit 8 distinct
// bank groups in ALL the access patterns of the x pass (stride 1, stride 64 blocks walked block-fastest,
// stride-4 radix-4 groups, and the digit-reversed gather of the split, whose lane stride is 64) —
// ncu had shown 54 % of the shared-memory wavefronts of the un-skewed kernel to be bank conflicts.
__host__ __device__ __forceinline__ constexpr int skew(int n) { return n + (n >> 3) + 5 * (n >> 6); }

// One DIF pass of radix R on one line: butterfly jb of the line, elements at sm[line_off + pos(n) * sn].
// BLKFAST: consecutive jb walk the N/L blocks first (their address stride is odd in bank rows under `skew`).
template <int R, int LOGN, bool SKEW, bool BLKFAST>
__device__ __forceinline__ void fft_pass(double2* __restrict__ sm, int line_off, int sn, int L, int jb,
                                         const double2* __restrict__ tw, bool twiddle) {
    constexpr int N = 1 << LOGN;
    const int per = L / R;
    int blk, j;
    if (BLKFAST) {
        const int nblk = N / L;
        j = jb / nblk, blk = jb - j * nblk;
    } else {
        blk = jb / per, j = jb - blk * per;
    }
    const int base = blk * L + j;
    double2 v[R];
#pragma unroll
    for (int m = 0; m < R; ++m) {
        const int n = base + m * per;
        v[m] = sm[line_off + (SKEW ? skew(n) : n) * sn];
    }
    dft<R>(v);
    if (twiddle) {
        const int s = j * (N / L);
#pragma unroll
        for (int q = 1; q < R; ++q) v[q] = cmul(v[q], __ldg(tw + q * s));
    }
#pragma unroll
    for (int q = 0; q < R; ++q) {
        const int n = base + q * per;
        sm[line_off + (SKEW ? skew(n) : n) * sn] = v[q];
    }
}

// All passes of `nlines` lines held in shared memory.  COLS layout: element n of line c at sm[n*nlines + c]
// (threads walk c fastest); ROWS layout: sm[c*pitch + skew(n)] (threads walk butterflies fastest).
template <int LOGN, bool COLS>
__device__ __forceinline__ void fft_lines_smem(double2* __restrict__ sm, int nlines, int pitch,
                                               const double2* __restrict__ tw) {
    using P = FftPlan<LOGN>;
    constexpr int N = P::N;
    int L = N;
#pragma unroll
    for (int s = 0; s < P::npass; ++s) {
        const int r = P::radix(s);
        const int nb = N / r;  // butterflies per line
        const bool twd = s + 1 < P::npass;
        for (int item = threadIdx.x; item < nb * nlines; item += blockDim.x) {
            int c, jb;
            if (COLS) c = item % nlines, jb = item / nlines;
            else jb = item % nb, c = item / nb;
            const int off = COLS ? c : c * pitch;
            const int sn = COLS ? nlines : 1;
            // ROWS layout: the first pass walks butterflies j-fastest (stride 1), later passes block-fastest
            if (COLS || s == 0) {
                if (r == 16) fft_pass<16, LOGN, !COLS, false>(sm, off, sn, L, jb, tw, twd);
                else if (r == 8) fft_pass<8, LOGN, !COLS, false>(sm, off, sn, L, jb, tw, twd);
                else if (r == 4) fft_pass<4, LOGN, !COLS, false>(sm, off, sn, L, jb, tw, twd);
                else fft_pass<2, LOGN, !COLS, false>(sm, off, sn, L, jb, tw, twd);
            } else {
                if (r == 16) fft_pass<16, LOGN, !COLS, true>(sm, off, sn, L, jb, tw, twd);
                else if (r == 8) fft_pass<8, LOGN, !COLS, true>(sm, off, sn, L, jb, tw, twd);
                else if (r == 4) fft_pass<4, LOGN, !COLS, true>(sm, off, sn, L, jb, tw, twd);
                else fft_pass<2, LOGN, !COLS, true>(sm, off, sn, L, jb, tw, twd);
            }
        }
        __syncthreads();
        L /= r;
    }
}

// ---- x pass fused with the weighting --------------------------------------------------------------------
constexpr int kFftThreads = 256;

template <typename T, int LOGN>
__global__ void __launch_bounds__(kFftThreads, 2)
    k_fft_x_weight(const T* __restrict__ rho, const T* __restrict__ ux, const T* __restrict__ uy,
                   const T* __restrict__ uz, int64_t nrows, const double2* __restrict__ tw, double2* __restrict__ fx,
                   double2* __restrict__ fy, double2* __restrict__ fz) {
    constexpr int N = 1 << LOGN, NH = N / 2 + 1;
    constexpr int PITCH = skew(N - 1) + 2;  // skewed line length, padded
    extern __shared__ __align__(16) unsigned char fft_smem[];
    double2* sm = reinterpret_cast<double2*>(fft_smem);  // [3][PITCH]
    const int64_t r0 = (int64_t)blockIdx.x * 2;
    const bool two = r0 + 1 < nrows;
    const T* u[3] = {ux, uy, uz};
    // z_c[n] = w_c[r0][n] + i w_c[r0+1][n],  w = sqrt(rho) u; lane <-> n: coalesced loads, conflict-free stores
    for (int n = threadIdx.x; n < N; n += kFftThreads) {
        const double ra = (double)__ldcs(rho + r0 * N + n);
        const double rb = two ? (double)__ldcs(rho + (r0 + 1) * N + n) : 0.0;
        const double sa = sqrt(ra), sb = sqrt(rb);
        const int p = skew(n);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double a = (double)__ldcs(u[c] + r0 * N + n);
            const double b = two ? (double)__ldcs(u[c] + (r0 + 1) * N + n) : 0.0;
            sm[c * PITCH + p] = make_double2(sa * a, sb * b);
        }
    }
    __syncthreads();
    fft_lines_smem<LOGN, false>(sm, 3, PITCH, tw);
    // split: row r0 gets (Z[k] + conj Z[N-k]) / 2, row r0+1 gets (Z[k] - conj Z[N-k]) / (2i)
    double2* out[3] = {fx, fy, fz};
    for (int k = threadIdx.x; k < NH; k += kFftThreads) {
        const int ia = skew(freq_to_pos<LOGN>(k)), ib = skew(freq_to_pos<LOGN>((N - k) & (N - 1)));
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double2 A = sm[c * PITCH + ia], B = sm[c * PITCH + ib];
            const double2 e = make_double2(0.5 * (A.x + B.x), 0.5 * (A.y - B.y));
            const double2 o = make_double2(0.5 * (A.y + B.y), 0.5 * (B.x - A.x));
            out[c][r0 * NH + k] = e;
            if (two) out[c][(r0 + 1) * NH + k] = o;
        }
    }
}

// Persistent form of the fused x pass: one CTA per SM walks row pairs; the eight input rows of the NEXT pair
// (rho, ux, uy, uz x 2 rows = four contiguous 2N-element spans) are fetched by bulk copies (TMA, cp.async.bulk +
// mbarrier) into a raw staging buffer while the current pair is transformed and written, so the global loads
// overlap the shared-memory phases inside the CTA instead of relying on a second resident CTA.
template <typename T, int LOGN>
__global__ void __launch_bounds__(512, 1)
    k_fft_x_weight_tma(const T* __restrict__ rho, const T* __restrict__ ux, const T* __restrict__ uy,
                       const T* __restrict__ uz, int64_t nrows, const double2* __restrict__ tw, double2* __restrict__ fx,
                       double2* __restrict__ fy, double2* __restrict__ fz) {
    constexpr int N = 1 << LOGN, NH = N / 2 + 1;
    constexpr int PITCH = skew(N - 1) + 2;
    extern __shared__ __align__(16) unsigned char fft_smem[];
    double2* sm = reinterpret_cast<double2*>(fft_smem);                           // [3][PITCH]
    T* raw = reinterpret_cast<T*>(fft_smem + sizeof(double2) * 3 * PITCH);        // [4 fields][2 rows][N]
    uint64_t* bar = reinterpret_cast<uint64_t*>(fft_smem + sizeof(double2) * 3 * PITCH + sizeof(T) * 8 * N);
    const int nthreads = blockDim.x;
    const int64_t npairs = (nrows + 1) / 2;
    const T* src[4] = {rho, ux, uy, uz};

    auto issue = [&](int64_t pair) {  // thread 0 only
        const int64_t r0 = 2 * pair;
        const unsigned bytes = (unsigned)(min((int64_t)2, nrows - r0) * N * sizeof(T));
        mbar_expect_tx(bar, 4 * bytes);
#pragma unroll
        for (int f = 0; f < 4; ++f) bulk_load(raw + f * 2 * N, src[f] + r0 * N, bytes, bar);
    };

    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async;\n" ::: "memory");
        if ((int64_t)blockIdx.x < npairs) issue(blockIdx.x);
    }
    __syncthreads();

    unsigned parity = 0;
    double2* out[3] = {fx, fy, fz};
    for (int64_t pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
        const int64_t r0 = 2 * pair;
        const bool two = r0 + 1 < nrows;
        mbar_wait(bar, parity);
        parity ^= 1u;
        // z_c[n] = w_c[r0][n] + i w_c[r0+1][n],  w = sqrt(rho) u, from the staged rows
        for (int n = threadIdx.x; n < N; n += nthreads) {
            const double ra = (double)raw[n];
            const double rb = two ? (double)raw[N + n] : 0.0;
            const double sa = sqrt(ra), sb = sqrt(rb);
            const int p = skew(n);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const double a = (double)raw[(c + 1) * 2 * N + n];
                const double b = two ? (double)raw[(c + 1) * 2 * N + N + n] : 0.0;
                sm[c * PITCH + p] = make_double2(sa * a, sb * b);
            }
        }
        __syncthreads();  // the staged rows are consumed: refill them for the next pair while this one is transformed
        if (threadIdx.x == 0 && pair + gridDim.x < npairs) issue(pair + gridDim.x);
        fft_lines_smem<LOGN, false>(sm, 3, PITCH, tw);
        for (int k = threadIdx.x; k < NH; k += nthreads) {
            const int ia = skew(freq_to_pos<LOGN>(k)), ib = skew(freq_to_pos<LOGN>((N - k) & (N - 1)));
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const double2 A = sm[c * PITCH + ia], B = sm[c * PITCH + ib];
                const double2 e = make_double2(0.5 * (A.x + B.x), 0.5 * (A.y - B.y));
                const double2 o = make_double2(0.5 * (A.y + B.y), 0.5 * (B.x - A.x));
                __stcs(&out[c][r0 * NH + k], e);
                if (two) __stcs(&out[c][(r0 + 1) * NH + k], o);
            }
        }
        __syncthreads();  // the transform buffer is free again
    }
}

// ---- strided column pass ------------------------------------------------------------------------------------
// data: complex [nbatch][N][ncols]; transform along the middle axis for every (batch, column).
template <int LOGN, int C, int THREADS>
__global__ void __launch_bounds__(THREADS)
    k_fft_cols(double2* __restrict__ data, int64_t ncols, int64_t ntiles_per_batch, const double2* __restrict__ tw,
               int prune_kmax2, int prune_nxh, int prune_n, const int32_t* __restrict__ ky_of_local, int debug_nopass) {
    constexpr int N = 1 << LOGN;
    extern __shared__ __align__(16) unsigned char fft_smem[];
    double2* sm = reinterpret_cast<double2*>(fft_smem);  // [N][C]
    const int64_t batch = blockIdx.x / ntiles_per_batch;
    const int64_t c0 = (blockIdx.x - batch * ntiles_per_batch) * C;
    const int nc = (int)min((int64_t)C, ncols - c0);
    if (prune_kmax2 >= 0) {
        // columns are (ky_local, kx) pairs, kx fastest; skip the tile if every column is outside the disc
        const int64_t jl = c0 / prune_nxh;
        const int kx0 = (int)(c0 - jl * prune_nxh);
        const int64_t jl1 = (c0 + nc - 1) / prune_nxh;
        bool any = false;
        for (int64_t q = jl; q <= jl1; ++q) {
            const int j = ky_of_local ? ky_of_local[q] : (int)q;
            if (j < 0) continue;
            const int ky = j < prune_n / 2 ? j : j - prune_n;
            const int kx = q == jl ? kx0 : 0;
            if (kx * kx + ky * ky <= prune_kmax2) any = true;
        }
        if (!any) return;
    }
    double2* base = data + batch * (int64_t)N * ncols + c0;
    if (nc == C) {
        for (int item = threadIdx.x; item < N * C; item += THREADS) {
            const int c = item % C, n = item / C;
            sm[item] = __ldcs(base + (int64_t)n * ncols + c);
        }
    } else {
        for (int item = threadIdx.x; item < N * C; item += THREADS) {
            const int c = item % C, n = item / C;
            sm[item] = c < nc ? __ldcs(base + (int64_t)n * ncols + c) : make_double2(0.0, 0.0);
        }
    }
    __syncthreads();
    if (!debug_nopass) fft_lines_smem<LOGN, true>(sm, C, 0, tw);
    for (int item = threadIdx.x; item < N * C; item += THREADS) {
        const int c = item % C, p = item / C;
        if (c < nc) base[(int64_t)pos_to_freq<LOGN>(p) * ncols + c] = sm[item];
    }
}

// ---- host side ---------------------------------------------------------------------------------------------
static int get_twiddles(fava_ctx* ctx, int64_t n, const double2** out) {
    auto it = ctx->twiddles.find(n);
    if (it != ctx->twiddles.end()) {
        *out = (const double2*)it->second;
        return FAVA_OK;
    }
    std::vector<double2> h((size_t)n);
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int64_t m = 0; m < n; ++m) {
        const long double a = -two_pi * (long double)m / (long double)n;
        h[(size_t)m] = make_double2((double)cosl(a), (double)sinl(a));
    }
    void* d = nullptr;
    FAVA_CHECK_CUDA(cudaMalloc(&d, sizeof(double2) * (size_t)n));
    FAVA_CHECK_CUDA(cudaMemcpy(d, h.data(), sizeof(double2) * (size_t)n, cudaMemcpyHostToDevice));
    ctx->twiddles[n] = d;
    *out = (const double2*)d;
    return FAVA_OK;
}

static int ilog2_pow2(int64_t v) {
    if (v <= 0 || (v & (v - 1))) return -1;
    int l = 0;
    while ((int64_t(1) << l) < v) ++l;
    return l;
}

bool fft_native_supported(int64_t n) {
    // Opt-in (FAVA_FFT=native).  Measured at 1024^3 fp64 on B200 (tools/fft_bench.py, profiles/r01_fft_native.txt): the
    // fused x pass takes 16.9 ms with the TMA-fed persistent kernel (21.0 ms for the two-CTA kernel) against 21.9 ms
    // for K4 + cuFFT's x pass, but the strided passes (9.0 / 7.3 ms per component) are slower than cuFFT's 5.4 ms, and
    // cuFFT offers no efficient plan for the remaining (z, y) passes on their own (a rank-2 strided plan over the two
    // slow axes takes 59 ms per component), so the x pass cannot be combined with cuFFT's column passes either.
    const int l = ilog2_pow2(n);
    const char* e = getenv("FAVA_FFT");
    if (!e || e[0] != 'n') return false;
    return l >= 6 && l <= 12;
}

template <typename T, int LOGN>
static int launch_x(const T* rho, const T* ux, const T* uy, const T* uz, int64_t nrows, const double2* tw, double2* fx,
                    double2* fy, double2* fz, cudaStream_t st) {
    constexpr int N = 1 << LOGN;
    const size_t smem = sizeof(double2) * 3 * (skew(N - 1) + 2);
    FAVA_CHECK_CUDA(cudaFuncSetAttribute(k_fft_x_weight<T, LOGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_fft_x_weight<T, LOGN><<<(unsigned)((nrows + 1) / 2), kFftThreads, smem, st>>>(rho, ux, uy, uz, nrows, tw, fx, fy, fz);
    FAVA_LAUNCHED();
    return FAVA_OK;
}

// Persistent TMA-prefetching x pass when its buffers (transform lines + staged rows) fit one CTA's shared memory
// (N <= 1024 for f64 input, N <= 2048 for f32); FAVA_FFT_X=plain keeps the two-CTA kernel, FAVA_FFT_X_THREADS sets
// the CTA size (default 384).
template <typename T, int LOGN>
static int launch_x_tma(fava_ctx* ctx, const T* rho, const T* ux, const T* uy, const T* uz, int64_t nrows,
                        const double2* tw, double2* fx, double2* fy, double2* fz, cudaStream_t st, bool* done) {
    constexpr int N = 1 << LOGN;
    const size_t smem = sizeof(double2) * 3 * (skew(N - 1) + 2) + sizeof(T) * 8 * N + 16;
    *done = false;
    const char* e = getenv("FAVA_FFT_X");
    if (smem > 227 * 1024 || (e && e[0] == 'p')) return FAVA_OK;
    int threads = 384;  // measured: 256 / 384 / 512 threads -> 17.7 / 16.9 / 17.3 ms
    if (const char* t = getenv("FAVA_FFT_X_THREADS")) threads = atoi(t) == 256 ? 256 : (atoi(t) == 512 ? 512 : 384);
    FAVA_CHECK_CUDA(cudaFuncSetAttribute(k_fft_x_weight_tma<T, LOGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    FAVA_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fft_x_weight_tma<T, LOGN>, threads, smem));
    const int64_t npairs = (nrows + 1) / 2;
    const unsigned grid = (unsigned)std::min<int64_t>(npairs, (int64_t)ctx->num_sms * std::max(per_sm, 1));
    k_fft_x_weight_tma<T, LOGN><<<grid, threads, smem, st>>>(rho, ux, uy, uz, nrows, tw, fx, fy, fz);
    FAVA_LAUNCHED();
    *done = true;
    return FAVA_OK;
}

template <int LOGN, int C>
static int launch_cols_c(double2* data, int64_t ncols, int64_t nbatch, const double2* tw, int kmax2, int nxh, int n,
                         const int32_t* ky_of_local, cudaStream_t st) {
    constexpr int N = 1 << LOGN;
    const size_t smem = sizeof(double2) * (size_t)N * C;
    const int64_t tiles = (ncols + C - 1) / C;
    if (tiles * nbatch > 0x7fffffffLL) return set_error(FAVA_EINVAL, "fft_cols: too many tiles");
    constexpr int THREADS = (N / 16) * C < 256 ? 256 : ((N / 16) * C > 1024 ? 1024 : (N / 16) * C);
    FAVA_CHECK_CUDA(cudaFuncSetAttribute(k_fft_cols<LOGN, C, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const char* e = getenv("FAVA_FFT_DEBUG_NOPASS");
    k_fft_cols<LOGN, C, THREADS><<<(unsigned)(tiles * nbatch), THREADS, smem, st>>>(data, ncols, tiles, tw, kmax2, nxh, n,
                                                                             ky_of_local, e ? atoi(e) : 0);
    FAVA_LAUNCHED();
    return FAVA_OK;
}

template <int LOGN>
static int launch_cols(double2* data, int64_t ncols, int64_t nbatch, const double2* tw, int kmax2, int nxh, int n,
                       const int32_t* ky_of_local, cudaStream_t st) {
    constexpr int N = 1 << LOGN;
    if constexpr (N <= 1024) {
        const char* e = getenv("FAVA_FFT_C");
        if (e && atoi(e) == 8) return launch_cols_c<LOGN, 8>(data, ncols, nbatch, tw, kmax2, nxh, n, ky_of_local, st);
        if (e && atoi(e) == 2) return launch_cols_c<LOGN, 2>(data, ncols, nbatch, tw, kmax2, nxh, n, ky_of_local, st);
        return launch_cols_c<LOGN, 4>(data, ncols, nbatch, tw, kmax2, nxh, n, ky_of_local, st);
    } else if constexpr (N == 2048) {
        return launch_cols_c<LOGN, 2>(data, ncols, nbatch, tw, kmax2, nxh, n, ky_of_local, st);
    } else {
        return launch_cols_c<LOGN, 1>(data, ncols, nbatch, tw, kmax2, nxh, n, ky_of_local, st);
    }
}

#define FAVA_LOGN_SWITCH(l, CALL)                                                              \
    switch (l) {                                                                               \
        case 6: return CALL(6);                                                                \
        case 7: return CALL(7);                                                                \
        case 8: return CALL(8);                                                                \
        case 9: return CALL(9);                                                                \
        case 10: return CALL(10);                                                              \
        case 11: return CALL(11);                                                              \
        case 12: return CALL(12);                                                              \
        default: return set_error(FAVA_EINVAL, "native FFT: N = 2^%d is not supported", l);    \
    }

template <typename T, int LOGN>
static int launch_x_any(fava_ctx* ctx, const T* rho, const T* ux, const T* uy, const T* uz, int64_t nrows,
                        const double2* tw, double2* fx, double2* fy, double2* fz, cudaStream_t st) {
    bool done = false;
    const int rc = launch_x_tma<T, LOGN>(ctx, rho, ux, uy, uz, nrows, tw, fx, fy, fz, st, &done);
    if (rc || done) return rc;
    return launch_x<T, LOGN>(rho, ux, uy, uz, nrows, tw, fx, fy, fz, st);
}

template <typename T>
static int dispatch_x(fava_ctx* ctx, int l, const T* rho, const T* ux, const T* uy, const T* uz, int64_t nrows,
                      const double2* tw, double2* fx, double2* fy, double2* fz, cudaStream_t st) {
#define CALL_X(L) launch_x_any<T, L>(ctx, rho, ux, uy, uz, nrows, tw, fx, fy, fz, st)
    FAVA_LOGN_SWITCH(l, CALL_X)
#undef CALL_X
}

static int dispatch_cols(int l, double2* data, int64_t ncols, int64_t nbatch, const double2* tw, int kmax2, int nxh,
                         int n, const int32_t* ky_of_local, cudaStream_t st) {
#define CALL_C(L) launch_cols<L>(data, ncols, nbatch, tw, kmax2, nxh, n, ky_of_local, st)
    FAVA_LOGN_SWITCH(l, CALL_C)
#undef CALL_C
}

}  // namespace fava

using namespace fava;

extern "C" {

int fava_fft_native_supported(int64_t n) { return fft_native_supported(n) ? 1 : 0; }

int fava_fft_x_weight3(fava_ctx* ctx, const void* d_rho, const void* d_ux, const void* d_uy, const void* d_uz, int dtype,
                       int64_t nrows, int64_t nx, double* d_fx, double* d_fy, double* d_fz, void* stream) {
    FAVA_REQUIRE(ctx && d_rho && d_ux && d_uy && d_uz && d_fx && d_fy && d_fz, "fava_fft_x_weight3: NULL argument");
    FAVA_REQUIRE(nrows > 0, "fava_fft_x_weight3: no rows");
    FAVA_REQUIRE(dtype == FAVA_F32 || dtype == FAVA_F64, "fava_fft_x_weight3: bad dtype %d", dtype);
    const int l = ilog2_pow2(nx);
    FAVA_REQUIRE(l >= 6 && l <= 12, "fava_fft_x_weight3: nx = %lld is not a power of two in [64, 4096]", (long long)nx);
    DeviceGuard g(ctx->device);
    const double2* tw;
    int rc = get_twiddles(ctx, nx, &tw);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == FAVA_F64)
        return dispatch_x<double>(ctx, l, (const double*)d_rho, (const double*)d_ux, (const double*)d_uy, (const double*)d_uz,
                                  nrows, tw, (double2*)d_fx, (double2*)d_fy, (double2*)d_fz, st);
    return dispatch_x<float>(ctx, l, (const float*)d_rho, (const float*)d_ux, (const float*)d_uy, (const float*)d_uz, nrows, tw,
                             (double2*)d_fx, (double2*)d_fy, (double2*)d_fz, st);
}

int fava_fft_cols(fava_ctx* ctx, double* d_data, int64_t n, int64_t ncols, int64_t nbatch, int64_t prune_grid_n,
                  const int32_t* d_ky_of_local, void* stream) {
    FAVA_REQUIRE(ctx && d_data, "fava_fft_cols: NULL argument");
    FAVA_REQUIRE(ncols > 0 && nbatch > 0, "fava_fft_cols: bad shape");
    const int l = ilog2_pow2(n);
    FAVA_REQUIRE(l >= 6 && l <= 12, "fava_fft_cols: n = %lld is not a power of two in [64, 4096]", (long long)n);
    int kmax2 = -1, nxh = 0;
    if (prune_grid_n > 0) {
        nxh = (int)(prune_grid_n / 2 + 1);
        kmax2 = (int)(prune_grid_n * prune_grid_n / 4 - 3 * prune_grid_n / 2 + 2);
        FAVA_REQUIRE(ncols % nxh == 0, "fava_fft_cols: pruning needs columns = (ky rows) x (N/2+1)");
    }
    DeviceGuard g(ctx->device);
    const double2* tw;
    int rc = get_twiddles(ctx, n, &tw);
    if (rc) return rc;
    return dispatch_cols(l, (double2*)d_data, ncols, nbatch, tw, kmax2, nxh, (int)prune_grid_n, d_ky_of_local,
                         (cudaStream_t)stream);
}

}  // extern "C"
