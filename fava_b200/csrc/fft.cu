// Hand-written fp64 line FFTs of the kinetic-energy spectrum for power-of-two grids (256 <= N <= 2048).
//
// The reference calls np.fft.fftn on three complex128 N^3 arrays (fava/mesh/FLASH/FlashUniform.py:266-270).  Here the
// separable transform is three passes over Hermitian-half storage, each one read and one write of its data, built
// on the register-resident radix-16 core of fft_core.cuh (a line lives in the registers of N/16 threads and changes
// owner through shared memory twice):
//
//   k_fft_x_row /    x pass FUSED with the weighting w = sqrt(rho) u (FlashUniform.py:266-268).  k_fft_x_row (N >= 512): a
//   k_fft_x_weight   CTA owns ONE row, transformed as a half-length complex line (see the kernel).  k_fft_x_weight (N = 256):
//                    a CTA owns pairs of rows: the eight input rows rho,ux,uy,uz x 2 arrive by TMA bulk copies
//                    (cp.async.bulk + mbarrier) while the previous pair is transformed; z_c = w_c[row a] + i w_c[row b]
//                    for the three components is transformed in registers and split into the two rows' half spectra
//                    (two-for-one real transform).  The weighted real fields never touch HBM: 32 B read + 24 B
//                    written per cell.  The Nyquist column kx = N/2 is NOT stored: it lies beyond the last bin edge
//                    N/2 - 1.5 (FlashUniform.py:273-276), and without it a row is N/2 complex numbers = a whole number
//                    of 128-byte lines (pitch N/2+1 shifts every row by 16 bytes: measured 5.3 vs 3.4 ms per pass).
//   k_fft_cols       in-place transform along y or z.  One persistent CTA per SM (512 threads) walks tiles of
//                    C = 8192/N adjacent columns x N rows (128 KB): the NEXT tile lands in shared memory through a
//                    3-D tensor map (cp.async.bulk.tensor, boxes of C x 256 rows) while the current one is
//                    transformed in registers and stored straight from them.  Pruning: a z-pass tile whose columns
//                    all lie outside the spectral disc kx^2 + ky^2 <= kmax^2 is skipped, and both passes skip the
//                    OUTPUT rows no bin can read (ky^2 + kx0^2 > kmax^2, resp. kz^2 + ky^2 + kx0^2 > kmax^2): 11 % /
//                    35 % of the traffic of the y / z pass.  Unwritten elements keep stale values; the binning
//                    kernel never uses an element outside the sphere (spectrum.cu: k2 <= kmax2).
//                    On several GPUs the y pass IS the slab -> pencil exchange: its output rows go straight into the
//                    owners' peer-mapped receive buffers over NVLink (ColScatter), no send buffer, no pack kernel.
//
// Measured on B200 at 1024^3 (tools/lab/fft_lab.cu, profiles/r02_fft_lab*.txt): y pass 3.40 ms, z pass 2.58 ms per
// component against 5.39 ms for cuFFT's strided passes; the results agree with cuFFT to 6e-16 (max-norm).
// Other even N take the cuFFT path of spectrum.cu (the reference accepts any even N).
#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"
#include "fft_core.cuh"

namespace fava {

using namespace fftc;

// ---- twiddle tables ------------------------------------------------------------------------------------------
template <int LOGN>
static void fill_tables(std::vector<double2>& h) {
    using P = RegPlan<LOGN>;
    h.assign(P::T1_LEN + (P::T2_LEN > 0 ? P::T2_LEN : 1), make_double2(1.0, 0.0));
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int q = 0; q < 16; ++q)
        for (int j = 0; j < P::M1; ++j) {
            const long double a = -two_pi * (long double)((j * q) % P::N) / (long double)P::N;
            h[q * P::M1 + j] = make_double2((double)cosl(a), (double)sinl(a));
        }
    for (int q = 0; q < 16; ++q)
        for (int j = 0; j < P::M2; ++j) {
            const long double a = -two_pi * (long double)((j * q) % P::M1) / (long double)P::M1;
            h[P::T1_LEN + q * P::M2 + j] = make_double2((double)cosl(a), (double)sinl(a));
        }
}

static int get_tables(fava_ctx* ctx, int logn, const double2** t1, const double2** t2) {
    const int64_t n = int64_t(1) << logn;
    auto it = ctx->twiddles.find(n);
    if (it == ctx->twiddles.end()) {
        std::vector<double2> h;
        switch (logn) {
            case 8: fill_tables<8>(h); break;
            case 9: fill_tables<9>(h); break;
            case 10: fill_tables<10>(h); break;
            case 11: fill_tables<11>(h); break;
            default: return set_error(FAVA_EINVAL, "native FFT: N = 2^%d is not supported", logn);
        }
        void* d = nullptr;
        FAVA_CHECK_CUDA(cudaMalloc(&d, sizeof(double2) * h.size()));
        FAVA_CHECK_CUDA(cudaMemcpy(d, h.data(), sizeof(double2) * h.size(), cudaMemcpyHostToDevice));
        it = ctx->twiddles.emplace(n, d).first;
    }
    *t1 = (const double2*)it->second;
    *t2 = *t1 + 16 * (n / 16);
    return FAVA_OK;
}

// exp(-2 pi i k / N), k < N/4: the factors that turn the half-length transform of an even/odd-packed real row into the
// row's spectrum (k_fft_x_row)
static int get_unpack_table(fava_ctx* ctx, int64_t n, const double2** tw) {
    auto it = ctx->twiddles.find(-n);
    if (it == ctx->twiddles.end()) {
        std::vector<double2> h((size_t)(n / 4));
        const long double two_pi = 6.283185307179586476925286766559005768L;
        for (int64_t k = 0; k < n / 4; ++k) {
            const long double a = -two_pi * (long double)k / (long double)n;
            h[(size_t)k] = make_double2((double)cosl(a), (double)sinl(a));
        }
        void* d = nullptr;
        FAVA_CHECK_CUDA(cudaMalloc(&d, sizeof(double2) * h.size()));
        FAVA_CHECK_CUDA(cudaMemcpy(d, h.data(), sizeof(double2) * h.size(), cudaMemcpyHostToDevice));
        it = ctx->twiddles.emplace(-n, d).first;
    }
    *tw = (const double2*)it->second;
    return FAVA_OK;
}

// ---- async-copy wrappers not in common.cuh ---------------------------------------------------------------------
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_setup(uint64_t* bar) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async;\n" ::: "memory");
}

// ---- x pass fused with the weighting ---------------------------------------------------------------------------
template <typename T, int LOGN, int PAIRS>
struct XLayout {
    static constexpr int N = 1 << LOGN, M1 = N / 16, LINES = 3 * PAIRS, THREADS = LINES * M1, ROWS = 2 * PAIRS;
    static constexpr int LP = line_pitch(N);
    static constexpr size_t land_bytes = sizeof(T) * 4 * ROWS * N;       // rho, ux, uy, uz rows of the tile
    static constexpr size_t srho_bytes = sizeof(double) * ROWS * N;      // sqrt(rho), computed once per cell
    static constexpr size_t xb_bytes = sizeof(double) * LINES * LP;      // exchange words (8 bytes)
    static constexpr size_t total = land_bytes + srho_bytes + xb_bytes + 16;
};

template <typename T, int LOGN, int PAIRS, int CTAS>
__global__ void __launch_bounds__(XLayout<T, LOGN, PAIRS>::THREADS, CTAS)
    k_fft_x_weight(const T* __restrict__ rho, const T* __restrict__ ux, const T* __restrict__ uy, const T* __restrict__ uz,
                   int64_t ntiles, const double2* __restrict__ t1, const double2* __restrict__ t2, double2* __restrict__ fx,
                   double2* __restrict__ fy, double2* __restrict__ fz, int64_t out_pitch) {
    using L = XLayout<T, LOGN, PAIRS>;
    constexpr int N = L::N, M1 = L::M1;
    extern __shared__ __align__(128) unsigned char fft_smem[];
    T* land = reinterpret_cast<T*>(fft_smem);                                         // [4][ROWS][N]
    double* srho = reinterpret_cast<double*>(fft_smem + L::land_bytes);                // [ROWS][N]
    double* xb = reinterpret_cast<double*>(fft_smem + L::land_bytes + L::srho_bytes);   // [LINES][LP]
    uint64_t* bar = reinterpret_cast<uint64_t*>(fft_smem + L::land_bytes + L::srho_bytes + L::xb_bytes);
    const int line = threadIdx.x / M1, u = threadIdx.x - line * M1;
    const int pair = line / 3, comp = line - 3 * pair;
    const T* src[4] = {rho, ux, uy, uz};
    double2* out = comp == 0 ? fx : (comp == 1 ? fy : fz);
    const LineAddr at{line * L::LP};
    static_assert(M1 <= 32 || L::LINES <= 3, "named barriers 1..3");
    const LineSync<M1> line_sync{line};  // exchanges involve the threads of one line only

    auto issue = [&](int64_t t) {  // one thread: the 2 PAIRS rows of a tile are contiguous in every field
        constexpr unsigned bytes = (unsigned)(sizeof(T) * L::ROWS * N);
        mbar_expect_tx(bar, 4 * bytes);
#pragma unroll
        for (int f = 0; f < 4; ++f) bulk_load(land + f * L::ROWS * N, src[f] + t * L::ROWS * N, bytes, bar);
    };
    int64_t t = blockIdx.x;
    if (threadIdx.x == 0) {
        mbar_setup(bar);
        if (t < ntiles) issue(t);
    }
    __syncthreads();
    unsigned parity = 0;
    for (; t < ntiles; t += gridDim.x) {
        mbar_wait(bar, parity);
        parity ^= 1u;
        for (int i = threadIdx.x; i < L::ROWS * N; i += L::THREADS) srho[i] = sqrt((double)land[i]);
        __syncthreads();
        double2 v[16];
        {
            const double* sa = srho + (2 * pair) * N;
            const T* ua = land + ((comp + 1) * L::ROWS + 2 * pair) * N;
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const int n = u + M1 * m;
                v[m] = make_double2(sa[n] * (double)ua[n], sa[N + n] * (double)ua[N + n]);
            }
        }
        __syncthreads();  // landing rows and sqrt(rho) consumed: refill them while this tile is transformed
        if (threadIdx.x == 0 && t + gridDim.x < ntiles) issue(t + gridDim.x);
        double2 e[8], o[8];
        fft_regs_half<LOGN>(v, u, at, xb, t1, t2, line_sync);
        split_two_for_one<LOGN>(v, u, at, xb, e, o, line_sync);
        double2* oa = out + (t * PAIRS + pair) * 2 * out_pitch;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int k = u + M1 * m;  // k < N/2: the Nyquist column is not stored
            __stcs(oa + k, e[m]);
            __stcs(oa + out_pitch + k, o[m]);
        }
    }
}

// ---- x pass, one ROW per CTA (N >= 512): the real row as a complex transform of half its length -------------------------
// z[m] = w[2m] + i w[2m+1] (w = sqrt(rho) u, even/odd packing), Z = FFT_{N/2}(z), then with E = (Z[k] + conj Z[N/2-k]) / 2,
// O = (Z[k] - conj Z[N/2-k]) / (2i):  X[k] = E + w_N^k O  and  X[N/2-k] = conj(E - w_N^k O),  X[N/4] = conj Z[N/4].
// Against the two-rows-at-once kernel above: a line is N/32 threads (ONE warp at N = 1024: its exchanges need __syncwarp
// only), a CTA is the three components of one row (96 threads, 53 KB), four CTAs per SM run independently with their own
// TMA pipelines, and the transform has 10 % fewer flops.  Measured at 1024^3 fp64: 10.7 ms against 12.5 ms.
template <typename T, int LOGN>
struct XRowLayout {
    static constexpr int N = 1 << LOGN, H = N / 2, LOGH = LOGN - 1, M1 = H / 16, THREADS = 3 * M1;
    static constexpr int LP = line_pitch(H);
    static constexpr size_t land_bytes = sizeof(T) * 4 * N;
    static constexpr size_t srho_bytes = sizeof(double) * N;
    static constexpr size_t xb_bytes = sizeof(double) * 3 * LP;
    static constexpr size_t total = land_bytes + srho_bytes + xb_bytes + 16;
};
template <typename T> struct Vec2;
template <> struct Vec2<double> { typedef double2 type; };
template <> struct Vec2<float> { typedef float2 type; };

template <typename T, int LOGN, int CTAS>
__global__ void __launch_bounds__(XRowLayout<T, LOGN>::THREADS, CTAS)
    k_fft_x_row(const T* __restrict__ rho, const T* __restrict__ ux, const T* __restrict__ uy, const T* __restrict__ uz,
                int64_t nrows, const double2* __restrict__ t1, const double2* __restrict__ t2, const double2* __restrict__ tw,
                double2* __restrict__ fx, double2* __restrict__ fy, double2* __restrict__ fz, int64_t out_pitch) {
    using L = XRowLayout<T, LOGN>;
    constexpr int N = L::N, H = L::H, M1 = L::M1, LOGH = L::LOGH;
    typedef typename Vec2<T>::type T2;
    extern __shared__ __align__(128) unsigned char row_smem[];
    T* land = reinterpret_cast<T*>(row_smem);                                         // [4][N]
    double* srho = reinterpret_cast<double*>(row_smem + L::land_bytes);                // [N]
    double* xb = reinterpret_cast<double*>(row_smem + L::land_bytes + L::srho_bytes);   // [3][LP]
    uint64_t* bar = reinterpret_cast<uint64_t*>(row_smem + L::land_bytes + L::srho_bytes + L::xb_bytes);
    const int comp = threadIdx.x / M1, u = threadIdx.x - comp * M1;
    const T* src[4] = {rho, ux, uy, uz};
    double2* out = comp == 0 ? fx : (comp == 1 ? fy : fz);
    const LineAddr at{comp * L::LP};
    const LineSync<M1> line_sync{comp};

    auto issue = [&](int64_t r) {  // one thread
        constexpr unsigned bytes = (unsigned)(sizeof(T) * N);
        mbar_expect_tx(bar, 4 * bytes);
#pragma unroll
        for (int f = 0; f < 4; ++f) bulk_load(land + f * N, src[f] + r * N, bytes, bar);
    };
    int64_t row = blockIdx.x;
    if (threadIdx.x == 0) {
        mbar_setup(bar);
        if (row < nrows) issue(row);
    }
    __syncthreads();
    unsigned parity = 0;
    for (; row < nrows; row += gridDim.x) {
        mbar_wait(bar, parity);
        parity ^= 1u;
        for (int i = threadIdx.x; i < N; i += L::THREADS) srho[i] = sqrt((double)land[i]);
        __syncthreads();
        double2 v[16];
        {
            const double2* s2 = reinterpret_cast<const double2*>(srho);
            const T2* u2 = reinterpret_cast<const T2*>(land + (comp + 1) * N);
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const int idx = u + M1 * m;  // complex point idx = real samples 2 idx, 2 idx + 1
                const double2 s = s2[idx];
                const T2 a = u2[idx];
                v[m] = make_double2(s.x * (double)a.x, s.y * (double)a.y);
            }
        }
        __syncthreads();  // landing row and sqrt(rho) consumed: refill them while this row is transformed
        if (threadIdx.x == 0 && row + gridDim.x < nrows) issue(row + gridDim.x);
        double2 e[8], o[8], mid;
        // second exchange through warp shuffles (the M2 partner threads are adjacent lanes): 10.67 -> 10.09 ms at 1024^3
        fft_regs_half<LOGH, LineAddr, LineSync<M1>, 1>(v, u, at, xb, t1, t2, line_sync);
        split_two_for_one<LOGH>(v, u, at, xb, e, o, line_sync, &mid);
        double2* orow = out + row * out_pitch;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int k = u + M1 * m;  // k < N/4
            const double2 t = cmul(o[m], __ldg(tw + k));
            __stcs(orow + k, cadd(e[m], t));                                // X[k] = E + w^k O
            const double2 d = csub(e[m], t);
            if (k != 0) __stcs(orow + (H - k), make_double2(d.x, -d.y));     // X[N/2 - k] = conj(E - w^k O); k = 0 -> Nyquist, not stored
        }
        if (u == 0) __stcs(orow + H / 2, make_double2(mid.x, -mid.y));       // X[N/4] = conj Z[N/4]
    }
}

// ---- strided column pass ---------------------------------------------------------------------------------------
struct ColPrune {
    int mode;   // 0 none; 1 y pass: keep output rows with ky^2 + kx0^2 <= kmax2; 2 z pass: tile skip + output rows
    int kmax2;  // floor((N/2 - 1.5)^2)
    int n;      // grid size (wavenumber of index k: k < n/2 ? k : k - n)
    const int32_t* ky_of_batch;  // z pass: global ky index of batch row b (NULL = identity, -1 = padding row)
};

__device__ __forceinline__ int wavenumber(int k, int n) { return k < n / 2 ? k : k - n; }

// Slab -> ky-pencil exchange FUSED into the y pass (several GPUs): output row ky of local plane z is not written back
// in place but straight into the receive buffer of the rank that owns ky - a peer-mapped pointer (CUDA IPC, NVLink 5 /
// NVSwitch) or local memory - at [me nz_local + z][row of ky on its owner][kx] of that rank's complex
// [n][nyl][pitch] array, i.e. already in the layout the z pass and the binning kernel consume.  128-byte chunks per
// quarter-warp; the NVLink stores of one tile overlap the transform of the next.
struct ColScatter {
    double2* const* peer_recv;   // [nranks] receive buffers
    const int32_t* owner_of_ky;  // [n] rank that owns global ky row k (-1: nobody, the Nyquist row)
    const int32_t* row_of_ky;    // [n] row index of k inside its owner's ky set
    int me, nz_local, nyl, z_offset;  // z_offset: first local plane of this call (chunked slabs)
};

// data: complex [d2][d1][pitch]; LINE_DIM = 1: lines run along d1, batches are d2 (y pass);
//                                LINE_DIM = 2: lines run along d2, batches are d1 (z pass).
template <int LOGN, int LINE_DIM, bool SCATTER = false>
__global__ void __launch_bounds__(512, 1)
    k_fft_cols(const __grid_constant__ CUtensorMap tmap, double2* __restrict__ data, int64_t rstride, int64_t bstride,
               int ntile_cols, int64_t ntiles, const double2* __restrict__ t1, const double2* __restrict__ t2, ColPrune pr,
               ColScatter sc, unsigned long long* __restrict__ tile_counter) {
    static_assert(!SCATTER || LINE_DIM == 1, "the exchange is fused into the y pass");
    using P = RegPlan<LOGN>;
    using O = Owner<LOGN>;
    constexpr int N = P::N, C = 8192 / N;
    constexpr int BOX = N < 256 ? N : 256;  // rows per TMA box (<= 256)
    extern __shared__ __align__(1024) unsigned char col_smem[];
    double2* land = reinterpret_cast<double2*>(col_smem);                        // [N][C] complex, rows of 16 C bytes
    double* xb = reinterpret_cast<double*>(col_smem + sizeof(double2) * N * C);  // [N][C] exchange words (8 bytes)
    uint64_t* bar = reinterpret_cast<uint64_t*>(col_smem + sizeof(double2) * N * C + sizeof(double) * N * C);
    int64_t* next_slot = reinterpret_cast<int64_t*>(bar + 2);  // [2]: the tile a CTA works on / the one being prefetched
    double2** dst_row = reinterpret_cast<double2**>(col_smem + sizeof(double2) * N * C + sizeof(double) * N * C + 64);  // [N]
    const int c = threadIdx.x % C, u = threadIdx.x / C;
    const ColAddr<C> at{c};
    if (SCATTER) {  // where output row ky of plane 0 goes (rstride = pitch in the y pass)
        for (int k = threadIdx.x; k < N; k += blockDim.x) {
            const int r = sc.owner_of_ky[k];
            dst_row[k] = r < 0 ? nullptr
                               : sc.peer_recv[r] + (((int64_t)sc.me * sc.nz_local + sc.z_offset) * sc.nyl + sc.row_of_ky[k]) * rstride;
        }
    }

    auto tile_k2 = [&](int64_t t, int64_t* b_out, int* kx0_out) {  // kx0^2 (+ ky^2 in the z pass); -1 = skip the tile
        const int64_t b = t / ntile_cols;
        const int kx0 = (int)(t - b * ntile_cols) * C;
        *b_out = b, *kx0_out = kx0;
        int k2 = kx0 * kx0;
        if (pr.mode == 2) {
            const int j = pr.ky_of_batch ? pr.ky_of_batch[b] : (int)b;
            if (j < 0) return -1;
            const int ky = wavenumber(j, pr.n);
            k2 += ky * ky;
            if (k2 > pr.kmax2) return -1;
        }
        return k2;
    };
    // Tiles are handed out by a global counter (one atomic per tile, by the CTA's elected thread): a CTA that starts late
    // because another kernel still holds its SM simply takes fewer tiles, and pruned tiles cost nobody a turn.
    auto next_tile = [&]() {  // one thread: next tile that is not skipped, ntiles when the work is exhausted
        int64_t b;
        int kx0;
        for (;;) {
            const int64_t t = (int64_t)atomicAdd(tile_counter, 1ull);
            if (t >= ntiles) return ntiles;
            if (tile_k2(t, &b, &kx0) >= 0) return t;
        }
    };
    auto issue = [&](int64_t t) {  // one thread
        const int64_t b = t / ntile_cols;
        const int ct = (int)(t - b * ntile_cols);
        mbar_expect_tx(bar, (unsigned)(sizeof(double2) * N * C));
#pragma unroll
        for (int i = 0; i < N / BOX; ++i) {
            if (LINE_DIM == 1) tma_load_3d(land + i * BOX * C, &tmap, ct * 2 * C, i * BOX, (int)b, bar);
            else tma_load_3d(land + i * BOX * C, &tmap, ct * 2 * C, (int)b, i * BOX, bar);
        }
    };

    if (threadIdx.x == 0) {
        mbar_setup(bar);
        const int64_t t0 = next_tile();
        next_slot[0] = t0;
        if (t0 < ntiles) issue(t0);
    }
    __syncthreads();
    unsigned parity = 0;
    for (int it = 0;; ++it) {
        const int64_t t = next_slot[it & 1];
        if (t >= ntiles) break;
        int64_t b;
        int kx0;
        const int base2 = tile_k2(t, &b, &kx0);
        mbar_wait(bar, parity);
        parity ^= 1u;
        double2 v[16];
#pragma unroll
        for (int m = 0; m < 16; ++m) v[m] = land[O::in_index(u, m) * C + c];
        __syncthreads();  // landing buffer consumed (and the previous tile's exchange reads are complete)
        if (threadIdx.x == 0) {  // the other threads read the slot in the next iteration, several barriers from here
            const int64_t tn = next_tile();
            next_slot[(it + 1) & 1] = tn;
            if (tn < ntiles) issue(tn);
        }
        // second exchange through warp shuffles in the y pass (M2 C = 32 lanes hold one exchange group; measured
        // 2.95 -> 2.88 ms at 1024^3, profiles/r02_fft_lab5_*.txt); the z pass measured no gain and keeps shared memory
        fft_regs_half<LOGN, ColAddr<C>, CtaSync, LINE_DIM == 1 ? C : 0>(v, u, at, xb, t1, t2);
        if (SCATTER) {
            const int64_t zoff = b * (int64_t)sc.nyl * rstride + kx0 + c;  // plane b of the destination's [z][row][kx]
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const int k = O::out_freq(u, r);
                const int w = wavenumber(k, pr.n);
                double2* row = dst_row[k];
                if (row != nullptr && w * w + base2 <= pr.kmax2) row[zoff] = v[r];
            }
        } else {
            double2* base = data + b * bstride + kx0 + c;
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const int k = O::out_freq(u, r);
                bool keep = true;
                if (pr.mode) {
                    const int w = wavenumber(k, pr.n);
                    keep = w * w + base2 <= pr.kmax2;
                }
                if (keep) __stcs(base + (int64_t)k * rstride, v[r]);
            }
        }
    }
    if (SCATTER) __threadfence_system();  // peer stores ordered before the kernel's completion is observed
}

// ---- host side -------------------------------------------------------------------------------------------------
static int ilog2_pow2(int64_t v) {
    if (v <= 0 || (v & (v - 1))) return -1;
    int l = 0;
    while ((int64_t(1) << l) < v) ++l;
    return l;
}

bool fft_native_supported(int64_t n) {
    const int l = ilog2_pow2(n);
    return l >= 8 && l <= 11;
}

// tensor map of complex [d2][d1][pitch] seen as doubles [d2][d1][2 pitch]; box = C complex x `rows` along the line dim
static int get_tensor_map(fava_ctx* ctx, double2* data, int64_t pitch, int64_t d1, int64_t d2, int line_dim, int C, int rows,
                          CUtensorMap* out) {
    const uint64_t dims[3] = {(uint64_t)(2 * pitch), (uint64_t)d1, (uint64_t)d2};
    const uint64_t strides[2] = {(uint64_t)(pitch * 16), (uint64_t)(pitch * 16 * d1)};
    const uint32_t box[3] = {(uint32_t)(2 * C), line_dim == 1 ? (uint32_t)rows : 1u, line_dim == 2 ? (uint32_t)rows : 1u};
    return ctx_tensor_map(ctx, data, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, dims, strides, box, out);
}

template <typename T, int LOGN, int PAIRS, int CTAS>
static int launch_x(fava_ctx* ctx, const T* rho, const T* ux, const T* uy, const T* uz, int64_t nrows, const double2* t1,
                    const double2* t2, double2* fx, double2* fy, double2* fz, int64_t pitch, cudaStream_t st) {
    using L = XLayout<T, LOGN, PAIRS>;
    static_assert(L::total <= 227 * 1024, "x pass: shared memory");
    if (nrows % L::ROWS) return set_error(FAVA_EINVAL, "fava_fft_x_weight3: %lld rows are not a multiple of %d", (long long)nrows, L::ROWS);
    auto kern = k_fft_x_weight<T, LOGN, PAIRS, CTAS>;
    FAVA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::total));
    const int64_t ntiles = nrows / L::ROWS;
    const unsigned grid = (unsigned)std::min<int64_t>(ntiles, (int64_t)ctx->num_sms * CTAS);
    kern<<<grid, L::THREADS, L::total, st>>>(rho, ux, uy, uz, ntiles, t1, t2, fx, fy, fz, pitch);
    FAVA_LAUNCHED();
    return FAVA_OK;
}

template <typename T, int LOGN, int CTAS>
static int launch_x_row(fava_ctx* ctx, const T* rho, const T* ux, const T* uy, const T* uz, int64_t nrows, double2* fx,
                        double2* fy, double2* fz, int64_t pitch, cudaStream_t st) {
    using L = XRowLayout<T, LOGN>;
    static_assert(L::total * CTAS <= 227 * 1024, "x pass: shared memory");
    const double2 *t1, *t2, *tw;  // tables of the half-length plan + the unpack factors of the full length
    int rc = get_tables(ctx, LOGN - 1, &t1, &t2);
    if (rc) return rc;
    rc = get_unpack_table(ctx, L::N, &tw);
    if (rc) return rc;
    auto kern = k_fft_x_row<T, LOGN, CTAS>;
    FAVA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::total));
    const unsigned grid = (unsigned)std::min<int64_t>(nrows, (int64_t)ctx->num_sms * CTAS);
    kern<<<grid, L::THREADS, L::total, st>>>(rho, ux, uy, uz, nrows, t1, t2, tw, fx, fy, fz, pitch);
    FAVA_LAUNCHED();
    return FAVA_OK;
}

template <typename T>
static int dispatch_x(fava_ctx* ctx, int l, const T* rho, const T* ux, const T* uy, const T* uz, int64_t nrows, const double2* t1,
                      const double2* t2, double2* fx, double2* fy, double2* fz, int64_t pitch, cudaStream_t st) {
    switch (l) {  // N = 256: four row pairs per CTA (the half-length plan starts at 256 points); N >= 512: one row per CTA
        case 8: return launch_x<T, 8, 4, 2>(ctx, rho, ux, uy, uz, nrows, t1, t2, fx, fy, fz, pitch, st);
        case 9: return launch_x_row<T, 9, 8>(ctx, rho, ux, uy, uz, nrows, fx, fy, fz, pitch, st);
        case 10: return launch_x_row<T, 10, 4>(ctx, rho, ux, uy, uz, nrows, fx, fy, fz, pitch, st);
        case 11: return launch_x_row<T, 11, 2>(ctx, rho, ux, uy, uz, nrows, fx, fy, fz, pitch, st);
        default: return set_error(FAVA_EINVAL, "native FFT: N = 2^%d is not supported", l);
    }
}

template <int LOGN>
static int launch_cols(fava_ctx* ctx, double2* data, int64_t pitch, int64_t ncols, int64_t d1, int64_t d2, int line_dim,
                       const double2* t1, const double2* t2, ColPrune pr, cudaStream_t st, const ColScatter* sc = nullptr,
                       int max_ctas = 0) {
    constexpr int N = 1 << LOGN, C = 8192 / N, BOX = N < 256 ? N : 256;
    if (ncols % C) return set_error(FAVA_EINVAL, "fava_fft_cols: %lld columns are not a multiple of %d", (long long)ncols, C);
    const int ntc = (int)(ncols / C);
    const int64_t nbatch = line_dim == 1 ? d2 : d1;
    const int64_t ntiles = nbatch * ntc;
    const int64_t rstride = line_dim == 1 ? pitch : pitch * d1, bstride = line_dim == 1 ? pitch * d1 : pitch;
    const size_t smem = sizeof(double2) * N * C + sizeof(double) * N * C + 64 + (sc ? sizeof(double2*) * N : 0);
    if (smem > 227 * 1024) return set_error(FAVA_EINVAL, "fava_fft_cols: n = %d needs %zu bytes of shared memory", N, smem);
    CUtensorMap map;
    int rc = get_tensor_map(ctx, data, pitch, d1, d2, line_dim, C, BOX, &map);
    if (rc) return rc;
    int64_t ctas = ctx->num_sms;
    if (max_ctas > 0) ctas = std::min<int64_t>(ctas, max_ctas);
    const unsigned grid = (unsigned)std::min<int64_t>(ntiles, ctas);
    const ColScatter none = {};
    // a zeroed tile counter per launch: 64 counters used in turn, so that launches in flight on several streams differ
    if (!ctx->tile_counters) FAVA_CHECK_CUDA(cudaMalloc(&ctx->tile_counters, sizeof(unsigned long long) * 64));
    unsigned long long* counter = (unsigned long long*)ctx->tile_counters + (ctx->tile_counter_next++ & 63);
    FAVA_CHECK_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st));
    if (sc) {
        auto kern = k_fft_cols<LOGN, 1, true>;
        FAVA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, 512, smem, st>>>(map, data, rstride, bstride, ntc, ntiles, t1, t2, pr, *sc, counter);
    } else if (line_dim == 1) {
        auto kern = k_fft_cols<LOGN, 1>;
        FAVA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, 512, smem, st>>>(map, data, rstride, bstride, ntc, ntiles, t1, t2, pr, none, counter);
    } else {
        auto kern = k_fft_cols<LOGN, 2>;
        FAVA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, 512, smem, st>>>(map, data, rstride, bstride, ntc, ntiles, t1, t2, pr, none, counter);
    }
    FAVA_LAUNCHED();
    return FAVA_OK;
}

}  // namespace fava

using namespace fava;

extern "C" {

int fava_fft_native_supported(int64_t n) { return fft_native_supported(n) ? 1 : 0; }

int fava_fft_x_weight3(fava_ctx* ctx, const void* d_rho, const void* d_ux, const void* d_uy, const void* d_uz, int dtype,
                       int64_t nrows, int64_t nx, int64_t pitch, double* d_fx, double* d_fy, double* d_fz, void* stream) {
    FAVA_REQUIRE(ctx && d_rho && d_ux && d_uy && d_uz && d_fx && d_fy && d_fz, "fava_fft_x_weight3: NULL argument");
    FAVA_REQUIRE(nrows > 0, "fava_fft_x_weight3: no rows");
    FAVA_REQUIRE(dtype == FAVA_F32 || dtype == FAVA_F64, "fava_fft_x_weight3: bad dtype %d", dtype);
    FAVA_REQUIRE(fft_native_supported(nx), "fava_fft_x_weight3: nx = %lld is not a power of two in [256, 2048]", (long long)nx);
    FAVA_REQUIRE(pitch >= nx / 2, "fava_fft_x_weight3: pitch %lld < nx/2", (long long)pitch);
    const int l = ilog2_pow2(nx);
    DeviceGuard g(ctx->device);
    const double2 *t1, *t2;
    int rc = get_tables(ctx, l, &t1, &t2);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == FAVA_F64)
        return dispatch_x<double>(ctx, l, (const double*)d_rho, (const double*)d_ux, (const double*)d_uy, (const double*)d_uz,
                                  nrows, t1, t2, (double2*)d_fx, (double2*)d_fy, (double2*)d_fz, pitch, st);
    return dispatch_x<float>(ctx, l, (const float*)d_rho, (const float*)d_ux, (const float*)d_uy, (const float*)d_uz, nrows, t1,
                             t2, (double2*)d_fx, (double2*)d_fy, (double2*)d_fz, pitch, st);
}

int fava_fft_cols(fava_ctx* ctx, double* d_data, int64_t n, int64_t pitch, int64_t ncols, int64_t d1, int64_t d2,
                  int line_dim, int prune_mode, const int32_t* d_ky_of_batch, void* stream) {
    FAVA_REQUIRE(ctx && d_data, "fava_fft_cols: NULL argument");
    FAVA_REQUIRE(line_dim == 1 || line_dim == 2, "fava_fft_cols: line_dim must be 1 (y) or 2 (z)");
    FAVA_REQUIRE(d1 > 0 && d2 > 0 && (line_dim == 1 ? d1 : d2) == n, "fava_fft_cols: the transformed dimension must have length n");
    FAVA_REQUIRE(ncols > 0 && ncols <= pitch, "fava_fft_cols: need 0 < ncols <= pitch");
    FAVA_REQUIRE(prune_mode >= 0 && prune_mode <= 2, "fava_fft_cols: bad prune mode %d", prune_mode);
    FAVA_REQUIRE(fft_native_supported(n), "fava_fft_cols: n = %lld is not a power of two in [256, 2048]", (long long)n);
    FAVA_REQUIRE(d1 < (int64_t(1) << 31) && d2 < (int64_t(1) << 31), "fava_fft_cols: dimension too large");
    const int l = ilog2_pow2(n);
    DeviceGuard g(ctx->device);
    const double2 *t1, *t2;
    int rc = get_tables(ctx, l, &t1, &t2);
    if (rc) return rc;
    ColPrune pr;
    pr.mode = prune_mode, pr.n = (int)n, pr.kmax2 = (int)(n * n / 4 - 3 * n / 2 + 2), pr.ky_of_batch = d_ky_of_batch;
    cudaStream_t st = (cudaStream_t)stream;
    switch (l) {
        case 8: return launch_cols<8>(ctx, (double2*)d_data, pitch, ncols, d1, d2, line_dim, t1, t2, pr, st);
        case 9: return launch_cols<9>(ctx, (double2*)d_data, pitch, ncols, d1, d2, line_dim, t1, t2, pr, st);
        case 10: return launch_cols<10>(ctx, (double2*)d_data, pitch, ncols, d1, d2, line_dim, t1, t2, pr, st);
        default: return launch_cols<11>(ctx, (double2*)d_data, pitch, ncols, d1, d2, line_dim, t1, t2, pr, st);
    }
}

int fava_fft_y_scatter(fava_ctx* ctx, double* d_data, int64_t n, int64_t nz_chunk, double* const* d_peer_recv,
                       const int32_t* d_owner_of_ky, const int32_t* d_row_of_ky, int my_rank, int64_t nz_local, int64_t nyl,
                       int64_t z_offset, int max_ctas, void* stream) {
    FAVA_REQUIRE(ctx && d_data && d_peer_recv && d_owner_of_ky && d_row_of_ky, "fava_fft_y_scatter: NULL argument");
    FAVA_REQUIRE(fft_native_supported(n), "fava_fft_y_scatter: n = %lld is not a power of two in [256, 2048]", (long long)n);
    FAVA_REQUIRE(nz_chunk > 0 && z_offset >= 0 && z_offset + nz_chunk <= nz_local && nyl > 0 && my_rank >= 0,
                 "fava_fft_y_scatter: bad plane range");
    const int l = ilog2_pow2(n);
    DeviceGuard g(ctx->device);
    const double2 *t1, *t2;
    int rc = get_tables(ctx, l, &t1, &t2);
    if (rc) return rc;
    ColPrune pr;
    pr.mode = 1, pr.n = (int)n, pr.kmax2 = (int)(n * n / 4 - 3 * n / 2 + 2), pr.ky_of_batch = nullptr;
    ColScatter sc;
    sc.peer_recv = (double2* const*)d_peer_recv, sc.owner_of_ky = d_owner_of_ky, sc.row_of_ky = d_row_of_ky;
    sc.me = my_rank, sc.nz_local = (int)nz_local, sc.nyl = (int)nyl, sc.z_offset = (int)z_offset;
    cudaStream_t st = (cudaStream_t)stream;
    double2* data = (double2*)d_data;
    switch (l) {
        case 8: return launch_cols<8>(ctx, data, n / 2, n / 2, n, nz_chunk, 1, t1, t2, pr, st, &sc, max_ctas);
        case 9: return launch_cols<9>(ctx, data, n / 2, n / 2, n, nz_chunk, 1, t1, t2, pr, st, &sc, max_ctas);
        case 10: return launch_cols<10>(ctx, data, n / 2, n / 2, n, nz_chunk, 1, t1, t2, pr, st, &sc, max_ctas);
        default: return launch_cols<11>(ctx, data, n / 2, n / 2, n, nz_chunk, 1, t1, t2, pr, st, &sc, max_ctas);
    }
}

}  // extern "C"
