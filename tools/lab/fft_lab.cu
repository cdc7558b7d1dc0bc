// Design lab for the strided column transforms (NOT part of libfava_b200): candidate kernels around the
// register-resident FFT core (csrc/fft_core.cuh), each checked against cuFFT on a small batch and timed at full
// size, plus "copy-only" builds of the same kernels that isolate the cost of the access pattern.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -I fava_b200/csrc tools/lab/fft_lab.cu \
//        -lcufft -o tools/lab/fft_lab
//   tools/lab/fft_lab [nz_full=1024]
#include <cuda.h>
#include <cuda_runtime.h>
#include <cufft.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <type_traits>
#include <vector>

#include "fft_core.cuh"

using namespace fava::fftc;

#define CK(x)                                                                                     \
    do {                                                                                          \
        cudaError_t e_ = (x);                                                                     \
        if (e_ != cudaSuccess) {                                                                  \
            printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_)); \
            exit(2);                                                                              \
        }                                                                                         \
    } while (0)

// ---------------------------------------------------------------------------------------------------------
struct Prune {
    int mode;   // 0 none, 1 y pass (keep output rows ky^2 + kx0^2 <= kmax2), 2 z pass (tile skip + output rows)
    int kmax2;  // (N/2 - 1.5)^2 rounded down
    int n;
};

__device__ __forceinline__ int wavenumber(int k, int n) { return k < n / 2 ? k : k - n; }

template <int LOGN, int C, int MODE>
__global__ void __launch_bounds__(RegPlan<LOGN>::M1* C, (C >= 8 ? 1 : 2))
    k_cols_direct(double2* __restrict__ data, int64_t rstride, int64_t bstride, int ntile_cols, int64_t ntiles,
                  const double2* __restrict__ t1, const double2* __restrict__ t2, Prune pr) {
    using P = RegPlan<LOGN>;
    using O = Owner<LOGN>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* xb = reinterpret_cast<double2*>(smem_raw);
    const int c = threadIdx.x % C, u = threadIdx.x / C;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t b = t / ntile_cols;
        const int ct = (int)(t - b * ntile_cols);
        const int kx0 = ct * C;
        int base2 = kx0 * kx0;
        if (pr.mode == 2) {
            const int ky = wavenumber((int)b, pr.n);
            base2 += ky * ky;
            if (base2 > pr.kmax2) continue;
        }
        double2* base = data + b * bstride + kx0 + c;
        double2 v[16];
#pragma unroll
        for (int m = 0; m < 16; ++m) v[m] = __ldcs(base + (int64_t)O::in_index(u, m) * rstride);
        if (MODE == 0) {
            __syncthreads();  // previous tile's exchange reads are complete
            fft_regs_full<LOGN>(v, u, ColAddr<C>{c}, xb, t1, t2);
        }
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const int k = MODE == 0 ? O::out_freq(u, r) : O::in_index(u, r);
            bool keep = true;
            if (pr.mode) {
                const int w = wavenumber(k, pr.n);
                keep = w * w + base2 <= pr.kmax2;
            }
            if (keep) __stcs(base + (int64_t)k * rstride, v[r]);
        }
    }
}

// ---- TMA-fed persistent variant: next tile lands in shared memory while this one is transformed in registers ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    for (unsigned spin = 0; !mbar_try(bar, parity); ++spin)
        if (spin > (1u << 26)) __trap();  // lab safety: never hang the box
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}

// LINE_DIM = which tensor dimension the transform runs along (1: y pass, 2: z pass); C = 8, 16 doubles per row
template <int LOGN, int C, int MODE, int LINE_DIM>
__global__ void __launch_bounds__(RegPlan<LOGN>::M1 * C, 1)
    k_cols_tma(const __grid_constant__ CUtensorMap tmap, double2* __restrict__ data, int64_t rstride, int64_t bstride,
               int ntile_cols, int64_t ntiles, const double2* __restrict__ t1, const double2* __restrict__ t2, Prune pr) {
    using P = RegPlan<LOGN>;
    using O = Owner<LOGN>;
    constexpr int N = P::N;
    constexpr int BOX = 256;  // rows per TMA box
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double2* land = reinterpret_cast<double2*>(smem_raw);                          // [N][8] complex, 128 B rows
    double* xb = reinterpret_cast<double*>(smem_raw + sizeof(double2) * N * C);    // [N][8] doubles
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + sizeof(double2) * N * C + sizeof(double) * N * C);
    const int c = threadIdx.x % C, u = threadIdx.x / C;

    auto next_tile = [&](int64_t t) {  // first tile >= t that survives the tile-level pruning
        for (; t < ntiles; t += gridDim.x) {
            if (pr.mode != 2) break;
            const int64_t b = t / ntile_cols;
            const int kx0 = (int)(t - b * ntile_cols) * C;
            const int ky = wavenumber((int)b, pr.n);
            if (kx0 * kx0 + ky * ky <= pr.kmax2) break;
        }
        return t;
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

    int64_t t = next_tile(blockIdx.x);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async;\n" ::: "memory");
        if (t < ntiles) issue(t);
    }
    __syncthreads();
    unsigned parity = 0;
    while (t < ntiles) {
        const int64_t b = t / ntile_cols;
        const int kx0 = (int)(t - b * ntile_cols) * C;
        int base2 = kx0 * kx0;
        if (pr.mode == 2) {
            const int ky = wavenumber((int)b, pr.n);
            base2 += ky * ky;
        }
        mbar_wait(bar, parity);
        parity ^= 1u;
        double2 v[16];
#pragma unroll
        for (int m = 0; m < 16; ++m) v[m] = land[O::in_index(u, m) * C + c];
        __syncthreads();  // landing buffer consumed (and the previous tile's exchange reads are complete)
        const int64_t tn = next_tile(t + gridDim.x);
        if (threadIdx.x == 0 && tn < ntiles) issue(tn);
#ifdef FAVA_LAB_SHFL
        if (MODE == 0) fft_regs_half<LOGN, ColAddr<C>, CtaSync, C>(v, u, ColAddr<C>{c}, xb, t1, t2);
#else
        if (MODE == 0) fft_regs_half<LOGN>(v, u, ColAddr<C>{c}, xb, t1, t2);
#endif
        double2* base = data + b * bstride + kx0 + c;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const int k = MODE == 0 ? O::out_freq(u, r) : O::in_index(u, r);
            bool keep = true;
            if (pr.mode) {
                const int w = wavenumber(k, pr.n);
                keep = w * w + base2 <= pr.kmax2;
            }
            if (keep) __stcs(base + (int64_t)k * rstride, v[r]);
        }
        t = tn;
    }
}


// ---- x pass fused with the weighting: z = w[row a] + i w[row b], w = sqrt(rho) u_c; two-for-one real transform ----
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <typename T, int LOGN, int PAIRS>
struct XLayout {
    static constexpr int N = 1 << LOGN, M1 = N / 16, LINES = 3 * PAIRS, THREADS = LINES * M1, ROWS = 2 * PAIRS;
    static constexpr int LP = line_pitch(N);
    static constexpr size_t land_bytes = sizeof(T) * 4 * ROWS * N;
    static constexpr size_t srho_bytes = sizeof(double) * ROWS * N;
    static constexpr size_t xb_bytes = sizeof(double) * LINES * LP;
    static constexpr size_t total = land_bytes + srho_bytes + xb_bytes + 16;
};

template <typename T, int LOGN, int PAIRS, int CTAS, int MODE>
__global__ void __launch_bounds__(XLayout<T, LOGN, PAIRS>::THREADS, CTAS)
    k_xpass(const T* __restrict__ rho, const T* __restrict__ ux, const T* __restrict__ uy, const T* __restrict__ uz,
            int64_t ntiles, const double2* __restrict__ t1, const double2* __restrict__ t2, double2* __restrict__ fx,
            double2* __restrict__ fy, double2* __restrict__ fz, int64_t out_pitch) {
    using L = XLayout<T, LOGN, PAIRS>;
    using P = RegPlan<LOGN>;
    constexpr int N = L::N, M1 = L::M1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* land = reinterpret_cast<T*>(smem_raw);                                        // [4][ROWS][N]
    double* srho = reinterpret_cast<double*>(smem_raw + L::land_bytes);               // [ROWS][N]
    double* xb = reinterpret_cast<double*>(smem_raw + L::land_bytes + L::srho_bytes);  // [LINES][LP]
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + L::land_bytes + L::srho_bytes + L::xb_bytes);
    const int line = threadIdx.x / M1, u = threadIdx.x - line * M1;
    const int pair = line / 3, comp = line - 3 * pair;
    const T* src[4] = {rho, ux, uy, uz};
    double2* out = comp == 0 ? fx : (comp == 1 ? fy : fz);
    const LineAddr at{line * L::LP};

    auto issue = [&](int64_t t) {  // one thread: the 2 PAIRS rows of a tile are contiguous in every field
        constexpr unsigned bytes = (unsigned)(sizeof(T) * L::ROWS * N);
        mbar_expect_tx(bar, 4 * bytes);
#pragma unroll
        for (int f = 0; f < 4; ++f) bulk_load_1d(land + f * L::ROWS * N, src[f] + t * L::ROWS * N, bytes, bar);
    };
    int64_t t = blockIdx.x;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async;\n" ::: "memory");
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
        __syncthreads();  // landing rows and sqrt(rho) consumed
        if (threadIdx.x == 0 && t + gridDim.x < ntiles) issue(t + gridDim.x);
        double2 e[8], o[8];
        if (MODE == 0) {
            fft_regs_half<LOGN>(v, u, at, xb, t1, t2);
            split_two_for_one<LOGN>(v, u, at, xb, e, o);
        } else {
#pragma unroll
            for (int m = 0; m < 8; ++m) e[m] = v[m], o[m] = v[m + 8];
        }
        const int64_t row_a = (t * PAIRS + pair) * 2;
        double2* oa = out + row_a * out_pitch;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int k = u + M1 * m;
            __stcs(oa + k, e[m]);
            __stcs(oa + out_pitch + k, o[m]);
        }
    }
}


// ---- x pass, one ROW per warp-line: real row of N points = complex transform of N/2 points of the even/odd packing ----
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

template <typename T, int LOGN, int CTAS, int MODE>
__global__ void __launch_bounds__(XRowLayout<T, LOGN>::THREADS, CTAS)
    k_xrow(const T* __restrict__ rho, const T* __restrict__ ux, const T* __restrict__ uy, const T* __restrict__ uz, int64_t nrows,
           const double2* __restrict__ t1, const double2* __restrict__ t2, const double2* __restrict__ tw,
           double2* __restrict__ fx, double2* __restrict__ fy, double2* __restrict__ fz, int64_t out_pitch) {
    using L = XRowLayout<T, LOGN>;
    constexpr int N = L::N, H = L::H, M1 = L::M1, LOGH = L::LOGH;
    typedef typename Vec2<T>::type T2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* land = reinterpret_cast<T*>(smem_raw);                                        // [4][N]
    double* srho = reinterpret_cast<double*>(smem_raw + L::land_bytes);               // [N]
    double* xb = reinterpret_cast<double*>(smem_raw + L::land_bytes + L::srho_bytes);  // [3][LP]
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + L::land_bytes + L::srho_bytes + L::xb_bytes);
    const int comp = threadIdx.x / M1, u = threadIdx.x - comp * M1;
    const T* src[4] = {rho, ux, uy, uz};
    double2* out = comp == 0 ? fx : (comp == 1 ? fy : fz);
    const LineAddr at{comp * L::LP};
    const LineSync<M1> line_sync{comp};

    auto issue = [&](int64_t r) {
        constexpr unsigned bytes = (unsigned)(sizeof(T) * N);
        mbar_expect_tx(bar, 4 * bytes);
#pragma unroll
        for (int f = 0; f < 4; ++f) bulk_load_1d(land + f * N, src[f] + r * N, bytes, bar);
    };
    int64_t row = blockIdx.x;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async;\n" ::: "memory");
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
        __syncthreads();
        if (threadIdx.x == 0 && row + gridDim.x < nrows) issue(row + gridDim.x);
        double2* orow = out + row * out_pitch;
        if (MODE == 0) {
            double2 e[8], o[8], mid;
#ifdef FAVA_LAB_SHFL
            fft_regs_half<LOGH, LineAddr, LineSync<M1>, 1>(v, u, at, xb, t1, t2, line_sync);
#else
            fft_regs_half<LOGH>(v, u, at, xb, t1, t2, line_sync);
#endif
            split_two_for_one<LOGH>(v, u, at, xb, e, o, line_sync, &mid);
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const int k = u + M1 * m;  // k < H/2
                const double2 t = cmul(o[m], __ldg(tw + k));
                __stcs(orow + k, cadd(e[m], t));                                  // X[k] = E + w^k O
                const double2 d = csub(e[m], t);
                if (k != 0) __stcs(orow + (H - k), make_double2(d.x, -d.y));       // X[H-k] = conj(E - w^k O)
            }
            if (u == 0) __stcs(orow + H / 2, make_double2(mid.x, -mid.y));         // X[H/2] = conj Z[H/2]
        } else {
#pragma unroll
            for (int m = 0; m < 16; ++m) __stcs(orow + u + M1 * m, v[m]);
        }
    }
}

template <typename T>
__global__ void k_fill_fields(T* rho, T* ux, T* uy, T* uz, int64_t n, uint64_t seed) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        rho[i] = (T)(1.0 + 0.5 * (double)(z & 0xffff) / 65536.0);
        ux[i] = (T)((double)((z >> 16) & 0xffff) / 65536.0 - 0.5);
        uy[i] = (T)((double)((z >> 32) & 0xffff) / 65536.0 - 0.5);
        uz[i] = (T)((double)((z >> 48) & 0xffff) / 65536.0 - 0.5);
    }
}
template <typename T>
__global__ void k_weight_ref(const T* rho, const T* u, double* w, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        w[i] = sqrt((double)rho[i]) * (double)u[i];
}

// ---------------------------------------------------------------------------------------------------------
__global__ void k_fill(double2* d, int64_t n, uint64_t seed) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        d[i] = make_double2((double)(z & 0xffffffffu) / 4294967296.0 - 0.5, (double)(z >> 32) / 4294967296.0 - 0.5);
    }
}

template <int LOGN>
static void make_tables(double2** d_t1, double2** d_t2) {
    using P = RegPlan<LOGN>;
    std::vector<double2> t1(P::T1_LEN), t2(P::T2_LEN > 0 ? P::T2_LEN : 1);
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int q = 0; q < 16; ++q)
        for (int j = 0; j < P::M1; ++j) {
            const long double a = -two_pi * (long double)((j * q) % P::N) / (long double)P::N;
            t1[q * P::M1 + j] = make_double2((double)cosl(a), (double)sinl(a));
        }
    for (int q = 0; q < 16; ++q)
        for (int j = 0; j < P::M2; ++j) {
            const long double a = -two_pi * (long double)((j * q) % P::M1) / (long double)P::M1;
            t2[q * P::M2 + j] = make_double2((double)cosl(a), (double)sinl(a));
        }
    CK(cudaMalloc(d_t1, sizeof(double2) * t1.size()));
    CK(cudaMalloc(d_t2, sizeof(double2) * t2.size()));
    CK(cudaMemcpy(*d_t1, t1.data(), sizeof(double2) * t1.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(*d_t2, t2.data(), sizeof(double2) * t2.size(), cudaMemcpyHostToDevice));
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiled get_encode() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) {
        printf("cuTensorMapEncodeTiled not available\n");
        exit(2);
    }
    return (EncodeTiled)fn;
}

// data: complex [d2][d1][pitch]; box = C complex x min(256, n) along line_dim
static CUtensorMap make_map(double2* data, int64_t pitch, int64_t d1, int64_t d2, int line_dim, int C, int n) {
    static EncodeTiled enc = get_encode();
    CUtensorMap m;
    const cuuint32_t rows = (cuuint32_t)std::min(256, n);
    cuuint64_t dims[3] = {(cuuint64_t)(2 * pitch), (cuuint64_t)d1, (cuuint64_t)d2};
    cuuint64_t strides[2] = {(cuuint64_t)(pitch * 16), (cuuint64_t)(pitch * 16 * d1)};
    cuuint32_t box[3] = {(cuuint32_t)(2 * C), line_dim == 1 ? rows : 1u, line_dim == 2 ? rows : 1u};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, data, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
        exit(2);
    }
    return m;
}

struct Shape {
    int64_t pitch, d1, d2;  // complex [d2][d1][pitch]
    int line_dim;           // 1: transform along d1 (y pass), 2: along d2 (z pass)
    int64_t rstride() const { return line_dim == 1 ? pitch : pitch * d1; }
    int64_t bstride() const { return line_dim == 1 ? pitch * d1 : pitch; }
    int64_t nbatch() const { return line_dim == 1 ? d2 : d1; }
    int64_t elems() const { return pitch * d1 * d2; }
};

static int g_sms = 148;

template <int LOGN, int C, int MODE>
static void run_direct(double2* d, const Shape& s, const double2* t1, const double2* t2, Prune pr, int ctas_per_sm) {
    constexpr int N = 1 << LOGN;
    const int ntc = (N / 2) / C;
    const int64_t ntiles = s.nbatch() * ntc;
    const size_t smem = sizeof(double2) * N * C;
    auto kern = k_cols_direct<LOGN, C, MODE>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::min<int64_t>(ntiles, (int64_t)g_sms * ctas_per_sm);
    kern<<<grid, RegPlan<LOGN>::M1 * C, smem>>>(d, s.rstride(), s.bstride(), ntc, ntiles, t1, t2, pr);
    CK(cudaGetLastError());
}

template <int LOGN, int MODE>
static void run_tma(double2* d, const Shape& s, const double2* t1, const double2* t2, Prune pr) {
    constexpr int N = 1 << LOGN;
    constexpr int C = 8192 / N;  // 512 threads x 16 points
    const int ntc = (N / 2) / C;
    const int64_t ntiles = s.nbatch() * ntc;
    const size_t smem = sizeof(double2) * N * C + sizeof(double) * N * C + 64;
    CUtensorMap map = make_map(d, s.pitch, s.d1, s.d2, s.line_dim, C, N);
    const int grid = (int)std::min<int64_t>(ntiles, (int64_t)g_sms);
    if (s.line_dim == 1) {
        auto kern = k_cols_tma<LOGN, C, MODE, 1>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, RegPlan<LOGN>::M1 * C, smem>>>(map, d, s.rstride(), s.bstride(), ntc, ntiles, t1, t2, pr);
    } else {
        auto kern = k_cols_tma<LOGN, C, MODE, 2>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, RegPlan<LOGN>::M1 * C, smem>>>(map, d, s.rstride(), s.bstride(), ntc, ntiles, t1, t2, pr);
    }
    CK(cudaGetLastError());
}

static void cufft_ref(double2* d, const Shape& s, int N) {  // in place, all columns
    cufftHandle h;
    int n[1] = {N};
    int embed[1] = {N};
    if (s.line_dim == 1) {
        if (cufftPlanMany(&h, 1, n, embed, (int)s.pitch, 1, embed, (int)s.pitch, 1, CUFFT_Z2Z, (int)s.pitch) != CUFFT_SUCCESS) exit(3);
        for (int64_t z = 0; z < s.d2; ++z) {
            cufftDoubleComplex* p = (cufftDoubleComplex*)(d + z * s.pitch * s.d1);
            if (cufftExecZ2Z(h, p, p, CUFFT_FORWARD) != CUFFT_SUCCESS) exit(3);
        }
    } else {
        const int cols = (int)(s.pitch * s.d1);
        if (cufftPlanMany(&h, 1, n, embed, cols, 1, embed, cols, 1, CUFFT_Z2Z, cols) != CUFFT_SUCCESS) exit(3);
        if (cufftExecZ2Z(h, (cufftDoubleComplex*)d, (cufftDoubleComplex*)d, CUFFT_FORWARD) != CUFFT_SUCCESS) exit(3);
    }
    CK(cudaDeviceSynchronize());
    cufftDestroy(h);
}

// max |a - b| over the elements a pruned transform must produce, relative to max |b|
static double compare(const std::vector<double2>& a, const std::vector<double2>& b, const Shape& s, Prune pr, int C, int N) {
    double err = 0, ref = 0;
    const int kmax2 = pr.kmax2;
    for (int64_t i2 = 0; i2 < s.d2; ++i2)
        for (int64_t i1 = 0; i1 < s.d1; ++i1)
            for (int64_t x = 0; x < N / 2; ++x) {
                const int64_t line = s.line_dim == 1 ? i1 : i2, batch = s.line_dim == 1 ? i2 : i1;
                if (pr.mode) {
                    const int kx0 = (int)(x / C) * C;
                    int w = (int)(line < N / 2 ? line : line - N);
                    int k2 = w * w + kx0 * kx0;
                    if (pr.mode == 2) {
                        const int ky = (int)(batch < N / 2 ? batch : batch - N);
                        k2 += ky * ky;
                    }
                    if (k2 > kmax2) continue;
                }
                const int64_t idx = (i2 * s.d1 + i1) * s.pitch + x;
                err = std::max(err, std::max(fabs(a[idx].x - b[idx].x), fabs(a[idx].y - b[idx].y)));
                ref = std::max(ref, std::max(fabs(b[idx].x), fabs(b[idx].y)));
            }
    return err / ref;
}

template <typename F>
static float time_ms(F&& f, int reps = 5) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(e0));
        f();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, ms);
    }
    return best;
}

template <int LOGN>
static void check_cols() {
    constexpr int N = 1 << LOGN;
    constexpr int CT = 8192 / N;
    double2 *t1, *t2;
    make_tables<LOGN>(&t1, &t2);
    const int kmax2 = N * N / 4 - 3 * N / 2 + 2;
    const int pitch = N / 2;
    for (int line_dim : {1, 2}) {
        Shape s{pitch, line_dim == 1 ? (int64_t)N : 6, line_dim == 1 ? 3 : (int64_t)N, line_dim};
        const int64_t ne = s.elems();
        double2 *d_in, *d_ref, *d_out;
        CK(cudaMalloc(&d_in, sizeof(double2) * ne));
        CK(cudaMalloc(&d_ref, sizeof(double2) * ne));
        CK(cudaMalloc(&d_out, sizeof(double2) * ne));
        k_fill<<<1024, 256>>>(d_in, ne, 42);
        CK(cudaMemcpy(d_ref, d_in, sizeof(double2) * ne, cudaMemcpyDeviceToDevice));
        cufft_ref(d_ref, s, N);
        std::vector<double2> h_ref(ne), h_out(ne);
        CK(cudaMemcpy(h_ref.data(), d_ref, sizeof(double2) * ne, cudaMemcpyDeviceToHost));
        for (int prm : {0, line_dim}) {
            Prune pr{prm, kmax2, N};
            CK(cudaMemcpy(d_out, d_in, sizeof(double2) * ne, cudaMemcpyDeviceToDevice));
            run_tma<LOGN, 0>(d_out, s, t1, t2, pr);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
                printf("CHECK cols N %d dim %d prune %d: CUDA error %s\n", N, line_dim, prm, cudaGetErrorString(e));
                exit(4);
            }
            CK(cudaMemcpy(h_out.data(), d_out, sizeof(double2) * ne, cudaMemcpyDeviceToHost));
            const double err = compare(h_out, h_ref, s, pr, CT, N);
            printf("CHECK cols tma N %d C %d dim %d prune %d: rel err %.3e %s\n", N, CT, line_dim, prm, err, err < 1e-13 ? "ok" : "FAIL");
        }
        CK(cudaFree(d_in));
        CK(cudaFree(d_ref));
        CK(cudaFree(d_out));
    }
    CK(cudaFree(t1));
    CK(cudaFree(t2));
}

template <typename T, int LOGN, int PAIRS, int CTAS, int MODE>
static void run_x(const T* rho, const T* ux, const T* uy, const T* uz, int64_t nrows, const double2* t1, const double2* t2,
                  double2* fx, double2* fy, double2* fz, int64_t pitch) {
    using L = XLayout<T, LOGN, PAIRS>;
    auto kern = k_xpass<T, LOGN, PAIRS, CTAS, MODE>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::total));
    const int64_t ntiles = nrows / L::ROWS;
    const int grid = (int)std::min<int64_t>(ntiles, (int64_t)g_sms * CTAS);
    kern<<<grid, L::THREADS, L::total>>>(rho, ux, uy, uz, ntiles, t1, t2, fx, fy, fz, pitch);
    CK(cudaGetLastError());
}

template <typename T, int LOGN, int PAIRS, int CTAS>
static void check_x(int64_t nrows_full) {
    constexpr int N = 1 << LOGN;
    using L = XLayout<T, LOGN, PAIRS>;
    double2 *t1, *t2;
    make_tables<LOGN>(&t1, &t2);
    const int64_t pitch = N / 2;
    {
        const int64_t nrows = 96;  // divisible by 2 PAIRS for PAIRS in {1,2,4}... and 8
        const int64_t ne = nrows * N;
        T *rho, *ux, *uy, *uz;
        CK(cudaMalloc(&rho, sizeof(T) * ne));
        CK(cudaMalloc(&ux, sizeof(T) * ne));
        CK(cudaMalloc(&uy, sizeof(T) * ne));
        CK(cudaMalloc(&uz, sizeof(T) * ne));
        k_fill_fields<T><<<256, 256>>>(rho, ux, uy, uz, ne, 99);
        double2* f[3];
        for (auto& p : f) CK(cudaMalloc(&p, sizeof(double2) * nrows * pitch));
        run_x<T, LOGN, PAIRS, CTAS, 0>(rho, ux, uy, uz, nrows, t1, t2, f[0], f[1], f[2], pitch);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("CHECK x N %d: CUDA error %s\n", N, cudaGetErrorString(e));
            exit(4);
        }
        double* w;
        double2* ref;
        CK(cudaMalloc(&w, sizeof(double) * ne));
        CK(cudaMalloc(&ref, sizeof(double2) * nrows * (N / 2 + 1)));
        cufftHandle h;
        int n[1] = {N};
        if (cufftPlanMany(&h, 1, n, nullptr, 1, N, nullptr, 1, N / 2 + 1, CUFFT_D2Z, (int)nrows) != CUFFT_SUCCESS) exit(3);
        const T* u[3] = {ux, uy, uz};
        std::vector<double2> h_ref(nrows * (N / 2 + 1)), h_out(nrows * pitch);
        for (int c = 0; c < 3; ++c) {
            k_weight_ref<T><<<256, 256>>>(rho, u[c], w, ne);
            if (cufftExecD2Z(h, w, (cufftDoubleComplex*)ref) != CUFFT_SUCCESS) exit(3);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h_ref.data(), ref, sizeof(double2) * h_ref.size(), cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(h_out.data(), f[c], sizeof(double2) * h_out.size(), cudaMemcpyDeviceToHost));
            double err = 0, mx = 0;
            for (int64_t r = 0; r < nrows; ++r)
                for (int k = 0; k < N / 2; ++k) {
                    const double2 a = h_out[r * pitch + k], b = h_ref[r * (N / 2 + 1) + k];
                    err = std::max(err, std::max(fabs(a.x - b.x), fabs(a.y - b.y)));
                    mx = std::max(mx, std::max(fabs(b.x), fabs(b.y)));
                }
            printf("CHECK x %s N %d pairs %d ctas %d smem %zu comp %d: rel err %.3e %s\n", sizeof(T) == 8 ? "f64" : "f32", N, PAIRS,
                   CTAS, (size_t)L::total, c, err / mx, err / mx < 1e-13 ? "ok" : "FAIL");
        }
        cufftDestroy(h);
        for (auto p : {(void*)rho, (void*)ux, (void*)uy, (void*)uz, (void*)w, (void*)ref, (void*)f[0], (void*)f[1], (void*)f[2]}) CK(cudaFree(p));
    }
    if (nrows_full > 0) {
        const int64_t ne = nrows_full * N;
        T *rho, *ux, *uy, *uz;
        CK(cudaMalloc(&rho, sizeof(T) * ne));
        CK(cudaMalloc(&ux, sizeof(T) * ne));
        CK(cudaMalloc(&uy, sizeof(T) * ne));
        CK(cudaMalloc(&uz, sizeof(T) * ne));
        k_fill_fields<T><<<4096, 256>>>(rho, ux, uy, uz, ne, 5);
        double2* f[3];
        for (auto& p : f) CK(cudaMalloc(&p, sizeof(double2) * nrows_full * pitch));
        CK(cudaDeviceSynchronize());
        const double gb = (4.0 * sizeof(T) + 24.0) * (double)ne / 1e9;
        float ms = time_ms([&] { run_x<T, LOGN, PAIRS, CTAS, 1>(rho, ux, uy, uz, nrows_full, t1, t2, f[0], f[1], f[2], pitch); });
        printf("TIME x %s N %d pairs %d ctas %d copy-only %8.3f ms %7.1f GB/s\n", sizeof(T) == 8 ? "f64" : "f32", N, PAIRS, CTAS, ms,
               gb / (ms * 1e-3));
        ms = time_ms([&] { run_x<T, LOGN, PAIRS, CTAS, 0>(rho, ux, uy, uz, nrows_full, t1, t2, f[0], f[1], f[2], pitch); });
        printf("TIME x %s N %d pairs %d ctas %d fft       %8.3f ms %7.1f GB/s (algorithmic %0.1f GB)\n", sizeof(T) == 8 ? "f64" : "f32", N,
               PAIRS, CTAS, ms, gb / (ms * 1e-3), gb);
        fflush(stdout);
        for (auto p : {(void*)rho, (void*)ux, (void*)uy, (void*)uz, (void*)f[0], (void*)f[1], (void*)f[2]}) CK(cudaFree(p));
    }
    CK(cudaFree(t1));
    CK(cudaFree(t2));
}


template <typename T, int LOGN, int CTAS>
static void check_xrow(int64_t nrows_full) {
    constexpr int N = 1 << LOGN, H = N / 2;
    using L = XRowLayout<T, LOGN>;
    double2 *t1, *t2;
    make_tables<LOGN - 1>(&t1, &t2);
    std::vector<double2> htw(H / 2);
    for (int k = 0; k < H / 2; ++k) {
        const long double a = -6.283185307179586476925286766559005768L * k / N;
        htw[k] = make_double2((double)cosl(a), (double)sinl(a));
    }
    double2* tw;
    CK(cudaMalloc(&tw, sizeof(double2) * htw.size()));
    CK(cudaMemcpy(tw, htw.data(), sizeof(double2) * htw.size(), cudaMemcpyHostToDevice));
    const int64_t pitch = H;
    auto run = [&](auto mode, const T* rho, const T* ux, const T* uy, const T* uz, int64_t nrows, double2** f) {
        constexpr int MODE = decltype(mode)::value;
        auto kern = k_xrow<T, LOGN, CTAS, MODE>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::total));
        const int grid = (int)std::min<int64_t>(nrows, (int64_t)g_sms * CTAS);
        kern<<<grid, L::THREADS, L::total>>>(rho, ux, uy, uz, nrows, t1, t2, tw, f[0], f[1], f[2], pitch);
        CK(cudaGetLastError());
    };
    for (int64_t nrows : {(int64_t)97, nrows_full}) {
        if (nrows <= 0) continue;
        const int64_t ne = nrows * N;
        T *rho, *ux, *uy, *uz;
        for (auto pp : {&rho, &ux, &uy, &uz}) CK(cudaMalloc(pp, sizeof(T) * ne));
        k_fill_fields<T><<<4096, 256>>>(rho, ux, uy, uz, ne, 99);
        double2* f[3];
        for (auto& p : f) CK(cudaMalloc(&p, sizeof(double2) * nrows * pitch));
        if (nrows == 97) {
            run(std::integral_constant<int, 0>(), rho, ux, uy, uz, nrows, f);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CHECK xrow N %d: CUDA error %s\n", N, cudaGetErrorString(e)); exit(4); }
            double* w;
            double2* ref;
            CK(cudaMalloc(&w, sizeof(double) * ne));
            CK(cudaMalloc(&ref, sizeof(double2) * nrows * (N / 2 + 1)));
            cufftHandle h;
            int n[1] = {N};
            if (cufftPlanMany(&h, 1, n, nullptr, 1, N, nullptr, 1, N / 2 + 1, CUFFT_D2Z, (int)nrows) != CUFFT_SUCCESS) exit(3);
            const T* uu[3] = {ux, uy, uz};
            std::vector<double2> h_ref(nrows * (N / 2 + 1)), h_out(nrows * pitch);
            for (int c = 0; c < 3; ++c) {
                k_weight_ref<T><<<256, 256>>>(rho, uu[c], w, ne);
                if (cufftExecD2Z(h, w, (cufftDoubleComplex*)ref) != CUFFT_SUCCESS) exit(3);
                CK(cudaDeviceSynchronize());
                CK(cudaMemcpy(h_ref.data(), ref, sizeof(double2) * h_ref.size(), cudaMemcpyDeviceToHost));
                CK(cudaMemcpy(h_out.data(), f[c], sizeof(double2) * h_out.size(), cudaMemcpyDeviceToHost));
                double err = 0, mx = 0;
                for (int64_t r = 0; r < nrows; ++r)
                    for (int k = 0; k < N / 2; ++k) {
                        const double2 a = h_out[r * pitch + k], b = h_ref[r * (N / 2 + 1) + k];
                        err = std::max(err, std::max(fabs(a.x - b.x), fabs(a.y - b.y)));
                        mx = std::max(mx, std::max(fabs(b.x), fabs(b.y)));
                    }
                printf("CHECK xrow %s N %d ctas %d smem %zu comp %d: rel err %.3e %s\n", sizeof(T) == 8 ? "f64" : "f32", N, CTAS,
                       (size_t)L::total, c, err / mx, err / mx < 1e-13 ? "ok" : "FAIL");
            }
            cufftDestroy(h);
            CK(cudaFree(w));
            CK(cudaFree(ref));
        } else {
            CK(cudaDeviceSynchronize());
            const double gb = (4.0 * sizeof(T) + 24.0) * (double)ne / 1e9;
            float ms = time_ms([&] { run(std::integral_constant<int, 1>(), rho, ux, uy, uz, nrows, f); });
            printf("TIME xrow %s N %d ctas %d copy-only %8.3f ms %7.1f GB/s\n", sizeof(T) == 8 ? "f64" : "f32", N, CTAS, ms, gb / (ms * 1e-3));
            ms = time_ms([&] { run(std::integral_constant<int, 0>(), rho, ux, uy, uz, nrows, f); });
            printf("TIME xrow %s N %d ctas %d fft       %8.3f ms %7.1f GB/s (algorithmic %0.1f GB)\n", sizeof(T) == 8 ? "f64" : "f32", N, CTAS,
                   ms, gb / (ms * 1e-3), gb);
            fflush(stdout);
        }
        for (auto p : {(void*)rho, (void*)ux, (void*)uy, (void*)uz, (void*)f[0], (void*)f[1], (void*)f[2]}) CK(cudaFree(p));
    }
    CK(cudaFree(t1));
    CK(cudaFree(t2));
    CK(cudaFree(tw));
}

static void time_cols(int64_t nfull, bool all) {
    constexpr int LOGN = 10, N = 1024;
    double2 *t1, *t2;
    make_tables<LOGN>(&t1, &t2);
    const int kmax2 = N * N / 4 - 3 * N / 2 + 2;
    for (int pitch : {512}) {
        double2* d;
        const int64_t ne = (int64_t)pitch * N * nfull;
        CK(cudaMalloc(&d, sizeof(double2) * ne));
        k_fill<<<4096, 256>>>(d, ne, 7);
        CK(cudaDeviceSynchronize());
        const double gb_full = 2.0 * 16.0 * (double)(N / 2) * N * nfull / 1e9;
        for (int line_dim : {1, 2}) {
            Shape s{pitch, line_dim == 1 ? (int64_t)N : nfull, line_dim == 1 ? nfull : (int64_t)N, line_dim};
            for (int prm : {0, line_dim}) {
                Prune pr{prm, kmax2, N};
                auto report = [&](const char* name, float ms) {
                    printf("TIME pitch %d dim %d prune %d %-16s %8.3f ms  %7.1f GB/s (of the unpruned %0.1f GB)\n", pitch, line_dim,
                           prm, name, ms, gb_full / (ms * 1e-3), gb_full);
                    fflush(stdout);
                };
                if (all) {
                    report("direct C8 copy", time_ms([&] { run_direct<LOGN, 8, 1>(d, s, t1, t2, pr, 1); }));
                    report("direct C8 fft", time_ms([&] { run_direct<LOGN, 8, 0>(d, s, t1, t2, pr, 1); }));
                    report("tma C8 copy", time_ms([&] { run_tma<LOGN, 1>(d, s, t1, t2, pr); }));
                }
                report("tma C8 fft", time_ms([&] { run_tma<LOGN, 0>(d, s, t1, t2, pr); }));
            }
        }
        CK(cudaFree(d));
    }
}

int main(int argc, char** argv) {
    const std::string what = argc > 1 ? argv[1] : "all";
    const int64_t nfull = argc > 2 ? atoll(argv[2]) : 1024;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    g_sms = prop.multiProcessorCount;
    printf("device %s, %d SMs, mode %s\n", prop.name, g_sms, what.c_str());
    if (what == "prof") {  // for ncu: exactly one pruned y pass, one pruned z pass and one fused x pass at full size
        constexpr int N = 1024;
        double2 *t1, *t2;
        make_tables<10>(&t1, &t2);
        const int kmax2 = N * N / 4 - 3 * N / 2 + 2;
        const int64_t ne = (int64_t)512 * N * nfull;
        double2* f[3];
        for (auto& p : f) CK(cudaMalloc(&p, sizeof(double2) * ne));
        double *rho, *ux, *uy, *uz;
        for (auto pp : {&rho, &ux, &uy, &uz}) CK(cudaMalloc(pp, sizeof(double) * N * N * nfull));
        k_fill_fields<double><<<4096, 256>>>(rho, ux, uy, uz, (int64_t)N * N * nfull, 5);
        run_x<double, 10, 1, 2, 0>(rho, ux, uy, uz, N * nfull, t1, t2, f[0], f[1], f[2], 512);
        Shape sy{512, N, nfull, 1}, sz{512, nfull, N, 2};
        run_tma<10, 0>(f[0], sy, t1, t2, Prune{1, kmax2, N});
        run_tma<10, 0>(f[0], sz, t1, t2, Prune{2, kmax2, N});
        CK(cudaDeviceSynchronize());
        printf("PROF DONE\n");
        return 0;
    }
    if (what == "xrow") {
        check_xrow<double, 9, 4>(0);
        check_xrow<float, 9, 4>(0);
        check_xrow<double, 11, 2>(0);
        check_xrow<float, 11, 2>(0);
        check_xrow<float, 10, 4>(1024 * nfull);
        check_xrow<double, 10, 4>(1024 * nfull);
        check_xrow<double, 10, 3>(1024 * nfull);
        check_xrow<double, 10, 2>(1024 * nfull);
        check_x<double, 10, 1, 2>(1024 * nfull);
        printf("LAB DONE\n");
        return 0;
    }
    if (what == "all" || what == "cols") {
        check_cols<8>();
        check_cols<9>();
        check_cols<10>();
        check_cols<11>();
        time_cols(nfull, true);
    }
    if (what == "all" || what == "x") {
        check_x<double, 8, 4, 2>(0);
        check_x<float, 8, 4, 2>(0);
        check_x<double, 9, 2, 2>(0);
        check_x<float, 9, 2, 2>(0);
        check_x<double, 11, 1, 1>(0);
        check_x<float, 11, 1, 1>(0);
        check_x<float, 10, 1, 2>(1024 * nfull);
        check_x<double, 10, 1, 2>(1024 * nfull);
        check_x<double, 10, 2, 1>(1024 * nfull);
        check_x<double, 10, 1, 1>(1024 * nfull);
    }
    printf("LAB DONE\n");
    return 0;
}
