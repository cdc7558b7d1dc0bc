// Register-resident fp64 line FFT core shared by the transform kernels (csrc/fft.cu) and the design lab
// (tools/lab/fft_lab.cu).
//
// A line of N = 2^LOGN complex points (256 <= N <= 4096) is transformed by N/16 threads holding 16 points each:
//   pass 1  radix 16 over the elements  n = j + (N/16) m,              thread j           -> sub-line q (= m after DFT)
//   pass 2  radix 16 over the elements  j = j2 + M2 m2 of sub-line q,  thread (q, j2)     -> (q, q2)
//   pass 3  radix M2 = N/256 over j2 of (q, q2),                       thread (q, h) owns q2 = h (16/M2) + i
// and the output frequency of element (q, q2, k3) is  k = q + 16 q2 + 256 k3  (decimation in frequency, natural
// order in, natural order out - the digit reversal is absorbed by where each thread reads and writes).  Between
// passes the points change owner through shared memory: two exchanges in all, each point written once and read once.
// Shared memory is touched for nothing else, so a 1024-point line costs 2 x 32 B of shared-memory traffic per
// point against 6 x 32 B for a transform that keeps the line in shared memory between radix-16 passes.
//
// `C` lines are interleaved (line index fastest in the thread index and in the exchange buffer), so that the C
// threads working on the same element index of adjacent lines touch C consecutive 16-byte words: with C = 8 every
// quarter-warp access is one 128-byte row and there are no bank conflicts for any access pattern.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace fava {
namespace fftc {

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
__device__ __forceinline__ void dft8(double2& v0, double2& v1, double2& v2, double2& v3, double2& v4, double2& v5,
                                     double2& v6, double2& v7) {
    // m = 2a + b: 4-point DFTs over a for even and odd m, then X[c + 4d] = y0[c] + (-1)^d w8^c y1[c]
    dft4(v0, v2, v4, v6);
    dft4(v1, v3, v5, v7);
    const double h = 0.70710678118654752440;
    const double2 y0[4] = {v0, v2, v4, v6};
    double2 y1[4] = {v1, v3, v5, v7};
    y1[1] = make_double2(h * (y1[1].x + y1[1].y), h * (y1[1].y - y1[1].x));   // * (1 - i)/sqrt2
    y1[2] = mul_mi(y1[2]);                                                    // * (-i)
    y1[3] = make_double2(h * (y1[3].y - y1[3].x), -h * (y1[3].x + y1[3].y));  // * (-1 - i)/sqrt2
    v0 = cadd(y0[0], y1[0]), v4 = csub(y0[0], y1[0]);
    v1 = cadd(y0[1], y1[1]), v5 = csub(y0[1], y1[1]);
    v2 = cadd(y0[2], y1[2]), v6 = csub(y0[2], y1[2]);
    v3 = cadd(y0[3], y1[3]), v7 = csub(y0[3], y1[3]);
}
__device__ __forceinline__ void dft16(double2 (&v)[16]) {
    // m = 4a + b, q = c + 4d : y[b][c] = DFT4 over a of v[4a+b];  y[b][c] *= w16^(b c);  X[c+4d] = DFT4 over b
    double2 y[4][4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        y[b][0] = v[b], y[b][1] = v[4 + b], y[b][2] = v[8 + b], y[b][3] = v[12 + b];
        dft4(y[b][0], y[b][1], y[b][2], y[b][3]);
    }
    const double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173, h = 0.70710678118654752440;
    const double2 w1 = make_double2(c1, -s1), w3 = make_double2(s1, -c1), w9 = make_double2(-c1, s1);
    y[1][1] = cmul(y[1][1], w1);
    y[1][2] = make_double2(h * (y[1][2].x + y[1][2].y), h * (y[1][2].y - y[1][2].x));   // w16^2 = (1 - i)/sqrt2
    y[1][3] = cmul(y[1][3], w3);
    y[2][1] = make_double2(h * (y[2][1].x + y[2][1].y), h * (y[2][1].y - y[2][1].x));   // w16^2
    y[2][2] = mul_mi(y[2][2]);                                                           // w16^4 = -i
    y[2][3] = make_double2(h * (y[2][3].y - y[2][3].x), -h * (y[2][3].x + y[2][3].y));  // w16^6 = (-1 - i)/sqrt2
    y[3][1] = cmul(y[3][1], w3);
    y[3][2] = make_double2(h * (y[3][2].y - y[3][2].x), -h * (y[3][2].x + y[3][2].y));  // w16^6
    y[3][3] = cmul(y[3][3], w9);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        dft4(y[0][c], y[1][c], y[2][c], y[3][c]);
#pragma unroll
        for (int d = 0; d < 4; ++d) v[c + 4 * d] = y[d][c];
    }
}

// DFTs of length R over consecutive groups of R registers of v[16] (pass 3)
template <int R>
__device__ __forceinline__ void dft_groups(double2 (&v)[16]) {
    if constexpr (R == 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dft2(v[2 * i], v[2 * i + 1]);
    } else if constexpr (R == 4) {
#pragma unroll
        for (int i = 0; i < 4; ++i) dft4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else if constexpr (R == 8) {
#pragma unroll
        for (int i = 0; i < 2; ++i)
            dft8(v[8 * i], v[8 * i + 1], v[8 * i + 2], v[8 * i + 3], v[8 * i + 4], v[8 * i + 5], v[8 * i + 6], v[8 * i + 7]);
    } else if constexpr (R == 16) {
        dft16(v);
    }
}

template <int LOGN>
struct RegPlan {
    static_assert(LOGN >= 8 && LOGN <= 12, "register-resident FFT: 256 <= N <= 4096");
    static constexpr int N = 1 << LOGN;
    static constexpr int M1 = N / 16;   // threads per line = length of the sub-lines after pass 1
    static constexpr int M2 = M1 / 16;  // radix of pass 3 (1 = no third pass)
    static constexpr int G = 16 / M2;   // pass-3 butterflies per thread
    // twiddle tables (doubles2, built on the host): T1[q][j] = exp(-2 pi i j q / N), q < 16, j < M1;
    //                                               T2[q2][j2] = exp(-2 pi i j2 q2 / M1), q2 < 16, j2 < M2
    static constexpr int T1_LEN = 16 * M1, T2_LEN = 16 * M2;
};

// exchange-buffer position of element index e of a line: bit 4 folded into bit 0, so that threads whose
// element indices differ by 1 OR by 16 land in words of different parity (needed when 2 element indices share
// one shared-memory access phase: C = 4 with 16-byte words, C = 8 with 8-byte words)
__device__ __forceinline__ int phi(int e) { return e ^ ((e >> 4) & 1); }

// Element indices (within a line) that thread u of a line reads / writes in each phase.
template <int LOGN>
struct Owner {
    using P = RegPlan<LOGN>;
    // pass 1: thread j holds n = j + M1 m (m = register index)
    __device__ static __forceinline__ int in_index(int u, int m) { return u + P::M1 * m; }
    // exchange 1: thread j writes register q to slot (q, j); thread (q, j2) = u reads m2 from slot (q, j2 + M2 m2)
    __device__ static __forceinline__ int x1_write(int u, int q) { return q * P::M1 + u; }
    __device__ static __forceinline__ int x1_read(int u, int m2) {
        const int q = u / P::M2, j2 = u - q * P::M2;
        return q * P::M1 + j2 + P::M2 * m2;
    }
    // exchange 2: thread (q, j2) writes register q2 to slot (q, q2, j2) - exactly the slots it read in exchange 1;
    // thread (q, h) reads register i M2 + j2 from slot (q, h G + i, j2)
    __device__ static __forceinline__ int x2_write(int u, int q2) { return x1_read(u, q2); }
    __device__ static __forceinline__ int x2_read(int u, int r) {
        const int q = u / P::M2, h = u - q * P::M2;
        const int i = r / P::M2, j2 = r - i * P::M2;
        return q * P::M1 + (h * P::G + i) * P::M2 + j2;
    }
    // output frequency of register r = i M2 + k3 of thread (q, h)
    __device__ static __forceinline__ int out_freq(int u, int r) {
        if constexpr (P::M2 == 1) {
            return u + 16 * r;  // no third pass: thread (q) holds q2 = r
        } else {
            const int q = u / P::M2, h = u - q * P::M2;
            const int i = r / P::M2, k3 = r - i * P::M2;
            return q + 16 * (h * P::G + i) + 256 * k3;
        }
    }
};

// twiddles after pass 1 (thread j, register q) and pass 2 (thread (q, j2), register q2): v[q] *= w^(j q).
// Four table loads per pass (w^j, w^2j, w^4j, w^8j, each correctly rounded) and eleven products give the fifteen
// factors: the loads of a full table row per factor were the top stall of the column kernels (long scoreboard) and
// 15 % of the shared-memory / L1 wavefronts of the x pass, while the fp64 pipe sits below 30 %.  A factor is the
// product of at most four rounded values (relative error < 5e-16).
__device__ __forceinline__ void twiddle_powers(double2 (&v)[16], double2 w1, double2 w2, double2 w4, double2 w8) {
    const double2 w3 = cmul(w1, w2), w5 = cmul(w4, w1), w6 = cmul(w4, w2), w7 = cmul(w4, w3);
    v[1] = cmul(v[1], w1);
    v[2] = cmul(v[2], w2);
    v[3] = cmul(v[3], w3);
    v[4] = cmul(v[4], w4);
    v[5] = cmul(v[5], w5);
    v[6] = cmul(v[6], w6);
    v[7] = cmul(v[7], w7);
    v[8] = cmul(v[8], w8);
    v[9] = cmul(v[9], cmul(w8, w1));
    v[10] = cmul(v[10], cmul(w8, w2));
    v[11] = cmul(v[11], cmul(w8, w3));
    v[12] = cmul(v[12], cmul(w8, w4));
    v[13] = cmul(v[13], cmul(w8, w5));
    v[14] = cmul(v[14], cmul(w8, w6));
    v[15] = cmul(v[15], cmul(w8, w7));
}
template <int LOGN>
__device__ __forceinline__ void twiddle1(double2 (&v)[16], int u, const double2* __restrict__ t1) {
    using P = RegPlan<LOGN>;
#ifdef FAVA_TW_TABLE
#pragma unroll
    for (int q = 1; q < 16; ++q) v[q] = cmul(v[q], __ldg(t1 + q * P::M1 + u));
#else
    twiddle_powers(v, __ldg(t1 + P::M1 + u), __ldg(t1 + 2 * P::M1 + u), __ldg(t1 + 4 * P::M1 + u), __ldg(t1 + 8 * P::M1 + u));
#endif
}
template <int LOGN>
__device__ __forceinline__ void twiddle2(double2 (&v)[16], int u, const double2* __restrict__ t2) {
    using P = RegPlan<LOGN>;
    if constexpr (P::M2 > 1) {
        const int j2 = u % P::M2;
#ifdef FAVA_TW_TABLE
#pragma unroll
        for (int q2 = 1; q2 < 16; ++q2) v[q2] = cmul(v[q2], __ldg(t2 + q2 * P::M2 + j2));
#else
        twiddle_powers(v, __ldg(t2 + P::M2 + j2), __ldg(t2 + 2 * P::M2 + j2), __ldg(t2 + 4 * P::M2 + j2), __ldg(t2 + 8 * P::M2 + j2));
#endif
    }
}

// Where element index e of this thread's line lives in the exchange buffer.
//   ColAddr  - C interleaved lines (column passes): word phi(e) C + c; conflict-free for C = 8 (16- and 8-byte words)
//              and C = 4 (16-byte words).
//   LineAddr - one line after the other (x pass, lanes run over the element index): word off + skew(e), where the
//              skew shifts every block of 64 by 4 and every block of 16 by one more; with 8-byte words every access
//              pattern of the three exchanges is then conflict-free except the mirrored read of the split (2-way on
//              half of its phases) - checked by enumeration for N = 256 ... 2048 (tools/lab/bank_conflicts.py).
__host__ __device__ __forceinline__ constexpr int line_skew(int e) { return e + 4 * (e >> 6) + ((e >> 4) & 3); }
__host__ __device__ constexpr int line_pitch(int n) { return n < 64 ? 68 : (n / 64) * 68; }

template <int C>
struct ColAddr {
    int c;
    __device__ __forceinline__ int operator()(int e) const { return phi(e) * C + c; }
};
struct LineAddr {
    int off;
    __device__ __forceinline__ int operator()(int e) const { return off + line_skew(e); }
};

// Who must arrive before exchanged words may be read: the whole CTA when the lines of a tile are interleaved over all
// warps (column passes), or only the warps of ONE line when a line owns whole warps (x pass) - the lines of a CTA then
// drift apart and one line's butterflies overlap another line's shared-memory phase.
struct CtaSync {
    __device__ __forceinline__ void operator()() const { __syncthreads(); }
};
template <int THREADS>  // threads of one line; `line` < 3 selects the named barrier 1..3 of this line
struct LineSync {
    int line;
    __device__ __forceinline__ void operator()() const {
        if constexpr (THREADS <= 32) {
            __syncwarp();
        } else {  // immediate barrier ids: with a register id ptxas reserves all 16 barriers of the CTA
            if (line == 0) asm volatile("bar.sync 1, %0;\n" ::"n"(THREADS) : "memory");
            else if (line == 1) asm volatile("bar.sync 2, %0;\n" ::"n"(THREADS) : "memory");
            else asm volatile("bar.sync 3, %0;\n" ::"n"(THREADS) : "memory");
        }
    }
};

// Exchange 2 without shared memory.  The M2 threads (q, j2 = 0..M2-1) that must swap points before pass 3 always sit in
// ONE warp, `stride` lanes apart (column kernels: stride = C, a warp is exactly one q; x-row kernel: stride = 1): thread
// j2 holds q2 = 0..15 and needs q2 = h G + i (h = its own index, i < G) from every j2.  That is the transpose of an
// M2 x M2 matrix of G-point blocks, one row per thread: log2(M2) rounds of __shfl_xor, each swapping the blocks whose
// index bit differs from the thread's, then a register renaming (block j, point i) -> i M2 + j.  No barrier, no bank.
template <int M2>
__device__ __forceinline__ void exchange2_shuffle(double2 (&v)[16], int j2, int stride) {
    constexpr int G = 16 / M2;
#pragma unroll
    for (int bit = 1; bit < M2; bit <<= 1) {
        const bool up = (j2 & bit) != 0;
#pragma unroll
        for (int h0 = 0; h0 < M2; ++h0) {
            if (h0 & bit) continue;
            const int h1 = h0 | bit;
#pragma unroll
            for (int i = 0; i < G; ++i) {
                const double2 send = up ? v[h0 * G + i] : v[h1 * G + i];
                double2 recv;
                recv.x = __shfl_xor_sync(0xffffffffu, send.x, stride * bit);
                recv.y = __shfl_xor_sync(0xffffffffu, send.y, stride * bit);
                if (up) v[h0 * G + i] = recv;
                else v[h1 * G + i] = recv;
            }
        }
    }
    double2 w[16];
#pragma unroll
    for (int j = 0; j < M2; ++j)
#pragma unroll
        for (int i = 0; i < G; ++i) w[i * M2 + j] = v[j * G + i];
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = w[r];
}

// Everything between "v holds the inputs of pass 1" and "v holds the outputs": two exchanges through `xb`
// (complex words).  u = thread's index within the line.  All threads that `sync` joins must call it.
template <int LOGN, class Addr, class Sync = CtaSync>
__device__ __forceinline__ void fft_regs_full(double2 (&v)[16], int u, const Addr& at, double2* __restrict__ xb,
                                              const double2* __restrict__ t1, const double2* __restrict__ t2,
                                              const Sync& sync = Sync()) {
    using P = RegPlan<LOGN>;
    using O = Owner<LOGN>;
    dft16(v);
    twiddle1<LOGN>(v, u, t1);
#pragma unroll
    for (int q = 0; q < 16; ++q) xb[at(O::x1_write(u, q))] = v[q];
    sync();
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = xb[at(O::x1_read(u, m))];
    dft16(v);
    if constexpr (P::M2 > 1) {
        twiddle2<LOGN>(v, u, t2);
#pragma unroll
        for (int q = 0; q < 16; ++q) xb[at(O::x2_write(u, q))] = v[q];  // own slots: no barrier needed before
        sync();
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = xb[at(O::x2_read(u, r))];
        dft_groups<P::M2>(v);
    }
}

// Same with an exchange buffer of 8-byte words (half the size): real and imaginary parts go through it one after the
// other (twice the barriers, same traffic).
template <int LOGN, class Addr, class Sync = CtaSync, int SHFL_STRIDE = 0>
__device__ __forceinline__ void fft_regs_half(double2 (&v)[16], int u, const Addr& at, double* __restrict__ xb,
                                              const double2* __restrict__ t1, const double2* __restrict__ t2,
                                              const Sync& sync = Sync()) {
    using P = RegPlan<LOGN>;
    using O = Owner<LOGN>;
    dft16(v);
    twiddle1<LOGN>(v, u, t1);
#pragma unroll
    for (int q = 0; q < 16; ++q) xb[at(O::x1_write(u, q))] = v[q].x;
    sync();
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m].x = xb[at(O::x1_read(u, m))];  // v[].y still holds pass-1 ownership
    sync();
#pragma unroll
    for (int q = 0; q < 16; ++q) xb[at(O::x1_write(u, q))] = v[q].y;
    sync();
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m].y = xb[at(O::x1_read(u, m))];
    dft16(v);
    if constexpr (P::M2 > 1) {
        twiddle2<LOGN>(v, u, t2);
        if constexpr (SHFL_STRIDE > 0) {
            static_assert(P::M2 * SHFL_STRIDE <= 32, "the exchange group must sit inside one warp");
            exchange2_shuffle<P::M2>(v, u % P::M2, SHFL_STRIDE);
        } else {
#pragma unroll
            for (int q = 0; q < 16; ++q) xb[at(O::x2_write(u, q))] = v[q].x;  // own slots of the last read
            sync();
#pragma unroll
            for (int r = 0; r < 16; ++r) v[r].x = xb[at(O::x2_read(u, r))];  // v[].y still holds pass-2 ownership
            sync();
#pragma unroll
            for (int q = 0; q < 16; ++q) xb[at(O::x2_write(u, q))] = v[q].y;
            sync();
#pragma unroll
            for (int r = 0; r < 16; ++r) v[r].y = xb[at(O::x2_read(u, r))];
        }
        dft_groups<P::M2>(v);
    }
}

// Two-for-one split after the transform of z = a + i b (a, b real lines): the thread gives up its outputs
// (frequency out_freq(u, r)) and receives Z[k] and Z[N - k] for k = u + M1 m, m < 8, i.e. k < N/2; on return
// e[m] = A^[k] = (Z[k] + conj Z[N-k]) / 2 and o[m] = B^[k] = (Z[k] - conj Z[N-k]) / (2i).  8-byte exchange words.
template <int LOGN, class Addr, class Sync = CtaSync>
__device__ __forceinline__ void split_two_for_one(double2 (&v)[16], int u, const Addr& at, double* __restrict__ xb,
                                                  double2 (&e)[8], double2 (&o)[8], const Sync& sync = Sync(),
                                                  double2* mid = nullptr) {
    using P = RegPlan<LOGN>;
    using O = Owner<LOGN>;
    sync();  // the last reads of the transform's exchange are complete
#pragma unroll
    for (int r = 0; r < 16; ++r) xb[at(O::out_freq(u, r))] = v[r].x;
    sync();
    double are[8], bre[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        const int k = u + P::M1 * m;
        are[m] = xb[at(k)];
        bre[m] = xb[at((P::N - k) & (P::N - 1))];
    }
    if (mid) mid->x = xb[at(P::N / 2)];  // Z[N/2], the one element the pairs (k, N - k), k < N/2, leave out
    sync();
#pragma unroll
    for (int r = 0; r < 16; ++r) xb[at(O::out_freq(u, r))] = v[r].y;
    sync();
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        const int k = u + P::M1 * m;
        const double aim = xb[at(k)], bim = xb[at((P::N - k) & (P::N - 1))];
        e[m] = make_double2(0.5 * (are[m] + bre[m]), 0.5 * (aim - bim));
        o[m] = make_double2(0.5 * (aim + bim), 0.5 * (bre[m] - are[m]));
    }
    if (mid) mid->y = xb[at(P::N / 2)];
}

}  // namespace fftc
}  // namespace fava
