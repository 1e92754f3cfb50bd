// k1_explore.cu -- standalone timing harness for variants of the Gray-code Glynn kernel (K1).
// Not part of the library: used on the GPU box to choose the production configuration.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o k1_explore k1_explore.cu
//   ./k1_explore matrix_n30.bin <re> <im>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../theboss_b200/csrc/bp_common.cuh"

#define N 30
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int M, int NCH>
__device__ __forceinline__ cplx prod_rr(const double (&sr)[M], const double (&si)[M]) {
    cplx p[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) { p[c].re = sr[c]; p[c].im = si[c]; }
#pragma unroll
    for (int j = NCH; j < M; ++j) { cplx s = {sr[j], si[j]}; p[j % NCH] = cmul(p[j % NCH], s); }
#pragma unroll
    for (int stride = 1; stride < NCH; stride <<= 1)
#pragma unroll
        for (int c = 0; c + stride < NCH; c += 2 * stride) p[c] = cmul(p[c], p[c + stride]);
    return p[0];
}

__device__ __forceinline__ int ctz64(uint64_t I) {
    const uint32_t Il = (uint32_t)I;
    return Il ? (__ffs((int)Il) - 1) : (31 + __ffs((int)(uint32_t)(I >> 32)));
}

// ---------------------------------------------------------------------------------------------
// variant A: one thread per term stream (production layout), NCH chains
// ---------------------------------------------------------------------------------------------
template <int NCH, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k1_a(const double *__restrict__ A, uint64_t lo, uint64_t hi, uint64_t span, double *__restrict__ partials) {
    __shared__ double2 sA2[N * N];
    __shared__ double red[4 * (THREADS / 32)];
    for (int e = threadIdx.x; e < N * N; e += THREADS) {
        double2 v = reinterpret_cast<const double2 *>(A)[e];
        sA2[e] = make_double2(2.0 * v.x, 2.0 * v.y);
    }
    __syncthreads();
    const uint64_t gtid = (uint64_t)blockIdx.x * THREADS + threadIdx.x;
    const uint64_t start = lo + gtid * span;
    dd acc_re = {0.0, 0.0}, acc_im = {0.0, 0.0};
    if (start < hi) {
        const uint64_t end = (hi - start < span) ? hi : start + span;
        double sr[N], si[N];
#pragma unroll
        for (int j = 0; j < N; ++j) { sr[j] = 0.0; si[j] = 0.0; }
        const uint64_t g0 = start ^ (start >> 1);
#pragma unroll 1
        for (int i = 0; i < N; ++i) {
            const double sg = ((g0 >> i) & 1ull) ? -0.5 : 0.5;
            const double2 *row = sA2 + i * N;
#pragma unroll
            for (int j = 0; j < N; ++j) { const double2 a = row[j]; sr[j] = fma(sg, a.x, sr[j]); si[j] = fma(sg, a.y, si[j]); }
        }
        cplx p = prod_rr<N, NCH>(sr, si);
        const double ts0 = (start & 1ull) ? -1.0 : 1.0;
        double wr = ts0 * p.re, wi = ts0 * p.im;
#pragma unroll 1
        for (uint64_t I = start + 1; I < end; ++I) {
            const uint32_t Il = (uint32_t)I;
            if ((Il & 63u) == 0u) { acc_re = dd_add_d(acc_re, wr); acc_im = dd_add_d(acc_im, wi); wr = 0.0; wi = 0.0; }
            const int r = ctz64(I);
            const double sg = ((I >> (r + 1)) & 1ull) ? 1.0 : -1.0;
            const double2 *row = sA2 + r * N;
#pragma unroll
            for (int j = 0; j < N; ++j) { const double2 a = row[j]; sr[j] = fma(sg, a.x, sr[j]); si[j] = fma(sg, a.y, si[j]); }
            p = prod_rr<N, NCH>(sr, si);
            const double ts = (Il & 1u) ? -1.0 : 1.0;
            wr = fma(ts, p.re, wr); wi = fma(ts, p.im, wi);
        }
        acc_re = dd_add_d(acc_re, wr); acc_im = dd_add_d(acc_im, wi);
    }
    block_reduce_dd(acc_re, acc_im, red);
    if (threadIdx.x == 0) { double *o = partials + 4 * (size_t)blockIdx.x; o[0] = acc_re.hi; o[1] = acc_re.lo; o[2] = acc_im.hi; o[3] = acc_im.lo; }
}

// ---------------------------------------------------------------------------------------------
// variant C: one thread per stream, two terms per iteration (even: general row, odd: row 0)
// ---------------------------------------------------------------------------------------------
template <int NCH, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k1_c(const double *__restrict__ A, uint64_t lo, uint64_t hi, uint64_t span, double *__restrict__ partials) {
    __shared__ double2 sA2[N * N];
    __shared__ double red[4 * (THREADS / 32)];
    for (int e = threadIdx.x; e < N * N; e += THREADS) {
        double2 v = reinterpret_cast<const double2 *>(A)[e];
        sA2[e] = make_double2(2.0 * v.x, 2.0 * v.y);
    }
    __syncthreads();
    const uint64_t gtid = (uint64_t)blockIdx.x * THREADS + threadIdx.x;
    const uint64_t start = lo + gtid * span;   // span and lo are even
    dd acc_re = {0.0, 0.0}, acc_im = {0.0, 0.0};
    if (start < hi) {
        const uint64_t end = (hi - start < span) ? hi : start + span;
        double sr[N], si[N];
#pragma unroll
        for (int j = 0; j < N; ++j) { sr[j] = 0.0; si[j] = 0.0; }
        const uint64_t g0 = start ^ (start >> 1);
#pragma unroll 1
        for (int i = 0; i < N; ++i) {
            const double sg = ((g0 >> i) & 1ull) ? -0.5 : 0.5;
            const double2 *row = sA2 + i * N;
#pragma unroll
            for (int j = 0; j < N; ++j) { const double2 a = row[j]; sr[j] = fma(sg, a.x, sr[j]); si[j] = fma(sg, a.y, si[j]); }
        }
        double wr = 0.0, wi = 0.0;
        bool first = true;
#pragma unroll 1
        for (uint64_t I = start; I < end; I += 2) {   // I even, I + 1 odd
            const uint32_t Il = (uint32_t)I;
            if ((Il & 63u) == 0u) { acc_re = dd_add_d(acc_re, wr); acc_im = dd_add_d(acc_im, wi); wr = 0.0; wi = 0.0; }
            const int r = first ? 0 : ctz64(I);
            const double sgr = first ? 0.0 : (((I >> (r + 1)) & 1ull) ? 1.0 : -1.0);
            first = false;
            const double sg0 = (((I + 1) >> 1) & 1ull) ? 1.0 : -1.0;
            const double2 *row = sA2 + r * N;
            cplx pa[NCH], pb[NCH];
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const double2 a = row[j];
                const double2 b = sA2[j];
                sr[j] = fma(sgr, a.x, sr[j]); si[j] = fma(sgr, a.y, si[j]);
                cplx s = {sr[j], si[j]};
                if (j < NCH) pa[j] = s; else pa[j % NCH] = cmul(pa[j % NCH], s);
                sr[j] = fma(sg0, b.x, sr[j]); si[j] = fma(sg0, b.y, si[j]);
                cplx s2 = {sr[j], si[j]};
                if (j < NCH) pb[j] = s2; else pb[j % NCH] = cmul(pb[j % NCH], s2);
            }
#pragma unroll
            for (int stride = 1; stride < NCH; stride <<= 1)
#pragma unroll
                for (int c = 0; c + stride < NCH; c += 2 * stride) { pa[c] = cmul(pa[c], pa[c + stride]); pb[c] = cmul(pb[c], pb[c + stride]); }
            wr += pa[0].re - pb[0].re;
            wi += pa[0].im - pb[0].im;
        }
        acc_re = dd_add_d(acc_re, wr); acc_im = dd_add_d(acc_im, wi);
    }
    block_reduce_dd(acc_re, acc_im, red);
    if (threadIdx.x == 0) { double *o = partials + 4 * (size_t)blockIdx.x; o[0] = acc_re.hi; o[1] = acc_re.lo; o[2] = acc_im.hi; o[3] = acc_im.lo; }
}

// ---------------------------------------------------------------------------------------------
// variant B / D: two lanes per term stream (15 columns each); PAIR = two terms per iteration
// ---------------------------------------------------------------------------------------------
template <int NCH, int THREADS, int MINB, bool PAIR>
__global__ void __launch_bounds__(THREADS, MINB)
k1_b(const double *__restrict__ A, uint64_t lo, uint64_t hi, uint64_t span, double *__restrict__ partials) {
    constexpr int H = (N + 1) / 2;
    __shared__ double2 sA2[N * 2 * H];
    __shared__ double red[4 * (THREADS / 32)];
    for (int e = threadIdx.x; e < N * 2 * H; e += THREADS) {
        const int i = e / (2 * H), j = e - i * 2 * H;
        double2 v = make_double2(0.0, 0.0);
        if (j < N) v = reinterpret_cast<const double2 *>(A)[i * N + j];
        sA2[e] = make_double2(2.0 * v.x, 2.0 * v.y);
    }
    __syncthreads();
    const int half = threadIdx.x & 1;
    const uint64_t stream = ((uint64_t)blockIdx.x * THREADS + threadIdx.x) >> 1;
    const uint64_t start = lo + stream * span;
    dd acc_re = {0.0, 0.0}, acc_im = {0.0, 0.0};
    if (start < hi) {   // uniform for both lanes of a stream
        const uint64_t end = (hi - start < span) ? hi : start + span;
        const unsigned pm = 3u << ((threadIdx.x & 31) & ~1);
        double sr[H], si[H];
#pragma unroll
        for (int j = 0; j < H; ++j) { sr[j] = 0.0; si[j] = 0.0; }
        const uint64_t g0 = start ^ (start >> 1);
#pragma unroll 1
        for (int i = 0; i < N; ++i) {
            const double sg = ((g0 >> i) & 1ull) ? -0.5 : 0.5;
            const double2 *row = sA2 + i * 2 * H + half * H;
#pragma unroll
            for (int j = 0; j < H; ++j) { const double2 a = row[j]; sr[j] = fma(sg, a.x, sr[j]); si[j] = fma(sg, a.y, si[j]); }
        }
        if ((N & 1) && half) { sr[H - 1] = 1.0; si[H - 1] = 0.0; }   // padding column
        double wr = 0.0, wi = 0.0;
        if (!PAIR) {
            cplx p = prod_rr<H, NCH>(sr, si);
            cplx q = {__shfl_xor_sync(pm, p.re, 1), __shfl_xor_sync(pm, p.im, 1)};
            cplx f = cmul(p, q);
            const double ts0 = (start & 1ull) ? -1.0 : 1.0;
            wr = ts0 * f.re; wi = ts0 * f.im;
#pragma unroll 1
            for (uint64_t I = start + 1; I < end; ++I) {
                const uint32_t Il = (uint32_t)I;
                if ((Il & 63u) == 0u) { acc_re = dd_add_d(acc_re, wr); acc_im = dd_add_d(acc_im, wi); wr = 0.0; wi = 0.0; }
                const int r = ctz64(I);
                const double sg = ((I >> (r + 1)) & 1ull) ? 1.0 : -1.0;
                const double2 *row = sA2 + r * 2 * H + half * H;
#pragma unroll
                for (int j = 0; j < H; ++j) { const double2 a = row[j]; sr[j] = fma(sg, a.x, sr[j]); si[j] = fma(sg, a.y, si[j]); }
                p = prod_rr<H, NCH>(sr, si);
                q.re = __shfl_xor_sync(pm, p.re, 1); q.im = __shfl_xor_sync(pm, p.im, 1);
                f = cmul(p, q);
                const double ts = (Il & 1u) ? -1.0 : 1.0;
                wr = fma(ts, f.re, wr); wi = fma(ts, f.im, wi);
            }
        } else {
            bool first = true;
            const double2 *row0 = sA2 + half * H;
#pragma unroll 1
            for (uint64_t I = start; I < end; I += 2) {
                const uint32_t Il = (uint32_t)I;
                if ((Il & 63u) == 0u) { acc_re = dd_add_d(acc_re, wr); acc_im = dd_add_d(acc_im, wi); wr = 0.0; wi = 0.0; }
                const int r = first ? 0 : ctz64(I);
                const double sgr = first ? 0.0 : (((I >> (r + 1)) & 1ull) ? 1.0 : -1.0);
                first = false;
                const double sg0 = (((I + 1) >> 1) & 1ull) ? 1.0 : -1.0;
                const double2 *row = sA2 + r * 2 * H + half * H;
                cplx pa[NCH], pb[NCH];
#pragma unroll
                for (int j = 0; j < H; ++j) {
                    const double2 a = row[j];
                    const double2 b = row0[j];
                    sr[j] = fma(sgr, a.x, sr[j]); si[j] = fma(sgr, a.y, si[j]);
                    cplx s = {sr[j], si[j]};
                    if (j < NCH) pa[j] = s; else pa[j % NCH] = cmul(pa[j % NCH], s);
                    sr[j] = fma(sg0, b.x, sr[j]); si[j] = fma(sg0, b.y, si[j]);
                    cplx s2 = {sr[j], si[j]};
                    if (j < NCH) pb[j] = s2; else pb[j % NCH] = cmul(pb[j % NCH], s2);
                }
#pragma unroll
                for (int stride = 1; stride < NCH; stride <<= 1)
#pragma unroll
                    for (int c = 0; c + stride < NCH; c += 2 * stride) { pa[c] = cmul(pa[c], pa[c + stride]); pb[c] = cmul(pb[c], pb[c + stride]); }
                cplx qa = {__shfl_xor_sync(pm, pa[0].re, 1), __shfl_xor_sync(pm, pa[0].im, 1)};
                cplx qb = {__shfl_xor_sync(pm, pb[0].re, 1), __shfl_xor_sync(pm, pb[0].im, 1)};
                cplx fa = cmul(pa[0], qa), fb = cmul(pb[0], qb);
                wr += fa.re - fb.re;
                wi += fa.im - fb.im;
            }
        }
        acc_re = dd_add_d(acc_re, wr); acc_im = dd_add_d(acc_im, wi);
        if (half) { acc_re.hi = acc_re.lo = acc_im.hi = acc_im.lo = 0.0; }   // both lanes hold the same sum
    }
    block_reduce_dd(acc_re, acc_im, red);
    if (threadIdx.x == 0) { double *o = partials + 4 * (size_t)blockIdx.x; o[0] = acc_re.hi; o[1] = acc_re.lo; o[2] = acc_im.hi; o[3] = acc_im.lo; }
}


// ---------------------------------------------------------------------------------------------
// variant E: UNR terms per iteration.  Inside an aligned block of UNR Gray steps the flipped rows
// are compile-time (ctz(u)), so rows 0 .. log2(UNR)-1 are read as CONSTANT-BANK OPERANDS of the
// FP64 instruction (no load instruction, no address arithmetic); only the last step of a block
// flips a run-time row, fetched from shared memory.
// ---------------------------------------------------------------------------------------------
__constant__ double2 cA2[N * N];

__host__ __device__ constexpr int cx_ctz(int u) { return (u & 1) ? 0 : 1 + cx_ctz(u >> 1); }
__host__ __device__ constexpr int cx_log2(int u) { return (u <= 1) ? 0 : 1 + cx_log2(u >> 1); }

template <int ROW>
__device__ __forceinline__ void upd_const(double (&sr)[N], double (&si)[N], double sg) {
#pragma unroll
    for (int j = 0; j < N; ++j) { sr[j] = fma(sg, cA2[ROW * N + j].x, sr[j]); si[j] = fma(sg, cA2[ROW * N + j].y, si[j]); }
}
template <int ROW, bool PLUS>
__device__ __forceinline__ void upd_const_fixed(double (&sr)[N], double (&si)[N]) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
        if (PLUS) { sr[j] += cA2[ROW * N + j].x; si[j] += cA2[ROW * N + j].y; }
        else      { sr[j] -= cA2[ROW * N + j].x; si[j] -= cA2[ROW * N + j].y; }
    }
}

template <int NCH, int UNR, int U>
struct EStep {
    static __device__ __forceinline__ void run(double (&sr)[N], double (&si)[N], double &wr, double &wi, double sg_half) {
        if constexpr (U < UNR) {
            constexpr int ROW = cx_ctz(U);
            if constexpr (U == UNR / 2) {
                upd_const<ROW>(sr, si, sg_half);            // sign = bit log2(UNR) of the block base (run-time)
            } else {
                constexpr bool PLUS = ((U >> (ROW + 1)) & 1) != 0;
                upd_const_fixed<ROW, PLUS>(sr, si);
            }
            cplx p = prod_rr<N, NCH>(sr, si);
            if constexpr (U & 1) { wr -= p.re; wi -= p.im; } else { wr += p.re; wi += p.im; }
            EStep<NCH, UNR, U + 1>::run(sr, si, wr, wi, sg_half);
        }
    }
};

template <int NCH, int UNR, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k1_e(const double *__restrict__ A, uint64_t lo, uint64_t hi, uint64_t span, double *__restrict__ partials) {
    __shared__ double2 sA2[N * N];
    __shared__ double red[4 * (THREADS / 32)];
    for (int e = threadIdx.x; e < N * N; e += THREADS) {
        double2 v = reinterpret_cast<const double2 *>(A)[e];
        sA2[e] = make_double2(2.0 * v.x, 2.0 * v.y);
    }
    __syncthreads();
    const uint64_t gtid = (uint64_t)blockIdx.x * THREADS + threadIdx.x;
    const uint64_t start = lo + gtid * span;     // lo, span: multiples of 64
    dd acc_re = {0.0, 0.0}, acc_im = {0.0, 0.0};
    if (start < hi) {
        const uint64_t end = (hi - start < span) ? hi : start + span;   // multiple of UNR
        double sr[N], si[N];
#pragma unroll
        for (int j = 0; j < N; ++j) { sr[j] = 0.0; si[j] = 0.0; }
        const uint64_t g0 = start ^ (start >> 1);
#pragma unroll 1
        for (int i = 0; i < N; ++i) {
            const double sg = ((g0 >> i) & 1ull) ? -0.5 : 0.5;
            const double2 *row = sA2 + i * N;
#pragma unroll
            for (int j = 0; j < N; ++j) { const double2 a = row[j]; sr[j] = fma(sg, a.x, sr[j]); si[j] = fma(sg, a.y, si[j]); }
        }
        cplx p = prod_rr<N, NCH>(sr, si);
        double wr = p.re, wi = p.im;             // term `start` (even index: +)
#pragma unroll 1
        for (uint64_t I0 = start;;) {
            const double sg_half = ((I0 >> cx_log2(UNR)) & 1ull) ? 1.0 : -1.0;
            EStep<NCH, UNR, 1>::run(sr, si, wr, wi, sg_half);
            I0 += UNR;
            if (I0 >= end) break;
            if (((uint32_t)I0 & 63u) == 0u) { acc_re = dd_add_d(acc_re, wr); acc_im = dd_add_d(acc_im, wi); wr = 0.0; wi = 0.0; }
            const int r = ctz64(I0);
            const double sg = ((I0 >> (r + 1)) & 1ull) ? 1.0 : -1.0;
            const double2 *row = sA2 + r * N;
#pragma unroll
            for (int j = 0; j < N; ++j) { const double2 a = row[j]; sr[j] = fma(sg, a.x, sr[j]); si[j] = fma(sg, a.y, si[j]); }
            p = prod_rr<N, NCH>(sr, si);
            wr += p.re; wi += p.im;
        }
        acc_re = dd_add_d(acc_re, wr); acc_im = dd_add_d(acc_im, wi);
    }
    block_reduce_dd(acc_re, acc_im, red);
    if (threadIdx.x == 0) { double *o = partials + 4 * (size_t)blockIdx.x; o[0] = acc_re.hi; o[1] = acc_re.lo; o[2] = acc_im.hi; o[3] = acc_im.lo; }
}


// variant E3: E with 32-bit loop bookkeeping
template <int NCH, int UNR, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k1_e3(const double *__restrict__ A, uint64_t lo, uint64_t hi, uint64_t span, double *__restrict__ partials) {
    __shared__ double2 sA2[N * N];
    __shared__ double red[4 * (THREADS / 32)];
    for (int e = threadIdx.x; e < N * N; e += THREADS) {
        double2 v = reinterpret_cast<const double2 *>(A)[e];
        sA2[e] = make_double2(2.0 * v.x, 2.0 * v.y);
    }
    __syncthreads();
    const uint64_t gtid = (uint64_t)blockIdx.x * THREADS + threadIdx.x;
    const uint64_t start = lo + gtid * span;     // lo, span: multiples of 64
    dd acc_re = {0.0, 0.0}, acc_im = {0.0, 0.0};
    if (start < hi) {
        const uint64_t end = (hi - start < span) ? hi : start + span;   // multiple of UNR
        double sr[N], si[N];
#pragma unroll
        for (int j = 0; j < N; ++j) { sr[j] = 0.0; si[j] = 0.0; }
        const uint64_t g0 = start ^ (start >> 1);
#pragma unroll 1
        for (int i = 0; i < N; ++i) {
            const double sg = ((g0 >> i) & 1ull) ? -0.5 : 0.5;
            const double2 *row = sA2 + i * N;
#pragma unroll
            for (int j = 0; j < N; ++j) { const double2 a = row[j]; sr[j] = fma(sg, a.x, sr[j]); si[j] = fma(sg, a.y, si[j]); }
        }
        cplx p = prod_rr<N, NCH>(sr, si);
        double wr = p.re, wi = p.im;             // term `start` (even index: +)
        // 32-bit bookkeeping: block counter counts down, the low word of the step index drives ctz / signs;
        // the (rare) blocks whose low word is zero take the 64-bit path.
        uint32_t nblk = (uint32_t)((end - start) / UNR);
        uint32_t lo32 = (uint32_t)start;
        uint32_t hi32 = (uint32_t)(start >> 32);
#pragma unroll 1
        for (;;) {
            const double sg_half = ((lo32 >> cx_log2(UNR)) & 1u) ? 1.0 : -1.0;
            EStep<NCH, UNR, 1>::run(sr, si, wr, wi, sg_half);
            lo32 += UNR;
            if (--nblk == 0) break;
            if ((lo32 & 63u) == 0u) { acc_re = dd_add_d(acc_re, wr); acc_im = dd_add_d(acc_im, wi); wr = 0.0; wi = 0.0; }
            int r; double sg;
            if (lo32 & 0x7fffffffu) {
                r = __ffs((int)lo32) - 1;                       // <= 30, so bit r + 1 is in the low word
                sg = ((lo32 >> (r + 1)) & 1u) ? 1.0 : -1.0;
            } else {
                if (lo32 == 0u) ++hi32;
                const uint64_t I0 = ((uint64_t)hi32 << 32) | lo32;
                r = ctz64(I0);
                sg = ((I0 >> (r + 1)) & 1ull) ? 1.0 : -1.0;
            }
            const double2 *row = sA2 + r * N;
#pragma unroll
            for (int j = 0; j < N; ++j) { const double2 a = row[j]; sr[j] = fma(sg, a.x, sr[j]); si[j] = fma(sg, a.y, si[j]); }
            p = prod_rr<N, NCH>(sr, si);
            wr += p.re; wi += p.im;
        }
        acc_re = dd_add_d(acc_re, wr); acc_im = dd_add_d(acc_im, wi);
    }
    block_reduce_dd(acc_re, acc_im, red);
    if (threadIdx.x == 0) { double *o = partials + 4 * (size_t)blockIdx.x; o[0] = acc_re.hi; o[1] = acc_re.lo; o[2] = acc_im.hi; o[3] = acc_im.lo; }
}



// ---------------------------------------------------------------------------------------------
// variant E4: like E (UNR = 4), but inside a 64-step window the block-closing flips (rows 2..5) also
// come from the constant bank: a warp-uniform switch over the four possible rows gives every case
// compile-time addresses, so those loads go through the uniform datapath (no vector-register-file
// write traffic).  Only the window-boundary flip (row >= 6, once per 64 steps) uses shared memory.
// ---------------------------------------------------------------------------------------------
template <int NCH, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k1_e4(const double *__restrict__ A, uint64_t lo, uint64_t hi, uint64_t span, double *__restrict__ partials) {
    __shared__ double2 sA2[N * N];
    __shared__ double red[4 * (THREADS / 32)];
    for (int e = threadIdx.x; e < N * N; e += THREADS) {
        double2 v = reinterpret_cast<const double2 *>(A)[e];
        sA2[e] = make_double2(2.0 * v.x, 2.0 * v.y);
    }
    __syncthreads();
    const uint64_t gtid = (uint64_t)blockIdx.x * THREADS + threadIdx.x;
    const uint64_t start = lo + gtid * span;     // lo, span, hi: multiples of 64
    dd acc_re = {0.0, 0.0}, acc_im = {0.0, 0.0};
    if (start < hi) {
        const uint64_t end = (hi - start < span) ? hi : start + span;
        double sr[N], si[N];
#pragma unroll
        for (int j = 0; j < N; ++j) { sr[j] = 0.0; si[j] = 0.0; }
        const uint64_t g0 = start ^ (start >> 1);
#pragma unroll 1
        for (int i = 0; i < N; ++i) {
            const double sg = ((g0 >> i) & 1ull) ? -0.5 : 0.5;
            const double2 *row = sA2 + i * N;
#pragma unroll
            for (int j = 0; j < N; ++j) { const double2 a = row[j]; sr[j] = fma(sg, a.x, sr[j]); si[j] = fma(sg, a.y, si[j]); }
        }
        double wr = 0.0, wi = 0.0;
        const uint32_t nwin = (uint32_t)((end - start) >> 6);
#pragma unroll 1
        for (uint32_t w = 0; w < nwin; ++w) {
            const uint64_t Iw = start + ((uint64_t)w << 6);
            if (w > 0) {
                acc_re = dd_add_d(acc_re, wr); acc_im = dd_add_d(acc_im, wi); wr = 0.0; wi = 0.0;
                const int r = ctz64(Iw);
                const double sg = ((Iw >> (r + 1)) & 1ull) ? 1.0 : -1.0;
                const double2 *row = sA2 + r * N;
#pragma unroll
                for (int j = 0; j < N; ++j) { const double2 a = row[j]; sr[j] = fma(sg, a.x, sr[j]); si[j] = fma(sg, a.y, si[j]); }
            }
            const double sg_top = ((Iw >> 6) & 1ull) ? 1.0 : -1.0;
#pragma unroll 1
            for (uint32_t q = 0; q < 16; ++q) {
                if (q > 0) {
                    const uint32_t t = q << 2;
                    const int c = __ffs((int)q) - 1;                         // 0..3 -> rows 2..5
                    const double sgu = ((t >> (c + 3)) & 1u) ? 1.0 : -1.0;
                    switch (c) {
                        case 0: upd_const<2>(sr, si, sgu); break;
                        case 1: upd_const<3>(sr, si, sgu); break;
                        case 2: upd_const<4>(sr, si, sgu); break;
                        default: upd_const<5>(sr, si, sg_top); break;
                    }
                }
                cplx p = prod_rr<N, NCH>(sr, si);
                wr += p.re; wi += p.im;
                const double sg_half = (q & 1u) ? 1.0 : -1.0;                // bit 2 of the step offset
                EStep<NCH, 4, 1>::run(sr, si, wr, wi, sg_half);
            }
        }
        acc_re = dd_add_d(acc_re, wr); acc_im = dd_add_d(acc_im, wi);
    }
    block_reduce_dd(acc_re, acc_im, red);
    if (threadIdx.x == 0) { double *o = partials + 4 * (size_t)blockIdx.x; o[0] = acc_re.hi; o[1] = acc_re.lo; o[2] = acc_im.hi; o[3] = acc_im.lo; }
}


// ---------------------------------------------------------------------------------------------
// variant E5: E (UNR = 4) with the loop nest of E4 -- outer loop over 64-step windows (64-bit, per thread),
// inner loop over the 16 blocks of a window with a small uniform counter that yields the block-closing row
// (2 + ctz(q)) and its sign without 64-bit arithmetic; the row itself still comes from shared memory.
// ---------------------------------------------------------------------------------------------
template <int NCH, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k1_e5(const double *__restrict__ A, uint64_t lo, uint64_t hi, uint64_t span, double *__restrict__ partials) {
    __shared__ double2 sA2[N * N];
    __shared__ double red[4 * (THREADS / 32)];
    for (int e = threadIdx.x; e < N * N; e += THREADS) {
        double2 v = reinterpret_cast<const double2 *>(A)[e];
        sA2[e] = make_double2(2.0 * v.x, 2.0 * v.y);
    }
    __syncthreads();
    const uint64_t gtid = (uint64_t)blockIdx.x * THREADS + threadIdx.x;
    const uint64_t start = lo + gtid * span;     // lo, span, hi: multiples of 64
    dd acc_re = {0.0, 0.0}, acc_im = {0.0, 0.0};
    if (start < hi) {
        const uint64_t end = (hi - start < span) ? hi : start + span;
        double sr[N], si[N];
#pragma unroll
        for (int j = 0; j < N; ++j) { sr[j] = 0.0; si[j] = 0.0; }
        const uint64_t g0 = start ^ (start >> 1);
#pragma unroll 1
        for (int i = 0; i < N; ++i) {
            const double sg = ((g0 >> i) & 1ull) ? -0.5 : 0.5;
            const double2 *row = sA2 + i * N;
#pragma unroll
            for (int j = 0; j < N; ++j) { const double2 a = row[j]; sr[j] = fma(sg, a.x, sr[j]); si[j] = fma(sg, a.y, si[j]); }
        }
        double wr = 0.0, wi = 0.0;
        const uint32_t nwin = (uint32_t)((end - start) >> 6);
        uint64_t Iw = start;
#pragma unroll 1
        for (uint32_t w = 0; w < nwin; ++w, Iw += 64) {
            int r = 0;
            double sg = 0.0;
            if (w > 0) {
                acc_re = dd_add_d(acc_re, wr); acc_im = dd_add_d(acc_im, wi); wr = 0.0; wi = 0.0;
                r = ctz64(Iw);
                sg = ((Iw >> (r + 1)) & 1ull) ? 1.0 : -1.0;
            }
            const double sg_top = ((Iw >> 6) & 1ull) ? 1.0 : -1.0;
#pragma unroll 1
            for (uint32_t q = 0; q < 16; ++q) {
                // block-closing flip of the previous block: q = 0 -> window boundary (r, sg above; nothing for the very first block)
                if (q > 0) {
                    const int c = __ffs((int)q) - 1;                         // 0..3 -> rows 2..5
                    r = 2 + c;
                    sg = (c == 3) ? sg_top : (((q >> (c + 1)) & 1u) ? 1.0 : -1.0);
                }
                const double2 *row = sA2 + r * N;
#pragma unroll
                for (int j = 0; j < N; ++j) { const double2 a = row[j]; sr[j] = fma(sg, a.x, sr[j]); si[j] = fma(sg, a.y, si[j]); }
                cplx p = prod_rr<N, NCH>(sr, si);
                wr += p.re; wi += p.im;
                const double sg_half = (q & 1u) ? 1.0 : -1.0;                // bit 2 of the step offset
                EStep<NCH, 4, 1>::run(sr, si, wr, wi, sg_half);
            }
        }
        acc_re = dd_add_d(acc_re, wr); acc_im = dd_add_d(acc_im, wi);
    }
    block_reduce_dd(acc_re, acc_im, red);
    if (threadIdx.x == 0) { double *o = partials + 4 * (size_t)blockIdx.x; o[0] = acc_re.hi; o[1] = acc_re.lo; o[2] = acc_im.hi; o[3] = acc_im.lo; }
}


// ---------------------------------------------------------------------------------------------
// variant E6: E (UNR = 4) tuned for 3 warps/SMSP (384 threads, 168 registers): ONE product chain whose last
// multiplication is fused into the window accumulator; SMEMACC keeps the per-thread double-double
// accumulators in shared memory (touched once per 64 steps) to free 8 registers.
// ---------------------------------------------------------------------------------------------
template <bool PLUS>
__device__ __forceinline__ void e6_prod_acc(const double (&sr)[N], const double (&si)[N], double &wr, double &wi) {
    cplx p = {sr[0], si[0]};
#pragma unroll
    for (int j = 1; j < N - 1; ++j) { cplx s = {sr[j], si[j]}; p = cmul(p, s); }
    cplx last = {sr[N - 1], si[N - 1]};
    if (PLUS) cmul_acc(wr, wi, p, last); else cmul_sub(wr, wi, p, last);
}

template <int THREADS, bool SMEMACC>
__global__ void __launch_bounds__(THREADS, 1)
k1_e6(const double *__restrict__ A, uint64_t lo, uint64_t hi, uint64_t span, double *__restrict__ partials) {
    __shared__ double2 sA2[N * N];
    __shared__ double red[4 * (THREADS / 32)];
    __shared__ double accs[SMEMACC ? 4 * THREADS : 4];
    for (int e = threadIdx.x; e < N * N; e += THREADS) {
        double2 v = reinterpret_cast<const double2 *>(A)[e];
        sA2[e] = make_double2(2.0 * v.x, 2.0 * v.y);
    }
    if (SMEMACC) { accs[threadIdx.x] = 0.0; accs[THREADS + threadIdx.x] = 0.0; accs[2 * THREADS + threadIdx.x] = 0.0; accs[3 * THREADS + threadIdx.x] = 0.0; }
    __syncthreads();
    const uint64_t gtid = (uint64_t)blockIdx.x * THREADS + threadIdx.x;
    const uint64_t start = lo + gtid * span;
    dd acc_re = {0.0, 0.0}, acc_im = {0.0, 0.0};
    if (start < hi) {
        const uint64_t end = (hi - start < span) ? hi : start + span;
        double sr[N], si[N];
#pragma unroll
        for (int j = 0; j < N; ++j) { sr[j] = 0.0; si[j] = 0.0; }
        const uint64_t g0 = start ^ (start >> 1);
#pragma unroll 1
        for (int i = 0; i < N; ++i) {
            const double sg = ((g0 >> i) & 1ull) ? -0.5 : 0.5;
            const double2 *row = sA2 + i * N;
#pragma unroll
            for (int j = 0; j < N; ++j) { const double2 a = row[j]; sr[j] = fma(sg, a.x, sr[j]); si[j] = fma(sg, a.y, si[j]); }
        }
        double wr = 0.0, wi = 0.0;
        e6_prod_acc<true>(sr, si, wr, wi);
#pragma unroll 1
        for (uint64_t I0 = start;;) {
            upd_const_fixed<0, false>(sr, si);
            e6_prod_acc<false>(sr, si, wr, wi);
            upd_const<1>(sr, si, ((I0 >> 2) & 1ull) ? 1.0 : -1.0);
            e6_prod_acc<true>(sr, si, wr, wi);
            upd_const_fixed<0, true>(sr, si);
            e6_prod_acc<false>(sr, si, wr, wi);
            I0 += 4;
            if (I0 >= end) break;
            if (((uint32_t)I0 & 63u) == 0u) {
                if (SMEMACC) {
                    dd a = {accs[threadIdx.x], accs[THREADS + threadIdx.x]}, b = {accs[2 * THREADS + threadIdx.x], accs[3 * THREADS + threadIdx.x]};
                    a = dd_add_d(a, wr); b = dd_add_d(b, wi);
                    accs[threadIdx.x] = a.hi; accs[THREADS + threadIdx.x] = a.lo; accs[2 * THREADS + threadIdx.x] = b.hi; accs[3 * THREADS + threadIdx.x] = b.lo;
                } else { acc_re = dd_add_d(acc_re, wr); acc_im = dd_add_d(acc_im, wi); }
                wr = 0.0; wi = 0.0;
            }
            const int r = ctz64(I0);
            const double sg = ((I0 >> (r + 1)) & 1ull) ? 1.0 : -1.0;
            const double2 *row = sA2 + r * N;
#pragma unroll
            for (int j = 0; j < N; ++j) { const double2 a = row[j]; sr[j] = fma(sg, a.x, sr[j]); si[j] = fma(sg, a.y, si[j]); }
            e6_prod_acc<true>(sr, si, wr, wi);
        }
        if (SMEMACC) { acc_re.hi = accs[threadIdx.x]; acc_re.lo = accs[THREADS + threadIdx.x]; acc_im.hi = accs[2 * THREADS + threadIdx.x]; acc_im.lo = accs[3 * THREADS + threadIdx.x]; }
        acc_re = dd_add_d(acc_re, wr); acc_im = dd_add_d(acc_im, wi);
    }
    block_reduce_dd(acc_re, acc_im, red);
    if (threadIdx.x == 0) { double *o = partials + 4 * (size_t)blockIdx.x; o[0] = acc_re.hi; o[1] = acc_re.lo; o[2] = acc_im.hi; o[3] = acc_im.lo; }
}

// variant E7: E6 with 32-bit loop bookkeeping
template <int THREADS, bool SMEMACC>
__global__ void __launch_bounds__(THREADS, 1)
k1_e7(const double *__restrict__ A, uint64_t lo, uint64_t hi, uint64_t span, double *__restrict__ partials) {
    __shared__ double2 sA2[N * N];
    __shared__ double red[4 * (THREADS / 32)];
    __shared__ double accs[SMEMACC ? 4 * THREADS : 4];
    for (int e = threadIdx.x; e < N * N; e += THREADS) {
        double2 v = reinterpret_cast<const double2 *>(A)[e];
        sA2[e] = make_double2(2.0 * v.x, 2.0 * v.y);
    }
    if (SMEMACC) { accs[threadIdx.x] = 0.0; accs[THREADS + threadIdx.x] = 0.0; accs[2 * THREADS + threadIdx.x] = 0.0; accs[3 * THREADS + threadIdx.x] = 0.0; }
    __syncthreads();
    const uint64_t gtid = (uint64_t)blockIdx.x * THREADS + threadIdx.x;
    const uint64_t start = lo + gtid * span;
    dd acc_re = {0.0, 0.0}, acc_im = {0.0, 0.0};
    if (start < hi) {
        const uint64_t end = (hi - start < span) ? hi : start + span;
        double sr[N], si[N];
#pragma unroll
        for (int j = 0; j < N; ++j) { sr[j] = 0.0; si[j] = 0.0; }
        const uint64_t g0 = start ^ (start >> 1);
#pragma unroll 1
        for (int i = 0; i < N; ++i) {
            const double sg = ((g0 >> i) & 1ull) ? -0.5 : 0.5;
            const double2 *row = sA2 + i * N;
#pragma unroll
            for (int j = 0; j < N; ++j) { const double2 a = row[j]; sr[j] = fma(sg, a.x, sr[j]); si[j] = fma(sg, a.y, si[j]); }
        }
        double wr = 0.0, wi = 0.0;
        e6_prod_acc<true>(sr, si, wr, wi);
        // 32-bit bookkeeping: blocks count down; the low word of the step index gives row and sign unless it is 0
        uint32_t nblk = (uint32_t)((end - start) >> 2);
        uint32_t lo32 = (uint32_t)start, hi32 = (uint32_t)(start >> 32);
#pragma unroll 1
        for (;;) {
            upd_const_fixed<0, false>(sr, si);
            e6_prod_acc<false>(sr, si, wr, wi);
            upd_const<1>(sr, si, (lo32 & 4u) ? 1.0 : -1.0);
            e6_prod_acc<true>(sr, si, wr, wi);
            upd_const_fixed<0, true>(sr, si);
            e6_prod_acc<false>(sr, si, wr, wi);
            lo32 += 4;
            if (--nblk == 0) break;
            if ((lo32 & 63u) == 0u) { acc_re = dd_add_d(acc_re, wr); acc_im = dd_add_d(acc_im, wi); wr = 0.0; wi = 0.0; }
            int r; double sg;
            if (lo32 & 0x7fffffffu) { r = __ffs((int)lo32) - 1; sg = ((lo32 >> (r + 1)) & 1u) ? 1.0 : -1.0; }
            else {
                if (lo32 == 0u) ++hi32;
                const uint64_t I0 = ((uint64_t)hi32 << 32) | lo32;
                r = ctz64(I0); sg = ((I0 >> (r + 1)) & 1ull) ? 1.0 : -1.0;
            }
            const double2 *row = sA2 + r * N;
#pragma unroll
            for (int j = 0; j < N; ++j) { const double2 a = row[j]; sr[j] = fma(sg, a.x, sr[j]); si[j] = fma(sg, a.y, si[j]); }
            e6_prod_acc<true>(sr, si, wr, wi);
        }
        if (SMEMACC) { acc_re.hi = accs[threadIdx.x]; acc_re.lo = accs[THREADS + threadIdx.x]; acc_im.hi = accs[2 * THREADS + threadIdx.x]; acc_im.lo = accs[3 * THREADS + threadIdx.x]; }
        acc_re = dd_add_d(acc_re, wr); acc_im = dd_add_d(acc_im, wi);
    }
    block_reduce_dd(acc_re, acc_im, red);
    if (threadIdx.x == 0) { double *o = partials + 4 * (size_t)blockIdx.x; o[0] = acc_re.hi; o[1] = acc_re.lo; o[2] = acc_im.hi; o[3] = acc_im.lo; }
}

// ---------------------------------------------------------------------------------------------
// variant F: fully warp-uniform inner structure.  Every thread owns `nwin` whole 64-step windows;
// inside a window the loop counters are block-uniform, so every flipped row (0..5) is addressed
// through the uniform datapath (LDCU from the constant bank into uniform registers that the FP64
// instructions take as operands); only the window-boundary flip (row >= 6, once per 64 steps) uses
// a per-thread shared-memory row.
// ---------------------------------------------------------------------------------------------
template <int NCH, int UNR, int U>
struct FStep {
    static __device__ __forceinline__ void run(double (&sr)[N], double (&si)[N], double &wr, double &wi, double sg_half) {
        if constexpr (U < UNR) {
            constexpr int ROW = cx_ctz(U);
            if constexpr (U == UNR / 2) {
                upd_const<ROW>(sr, si, sg_half);
            } else {
                constexpr bool PLUS = ((U >> (ROW + 1)) & 1) != 0;
                upd_const_fixed<ROW, PLUS>(sr, si);
            }
            cplx p = prod_rr<N, NCH>(sr, si);
            if constexpr (U & 1) { wr -= p.re; wi -= p.im; } else { wr += p.re; wi += p.im; }
            FStep<NCH, UNR, U + 1>::run(sr, si, wr, wi, sg_half);
        }
    }
};

template <int NCH, int UNR, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k1_f(const double *__restrict__ A, uint64_t lo, uint32_t nwin, double *__restrict__ partials) {
    constexpr int LOGU = cx_log2(UNR);
    __shared__ double2 sA2[N * N];
    __shared__ double red[4 * (THREADS / 32)];
    for (int e = threadIdx.x; e < N * N; e += THREADS) {
        double2 v = reinterpret_cast<const double2 *>(A)[e];
        sA2[e] = make_double2(2.0 * v.x, 2.0 * v.y);
    }
    __syncthreads();
    const uint64_t gtid = (uint64_t)blockIdx.x * THREADS + threadIdx.x;
    const uint64_t start = lo + gtid * ((uint64_t)nwin << 6);
    dd acc_re = {0.0, 0.0}, acc_im = {0.0, 0.0};
    double sr[N], si[N];
#pragma unroll
    for (int j = 0; j < N; ++j) { sr[j] = 0.0; si[j] = 0.0; }
    const uint64_t g0 = start ^ (start >> 1);
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
        const double sg = ((g0 >> i) & 1ull) ? -0.5 : 0.5;
        const double2 *row = sA2 + i * N;
#pragma unroll
        for (int j = 0; j < N; ++j) { const double2 a = row[j]; sr[j] = fma(sg, a.x, sr[j]); si[j] = fma(sg, a.y, si[j]); }
    }
    double wr = 0.0, wi = 0.0;
#pragma unroll 1
    for (uint32_t w = 0; w < nwin; ++w) {
        const uint64_t Iw = start + ((uint64_t)w << 6);          // first step of this window (per thread)
        if (w > 0) {
            // window boundary: per-thread row >= 6 from shared memory, fold the window into the dd sum
            acc_re = dd_add_d(acc_re, wr); acc_im = dd_add_d(acc_im, wi); wr = 0.0; wi = 0.0;
            const int r = ctz64(Iw);
            const double sg = ((Iw >> (r + 1)) & 1ull) ? 1.0 : -1.0;
            const double2 *row = sA2 + r * N;
#pragma unroll
            for (int j = 0; j < N; ++j) { const double2 a = row[j]; sr[j] = fma(sg, a.x, sr[j]); si[j] = fma(sg, a.y, si[j]); }
        }
        const double sg_top = ((Iw >> 6) & 1ull) ? 1.0 : -1.0;   // sign of the row-5 flip at step 32 of the window
#pragma unroll 1
        for (uint32_t q = 0; q < 64 / UNR; ++q) {                // block-uniform
            if (q > 0) {
                const uint32_t t = q * UNR;                      // uniform step offset inside the window
                const int r = __ffs((int)t) - 1;                 // uniform row LOGU .. 5
                const double sgu = ((t >> (r + 1)) & 1u) ? 1.0 : -1.0;
                const double sg = (r == 5) ? sg_top : sgu;
                const double2 *row = cA2 + r * N;                // uniform address -> LDCU
#pragma unroll
                for (int j = 0; j < N; ++j) { sr[j] = fma(sg, row[j].x, sr[j]); si[j] = fma(sg, row[j].y, si[j]); }
            }
            cplx p = prod_rr<N, NCH>(sr, si);
            wr += p.re; wi += p.im;
            const double sg_half = (q & 1u) ? 1.0 : -1.0;        // bit LOGU of the step offset
            FStep<NCH, UNR, 1>::run(sr, si, wr, wi, sg_half);
        }
    }
    acc_re = dd_add_d(acc_re, wr); acc_im = dd_add_d(acc_im, wi);
    block_reduce_dd(acc_re, acc_im, red);
    if (threadIdx.x == 0) { double *o = partials + 4 * (size_t)blockIdx.x; o[0] = acc_re.hi; o[1] = acc_re.lo; o[2] = acc_im.hi; o[3] = acc_im.lo; }
}


// variant E2: like E, compile-time rows from shared memory (immediate offsets) instead of the constant bank
template <int ROW>
__device__ __forceinline__ void upd_smem(const double2 *__restrict__ cA2, double (&sr)[N], double (&si)[N], double sg) {
#pragma unroll
    for (int j = 0; j < N; ++j) { sr[j] = fma(sg, cA2[ROW * N + j].x, sr[j]); si[j] = fma(sg, cA2[ROW * N + j].y, si[j]); }
}
template <int ROW, bool PLUS>
__device__ __forceinline__ void upd_smem_fixed(const double2 *__restrict__ cA2, double (&sr)[N], double (&si)[N]) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
        if (PLUS) { sr[j] += cA2[ROW * N + j].x; si[j] += cA2[ROW * N + j].y; }
        else      { sr[j] -= cA2[ROW * N + j].x; si[j] -= cA2[ROW * N + j].y; }
    }
}

template <int NCH, int UNR, int U>
struct E2Step {
    static __device__ __forceinline__ void run(const double2 *__restrict__ rows, double (&sr)[N], double (&si)[N], double &wr, double &wi, double sg_half) {
        if constexpr (U < UNR) {
            constexpr int ROW = cx_ctz(U);
            if constexpr (U == UNR / 2) {
                upd_smem<ROW>(rows, sr, si, sg_half);            // sign = bit log2(UNR) of the block base (run-time)
            } else {
                constexpr bool PLUS = ((U >> (ROW + 1)) & 1) != 0;
                upd_smem_fixed<ROW, PLUS>(rows, sr, si);
            }
            cplx p = prod_rr<N, NCH>(sr, si);
            if constexpr (U & 1) { wr -= p.re; wi -= p.im; } else { wr += p.re; wi += p.im; }
            E2Step<NCH, UNR, U + 1>::run(rows, sr, si, wr, wi, sg_half);
        }
    }
};

template <int NCH, int UNR, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k1_e2(const double *__restrict__ A, uint64_t lo, uint64_t hi, uint64_t span, double *__restrict__ partials) {
    __shared__ double2 sA2[N * N];
    __shared__ double red[4 * (THREADS / 32)];
    for (int e = threadIdx.x; e < N * N; e += THREADS) {
        double2 v = reinterpret_cast<const double2 *>(A)[e];
        sA2[e] = make_double2(2.0 * v.x, 2.0 * v.y);
    }
    __syncthreads();
    const uint64_t gtid = (uint64_t)blockIdx.x * THREADS + threadIdx.x;
    const uint64_t start = lo + gtid * span;     // lo, span: multiples of 64
    dd acc_re = {0.0, 0.0}, acc_im = {0.0, 0.0};
    if (start < hi) {
        const uint64_t end = (hi - start < span) ? hi : start + span;   // multiple of UNR
        double sr[N], si[N];
#pragma unroll
        for (int j = 0; j < N; ++j) { sr[j] = 0.0; si[j] = 0.0; }
        const uint64_t g0 = start ^ (start >> 1);
#pragma unroll 1
        for (int i = 0; i < N; ++i) {
            const double sg = ((g0 >> i) & 1ull) ? -0.5 : 0.5;
            const double2 *row = sA2 + i * N;
#pragma unroll
            for (int j = 0; j < N; ++j) { const double2 a = row[j]; sr[j] = fma(sg, a.x, sr[j]); si[j] = fma(sg, a.y, si[j]); }
        }
        cplx p = prod_rr<N, NCH>(sr, si);
        double wr = p.re, wi = p.im;             // term `start` (even index: +)
#pragma unroll 1
        for (uint64_t I0 = start;;) {
            const double sg_half = ((I0 >> cx_log2(UNR)) & 1ull) ? 1.0 : -1.0;
            E2Step<NCH, UNR, 1>::run(sA2, sr, si, wr, wi, sg_half);
            I0 += UNR;
            if (I0 >= end) break;
            if (((uint32_t)I0 & 63u) == 0u) { acc_re = dd_add_d(acc_re, wr); acc_im = dd_add_d(acc_im, wi); wr = 0.0; wi = 0.0; }
            const int r = ctz64(I0);
            const double sg = ((I0 >> (r + 1)) & 1ull) ? 1.0 : -1.0;
            const double2 *row = sA2 + r * N;
#pragma unroll
            for (int j = 0; j < N; ++j) { const double2 a = row[j]; sr[j] = fma(sg, a.x, sr[j]); si[j] = fma(sg, a.y, si[j]); }
            p = prod_rr<N, NCH>(sr, si);
            wr += p.re; wi += p.im;
        }
        acc_re = dd_add_d(acc_re, wr); acc_im = dd_add_d(acc_im, wi);
    }
    block_reduce_dd(acc_re, acc_im, red);
    if (threadIdx.x == 0) { double *o = partials + 4 * (size_t)blockIdx.x; o[0] = acc_re.hi; o[1] = acc_re.lo; o[2] = acc_im.hi; o[3] = acc_im.lo; }
}


// ---------------------------------------------------------------------------------------------
// variant G: column-major sweep over an aligned block of UNR = 2^LOGU Gray steps.
// For every column j the UNR successive factor values f_1..f_UNR are produced by the same chain of
// additions the reference performs (f_k = f_{k-1} +- 2A[ctz(k)][j]), each feeding its own running
// product P_k.  Rows 0..LOGU-1 are compile-time (constant bank, one load per column per block), the
// block-closing flip uses a run-time row from shared memory: (LOGU+1)*N loads per UNR terms.
// ---------------------------------------------------------------------------------------------
template <int UNR, int K>
struct GCol {
    // advances factor value f (column j) through steps K..UNR and multiplies it into P[K-1..UNR-1]
    template <int J>
    static __device__ __forceinline__ void run(double &fr, double &fi, cplx (&P)[UNR], const double2 (&rowv)[cx_log2(UNR) + 1],
                                               double sg_half, double sg_last) {
        if constexpr (K <= UNR) {
            constexpr int LOGU = cx_log2(UNR);
            if constexpr (K == UNR) {                       // block-closing flip: run-time row, per-thread sign
                fr = fma(sg_last, rowv[LOGU].x, fr); fi = fma(sg_last, rowv[LOGU].y, fi);
            } else {
                constexpr int ROW = cx_ctz(K);
                if constexpr (K == UNR / 2) {               // sign = bit LOGU of the block base
                    fr = fma(sg_half, rowv[ROW].x, fr); fi = fma(sg_half, rowv[ROW].y, fi);
                } else if constexpr (((K >> (ROW + 1)) & 1) != 0) {
                    fr += rowv[ROW].x; fi += rowv[ROW].y;
                } else {
                    fr -= rowv[ROW].x; fi -= rowv[ROW].y;
                }
            }
            if constexpr (J == 0) { P[K - 1].re = fr; P[K - 1].im = fi; }
            else { cplx f = {fr, fi}; P[K - 1] = cmul(P[K - 1], f); }
            GCol<UNR, K + 1>::template run<J>(fr, fi, P, rowv, sg_half, sg_last);
        }
    }
};

template <int UNR, int J>
struct GSweep {
    // software-pipelined: the rows of column J+1 are fetched before column J is processed, and a
    // compiler memory barrier per column keeps ptxas from hoisting all loads to the top (which
    // blows the register budget).
    static __device__ __forceinline__ void run(double (&sr)[N], double (&si)[N], cplx (&P)[UNR], const double2 *__restrict__ rowlast,
                                               const double2 *__restrict__ rows, const double2 (&cur)[cx_log2(UNR) + 1],
                                               double sg_half, double sg_last) {
        if constexpr (J < N) {
            constexpr int LOGU = cx_log2(UNR);
            double2 nxt[LOGU + 1];
            if constexpr (J + 1 < N) {
#pragma unroll
                for (int r = 0; r < LOGU; ++r) nxt[r] = rows[r * N + J + 1];
                nxt[LOGU] = rowlast[J + 1];
            }
            asm volatile("" ::: "memory");
            GCol<UNR, 1>::template run<J>(sr[J], si[J], P, cur, sg_half, sg_last);
            GSweep<UNR, J + 1>::run(sr, si, P, rowlast, rows, nxt, sg_half, sg_last);
        }
    }
};

template <int UNR, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k1_g(const double *__restrict__ A, uint64_t lo, uint64_t hi, uint64_t span, double *__restrict__ partials) {
    constexpr int LOGU = cx_log2(UNR);
    __shared__ double2 sA2[N * N];
    __shared__ double red[4 * (THREADS / 32)];
    for (int e = threadIdx.x; e < N * N; e += THREADS) {
        double2 v = reinterpret_cast<const double2 *>(A)[e];
        sA2[e] = make_double2(2.0 * v.x, 2.0 * v.y);
    }
    __syncthreads();
    const uint64_t gtid = (uint64_t)blockIdx.x * THREADS + threadIdx.x;
    const uint64_t start = lo + gtid * span;     // lo, span, hi: multiples of 64
    dd acc_re = {0.0, 0.0}, acc_im = {0.0, 0.0};
    if (start < hi) {
        const uint64_t end = (hi - start < span) ? hi : start + span;
        double sr[N], si[N];
#pragma unroll
        for (int j = 0; j < N; ++j) { sr[j] = 0.0; si[j] = 0.0; }
        const uint64_t g0 = start ^ (start >> 1);
#pragma unroll 1
        for (int i = 0; i < N; ++i) {
            const double sg = ((g0 >> i) & 1ull) ? -0.5 : 0.5;
            const double2 *row = sA2 + i * N;
#pragma unroll
            for (int j = 0; j < N; ++j) { const double2 a = row[j]; sr[j] = fma(sg, a.x, sr[j]); si[j] = fma(sg, a.y, si[j]); }
        }
        cplx p0 = prod_rr<N, 4>(sr, si);
        double wr = p0.re, wi = p0.im;             // term `start` (even index: +)
#pragma unroll 1
        for (uint64_t I0 = start; I0 < end; I0 += UNR) {
            // steps I0+1 .. I0+UNR; the last one (I0+UNR) belongs to this thread only if it is < end
            const uint64_t In = I0 + UNR;
            const bool last_block = (In >= end);
            const int r = last_block ? LOGU : ctz64(In);
            const double sg_last = last_block ? 0.0 : (((In >> (r + 1)) & 1ull) ? 1.0 : -1.0);
            const double sg_half = ((I0 >> LOGU) & 1ull) ? 1.0 : -1.0;
            cplx P[UNR];
            double2 first[LOGU + 1];
#pragma unroll
            for (int q = 0; q < LOGU; ++q) first[q] = sA2[q * N];
            first[LOGU] = sA2[r * N];
            GSweep<UNR, 0>::run(sr, si, P, sA2 + r * N, sA2, first, sg_half, sg_last);
            double br = 0.0, bi = 0.0;
#pragma unroll
            for (int k = 1; k < UNR; ++k) { if (k & 1) { br -= P[k - 1].re; bi -= P[k - 1].im; } else { br += P[k - 1].re; bi += P[k - 1].im; } }
            if (!last_block) { br += P[UNR - 1].re; bi += P[UNR - 1].im; }
            wr += br; wi += bi;
            if (((uint32_t)In & 63u) == 0u) { acc_re = dd_add_d(acc_re, wr); acc_im = dd_add_d(acc_im, wi); wr = 0.0; wi = 0.0; }
        }
        acc_re = dd_add_d(acc_re, wr); acc_im = dd_add_d(acc_im, wi);
    }
    block_reduce_dd(acc_re, acc_im, red);
    if (threadIdx.x == 0) { double *o = partials + 4 * (size_t)blockIdx.x; o[0] = acc_re.hi; o[1] = acc_re.lo; o[2] = acc_im.hi; o[3] = acc_im.lo; }
}


// ---------------------------------------------------------------------------------------------
// variant H: running column sums in SHARED memory, 2^LOGU product chains in registers, dynamic
// loop over the columns (N is a run-time value).  Per block of UNR = 2^LOGU Gray steps and per
// column: 1 LDS (sum) + LOGU LDS (compile-time rows, warp-uniform address) + 1 LDS (run-time row)
// + 1 STS (sum) feed 6*UNR FP64 instructions.
// ---------------------------------------------------------------------------------------------
template <int UNR, int K, bool FIRST>
struct HCol {
    static __device__ __forceinline__ void run(double &fr, double &fi, cplx (&P)[UNR], const double2 (&rowv)[cx_log2(UNR) + 1],
                                               double sg_half, double sg_last) {
        if constexpr (K <= UNR) {
            constexpr int LOGU = cx_log2(UNR);
            if constexpr (K == UNR) {
                fr = fma(sg_last, rowv[LOGU].x, fr); fi = fma(sg_last, rowv[LOGU].y, fi);
            } else {
                constexpr int ROW = cx_ctz(K);
                if constexpr (K == UNR / 2) {
                    fr = fma(sg_half, rowv[ROW].x, fr); fi = fma(sg_half, rowv[ROW].y, fi);
                } else if constexpr (((K >> (ROW + 1)) & 1) != 0) {
                    fr += rowv[ROW].x; fi += rowv[ROW].y;
                } else {
                    fr -= rowv[ROW].x; fi -= rowv[ROW].y;
                }
            }
            if constexpr (FIRST) { P[K - 1].re = fr; P[K - 1].im = fi; }
            else { cplx f = {fr, fi}; P[K - 1] = cmul(P[K - 1], f); }
            HCol<UNR, K + 1, FIRST>::run(fr, fi, P, rowv, sg_half, sg_last);
        }
    }
};

template <int UNR, int THREADS, int CUNR>
__global__ void __launch_bounds__(THREADS, 1)
k1_h(const double *__restrict__ A, int n, uint64_t lo, uint64_t hi, uint64_t span, double *__restrict__ partials) {
    constexpr int LOGU = cx_log2(UNR);
    extern __shared__ __align__(16) unsigned char hsm[];
    double2 *sA2 = reinterpret_cast<double2 *>(hsm);                 // [n][n]   2*A
    double2 *sums = sA2 + n * n;                                     // [n][THREADS]
    __shared__ double red[4 * (THREADS / 32)];
    for (int e = threadIdx.x; e < n * n; e += THREADS) {
        double2 v = reinterpret_cast<const double2 *>(A)[e];
        sA2[e] = make_double2(2.0 * v.x, 2.0 * v.y);
    }
    __syncthreads();
    const uint64_t gtid = (uint64_t)blockIdx.x * THREADS + threadIdx.x;
    const uint64_t start = lo + gtid * span;     // lo, span, hi: multiples of 64
    double2 *my = sums + threadIdx.x;
    dd acc_re = {0.0, 0.0}, acc_im = {0.0, 0.0};
    if (start < hi) {
        const uint64_t end = (hi - start < span) ? hi : start + span;
        const uint64_t g0 = start ^ (start >> 1);
        cplx p0 = {1.0, 0.0};
#pragma unroll 1
        for (int j = 0; j < n; ++j) {
            double xr = 0.0, xi = 0.0;
#pragma unroll 2
            for (int i = 0; i < n; ++i) {
                const double sg = ((g0 >> i) & 1ull) ? -0.5 : 0.5;
                const double2 a = sA2[i * n + j];
                xr = fma(sg, a.x, xr); xi = fma(sg, a.y, xi);
            }
            my[j * THREADS] = make_double2(xr, xi);
            cplx x = {xr, xi};
            p0 = cmul(p0, x);
        }
        double wr = p0.re, wi = p0.im;             // term `start` (even index: +)
#pragma unroll 1
        for (uint64_t I0 = start; I0 < end; I0 += UNR) {
            const uint64_t In = I0 + UNR;
            const bool last_block = (In >= end);
            const int r = last_block ? LOGU : ctz64(In);
            const double sg_last = last_block ? 0.0 : (((In >> (r + 1)) & 1ull) ? 1.0 : -1.0);
            const double sg_half = ((I0 >> LOGU) & 1ull) ? 1.0 : -1.0;
            const double2 *rowlast = sA2 + r * n;
            cplx P[UNR];
            {   // column 0 initialises the chains
                double2 rowv[LOGU + 1];
#pragma unroll
                for (int q = 0; q < LOGU; ++q) rowv[q] = sA2[q * n];
                rowv[LOGU] = rowlast[0];
                double2 sv = my[0];
                HCol<UNR, 1, true>::run(sv.x, sv.y, P, rowv, sg_half, sg_last);
                my[0] = sv;
            }
#pragma unroll CUNR
            for (int j = 1; j < n; ++j) {
                double2 rowv[LOGU + 1];
#pragma unroll
                for (int q = 0; q < LOGU; ++q) rowv[q] = sA2[q * n + j];
                rowv[LOGU] = rowlast[j];
                double2 sv = my[j * THREADS];
                HCol<UNR, 1, false>::run(sv.x, sv.y, P, rowv, sg_half, sg_last);
                my[j * THREADS] = sv;
            }
            double br = 0.0, bi = 0.0;
#pragma unroll
            for (int k = 1; k < UNR; ++k) { if (k & 1) { br -= P[k - 1].re; bi -= P[k - 1].im; } else { br += P[k - 1].re; bi += P[k - 1].im; } }
            if (!last_block) { br += P[UNR - 1].re; bi += P[UNR - 1].im; }
            wr += br; wi += bi;
            if (((uint32_t)In & 63u) == 0u) { acc_re = dd_add_d(acc_re, wr); acc_im = dd_add_d(acc_im, wi); wr = 0.0; wi = 0.0; }
        }
        acc_re = dd_add_d(acc_re, wr); acc_im = dd_add_d(acc_im, wi);
    }
    block_reduce_dd(acc_re, acc_im, red);
    if (threadIdx.x == 0) { double *o = partials + 4 * (size_t)blockIdx.x; o[0] = acc_re.hi; o[1] = acc_re.lo; o[2] = acc_im.hi; o[3] = acc_im.lo; }
}

// ---------------------------------------------------------------------------------------------
// FP64 peak probes
// ---------------------------------------------------------------------------------------------
template <int CH, int UNR>
__global__ void __launch_bounds__(256) peak_fma(int iters, double *sink) {
    double a[CH];
    const double x = 1.0 + 1e-9 * threadIdx.x, y = 1e-12 * (blockIdx.x + 1);
#pragma unroll
    for (int k = 0; k < CH; ++k) a[k] = k * 1e-3;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < UNR; ++u)
#pragma unroll
            for (int k = 0; k < CH; ++k) a[k] = fma(a[k], x, y);
    }
    double r = 0.0;
#pragma unroll
    for (int k = 0; k < CH; ++k) r += a[k];
    if (r == 123.456) sink[threadIdx.x] = r;
}


// ---------------------------------------------------------------------------------------------
// issue-mix probes: 64 independent DFMAs per iteration (16 chains x 4) + NLDS LDS.128 + NINT integer ops
// ---------------------------------------------------------------------------------------------
template <int NLDS, int NINT, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) mix_probe(int iters, double *sink) {
    __shared__ double2 sm[1024];
    for (int e = threadIdx.x; e < 1024; e += THREADS) sm[e] = make_double2(1e-12 * e, 1e-13 * e);
    __syncthreads();
    double a[16];
    const double x = 1.0 + 1e-9 * threadIdx.x;
    double2 ld[NLDS > 0 ? NLDS : 1];
#pragma unroll
    for (int k = 0; k < (NLDS > 0 ? NLDS : 1); ++k) ld[k] = make_double2(1e-12, 1e-12);
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = k * 1e-3;
    unsigned idx = blockIdx.x & 7, acc = threadIdx.x;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        const double2 *row = sm + (idx & 31) * 32;     // warp-uniform address, like a matrix row
#pragma unroll
        for (int k = 0; k < NLDS; ++k) ld[k] = row[k];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int q = (u * 16 + k);
                const double y = (NLDS > 0) ? ((q & 1) ? ld[(q >> 1) % (NLDS > 0 ? NLDS : 1)].y : ld[(q >> 1) % (NLDS > 0 ? NLDS : 1)].x) : 1e-12;
                a[k] = fma(a[k], x, y);
            }
#pragma unroll
        for (int k = 0; k < NINT; ++k) acc = acc * 1664525u + idx + k;
        idx += 1 + (acc >> 31);
    }
    double r = 0.0;
#pragma unroll
    for (int k = 0; k < 16; ++k) r += a[k];
    if (r == 123.456 || acc == 77u) sink[threadIdx.x] = r;
}

// ---------------------------------------------------------------------------------------------
typedef void (*kfn)(const double *, uint64_t, uint64_t, uint64_t, double *);
struct Variant { const char *name; kfn fn; int threads, minb, lanes_per_stream; };

static double dd_total(const std::vector<double> &p, int nblocks, int q) {
    long double s = 0;
    for (int b = 0; b < nblocks; ++b) s += (long double)p[4 * b + q] + (long double)p[4 * b + q + 1];
    return (double)s;
}

int main(int argc, char **argv) {
    if (argc < 4) { printf("usage: %s matrix.bin re im\n", argv[0]); return 1; }
    std::vector<double> A(2 * N * N);
    FILE *f = fopen(argv[1], "rb");
    if (!f || fread(A.data(), sizeof(double), A.size(), f) != A.size()) { printf("cannot read %s\n", argv[1]); return 1; }
    fclose(f);
    const double want_re = atof(argv[2]), want_im = atof(argv[3]);
    const char *filter = argc > 4 ? argv[4] : nullptr;
    auto skip = [&](const char *name) { return filter && !strstr(name, filter); };
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs\n", prop.name, sms);
    double *dA, *dP, *dSink;
    CK(cudaMalloc(&dA, A.size() * sizeof(double)));
    CK(cudaMalloc(&dP, sizeof(double) * 4 * 65536));
    CK(cudaMalloc(&dSink, sizeof(double) * 1024));
    CK(cudaMemcpy(dA, A.data(), A.size() * sizeof(double), cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));

    // ---- FP64 peak probes
    {
        struct P { const char *name; void (*fn)(int, double *); int ch, unr, blocks_per_sm; } probes[] = {
            {"fma 16ch x1, 8 blk/SM", peak_fma<16, 1>, 16, 1, 8}, {"fma 16ch x4, 8 blk/SM", peak_fma<16, 4>, 16, 4, 8},
            {"fma 8ch x8, 8 blk/SM", peak_fma<8, 8>, 8, 8, 8},    {"fma 8ch x8, 4 blk/SM", peak_fma<8, 8>, 8, 8, 4},
            {"fma 4ch x16, 8 blk/SM", peak_fma<4, 16>, 4, 16, 8}, {"fma 32ch x2, 4 blk/SM", peak_fma<32, 2>, 32, 2, 4},
            {"fma 8ch x8, 2 blk/SM", peak_fma<8, 8>, 8, 8, 2},    {"fma 8ch x8, 1 blk/SM", peak_fma<8, 8>, 8, 8, 1},
        };
        for (auto &p : probes) {
            if (skip(p.name)) continue;
            const int iters = 200000 / p.unr;
            p.fn<<<sms * p.blocks_per_sm, 256>>>(iters / 10, dSink);
            CK(cudaEventRecord(e0));
            p.fn<<<sms * p.blocks_per_sm, 256>>>(iters, dSink);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            const double flops = 2.0 * sms * p.blocks_per_sm * 256.0 * p.ch * p.unr * (double)iters;
            printf("PEAK %-26s %8.3f ms  %7.2f TFLOP/s\n", p.name, ms, flops / (ms * 1e-3) / 1e12);
        }
    }

    {
        struct M { const char *name; void (*fn)(int, double *); int threads; } mixes[] = {
            {"mix 64 DFMA            384thr", mix_probe<0, 0, 384>, 384},  {"mix 64 DFMA            1024thr", mix_probe<0, 0, 1024>, 1024},
            {"mix 64 DFMA +  5 LDS   384thr", mix_probe<5, 0, 384>, 384},  {"mix 64 DFMA +  5 LDS   1024thr", mix_probe<5, 0, 1024>, 1024},
            {"mix 64 DFMA + 11 LDS   384thr", mix_probe<11, 0, 384>, 384}, {"mix 64 DFMA + 11 LDS   1024thr", mix_probe<11, 0, 1024>, 1024},
            {"mix 64 DFMA + 22 LDS   384thr", mix_probe<22, 0, 384>, 384}, {"mix 64 DFMA + 22 LDS   1024thr", mix_probe<22, 0, 1024>, 1024},
            {"mix 64 DFMA + 11 INT   384thr", mix_probe<0, 11, 384>, 384}, {"mix 64 DFMA + 11 INT   1024thr", mix_probe<0, 11, 1024>, 1024},
            {"mix 64 DFMA + 22 INT   384thr", mix_probe<0, 22, 384>, 384}, {"mix 64 DFMA + 22 INT   1024thr", mix_probe<0, 22, 1024>, 1024},
            {"mix 64 DFMA + 11L+11I  384thr", mix_probe<11, 11, 384>, 384}, {"mix 64 DFMA + 11L+11I  1024thr", mix_probe<11, 11, 1024>, 1024},
            {"mix 64 DFMA            256thr", mix_probe<0, 0, 256>, 256},   {"mix 64 DFMA + 11L+11I  256thr", mix_probe<11, 11, 256>, 256},
            {"mix 64 DFMA            128thr", mix_probe<0, 0, 128>, 128},   {"mix 64 DFMA + 11L+11I  128thr", mix_probe<11, 11, 128>, 128},
        };
        for (auto &p : mixes) {
            if (skip(p.name)) continue;
            const int iters = 60000;
            p.fn<<<sms, p.threads>>>(iters / 10, dSink);
            CK(cudaEventRecord(e0));
            p.fn<<<sms, p.threads>>>(iters, dSink);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            const double flops = 2.0 * sms * p.threads * 64.0 * (double)iters;
            printf("MIX  %-34s %8.3f ms  %7.2f TFLOP/s\n", p.name, ms, flops / (ms * 1e-3) / 1e12);
        }
    }

    Variant vars[] = {
        {"A nch3 128x3", k1_a<3, 128, 3>, 128, 3, 1}, {"A nch4 128x3", k1_a<4, 128, 3>, 128, 3, 1},
        {"A nch5 128x3", k1_a<5, 128, 3>, 128, 3, 1}, {"A nch6 128x3", k1_a<6, 128, 3>, 128, 3, 1},
        {"A nch4 128x2", k1_a<4, 128, 2>, 128, 2, 1}, {"A nch6 128x2", k1_a<6, 128, 2>, 128, 2, 1},
        {"A nch4 64x6", k1_a<4, 64, 6>, 64, 6, 1},    {"A nch4 96x4", k1_a<4, 96, 4>, 96, 4, 1},
        {"A nch4 192x2", k1_a<4, 192, 2>, 192, 2, 1}, {"A nch4 384x1", k1_a<4, 384, 1>, 384, 1, 1},
        {"C nch2 128x3", k1_c<2, 128, 3>, 128, 3, 1}, {"C nch3 128x3", k1_c<3, 128, 3>, 128, 3, 1},
        {"C nch2 128x2", k1_c<2, 128, 2>, 128, 2, 1}, {"C nch3 128x2", k1_c<3, 128, 2>, 128, 2, 1},
        {"B nch2 128x4", k1_b<2, 128, 4, false>, 128, 4, 2}, {"B nch3 128x4", k1_b<3, 128, 4, false>, 128, 4, 2},
        {"B nch3 128x5", k1_b<3, 128, 5, false>, 128, 5, 2}, {"B nch4 128x4", k1_b<4, 128, 4, false>, 128, 4, 2},
        {"B nch3 128x6", k1_b<3, 128, 6, false>, 128, 6, 2}, {"B nch3 256x2", k1_b<3, 256, 2, false>, 256, 2, 2},
        {"D nch2 128x4", k1_b<2, 128, 4, true>, 128, 4, 2},  {"D nch3 128x4", k1_b<3, 128, 4, true>, 128, 4, 2},
        {"D nch2 128x5", k1_b<2, 128, 5, true>, 128, 5, 2},  {"D nch2 128x3", k1_b<2, 128, 3, true>, 128, 3, 2},
        {"D nch3 128x3", k1_b<3, 128, 3, true>, 128, 3, 2},
        {"E nch4 u2 384x1", k1_e<4, 2, 384, 1>, 384, 1, 1}, {"E nch4 u4 384x1", k1_e<4, 4, 384, 1>, 384, 1, 1},
        {"E nch4 u8 384x1", k1_e<4, 8, 384, 1>, 384, 1, 1}, {"E nch4 u16 384x1", k1_e<4, 16, 384, 1>, 384, 1, 1},
        {"E nch3 u4 384x1", k1_e<3, 4, 384, 1>, 384, 1, 1}, {"E nch5 u4 384x1", k1_e<5, 4, 384, 1>, 384, 1, 1},
        {"E nch4 u4 128x3", k1_e<4, 4, 128, 3>, 128, 3, 1}, {"E nch4 u8 128x3", k1_e<4, 8, 128, 3>, 128, 3, 1},
        {"E nch4 u4 192x2", k1_e<4, 4, 192, 2>, 192, 2, 1}, {"E nch4 u4 256x1", k1_e<4, 4, 256, 1>, 256, 1, 1},
        {"E nch6 u4 256x1", k1_e<6, 4, 256, 1>, 256, 1, 1}, {"E nch4 u8 256x1", k1_e<4, 8, 256, 1>, 256, 1, 1},
        {"E nch4 u4 288x1", k1_e<4, 4, 288, 1>, 288, 1, 1}, {"E nch4 u4 320x1", k1_e<4, 4, 320, 1>, 320, 1, 1},
        {"E nch3 u4 320x1", k1_e<3, 4, 320, 1>, 320, 1, 1}, {"E nch2 u4 320x1", k1_e<2, 4, 320, 1>, 320, 1, 1},
        {"E nch4 u4 352x1", k1_e<4, 4, 352, 1>, 352, 1, 1}, {"E nch3 u4 352x1", k1_e<3, 4, 352, 1>, 352, 1, 1},
        {"E nch3 u4 256x1", k1_e<3, 4, 256, 1>, 256, 1, 1}, {"E nch2 u4 256x1", k1_e<2, 4, 256, 1>, 256, 1, 1},
        {"E nch4 u8 288x1", k1_e<4, 8, 288, 1>, 288, 1, 1}, {"E nch4 u2 256x1", k1_e<4, 2, 256, 1>, 256, 1, 1},
        {"E nch4 u4 224x1", k1_e<4, 4, 224, 1>, 224, 1, 1}, {"E nch8 u4 256x1", k1_e<8, 4, 256, 1>, 256, 1, 1},
        {"E2 nch4 u4 256x1", k1_e2<4, 4, 256, 1>, 256, 1, 1}, {"E2 nch4 u8 256x1", k1_e2<4, 8, 256, 1>, 256, 1, 1},
        {"E2 nch3 u4 256x1", k1_e2<3, 4, 256, 1>, 256, 1, 1}, {"E2 nch4 u4 384x1", k1_e2<4, 4, 384, 1>, 384, 1, 1},
        {"E2 nch4 u2 256x1", k1_e2<4, 2, 256, 1>, 256, 1, 1}, {"E2 nch4 u2 384x1", k1_e2<4, 2, 384, 1>, 384, 1, 1},
        {"E3 nch2 u4 256x1", k1_e3<2, 4, 256, 1>, 256, 1, 1}, {"E3 nch4 u4 256x1", k1_e3<4, 4, 256, 1>, 256, 1, 1},
        {"E3 nch2 u8 256x1", k1_e3<2, 8, 256, 1>, 256, 1, 1}, {"E3 nch3 u4 256x1", k1_e3<3, 4, 256, 1>, 256, 1, 1},
        {"E4 nch2 256x1", k1_e4<2, 256, 1>, 256, 1, 1}, {"E4 nch4 256x1", k1_e4<4, 256, 1>, 256, 1, 1},
        {"E4 nch3 256x1", k1_e4<3, 256, 1>, 256, 1, 1}, {"E4 nch2 384x1", k1_e4<2, 384, 1>, 384, 1, 1},
        {"E5 nch2 256x1", k1_e5<2, 256, 1>, 256, 1, 1}, {"E5 nch4 256x1", k1_e5<4, 256, 1>, 256, 1, 1},
        {"E5 nch3 256x1", k1_e5<3, 256, 1>, 256, 1, 1},
        {"E nch1 u4 384x1", k1_e<1, 4, 384, 1>, 384, 1, 1}, {"E nch2 u4 384x1", k1_e<2, 4, 384, 1>, 384, 1, 1},
        {"E nch1 u2 384x1", k1_e<1, 2, 384, 1>, 384, 1, 1}, {"E nch2 u2 384x1", k1_e<2, 2, 384, 1>, 384, 1, 1},
        {"E nch1 u4 256x1", k1_e<1, 4, 256, 1>, 256, 1, 1},
        {"E6 384 regacc", k1_e6<384, false>, 384, 1, 1}, {"E6 384 smemacc", k1_e6<384, true>, 384, 1, 1},
        {"E6 256 regacc", k1_e6<256, false>, 256, 1, 1}, {"E6 512 smemacc", k1_e6<512, true>, 512, 1, 1},
        {"E7 384 regacc", k1_e7<384, false>, 384, 1, 1},
        {"G u4 256x1", k1_g<4, 256, 1>, 256, 1, 1},  {"G u8 256x1", k1_g<8, 256, 1>, 256, 1, 1},  {"G u16 256x1", k1_g<16, 256, 1>, 256, 1, 1},
        {"G u2 256x1", k1_g<2, 256, 1>, 256, 1, 1},  {"G u4 384x1", k1_g<4, 384, 1>, 384, 1, 1},  {"G u8 384x1", k1_g<8, 384, 1>, 384, 1, 1},
        {"G u8 128x2", k1_g<8, 128, 2>, 128, 2, 1},  {"G u8 128x3", k1_g<8, 128, 3>, 128, 3, 1},  {"G u4 128x3", k1_g<4, 128, 3>, 128, 3, 1},
        {"G u8 320x1", k1_g<8, 320, 1>, 320, 1, 1},  {"G u8 288x1", k1_g<8, 288, 1>, 288, 1, 1},  {"G u16 288x1", k1_g<16, 288, 1>, 288, 1, 1},
    };
    {
        std::vector<double> A2(A);
        for (auto &x : A2) x *= 2.0;
        CK(cudaMemcpyToSymbol(cA2, A2.data(), sizeof(double) * 2 * N * N));
    }
    const uint64_t total = 1ull << (N - 1);
    for (auto &v : vars) {
        if (skip(v.name)) continue;
        cudaFuncAttributes attr;
        CK(cudaFuncGetAttributes(&attr, (const void *)v.fn));
        int occ_blocks = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_blocks, (const void *)v.fn, v.threads, 0));
        const uint64_t streams_cap = (uint64_t)sms * occ_blocks * v.threads / v.lanes_per_stream;
        uint64_t span = (total + streams_cap - 1) / streams_cap;
        span = ((span + 63) / 64) * 64;
        const uint64_t nstreams = (total + span - 1) / span;
        const int grid = (int)((nstreams * v.lanes_per_stream + v.threads - 1) / v.threads);
        std::vector<double> hp(4 * (size_t)grid);
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaEventRecord(e0));
            v.fn<<<grid, v.threads>>>(dA, 0, total, span, dP);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
        }
        CK(cudaMemcpy(hp.data(), dP, sizeof(double) * 4 * grid, cudaMemcpyDeviceToHost));
        const double re = dd_total(hp, grid, 0) / (double)total, im = dd_total(hp, grid, 2) / (double)total;
        const double rel = hypot(re - want_re, im - want_im) / hypot(want_re, want_im);
        const double tf = (8.0 * N - 4) * (double)total / (best * 1e-3) / 1e12;
        printf("K1 %-14s regs %3d occ %d blk/SM grid %5d span %6llu  %7.3f ms  %6.2f TF useful  %6.1f perm/s  rel.err %.2e\n", v.name,
               attr.numRegs, occ_blocks, grid, (unsigned long long)span, best, tf, 1e3 / best, rel);
    }

    // ---- variant F: uniform windows + generic tail
    {
        typedef void (*ffn)(const double *, uint64_t, uint32_t, double *);
        struct FV { const char *name; ffn fn; int threads, minb; } fv[] = {
            {"F nch4 u4 256x1", k1_f<4, 4, 256, 1>, 256, 1}, {"F nch4 u2 256x1", k1_f<4, 2, 256, 1>, 256, 1},
            {"F nch4 u8 256x1", k1_f<4, 8, 256, 1>, 256, 1}, {"F nch3 u4 256x1", k1_f<3, 4, 256, 1>, 256, 1},
            {"F nch5 u4 256x1", k1_f<5, 4, 256, 1>, 256, 1}, {"F nch6 u4 256x1", k1_f<6, 4, 256, 1>, 256, 1},
            {"F nch4 u4 384x1", k1_f<4, 4, 384, 1>, 384, 1}, {"F nch4 u8 384x1", k1_f<4, 8, 384, 1>, 384, 1},
            {"F nch4 u4 128x2", k1_f<4, 4, 128, 2>, 128, 2}, {"F nch4 u4 128x3", k1_f<4, 4, 128, 3>, 128, 3},
            {"F nch4 u4 320x1", k1_f<4, 4, 320, 1>, 320, 1}, {"F nch4 u4 288x1", k1_f<4, 4, 288, 1>, 288, 1},
            {"F nch4 u16 256x1", k1_f<4, 16, 256, 1>, 256, 1},
        };
        const uint64_t total_f = 1ull << (N - 1), windows = total_f >> 6;
        for (auto &v : fv) {
            if (!filter || skip(v.name)) continue;
            cudaFuncAttributes attr;
            CK(cudaFuncGetAttributes(&attr, (const void *)v.fn));
            int occ_blocks = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_blocks, (const void *)v.fn, v.threads, 0));
            const int grid = sms * occ_blocks;
            const uint64_t nthreads = (uint64_t)grid * v.threads;
            const uint32_t nwin = (uint32_t)(windows / nthreads);
            const uint64_t main_terms = (uint64_t)nwin * 64 * nthreads;
            const uint64_t tail = total_f - main_terms;
            const int tail_grid = (int)((tail / 64 + 127) / 128);
            std::vector<double> hp(4 * (size_t)(grid + tail_grid));
            float best = 1e30f;
            for (int rep = 0; rep < 4; ++rep) {
                CK(cudaEventRecord(e0));
                v.fn<<<grid, v.threads>>>(dA, 0, nwin, dP);
                if (tail) k1_a<4, 128, 3><<<tail_grid, 128>>>(dA, main_terms, total_f, 64, dP + 4 * grid);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                CK(cudaGetLastError());
                float ms;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                if (rep > 0 && ms < best) best = ms;
            }
            CK(cudaMemcpy(hp.data(), dP, sizeof(double) * 4 * (grid + tail_grid), cudaMemcpyDeviceToHost));
            const double re = dd_total(hp, grid + tail_grid, 0) / (double)total_f, im = dd_total(hp, grid + tail_grid, 2) / (double)total_f;
            const double rel = hypot(re - want_re, im - want_im) / hypot(want_re, want_im);
            const double tf = (8.0 * N - 4) * (double)total_f / (best * 1e-3) / 1e12;
            printf("K1 %-16s regs %3d spill? occ %d grid %4d nwin %4u tail %7llu  %7.3f ms  %6.2f TF useful  %6.1f perm/s  rel.err %.2e\n", v.name,
                   attr.numRegs, occ_blocks, grid, nwin, (unsigned long long)tail, best, tf, 1e3 / best, rel);
        }
    }

    // ---- variant H: sums in shared memory
    {
        typedef void (*hfn)(const double *, int, uint64_t, uint64_t, uint64_t, double *);
        struct HV { const char *name; hfn fn; int threads; } hv[] = {
            {"H u8  c1 384", k1_h<8, 384, 1>, 384},   {"H u16 c1 384", k1_h<16, 384, 1>, 384}, {"H u32 c1 384", k1_h<32, 384, 1>, 384},
            {"H u16 c2 384", k1_h<16, 384, 2>, 384},  {"H u32 c2 384", k1_h<32, 384, 2>, 384}, {"H u64 c1 384", k1_h<64, 384, 1>, 384},
            {"H u16 c1 256", k1_h<16, 256, 1>, 256},  {"H u32 c1 256", k1_h<32, 256, 1>, 256}, {"H u32 c2 256", k1_h<32, 256, 2>, 256},
            {"H u16 c1 448", k1_h<16, 448, 1>, 448},  {"H u32 c1 320", k1_h<32, 320, 1>, 320}, {"H u16 c3 384", k1_h<16, 384, 3>, 384},
            {"H u32 c3 256", k1_h<32, 256, 3>, 256},  {"H u32 c4 256", k1_h<32, 256, 4>, 256}, {"H u32 c5 256", k1_h<32, 256, 5>, 256},
            {"H u64 c1 256", k1_h<64, 256, 1>, 256},  {"H u16 c1 128", k1_h<16, 128, 1>, 128}, {"H u32 c1 192", k1_h<32, 192, 1>, 192},
        };
        const uint64_t total_h = 1ull << (N - 1);
        for (auto &v : hv) {
            if (skip(v.name)) continue;
            const size_t smem = sizeof(double2) * ((size_t)N * N + (size_t)N * v.threads);
            CK(cudaFuncSetAttribute((const void *)v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cudaFuncAttributes attr;
            CK(cudaFuncGetAttributes(&attr, (const void *)v.fn));
            int occ_blocks = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_blocks, (const void *)v.fn, v.threads, smem));
            const uint64_t cap = (uint64_t)sms * occ_blocks * v.threads;
            uint64_t span = (total_h + cap - 1) / cap;
            span = ((span + 63) / 64) * 64;
            const uint64_t nthreads = (total_h + span - 1) / span;
            const int grid = (int)((nthreads + v.threads - 1) / v.threads);
            std::vector<double> hp(4 * (size_t)grid);
            float best = 1e30f;
            for (int rep = 0; rep < 4; ++rep) {
                CK(cudaEventRecord(e0));
                v.fn<<<grid, v.threads, smem>>>(dA, N, 0, total_h, span, dP);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                CK(cudaGetLastError());
                float ms;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                if (rep > 0 && ms < best) best = ms;
            }
            CK(cudaMemcpy(hp.data(), dP, sizeof(double) * 4 * grid, cudaMemcpyDeviceToHost));
            const double re = dd_total(hp, grid, 0) / (double)total_h, im = dd_total(hp, grid, 2) / (double)total_h;
            const double rel = hypot(re - want_re, im - want_im) / hypot(want_re, want_im);
            const double tf = (8.0 * N - 4) * (double)total_h / (best * 1e-3) / 1e12;
            printf("K1 %-14s regs %3d occ %d blk/SM smem %6zu grid %4d span %6llu  %7.3f ms  %6.2f TF useful  %6.1f perm/s  rel.err %.2e\n", v.name,
                   attr.numRegs, occ_blocks, smem, grid, (unsigned long long)span, best, tf, 1e3 / best, rel);
        }
    }
    return 0;
}
