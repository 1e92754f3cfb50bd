// glynn_kernel.cu -- K1: Gray-code Glynn permanent of an explicit N x N complex128 matrix.
//
// Replaces the loop of GlynnGrayPermanentCalculator.compute_permanent
// (reference: theboss/boson_sampling_utilities/permanent_calculators/
//  glynn_gray_permanent_calculator.py:55-71, init :73-84).
//
// Work decomposition (B200: 148 SMs, FP64 pipe 64 DFMA/clk/SM):
//   * the 2^(N-1) Gray steps [lo, hi) are cut into one contiguous span per thread; every span
//     starts on a multiple of 2^6, so that inside a 64-step window all lanes of a warp flip the
//     same row (ctz of the low bits) and the shared-memory row read is a broadcast;
//   * the N running column sums live in registers (static indexing: N is a template
//     parameter), the matrix A lives in shared memory.  The sums are kept HALVED (s/2, flips add
//     +-A[row] instead of +-2A[row]): no pre-doubling pass over the matrix; every product is then
//     exactly 2^-N times the reference's (powers of two: no rounding), undone when the partials
//     are combined;
//   * per step: N complex (sum += +-A[row]) updates and an (N-1)-multiply complex product in
//     three independent chains for ILP; (6N-2) FP64 issue slots for (8N-4) useful flops;
//   * terms are added in plain FP64 inside a 64-step window, windows are folded into a
//     double-double accumulator per thread, then warp-shuffle + shared-memory block reduction in
//     double-double; one partial per block, summed in block order (and scaled by 2^N) by the last
//     block to finish when one kernel covers the whole range, else by glynn_finish_kernel.
//
// Two kernels share this layout:
//   glynn_gray_kernel<N>    generic: any step range, one Gray step per loop iteration.
//   glynn_block4_kernel<N>  bulk path for 64-aligned ranges (K1B_MIN_N <= N <= K1B_MAX_N; smaller N spill under ptxas and are tiny anyway): four Gray steps per
//     iteration.  Inside an aligned block of four steps the flipped rows are 0, 1, 0 with
//     compile-time signs, so those rows are read from the CONSTANT bank into uniform registers that
//     the FP64 instructions take directly as operands (no vector registers, no address
//     arithmetic); only the block-closing flip fetches a run-time row from shared memory.
//     Measured on B200 (n = 30): every non-FP64 instruction costs about one FP64 issue slot, so
//     cutting the loop from 66 to ~29 non-FP64 instructions per step lifts the FP64 pipe from
//     70 % to 88 % busy (profiles/).
#include <mutex>
#include <stdlib.h>

#include "bp_common.cuh"

#define K1_THREADS 128
#define K1_WINDOW_LOG2 6

template <int N>
struct K1Cfg {
    static constexpr int MINB = (N <= 12) ? 6 : (N <= 20) ? 4 : (N <= 30) ? 3 : 2;
};

template <int N>
__device__ __forceinline__ void k1_product(const double (&sr)[N], const double (&si)[N], double &pr,
                                           double &pi) {
    constexpr int NCH = (N >= 13) ? 4 : (N >= 4 ? 2 : 1);   // independent chains: ILP for the FP64 pipe
    cplx p[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) { p[c].re = sr[c]; p[c].im = si[c]; }
#pragma unroll
    for (int j = NCH; j < N; ++j) {
        cplx s = {sr[j], si[j]};
        p[j % NCH] = cmul(p[j % NCH], s);
    }
#pragma unroll
    for (int stride = 1; stride < NCH; stride <<= 1)
#pragma unroll
        for (int c = 0; c + stride < NCH; c += 2 * stride) p[c] = cmul(p[c], p[c + stride]);
    pr = p[0].re;
    pi = p[0].im;
}

// Sums `nblocks` double-double complex partials in block order and scales by 2^scale_log2 (exact).  One warp; the result is
// returned in every lane and, when out_dd is given, written there by lane 0.
__device__ __forceinline__ void k1_sum_partials(const double *partials, int nblocks, int scale_log2, double *out_dd, double (&res)[4]) {
    const int lane = threadIdx.x & 31;
    dd re = {0.0, 0.0}, im = {0.0, 0.0};
    for (int b = lane; b < nblocks; b += 32) {
        dd a = {__ldcg(partials + 4 * b + 0), __ldcg(partials + 4 * b + 1)}, c = {__ldcg(partials + 4 * b + 2), __ldcg(partials + 4 * b + 3)};
        re = dd_add(re, a);
        im = dd_add(im, c);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        re = dd_add(re, dd_shfl_down(re, d));
        im = dd_add(im, dd_shfl_down(im, d));
    }
    res[0] = ldexp(re.hi, scale_log2); res[1] = ldexp(re.lo, scale_log2);
    res[2] = ldexp(im.hi, scale_log2); res[3] = ldexp(im.lo, scale_log2);
#pragma unroll
    for (int q = 0; q < 4; ++q) res[q] = __shfl_sync(0xffffffffu, res[q], 0);
    if (lane == 0 && out_dd) { out_dd[0] = res[0]; out_dd[1] = res[1]; out_dd[2] = res[2]; out_dd[3] = res[3]; }
}

// Partial exchange over peer memory (bp_glynn_matrix_range_exchange).  Slot buffers hold 2 x world slots of 8 doubles
// {re_hi, re_lo, im_hi, im_lo, call number, -, -, -}; calls alternate between the two halves, so a rank that is one call ahead
// never overwrites a slot a slower peer still has to read (it cannot be two ahead: its own wait needs the peer's arrival).
struct K1Exchange {
    double *peer[BP_MAX_PEERS];     // slot buffer of every rank as mapped into this process (peer[rank] = the local one)
    double *out_all;                // [world][4] on this device
    unsigned long long seq;         // call number, >= 1
    int world, rank;                // world == 0: no exchange
};
#define K1X_SLOT 8
#define K1X_TIMEOUT_NS 10000000000ull

__device__ __forceinline__ unsigned long long k1_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// One warp of the last block: lane r < world stores this rank's partial into rank r's buffer (payload, system-wide fence, call
// number), then waits for rank r's partial of the same call in the local buffer and copies it out.
__device__ __forceinline__ void k1_exchange(const K1Exchange &x, const double (&res)[4]) {
    const int lane = threadIdx.x & 31;
    if (lane >= x.world) return;
    const size_t half = (size_t)(x.seq & 1ull) * (size_t)x.world;
    volatile double *dst = x.peer[lane] + (half + (size_t)x.rank) * K1X_SLOT;
    dst[0] = res[0]; dst[1] = res[1]; dst[2] = res[2]; dst[3] = res[3];
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long *>(dst + 4) = x.seq;
    volatile double *src = x.peer[x.rank] + (half + (size_t)lane) * K1X_SLOT;
    volatile unsigned long long *flag = reinterpret_cast<volatile unsigned long long *>(src + 4);
    const unsigned long long t0 = k1_globaltimer();
    bool arrived = true;
    while (*flag != x.seq) {
        if (k1_globaltimer() - t0 > K1X_TIMEOUT_NS) { arrived = false; break; }
    }
    __threadfence_system();
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
#pragma unroll
    for (int q = 0; q < 4; ++q) x.out_all[4 * lane + q] = arrived ? src[q] : nan;
}

// Block epilogue shared by both kernels: thread 0 holds the block's partial.  With a counter (one kernel covers the whole
// range) the last block to arrive adds all partials in block order -- the same order and arithmetic as glynn_finish_kernel,
// so the result does not depend on which block is last -- resets the counter for the next launch, and runs the exchange.
__device__ __forceinline__ void k1_block_epilogue(dd acc_re, dd acc_im, double *__restrict__ partials, unsigned int *counter,
                                                  int scale_log2, double *__restrict__ out_dd, const K1Exchange &x) {
    __shared__ int is_last;
    if (threadIdx.x == 0) {
        double *o = partials + 4 * (size_t)blockIdx.x;
        o[0] = acc_re.hi; o[1] = acc_re.lo; o[2] = acc_im.hi; o[3] = acc_im.lo;
        int last = 0;
        if (counter) {
            __threadfence();
            last = (atomicAdd(counter, 1u) == gridDim.x - 1) ? 1 : 0;
        }
        is_last = last;
    }
    if (!counter) return;
    __syncthreads();
    if (is_last && threadIdx.x < 32) {
        __threadfence();
        double res[4];
        k1_sum_partials(partials, (int)gridDim.x, scale_log2, out_dd, res);
        if (threadIdx.x == 0) *counter = 0u;
        if (x.world > 0) k1_exchange(x, res);
    }
}

template <int N>
__global__ void __launch_bounds__(K1_THREADS, K1Cfg<N>::MINB)
glynn_gray_kernel(const double *__restrict__ A, uint64_t lo, uint64_t hi, uint64_t span,
                  double *__restrict__ partials, unsigned int *counter, double *__restrict__ out_dd, const K1Exchange x) {
    __shared__ double2 sA2[N * N];                    // A, row-major (the column sums are kept halved)
    __shared__ double red[4 * (K1_THREADS / 32)];

    for (int e = threadIdx.x; e < N * N; e += K1_THREADS) sA2[e] = reinterpret_cast<const double2 *>(A)[e];
    __syncthreads();

    const uint64_t gtid = (uint64_t)blockIdx.x * K1_THREADS + threadIdx.x;
    const uint64_t start = lo + gtid * span;
    dd acc_re = {0.0, 0.0}, acc_im = {0.0, 0.0};

    if (start < hi) {
        const uint64_t end = (hi - start < span) ? hi : start + span;
        double sr[N], si[N];
#pragma unroll
        for (int j = 0; j < N; ++j) { sr[j] = 0.0; si[j] = 0.0; }
        // sums_j = sum_i delta_i A[i][j]; delta from the Gray code of `start`; row N-1 is never
        // flipped (glynn_gray_permanent_calculator.py:57 iterates over N-1 rows only).
        const uint64_t g0 = start ^ (start >> 1);
#pragma unroll 1
        for (int i = 0; i < N; ++i) {
            const double sg = ((g0 >> i) & 1ull) ? -0.5 : 0.5;   // halved sums
            const double2 *row = sA2 + i * N;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const double2 a = row[j];
                sr[j] = fma(sg, a.x, sr[j]);
                si[j] = fma(sg, a.y, si[j]);
            }
        }
        double pr, pi;
        k1_product<N>(sr, si, pr, pi);
        const double ts0 = (start & 1ull) ? -1.0 : 1.0;
        double wr = ts0 * pr, wi = ts0 * pi;   // window accumulators (plain FP64)

#pragma unroll 1
        for (uint64_t I = start + 1; I < end; ++I) {
            const uint32_t Il = (uint32_t)I;
            if ((Il & ((1u << K1_WINDOW_LOG2) - 1u)) == 0u) {
                acc_re = dd_add_d(acc_re, wr);
                acc_im = dd_add_d(acc_im, wi);
                wr = 0.0; wi = 0.0;
            }
            const int r = Il ? (__ffs((int)Il) - 1) : (31 + __ffs((int)(uint32_t)(I >> 32)));
            // new delta_r = -1 iff bit r of gray(I) is set = bit_r(I) ^ bit_{r+1}(I) = !bit_{r+1}(I)
            const double sg = ((I >> (r + 1)) & 1ull) ? 1.0 : -1.0;
            const double2 *row = sA2 + r * N;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const double2 a = row[j];
                sr[j] = fma(sg, a.x, sr[j]);
                si[j] = fma(sg, a.y, si[j]);
            }
            k1_product<N>(sr, si, pr, pi);
            const double ts = (Il & 1u) ? -1.0 : 1.0;
            wr = fma(ts, pr, wr);
            wi = fma(ts, pi, wi);
        }
        acc_re = dd_add_d(acc_re, wr);
        acc_im = dd_add_d(acc_im, wi);
    }

    block_reduce_dd(acc_re, acc_im, red);
    k1_block_epilogue(acc_re, acc_im, partials, counter, N, out_dd, x);
}


// ---------------------------------------------------------------------------------------------
// bulk kernel: four Gray steps per iteration, rows 0 and 1 through the constant bank
// ---------------------------------------------------------------------------------------------
#define K1B_MIN_N 23
#define K1B_MAX_N 34

// Per-N launch shape of the bulk kernel (measured at N = 30, profiles/r01_k1_explore.txt): three warps per SMSP
// (384 threads, 168 registers) with a SINGLE product chain beat two warps (256 threads, 211 registers) with two
// chains, 5.57 vs 5.85 ms; beyond N = 30 the column sums alone need more than 168 registers.
template <int N>
struct K1BCfg {
    static constexpr int THREADS = (N <= 30) ? 384 : 256;
    static constexpr int NCH = (N <= 30) ? 1 : 2;
    // Run-time row flips (every fourth step) as DADDs of a SIGNED shared-memory image (A, then -A) instead of sign x row DFMAs: a
    // DADD takes 2.0 cycles of the pipe where a DFMA takes 2.18 (profiles/r02_rf_probe.txt).  Measured per N (profiles/r02_k1_wide_n.txt):
    // +1.5 .. +6 % for N = 25 .. 27, 29, 31 .. 34, no change at 23, 24, 28, and -3.3 % at N = 30, where ptxas pairs fewer DFMAs of the
    // product chain in the instantiation with the image -- that one keeps the DFMA form.
    static constexpr bool SIGNED_ROWS = (N != 30);
};

__constant__ double2 c_A2[BP_MAX_N * BP_MAX_N];   // A of the permanent in flight (row stride N)

// product of the N column sums in NCH (1 or 2) chains; the final multiplication is fused into the window
// accumulator (PLUS: w += p, else w -= p): 4(N-2) + 4 FP64 instructions per step
template <int N, bool PLUS>
__device__ __forceinline__ void k1b_product_acc(const double (&sr)[N], const double (&si)[N], double &wr, double &wi) {
    if constexpr (K1BCfg<N>::NCH == 1) {
        cplx p = {sr[0], si[0]};
#pragma unroll
        for (int j = 1; j < N - 1; ++j) { cplx s = {sr[j], si[j]}; p = cmul(p, s); }
        cplx last = {sr[N - 1], si[N - 1]};
        if (PLUS) cmul_acc(wr, wi, p, last); else cmul_sub(wr, wi, p, last);
    } else {
        cplx p0 = {sr[0], si[0]}, p1 = {sr[1], si[1]};
#pragma unroll
        for (int j = 2; j < N; ++j) {
            cplx s = {sr[j], si[j]};
            if (j & 1) p1 = cmul(p1, s); else p0 = cmul(p0, s);
        }
        if (PLUS) cmul_acc(wr, wi, p0, p1); else cmul_sub(wr, wi, p0, p1);
    }
}

template <int N, int ROW, int MODE>   // MODE 0: subtract, 1: add, 2: run-time sign
__device__ __forceinline__ void k1b_flip_const(double (&sr)[N], double (&si)[N], double sg) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
        if (MODE == 0)      { sr[j] -= c_A2[ROW * N + j].x; si[j] -= c_A2[ROW * N + j].y; }
        else if (MODE == 1) { sr[j] += c_A2[ROW * N + j].x; si[j] += c_A2[ROW * N + j].y; }
        else                { sr[j] = fma(sg, c_A2[ROW * N + j].x, sr[j]); si[j] = fma(sg, c_A2[ROW * N + j].y, si[j]); }
    }
}

// lo and hi are multiples of 64, span a multiple of 32.  Reads the matrix twice: c_A2 (constant bank, rows 0-1 in
// the loop) and A (global -> shared, run-time rows and the start state).
template <int N>
__global__ void __launch_bounds__(K1BCfg<N>::THREADS, 1)
glynn_block4_kernel(const double *__restrict__ A, uint64_t lo, uint64_t hi, uint64_t span,
                    double *__restrict__ partials, unsigned int *counter, double *__restrict__ out_dd, const K1Exchange x) {
    constexpr int THREADS = K1BCfg<N>::THREADS;
    constexpr bool SIGNED = K1BCfg<N>::SIGNED_ROWS;
    __shared__ double2 sA2[(SIGNED ? 2 : 1) * N * N];                // A (and -A)
    __shared__ double red[4 * (THREADS / 32)];
    for (int e = threadIdx.x; e < N * N; e += THREADS) {
        const double2 v = reinterpret_cast<const double2 *>(A)[e];
        sA2[e] = v;
        if constexpr (SIGNED) sA2[N * N + e] = make_double2(-v.x, -v.y);
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
            for (int j = 0; j < N; ++j) {
                const double2 a = row[j];
                sr[j] = fma(sg, a.x, sr[j]);
                si[j] = fma(sg, a.y, si[j]);
            }
        }
        double wr = 0.0, wi = 0.0;
        k1b_product_acc<N, true>(sr, si, wr, wi);      // step `start` (even: +)
#pragma unroll 1
        for (uint64_t I0 = start;;) {
            // steps I0+1, I0+2, I0+3 with I0 = 0 (mod 4): rows 0, 1, 0; new delta = -1, (bit 2 of I0 ? +1 : -1), +1
            k1b_flip_const<N, 0, 0>(sr, si, 0.0);
            k1b_product_acc<N, false>(sr, si, wr, wi);
            k1b_flip_const<N, 1, 2>(sr, si, ((I0 >> 2) & 1ull) ? 1.0 : -1.0);
            k1b_product_acc<N, true>(sr, si, wr, wi);
            k1b_flip_const<N, 0, 1>(sr, si, 0.0);
            k1b_product_acc<N, false>(sr, si, wr, wi);
            I0 += 4;
            if (I0 >= end) break;
            if (((uint32_t)I0 & 63u) == 0u) {
                acc_re = dd_add_d(acc_re, wr); acc_im = dd_add_d(acc_im, wi);
                wr = 0.0; wi = 0.0;
            }
            // step I0 (block-closing flip): run-time row >= 2, per-thread sign
            const uint32_t Il = (uint32_t)I0;
            const int r = Il ? (__ffs((int)Il) - 1) : (31 + __ffs((int)(uint32_t)(I0 >> 32)));
            if constexpr (SIGNED) {
                const double2 *row = sA2 + (r + (((I0 >> (r + 1)) & 1ull) ? 0 : N)) * N;
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    const double2 a = row[j];
                    sr[j] += a.x;
                    si[j] += a.y;
                }
            } else {
                const double sg = ((I0 >> (r + 1)) & 1ull) ? 1.0 : -1.0;
                const double2 *row = sA2 + r * N;
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    const double2 a = row[j];
                    sr[j] = fma(sg, a.x, sr[j]);
                    si[j] = fma(sg, a.y, si[j]);
                }
            }
            k1b_product_acc<N, true>(sr, si, wr, wi);
        }
        acc_re = dd_add_d(acc_re, wr);
        acc_im = dd_add_d(acc_im, wi);
    }
    block_reduce_dd(acc_re, acc_im, red);
    k1_block_epilogue(acc_re, acc_im, partials, counter, N, out_dd, x);
}

// ---------------------------------------------------------------------------------------------
// wide bulk kernel (K1W_MIN_N <= N <= BP_MAX_N): the block-4 layout with the COLUMNS of a Gray stream split over two warps
// ---------------------------------------------------------------------------------------------
// Beyond N = 34 the 4N registers of the column sums leave no room for the four-step body (profiles/r02_k1_wide_n.txt: spills,
// 0.33 .. 0.38 of peak against 0.46 for the generic kernel).  Here warps 2p and 2p + 1 of a block walk the SAME 32 Gray streams
// (lane = stream), the even warp over columns [0, H), the odd one over [H, N), H = ceil(N / 2): 2H <= 40 registers of sums per
// lane, row 0 still comes from the constant bank with compile-time addresses (the half is warp-uniform: each warp takes
// its own copy of the loop), and once per four steps the two warps trade half-products through shared memory under
// named barriers -- the even warp accumulates the first two steps of a block of four, the odd one the last two, so the FP64 work is
// the (6N - 4) instructions per step of the one-lane layout.  Two alternating exchange buffers make one rendezvous per
// iteration enough (a warp cannot be two iterations ahead of its partner).  Trip counts are block-uniform (span / 4
// iterations for every lane; steps at or beyond `end` are computed but not accumulated), so no lane skips a barrier.
#define K1W_MIN_N 35
#define K1W_THREADS 128
#define K1W_MINB 3
static_assert(4 * (K1W_THREADS / 64) <= 15, "four named barriers per warp pair, ids 1 .. 15");

// (H = ceil(N / 2) registers pairs per lane in both halves; a half owns the HC <= H columns [C0, C0 + HC))
template <int H, int HC>
__device__ __forceinline__ cplx k1w_half_product(const double (&sr)[H], const double (&si)[H]) {
    cplx p = {sr[0], si[0]};
#pragma unroll
    for (int j = 1; j < HC; ++j) {   // (both fused multiply-adds on p.re: the spelling with the fewest three-source DFMAs, scripts/sass_rf.py)
        cplx q = {sr[j], si[j]}, t;
        t.re = fma(p.re, q.re, -(p.im * q.im)); t.im = fma(p.re, q.im, p.im * q.re);
        p = t;
    }
    return p;
}
// Row 0 as constant-bank operands of the DADDs.  `r0` is c_A2 plus an offset that is always zero but formally depends on the loop
// counter: loop-INVARIANT constant loads are hoisted out of the step loop by ptxas -- into uniform registers while they last, then
// into vector registers that it spills (measured: 230 .. 690 bytes of spills at every register cap) -- whereas loop-variant ones
// are re-issued per use as uniform loads (LDCU.128) feeding the FP64 instructions directly, the block-4 kernel's pattern.
template <int H, int C0, int HC, int MODE>   // MODE 0: subtract, 1: add
__device__ __forceinline__ void k1w_flip_row0(const double2 *r0, double (&sr)[H], double (&si)[H]) {
#pragma unroll
    for (int j = 0; j < HC; ++j) {
        if (MODE == 0) { sr[j] -= r0[C0 + j].x; si[j] -= r0[C0 + j].y; }
        else           { sr[j] += r0[C0 + j].x; si[j] += r0[C0 + j].y; }
    }
}
// sums += sg * row (shared memory), or += row when SIGNED rows are stored
template <int H, int HC, bool FMA>
__device__ __forceinline__ void k1w_add_row(const double2 *row, double sg, double (&sr)[H], double (&si)[H]) {
#pragma unroll
    for (int j = 0; j < HC; ++j) {
        const double2 v = row[j];
        if (FMA) { sr[j] = fma(sg, v.x, sr[j]); si[j] = fma(sg, v.y, si[j]); }
        else     { sr[j] += v.x; si[j] += v.y; }
    }
}
// steps I0 + 1 .. I0 + 3 of an aligned block of four on one half's columns: rows 0, 1, 0; new delta = -1, (bit 2 of I0 ? +1 : -1), +1.
// Row 1 comes from a signed image in shared memory (+row 1, -row 1): as a constant-bank operand of a run-time-signed DFMA (the
// block-4 kernel's form) ptxas hoists the whole row out of the loop into registers and spills them here.
template <int N, int H, int C0, int HC>
__device__ __forceinline__ void k1w_three_steps(uint64_t I0, const double2 *r0, const double2 *srow1, double (&sr)[H], double (&si)[H],
                                                cplx &p1, cplx &p2, cplx &p3) {
    k1w_flip_row0<H, C0, HC, 0>(r0, sr, si);
    p1 = k1w_half_product<H, HC>(sr, si);
    k1w_add_row<H, HC, false>(srow1 + (((I0 >> 2) & 1ull) ? 0 : N) + C0, 0.0, sr, si);
    p2 = k1w_half_product<H, HC>(sr, si);
    k1w_flip_row0<H, C0, HC, 1>(r0, sr, si);
    p3 = k1w_half_product<H, HC>(sr, si);
}

// The walk of one warp.  HALF = 0: columns [0, H), accumulates the first two steps of every block of four; HALF = 1: columns [H, N),
// the last two.  Each half is its own straight-line loop (compile-time column offsets).  The exchange is the producer / consumer
// pattern of named barriers: a warp ARRIVES on the barrier its partner waits on (its two half-products are in shared memory) and
// SYNCS on its own (the partner's are) -- 32 arriving + 32 waiting threads per barrier.  Like the exchange buffers the barrier ids
// alternate between two sets with the parity of the iteration: a warp that runs ahead arrives on the OTHER set, never on a barrier
// whose previous phase its partner has not left yet (four ids per pair).
// xmine / xpeer: this lane's slots [buffer][product] in the exchange area, written by this warp / by the partner warp.
template <int N, int HALF>
__device__ __forceinline__ void k1w_walk(const double2 *sA2, const double2 *srow1, uint64_t start, uint64_t end, unsigned iters,
                                         double2 *xmine, const double2 *xpeer, int bar_mine, int bar_peer, dd &acc_re, dd &acc_im) {
    constexpr int H = (N + 1) / 2, C0 = HALF ? H : 0, HC = HALF ? N - H : H;
    constexpr int XS = 2 * 32;                       // double2s between the two buffers of a lane
    double sr[H], si[H];
#pragma unroll
    for (int j = 0; j < H; ++j) { sr[j] = 0.0; si[j] = 0.0; }
    const uint64_t g0 = start ^ (start >> 1);
#pragma unroll 1
    for (int i = 0; i < N; ++i)
        k1w_add_row<H, HC, true>(sA2 + i * N + C0, ((g0 >> i) & 1ull) ? -0.5 : 0.5, sr, si);
    double wr = 0.0, wi = 0.0;
    uint64_t I0 = start;
    cplx p0 = k1w_half_product<H, HC>(sr, si);                           // step I0 (even: +)
#pragma unroll 1
    for (unsigned it = 0; it < iters; ++it) {
        const bool live = I0 < end;                                       // spans end on multiples of 4: a block of steps is in or out
        const double2 *r0 = c_A2 + __any_sync(0xffffffffu, it >> 30);     // = c_A2: a warp vote keeps the zero in the uniform datapath (see k1w_flip_row0)
        cplx p1, p2, p3;
        k1w_three_steps<N, H, C0, HC>(I0, r0, srow1, sr, si, p1, p2, p3);
        double2 *mine = xmine + (it & 1u) * XS;
        const double2 *peer = xpeer + (it & 1u) * XS;
        if (HALF == 0) { mine[0] = make_double2(p2.re, p2.im); mine[32] = make_double2(p3.re, p3.im); }
        else           { mine[0] = make_double2(p0.re, p0.im); mine[32] = make_double2(p1.re, p1.im); }
        __threadfence_block();                                             // the stores above before the arrival
        __syncwarp();                                                      // (the flush / `live` branches of the previous iteration are per-lane)
        asm volatile("bar.arrive %0, 64;\n\tbar.sync %1, 64;" :: "r"(bar_peer + 2 * (int)(it & 1u)), "r"(bar_mine + 2 * (int)(it & 1u)) : "memory");
        const double2 qa = peer[0], qb = peer[32];
        const cplx a = {qa.x, qa.y}, b = {qb.x, qb.y};
        if (live) {
            if (HALF == 0) { cmul_acc(wr, wi, p0, a); cmul_sub(wr, wi, p1, b); }
            else           { cmul_acc(wr, wi, p2, a); cmul_sub(wr, wi, p3, b); }
        }
        I0 += 4;
        if (((uint32_t)I0 & 63u) == 0u) {
            acc_re = dd_add_d(acc_re, wr); acc_im = dd_add_d(acc_im, wi);
            wr = 0.0; wi = 0.0;
        }
        // step I0 (block-closing flip): run-time row >= 2, per-lane sign
        const uint32_t Il = (uint32_t)I0;
        const int r = Il ? (__ffs((int)Il) - 1) : (31 + __ffs((int)(uint32_t)(I0 >> 32)));
        const double2 *row = sA2 + ((r < N ? r : 0) + (((I0 >> (r + 1)) & 1ull) ? 0 : N)) * N + C0;   // (r >= N only past the end of the term space: not accumulated)
        k1w_add_row<H, HC, false>(row, 0.0, sr, si);
        p0 = k1w_half_product<H, HC>(sr, si);
    }
    acc_re = dd_add_d(acc_re, wr);
    acc_im = dd_add_d(acc_im, wi);
}

// lo and hi are multiples of 64, span a multiple of 32; one Gray stream per lane of a warp PAIR (K1W_THREADS / 2 streams per block)
template <int N>
__global__ void __launch_bounds__(K1W_THREADS, K1W_MINB)
glynn_pair4_kernel(const double *__restrict__ A, uint64_t lo, uint64_t hi, uint64_t span,
                   double *__restrict__ partials, unsigned int *counter, double *__restrict__ out_dd, const K1Exchange x) {
    extern __shared__ __align__(16) unsigned char k1w_dyn[];
    double2 *sA2 = reinterpret_cast<double2 *>(k1w_dyn);                 // A, then -A: run-time row flips add a signed row (K1W_SMEM bytes)
    __shared__ double2 xbuf[(K1W_THREADS / 64) * 2 * 2 * 2 * 32];       // [pair][half][buffer][product][lane]
    __shared__ double red[4 * (K1W_THREADS / 32)];
    __shared__ double2 srow1[2 * N];                                       // +row 1, -row 1
    for (int e = threadIdx.x; e < N * N; e += K1W_THREADS) {
        const double2 v = reinterpret_cast<const double2 *>(A)[e];
        sA2[e] = v; sA2[N * N + e] = make_double2(-v.x, -v.y);
    }
    for (int e = threadIdx.x; e < N; e += K1W_THREADS) {
        const double2 v = reinterpret_cast<const double2 *>(A)[N + e];
        srow1[e] = v; srow1[N + e] = make_double2(-v.x, -v.y);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, half = warp & 1, pair = warp >> 1;
    const uint64_t stream = (uint64_t)blockIdx.x * (K1W_THREADS / 2) + (uint64_t)(pair * 32 + lane);
    const uint64_t start = lo + stream * span;                            // (may lie beyond hi: the lane still keeps its partner company)
    const uint64_t end = (start >= hi) ? start : ((hi - start < span) ? hi : start + span);
    const unsigned iters = (unsigned)(span >> 2);
    double2 *xmine = xbuf + ((pair * 2 + half) * 4) * 32 + lane;
    const double2 *xpeer = xbuf + ((pair * 2 + (half ^ 1)) * 4) * 32 + lane;
    dd acc_re = {0.0, 0.0}, acc_im = {0.0, 0.0};
    // barrier ids 1 + 4 pair (+ 2 on odd iterations): the even warp waits on it; 2 + 4 pair (+ 2): the odd warp does; id 0 is __syncthreads
    if (half == 0) k1w_walk<N, 0>(sA2, srow1, start, end, iters, xmine, xpeer, 1 + 4 * pair, 2 + 4 * pair, acc_re, acc_im);
    else           k1w_walk<N, 1>(sA2, srow1, start, end, iters, xmine, xpeer, 2 + 4 * pair, 1 + 4 * pair, acc_re, acc_im);
    block_reduce_dd(acc_re, acc_im, red);
    k1_block_epilogue(acc_re, acc_im, partials, counter, N, out_dd, x);
}

// Sums the partials of several kernels (bulk + unaligned head / tail) in block order; one warp.
__global__ void glynn_finish_kernel(const double *__restrict__ partials, int nblocks, int scale_log2,
                                    double *__restrict__ out_dd, const K1Exchange x) {
    double res[4];
    k1_sum_partials(partials, nblocks, scale_log2, out_dd, res);
    if (x.world > 0) k1_exchange(x, res);
}

// ---------------------------------------------------------------------------------------------
// host-side launch
// ---------------------------------------------------------------------------------------------
typedef void (*k1_fn)(const double *, uint64_t, uint64_t, uint64_t, double *, unsigned int *, double *, const K1Exchange);

static k1_fn g_k1_fn[BP_MAX_N + 1], g_k1_bulk[BP_MAX_N + 1];
static int g_k1_minb[BP_MAX_N + 1], g_k1_bulk_threads[BP_MAX_N + 1];
static int g_k1_bulk_streams[BP_MAX_N + 1], g_k1_bulk_blocks[BP_MAX_N + 1];   // Gray streams per block, resident blocks per SM
static int g_k1_bulk_smem[BP_MAX_N + 1];                                       // dynamic shared memory of the bulk kernel (wide kernel: signed image of A)

template <int N>
static void k1_entry(k1_fn *fn, int *minb, k1_fn *bulk) {
    fn[N] = glynn_gray_kernel<N>;
    minb[N] = K1Cfg<N>::MINB;
    if constexpr (N >= K1B_MIN_N && N <= K1B_MAX_N) { bulk[N] = glynn_block4_kernel<N>; g_k1_bulk_threads[N] = K1BCfg<N>::THREADS; g_k1_bulk_streams[N] = K1BCfg<N>::THREADS; g_k1_bulk_blocks[N] = 1; }
    if constexpr (N >= K1W_MIN_N) { bulk[N] = glynn_pair4_kernel<N>; g_k1_bulk_threads[N] = K1W_THREADS; g_k1_bulk_streams[N] = K1W_THREADS / 2; g_k1_bulk_blocks[N] = K1W_MINB; g_k1_bulk_smem[N] = 2 * N * N * (int)sizeof(double2); }
    if constexpr (N > 1) k1_entry<N - 1>(fn, minb, bulk);
}

// c_A2 is one slot per device: uses are ordered by a host mutex plus an event the next writer waits on.
static std::mutex g_const_mutex;
static cudaEvent_t g_const_event[64];
static bool g_const_event_valid[64];

static int k1_generic(bp_context *h, const double *dA, int N, uint64_t lo, uint64_t hi, double *d_partials, int *grid_out,
                      unsigned int *d_counter, double *d_out_dd, const K1Exchange &x) {
    const uint64_t window = 1ull << K1_WINDOW_LOG2;
    const uint64_t total = hi - lo;
    const uint64_t max_threads = (uint64_t)h->sm_count * g_k1_minb[N] * K1_THREADS;
    uint64_t span = (total + max_threads - 1) / max_threads;
    span = ((span + window - 1) / window) * window;
    if (span == 0) span = window;
    uint64_t nthreads = (total + span - 1) / span;
    if (nthreads == 0) nthreads = 1;
    const int grid = (int)((nthreads + K1_THREADS - 1) / K1_THREADS);
    g_k1_fn[N]<<<grid, K1_THREADS, 0, h->stream>>>(dA, lo, hi, span, d_partials, d_counter, d_out_dd, x);
    BP_CHECK_LAUNCH(h);
    *grid_out = grid;
    return BP_OK;
}

// c_A2 currently holds the image of this (handle, matrix, declared-resident generation); guarded by g_const_mutex
struct K1ConstOwner { const bp_context *h; const double *dA; int N; uint64_t gen; };
static K1ConstOwner g_const_owner[64];

// Enqueue K1 over Gray steps [lo, hi) of an N x N device matrix; d_out_dd (may be NULL) receives the un-normalised
// double-double partial.  d_exchange_out != NULL: the kernel that finishes the sum also runs the peer-memory exchange of the
// handle (bp_exchange_*) and writes all ranks' partials there.
int bp_k1_launch(bp_context *h, const double *dA, int N, uint64_t lo, uint64_t hi, double *d_out_dd, double *d_exchange_out) {
    static const bool ready = [] {   // thread-safe one-time registration of the template instances
        k1_entry<BP_MAX_N>(g_k1_fn, g_k1_minb, g_k1_bulk);
        if (const char *e = getenv("BP_K1_BULK_MAX_N"))   // tuning: generic kernel above this N
            for (int n = atoi(e) + 1; n <= BP_MAX_N; ++n) if (n >= 1) g_k1_bulk[n] = nullptr;
        return true;
    }();
    (void)ready;
    if (N < 1 || N > BP_MAX_N) return bp_fail(h, BP_ERR_UNSUPPORTED, "K1 supports 1 <= N <= %d, got %d", BP_MAX_N, N);
    const uint64_t total_terms = 1ull << (N - 1);
    if (lo > hi || hi > total_terms) return bp_fail(h, BP_ERR_INVALID, "Gray step range [%llu, %llu) outside [0, 2^%d)",
                                                    (unsigned long long)lo, (unsigned long long)hi, N - 1);
    // aligned bulk [blo, bhi) for the block-4 kernel, unaligned head / tail for the generic one
    uint64_t blo = (lo + 63) & ~63ull, bhi = hi & ~63ull;
    const bool bulk = g_k1_bulk[N] != nullptr && bhi > blo && (bhi - blo) >= (1ull << 16) && h->device < 64;
    if (!bulk) { blo = hi; bhi = hi; }
    const int bulk_grid_max = h->sm_count * (g_k1_bulk[N] ? g_k1_bulk_blocks[N] : 1);
    const int gen_grid_max = h->sm_count * g_k1_minb[N];
    int rc = bp_reserve(h, BP_SLOT_PARTIALS, sizeof(double) * 4 * (size_t)(bulk_grid_max + 2 * gen_grid_max + 8));
    if (rc) return rc;
    double *d_partials = (double *)h->d_buf[BP_SLOT_PARTIALS];
    int nparts = 0;
    if (!h->d_counter) {   // arrival counter of the fused finish: zero once, every launch leaves it at zero
        BP_CUDA(h, cudaMalloc((void **)&h->d_counter, sizeof(unsigned int)));
        BP_CUDA(h, cudaMemsetAsync(h->d_counter, 0, sizeof(unsigned int), h->stream));
    }
    // one kernel covers the whole range (the usual case: full permanents and 64-aligned shards): its last block finishes
    const bool single = bulk ? (blo == lo && bhi == hi) : (hi > lo);
    unsigned int *d_counter = single ? h->d_counter : nullptr;
    K1Exchange none, xchg;
    none.world = 0; none.rank = 0; none.seq = 0; none.out_all = nullptr;
    xchg = none;
    if (d_exchange_out) {
        if (h->xchg_world < 1) return bp_fail(h, BP_ERR_INVALID, "bp_glynn_matrix_range_exchange: call bp_exchange_create / bp_exchange_connect first");
        for (int r = 0; r < h->xchg_world; ++r) {
            if (!h->xchg_peer[r]) return bp_fail(h, BP_ERR_INVALID, "bp_glynn_matrix_range_exchange: rank %d is not connected", r);
            xchg.peer[r] = h->xchg_peer[r];
        }
        xchg.world = h->xchg_world; xchg.rank = h->xchg_rank; xchg.seq = ++h->xchg_seq; xchg.out_all = d_exchange_out;
    }
    const K1Exchange &xs = single ? xchg : none;     // the exchange runs in whichever kernel finishes the sum
    if (bulk) {
        const size_t bytes = sizeof(double2) * (size_t)N * N;
        const uint64_t total = bhi - blo;
        const int bthreads = g_k1_bulk_threads[N], bstreams = g_k1_bulk_streams[N];
        const uint64_t threads = (uint64_t)bulk_grid_max * bstreams;                 // Gray streams in flight
        // spans are multiples of 32 steps: lanes of a warp then differ by a multiple of 32, so inside a warp only the
        // block-closing flips at multiples of 32 (1 block in 8) read two different rows; the finer grain keeps the load
        // balance above 99 % when the range is cut over 8 GPUs (2^26 steps over 56832 threads = 1180.8 per thread)
        uint64_t span = (total + threads - 1) / threads;
        span = ((span + 31) / 32) * 32;
        const int grid = (int)(((total + span - 1) / span + bstreams - 1) / bstreams);
        {
            std::lock_guard<std::mutex> g(g_const_mutex);
            K1ConstOwner &own = g_const_owner[h->device];
            // a matrix declared resident whose image is still the last one written: nothing to copy (and nothing to wait for --
            // every launch that read or wrote c_A2 since then came from this handle's own stream)
            const bool keep = g_const_event_valid[h->device] && h->resident_A == dA && own.h == h && own.dA == dA && own.N == N &&
                              own.gen == h->resident_gen;
            if (!keep) {
                if (!g_const_event_valid[h->device]) {
                    BP_CUDA(h, cudaEventCreateWithFlags(&g_const_event[h->device], cudaEventDisableTiming));
                    g_const_event_valid[h->device] = true;
                } else {
                    BP_CUDA(h, cudaStreamWaitEvent(h->stream, g_const_event[h->device], 0));   // previous user of c_A2
                }
                BP_CUDA(h, cudaMemcpyToSymbolAsync(c_A2, dA, bytes, 0, cudaMemcpyDeviceToDevice, h->stream));
                own.h = h; own.dA = dA; own.N = N; own.gen = (h->resident_A == dA) ? h->resident_gen : ~0ull;
            }
            if (g_k1_bulk_smem[N] > 0) {   // (up to 51 KB of dynamic + 10 KB of static shared memory: beyond the 48 KB that need no opt-in)
                cudaError_t e = cudaFuncSetAttribute((const void *)g_k1_bulk[N], cudaFuncAttributeMaxDynamicSharedMemorySize, g_k1_bulk_smem[N]);
                if (e != cudaSuccess) return bp_fail(h, BP_ERR_CUDA, "cudaFuncSetAttribute(%d): %s", g_k1_bulk_smem[N], cudaGetErrorString(e));
            }
            g_k1_bulk[N]<<<grid, bthreads, (size_t)g_k1_bulk_smem[N], h->stream>>>(dA, blo, bhi, span, d_partials, d_counter, d_out_dd, xs);
            BP_CHECK_LAUNCH(h);
            BP_CUDA(h, cudaEventRecord(g_const_event[h->device], h->stream));
        }
        nparts += grid;
    }
    if (blo > lo) {   // head (everything when the bulk path is not taken)
        int g = 0;
        if ((rc = k1_generic(h, dA, N, lo, blo, d_partials + 4 * nparts, &g, d_counter, d_out_dd, xs))) return rc;
        nparts += g;
    }
    if (hi > bhi) {   // tail
        int g = 0;
        if ((rc = k1_generic(h, dA, N, bhi, hi, d_partials + 4 * nparts, &g, nullptr, d_out_dd, none))) return rc;
        nparts += g;
    }
    if (nparts == 0 && !d_exchange_out) {   // empty range
        if (d_out_dd) BP_CUDA(h, cudaMemsetAsync(d_out_dd, 0, sizeof(double) * 4, h->stream));
        return BP_OK;
    }
    if (single) return BP_OK;   // the kernel's last block has written d_out_dd (and exchanged it)
    glynn_finish_kernel<<<1, 32, 0, h->stream>>>(d_partials, nparts, N, d_out_dd, xchg);   // (an empty range contributes zero)
    BP_CHECK_LAUNCH(h);
    return BP_OK;
}
