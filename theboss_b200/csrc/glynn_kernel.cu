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
//     parameter), the pre-doubled matrix 2A lives in shared memory;
//   * per step: N complex (sum += +-2A[row]) updates and an (N-1)-multiply complex product in
//     three independent chains for ILP; (6N-2) FP64 issue slots for (8N-4) useful flops;
//   * terms are added in plain FP64 inside a 64-step window, windows are folded into a
//     double-double accumulator per thread, then warp-shuffle + shared-memory block reduction in
//     double-double; one partial per block, summed in block order by glynn_finish_kernel.
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
    constexpr int NCH = (N >= 9) ? 3 : (N >= 4 ? 2 : 1);
    cplx p[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) { p[c].re = sr[c]; p[c].im = si[c]; }
#pragma unroll
    for (int j = NCH; j < N; ++j) {
        cplx s = {sr[j], si[j]};
        p[j % NCH] = cmul(p[j % NCH], s);
    }
    cplx r = p[0];
#pragma unroll
    for (int c = 1; c < NCH; ++c) r = cmul(r, p[c]);
    pr = r.re;
    pi = r.im;
}

template <int N>
__global__ void __launch_bounds__(K1_THREADS, K1Cfg<N>::MINB)
glynn_gray_kernel(const double *__restrict__ A, uint64_t lo, uint64_t hi, uint64_t span,
                  double *__restrict__ partials) {
    __shared__ double2 sA2[N * N];                    // 2*A, row-major
    __shared__ double red[4 * (K1_THREADS / 32)];

    for (int e = threadIdx.x; e < N * N; e += K1_THREADS) {
        double2 v = reinterpret_cast<const double2 *>(A)[e];
        sA2[e] = make_double2(2.0 * v.x, 2.0 * v.y);
    }
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
            const double sg = ((g0 >> i) & 1ull) ? -0.5 : 0.5;   // times the pre-doubled entry
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
    if (threadIdx.x == 0) {
        double *o = partials + 4 * (size_t)blockIdx.x;
        o[0] = acc_re.hi; o[1] = acc_re.lo; o[2] = acc_im.hi; o[3] = acc_im.lo;
    }
}

// Sums `nblocks` double-double complex partials in block order.  One warp.
__global__ void glynn_finish_kernel(const double *__restrict__ partials, int nblocks,
                                    double *__restrict__ out_dd) {
    dd re = {0.0, 0.0}, im = {0.0, 0.0};
    for (int b = threadIdx.x; b < nblocks; b += 32) {
        dd a = {partials[4 * b + 0], partials[4 * b + 1]}, c = {partials[4 * b + 2], partials[4 * b + 3]};
        re = dd_add(re, a);
        im = dd_add(im, c);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        re = dd_add(re, dd_shfl_down(re, d));
        im = dd_add(im, dd_shfl_down(im, d));
    }
    if (threadIdx.x == 0) {
        out_dd[0] = re.hi; out_dd[1] = re.lo; out_dd[2] = im.hi; out_dd[3] = im.lo;
    }
}

// ---------------------------------------------------------------------------------------------
// host-side launch
// ---------------------------------------------------------------------------------------------
typedef void (*k1_fn)(const double *, uint64_t, uint64_t, uint64_t, double *);

template <int N>
static void k1_entry(k1_fn *fn, int *minb) {
    fn[N] = glynn_gray_kernel<N>;
    minb[N] = K1Cfg<N>::MINB;
    if constexpr (N > 1) k1_entry<N - 1>(fn, minb);
}

static k1_fn g_k1_fn[BP_MAX_N + 1];
static int g_k1_minb[BP_MAX_N + 1];
static bool g_k1_init = false;

// Enqueue K1 over Gray steps [lo, hi) of an N x N device matrix; d_out_dd receives the
// un-normalised double-double partial.  d_partials must hold 4 * grid doubles.
int bp_k1_launch(bp_context *h, const double *dA, int N, uint64_t lo, uint64_t hi, double *d_out_dd) {
    if (!g_k1_init) { k1_entry<BP_MAX_N>(g_k1_fn, g_k1_minb); g_k1_init = true; }
    if (N < 1 || N > BP_MAX_N) return bp_fail(h, BP_ERR_UNSUPPORTED, "K1 supports 1 <= N <= %d, got %d", BP_MAX_N, N);
    const uint64_t total_terms = 1ull << (N - 1);
    if (lo > hi || hi > total_terms) return bp_fail(h, BP_ERR_INVALID, "Gray step range [%llu, %llu) outside [0, 2^%d)",
                                                    (unsigned long long)lo, (unsigned long long)hi, N - 1);
    const uint64_t window = 1ull << K1_WINDOW_LOG2;
    const uint64_t total = hi - lo;
    const uint64_t max_threads = (uint64_t)h->sm_count * g_k1_minb[N] * K1_THREADS;
    uint64_t span = (total + max_threads - 1) / max_threads;
    span = ((span + window - 1) / window) * window;
    if (span == 0) span = window;
    uint64_t nthreads = (total + span - 1) / span;
    if (nthreads == 0) nthreads = 1;
    const int grid = (int)((nthreads + K1_THREADS - 1) / K1_THREADS);
    int rc = bp_reserve(h, BP_SLOT_PARTIALS, sizeof(double) * 4 * (size_t)grid);
    if (rc) return rc;
    double *d_partials = (double *)h->d_buf[BP_SLOT_PARTIALS];
    g_k1_fn[N]<<<grid, K1_THREADS, 0, h->stream>>>(dA, lo, hi, span, d_partials);
    BP_CHECK_LAUNCH(h);
    glynn_finish_kernel<<<1, 32, 0, h->stream>>>(d_partials, grid, d_out_dd);
    BP_CHECK_LAUNCH(h);
    return BP_OK;
}
