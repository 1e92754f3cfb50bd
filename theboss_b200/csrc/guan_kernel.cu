// guan_kernel.cu -- K2: batched single permanents with input/output multiplicities.
//
// Replaces B calls of compute_permanent() of the Guan-code calculators
// (reference: theboss/boson_sampling_utilities/permanent_calculators/
//  bs_permanent_calculator_base.py:166-209, chin_huh_permanent_calculator.py:38-59,
//  ryser_permanent_calculator.py:45-64; also glynn_gray_permanent_calculator.py:41-71, which
//  evaluates the same quantity on the expanded matrix) that share one interferometer U.
//
// Chin-Huh form, generalised Gray (Guan) walk over the side with FEWER terms (perm A = perm A^T):
//   perm = 2^-n sum_r (-1)^{sum r} prod_v C(w_v, r_v) prod_{j=1..n} ( sum_v (w_v - 2 r_v) X[v][j] )
// with the product side expanded to its n particles (columns repeated by multiplicity), and the
// r <-> w - r symmetry halving the walk (guan_walker.cuh).
//
// Scheduling: one thread block per item; a counting sort orders items by (n, descending cost) so
// that one templated launch per distinct n runs longest-first (LPT) over the 148 SMs.
#include "bp_common.cuh"
#include "guan_walker.cuh"

struct K2Meta {            // per item, written by k2_prep_kernel
    int n;                 // particles (-1: sum(s) != sum(t))
    int walk_outputs;      // 1: walk over output modes (rows), 0: over input modes (columns)
    float log2cost;
};

// ---------------------------------------------------------------------------------------------
// prep: per item particle numbers, walk side, cost
// ---------------------------------------------------------------------------------------------
__global__ void k2_prep_kernel(const unsigned char *__restrict__ S, const unsigned char *__restrict__ T, int m,
                               long long B, K2Meta *__restrict__ meta, int *__restrict__ bins /*[41*64]*/) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const unsigned char *s = S + b * m, *t = T + b * m;
    int ns = 0, nt = 0;
    for (int v = 0; v < m; ++v) { ns += s[v]; nt += t[v]; }
    K2Meta me;
    if (ns != nt || ns > BP_MAX_N) { me.n = -1; me.walk_outputs = 0; me.log2cost = 0.f; meta[b] = me; return; }
    const double cs = guan_terms_of(s, m), ct = guan_terms_of(t, m);
    me.n = ns;
    me.walk_outputs = (ct < cs) ? 1 : 0;
    me.log2cost = (float)log2(ct < cs ? ct : cs);
    meta[b] = me;
    int bucket = (int)me.log2cost;
    if (bucket > 63) bucket = 63;
    atomicAdd(&bins[ns * 64 + (63 - bucket)], 1);   // descending cost inside each n
}

// exclusive scan of the 41*64 bins + per-n offsets: one block of 1024 threads, three bins per thread, warp-shuffle scans
#define K2_NBINS ((BP_MAX_N + 1) * 64)
__global__ void __launch_bounds__(1024) k2_scan_kernel(int *__restrict__ bins, int *__restrict__ n_offsets /*[42]*/) {
    constexpr int PER = (K2_NBINS + 1023) / 1024;
    __shared__ int warp_sum[32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    int v[PER], mine = 0;
#pragma unroll
    for (int q = 0; q < PER; ++q) { const int i = t * PER + q; v[q] = i < K2_NBINS ? bins[i] : 0; mine += v[q]; }
    int x = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
    if (lane == 31) warp_sum[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int w = warp_sum[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, d); if (lane >= d) w += y; }
        warp_sum[lane] = w;
    }
    __syncthreads();
    int run = x - mine + (warp ? warp_sum[warp - 1] : 0);   // exclusive prefix of this thread's first bin
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const int i = t * PER + q;
        if (i < K2_NBINS) {
            bins[i] = run;
            if ((i & 63) == 0) n_offsets[i >> 6] = run;
        }
        run += v[q];
    }
    if (t == 1023) n_offsets[BP_MAX_N + 1] = run;
}

__global__ void k2_scatter_kernel(const K2Meta *__restrict__ meta, long long B, int *__restrict__ bins,
                                  int *__restrict__ order, double *__restrict__ out) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const K2Meta me = meta[b];
    if (me.n < 0) {   // shape error: the host API refuses earlier; the _dev API marks the item
        out[2 * b] = __longlong_as_double(0x7ff8000000000000ll); out[2 * b + 1] = out[2 * b];
        return;
    }
    int bucket = (int)me.log2cost;
    if (bucket > 63) bucket = 63;
    const int pos = atomicAdd(&bins[me.n * 64 + (63 - bucket)], 1);
    order[pos] = (int)b;
}

// ---------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------
#define K2_PMAX 512
#ifndef K2_PERIODS_PER_THREAD
#define K2_PERIODS_PER_THREAD 8    // a thread owns at least this many periods (thread ranges differ by at most one period; 16 -> 8: config 2 -1.1 %)
#endif
struct __align__(16) K2Step { double blow; int off; int pad; };   // binomial product at the position; byte offset of the signed row that leads there

// 16-byte shared-memory load at a 32-bit shared-window address plus a compile-time byte offset (one address register per row)
template <int OFF>
__device__ __forceinline__ double2 k2_lds(unsigned addr) {
    double2 v;
    asm("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "r"(addr), "n"(OFF));
    return v;
}
// sums[j] += row[j], j = 0 .. N-1
template <int N, int J = 0>
__device__ __forceinline__ void k2_row_add(unsigned row, double (&sr)[N], double (&si)[N]) {
    if constexpr (J < N) {
        const double2 a = k2_lds<J * (int)sizeof(double2)>(row);
        sr[J] += a.x;
        si[J] += a.y;
        k2_row_add<N, J + 1>(row, sr, si);
    }
}

template <int N>
struct K2Cfg {
    static constexpr int MINB = (N <= 8) ? 4 : (N <= 26) ? 3 : 2;
};
__host__ __device__ inline size_t k2_smem_bytes(int N) {
    return (size_t)(2 * N + 1) * N * sizeof(double2) + (size_t)N * GW_THREADS + 16;
}

// Product of the N column sums as a balanced tree, evaluated depth-first (at most log2 N partial products are live).
// The columns of one mode are adjacent (the product side is expanded in mode order), so a mode that holds w particles
// contributes its factor through ~log2 w squarings -- the rounding behaviour of the reference's `pow`
// (bs_permanent_calculator_base.py:200-209) instead of w - 1 sequential multiplications.
template <int N, int LO, int HI>
__device__ __forceinline__ cplx k2_tree(const double (&sr)[N], const double (&si)[N]) {
    if constexpr (HI - LO == 1) { cplx v = {sr[LO], si[LO]}; return v; }
    else {
        constexpr int MID = (LO + HI) / 2;
        return cmul(k2_tree<N, LO, MID>(sr, si), k2_tree<N, MID, HI>(sr, si));
    }
}

// The walk of one item is organised like the minors kernel's (minors_kernel.cu): ONE uniform term loop.  The lowest digits
// (digit 0 = the largest multiplicity first) follow a per-block step table of period P; inside a period every term fetches
// the 16-byte entry of the next position, multiplies the N column sums, accumulates, and ADDS the signed row of the digit that
// changes (shared-memory image: +2X rows, -2X rows, a zero row for the sentinel steps at both table ends).  The generic Guan
// stepper only carries into the digits above the table, once per period.  Threads own whole periods.
template <int N>
__global__ void __launch_bounds__(GW_THREADS, K2Cfg<N>::MINB)
k2_perm_kernel(const double *__restrict__ U, int m, const unsigned char *__restrict__ S,
               const unsigned char *__restrict__ T, const K2Meta *__restrict__ meta,
               const int *__restrict__ order, int first, double *__restrict__ partials, double *__restrict__ out) {
    constexpr int ROWBYTES = N * (int)sizeof(double2);
    extern __shared__ __align__(16) unsigned char k2_smem[];
    __shared__ GuanItem item;
    __shared__ short col_mode[N];
    __shared__ double red[4 * (GW_THREADS / 32)];
    __shared__ K2Step fwd[K2_PMAX + 2], bwd[K2_PMAX + 2];   // indexed by destination position + 1; entries 0 and P + 1: sentinels
    __shared__ int low_digits;
    __shared__ unsigned period;

    const int b = order[first + blockIdx.x];
    const K2Meta me = meta[b];   // (blockIdx.y = chunk of this item's walk)
    const unsigned char *walk = (me.walk_outputs ? T : S) + (long long)b * m;
    const unsigned char *prod = (me.walk_outputs ? S : T) + (long long)b * m;
    if (threadIdx.x < 32) guan_item_build_warp(item, walk, m, /*inner_first=*/true);       // block setup by whole warps
    else if (threadIdx.x < 64) guan_expand_columns_warp(col_mode, prod, m, N);
    __syncthreads();
    const int D = item.D;
    double2 *X2 = reinterpret_cast<double2 *>(k2_smem);             // rows [0, D): +2X, [D, 2D): -2X, row 2D: 0
    unsigned char *rdig = k2_smem + (size_t)(2 * N + 1) * N * sizeof(double2);   // per-thread digit vectors (column = thread)
    const double2 *U2 = reinterpret_cast<const double2 *>(U);
    for (int e = threadIdx.x; e < D * N; e += GW_THREADS) {
        const int v = e / N, j = e - v * N;
        const int wm = item.mode[v], pm = col_mode[j];
        // U[out_mode][in_mode]: walking outputs -> rows of U, walking inputs -> columns of U
        const double2 u = me.walk_outputs ? U2[wm * m + pm] : U2[pm * m + wm];
        X2[e] = make_double2(2.0 * u.x, 2.0 * u.y);
        X2[D * N + e] = make_double2(-2.0 * u.x, -2.0 * u.y);
    }
    for (int j = threadIdx.x; j < N; j += GW_THREADS) X2[2 * D * N + j] = make_double2(0.0, 0.0);
    const unsigned long long terms = item.terms;
    const unsigned long long nthreads = (unsigned long long)gridDim.y * GW_THREADS;
    if (threadIdx.x == 0) {
        unsigned long long raw = (terms + nthreads - 1) / nthreads, P = 1;
        int a = -1;
        for (int v = 0; v < D; ++v) {
            const unsigned long long nxt = P * (unsigned long long)(item.lim[v] + 1);
            if (v > 0 && (nxt * K2_PERIODS_PER_THREAD > raw || nxt > K2_PMAX)) break;   // digit 0 always: work is dealt out in whole sweeps of it
            P = nxt; a = v;
        }
        period = (unsigned)P;
        low_digits = a;
    }
    __syncthreads();
    const unsigned P = period;
    const int a_low = low_digits;
    for (unsigned p = threadIdx.x; p < P; p += GW_THREADS) {
        double bprod = 1.0;
        unsigned q = p, qm = p ? p - 1 : 0;
        int chg = 0, up = 0;
        for (int v = 0; v <= a_low; ++v) {
            const unsigned R = (unsigned)item.lim[v] + 1u;
            const unsigned d = gw_divmod(q, R), dm = gw_divmod(qm, R);
            const int rv = (q & 1u) ? (int)item.lim[v] - (int)d : (int)d;
            const int rm = (qm & 1u) ? (int)item.lim[v] - (int)dm : (int)dm;
            if (p && rv != rm) { chg = v; up = rv > rm; }
            double c = gw_binom(item.mult[v], rv);
            if (v == D - 1) c *= gw_top_weight(item, rv);
            bprod *= c;
        }
        // a digit that goes UP lowers its coefficient w - 2 r by 2: the negated row
        const int zero_row = 2 * D * ROWBYTES;
        fwd[p + 1].blow = bprod; fwd[p + 1].off = p ? (chg + (up ? D : 0)) * ROWBYTES : zero_row; fwd[p + 1].pad = 0;
        bwd[p + 1].blow = bprod;
        if (p == P - 1) { bwd[p + 1].off = zero_row; bwd[p + 1].pad = 0; }
        if (p) { bwd[p].off = (chg + (up ? 0 : D)) * ROWBYTES; bwd[p].pad = 0; }
        if (p == 0) {
            fwd[P + 1].blow = 0.0; fwd[P + 1].off = zero_row; fwd[P + 1].pad = 0;
            bwd[0].blow = 0.0; bwd[0].off = zero_row; bwd[0].pad = 0;
        }
    }
    __syncthreads();

    // whole periods are dealt out to the threads of all chunk blocks of this item
    const unsigned long long nper = terms / P;
    const unsigned long long tidx = (unsigned long long)blockIdx.y * GW_THREADS + threadIdx.x;
    const unsigned long long pbase = nper / nthreads, prem = nper % nthreads;
    const unsigned my_periods = (unsigned)(pbase + (tidx < prem ? 1ull : 0ull));   // (<= 2^39 terms over >= 128 threads: fits 32 bits)
    const unsigned long long hi0 = tidx * pbase + (tidx < prem ? tidx : prem);
    dd acc_re = {0.0, 0.0}, acc_im = {0.0, 0.0};

    if (my_periods > 0) {
        unsigned char *r = rdig + threadIdx.x;
        GuanState st;
        guan_seek(item, hi0, r, st, /*v0=*/a_low + 1);
        int pos = (hi0 & 1ull) ? (int)P - 1 : 0;
        int pdir = (hi0 & 1ull) ? -1 : 1;
        double sr[N], si[N];
#pragma unroll
        for (int j = 0; j < N; ++j) { sr[j] = 0.0; si[j] = 0.0; }
        int par = 0;
        {
            unsigned q = (unsigned)pos;
#pragma unroll 1
            for (int v = 0; v < D; ++v) {
                int rv;
                if (v <= a_low) {
                    const unsigned R = (unsigned)item.lim[v] + 1u;
                    const unsigned d = gw_divmod(q, R);
                    rv = (q & 1u) ? (int)item.lim[v] - (int)d : (int)d;
                } else rv = (int)r[v * GW_THREADS];
                par += rv;
                const double c = 0.5 * (double)((int)item.mult[v] - 2 * rv);
                const double2 *row = X2 + v * N;
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    const double2 a = row[j];
                    sr[j] = fma(c, a.x, sr[j]);
                    si[j] = fma(c, a.y, si[j]);
                }
            }
        }
        double bsgn = (par & 1) ? -st.binom : st.binom;      // sign of the term x binomial product of the digits above the table
        const K2Step *tab = (pdir > 0) ? fwd : bwd;
        double blow = fwd[pos + 1].blow;
        // shared-window address of the image, kept opaque so that row loads use ONE address register plus immediates
        unsigned x2base = (unsigned)__cvta_generic_to_shared(X2);
        asm volatile("" : "+r"(x2base));
        double wr = 0.0, wi = 0.0;

#pragma unroll 1
        for (unsigned per = 0;;) {
#pragma unroll 1
            for (unsigned i = 0; i < P; ++i) {
                pos += pdir;
                const K2Step e = tab[pos + 1];               // step to the NEXT position, fetched ahead of the product
                const double w = bsgn * blow;
                const cplx pr = k2_tree<N, 0, N>(sr, si);
                wr = fma(w, pr.re, wr);
                wi = fma(w, pr.im, wi);
                bsgn = -bsgn;
                blow = e.blow;
                k2_row_add<N>(x2base + (unsigned)e.off, sr, si);
            }
            // ---- period boundary: the period's plain FP64 sum (P <= 512 terms) goes into the double-double accumulator; then
            // one Guan step of the digits above the table; the table digits stay and reverse
            acc_re = dd_add_d(acc_re, wr); acc_im = dd_add_d(acc_im, wi);
            wr = 0.0; wi = 0.0;
            if (++per >= my_periods) break;
            pos -= pdir;
            pdir = -pdir;
            tab = (pdir > 0) ? fwd : bwd;
            blow = fwd[pos + 1].blow;
            int delta;
            const int v = guan_step(item, r, st, delta, /*v0=*/a_low + 1);
            k2_row_add<N>(x2base + (unsigned)((v + (delta > 0 ? D : 0)) * ROWBYTES), sr, si);
            bsgn = (bsgn < 0.0) ? -st.binom : st.binom;
        }
    }
    block_reduce_dd(acc_re, acc_im, red);
    if (threadIdx.x == 0) {
        if (gridDim.y == 1) {   // the whole walk of the item in this block: finish here, 2^-n (chin_huh_permanent_calculator.py:41)
            const double scale = ldexp(1.0, -N);
            out[2 * (size_t)b] = (acc_re.hi + acc_re.lo) * scale;
            out[2 * (size_t)b + 1] = (acc_im.hi + acc_im.lo) * scale;
        } else {
            double *o = partials + 4 * ((size_t)blockIdx.x * gridDim.y + blockIdx.y);
            o[0] = acc_re.hi; o[1] = acc_re.lo; o[2] = acc_im.hi; o[3] = acc_im.lo;
        }
    }
}

// out[item] = 2^-n * (sum of the item's chunk partials, in chunk order)
__global__ void k2_finish_kernel(const int *__restrict__ order, int first, int count, int chunks, int n,
                                 const double *__restrict__ partials, double *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    dd re = {0.0, 0.0}, im = {0.0, 0.0};
    const double *p = partials + 4 * (size_t)i * chunks;
    for (int c = 0; c < chunks; ++c) {
        dd a = {p[4 * c + 0], p[4 * c + 1]}, b = {p[4 * c + 2], p[4 * c + 3]};
        re = dd_add(re, a);
        im = dd_add(im, b);
    }
    const double scale = ldexp(1.0, -n);                      // 2^-n (chin_huh_permanent_calculator.py:41)
    const long long b = order[first + i];
    out[2 * b] = (re.hi + re.lo) * scale;
    out[2 * b + 1] = (im.hi + im.lo) * scale;
}

// items without particles: permanent of the empty matrix = 1
__global__ void k2_empty_kernel(const int *__restrict__ order, int first, int count, double *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const long long b = order[first + i];
    out[2 * b] = 1.0; out[2 * b + 1] = 0.0;
}

typedef void (*k2_fn)(const double *, int, const unsigned char *, const unsigned char *, const K2Meta *, const int *, int, double *, double *);
template <int N>
static void k2_entry(k2_fn *fn) {
    fn[N] = k2_perm_kernel<N>;
    if constexpr (N > 1) k2_entry<N - 1>(fn);
}
static k2_fn g_k2_fn[BP_MAX_N + 1];

#include <algorithm>
#include <vector>

// Batches up to this size are prepared (particle numbers, walk side, cost order) on the host while the matrix uploads: no
// prep / scan / scatter launches -- a lone compute_permanent() is one copy in, ONE kernel, one copy out.
#define K2_HOST_PREP_MAX 256

// Particle numbers, walk side and cost order of a small batch, on the host: what k2_prep / k2_scan / k2_scatter compute.
static void k2_host_prep(const unsigned char *hS, const unsigned char *hT, int m, long long B, K2Meta *hm, int *ho, int off[BP_MAX_N + 2]) {
    int counts[BP_MAX_N + 2] = {0};
    for (long long b = 0; b < B; ++b) {
        const unsigned char *s = hS + b * m, *t = hT + b * m;
        int ns = 0;
        for (int v = 0; v < m; ++v) ns += s[v];
        const double cs = guan_terms_of(s, m), ct = guan_terms_of(t, m);
        hm[b].n = ns;                                  // (sum(s) == sum(t) <= BP_MAX_N: checked by the caller)
        hm[b].walk_outputs = (ct < cs) ? 1 : 0;
        hm[b].log2cost = (float)log2(ct < cs ? ct : cs);
        ho[b] = (int)b;
        counts[ns]++;
    }
    std::stable_sort(ho, ho + B, [hm](int x, int y) {
        return hm[x].n != hm[y].n ? hm[x].n < hm[y].n : hm[x].log2cost > hm[y].log2cost;   // by n, longest first inside
    });
    off[0] = 0;
    for (int n = 0; n <= BP_MAX_N; ++n) off[n + 1] = off[n] + counts[n];
}

// One templated launch per distinct particle number over the items order[off[n] .. off[n + 1]).
static int k2_run(bp_context *h, const double *dU, int m, const unsigned char *dS, const unsigned char *dT, const K2Meta *meta,
                  const int *order, const int off[BP_MAX_N + 2], double *d_out) {
    int rc;
    for (int n = 0; n <= BP_MAX_N; ++n) {
        const int count = off[n + 1] - off[n];
        if (count <= 0) continue;
        if (n == 0) {
            k2_empty_kernel<<<(count + 255) / 256, 256, 0, h->stream>>>(order, off[n], count, d_out);
            BP_CHECK_LAUNCH(h);
            continue;
        }
        // chunk blocks per item: enough blocks to fill the GPU a few times over; ~512 terms per thread when the batch fills
        // the GPU anyway (every block pays its setup), down to ~64 when a few items must spread over all SMs
        const double worst = ldexp(1.0, n - 1);
        long long by_fill = ((long long)h->sm_count * 12 + count - 1) / count;
        const double per_thread = (count >= h->sm_count * 2) ? 512.0 : 64.0;
        long long by_work = (long long)(worst / (GW_THREADS * per_thread));
        long long chunks = by_fill < by_work ? by_fill : by_work;
        if (chunks < 1) chunks = 1;
        if (chunks > 4096) chunks = 4096;
        double *d_partials = nullptr;
        if (chunks > 1) {
            if ((rc = bp_reserve(h, BP_SLOT_PARTIALS, sizeof(double) * 4 * (size_t)count * (size_t)chunks))) return rc;
            d_partials = (double *)h->d_buf[BP_SLOT_PARTIALS];
        }
        const size_t smem = k2_smem_bytes(n);
        if (smem > 20 * 1024) {   // static shared memory (step tables, item) takes ~18 KB of the 48 KB that need no opt-in
            cudaError_t e = cudaFuncSetAttribute((const void *)g_k2_fn[n], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return bp_fail(h, BP_ERR_CUDA, "cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
        }
        dim3 grid((unsigned)count, (unsigned)chunks);
        g_k2_fn[n]<<<grid, GW_THREADS, smem, h->stream>>>(dU, m, dS, dT, meta, order, off[n], d_partials, d_out);
        BP_CHECK_LAUNCH(h);
        if (chunks > 1) {
            k2_finish_kernel<<<(count + 127) / 128, 128, 0, h->stream>>>(order, off[n], count, (int)chunks, n, d_partials, d_out);
            BP_CHECK_LAUNCH(h);
        }
    }
    return BP_OK;
}

static void k2_register() {
    static const bool ready = [] { k2_entry<BP_MAX_N>(g_k2_fn); return true; }();   // thread-safe one-time registration
    (void)ready;
}

// Host-pointer batch of at most K2_HOST_PREP_MAX items (a lone compute_permanent() above all): matrix, occupations, item
// metadata and order travel in ONE staged upload, the results in one download; no prep kernels, no synchronisation but the last.
int bp_k2_small_host(bp_context *h, const double *U, int m, const unsigned char *S, const unsigned char *T, long long B, double *out) {
    k2_register();
    auto up16 = [](size_t x) { return (x + 15) / 16 * 16; };
    const size_t ub = sizeof(double) * 2 * (size_t)m * m, sb = (size_t)B * m;
    const size_t o_S = ub, o_T = o_S + up16(sb), o_meta = o_T + up16(sb), o_order = o_meta + up16(sizeof(K2Meta) * (size_t)B);
    const size_t in_bytes = o_order + up16(sizeof(int) * (size_t)B), out_bytes = sizeof(double) * 2 * (size_t)B;
    int rc;
    if ((rc = bp_reserve(h, BP_SLOT_ITEMS, in_bytes + out_bytes))) return rc;
    if ((rc = bp_reserve_pinned(h, in_bytes + out_bytes))) return rc;
    // (every entry point ends with a stream synchronisation: the staging buffer is never in flight here)
    char *hp = (char *)h->h_pin, *dp = (char *)h->d_buf[BP_SLOT_ITEMS];
    memcpy(hp, U, ub);
    memcpy(hp + o_S, S, sb);
    memcpy(hp + o_T, T, sb);
    int off[BP_MAX_N + 2];
    k2_host_prep(S, T, m, B, (K2Meta *)(hp + o_meta), (int *)(hp + o_order), off);
    BP_CUDA(h, cudaMemcpyAsync(dp, hp, in_bytes, cudaMemcpyHostToDevice, h->stream));
    double *d_out = (double *)(dp + in_bytes);
    if ((rc = k2_run(h, (const double *)dp, m, (const unsigned char *)(dp + o_S), (const unsigned char *)(dp + o_T),
                     (const K2Meta *)(dp + o_meta), (const int *)(dp + o_order), off, d_out))) return rc;
    BP_CUDA(h, cudaMemcpyAsync(hp + in_bytes, d_out, out_bytes, cudaMemcpyDeviceToHost, h->stream));
    BP_CUDA(h, cudaStreamSynchronize(h->stream));
    memcpy(out, hp + in_bytes, out_bytes);
    return BP_OK;
}

// dU, dS, dT, d_out are device pointers.  hS: the input occupations in host memory too, or NULL (device-pointer entry point).
// With the host copy the per-n item counts are known without asking the device, so nothing synchronises the stream;
// without it one small D2H copy (the per-n offsets) does.
int bp_k2_launch(bp_context *h, const double *dU, int m, const unsigned char *dS, const unsigned char *dT,
                 long long B, double *d_out, const unsigned char *hS) {
    k2_register();
    if (B <= 0) return BP_OK;
    if (B > 0x7fffffffll) return bp_fail(h, BP_ERR_UNSUPPORTED, "bp_perm_batched: B=%lld items exceeds 2^31-1", B);
    const size_t meta_bytes = sizeof(K2Meta) * (size_t)B, order_bytes = sizeof(int) * (size_t)B;
    const size_t bins_bytes = sizeof(int) * (K2_NBINS + BP_MAX_N + 2);
    int rc = bp_reserve(h, BP_SLOT_ITEMS, meta_bytes + order_bytes + bins_bytes + 64);
    if (rc) return rc;
    char *base = (char *)h->d_buf[BP_SLOT_ITEMS];
    K2Meta *meta = (K2Meta *)base;
    int *order = (int *)(base + ((meta_bytes + 15) / 16) * 16);
    int *bins = (int *)((char *)order + ((order_bytes + 15) / 16) * 16);
    int *n_offsets = bins + K2_NBINS;
    int off[BP_MAX_N + 2];
    BP_CUDA(h, cudaMemsetAsync(bins, 0, bins_bytes, h->stream));
    const int tb = 256, gb = (int)((B + tb - 1) / tb);
    k2_prep_kernel<<<gb, tb, 0, h->stream>>>(dS, dT, m, B, meta, bins);
    BP_CHECK_LAUNCH(h);
    k2_scan_kernel<<<1, 1024, 0, h->stream>>>(bins, n_offsets);
    BP_CHECK_LAUNCH(h);
    k2_scatter_kernel<<<gb, tb, 0, h->stream>>>(meta, B, bins, order, d_out);
    BP_CHECK_LAUNCH(h);
    if (hS) {
        // item counts per particle number from the host copy: no round trip to the device
        int counts[BP_MAX_N + 2] = {0};
        for (long long b = 0; b < B; ++b) {
            int ns = 0;
            const unsigned char *s = hS + b * m;
            for (int v = 0; v < m; ++v) ns += s[v];
            counts[ns]++;
        }
        off[0] = 0;
        for (int n = 0; n <= BP_MAX_N; ++n) off[n + 1] = off[n] + counts[n];
    } else {
        if ((rc = bp_reserve_pinned(h, 4096))) return rc;
        int *h_off = (int *)h->h_pin;
        BP_CUDA(h, cudaMemcpyAsync(h_off, n_offsets, sizeof(int) * (BP_MAX_N + 2), cudaMemcpyDeviceToHost, h->stream));
        BP_CUDA(h, cudaStreamSynchronize(h->stream));
        for (int i = 0; i < BP_MAX_N + 2; ++i) off[i] = h_off[i];
    }
    return k2_run(h, dU, m, dS, dT, meta, order, off, d_out);
}
