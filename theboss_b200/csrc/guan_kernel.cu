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

// exclusive scan of the 41*64 bins (one block) + per-n offsets
__global__ void k2_scan_kernel(int *__restrict__ bins, int *__restrict__ n_offsets /*[42]*/) {
    if (threadIdx.x == 0) {
        int run = 0;
        for (int n = 0; n <= BP_MAX_N; ++n) {
            n_offsets[n] = run;
            for (int c = 0; c < 64; ++c) { const int v = bins[n * 64 + c]; bins[n * 64 + c] = run; run += v; }
        }
        n_offsets[BP_MAX_N + 1] = run;
    }
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
struct __align__(16) K2Step { double blow; int off; int pad; };   // binomial product at the position; row byte offset | went-up bit

// 16-byte shared-memory load at a 32-bit shared-window address plus a compile-time byte offset
// (PIN: volatile, so that loads of a loop-invariant row stay inside the term loop instead of occupying registers)
template <int OFF, bool PIN>
__device__ __forceinline__ double2 k2_lds(unsigned addr) {
    double2 v;
    if (PIN) asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "r"(addr), "n"(OFF));
    else     asm("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "r"(addr), "n"(OFF));
    return v;
}
// sums[j] += sg * X2row[j], j = 0 .. N-1 (row given by its shared-window address: one register + immediates)
template <int N, bool PIN, int J = 0>
__device__ __forceinline__ void k2_row_update(unsigned row, double sg, double (&sr)[N], double (&si)[N]) {
    if constexpr (J < N) {
        const double2 a = k2_lds<J * (int)sizeof(double2), PIN>(row);
        sr[J] = fma(sg, a.x, sr[J]);
        si[J] = fma(sg, a.y, si[J]);
        k2_row_update<N, PIN, J + 1>(row, sg, sr, si);
    }
}
#ifndef K2_REGROW_MAX_N
#define K2_REGROW_MAX_N 12   // up to this N the row of the inner digit is kept in registers (4N extra registers); measured at N=20: 3 warps/SMSP with LDS rows beat 2 warps with register rows (11.8 vs 12.3 ms on config 2)
#endif

template <int N>
struct K2Cfg {
    static constexpr bool REGROW = (N <= K2_REGROW_MAX_N);
#ifndef K2_MINB_MID
#define K2_MINB_MID 3
#endif
#ifndef K2_NCH_MID
#define K2_NCH_MID 1
#endif
    static constexpr int MINB = (N <= 8) ? 4 : (N <= 20) ? K2_MINB_MID : (N <= 26) ? 3 : 2;
};

template <int N>
__device__ __forceinline__ void k2_product(const double (&sr)[N], const double (&si)[N], double &pr, double &pi) {
    constexpr int NCH = (N > 20) ? 4 : (N > K2_REGROW_MAX_N) ? K2_NCH_MID : (N >= 4 ? 2 : 1);   // fewer chains where the inner row occupies registers
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
    pr = p[0].re; pi = p[0].im;
}


// The walk of one item is organised like the minors kernel's (minors_kernel.cu): rows of L0 + 1 terms that
// differ only in the inner digit 0 (largest multiplicity; swept up on even rows, down on odd ones), rows
// indexed by the sub-walk over digits 1 .. D-1; the low digits of that sub-walk follow a per-block table
// (period P), the generic Guan stepper only carries into the digits above it.  Threads own whole periods.
template <int N>
__global__ void __launch_bounds__(GW_THREADS, K2Cfg<N>::MINB)
k2_perm_kernel(const double *__restrict__ U, int m, const unsigned char *__restrict__ S,
               const unsigned char *__restrict__ T, const K2Meta *__restrict__ meta,
               const int *__restrict__ order, int first, double *__restrict__ partials) {
    constexpr bool REGROW = K2Cfg<N>::REGROW;
    __shared__ GuanItem item;
    __shared__ short col_mode[N];
    __shared__ double2 X2[N * N];                       // 2 * X[v][j], D <= N rows
    __shared__ unsigned char rdig[N * GW_THREADS];      // per-thread digit vectors (column = thread)
    __shared__ double red[4 * (GW_THREADS / 32)];
    __shared__ double bin0[BP_MAX_N + 2];
    // step tables of the low digits, indexed by the DESTINATION position inside a period (see minors_kernel.cu)
    __shared__ K2Step fwd[K2_PMAX], bwd[K2_PMAX];
    __shared__ int low_digits;
    __shared__ unsigned period;

    const int b = order[first + blockIdx.x];
    const K2Meta me = meta[b];   // (blockIdx.y = chunk of this item's walk)
    const unsigned char *walk = (me.walk_outputs ? T : S) + (long long)b * m;
    const unsigned char *prod = (me.walk_outputs ? S : T) + (long long)b * m;
    if (threadIdx.x == 0) {
        guan_item_build(item, walk, m, /*inner_first=*/true);
        int c = 0;
        for (int v = 0; v < m; ++v)
            for (int a = 0; a < prod[v] && c < N; ++a) col_mode[c++] = (short)v;
        for (int r = 0; r <= (int)item.lim[0]; ++r)
            bin0[r] = gw_binom(item.mult[0], r) * (item.D == 1 ? gw_top_weight(item, r) : 1.0);
    }
    __syncthreads();
    const int D = item.D;
    const double2 *U2 = reinterpret_cast<const double2 *>(U);
    for (int e = threadIdx.x; e < D * N; e += GW_THREADS) {
        const int v = e / N, j = e - v * N;
        const int wm = item.mode[v], pm = col_mode[j];
        // U[out_mode][in_mode]: walking outputs -> rows of U, walking inputs -> columns of U
        const double2 u = me.walk_outputs ? U2[wm * m + pm] : U2[pm * m + wm];
        X2[e] = make_double2(2.0 * u.x, 2.0 * u.y);
    }
    const int L0 = item.lim[0];
    const unsigned long long rows = item.terms / (unsigned long long)(L0 + 1);
    const unsigned long long nthreads = (unsigned long long)gridDim.y * GW_THREADS;
    if (threadIdx.x == 0) {
        unsigned long long raw = (rows + nthreads - 1) / nthreads, P = 1;
        int a = 0;
        for (int v = 1; v < D; ++v) {
            const unsigned long long nxt = P * (unsigned long long)(item.lim[v] + 1);
            if (nxt * 16 > raw || nxt > K2_PMAX) break;
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
        for (int v = 1; v <= a_low; ++v) {
            const unsigned R = (unsigned)item.lim[v] + 1u;
            unsigned d = q % R; q /= R;
            unsigned dm = qm % R; qm /= R;
            const int rv = (q & 1u) ? (int)item.lim[v] - (int)d : (int)d;
            const int rm = (qm & 1u) ? (int)item.lim[v] - (int)dm : (int)dm;
            if (p && rv != rm) { chg = v; up = rv > rm; }
            double c = gw_binom(item.mult[v], rv);
            if (v == D - 1) c *= gw_top_weight(item, rv);
            bprod *= c;
        }
        const int rowbytes = N * (int)sizeof(double2);
        fwd[p].blow = bprod; fwd[p].off = chg * rowbytes | up; fwd[p].pad = 0;
        bwd[p].blow = bprod;
        if (p == P - 1) { bwd[p].off = 0; bwd[p].pad = 0; }
        if (p) { bwd[p - 1].off = chg * rowbytes | (up ^ 1); bwd[p - 1].pad = 0; }
    }
    __syncthreads();

    // whole periods are dealt out to the threads of all chunk blocks of this item
    const unsigned long long nper = rows / P;
    const unsigned long long tidx = (unsigned long long)blockIdx.y * GW_THREADS + threadIdx.x;
    const unsigned long long pbase = nper / nthreads, prem = nper % nthreads;
    const unsigned long long my_periods = pbase + (tidx < prem ? 1ull : 0ull);
    const unsigned long long row_start = (tidx * pbase + (tidx < prem ? tidx : prem)) * P;
    dd acc_re = {0.0, 0.0}, acc_im = {0.0, 0.0};

    if (my_periods > 0) {
        const unsigned my_rows = (unsigned)(my_periods * P);   // per-thread row counts fit 32 bits (<= 2^39 terms over >= 128 threads)
        unsigned char *r = rdig + threadIdx.x;
        GuanState st;
        const unsigned long long hi0 = row_start / P;
        guan_seek(item, hi0, r, st, /*v0=*/a_low + 1);
        int pos = (hi0 & 1ull) ? (int)P - 1 : 0;
        int pdir = (hi0 & 1ull) ? -1 : 1;
        unsigned off = 0;
        int r0 = (row_start & 1ull) ? L0 : 0;
        int dir0 = (row_start & 1ull) ? -1 : 1;
        double sr[N], si[N];
        double x0r[REGROW ? N : 1], x0i[REGROW ? N : 1];
#pragma unroll
        for (int j = 0; j < N; ++j) { sr[j] = 0.0; si[j] = 0.0; }
        int par = r0;
        {
            unsigned q = (unsigned)pos;
#pragma unroll 1
            for (int v = 0; v < D; ++v) {
                int rv;
                if (v == 0) rv = r0;
                else if (v <= a_low) {
                    const unsigned R = (unsigned)item.lim[v] + 1u;
                    const unsigned d = q % R; q /= R;
                    rv = (q & 1u) ? (int)item.lim[v] - (int)d : (int)d;
                } else rv = (int)r[v * GW_THREADS];
                if (v > 0) par += rv;
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
        if (REGROW) {
#pragma unroll
            for (int j = 0; j < N; ++j) { const double2 a = X2[j]; x0r[j] = a.x; x0i[j] = a.y; }
        }
        double sgn = (par & 1) ? -1.0 : 1.0;
        double bout = st.binom * fwd[pos].blow;
        const K2Step *tab = (pdir > 0) ? fwd : bwd;
        // shared-window address of X2, kept opaque so that row loads use ONE address register plus immediates
        unsigned x2base = (unsigned)__cvta_generic_to_shared(X2);
        asm volatile("" : "+r"(x2base));
        double w0 = bin0[r0];                                // weight of digit 0, fetched one term ahead
        double wr = 0.0, wi = 0.0;
        unsigned cnt = 0;

#pragma unroll 1
        for (unsigned q = 0;;) {
            // ---- plan the step to the next row now, so that its table / digit loads overlap the sweep below
            const bool have_next = q + 1 < my_rows;
            int off_next = 0;
            double bout_next = 0.0;
            if (have_next) {
                if (++off < P) {
                    pos += pdir;
                    const K2Step e = tab[pos];
                    off_next = e.off;
                    bout_next = st.binom * e.blow;
                } else {
                    int delta;
                    const int v = guan_step(item, r, st, delta, /*v0=*/a_low + 1);
                    off_next = v * (N * (int)sizeof(double2)) | (delta > 0 ? 1 : 0);
                    off = 0;
                    pdir = -pdir;
                    tab = (pdir > 0) ? fwd : bwd;
                    bout_next = st.binom * fwd[pos].blow;
                }
            }
            // ---- inner sweep over digit 0
#pragma unroll 1
            for (int step = 0;; ++step) {
                const double w = sgn * bout * w0;
                double pr, pi;
                k2_product<N>(sr, si, pr, pi);
                wr = fma(w, pr, wr);
                wi = fma(w, pi, wi);
                if (++cnt == 64u) {
                    cnt = 0;
                    acc_re = dd_add_d(acc_re, wr); acc_im = dd_add_d(acc_im, wi);
                    wr = 0.0; wi = 0.0;
                }
                const bool last = (step == L0);
                if (!last) r0 += dir0;
                w0 = bin0[r0];
                if (last) break;
                sgn = -sgn;
                const double sg = (dir0 > 0) ? -1.0 : 1.0;        // sums -= 2 * dir0 * X[0]
                if (REGROW) {
#pragma unroll
                    for (int j = 0; j < N; ++j) { sr[j] = fma(sg, x0r[j], sr[j]); si[j] = fma(sg, x0i[j], si[j]); }
                } else {
                    k2_row_update<N, true>(x2base, sg, sr, si);
                }
            }
            dir0 = -dir0;
            // ---- next row
            if (!have_next) break;
            ++q;
            sgn = -sgn;
            bout = bout_next;
            k2_row_update<N, false>(x2base + (unsigned)(off_next & ~1), (off_next & 1) ? -1.0 : 1.0, sr, si);   // sums -= 2 * delta * X[v]
        }
        acc_re = dd_add_d(acc_re, wr);
        acc_im = dd_add_d(acc_im, wi);
    }
    block_reduce_dd(acc_re, acc_im, red);
    if (threadIdx.x == 0) {
        double *o = partials + 4 * ((size_t)blockIdx.x * gridDim.y + blockIdx.y);
        o[0] = acc_re.hi; o[1] = acc_re.lo; o[2] = acc_im.hi; o[3] = acc_im.lo;
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

typedef void (*k2_fn)(const double *, int, const unsigned char *, const unsigned char *, const K2Meta *, const int *, int, double *);
template <int N>
static void k2_entry(k2_fn *fn) {
    fn[N] = k2_perm_kernel<N>;
    if constexpr (N > 1) k2_entry<N - 1>(fn);
}
static k2_fn g_k2_fn[BP_MAX_N + 1];

// All pointers are device pointers.  Enqueues prep + sort + one launch per distinct n; needs one
// small D2H copy (per-n offsets) in the middle, so it synchronises the stream once.
int bp_k2_launch(bp_context *h, const double *dU, int m, const unsigned char *dS, const unsigned char *dT,
                 long long B, double *d_out) {
    static const bool ready = [] { k2_entry<BP_MAX_N>(g_k2_fn); return true; }();   // thread-safe one-time registration
    (void)ready;
    if (B <= 0) return BP_OK;
    if (B > 0x7fffffffll) return bp_fail(h, BP_ERR_UNSUPPORTED, "bp_perm_batched: B=%lld items exceeds 2^31-1", B);
    const size_t meta_bytes = sizeof(K2Meta) * (size_t)B, order_bytes = sizeof(int) * (size_t)B;
    const size_t bins_bytes = sizeof(int) * ((BP_MAX_N + 1) * 64 + BP_MAX_N + 2);
    int rc = bp_reserve(h, BP_SLOT_ITEMS, meta_bytes + order_bytes + bins_bytes + 64);
    if (rc) return rc;
    if ((rc = bp_reserve_pinned(h, 4096))) return rc;
    char *base = (char *)h->d_buf[BP_SLOT_ITEMS];
    K2Meta *meta = (K2Meta *)base;
    int *order = (int *)(base + ((meta_bytes + 15) / 16) * 16);
    int *bins = (int *)((char *)order + ((order_bytes + 15) / 16) * 16);
    int *n_offsets = bins + (BP_MAX_N + 1) * 64;
    BP_CUDA(h, cudaMemsetAsync(bins, 0, bins_bytes, h->stream));
    const int tb = 256, gb = (int)((B + tb - 1) / tb);
    k2_prep_kernel<<<gb, tb, 0, h->stream>>>(dS, dT, m, B, meta, bins);
    BP_CHECK_LAUNCH(h);
    k2_scan_kernel<<<1, 32, 0, h->stream>>>(bins, n_offsets);
    BP_CHECK_LAUNCH(h);
    k2_scatter_kernel<<<gb, tb, 0, h->stream>>>(meta, B, bins, order, d_out);
    BP_CHECK_LAUNCH(h);
    int *h_off = (int *)h->h_pin;
    BP_CUDA(h, cudaMemcpyAsync(h_off, n_offsets, sizeof(int) * (BP_MAX_N + 2), cudaMemcpyDeviceToHost, h->stream));
    BP_CUDA(h, cudaStreamSynchronize(h->stream));
    int off[BP_MAX_N + 2];
    for (int i = 0; i < BP_MAX_N + 2; ++i) off[i] = h_off[i];
    for (int n = 0; n <= BP_MAX_N; ++n) {
        const int count = off[n + 1] - off[n];
        if (count <= 0) continue;
        if (n == 0) {
            k2_empty_kernel<<<(count + 255) / 256, 256, 0, h->stream>>>(order, off[n], count, d_out);
            BP_CHECK_LAUNCH(h);
            continue;
        }
        // chunks per item: enough blocks to fill the GPU a few times over, but at least ~256 terms per thread
        long long by_fill = ((long long)h->sm_count * 12 + count - 1) / count;
        long long by_work = (long long)(ldexp(1.0, n - 1) / (GW_THREADS * 512.0));
        long long chunks = by_fill < by_work ? by_fill : by_work;
        if (chunks < 1) chunks = 1;
        if (chunks > 4096) chunks = 4096;
        if ((rc = bp_reserve(h, BP_SLOT_PARTIALS, sizeof(double) * 4 * (size_t)count * (size_t)chunks))) return rc;
        double *d_partials = (double *)h->d_buf[BP_SLOT_PARTIALS];
        dim3 grid((unsigned)count, (unsigned)chunks);
        g_k2_fn[n]<<<grid, GW_THREADS, 0, h->stream>>>(dU, m, dS, dT, meta, order, off[n], d_partials);
        BP_CHECK_LAUNCH(h);
        k2_finish_kernel<<<(count + 127) / 128, 128, 0, h->stream>>>(order, off[n], count, (int)chunks, n, d_partials, d_out);
        BP_CHECK_LAUNCH(h);
    }
    return BP_OK;
}
