// guan_walker.cuh -- mixed-radix reflected Gray ("Guan") code walker shared by K2 and K3.
//
// The reference iterates Guan codes sequentially from r = 0
// (theboss/boson_sampling_utilities/permanent_calculators/bs_permanent_calculator_base.py:123-149,
//  binomial product :151-164).  On the GPU every thread owns a contiguous range of term indices
// and therefore needs (a) random access: term index -> digit vector, directions and binomial
// product and (b) the same single-digit +-1 step as the reference.
//
// Chin-Huh symmetry: the term of r equals the term of (w - r), so the LAST digit is only walked
// over [0, floor(w_top/2)] and weighted 2 (1 on the self-paired middle plane when w_top is even).
// For collision-free input this is exactly Glynn's "one delta fixed" 2^(n-1) space.
#pragma once
#include "bp_common.cuh"

#define GW_THREADS 128

struct GuanItem {                 // block-uniform description of one walk, lives in shared memory
    int D;                        // number of digits (distinct occupied modes on the walk side)
    int n;                        // particles
    unsigned long long terms;     // prod (lim + 1)
    unsigned char lim[BP_MAX_N];  // largest value of digit v
    unsigned char mult[BP_MAX_N]; // multiplicity w_v of digit v (binomial C(w_v, r_v), coefficient w_v - 2 r_v)
    short mode[BP_MAX_N];         // mode index of digit v
};

// Orders the digits of an occupation vector: all occupied modes in mode order, except that the digit chosen as "top" (an odd
// multiplicity if there is one, else the largest; the first candidate wins ties) is moved last and -- inner_first -- the largest
// multiplicity among the others becomes digit 0 (the fastest digit of the step tables).  Built by ONE WARP (all 32 lanes of
// the calling warp must enter; `occ` may live in global memory): the occupied modes are found 32 at a time with a ballot, the
// top / inner digits by warp arg-max, the halved term count by a product reduction -- ~m/32 rounds of the warp where one thread
// looping over m modes was half of the instructions of a small sampling step.
__device__ inline void guan_item_build_warp(GuanItem &it, const unsigned char *occ, int m, bool inner_first = false) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    int D = 0, n = 0;
    // key of the "top" candidate: odd multiplicity first, then the larger one, then the EARLIER digit
    int best_key = -1;
    for (int base = 0; base < m; base += 32) {
        const int v = base + lane;
        const int w = (v < m) ? (int)occ[v] : 0;
        const unsigned mask = __ballot_sync(0xffffffffu, w > 0);
        const int d = D + __popc(mask & lt);
        if (w > 0) {
            if (d < BP_MAX_N) { it.mode[d] = (short)v; it.mult[d] = (unsigned char)w; }
            const int key = ((((w & 1) << 8) | w) << 8) | (255 - (d < 255 ? d : 255));
            best_key = key > best_key ? key : best_key;
        }
        D += __popc(mask);
        n += w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n += __shfl_xor_sync(0xffffffffu, n, o);
        const int other = __shfl_xor_sync(0xffffffffu, best_key, o);
        best_key = other > best_key ? other : best_key;
    }
    if (lane == 0) { it.D = D; it.n = n; }
    if (D == 0 || D > BP_MAX_N) { if (lane == 0) it.terms = (D == 0) ? 1ull : 0ull; __syncwarp(); return; }
    __syncwarp();
    const int top = 255 - (best_key & 255);
    // move the top digit to the end (digits top+1 .. D-1 shift down by one); D <= 40: two rounds of the warp
    short tm = it.mode[top]; unsigned char tw = it.mult[top];
    short mv[2]; unsigned char wv[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int v = lane + 32 * q;
        if (v >= top && v + 1 < D) { mv[q] = it.mode[v + 1]; wv[q] = it.mult[v + 1]; }
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int v = lane + 32 * q;
        if (v >= top && v + 1 < D) { it.mode[v] = mv[q]; it.mult[v] = wv[q]; }
    }
    if (lane == 0) { it.mode[D - 1] = tm; it.mult[D - 1] = tw; }
    __syncwarp();
    if (inner_first && D > 2) {
        // digit 0 becomes the inner loop: the largest multiplicity among digits 0 .. D-2 (the first one on ties)
        int key = -1;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int v = lane + 32 * q;
            if (v < D - 1) { const int kq = ((int)it.mult[v] << 8) | (255 - v); key = kq > key ? kq : key; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const int other = __shfl_xor_sync(0xffffffffu, key, o); key = other > key ? other : key; }
        const int best = 255 - (key & 255);
        if (lane == 0 && best != 0) {
            const short bm = it.mode[best]; const unsigned char bw = it.mult[best];
            it.mode[best] = it.mode[0]; it.mult[best] = it.mult[0];
            it.mode[0] = bm; it.mult[0] = bw;
        }
        __syncwarp();
    }
    unsigned long long terms = 1ull;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int v = lane + 32 * q;
        if (v < D) {
            const unsigned char l = (v == D - 1) ? (unsigned char)(it.mult[v] >> 1) : it.mult[v];
            it.lim[v] = l;
            terms *= (unsigned long long)(l + 1);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) terms *= __shfl_xor_sync(0xffffffffu, terms, o);
    if (lane == 0) it.terms = terms;
    __syncwarp();
}


// Expands an occupation vector into its particle list by ONE WARP: col_mode[c] = mode of particle c (mode order), entries beyond
// the particles up to `width` are -1.  32 modes per round, positions from a warp scan.
__device__ inline void guan_expand_columns_warp(short *col_mode, const unsigned char *occ, int m, int width) {
    const int lane = threadIdx.x & 31;
    int carry = 0;
    for (int base = 0; base < m; base += 32) {
        const int v = base + lane;
        const int w = (v < m) ? (int)occ[v] : 0;
        int x = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        const int start = carry + x - w;
        for (int a = 0; a < w; ++a)
            if (start + a < width) col_mode[start + a] = (short)v;
        carry += __shfl_sync(0xffffffffu, x, 31);
    }
    for (int c = carry + lane; c < width; c += 32) col_mode[c] = -1;
    __syncwarp();
}

// Term count of the halved walk without building the item (cost model of the scheduler).
__host__ __device__ inline double guan_terms_of(const unsigned char *occ, int m) {
    double full = 1.0;
    int best_w = 0;
    bool any = false;
    for (int v = 0; v < m; ++v) {
        const int w = occ[v];
        if (!w) continue;
        full *= (double)(w + 1);
        const bool better = !any || ((w & 1) && !(best_w & 1)) || (((w & 1) == (best_w & 1)) && w > best_w);
        if (better) { best_w = w; any = true; }
    }
    if (!any) return 1.0;
    return full / (double)(best_w + 1) * (double)((best_w >> 1) + 1);
}

// 1 / q for q = 1 .. 41 (entry 0 unused): reciprocals of the step denominators of guan_step
static __constant__ double gw_rcp[BP_MAX_N + 2] = {
    0.0, 1.0 / 1, 1.0 / 2, 1.0 / 3, 1.0 / 4, 1.0 / 5, 1.0 / 6, 1.0 / 7, 1.0 / 8, 1.0 / 9, 1.0 / 10, 1.0 / 11, 1.0 / 12, 1.0 / 13, 1.0 / 14,
    1.0 / 15, 1.0 / 16, 1.0 / 17, 1.0 / 18, 1.0 / 19, 1.0 / 20, 1.0 / 21, 1.0 / 22, 1.0 / 23, 1.0 / 24, 1.0 / 25, 1.0 / 26, 1.0 / 27,
    1.0 / 28, 1.0 / 29, 1.0 / 30, 1.0 / 31, 1.0 / 32, 1.0 / 33, 1.0 / 34, 1.0 / 35, 1.0 / 36, 1.0 / 37, 1.0 / 38, 1.0 / 39, 1.0 / 40, 1.0 / 41};

// d = q mod R, q = q div R for a digit radix R = lim + 1 (block-uniform).  Collision-free digits (R = 2, the usual case) and
// single-valued ones (R = 1) take no division: the position decodes of the block setup were mostly 32-bit divisions.
__device__ __forceinline__ unsigned gw_divmod(unsigned &q, unsigned R) {
    unsigned d;
    if (R == 2u)      { d = q & 1u; q >>= 1; }
    else if (R == 1u) { d = 0u; }
    else              { d = q % R; q /= R; }
    return d;
}

struct GuanState {                // per-thread
    unsigned long long dirmask;   // bit v set: digit v currently moves downwards
    double binom;                 // prod_v C(w_v, r_v) * (top weight 1 or 2), exact integer
};

__device__ __forceinline__ double gw_binom(int w, int r) {   // exact for w <= 40
    if (r > w - r) r = w - r;
    double c = 1.0;
    for (int q = 1; q <= r; ++q) c = rint(c * (double)(w - q + 1) / (double)q);
    return c;
}

__device__ __forceinline__ double gw_top_weight(const GuanItem &it, int r_top) {
    return (2 * r_top < (int)it.mult[it.D - 1]) ? 2.0 : 1.0;
}

// Random access (Appendix A.6): digits r[v * GW_THREADS] (caller passes its own column), directions
// and binomial product of term index I.
// With v0 > 0 the walk runs over digits v0 .. D-1 only (I indexes that sub-walk).
// STRIDE = threads per block of the caller (the digit vectors are stored column-per-thread).
template <int STRIDE = GW_THREADS>
__device__ inline void guan_seek(const GuanItem &it, unsigned long long I, unsigned char *r, GuanState &st, int v0 = 0) {
    unsigned long long q = I;
    st.dirmask = 0ull;
    double b = 1.0;
    for (int v = v0; v < it.D; ++v) {
        const unsigned R = (unsigned)it.lim[v] + 1u;
        unsigned d;
        if (q >> 32) { d = (unsigned)(q % R); q /= R; }                                   // 64-bit division: ~10x the cost, rare
        else         { unsigned q32 = (unsigned)q; d = gw_divmod(q32, R); q = q32; }
        int rv;
        if (q & 1ull) { rv = (int)it.lim[v] - (int)d; st.dirmask |= (1ull << v); }
        else          { rv = (int)d; }
        r[v * STRIDE] = (unsigned char)rv;
        if (it.mult[v] > 1) b *= gw_binom(it.mult[v], rv);
    }
    st.binom = (it.D - 1 >= v0) ? b * gw_top_weight(it, r[(it.D - 1) * STRIDE]) : b;
}

// One Guan step.  Returns the digit that changed; `delta` = +1 / -1.  Must not be called on the
// last term of the walk.  __ldg-free: everything is in shared memory / registers.
template <int STRIDE = GW_THREADS>
__device__ __forceinline__ int guan_step(const GuanItem &it, unsigned char *r, GuanState &st, int &delta, int v0 = 0) {
    int v = v0;
    int cur, nxt, dir;
    for (;;) {
        cur = r[v * STRIDE];
        dir = ((st.dirmask >> v) & 1ull) ? -1 : 1;
        nxt = cur + dir;
        if (nxt >= 0 && nxt <= (int)it.lim[v]) break;
        st.dirmask ^= (1ull << v);
        ++v;
    }
    r[v * STRIDE] = (unsigned char)nxt;
    delta = dir;
    const int w = it.mult[v];
    if (w > 1) {
        // C(w, nxt) from C(w, cur): exact integer arithmetic in doubles (values < 2^46).  The quotient is an integer, so the
        // division is a multiplication by the rounded reciprocal followed by rint (error < 2^46 * 2^-52: the nearest integer
        // is the exact quotient) -- no FP64 division sequence at the period boundaries of the term loops.
        double b = st.binom;
        if (v == it.D - 1) b *= (2 * cur < w) ? 0.5 : 1.0;   // divide by the top weight (2 or 1): exact
        if (dir > 0) b = rint(b * (double)(w - cur) * gw_rcp[nxt]);
        else         b = rint(b * (double)cur * gw_rcp[w - nxt]);
        if (v == it.D - 1) b *= gw_top_weight(it, nxt);
        st.binom = b;
    }
    return v;
}
