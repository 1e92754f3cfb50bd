// minors_kernel.cu -- K3: all one-input-particle-removed permanents of a GCC-B step, batched over
// samples, plus the finish kernel (chunk reduction, Laplace combine into the step's pmf, and the
// categorical draw / state update of the sampling loop).
//
// Replaces BSCC{Ryser,CH}SubmatricesPermanentCalculator.compute_permanents
// (reference: theboss/boson_sampling_utilities/permanent_calculators/
//  bs_submatrices_permanent_calculator_base.py:150-189, bs_cc_ryser_submatrices_permanent_calculator.py:73-119,
//  bs_cc_ch_submatrices_permanent_calculator.py:51-105) and the consumer
// GeneralizedCliffordsBSimulationStrategy._compute_pmf / _sample_from_pmf / _update_current_input
// (theboss/simulation_strategies/generalized_cliffords_b_simulation_strategy.py:69-110).
//
// Layout (perm A = perm A^T): the Guan walk runs over the k-1 already
// sampled OUTPUT particles (multiplicities t, r <-> t - r symmetry halved), the product runs over
// the k INPUT particles (columns of U repeated by s):
//     c_col(rho) = sum_j (t_j - 2 rho_j) U[j][mode(col)]
//     P_col      = 2^-(k-1) sum_rho (-1)^{sum rho} prod_j C(t_j, rho_j) prod_{col' != col} c_col'
// All k leave-one-out products of a term come from a balanced product tree: ~3k complex multiplies per
// term instead of k^2.
//
// Thread mapping: LPG (1, 2 or 4) adjacent lanes form a group that walks one contiguous range of terms; each
// lane owns C columns (LPG * C >= k, padding columns are the constant 1).  A lane keeps c[C], the tree nodes
// and its C accumulators in registers; the product of the OTHER lanes' column products reaches it through an
// xor-butterfly of warp shuffles and seeds its downward pass, so splitting costs only (log2 LPG) complex
// multiplies per term.
//
// What bounds the term loop (measured, scripts/rf_probe.cu -> profiles/r02_rf_probe.txt): an SMSP reads ONE 64-bit
// vector-register operand per cycle, so a DFMA whose three sources are distinct registers costs 3 cycles where the FP64
// pipe needs 2; only operands repeated from the previous instruction (operand reuse cache) are free.  The loop below is a
// DFMA / DMUL / DADD mix of 90 / 45 / 25 instructions per term at C = 12, of which ptxas pairs 37 DFMAs: 380 cycles per
// term against 320 of pure pipe time (scripts/sass_rf.py) -- exactly the measured rate.  More warps per SMSP do not help
// (profiles/r02_k3_history.txt); fewer distinct operands per instruction would.
#include "bp_common.cuh"
#include "guan_walker.cuh"
#include "minors.cuh"

#include <stdlib.h>

// xor-shuffle of a complex value; callers keep the trip counts warp-uniform and pass the full mask
__device__ __forceinline__ cplx cshfl_xor(unsigned gmask, cplx a, int mask) {
    cplx r;
    r.re = __shfl_xor_sync(gmask, a.re, mask);
    r.im = __shfl_xor_sync(gmask, a.im, mask);
    return r;
}

// 16-byte shared-memory load at a 32-bit shared-window address plus a compile-time byte offset: the C loads of a row
// use ONE address register plus immediates (the compiler otherwise keeps a pointer per column)
template <int OFF>
__device__ __forceinline__ double2 lds_f64x2(unsigned addr) {
    double2 v;
    asm("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "r"(addr), "n"(OFF));
    return v;
}

// ---------------------------------------------------------------------------------------------
// Balanced product tree.  All C leave-one-out products of a term from an upward pass (node products,
// C - 1 complex multiplies, depth ceil(log2 C)) and a downward pass (outside(child) = outside(parent) x sibling, the
// leaf level fused into the accumulators): the same ~3C complex multiplies as prefix x suffix scans, but the dependency
// depth is 2 log2 C instead of C and up to C / 2 multiplies are independent, and a lane can own up to 17 columns
// (no shuffles, no butterfly multiplies for k <= 17).
// node[MID] holds the product of the columns [LO, HI) with MID = (LO + HI) / 2 -- every internal node of the split tree
// has its own MID in 1 .. C-1, so the indices are compile-time constants and the nodes live in registers.
// ---------------------------------------------------------------------------------------------
// Complex product of the tree nodes.  The explicit form -- both fused multiply-adds take a.re as their first factor -- is the one of
// four equivalent spellings for which ptxas schedules the term loop with the fewest three-source DFMAs (scripts/sass_rf.py:
// 378 / 348 / 311 cycles per term at C = 12 / 11 / 10 against 386 / 350 / 311 for bp_common's cmul; the loop is register-read bound).
#ifndef K3_SPELL
#define K3_SPELL 0
#endif
template <int F>
__device__ __forceinline__ cplx k3_cmul_f(cplx a, cplx b) {
    cplx r;
    if constexpr ((F & 1) == 0) r.re = fma(a.re, b.re, -(a.im * b.im)); else r.re = fma(-a.im, b.im, a.re * b.re);
    if constexpr ((F & 2) == 0) r.im = fma(a.re, b.im, a.im * b.re); else r.im = fma(a.im, b.re, a.re * b.im);
    return r;
}
__device__ __forceinline__ cplx k3_cmul(cplx a, cplx b) { return k3_cmul_f<0>(a, b); }
__device__ __forceinline__ cplx k3_cmul_up(cplx a, cplx b) {
    if constexpr ((K3_SPELL >> 2) & 1) return k3_cmul_f<K3_SPELL & 3>(b, a); else return k3_cmul_f<K3_SPELL & 3>(a, b);
}
__device__ __forceinline__ cplx k3_cmul_down(cplx a, cplx b) {
    if constexpr ((K3_SPELL >> 5) & 1) return k3_cmul_f<(K3_SPELL >> 3) & 3>(b, a); else return k3_cmul_f<(K3_SPELL >> 3) & 3>(a, b);
}
__device__ __forceinline__ void k3_leaf_acc(double &acc_re, double &acc_im, cplx x0, cplx y0) {
    const cplx x = ((K3_SPELL >> 6) & 1) ? y0 : x0, y = ((K3_SPELL >> 6) & 1) ? x0 : y0;
    if constexpr (((K3_SPELL >> 7) & 3) == 0) {
        acc_re = fma(x.re, y.re, acc_re); acc_re = fma(-x.im, y.im, acc_re); acc_im = fma(x.re, y.im, acc_im); acc_im = fma(x.im, y.re, acc_im);
    } else if constexpr (((K3_SPELL >> 7) & 3) == 1) {
        acc_re = fma(x.re, y.re, acc_re); acc_im = fma(x.re, y.im, acc_im); acc_re = fma(-x.im, y.im, acc_re); acc_im = fma(x.im, y.re, acc_im);
    } else if constexpr (((K3_SPELL >> 7) & 3) == 2) {
        acc_re = fma(x.re, y.re, acc_re); acc_im = fma(y.re, x.im, acc_im); acc_re = fma(-x.im, y.im, acc_re); acc_im = fma(y.im, x.re, acc_im);
    } else {
        acc_re = fma(x.re, y.re, acc_re); acc_im = fma(x.im, y.re, acc_im); acc_im = fma(x.re, y.im, acc_im); acc_re = fma(-x.im, y.im, acc_re);
    }
}
template <int C, int LO, int HI>
__device__ __forceinline__ cplx k3_tree_val(const double (&cr)[C], const double (&ci)[C], const cplx (&node)[C]) {
    if constexpr (HI - LO == 1) { cplx v = {cr[LO], ci[LO]}; return v; }
    else return node[(LO + HI) / 2];
}
template <int C, int LO, int HI, bool ROOT>
__device__ __forceinline__ void k3_tree_up(const double (&cr)[C], const double (&ci)[C], cplx (&node)[C]) {
    if constexpr (HI - LO >= 2) {
        constexpr int MID = (LO + HI) / 2;
        k3_tree_up<C, LO, MID, true>(cr, ci, node);
        k3_tree_up<C, MID, HI, true>(cr, ci, node);
        if constexpr (ROOT) node[MID] = k3_cmul_up(k3_tree_val<C, LO, MID>(cr, ci, node), k3_tree_val<C, MID, HI>(cr, ci, node));
    }
}
template <int C, int LO, int HI>
__device__ __forceinline__ void k3_tree_down(cplx out, const double (&cr)[C], const double (&ci)[C], const cplx (&node)[C],
                                             double (&ar)[C], double (&ai)[C]) {
    if constexpr (HI - LO == 1) { ar[LO] += out.re; ai[LO] += out.im; }
    else {
        constexpr int MID = (LO + HI) / 2;
        const cplx L = k3_tree_val<C, LO, MID>(cr, ci, node), R = k3_tree_val<C, MID, HI>(cr, ci, node);
        if constexpr (((K3_SPELL >> 9) & 1) == 0) {
            if constexpr (MID - LO == 1) k3_leaf_acc(ar[LO], ai[LO], out, R);
            else k3_tree_down<C, LO, MID>(k3_cmul_down(out, R), cr, ci, node, ar, ai);
            if constexpr (HI - MID == 1) k3_leaf_acc(ar[MID], ai[MID], out, L);
            else k3_tree_down<C, MID, HI>(k3_cmul_down(out, L), cr, ci, node, ar, ai);
        } else {
            if constexpr (HI - MID == 1) k3_leaf_acc(ar[MID], ai[MID], out, L);
            else k3_tree_down<C, MID, HI>(k3_cmul_down(out, L), cr, ci, node, ar, ai);
            if constexpr (MID - LO == 1) k3_leaf_acc(ar[LO], ai[LO], out, R);
            else k3_tree_down<C, LO, MID>(k3_cmul_down(out, R), cr, ci, node, ar, ai);
        }
    }
}
// root of the downward pass when the outside factor is the real term weight w (a lane that owns every column)
template <int C>
__device__ __forceinline__ void k3_tree_down_real(double w, const double (&cr)[C], const double (&ci)[C], const cplx (&node)[C],
                                                  double (&ar)[C], double (&ai)[C]) {
    if constexpr (C == 1) { ar[0] += w; }
    else {
        constexpr int MID = C / 2;
        const cplx L = k3_tree_val<C, 0, MID>(cr, ci, node), R = k3_tree_val<C, MID, C>(cr, ci, node);
        if constexpr (MID == 1) { ar[0] = fma(w, R.re, ar[0]); ai[0] = fma(w, R.im, ai[0]); }
        else { cplx o = {w * R.re, w * R.im}; k3_tree_down<C, 0, MID>(o, cr, ci, node, ar, ai); }
        if constexpr (C - MID == 1) { ar[MID] = fma(w, L.re, ar[MID]); ai[MID] = fma(w, L.im, ai[MID]); }
        else { cplx o = {w * L.re, w * L.im}; k3_tree_down<C, MID, C>(o, cr, ci, node, ar, ai); }
    }
}

// Number of chunk blocks that actually work on a sample whose walk has `terms` terms: about
// K3_TERMS_PER_GROUP terms per lane group, at most the launched `chunks`.  Bunched outputs shrink the
// walk by orders of magnitude, so the grid is sized for the collision-free worst case and blocks beyond
// `active` exit at once; the finish kernel applies the same rule.
#ifndef K3_TERMS_PER_GROUP
#define K3_TERMS_PER_GROUP 192ull
#endif
#ifndef K3_PERIODS_PER_GROUP
#define K3_PERIODS_PER_GROUP 8    // a lane group owns at least this many table periods (group ranges differ by at most one period; 16 -> 8: n = 24 run -0.5 %, n = 16 run -3 %)
#endif
#define K3_PMAX 512
// One-warp blocks (THREADS = 32) serve the steps k <= 16, where a sample's whole walk fits one block and the
// per-block setup (item build, tables, seek: largely single-thread work) dominates: with 4x more, 4x smaller
// blocks resident per SM the setup of one block overlaps the term loops of the others instead of idling
// three of its own four warps.
#define K3_WARP_THREADS 32
#define K3_WARP_PMAX 128
struct __align__(16) K3Step { double blow; int off; int pad; };
__host__ __device__ inline int k3_active_chunks(unsigned long long terms, int chunks, unsigned long long per_block) {
    unsigned long long a = (terms + per_block - 1) / per_block;
    if (a < 1) a = 1;
    if (a > (unsigned long long)chunks) a = (unsigned long long)chunks;
    return (int)a;
}

// ---------------------------------------------------------------------------------------------
// The term loop is ONE uniform loop.  The walk of a lane group is a sequence of PERIODS of P = prod_{v <= a_low} (lim_v + 1) terms;
// inside a period every term does the same straight-line work -- fetch the 16-byte step entry of the next position, evaluate
// the term (product tree), ADD the signed row of the digit that changes (the shared-memory image holds +2U rows, -2U rows and
// a zero row, so a step is C complex additions from one address: no sign logic, no multiplication) -- and the branchy part
// (the Guan step of the digits above the table, the reversal of the table direction) runs once per period in an outer loop.
// Sentinel entries (weight 0, zero row) beyond both table ends make the last term of a period the same code as the others.
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline size_t k3_smem_bytes(int rows, int W, int C, int threads) {
    size_t x2 = (size_t)(2 * rows + 1) * W * sizeof(double2);
    size_t red = (size_t)threads * C * sizeof(double2);
    size_t dig = (size_t)rows * threads;
    return ((x2 > red ? x2 : red) + dig + 15) / 16 * 16 + 16;
}
// c[j] += row[j] for the C columns of a lane (row given by its shared-window address)
template <int C, int J = 0>
__device__ __forceinline__ void k3_row_add(unsigned row, double (&cr)[C], double (&ci)[C]) {
    if constexpr (J < C) {
        const double2 a = lds_f64x2<J * (int)sizeof(double2)>(row);
        cr[J] += a.x;
        ci[J] += a.y;
        k3_row_add<C, J + 1>(row, cr, ci);
    }
}
// resident 128-thread blocks per SM by column count (register caps 128 / 168 / 255): the widest lanes that compile without
// spills under each cap (C = 6: 124 registers, C = 10: 162)
#ifndef K3_MINB4_MAX_C
#define K3_MINB4_MAX_C 6
#endif
#ifndef K3_MINB3_MAX_C
#define K3_MINB3_MAX_C 10
#endif
template <int LPG, int C, int THREADS>
struct K3Cfg {
    static constexpr int MINB_128 = (C <= K3_MINB4_MAX_C) ? 4 : (C <= K3_MINB3_MAX_C) ? 3 : 2;
    static constexpr int MINB = (MINB_128 * GW_THREADS / THREADS) < 1 ? 1 : (MINB_128 * GW_THREADS / THREADS);
    static constexpr int PMAX = (THREADS >= GW_THREADS) ? K3_PMAX : K3_WARP_PMAX;
};

template <int LPG, int C, int THREADS>
__global__ void __launch_bounds__(THREADS, K3Cfg<LPG, C, THREADS>::MINB)
k3_minors_kernel(const double *__restrict__ U0, size_t u_stride, int m, const unsigned char *__restrict__ occ_s,
                  const unsigned char *__restrict__ occ_t, const int *__restrict__ steps_total, const int *__restrict__ order,
                  int step, double *__restrict__ partials, unsigned long long *__restrict__ terms_out, unsigned long long per_block) {
    constexpr int W = LPG * C;
    constexpr int GROUPS = THREADS / LPG;
    constexpr int PMAX = K3Cfg<LPG, C, THREADS>::PMAX;
    constexpr int ROWBYTES = W * (int)sizeof(double2);
    extern __shared__ __align__(16) unsigned char k3_smem[];
    __shared__ GuanItem item;
    __shared__ short col_mode[W];
    // step tables, indexed by the DESTINATION position p + 1 inside a period: binomial product of the table digits at p and
    // the byte offset of the signed row that leads there (fwd: from p - 1, bwd: from p + 1); entries 0 and P + 1 are sentinels
    __shared__ K3Step fwd[PMAX + 2], bwd[PMAX + 2];
    __shared__ int low_digits;                 // digits 0 .. low_digits are driven by the table
    __shared__ unsigned period;                // P = prod_{v <= low_digits} (lim_v + 1): terms per table period

    const int slot = blockIdx.y, chunk = blockIdx.x, chunks = gridDim.x;
    const int sample = order ? order[slot] : slot;
    const double *U = U0 + (size_t)sample * u_stride;   // u_stride = 0: one interferometer for all samples
    double *my_part = partials + ((size_t)slot * chunks + chunk) * (size_t)(W * 4);
    if (steps_total && step >= steps_total[sample]) return;   // uniform-loss variant: this sample is complete

    const unsigned char *s = occ_s + (size_t)sample * m, *t = occ_t + (size_t)sample * m;
    // block setup by whole warps (ballots and scans over 32 modes at a time instead of one thread looping over m modes: the
    // serial version was half of the instructions of a small step, profiles/r02_k3_small_step_ncu.txt)
    if (threadIdx.x < 32) {
        guan_item_build_warp(item, t, m, /*inner_first=*/true);
        if (THREADS == 32) guan_expand_columns_warp(col_mode, s, m, W);
    } else if (threadIdx.x < 64) {                          // a second warp expands the input columns meanwhile
        guan_expand_columns_warp(col_mode, s, m, W);
    }
    __syncthreads();
    const int D = item.D;
    const int active = k3_active_chunks(item.terms, chunks, per_block);
    if (chunk == 0 && threadIdx.x == 0) terms_out[sample] = item.terms;
    if (chunk >= active) return;
    // shared-memory image: rows [0, D) = +2 U[mode_v][cols], rows [D, 2D) = -2 U[mode_v][cols], row 2D = 0
    double2 *X2 = reinterpret_cast<double2 *>(k3_smem);
    size_t x2_bytes = (size_t)(2 * D + 1) * W * sizeof(double2), red_bytes = (size_t)THREADS * C * sizeof(double2);
    unsigned char *rdig = k3_smem + ((x2_bytes > red_bytes ? x2_bytes : red_bytes) + 15) / 16 * 16;
    const double2 *U2 = reinterpret_cast<const double2 *>(U);
    for (int e = threadIdx.x; e < D * W; e += THREADS) {
        const int v = e / W, c = e - v * W;
        const int cm = col_mode[c];
        double2 x = make_double2(0.0, 0.0);
        if (cm >= 0) { const double2 u = U2[(int)item.mode[v] * m + cm]; x = make_double2(2.0 * u.x, 2.0 * u.y); }
        X2[e] = x;
        X2[D * W + e] = make_double2(-x.x, -x.y);
    }
    for (int c = threadIdx.x; c < W; c += THREADS) X2[2 * D * W + c] = make_double2(0.0, 0.0);
    const unsigned long long terms = item.terms;
    const unsigned long long ngroups = (unsigned long long)active * GROUPS;
    if (threadIdx.x == 0) {
        // table digits: digit 0 always (work is dealt out in whole sweeps of it), further digits while a lane group still
        // gets K3_PERIODS_PER_GROUP periods (balance: group ranges differ by at most one period) and the table fits
        unsigned long long raw = (terms + ngroups - 1) / ngroups, P = 1;
        int a = -1;
        for (int v = 0; v < D; ++v) {
            const unsigned long long nxt = P * (unsigned long long)(item.lim[v] + 1);
            if (v > 0 && (nxt * K3_PERIODS_PER_GROUP > raw || nxt > (unsigned long long)PMAX)) break;
            P = nxt; a = v;
        }
        // ... and beyond that while the periods still divide EVENLY among the lane groups (every group the same whole number of
        // periods: no imbalance to pay for, fewer period boundaries -- a walk of 2^10 terms over 16 groups becomes one 64-term period each)
        for (int v = a + 1; v < D; ++v) {
            const unsigned long long nxt = P * (unsigned long long)(item.lim[v] + 1);
            if (nxt > (unsigned long long)PMAX) break;
            const unsigned long long np = terms / nxt;
            if (np < ngroups || np % ngroups) break;
            P = nxt; a = v;
        }
        period = (unsigned)P;
        low_digits = a;
    }
    __syncthreads();
    const unsigned P = period;
    const int a_low = low_digits;
    for (unsigned p = threadIdx.x; p < P; p += THREADS) {
        // digits of position p and p - 1 of the reflected code over digits 0 .. a_low
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
        // a digit that goes UP lowers its coefficient t - 2 rho by 2: the negated row
        const int zero_row = 2 * D * ROWBYTES;
        fwd[p + 1].blow = bprod; fwd[p + 1].off = p ? (chg + (up ? D : 0)) * ROWBYTES : zero_row; fwd[p + 1].pad = 0;
        bwd[p + 1].blow = bprod;
        if (p == P - 1) { bwd[p + 1].off = zero_row; bwd[p + 1].pad = 0; }
        if (p) { bwd[p].off = (chg + (up ? 0 : D)) * ROWBYTES; bwd[p].pad = 0; }
        if (p == 0) {   // sentinels: the step "beyond the end" of a period adds the zero row and carries weight 0
            fwd[P + 1].blow = 0.0; fwd[P + 1].off = zero_row; fwd[P + 1].pad = 0;
            bwd[0].blow = 0.0; bwd[0].off = zero_row; bwd[0].pad = 0;
        }
    }
    __syncthreads();

    const int lane_in_group = threadIdx.x % LPG, group = threadIdx.x / LPG;
    // whole periods are dealt out to the lane groups; counts differ by at most one period
    const unsigned long long nper = terms / P;                    // terms is a multiple of P
    const unsigned long long gidx = (unsigned long long)chunk * GROUPS + group;
    const unsigned long long pbase = nper / ngroups, prem = nper % ngroups;
    const unsigned long long my_periods64 = pbase + (gidx < prem ? 1ull : 0ull);
    const unsigned long long hi0 = gidx * pbase + (gidx < prem ? gidx : prem);   // first period of this group
    // Trip counts are WARP-uniform (full-mask shuffles in the term loop): the first group of a warp has the most periods
    // (counts never grow with gidx); a group that owns one period less runs it with weight 0.
    const unsigned long long gidx_w = (unsigned long long)chunk * GROUPS + (threadIdx.x & ~31u) / LPG;
    const unsigned warp_periods = (unsigned)(pbase + (gidx_w < prem ? 1ull : 0ull));
    const unsigned my_periods = (unsigned)my_periods64;
    const int col0 = lane_in_group * C;

    double ar[C], ai[C];
#pragma unroll
    for (int j = 0; j < C; ++j) { ar[j] = 0.0; ai[j] = 0.0; }

    if (warp_periods > 0) {
        unsigned char *r = rdig + threadIdx.x;
        GuanState st;
        st.dirmask = 0ull; st.binom = 0.0;
        const bool mine = my_periods > 0;
        if (mine) guan_seek<THREADS>(item, hi0, r, st, /*v0=*/a_low + 1);      // digits above the table
        int pos = (hi0 & 1ull) ? (int)P - 1 : 0;            // position inside the period (odd periods run backwards: reflected code)
        int pdir = (hi0 & 1ull) ? -1 : 1;
        double cr[C], ci[C];
#pragma unroll
        for (int j = 0; j < C; ++j) { cr[j] = 0.0; ci[j] = 0.0; }
        int par = 0;                                         // parity of sum(rho) -> sign of the term
        if (mine) {
            unsigned q = (unsigned)pos;
#pragma unroll 1
            for (int v = 0; v < D; ++v) {
                int rv;
                if (v <= a_low) {
                    const unsigned R = (unsigned)item.lim[v] + 1u;
                    const unsigned d = gw_divmod(q, R);
                    rv = (q & 1u) ? (int)item.lim[v] - (int)d : (int)d;
                } else rv = (int)r[v * THREADS];
                par += rv;
                const double coef = 0.5 * (double)((int)item.mult[v] - 2 * rv);
                const double2 *row = X2 + v * W + col0;
#pragma unroll
                for (int j = 0; j < C; ++j) {
                    const double2 a = row[j];
                    cr[j] = fma(coef, a.x, cr[j]);
                    ci[j] = fma(coef, a.y, ci[j]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < C; ++j)
            if (col_mode[col0 + j] < 0) { cr[j] = 1.0; ci[j] = 0.0; }   // padding column: constant 1 (its rows are zero)
        // sign of the term x binomial product of the digits above the table; the sign flips with every term
        double bsgn = mine ? ((par & 1) ? -st.binom : st.binom) : 0.0;
        // shared-window addresses: this lane's first column of the image, and the two tables
        unsigned x2c0 = (unsigned)__cvta_generic_to_shared(X2 + col0);
        asm volatile("" : "+r"(x2c0));
        const K3Step *tab = (pdir > 0) ? fwd : bwd;
        double blow = fwd[pos + 1].blow;

#pragma unroll 1
        for (unsigned per = 0;;) {
#pragma unroll 1
            for (unsigned i = 0; i < P; ++i) {
                // step entry of the NEXT position, fetched before the term so that its latency hides behind the products
                pos += pdir;
                const K3Step e = tab[pos + 1];
                const double w = bsgn * blow;
                cplx node[C];
                if constexpr (LPG == 1) {
                    k3_tree_up<C, 0, C, false>(cr, ci, node);
                    k3_tree_down_real<C>(w, cr, ci, node, ar, ai);
                } else {
                    k3_tree_up<C, 0, C, true>(cr, ci, node);
                    cplx all = k3_tree_val<C, 0, C>(cr, ci, node), oth = {1.0, 0.0};
#pragma unroll
                    for (int mask = 1; mask < LPG; mask <<= 1) {
                        const cplx x = cshfl_xor(0xffffffffu, all, mask);
                        oth = (mask == 1) ? x : cmul(oth, x);
                        if ((mask << 1) < LPG) all = cmul(all, x);
                    }
                    cplx seed = {w * oth.re, w * oth.im};
                    k3_tree_down<C, 0, C>(seed, cr, ci, node, ar, ai);
                }
                bsgn = -bsgn;
                blow = e.blow;
                k3_row_add<C>(x2c0 + (unsigned)e.off, cr, ci);
            }
            // ---- period boundary: one Guan step of the digits above the table; the table digits stay and reverse
            if (++per >= warp_periods) break;
            pos -= pdir;                                     // the sentinel step overshot by one
            pdir = -pdir;
            tab = (pdir > 0) ? fwd : bwd;
            blow = fwd[pos + 1].blow;
            if (per < my_periods) {
                int delta;
                const int v = guan_step<THREADS>(item, r, st, delta, /*v0=*/a_low + 1);
                k3_row_add<C>(x2c0 + (unsigned)((v + (delta > 0 ? D : 0)) * ROWBYTES), cr, ci);
                bsgn = (bsgn < 0.0) ? -st.binom : st.binom;
            } else {
                bsgn = 0.0;                                  // dummy period of a group that owns one period less
            }
        }
    }

    // ---- block reduction: column (lane_in_group, j) over the GROUPS groups, double-double
    __syncthreads();   // the image is dead from here on; its storage is reused
    double2 *red = reinterpret_cast<double2 *>(k3_smem);   // [W][GROUPS]
#pragma unroll
    for (int j = 0; j < C; ++j) red[(col0 + j) * GROUPS + group] = make_double2(ar[j], ai[j]);
    __syncthreads();
    static_assert(THREADS >= W, "one reduction thread per column");
    constexpr int PARTS = (THREADS / W) > 8 ? 8 : (THREADS / W);
    constexpr int SLICE = (GROUPS + PARTS - 1) / PARTS;
    __shared__ double part_sum[THREADS * 4];
    {
        const int c = threadIdx.x / PARTS, part = threadIdx.x % PARTS;
        if (c < W) {
            dd re = {0.0, 0.0}, im = {0.0, 0.0};
            const int g1 = (part + 1) * SLICE < GROUPS ? (part + 1) * SLICE : GROUPS;
            for (int g = part * SLICE; g < g1; ++g) {
                const double2 x = red[c * GROUPS + g];
                re = dd_add_d(re, x.x);
                im = dd_add_d(im, x.y);
            }
            double *ps = part_sum + 4 * threadIdx.x;
            ps[0] = re.hi; ps[1] = re.lo; ps[2] = im.hi; ps[3] = im.lo;
        }
    }
    __syncthreads();
    if (threadIdx.x < W) {
        const int c = threadIdx.x;
        const double *ps = part_sum + 4 * (c * PARTS);
        dd re = {ps[0], ps[1]}, im = {ps[2], ps[3]};
#pragma unroll
        for (int q = 1; q < PARTS; ++q) {
            dd a = {ps[4 * q + 0], ps[4 * q + 1]}, b = {ps[4 * q + 2], ps[4 * q + 3]};
            re = dd_add(re, a);
            im = dd_add(im, b);
        }
        my_part[4 * c + 0] = re.hi; my_part[4 * c + 1] = re.lo;
        my_part[4 * c + 2] = im.hi; my_part[4 * c + 3] = im.lo;
    }
}

// ---------------------------------------------------------------------------------------------
// Finish kernel: one block per sample.
//   1. minors[mode] = 2^-(k-1) * sum over chunks of the partials of the first column of that mode
//      (k == 1: the occupations themselves, bs_submatrices_permanent_calculator_base.py:157-158)
//   2. pmf[j] = |sum_i s_i P_i U[j][i]|^2, then / total  (generalized_cliffords_b_simulation_strategy.py:82-92)
//   3. optional sampling update: numpy.random.choice semantics (cumsum, /= last, searchsorted right)
//      with the tape's u_choice (:107-110), r_sample[j] += 1, then the NEXT step's input particle:
//      pop index floor(u_pick * #remaining) of the remaining input particles (:102-105).
// ---------------------------------------------------------------------------------------------
#define K3F_THREADS 64
#define K3F_PARTS 16
// dynamic shared memory of the finish kernel: psum[(BP_MAX_N + 2) * parts * 4] doubles, P[m] double2, wgt[m] doubles, s_sh[m] bytes --
// sized by m (not BP_MAX_MODES) and 32 registers per thread, so that the 64-thread blocks of a 4096-sample batch (27.7 per SM) are
// all resident at once: one wave instead of two (24 blocks per SM before; profiles/r02_finish_ncu_summary.txt)
static inline size_t k3_finish_smem(int m, int parts) {
    return sizeof(double) * 4 * (BP_MAX_N + 2) * (size_t)parts + (sizeof(double2) + sizeof(double)) * (size_t)m + ((size_t)m + 15) / 16 * 16;
}
__global__ void __launch_bounds__(512, 4) k3_finish_kernel(K3Finish a, int parts_alloc) {
    extern __shared__ __align__(16) double psum[];   // [(BP_MAX_N + 2) * parts_alloc * 4]: parts_alloc = K3F_PARTS when chunks > 32, else 1
    double2 *P = reinterpret_cast<double2 *>(psum + 4 * (BP_MAX_N + 2) * parts_alloc);        // [m]
    double *wgt = reinterpret_cast<double *>(P + a.m);                                       // [m]
    unsigned char *s_sh = reinterpret_cast<unsigned char *>(wgt + a.m);                      // [m] input occupation of this sample
    __shared__ double total_sh;
    __shared__ int idx_sh;
    // chunk reduction of the few samples that had the whole GPU (up to 296 chunk blocks each): K3F_PARTS interleaved slices
    // per column summed by different threads, then added in slice order (fixed order); batches (<= 32 chunks) use one slice
    __shared__ short occ_mode[BP_MAX_N + 2], occ_col[BP_MAX_N + 2];
    __shared__ int nocc_sh;
    const int slot = blockIdx.x, m = a.m, k = a.step + 1;
    const int sample = a.order ? a.order[slot] : slot;
    if (a.steps_total && a.step >= a.steps_total[sample]) return;
    unsigned char *s = a.occ_s + (size_t)sample * m, *t = a.occ_t + (size_t)sample * m;
    __shared__ double sp_re[BP_MAX_N + 2], sp_im[BP_MAX_N + 2];  // s_i * P_i of the occupied input modes, mode order
    for (int v = threadIdx.x; v < m; v += blockDim.x) {
        const unsigned char sv = s[v];
        s_sh[v] = sv;
        P[v] = make_double2(k == 1 ? (double)sv : 0.0, 0.0);
    }
    if (threadIdx.x == 0) idx_sh = 0;
    __syncthreads();
    if (threadIdx.x < 32) {
        // the list of occupied input modes with the first column of each: warp scan over 32 modes at a time
        // (low half: particles = columns before the mode, high half: occupied modes before it)
        const int lane = threadIdx.x;
        int carry_c = 0, carry_o = 0;
        for (int base = 0; base < m; base += 32) {
            const int v = base + lane;
            const int sv = v < m ? (int)s_sh[v] : 0, occ = sv > 0 ? 1 : 0;
            int x = sv | (occ << 16);
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, x, d);
                if (lane >= d) x += y;
            }
            const int c = carry_c + (x & 0xffff) - sv, o = carry_o + (x >> 16) - occ;
            if (occ && o < BP_MAX_N + 2) { occ_mode[o] = (short)v; occ_col[o] = (short)c; }
            const int tot = __shfl_sync(0xffffffffu, x, 31);
            carry_c += tot & 0xffff;
            carry_o += tot >> 16;
        }
        if (lane == 0) nocc_sh = carry_o < BP_MAX_N + 2 ? carry_o : BP_MAX_N + 2;
    }
    __syncthreads();
    if (k > 1) {
        const double scale = ldexp(1.0, -(k - 1));
        const int nocc = nocc_sh;
        const int active = k3_active_chunks(a.terms[sample], a.chunks, a.per_block);
        const int parts = (a.chunks > 32 && active > 32) ? K3F_PARTS : 1;   // host: 512 threads and K3F_PARTS slices of psum when chunks > 32
        const double *base = a.partials + ((size_t)slot * a.chunks) * (size_t)(a.W * 4);
        for (int it = threadIdx.x; it < nocc * parts; it += blockDim.x) {
            const int i = it / parts, p = it - i * parts;
            const double *col = base + 4 * (int)occ_col[i];
            dd re = {0.0, 0.0}, im = {0.0, 0.0};
            for (int ch = p; ch < active; ch += parts) {   // fixed chunk order
                const double *q = col + (size_t)ch * (a.W * 4);
                dd x = {q[0], q[1]}, y = {q[2], q[3]};
                re = dd_add(re, x);
                im = dd_add(im, y);
            }
            double *ps = psum + 4 * it;
            ps[0] = re.hi; ps[1] = re.lo; ps[2] = im.hi; ps[3] = im.lo;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < nocc; i += blockDim.x) {
            const double *ps = psum + 4 * (i * parts);
            dd re = {ps[0], ps[1]}, im = {ps[2], ps[3]};
            for (int p = 1; p < parts; ++p) {
                dd x = {ps[4 * p + 0], ps[4 * p + 1]}, y = {ps[4 * p + 2], ps[4 * p + 3]};
                re = dd_add(re, x);
                im = dd_add(im, y);
            }
            const int mode = occ_mode[i];
            const double2 val = make_double2((re.hi + re.lo) * scale, (im.hi + im.lo) * scale);
            P[mode] = val;
            const double cnt = (double)s_sh[mode];               // permanent_added = s_i * P_i
            sp_re[i] = cnt * val.x; sp_im[i] = cnt * val.y;
        }
    } else {
        for (int i = threadIdx.x; i < nocc_sh; i += blockDim.x) {   // k = 1: P_i = s_i
            const double cnt = (double)s_sh[occ_mode[i]];
            sp_re[i] = cnt * cnt; sp_im[i] = cnt * 0.0;
        }
    }
    __syncthreads();
    if (a.minors_out)
        for (int v = threadIdx.x; v < m; v += blockDim.x) {
            a.minors_out[2 * ((size_t)sample * m + v)] = P[v].x; a.minors_out[2 * ((size_t)sample * m + v) + 1] = P[v].y;
        }
    if (!a.pmf_out && !a.tape) return;
    const double2 *U2 = reinterpret_cast<const double2 *>(a.U + (size_t)sample * a.u_stride);
    const int nocc_all = nocc_sh;
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        double re = 0.0, im = 0.0;
        const double2 *Uj = U2 + (size_t)j * m;
        // permanent_added = s_i * P_i; permanent_added *= U[j][i]; permanent += permanent_added -- occupied modes i in
        // ascending order (the reference skips nothing but adds exact zeros for the others); four loads in flight
        int i = 0;
        for (; i + 4 <= nocc_all; i += 4) {
            double2 u[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) u[q] = Uj[occ_mode[i + q]];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double sr = sp_re[i + q], si = sp_im[i + q];
                re += __dsub_rn(__dmul_rn(sr, u[q].x), __dmul_rn(si, u[q].y));
                im += __dadd_rn(__dmul_rn(sr, u[q].y), __dmul_rn(si, u[q].x));
            }
        }
        for (; i < nocc_all; ++i) {
            const double sr = sp_re[i], si = sp_im[i];
            const double2 u = Uj[occ_mode[i]];
            re += __dsub_rn(__dmul_rn(sr, u.x), __dmul_rn(si, u.y));
            im += __dadd_rn(__dmul_rn(sr, u.y), __dmul_rn(si, u.x));
        }
        const double ab = hypot(re, im);   // abs(permanent) ** 2
        wgt[j] = ab * ab;
    }
    __syncthreads();
    // the sums stay sequential (python sum() / numpy.cumsum order); the element-wise divisions run in parallel
    if (threadIdx.x == 0) {
        double total = 0.0;
        for (int j = 0; j < m; ++j) total += wgt[j];
        total_sh = total;
        if (a.err_flag && !(total > 0.0 && total < 1.0 / 0.0)) atomicOr(a.err_flag, 1);   // zero, NaN or infinite: no distribution
    }
    __syncthreads();
    const double total = total_sh;
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        const double w = wgt[j] / total;
        wgt[j] = w;
        if (a.pmf_out) a.pmf_out[(size_t)sample * m + j] = w;
    }
    if (!a.tape) return;
    __syncthreads();
    const double *tp = a.tape + (size_t)sample * a.tape_stride;
    // numpy.random.choice: cdf = cumsum(p); cdf /= cdf[-1]; searchsorted(cdf, u, side='right')
    if (threadIdx.x == 0) {
        double run = 0.0;
        for (int j = 0; j < m; ++j) { run += wgt[j]; wgt[j] = run; }
    }
    __syncthreads();
    {
        const double last = wgt[m - 1], u = tp[2 + 2 * a.step];
        int mine = 0;
        for (int j = threadIdx.x; j < m; j += blockDim.x) if (wgt[j] / last <= u) mine = j + 1;   // j ascending: last hit wins
        if (mine) atomicMax(&idx_sh, mine);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int idx = idx_sh;
        if (idx >= m) idx = m - 1;
        t[idx] += 1;
        // next step's input particle
        const int nsteps = a.steps_total ? a.steps_total[sample] : a.n;
        if (a.step + 1 < nsteps) {
            int nr = a.n_remaining[sample];
            unsigned char *rem = a.remaining + (size_t)sample * a.n;
            int pick = (int)(tp[1 + 2 * (a.step + 1)] * (double)nr);
            if (pick >= nr) pick = nr - 1;
            const int mode = rem[pick];
            for (int q = pick; q + 1 < nr; ++q) rem[q] = rem[q + 1];   // list.pop(pick)
            a.n_remaining[sample] = nr - 1;
            s[mode] += 1;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host-side dispatch
// ---------------------------------------------------------------------------------------------
typedef void (*k3_fn)(const double *, size_t, int, const unsigned char *, const unsigned char *, const int *, const int *, int, double *, unsigned long long *, unsigned long long);

struct K3Variant { k3_fn fn; int lpg, c, threads; };
#define K3_WARP_DEFAULT_MAX_K 16
// one lane owns all columns up to k = K3_MAX_C1, two (four) lanes share them beyond: the widest lanes that compile without spills
#define K3_MAX_C1 17
#define K3_MAX_C2 15
#define K3_MAX_C4 12
#define K3_WARP_MAX_C 16
static K3Variant g_k3[3][K3_MAX_C1 + 1];     // [log2 LPG][C], 128-thread blocks
static K3Variant g_k3w[K3_WARP_MAX_C + 1];   // LPG = 1, one-warp blocks
static K3Variant g_k3w2[K3_WARP_MAX_C / 2 + 1];   // LPG = 2, one-warp blocks (k <= 16): [C]

template <int C>
static void k3_reg() {
    g_k3[0][C] = K3Variant{k3_minors_kernel<1, C, GW_THREADS>, 1, C, GW_THREADS};
    if constexpr (C <= K3_WARP_MAX_C) g_k3w[C] = K3Variant{k3_minors_kernel<1, C, K3_WARP_THREADS>, 1, C, K3_WARP_THREADS};
    if constexpr (C <= K3_WARP_MAX_C / 2 && C >= 4) g_k3w2[C] = K3Variant{k3_minors_kernel<2, C, K3_WARP_THREADS>, 2, C, K3_WARP_THREADS};
    if constexpr (C <= K3_MAX_C2 && C >= 4) g_k3[1][C] = K3Variant{k3_minors_kernel<2, C, GW_THREADS>, 2, C, GW_THREADS};
    if constexpr (C <= K3_MAX_C4 && C >= 4) g_k3[2][C] = K3Variant{k3_minors_kernel<4, C, GW_THREADS>, 4, C, GW_THREADS};
    if constexpr (C > 1) k3_reg<C - 1>();
}

// Variant for k input columns: one lane per term stream with all k <= 17 columns, two lanes with ceil(k / 2) columns each for
// k = 18 .. 30, four lanes for k = 31 .. 48; one-warp blocks for the steps k <= K3_WARP_DEFAULT_MAX_K whose walk fits one block --
// with two lanes per term stream from k = 8 (half the registers per lane, twice the resident blocks: these steps are setup- and
// latency-bound; n = 16 run -5 %, n = 12 run -8 %, profiles/r02_k3_history.txt).
// Tuning knobs, read once: BP_K3_WARP_MAX_K (0 = never one-warp blocks), BP_K3_WARP2_MIN_K (first k of the two-lane one-warp blocks),
// BP_K3_TREE_MAX_C (column limit per lane, >= 6).
static K3Variant k3_pick(int k) {
    static int warp_max_k = K3_WARP_DEFAULT_MAX_K, max_c1 = K3_MAX_C1, warp2_min_k = 8;
    static const bool ready = [] {   // thread-safe one-time registration (C++11 static initialisation)
#ifdef K3_DEV_C   // development builds: a single instantiation (seconds to compile; for SASS inspection only)
        g_k3[0][K3_DEV_C] = K3Variant{k3_minors_kernel<K3_DEV_LPG, K3_DEV_C, K3_DEV_THREADS>, K3_DEV_LPG, K3_DEV_C, K3_DEV_THREADS};
#else
        k3_reg<K3_MAX_C1>();
#endif
        const char *e;
        if ((e = getenv("BP_K3_WARP_MAX_K"))) warp_max_k = atoi(e);
        if ((e = getenv("BP_K3_TREE_MAX_C"))) max_c1 = atoi(e);
        if ((e = getenv("BP_K3_WARP2_MIN_K"))) warp2_min_k = atoi(e);   // one-warp blocks with two lanes per term stream from this k
        if (max_c1 > K3_MAX_C1) max_c1 = K3_MAX_C1;
        if (max_c1 < 6) max_c1 = 6;
        return true;
    }();
    (void)ready;
    K3Variant none = {nullptr, 0, 0, 0};
    if (k < 1) return none;
    if (k >= warp2_min_k && k <= warp_max_k && k <= K3_WARP_MAX_C && (k + 1) / 2 >= 4 && g_k3w2[(k + 1) / 2].fn) return g_k3w2[(k + 1) / 2];
    // fewest lanes per group whose column count fits the lane limit
    for (int lg = 0; lg < 3; ++lg) {
        const int lpg = 1 << lg, c = (k + lpg - 1) / lpg;
        const int lim = lg == 0 ? max_c1 : lg == 1 ? (max_c1 < K3_MAX_C2 ? max_c1 : K3_MAX_C2) : K3_MAX_C4;
        if (c > lim) continue;
        if (lg == 0 && k <= warp_max_k && k <= K3_WARP_MAX_C) return g_k3w[k];
        if (g_k3[lg][c].fn) return g_k3[lg][c];
    }
    return none;
}

int bp_k3_width(int k) { K3Variant v = k3_pick(k); return v.fn ? v.lpg * v.c : 0; }
// Work sizing of step k over `samples` samples.  A sample whose walk has T terms is served by
// ceil(T / per_block) chunk blocks (at most `chunks`, the launched grid width); the other blocks exit at once.
//   * per_block = (terms per lane group) x (groups per block).  With few samples the target is 192 terms per
//     group (parallelism first); with enough samples to fill the GPU anyway it grows to 2048, because every
//     block pays ~18 us of setup and reduction (measured: +62 ns of kernel time per extra block at 4096 samples);
//   * chunks is sized for the collision-free worst case 2^(k-2), capped so that the heaviest samples still split
//     into blocks of bounded duration (tail of the launch) without flooding the grid with empty blocks.
// BP_K3_TPG / BP_K3_CAP override the two knobs (tuning).
static void k3_plan(bp_context *h, int k, long long samples, int *chunks_out, unsigned long long *per_block_out) {
    static int env_tpg = -1, env_cap = -1;
    if (env_tpg < 0) { const char *e = getenv("BP_K3_TPG"); env_tpg = e ? atoi(e) : 0; }
    if (env_cap < 0) { const char *e = getenv("BP_K3_CAP"); env_cap = e ? atoi(e) : 0; }
    K3Variant v = k3_pick(k);
    const int threads = v.threads ? v.threads : GW_THREADS;
    const int groups = threads / (v.lpg ? v.lpg : 1);
    const long long fill_blocks = (long long)h->sm_count * 8 * (GW_THREADS / threads);   // ~32 warps per SM
    long long tpg = (long long)K3_TERMS_PER_GROUP;
    if (samples >= fill_blocks) tpg = 2048;
    else if (2048 * samples / fill_blocks > tpg) tpg = 2048 * samples / fill_blocks;
    if (env_tpg > 0) tpg = env_tpg;
    const unsigned long long per_block = (unsigned long long)tpg * (unsigned long long)groups;
    int chunks = 1;
    if (k > 1) {
        const double max_terms = ldexp(1.0, k - 2);
        long long by_work = (long long)ceil(max_terms / (double)per_block);
        if (by_work < 1) by_work = 1;
        long long by_fill = ((long long)h->sm_count * 64 + samples - 1) / samples;
        const long long cap = env_cap > 0 ? env_cap : 16;   // measured optimum 8 .. 32 (profiles/r01_k3_sizing.txt)
        if (by_fill < cap) by_fill = cap;
        long long ch = by_work < by_fill ? by_work : by_fill;
        // one full wave of two resident blocks per SM when a single sample has the GPU to itself
        const long long wave = 2ll * h->sm_count;
        if (ch > wave) ch = wave;
        chunks = (int)ch;
    }
    *chunks_out = chunks;
    *per_block_out = per_block;
}

int bp_k3_chunks(bp_context *h, int k, long long samples) {
    int ch; unsigned long long pb;
    k3_plan(h, k, samples, &ch, &pb);
    return ch;
}
unsigned long long bp_k3_per_block(bp_context *h, int k, long long samples) {
    int ch; unsigned long long pb;
    k3_plan(h, k, samples, &ch, &pb);
    return pb;
}

// Enqueue the minors main kernel for step k (= particles in occ_s) over `samples` launch slots (d_order: slot -> sample,
// NULL = identity).
int bp_k3_launch(bp_context *h, const double *dU, size_t u_stride, int m, const unsigned char *d_s, const unsigned char *d_t,
                 const int *d_steps_total, const int *d_order, int k, long long samples, int chunks, double *d_partials,
                 unsigned long long *d_terms) {
    if (k <= 1) return BP_OK;   // handled by the finish kernel
    K3Variant v = k3_pick(k);
    if (!v.fn) return bp_fail(h, BP_ERR_UNSUPPORTED, "minors kernel supports k <= %d, got %d", 4 * K3_MAX_C4, k);
    if (k - 1 > BP_MAX_N) return bp_fail(h, BP_ERR_UNSUPPORTED, "minors kernel supports k - 1 <= %d, got k = %d", BP_MAX_N, k);
    if (samples > 65535) return bp_fail(h, BP_ERR_INVALID, "bp_k3_launch: at most 65535 samples per launch");
    const size_t smem = k3_smem_bytes(k - 1, v.lpg * v.c, v.c, v.threads);
    if (smem > 24 * 1024) {   // static shared memory (step tables) takes ~17 KB of the 48 KB that need no opt-in
        cudaError_t e = cudaFuncSetAttribute((const void *)v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return bp_fail(h, BP_ERR_CUDA, "cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
    }
    dim3 grid((unsigned)chunks, (unsigned)samples);
    v.fn<<<grid, v.threads, smem, h->stream>>>(dU, u_stride, m, d_s, d_t, d_steps_total, d_order, k - 1, d_partials, d_terms,
                                                bp_k3_per_block(h, k, samples));
    BP_CHECK_LAUNCH(h);
    return BP_OK;
}

int bp_k3_finish_launch(bp_context *h, const K3Finish &a, long long samples) {
    // few samples (< 2 x SM count) may have up to 2 x SM-count chunk partials per column to add: give their blocks more threads
    const bool many_chunks = a.chunks > 32;
    const int threads = many_chunks ? 512 : K3F_THREADS;
    const int parts = many_chunks ? K3F_PARTS : 1;
    k3_finish_kernel<<<(unsigned)samples, threads, k3_finish_smem(a.m, parts), h->stream>>>(a, parts);
    BP_CHECK_LAUNCH(h);
    return BP_OK;
}
