/*
 * oracle/oracle_impl.h -- TEST INFRASTRUCTURE ONLY (see bossperm_oracle.c for the header note).
 *
 * Precision-generic body of the CPU restatement.  Included twice by bossperm_oracle.c with
 *   REAL = double        , SFX(x) = x##_d    (same arithmetic type as the NumPy reference)
 *   REAL = long double   , SFX(x) = x##_ld   (80-bit x87 ground truth for n > ~16)
 * All reference citations are relative to /root/reference/theboss/.
 */

typedef struct { REAL re, im; } SFX(cplx);

static inline SFX(cplx) SFX(cmul)(SFX(cplx) a, SFX(cplx) b) {
    SFX(cplx) r;
    r.re = a.re * b.re - a.im * b.im;
    r.im = a.re * b.im + a.im * b.re;
    return r;
}
static inline SFX(cplx) SFX(cpowi)(SFX(cplx) a, int e) {
    /* pow(complex, small non-negative int) by repeated multiplication
       (bs_permanent_calculator_base.py:204, python pow on complex128 with int exponent). */
    SFX(cplx) r = {1, 0};
    SFX(cplx) b = a;
    while (e > 0) {
        if (e & 1) r = SFX(cmul)(r, b);
        e >>= 1;
        if (e) b = SFX(cmul)(b, b);
    }
    return r;
}

/* ------------------------------------------------------------------------------------------
 * A.1  Gray-code Glynn on an explicit N x N matrix A (row-major, interleaved re/im doubles).
 * Follows boson_sampling_utilities/permanent_calculators/glynn_gray_permanent_calculator.py:
 *   init   :73-84   delta = ones, sums[j] = sum_i delta[i]*A[i][j], perm = prod_j sums[j]
 *   loop   :60-67   mult=-mult; delta[i]=-delta[i]; sums[j] += 2*delta[i]*A[i][j];
 *                   perm += mult*prod(sums)
 *   scale  :69      perm /= 2^(N-1)
 * The flip index sequence (guancodes, :57) is the binary-reflected Gray ruler ctz(step).
 * This routine evaluates the step range [lo, hi) of the 2^(N-1) terms (step 0 = all-ones delta),
 * un-normalised, so that a caller can split the term space; SFX(orc_glynn_gray) below runs the
 * whole range sequentially exactly like the reference.
 * ------------------------------------------------------------------------------------------ */
static void SFX(glynn_range)(const double *A, int N, uint64_t lo, uint64_t hi, REAL out[2]) {
    SFX(cplx) sums[ORC_MAX_N];
    int delta[ORC_MAX_N];
    uint64_t g = lo ^ (lo >> 1);
    for (int i = 0; i < N; ++i) delta[i] = ((i < N - 1) && ((g >> i) & 1)) ? -1 : 1;
    for (int j = 0; j < N; ++j) {
        SFX(cplx) s = {0, 0};
        for (int i = 0; i < N; ++i) {
            s.re += delta[i] * (REAL)A[2 * (i * N + j)];
            s.im += delta[i] * (REAL)A[2 * (i * N + j) + 1];
        }
        sums[j] = s;
    }
    int mult = (lo & 1) ? -1 : 1;
    REAL pr = 0, pi = 0;
    {
        SFX(cplx) p = {1, 0};
        for (int j = 0; j < N; ++j) p = SFX(cmul)(p, sums[j]);
        pr += mult * p.re;
        pi += mult * p.im;
    }
    for (uint64_t step = lo + 1; step < hi; ++step) {
        int i = __builtin_ctzll(step);
        mult = -mult;
        delta[i] = -delta[i];
        REAL two_d = 2 * delta[i];
        const double *row = A + 2 * (size_t)i * N;
        SFX(cplx) p = {1, 0};
        for (int j = 0; j < N; ++j) {
            sums[j].re += two_d * (REAL)row[2 * j];
            sums[j].im += two_d * (REAL)row[2 * j + 1];
            p = SFX(cmul)(p, sums[j]);
        }
        pr += mult * p.re;
        pi += mult * p.im;
    }
    out[0] = pr;
    out[1] = pi;
}

int SFX(orc_glynn_gray)(const double *A, int N, double out[2]) {
    if (N < 0 || N > ORC_MAX_N) return -1;
    if (N == 0) { out[0] = 1; out[1] = 0; return 0; } /* glynn_gray_permanent_calculator.py:52-53 */
    REAL acc[2];
    uint64_t T = (uint64_t)1 << (N - 1);
    SFX(glynn_range)(A, N, 0, T, acc);
    REAL scale = (REAL)T;
    out[0] = (double)(acc[0] / scale);
    out[1] = (double)(acc[1] / scale);
    return 0;
}

/* Same quantity, term space split into `nchunks` contiguous Gray ranges evaluated by a pthread
 * pool (the reference's own parallel pattern is process pools over independent work, e.g.
 * simulation_strategies/nonuniform_losses_approximation_strategy.py:254-257); partials are
 * combined in chunk order so the result does not depend on the thread count. */
typedef struct { const double *A; int N; int nchunks; uint64_t T; REAL *part; } SFX(par_ctx);
static void SFX(par_body)(int c, void *vctx) {
    SFX(par_ctx) *x = (SFX(par_ctx) *)vctx;
    uint64_t q = x->T / x->nchunks, rem = x->T % x->nchunks;
    uint64_t lo = q * c + (rem < (uint64_t)c ? rem : (uint64_t)c);
    uint64_t len = q + ((uint64_t)c < rem ? 1 : 0);
    SFX(glynn_range)(x->A, x->N, lo, lo + len, x->part + 2 * c);
}
int SFX(orc_glynn_gray_par)(const double *A, int N, int nchunks, int nthreads, double out[2]) {
    if (N < 0 || N > ORC_MAX_N) return -1;
    if (N == 0) { out[0] = 1; out[1] = 0; return 0; }
    uint64_t T = (uint64_t)1 << (N - 1);
    if (nchunks < 1) nchunks = 1;
    if ((uint64_t)nchunks > T) nchunks = (int)T;
    REAL *part = (REAL *)malloc(sizeof(REAL) * 2 * (size_t)nchunks);
    if (!part) return -2;
    SFX(par_ctx) ctx = {A, N, nchunks, T, part};
    orc_parallel_for(nchunks, SFX(par_body), &ctx, nthreads);
    REAL sr = 0, si = 0;
    for (int c = 0; c < nchunks; ++c) { sr += part[2 * c]; si += part[2 * c + 1]; }
    free(part);
    out[0] = (double)(sr / (REAL)T);
    out[1] = (double)(si / (REAL)T);
    return 0;
}

/* Un-normalised partial over Gray steps [lo, hi) -- used to check bp_glynn_range shards. */
int SFX(orc_glynn_gray_range)(const double *A, int N, uint64_t lo, uint64_t hi, double out[2]) {
    if (N < 1 || N > ORC_MAX_N) return -1;
    REAL acc[2];
    SFX(glynn_range)(A, N, lo, hi, acc);
    out[0] = (double)acc[0];
    out[1] = (double)acc[1];
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * A.2  Guan-code driver state, shared by the single and the submatrices calculators.
 *   _update_guan_code          bs_permanent_calculator_base.py:123-149
 *                              (= bs_submatrices_permanent_calculator_base.py:107-133)
 *   _update_binomials_product  bs_permanent_calculator_base.py:151-164 (floating true division)
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int r[ORC_MAX_M], u[ORC_MAX_M], lim[ORC_MAX_M];
    int len, idx, last;
    REAL binom;
} SFX(guan);

static void SFX(guan_init)(SFX(guan) * g, const int *s, int m) {
    g->len = m;
    for (int i = 0; i < m; ++i) { g->r[i] = 0; g->u[i] = 1; g->lim[i] = s[i]; }
    g->idx = 0; g->last = 0; g->binom = 1;
}
/* returns 0 when the code is exhausted (index_to_update == len) */
static int SFX(guan_step)(SFX(guan) * g) {
    int i = 0;
    int k = g->r[0] + g->u[0];
    while (k > g->lim[i] || k < 0) {
        g->u[i] = -g->u[i];
        ++i;
        if (i == g->len) { g->idx = i; return 0; }
        k = g->r[i] + g->u[i];
    }
    g->last = g->r[i];
    g->r[i] = k;
    g->idx = i;
    return 1;
}
static void SFX(guan_binom)(SFX(guan) * g, const int *s) {
    int i = g->idx;
    if (g->r[i] > g->last)
        g->binom *= (REAL)(s[i] - g->last) / (REAL)g->r[i];
    else
        g->binom *= (REAL)g->last / (REAL)(s[i] - g->r[i]);
}

/* ------------------------------------------------------------------------------------------
 * A.2-A.4  Ryser / Chin-Huh single permanent with input/output multiplicities.
 *   compute_permanent   bs_permanent_calculator_base.py:166-198
 *   _update_permanent   bs_permanent_calculator_base.py:200-209
 *   Ryser   init ryser_permanent_calculator.py:45-53, sums update :55-64
 *   ChinHuh init chin_huh_permanent_calculator.py:38-48, sums update :50-59
 * U is m x m row-major interleaved; s, t have length m (callers pad shorter states with zeros,
 * which touches exactly the same matrix entries as the reference's shorter loops).
 * formula: 0 = Ryser, 1 = Chin-Huh.
 * ------------------------------------------------------------------------------------------ */
int SFX(orc_guan_permanent)(const double *U, int m, const int *s, const int *t, int formula,
                            double out[2]) {
    if (m < 1 || m > ORC_MAX_M) return -1;
    SFX(guan) g;
    SFX(guan_init)(&g, s, m);
    int cols[ORC_MAX_M], nc = 0, n = 0;
    for (int j = 0; j < m; ++j) { if (t[j] != 0) cols[nc++] = j; n += s[j]; }
    SFX(cplx) sums[ORC_MAX_M];
    REAL mult;
    if (formula == 0) {
        mult = (n & 1) ? -1 : 1;
        for (int c = 0; c < nc; ++c) { sums[c].re = 0; sums[c].im = 0; }
    } else {
        mult = 1;
        for (int i = 0; i < n; ++i) mult /= 2; /* 1 / pow(2, n), exact */
        for (int c = 0; c < nc; ++c) {
            int i = cols[c];
            SFX(cplx) a = {0, 0};
            for (int j = 0; j < m; ++j) {
                a.re += s[j] * (REAL)U[2 * (i * m + j)];
                a.im += s[j] * (REAL)U[2 * (i * m + j) + 1];
            }
            sums[c] = a;
        }
    }
    REAL pr = 0, pi = 0;
#define ORC_ADD_TERM()                                                              \
    do {                                                                            \
        SFX(cplx) p = {mult * g.binom, 0};                                          \
        for (int c = 0; c < nc; ++c) p = SFX(cmul)(p, SFX(cpowi)(sums[c], t[cols[c]])); \
        pr += p.re; pi += p.im;                                                     \
    } while (0)
    ORC_ADD_TERM(); /* initial term: ryser :53, chin-huh :48 */
    while (g.r[m - 1] <= g.lim[m - 1]) {
        if (!SFX(guan_step)(&g)) break;
        mult = -mult;
        int idx = g.idx;
        int dr = g.r[idx] - g.last;
        for (int c = 0; c < nc; ++c) {
            int j = cols[c];
            REAL ur = (REAL)U[2 * (j * m + idx)], ui = (REAL)U[2 * (j * m + idx) + 1];
            if (formula == 0) { sums[c].re += dr * ur; sums[c].im += dr * ui; }
            else              { sums[c].re -= 2 * dr * ur; sums[c].im -= 2 * dr * ui; }
        }
        SFX(guan_binom)(&g, s);
        ORC_ADD_TERM();
    }
#undef ORC_ADD_TERM
    out[0] = (double)pr;
    out[1] = (double)pi;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * A.5  Submatrices ("all one-input-particle-removed minors") calculators.
 *   compute_permanents  bs_submatrices_permanent_calculator_base.py:150-175 (k=1 case :157-158)
 *   init                :177-189  (mult0 = (-1)^sum(t))
 *   Ryser  variant  bs_cc_ryser_submatrices_permanent_calculator.py:73-119
 *   ChinHuh variant bs_cc_ch_submatrices_permanent_calculator.py:51-105
 * out: m complex (2m doubles).  formula: 0 = Ryser, 1 = Chin-Huh.
 * ------------------------------------------------------------------------------------------ */
int SFX(orc_submatrices)(const double *U, int m, const int *s, const int *t, int formula,
                         double *out) {
    if (m < 1 || m > ORC_MAX_M) return -1;
    int k = 0, nt = 0;
    for (int i = 0; i < m; ++i) { k += s[i]; nt += t[i]; }
    if (k == 1) {
        for (int i = 0; i < m; ++i) { out[2 * i] = (double)s[i]; out[2 * i + 1] = 0; }
        return 0;
    }
    SFX(cplx) P[ORC_MAX_M];
    for (int i = 0; i < m; ++i) { P[i].re = 0; P[i].im = 0; }
    SFX(guan) g;
    SFX(guan_init)(&g, s, m);
    int cols[ORC_MAX_M], nc = 0;
    for (int j = 0; j < m; ++j) if (t[j] != 0) cols[nc++] = j;
    SFX(cplx) sums[ORC_MAX_M];
    REAL mult = (nt & 1) ? -1 : 1;
    if (formula == 0) {
        for (int c = 0; c < nc; ++c) { sums[c].re = 0; sums[c].im = 0; }
    } else {
        mult = 1;
        for (int i = 0; i < k - 1; ++i) mult /= 2; /* 1 / pow(2, k-1) */
        for (int c = 0; c < nc; ++c) {
            int i = cols[c];
            SFX(cplx) a = {0, 0};
            for (int j = 0; j < m; ++j) {
                a.re += s[j] * (REAL)U[2 * (i * m + j)];
                a.im += s[j] * (REAL)U[2 * (i * m + j) + 1];
            }
            sums[c] = a;
        }
    }
#define ORC_CH_UPDATE()                                                                      \
    do {                                                                                     \
        for (int i = 0; i < m; ++i) {                                                        \
            if (s[i] == 0 || s[i] == g.r[i]) continue;                                       \
            REAL ub = g.binom / ((REAL)s[i] / (REAL)(s[i] - g.r[i]));                        \
            SFX(cplx) p = {mult * ub, 0};                                                    \
            for (int c = 0; c < nc; ++c) {                                                   \
                int j = cols[c];                                                             \
                SFX(cplx) v = {sums[c].re - (REAL)U[2 * (j * m + i)],                        \
                               sums[c].im - (REAL)U[2 * (j * m + i) + 1]};                   \
                p = SFX(cmul)(p, SFX(cpowi)(v, t[j]));                                       \
            }                                                                                \
            P[i].re += p.re; P[i].im += p.im;                                                \
        }                                                                                    \
    } while (0)
    if (formula == 1) ORC_CH_UPDATE(); /* bs_cc_ch...:64 initial r = 0 update */
    while (g.r[m - 1] <= g.lim[m - 1]) {
        if (!SFX(guan_step)(&g)) break;
        mult = -mult;
        SFX(guan_binom)(&g, s);
        int idx = g.idx;
        int dr = g.r[idx] - g.last;
        if (formula == 0) {
            /* _update_sums bs_cc_ryser...:80-94 */
            SFX(cplx) prod = {1, 0};
            for (int c = 0; c < nc; ++c) {
                int j = cols[c];
                sums[c].re += dr * (REAL)U[2 * (j * m + idx)];
                sums[c].im += dr * (REAL)U[2 * (j * m + idx) + 1];
                prod = SFX(cmul)(prod, SFX(cpowi)(sums[c], t[j]));
            }
            prod.re *= mult; prod.im *= mult;
            /* _update_permanents bs_cc_ryser...:96-119 */
            for (int i = 0; i < m; ++i) {
                if (s[i] == 0 || g.r[i] == s[i]) continue;
                REAL ub = g.binom / ((REAL)s[i] / (REAL)(s[i] - g.r[i]));
                P[i].re += prod.re * ub;
                P[i].im += prod.im * ub;
            }
        } else {
            for (int c = 0; c < nc; ++c) {
                int j = cols[c];
                sums[c].re -= 2 * dr * (REAL)U[2 * (j * m + idx)];
                sums[c].im -= 2 * dr * (REAL)U[2 * (j * m + idx) + 1];
            }
            ORC_CH_UPDATE();
        }
    }
#undef ORC_CH_UPDATE
    for (int i = 0; i < m; ++i) { out[2 * i] = (double)P[i].re; out[2 * i + 1] = (double)P[i].im; }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * GCC-B step pmf: simulation_strategies/generalized_cliffords_b_simulation_strategy.py:69-92
 *   minors via the Ryser submatrices calculator (:73-77), then for every output mode j
 *   p_j = | sum_i s_i * P_i * U[j][i] |^2 (:82-89), normalised by the sequential sum (:91-92).
 * Also returns the un-normalised weights (before :91) in `raw` when non-NULL.
 * ------------------------------------------------------------------------------------------ */
int SFX(orc_gccb_pmf)(const double *U, int m, const int *s_cur, const int *r_sample, double *pmf,
                      double *raw) {
    double minors[2 * ORC_MAX_M];
    int rc = SFX(orc_submatrices)(U, m, s_cur, r_sample, 0, minors);
    if (rc) return rc;
    REAL w[ORC_MAX_M];
    REAL total = 0;
    for (int j = 0; j < m; ++j) {
        SFX(cplx) acc = {0, 0};
        for (int i = 0; i < m; ++i) {
            SFX(cplx) a = {s_cur[i] * (REAL)minors[2 * i], s_cur[i] * (REAL)minors[2 * i + 1]};
            SFX(cplx) b = {(REAL)U[2 * (j * m + i)], (REAL)U[2 * (j * m + i) + 1]};
            SFX(cplx) p = SFX(cmul)(a, b);
            acc.re += p.re; acc.im += p.im;
        }
        /* abs(z) ** 2 in the reference: hypot then square */
        REAL a = SFX(orc_hypot)(acc.re, acc.im);
        w[j] = a * a;
        total += w[j];
    }
    for (int j = 0; j < m; ++j) {
        if (raw) raw[j] = (double)w[j];
        pmf[j] = (double)(w[j] / total);
    }
    return 0;
}
