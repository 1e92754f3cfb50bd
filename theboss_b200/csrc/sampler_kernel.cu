// sampler_kernel.cu -- K4: device-resident GCC-B sampling loop around the minors kernel (K3), and
// the C ABI entry points bp_minors / bp_gccb_pmf / bp_gccb_simulate.
//
// Replaces GeneralizedCliffordsBSimulationStrategy.simulate / _fill_r_sample
// (reference: theboss/simulation_strategies/generalized_cliffords_b_simulation_strategy.py:41-67, :94-110)
// and GeneralizedCliffordsBUniformLossesSimulationStrategy.simulate
// (theboss/simulation_strategies/generalized_cliffords_b_uniform_losses_simulation_strategy.py:50-121).
//
// All samples of a batch advance in lock step: step k launches the minors kernel over
// (chunks x samples) blocks and one finish kernel that reduces the chunks, forms the pmf, draws the
// output mode and admits the next input particle.  No host synchronisation inside the loop; every
// random decision comes from the decision tape (include/bossperm.h).
#include <math.h>
#include <string.h>

#include <new>
#include <vector>

#include "bp_common.cuh"
#include "guan_walker.cuh"
#include "minors.cuh"

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based generator: uniform(seed, sample, slot) does not depend on how the
// samples are split over launches or GPUs.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox_round(unsigned &c0, unsigned &c1, unsigned &c2, unsigned &c3, unsigned k0, unsigned k1) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const unsigned n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}
__device__ inline double philox_uniform(unsigned long long seed, unsigned long long sample, unsigned slot) {
    unsigned c0 = (unsigned)sample, c1 = (unsigned)(sample >> 32), c2 = slot, c3 = 0x5bd1e995u;
    unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    const unsigned long long bits = (((unsigned long long)c0 << 32) | c1) >> 11;   // 53 bits
    return (double)bits * (1.0 / 9007199254740992.0);                               // [0, 1)
}

__global__ void k4_fill_tape_kernel(double *__restrict__ tape, long long samples, int stride, unsigned long long seed,
                                    long long first_sample) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= samples * stride) return;
    const long long sample = i / stride;
    const int slot = (int)(i - sample * stride);
    tape[i] = philox_uniform(seed, (unsigned long long)(first_sample + sample), (unsigned)slot);
}

// One thread per sample: particle number (uniform-loss inverse CDF,
// generalized_cliffords_b_uniform_losses_simulation_strategy.py:67-85), remaining-particle list
// (mode assignment, boson_sampling_utilities.py:61-78), empty states, and the first input particle.
__global__ void k4_init_kernel(const unsigned char *__restrict__ s0_base, size_t s0_stride, int m, int n, const double *__restrict__ loss_weights,
                               const double *__restrict__ tape, int tape_stride, long long samples,
                               unsigned char *__restrict__ occ_s, unsigned char *__restrict__ occ_t,
                               unsigned char *__restrict__ remaining, int *__restrict__ n_remaining,
                               int *__restrict__ steps_total, int *__restrict__ err_flag) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= samples) return;
    const double *tp = tape + i * tape_stride;
    const unsigned char *s0 = s0_base + (size_t)i * s0_stride;   // s0_stride = 0: one input state for all samples
    int steps = 0;
    for (int v = 0; v < m; ++v) steps += s0[v];                   // particles of THIS sample (<= n)
    const int n_mine = steps;
    if (loss_weights) {
        steps = 0;
        double run = 0.0;
        const double u = tp[0];
        bool found = false;
        for (int l = 0; l <= n; ++l) {
            run += loss_weights[l];
            if (run > u) { found = true; break; }
            ++steps;
        }
        if (!found) { steps = n; atomicOr(err_flag, 2); }   // the CDF never exceeds u (weights that do not sum to 1): the reference indexes past the end
    }
    steps_total[i] = steps;
    unsigned char *s = occ_s + i * m, *t = occ_t + i * m, *rem = remaining + i * (long long)n;
    for (int v = 0; v < m; ++v) { s[v] = 0; t[v] = 0; }
    int c = 0;
    for (int v = 0; v < m; ++v)
        for (int a = 0; a < s0[v]; ++a) rem[c++] = (unsigned char)v;
    int nr = n_mine;
    if (steps > 0) {
        int pick = (int)(tp[1] * (double)nr);
        if (pick >= nr) pick = nr - 1;
        const int mode = rem[pick];
        for (int q = pick; q + 1 < nr; ++q) rem[q] = rem[q + 1];
        --nr;
        s[mode] = 1;
    }
    n_remaining[i] = nr;
}

__global__ void k4_output_kernel(const unsigned char *__restrict__ occ_t, long long count, int *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = (int)occ_t[i];
}

// Per-sample matrices of the BOBS strategies, built where they are used (row f4 of SURVEY.md section 8):
//   Us[s][r][c] = B[r][perm_s(c)]                                          for c >= a
//   Us[s][r][c] = sum_{q < a} B[r][perm_s(q)] * phase_s[q] * QFT[q][c]      for c <  a
// i.e. (B with its columns permuted) @ diag(phases, 1 ...) @ (QFT on the first a modes) --
// nonuniform_losses_approximation_strategy.py:331-347 (B = dilation template, no permutation, a = approximated modes) and
// lossy_state_approximated_simulation_strategy.py:329-362 (B = the unitary, one column permutation per sample).
// One block per sample; a row's phased entries are staged in shared memory, the QFT comes through L1/L2.
__global__ void __launch_bounds__(256) k4_bobs_build_kernel(const double2 *__restrict__ B, int m, const double2 *__restrict__ qft, int a,
                                                            const double2 *__restrict__ phases, const int *__restrict__ perms,
                                                            double2 *__restrict__ Us) {
    extern __shared__ double2 bb_t[];                    // [rows_per_pass][a]
    const long long smp = blockIdx.x;
    const double2 *ph = phases + smp * a;
    const int *perm = perms ? perms + smp * m : nullptr;
    double2 *out = Us + smp * (long long)m * m;
    const int rows_per_pass = a > 0 ? max(1, min(m, 2048 / a)) : m;
    for (int r0 = 0; r0 < m; r0 += rows_per_pass) {
        const int nr = min(rows_per_pass, m - r0);
        for (int e = threadIdx.x; e < nr * a; e += blockDim.x) {
            const int r = e / a, q = e - r * a;
            const double2 b = B[(long long)(r0 + r) * m + (perm ? perm[q] : q)], p = ph[q];
            bb_t[e] = make_double2(b.x * p.x - b.y * p.y, b.x * p.y + b.y * p.x);
        }
        __syncthreads();
        for (int e = threadIdx.x; e < nr * m; e += blockDim.x) {
            const int r = e / m, c = e - r * m;
            double2 v;
            if (c >= a) v = B[(long long)(r0 + r) * m + (perm ? perm[c] : c)];
            else {
                double re = 0.0, im = 0.0;
                const double2 *t = bb_t + r * a;
                for (int q = 0; q < a; ++q) {
                    const double2 x = t[q], w = qft[q * a + c];
                    re = fma(x.x, w.x, re); re = fma(-x.y, w.y, re);
                    im = fma(x.x, w.y, im); im = fma(x.y, w.x, im);
                }
                v = make_double2(re, im);
            }
            out[(long long)(r0 + r) * m + c] = v;
        }
        __syncthreads();
    }
}

struct BobsBuild {             // host description of a device-side matrix build (all pointers HOST memory)
    const double *B;           // [m][m] complex
    const double *qft;         // [a][a] complex
    int a;
    const double *phases;      // [n_samples][a] complex
    const int32_t *perms;      // [n_samples][m] or NULL
};

// ---------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------
static int occ_to_u8(bp_context *h, const int32_t *v, int m, unsigned char *dst, long *sum, const char *who) {
    long n = 0;
    for (int i = 0; i < m; ++i) {
        if (v[i] < 0 || v[i] > 255) return bp_fail(h, BP_ERR_INVALID, "%s: occupation %d out of range", who, v[i]);
        dst[i] = (unsigned char)v[i];
        n += v[i];
    }
    *sum = n;
    return BP_OK;
}

// Device-side build of S per-sample matrices (samples [done, done + S) of the request) into dU.  d_B / d_qft: the resident
// operands (uploaded once per request by bobs_upload_resident), scratch slot MISC: this batch's phases and permutations.
static int bobs_upload_resident(bp_context *h, const BobsBuild &bb, int m, double *d_B, double *d_qft) {
    BP_CUDA(h, cudaMemcpyAsync(d_B, bb.B, sizeof(double) * 2 * (size_t)m * m, cudaMemcpyHostToDevice, h->stream));
    if (bb.a > 0) BP_CUDA(h, cudaMemcpyAsync(d_qft, bb.qft, sizeof(double) * 2 * (size_t)bb.a * bb.a, cudaMemcpyHostToDevice, h->stream));
    return BP_OK;
}
static int bobs_build_batch(bp_context *h, const BobsBuild &bb, int m, long long done, long long S, const double *d_B,
                            const double *d_qft, double *dU) {
    const size_t ph_bytes = sizeof(double) * 2 * (size_t)S * (size_t)(bb.a > 0 ? bb.a : 1);
    const size_t pm_bytes = bb.perms ? sizeof(int) * (size_t)S * m : 0;
    int rc = bp_reserve(h, BP_SLOT_MISC, ((ph_bytes + 15) / 16) * 16 + pm_bytes + 64);
    if (rc) return rc;
    double *d_ph = (double *)h->d_buf[BP_SLOT_MISC];
    int *d_pm = bb.perms ? (int *)((char *)d_ph + ((ph_bytes + 15) / 16) * 16) : nullptr;
    if (bb.a > 0) BP_CUDA(h, cudaMemcpyAsync(d_ph, bb.phases + (size_t)done * 2 * bb.a, sizeof(double) * 2 * (size_t)S * bb.a, cudaMemcpyHostToDevice, h->stream));
    if (bb.perms) BP_CUDA(h, cudaMemcpyAsync(d_pm, bb.perms + (size_t)done * m, pm_bytes, cudaMemcpyHostToDevice, h->stream));
    const int rows_per_pass = bb.a > 0 ? ((2048 / bb.a) < 1 ? 1 : ((2048 / bb.a) > m ? m : (2048 / bb.a))) : m;
    const size_t smem = sizeof(double) * 2 * (size_t)rows_per_pass * (size_t)(bb.a > 0 ? bb.a : 1);
    k4_bobs_build_kernel<<<(unsigned)S, 256, smem, h->stream>>>((const double2 *)d_B, m, (const double2 *)d_qft, bb.a, (const double2 *)d_ph,
                                                               d_pm, (double2 *)dU);
    BP_CHECK_LAUNCH(h);
    return BP_OK;
}

// minors (+ optional pmf) of ONE (s, t) pair; host pointers.
static int minors_host(bp_context *h, const double *U, int m, const int32_t *s, const int32_t *t, double *minors,
                       double *pmf, const char *who) {
    if (!h || !U || !s || !t) return bp_fail(h, BP_ERR_INVALID, "%s: NULL argument", who);
    if (m < 1 || m > BP_MAX_MODES) return bp_fail(h, BP_ERR_UNSUPPORTED, "%s: m=%d outside [1, %d]", who, m, BP_MAX_MODES);
    BP_ON_DEVICE(h);
    const size_t ub = sizeof(double) * 2 * (size_t)m * m;
    int rc;
    if ((rc = bp_reserve_pinned(h, 2 * (size_t)m + 64 + sizeof(double) * 3 * (size_t)m + 64))) return rc;
    unsigned char *hs = (unsigned char *)h->h_pin, *ht = hs + m;
    long k = 0, kt = 0;
    if ((rc = occ_to_u8(h, s, m, hs, &k, who))) return rc;
    if ((rc = occ_to_u8(h, t, m, ht, &kt, who))) return rc;
    if (k < 1) return bp_fail(h, BP_ERR_SHAPE, "%s: the input state holds no particle", who);
    if (kt != k - 1) return bp_fail(h, BP_ERR_SHAPE, "%s: sum(t) = %ld must equal sum(s) - 1 = %ld", who, kt, k - 1);
    if (k - 1 > BP_MAX_N || bp_k3_width((int)k) == 0) return bp_fail(h, BP_ERR_UNSUPPORTED, "%s: k = %ld too large", who, k);
    const int chunks = bp_k3_chunks(h, (int)k, 1), W = bp_k3_width((int)k);
    if ((rc = bp_reserve(h, BP_SLOT_AUX, ub))) return rc;
    if ((rc = bp_reserve(h, BP_SLOT_STATE, 2 * (size_t)m + 32))) return rc;
    if ((rc = bp_reserve(h, BP_SLOT_PARTIALS, sizeof(double) * 4 * (size_t)(W > 0 ? W : 1) * chunks))) return rc;
    if ((rc = bp_reserve(h, BP_SLOT_OUT, sizeof(double) * 3 * (size_t)m))) return rc;
    if ((rc = bp_reserve(h, BP_SLOT_MISC, 64))) return rc;
    unsigned long long *d_terms = (unsigned long long *)h->d_buf[BP_SLOT_MISC];
    int *d_flag = (int *)(d_terms + 2);
    BP_CUDA(h, cudaMemsetAsync(d_flag, 0, sizeof(int), h->stream));
    BP_CUDA(h, cudaMemcpyAsync(h->d_buf[BP_SLOT_AUX], U, ub, cudaMemcpyHostToDevice, h->stream));
    unsigned char *d_s = (unsigned char *)h->d_buf[BP_SLOT_STATE], *d_t = d_s + m;
    BP_CUDA(h, cudaMemcpyAsync(d_s, hs, 2 * (size_t)m, cudaMemcpyHostToDevice, h->stream));
    const double *dU = (const double *)h->d_buf[BP_SLOT_AUX];
    double *d_part = (double *)h->d_buf[BP_SLOT_PARTIALS];
    if ((rc = bp_k3_launch(h, dU, 0, m, d_s, d_t, nullptr, nullptr, (int)k, 1, chunks, d_part, d_terms))) return rc;
    double *d_min = (double *)h->d_buf[BP_SLOT_OUT], *d_pmf = d_min + 2 * (size_t)m;
    K3Finish a;
    memset(&a, 0, sizeof(a));
    a.U = dU; a.m = m; a.W = W; a.chunks = chunks; a.step = (int)k - 1; a.partials = d_part;
    a.terms = d_terms; a.per_block = bp_k3_per_block(h, (int)k, 1);
    a.occ_s = d_s; a.occ_t = d_t; a.minors_out = d_min; a.pmf_out = pmf ? d_pmf : nullptr;
    a.err_flag = d_flag;
    if ((rc = bp_k3_finish_launch(h, a, 1))) return rc;
    double *hres = (double *)((char *)h->h_pin + ((2 * (size_t)m + 63) / 64) * 64);
    int *h_flag = (int *)(hres + 3 * (size_t)m);
    BP_CUDA(h, cudaMemcpyAsync(hres, d_min, sizeof(double) * 3 * (size_t)m, cudaMemcpyDeviceToHost, h->stream));
    BP_CUDA(h, cudaMemcpyAsync(h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    BP_CUDA(h, cudaStreamSynchronize(h->stream));
    if (pmf && *h_flag) return bp_fail(h, BP_ERR_DOMAIN, "%s: the probabilities of the step do not sum to a positive finite number", who);
    if (minors) memcpy(minors, hres, sizeof(double) * 2 * (size_t)m);
    if (pmf) memcpy(pmf, hres + 2 * (size_t)m, sizeof(double) * (size_t)m);
    return BP_OK;
}

// The C ABI never throws: host-side allocation failures of the implementation become BP_ERR_NOMEM.
template <typename F>
static int no_throw(bp_context *h, const char *who, F &&body) {
    try {
        return body();
    } catch (const std::bad_alloc &) {
        return bp_fail(h, BP_ERR_NOMEM, "%s: out of host memory", who);
    } catch (...) {
        return bp_fail(h, BP_ERR_INVALID, "%s: unexpected C++ exception", who);
    }
}

extern "C" {

int bp_minors(bp_handle h, const double *U, int m, const int32_t *s, const int32_t *t, int formula, double *out) {
    if (!out) return bp_fail(h, BP_ERR_INVALID, "bp_minors: out is NULL");
    if (formula < BP_FORMULA_RYSER || formula > BP_FORMULA_GLYNN) return bp_fail(h, BP_ERR_INVALID, "bp_minors: formula %d", formula);
    return minors_host(h, U, m, s, t, out, nullptr, "bp_minors");
}

int bp_gccb_pmf(bp_handle h, const double *U, int m, const int32_t *s, const int32_t *t, double *pmf, double *minors_out) {
    if (!pmf) return bp_fail(h, BP_ERR_INVALID, "bp_gccb_pmf: pmf is NULL");
    return minors_host(h, U, m, s, t, minors_out, pmf, "bp_gccb_pmf");
}

// Shared implementation.  per_sample = false: one matrix U (m x m) and one input state s for all samples;
// per_sample = true: U is [n_samples][m][m] and s is [n_samples][m] (row f1 of SURVEY.md section 8f: the
// BOBS strategies draw a new matrix and a new lossy input state for every sample).
static int gccb_simulate_impl(bp_context *h, const double *U, int m, const int32_t *s, bool per_sample, int64_t n_samples,
                              double eta, uint64_t seed, int64_t first_sample, const double *tape, int tape_n,
                              int32_t *out, const char *who, const BobsBuild *bb = nullptr) {
    if (!h || (!U && !bb) || !s || !out) return bp_fail(h, BP_ERR_INVALID, "%s: NULL argument", who);
    if (m < 1 || m > BP_MAX_MODES) return bp_fail(h, BP_ERR_UNSUPPORTED, "%s: m=%d outside [1, %d]", who, m, BP_MAX_MODES);
    if (n_samples < 0) return bp_fail(h, BP_ERR_INVALID, "%s: n_samples=%lld", who, (long long)n_samples);
    if (eta > 1.0) return bp_fail(h, BP_ERR_INVALID, "%s: eta=%g > 1", who, eta);
    if (n_samples == 0) return BP_OK;
    BP_ON_DEVICE(h);
    const size_t n_states = per_sample ? (size_t)n_samples : 1;
    std::vector<unsigned char> s8(n_states * (size_t)m);
    int rc, n = 0;
    for (size_t i = 0; i < n_states; ++i) {
        long nl = 0;
        if ((rc = occ_to_u8(h, s + i * m, m, s8.data() + i * m, &nl, who))) return rc;
        if ((int)nl > n) n = (int)nl;
    }
    if (tape_n > n) n = tape_n;   // the caller's tape may be laid out for more particles than any sample holds
    if (n == 0) { memset(out, 0, sizeof(int32_t) * (size_t)n_samples * m); return BP_OK; }
    if (n - 1 > BP_MAX_N || bp_k3_width(n) == 0) return bp_fail(h, BP_ERR_UNSUPPORTED, "%s: n=%d too large", who, n);
    const int stride = 1 + 2 * n;

    // binomial weights C(n,l) eta^l (1-eta)^(n-l), same expression as the reference
    // (generalized_cliffords_b_uniform_losses_simulation_strategy.py:62-65)
    std::vector<double> weights;
    if (eta >= 0.0) {
        weights.resize(n + 1);
        for (int l = 0; l <= n; ++l) {
            double b = 1.0;
            for (int q = 1; q <= l; ++q) b = b * (double)(n - l + q) / (double)q;   // exact for these sizes
            b = nearbyint(b);
            weights[l] = b * pow(eta, (double)l) * pow(1.0 - eta, (double)(n - l));
        }
    }

    const size_t ub = sizeof(double) * 2 * (size_t)m * m;
    long long batch_cap = 32768;
    if (per_sample) {   // keep the per-batch matrix upload below ~1 GiB
        long long by_mem = (long long)((1ull << 30) / ub);
        if (by_mem < 1) by_mem = 1;
        if (by_mem < batch_cap) batch_cap = by_mem;
    }
    const long long batch = n_samples < batch_cap ? n_samples : batch_cap;
    int max_chunks = 1, maxW = 1;
    for (int k = 2; k <= n; ++k) {
        const int ch = bp_k3_chunks(h, k, batch), W = bp_k3_width(k);
        if ((long long)ch * W > (long long)max_chunks * maxW) { max_chunks = ch; maxW = W; }
    }
    const size_t u_count = per_sample ? (size_t)batch : 1;
    // Samples of a run stop after different numbers of steps when particles are lost (eta >= 0: binomial draw on the
    // device) or when every sample has its own input state: such runs launch only the samples still active at step k
    // (launch slot -> sample through `order`, samples sorted by their step count, longest first).
    const bool ragged = per_sample || eta >= 0.0;
    const size_t state_bytes = (size_t)batch * (2 * (size_t)m + (size_t)n + 20) + u_count * (size_t)m + (size_t)m + 256 + 16;
    const size_t bb_bytes = bb ? ub + sizeof(double) * 2 * (size_t)bb->a * bb->a + 32 : 0;
    if ((rc = bp_reserve(h, BP_SLOT_AUX, ub * u_count + sizeof(double) * (size_t)(n + 4) + bb_bytes))) return rc;
    if ((rc = bp_reserve(h, BP_SLOT_STATE, state_bytes))) return rc;
    if ((rc = bp_reserve(h, BP_SLOT_TAPE, sizeof(double) * (size_t)batch * stride))) return rc;
    if (!ragged && (rc = bp_reserve(h, BP_SLOT_PARTIALS, sizeof(double) * 4 * (size_t)maxW * max_chunks * (size_t)batch))) return rc;
    if ((rc = bp_reserve_pinned(h, (ragged ? sizeof(int) * 2 * (size_t)batch : 0) + 128))) return rc;
    if ((rc = bp_reserve(h, BP_SLOT_OUT, sizeof(int) * (size_t)batch * m))) return rc;

    double *dU = (double *)h->d_buf[BP_SLOT_AUX];
    double *d_w = dU + 2 * (size_t)m * m * u_count;
    double *d_B = d_w + ((n + 3) & ~1), *d_qft = d_B + 2 * (size_t)m * m;   // resident operands of the device-side matrix build (16-byte aligned)
    if (bb && (rc = bobs_upload_resident(h, *bb, m, d_B, d_qft))) return rc;
    if (!per_sample) BP_CUDA(h, cudaMemcpyAsync(dU, U, ub, cudaMemcpyHostToDevice, h->stream));
    if (eta >= 0.0) BP_CUDA(h, cudaMemcpyAsync(d_w, weights.data(), sizeof(double) * (n + 1), cudaMemcpyHostToDevice, h->stream));
    // state carve-up: 8-byte and 4-byte items first (alignment), then bytes
    char *base = (char *)h->d_buf[BP_SLOT_STATE];
    unsigned long long *d_terms = (unsigned long long *)base;
    int *d_nrem = (int *)(d_terms + batch);
    int *d_steps = d_nrem + batch;
    int *d_order = d_steps + batch;
    int *d_flag = d_order + batch;                       // bit 0: a step without a distribution, bit 1: particle-number draw out of range
    unsigned char *d_occ_s = (unsigned char *)(d_flag + 4);
    unsigned char *d_occ_t = d_occ_s + (size_t)batch * m;
    unsigned char *d_rem = d_occ_t + (size_t)batch * m;
    unsigned char *d_s0 = d_rem + (size_t)batch * n;
    if (!per_sample) BP_CUDA(h, cudaMemcpyAsync(d_s0, s8.data(), (size_t)m, cudaMemcpyHostToDevice, h->stream));
    double *d_tape = (double *)h->d_buf[BP_SLOT_TAPE];
    int *d_out = (int *)h->d_buf[BP_SLOT_OUT];
    std::vector<long long> active(n + 2, 0);   // active[k] = samples that take at least k steps
    const size_t u_stride = per_sample ? 2 * (size_t)m * m : 0, s0_stride = per_sample ? (size_t)m : 0;

    int *h_flag = (int *)((char *)h->h_pin + (ragged ? sizeof(int) * 2 * (size_t)batch : 0) + 64);
    for (long long done = 0; done < n_samples; done += batch) {
        const long long S = (n_samples - done < batch) ? (n_samples - done) : batch;
        BP_CUDA(h, cudaMemsetAsync(d_flag, 0, sizeof(int), h->stream));
        if (per_sample) {
            if (bb) { if ((rc = bobs_build_batch(h, *bb, m, done, S, d_B, d_qft, dU))) return rc; }
            else BP_CUDA(h, cudaMemcpyAsync(dU, U + (size_t)done * 2 * m * m, ub * (size_t)S, cudaMemcpyHostToDevice, h->stream));
            BP_CUDA(h, cudaMemcpyAsync(d_s0, s8.data() + (size_t)done * m, (size_t)S * m, cudaMemcpyHostToDevice, h->stream));
        }
        if (tape) {
            BP_CUDA(h, cudaMemcpyAsync(d_tape, tape + (size_t)done * stride, sizeof(double) * (size_t)S * stride,
                                       cudaMemcpyHostToDevice, h->stream));
        } else {
            const long long tot = S * stride;
            k4_fill_tape_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(d_tape, S, stride, seed, first_sample + done);
            BP_CHECK_LAUNCH(h);
        }
        k4_init_kernel<<<(unsigned)((S + 127) / 128), 128, 0, h->stream>>>(d_s0, s0_stride, m, n, eta >= 0.0 ? d_w : nullptr, d_tape, stride, S,
                                                                          d_occ_s, d_occ_t, d_rem, d_nrem, d_steps, d_flag);
        BP_CHECK_LAUNCH(h);
        for (int k = 0; k <= n + 1; ++k) active[k] = (k <= n) ? S : 0;
        if (ragged) {
            // step counts back to the host (4 bytes per sample), counting sort by step count (descending, stable)
            int *h_steps = (int *)h->h_pin, *h_order = h_steps + S;
            BP_CUDA(h, cudaMemcpyAsync(h_steps, d_steps, sizeof(int) * (size_t)S, cudaMemcpyDeviceToHost, h->stream));
            BP_CUDA(h, cudaStreamSynchronize(h->stream));
            std::vector<long long> first(n + 2, 0);
            for (long long i = 0; i < S; ++i) {
                const int st = h_steps[i];
                if (st < 0 || st > n) return bp_fail(h, BP_ERR_CUDA, "%s: sample %lld reports %d steps (n = %d)", who, i, st, n);
                first[st]++;
            }
            long long run = 0;
            for (int st = n; st >= 0; --st) { const long long c = first[st]; first[st] = run; run += c; active[st] = run; }
            active[0] = S; active[n + 1] = 0;
            for (long long i = 0; i < S; ++i) h_order[first[h_steps[i]]++] = (int)i;
            BP_CUDA(h, cudaMemcpyAsync(d_order, h_order, sizeof(int) * (size_t)S, cudaMemcpyHostToDevice, h->stream));
            size_t need = 1;
            for (int k = 2; k <= n; ++k) {
                if (active[k] == 0) break;
                const size_t v = (size_t)bp_k3_chunks(h, k, active[k]) * (size_t)bp_k3_width(k) * (size_t)active[k];
                if (v > need) need = v;
            }
            if ((rc = bp_reserve(h, BP_SLOT_PARTIALS, sizeof(double) * 4 * need))) return rc;
        }
        double *d_part = (double *)h->d_buf[BP_SLOT_PARTIALS];
        for (int k = 1; k <= n; ++k) {
            const long long A = active[k];   // samples that reach step k (all of them unless the run is ragged)
            if (A == 0) break;
            const int chunks = bp_k3_chunks(h, k, A), W = bp_k3_width(k);
            if ((rc = bp_k3_launch(h, dU, u_stride, m, d_occ_s, d_occ_t, d_steps, ragged ? d_order : nullptr, k, A, chunks, d_part, d_terms)))
                return rc;
            K3Finish a;
            memset(&a, 0, sizeof(a));
            a.U = dU; a.u_stride = u_stride; a.m = m; a.W = W; a.chunks = chunks; a.step = k - 1; a.partials = d_part;
            a.terms = d_terms; a.per_block = bp_k3_per_block(h, k, A);
            a.occ_s = d_occ_s; a.occ_t = d_occ_t;
            a.tape = d_tape; a.tape_stride = stride; a.remaining = d_rem; a.n_remaining = d_nrem; a.n = n;
            a.steps_total = d_steps;
            a.order = ragged ? d_order : nullptr;
            a.err_flag = d_flag;
            if ((rc = bp_k3_finish_launch(h, a, A))) return rc;
        }
        const long long cnt = S * m;
        k4_output_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, h->stream>>>(d_occ_t, cnt, d_out);
        BP_CHECK_LAUNCH(h);
        BP_CUDA(h, cudaMemcpyAsync(out + (size_t)done * m, d_out, sizeof(int) * (size_t)cnt, cudaMemcpyDeviceToHost, h->stream));
        BP_CUDA(h, cudaMemcpyAsync(h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        BP_CUDA(h, cudaStreamSynchronize(h->stream));
        if (*h_flag & 1) return bp_fail(h, BP_ERR_DOMAIN, "%s: the probabilities of a sampling step do not sum to a positive finite number "
                                        "(numpy.random.choice raises ValueError in the reference)", who);
        if (*h_flag & 2) return bp_fail(h, BP_ERR_DOMAIN, "%s: the particle-number weights sum to less than a drawn uniform (eta = %g)", who, eta);
    }
    return BP_OK;
}

int bp_gccb_simulate(bp_handle h, const double *U, int m, const int32_t *s, int64_t n_samples, double eta, uint64_t seed,
                     int64_t first_sample, const double *tape, int32_t *out) {
    return no_throw(h, "bp_gccb_simulate", [&] {
        return gccb_simulate_impl(h, U, m, s, false, n_samples, eta, seed, first_sample, tape, 0, out, "bp_gccb_simulate");
    });
}

int bp_gccb_simulate_batch(bp_handle h, const double *Us, int m, const int32_t *states, int64_t n_samples, uint64_t seed,
                           int64_t first_sample, const double *tape, int tape_particles, int32_t *out) {
    if (tape && tape_particles < 0) return bp_fail(h, BP_ERR_INVALID, "bp_gccb_simulate_batch: tape_particles=%d", tape_particles);
    return no_throw(h, "bp_gccb_simulate_batch", [&] {
        return gccb_simulate_impl(h, Us, m, states, true, n_samples, -1.0, seed, first_sample, tape, tape ? tape_particles : 0, out,
                                  "bp_gccb_simulate_batch");
    });
}

static int bobs_check(bp_context *h, const double *B, int m, const double *qft, int a, const double *phases, int64_t n, const char *who) {
    if (!h || !B) return bp_fail(h, BP_ERR_INVALID, "%s: NULL argument", who);
    if (m < 1 || m > BP_MAX_MODES) return bp_fail(h, BP_ERR_UNSUPPORTED, "%s: m=%d outside [1, %d]", who, m, BP_MAX_MODES);
    if (a < 0 || a > m) return bp_fail(h, BP_ERR_INVALID, "%s: a=%d outside [0, m=%d]", who, a, m);
    if (a > 0 && (!qft || (n > 0 && !phases))) return bp_fail(h, BP_ERR_INVALID, "%s: NULL qft / phases with a=%d", who, a);
    if (n < 0) return bp_fail(h, BP_ERR_INVALID, "%s: n_samples=%lld", who, (long long)n);
    return BP_OK;
}

int bp_gccb_simulate_bobs(bp_handle h, const double *B, int m, const double *qft, int a, const double *phases, const int32_t *perms,
                          const int32_t *states, int64_t n_samples, uint64_t seed, int64_t first_sample, const double *tape,
                          int tape_particles, int32_t *out) {
    int rc = bobs_check(h, B, m, qft, a, phases, n_samples, "bp_gccb_simulate_bobs");
    if (rc) return rc;
    if (tape && tape_particles < 0) return bp_fail(h, BP_ERR_INVALID, "bp_gccb_simulate_bobs: tape_particles=%d", tape_particles);
    return no_throw(h, "bp_gccb_simulate_bobs", [&] {
        BobsBuild bb = {B, qft, a, phases, perms};
        return gccb_simulate_impl(h, nullptr, m, states, true, n_samples, -1.0, seed, first_sample, tape, tape ? tape_particles : 0, out,
                                  "bp_gccb_simulate_bobs", &bb);
    });
}

int bp_bobs_build(bp_handle h, const double *B, int m, const double *qft, int a, const double *phases, const int32_t *perms,
                  int64_t n_samples, double *Us_out) {
    int rc = bobs_check(h, B, m, qft, a, phases, n_samples, "bp_bobs_build");
    if (rc) return rc;
    if (!Us_out) return bp_fail(h, BP_ERR_INVALID, "bp_bobs_build: NULL argument");
    if (n_samples == 0) return BP_OK;
    BP_ON_DEVICE(h);
    const size_t ub = sizeof(double) * 2 * (size_t)m * m;
    long long batch = (long long)((256ull << 20) / ub);
    if (batch < 1) batch = 1;
    if (batch > n_samples) batch = n_samples;
    if ((rc = bp_reserve(h, BP_SLOT_AUX, ub * (size_t)batch + ub + sizeof(double) * 2 * (size_t)a * a + 64))) return rc;
    double *dU = (double *)h->d_buf[BP_SLOT_AUX], *d_B = dU + 2 * (size_t)m * m * (size_t)batch, *d_qft = d_B + 2 * (size_t)m * m;
    BobsBuild bb = {B, qft, a, phases, perms};
    if ((rc = bobs_upload_resident(h, bb, m, d_B, d_qft))) return rc;
    for (long long done = 0; done < n_samples; done += batch) {
        const long long S = n_samples - done < batch ? n_samples - done : batch;
        if ((rc = bobs_build_batch(h, bb, m, done, S, d_B, d_qft, dU))) return rc;
        BP_CUDA(h, cudaMemcpyAsync(Us_out + (size_t)done * 2 * m * m, dU, ub * (size_t)S, cudaMemcpyDeviceToHost, h->stream));
        BP_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return BP_OK;
}

}  // extern "C"
