/*
 * bossperm.h -- C ABI of libbossperm.so: B200 (sm_100a) kernels for the permanent hot path of
 * the boson-sampling simulator Tomev-CTP/theboss (v3.0.1).
 *
 * The reference is pure Python and has no FFI; its boundary for this path is the Python class
 * API (compute_permanent / compute_permanents / simulate).  Each entry point below is what a
 * ctypes binding placed *directly beneath* one reference method would call; the method it
 * replaces is cited as file:line relative to /root/reference/theboss/.  INTEGRATION.md shows
 * the reference-side stubs.
 *
 * Conventions
 *   - Plain C types only.  Complex numbers are interleaved (re, im) float64, i.e. the memory of
 *     a C-contiguous numpy.complex128 array.  Matrices are row-major U[out_mode][in_mode]
 *     (row = output mode, column = input mode: boson_sampling_utilities.py:595-626).
 *   - Unless a function name ends in `_dev`, every pointer is a HOST pointer owned by the
 *     caller and borrowed for the duration of the call; the call copies inputs to the device,
 *     runs, copies results back and returns after the stream has drained.
 *   - `_dev` variants take DEVICE pointers, enqueue on the handle's stream and return without
 *     synchronising (inputs resident in HBM; used by bench.py's `value` leg and by the
 *     multi-GPU path, which owns device buffers through torch).
 *   - Every function returns 0 on success or a negative bp_status; nothing throws.
 *     bp_last_error(h) returns a NUL-terminated description of the last failure on h.
 *   - A bp_handle is one device + one stream + scratch buffers.  Not thread-safe, not
 *     re-entrant (the reference calculators are not either: they mutate instance state during
 *     a compute, bs_permanent_calculator_base.py:104-121).  Cheap to create and destroy.
 *   - There is no CPU fallback: without a CUDA device bp_create fails with BP_ERR_CUDA.
 */
#ifndef BOSSPERM_H
#define BOSSPERM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BP_ABI_VERSION 1
#define BP_MAX_N 40        /* largest explicit matrix / particle number of the register kernels */
#define BP_MAX_MODES 256   /* largest interferometer dimension m */
#define BP_MAX_PEERS 16     /* most ranks (GPUs of one node) of a peer-memory exchange */

typedef struct bp_context *bp_handle;

typedef enum {
    BP_OK = 0,
    BP_ERR_INVALID = -1,      /* bad argument (NULL pointer, negative size, ...) */
    BP_ERR_SHAPE = -2,        /* shapes do not match: the reference raises AttributeError here
                                 (bs_permanent_calculator_base.py:61-72, :179-180) */
    BP_ERR_UNSUPPORTED = -3,  /* size beyond BP_MAX_N / BP_MAX_MODES */
    BP_ERR_CUDA = -4,         /* CUDA runtime failure, see bp_last_error */
    BP_ERR_NOMEM = -5,
    BP_ERR_DOMAIN = -6        /* a step's probabilities do not sum to a positive finite number (e.g. an all-zero column of U):
                               * numpy.random.choice raises ValueError in the reference's _sample_from_pmf; so does the binding */
} bp_status;

/* formula selector of the multiplicity / minors entry points.  All three name the same
 * mathematical quantity; RYSER and CHIN_HUH mirror PermanentCalculatorType
 * (bs_permanent_calculator_factory.py:29-33).  The device engine always evaluates the
 * Glynn/Chin-Huh form (Ryser-form float64 arithmetic cannot meet 1e-10 beyond n~22,
 * SURVEY.md Appendix C); the selector is kept so callers can pass the reference's enum. */
#define BP_FORMULA_RYSER 0
#define BP_FORMULA_CHIN_HUH 1
#define BP_FORMULA_GLYNN 2

int bp_abi_version(void);

/* ---- lifetime ------------------------------------------------------------------------------ */
int bp_create(int device, bp_handle *out);
/* Same, but enqueue on a caller-owned cudaStream_t (e.g. torch's current stream) so that the
 * caller's CUDA events bracket this library's kernels. */
int bp_create_on_stream(int device, void *cuda_stream, bp_handle *out);
int bp_destroy(bp_handle h);
const char *bp_last_error(bp_handle h);
int bp_synchronize(bp_handle h);
/* sm_count, compute capability major/minor, SM clock (kHz) of the handle's device. */
int bp_device_info(bp_handle h, int *sm_count, int *cc_major, int *cc_minor, int *clock_khz);
/* Number of kernels this handle has launched since creation (bench.py `gpu_launches`). */
int64_t bp_launch_count(bp_handle h);

/* CUDA-event timer on the handle's stream (bench.py roofline leg). */
int bp_timer_start(bp_handle h);
int bp_timer_stop(bp_handle h, float *elapsed_ms);

/* FP64 roofline denominator: runs a dependent-chain-free DFMA kernel on every SM for about
 * `target_ms` and reports the achieved FP64 rate (2 flops per DFMA).  */
int bp_fp64_peak(bp_handle h, double target_ms, double *tflops);

/* ---- K1: Gray-code Glynn on an explicit N x N matrix ---------------------------------------
 * Replaces GlynnGrayPermanentCalculator.compute_permanent after the effective matrix has been
 * built (permanent_calculators/glynn_gray_permanent_calculator.py:55-71).
 *   perm(A) = 2^-(N-1) * sum over the 2^(N-1) Gray steps (step 0 = all-ones delta).
 * N == 0 returns 1 (:52-53). */
int bp_glynn_matrix(bp_handle h, const double *A, int N, double out[2]);

/* Shardable partial: UN-normalised sum over Gray steps [step_lo, step_hi) of the 2^(N-1) terms,
 * as a double-double complex {re_hi, re_lo, im_hi, im_lo}.  Summing the partials of a disjoint
 * cover of [0, 2^(N-1)) in double-double and scaling by 2^-(N-1) gives bp_glynn_matrix. */
int bp_glynn_matrix_range(bp_handle h, const double *A, int N, uint64_t step_lo, uint64_t step_hi,
                          double out_dd[4]);
int bp_glynn_matrix_range_dev(bp_handle h, const double *dA, int N, uint64_t step_lo,
                              uint64_t step_hi, double *d_out_dd);

/* Declares the device matrix dA RESIDENT: its contents stay as they are until the next bp_glynn_set_resident call on this
 * handle (same pointer again = "the contents changed", NULL = no resident matrix).  Range calls on a resident matrix reuse the
 * kernel's constant-bank image of it instead of copying it again at every launch (12 us per call).  No reference counterpart:
 * the reference rebuilds its matrix view per call (glynn_gray_permanent_calculator.py:41-53). */
int bp_glynn_set_resident(bp_handle h, const double *dA);

/* ---- K1 sharded over the GPUs of one node: partial exchange over peer memory (NVLink) --------------------------------
 * One process per GPU.  The reference has no distributed path; this is the exchange step of the sharded
 * Glynn permanent of BASELINE.json configs[3]: every rank evaluates a slice of the Gray range and needs all ranks' 32-byte
 * double-double partials.  bp_exchange_create allocates this rank's slot buffer and returns its CUDA IPC handle; the caller
 * all-gathers the `world` handles (64 bytes each, rank order) by any means and passes them to bp_exchange_connect, which maps
 * the peers' buffers.  bp_glynn_matrix_range_exchange then is bp_glynn_matrix_range_dev fused with the exchange: the LAST block
 * of the kernel stores this rank's partial straight into every peer's slot buffer, waits for the `world` partials of this call
 * to arrive in its own, and writes them in rank order to d_out_all[world][4].  Collective: every rank must make the same
 * sequence of calls.  A peer that does not arrive within 10 s leaves NaNs in its row of d_out_all. */
int bp_exchange_create(bp_handle h, int world, int rank, unsigned char ipc_handle_out[64]);
int bp_exchange_connect(bp_handle h, const unsigned char *ipc_handles /* [world][64] */);
int bp_exchange_destroy(bp_handle h);
int bp_glynn_matrix_range_exchange(bp_handle h, const double *dA, int N, uint64_t step_lo, uint64_t step_hi,
                                   double *d_out_all);
/* The same collective with HOST buffers, as one call (what a sharded GlynnGrayPermanentCalculator.compute_permanent pays per
 * permanent, glynn_gray_permanent_calculator.py:41-71): A (N x N, row-major complex128 on the host) goes through the handle's pinned
 * staging buffer to the device, the kernel runs this rank's slice and the exchange, and all ranks' partials come back to
 * out_all[world][4] (host); returns after the stream has drained.  With world == 1 it equals bp_glynn_matrix_range. */
int bp_glynn_matrix_range_exchange_host(bp_handle h, const double *A, int N, uint64_t step_lo, uint64_t step_hi,
                                        double *out_all);

/* Full calculator call: builds the effective scattering matrix of (U, s, t) on the device
 * (boson_sampling_utilities.py:595-626) and evaluates it.  s, t: length m occupations.
 * Returns BP_ERR_SHAPE when sum(s) != sum(t); 1 when either side is empty (:52-53). */
int bp_glynn_single(bp_handle h, const double *U, int m, const int32_t *s, const int32_t *t,
                    double out[2]);

/* ---- K2: batched permanents with input/output multiplicities -------------------------------
 * Replaces B calls of {Ryser,ChinHuh,GlynnGray}PermanentCalculator.compute_permanent
 * (bs_permanent_calculator_base.py:166-209, ryser_permanent_calculator.py:45-64,
 * chin_huh_permanent_calculator.py:38-59) that share one interferometer U.
 * S, T: B x m occupation tables (uint8, row-major); out: B complex.
 * Item b with sum(S[b]) != sum(T[b]) makes the call fail with BP_ERR_SHAPE before any launch;
 * an item without particles yields 1. */
int bp_perm_batched(bp_handle h, const double *U, int m, const uint8_t *S, const uint8_t *T,
                    int64_t B, int formula, double *out);
int bp_perm_batched_dev(bp_handle h, const double *dU, int m, const uint8_t *dS, const uint8_t *dT,
                        int64_t B, int formula, double *d_out);

/* ---- K3: all one-input-particle-removed minors of one GCC-B step ---------------------------
 * Replaces BSCC{Ryser,CH}SubmatricesPermanentCalculator.compute_permanents
 * (bs_submatrices_permanent_calculator_base.py:150-175): sum(s) = k, sum(t) = k-1;
 * out[i] = perm with one particle removed from input mode i, 0 where s[i] == 0; for k == 1 the
 * occupations themselves as complex (:157-158).  out: m complex. */
int bp_minors(bp_handle h, const double *U, int m, const int32_t *s, const int32_t *t, int formula,
              double *out);

/* Minors + Laplace combine: GeneralizedCliffordsBSimulationStrategy._compute_pmf
 * (simulation_strategies/generalized_cliffords_b_simulation_strategy.py:69-92):
 * pmf[j] = |sum_i s_i P_i U[j][i]|^2 / total, j = 0..m-1.  minors_out may be NULL. */
int bp_gccb_pmf(bp_handle h, const double *U, int m, const int32_t *s, const int32_t *t,
                double *pmf, double *minors_out);

/* ---- K3+K4: device-resident GCC-B sampling loops --------------------------------------------
 * Replaces GeneralizedCliffordsBSimulationStrategy.simulate (:41-67, :94-110) and its
 * uniform-loss subclass (generalized_cliffords_b_uniform_losses_simulation_strategy.py:50-121).
 *
 * Decision tape.  The reference draws from numpy's global generator:
 * per step one randint (which remaining input particle enters, :102-105) and the single uniform
 * numpy.random.choice consumes (:107-110); the uniform-loss variant first draws one uniform for
 * the surviving particle number (:67-85).  Here every decision comes from
 *     tape[sample * tape_stride + 0]          particle-number uniform (uniform-loss only)
 *     tape[sample * tape_stride + 1 + 2*k]    u_pick   -> index floor(u_pick * #remaining)
 *     tape[sample * tape_stride + 2 + 2*k]    u_choice -> searchsorted(cdf/cdf[-1], u, 'right')
 * with tape_stride = 1 + 2*n.  tape == NULL: the library fills the tape on the device from a
 * counter-based generator keyed by (seed, sample, slot), so results do not depend on how
 * samples are split across GPUs (first_sample offsets the counter).
 *
 *   eta < 0      plain GCC-B: all n particles are sampled.
 *   0<=eta<=1    uniform losses: l ~ Binomial(n, eta) by inverse CDF, then l steps.
 * out: n_samples x m int32 occupations. */
int bp_gccb_simulate(bp_handle h, const double *U, int m, const int32_t *s, int64_t n_samples,
                     double eta, uint64_t seed, int64_t first_sample, const double *tape,
                     int32_t *out);

/* One interferometer AND one input state per sample: what the BOBS strategies need, which draw fresh
 * random phases (a new matrix) and a fresh lossy input state for every sample and then take ONE GCC-B sample
 * (simulation_strategies/nonuniform_losses_approximation_strategy.py:263-296,
 *  simulation_strategies/lossy_state_approximated_simulation_strategy.py:287-310).
 *   Us:     n_samples x m x m complex, states: n_samples x m occupations (particle numbers may differ).
 *   tape:   NULL (Philox as above) or n_samples x (1 + 2 * tape_particles) uniforms,
 *           tape_particles >= the largest particle number of any sample. */
int bp_gccb_simulate_batch(bp_handle h, const double *Us, int m, const int32_t *states, int64_t n_samples,
                           uint64_t seed, int64_t first_sample, const double *tape, int tape_particles,
                           int32_t *out);

/* The same, with the per-sample matrices built on the device instead of shipped (16 m^2 bytes per
 * sample): sample i runs on
 *     Us[i] = (B with its columns permuted by perms[i]) @ diag(phases[i], 1, ..., 1) @ (QFT on the first a modes),
 * what NonuniformLossesApproximationStrategy builds per sample (M0 @ random_phases @ QFT in its 2m-mode dilation,
 * nonuniform_losses_approximation_strategy.py:331-347: B = the dilation template, perms = NULL, a = approximated modes) and
 * LossyStateApproximationSimulationStrategy (U[:, random permutation] @ random_phases @ QFT,
 * lossy_state_approximated_simulation_strategy.py:329-362: B = U, a = m - hierarchy_level).
 * B: [m][m] complex; qft: [a][a] complex; phases: [n_samples][a] complex (unit modulus); perms: [n_samples][m] column indices or NULL.
 * bp_bobs_build returns the matrices themselves ([n_samples][m][m] complex) for inspection. */
int bp_gccb_simulate_bobs(bp_handle h, const double *B, int m, const double *qft, int a, const double *phases,
                          const int32_t *perms, const int32_t *states, int64_t n_samples, uint64_t seed,
                          int64_t first_sample, const double *tape, int tape_particles, int32_t *out);
int bp_bobs_build(bp_handle h, const double *B, int m, const double *qft, int a, const double *phases,
                  const int32_t *perms, int64_t n_samples, double *Us_out);

#ifdef __cplusplus
}
#endif
#endif /* BOSSPERM_H */
