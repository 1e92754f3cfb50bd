// minors.cuh -- interface between minors_kernel.cu (K3) and sampler_kernel.cu (K4 + C ABI).
#pragma once
#include "bp_common.cuh"

struct K3Finish {
    const double *U; size_t u_stride;                          // per-sample matrix stride in doubles (0 = shared)
    int m; int W; int chunks; int step;                        // step = k - 1 (0-based)
    const double *partials;
    const unsigned long long *terms; unsigned long long per_block; // per-sample walk length (written by K3), terms served by one chunk block
    unsigned char *occ_s, *occ_t;                              // [samples][m]
    double *minors_out;                                        // NULL or [samples][m] complex
    double *pmf_out;                                           // NULL or [samples][m]
    // sampling state (all NULL for the calculator entry points)
    const double *tape; int tape_stride;                       // [samples][1 + 2n]
    unsigned char *remaining; int *n_remaining; int n;         // [samples][n], [samples]
    const int *steps_total;                                    // [samples]
    // launch slot -> sample (NULL: identity).  Runs whose samples stop after different numbers of steps launch only
    // the samples still active; partials are indexed by launch slot, every per-sample array by sample.
    const int *order;
    // NULL or one int: bit 0 is set when a step's probabilities do not sum to a positive finite number (the reference's
    // numpy.random.choice raises there; the entry points turn the flag into BP_ERR_DOMAIN after their final synchronisation)
    int *err_flag;
};

int bp_k3_width(int k);
int bp_k3_chunks(bp_context *h, int k, long long samples);
unsigned long long bp_k3_per_block(bp_context *h, int k, long long samples);
int bp_k3_launch(bp_context *h, const double *dU, size_t u_stride, int m, const unsigned char *d_s, const unsigned char *d_t,
                 const int *d_steps_total, const int *d_order, int k, long long samples, int chunks, double *d_partials,
                 unsigned long long *d_terms);
int bp_k3_finish_launch(bp_context *h, const K3Finish &a, long long samples);
