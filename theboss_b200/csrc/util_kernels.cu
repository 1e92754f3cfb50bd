// util_kernels.cu -- small helper kernels: effective scattering matrix expansion, FP64 peak probe.
#include "bp_common.cuh"

// ---------------------------------------------------------------------------------------------
// Effective scattering matrix on the device.
// Replaces EffectiveScatteringMatrixCalculator.calculate
// (reference: theboss/boson_sampling_utilities/boson_sampling_utilities.py:595-626):
// columns of U repeated by the input occupation s, rows by the output occupation t.
// One block; thread (r, c) of the N x N result looks its source mode up in the expanded
// mode-assignment lists built in shared memory.
// ---------------------------------------------------------------------------------------------
__global__ void effective_matrix_kernel(const double *__restrict__ U, int m, const int32_t *__restrict__ s,
                                        const int32_t *__restrict__ t, int N, double *__restrict__ A) {
    __shared__ int16_t row_mode[BP_MAX_N], col_mode[BP_MAX_N];
    if (threadIdx.x == 0) {
        int r = 0, c = 0;
        for (int j = 0; j < m; ++j) {
            for (int a = 0; a < t[j] && r < N; ++a) row_mode[r++] = (int16_t)j;
            for (int a = 0; a < s[j] && c < N; ++a) col_mode[c++] = (int16_t)j;
        }
    }
    __syncthreads();
    const double2 *U2 = reinterpret_cast<const double2 *>(U);
    double2 *A2 = reinterpret_cast<double2 *>(A);
    for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
        const int r = e / N, c = e - r * N;
        A2[e] = U2[(int)row_mode[r] * m + (int)col_mode[c]];
    }
}

int bp_effective_matrix_launch(bp_context *h, const double *dU, int m, const int32_t *d_s, const int32_t *d_t,
                               int N, double *dA) {
    effective_matrix_kernel<<<1, 256, 0, h->stream>>>(dU, m, d_s, d_t, N, dA);
    BP_CHECK_LAUNCH(h);
    return BP_OK;
}

// ---------------------------------------------------------------------------------------------
// FP64 peak probe: 16 independent DFMA chains per thread, 64 DFMAs per loop iteration, 256 threads,
// 8 blocks per SM.  On B200 this reaches 37.1 TFLOP/s = 148 SM x 64 DFMA/clk x 1.965 GHz (a loop body of
// only 16 DFMAs loses 11 % to the three loop-control instructions, see profiles/r01_k1_explore.txt).
// The measured rate (2 flops per DFMA) is the roofline denominator bench.py reports against.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fp64_peak_kernel(int iters, double *__restrict__ sink) {
    double a[16];
    const double x = 1.0 + 1e-9 * (double)threadIdx.x, y = 1e-12 * (double)(blockIdx.x + 1);
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = (double)k * 1e-3;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < 16; ++k) a[k] = fma(a[k], x, y);
    }
    double r = 0.0;
#pragma unroll
    for (int k = 0; k < 16; ++k) r += a[k];
    if (r == 123.456) sink[threadIdx.x] = r;   // keeps the chains alive; practically never true
}

int bp_fp64_peak_launch(bp_context *h, int iters, double *d_sink) {
    fp64_peak_kernel<<<h->sm_count * 8, 256, 0, h->stream>>>(iters, d_sink);
    BP_CHECK_LAUNCH(h);
    return BP_OK;
}
