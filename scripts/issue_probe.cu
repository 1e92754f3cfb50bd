// issue_probe.cu -- what does a non-FP64 instruction cost inside an FP64 stream on B200?
// Body: 64 independent DFMAs (16 chains x 4) + COUNT instructions of one KIND, either clustered at the top of the
// iteration or spread evenly between the DFMAs.  384 threads x 1 block per SM (3 warps/SMSP, like K1).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__constant__ double2 cbuf[512];

// KIND: 0 none, 1 LDS.128, 2 LDS.64, 3 LDS.32, 4 independent IADD/LOP, 5 constant-bank operand (LDCU.128), 6 FADD (fp32 pipe)
template <int KIND, int COUNT, bool SPREAD, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) probe(int iters, double *sink, int salt) {
    __shared__ double2 sm[1024];
    for (int e = threadIdx.x; e < 1024; e += THREADS) sm[e] = make_double2(1e-12 * e, 1e-13 * e);
    __syncthreads();
    double a[16];
    const double x = 1.0 + 1e-9 * threadIdx.x;
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = k * 1e-3;
    unsigned iv[8];
    float fv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { iv[k] = threadIdx.x + k; fv[k] = 0.5f * k; }
    unsigned row = salt & 31;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        const double2 *p = sm + row * 32;
        double y[64];
#pragma unroll
        for (int q = 0; q < 64; ++q) y[q] = 1e-12;
        if (!SPREAD) {
#pragma unroll
            for (int c = 0; c < COUNT; ++c) {
                if (KIND == 1) { double2 v = p[c]; y[2 * c % 64] = v.x; y[(2 * c + 1) % 64] = v.y; }
                if (KIND == 2) { double v = reinterpret_cast<const double *>(p)[c]; y[c % 64] = v; }
                if (KIND == 3) { float v = reinterpret_cast<const float *>(p)[c]; fv[c & 7] += v; }
                if (KIND == 4) { iv[c & 7] = (iv[c & 7] ^ (row + c)) + 0x9e37u; }
                if (KIND == 5) { double2 v = cbuf[c]; y[2 * c % 64] = v.x; y[(2 * c + 1) % 64] = v.y; }
                if (KIND == 6) { fv[c & 7] = fv[c & 7] + 1.25f; }
            }
        }
#pragma unroll
        for (int q = 0; q < 64; ++q) {
            if (SPREAD && COUNT > 0 && (q % (64 / (COUNT > 64 ? 64 : COUNT))) == 0 && (q / (64 / (COUNT > 64 ? 64 : COUNT))) < COUNT) {
                const int c = q / (64 / (COUNT > 64 ? 64 : COUNT));
                if (KIND == 1) { double2 v = p[c]; y[(q + 8) % 64] = v.x; y[(q + 9) % 64] = v.y; }
                if (KIND == 2) { double v = reinterpret_cast<const double *>(p)[c]; y[(q + 8) % 64] = v; }
                if (KIND == 3) { float v = reinterpret_cast<const float *>(p)[c]; fv[c & 7] += v; }
                if (KIND == 4) { iv[c & 7] = (iv[c & 7] ^ (row + c)) + 0x9e37u; }
                if (KIND == 5) { double2 v = cbuf[c]; y[(q + 8) % 64] = v.x; y[(q + 9) % 64] = v.y; }
                if (KIND == 6) { fv[c & 7] = fv[c & 7] + 1.25f; }
            }
            a[q & 15] = fma(a[q & 15], x, y[q]);
        }
        row = (row + 1) & 31;
    }
    double r = 0.0;
#pragma unroll
    for (int k = 0; k < 16; ++k) r += a[k];
    unsigned s = 0; float f = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { s ^= iv[k]; f += fv[k]; }
    if (r == 123.456 || s == 0x12345u || f == 7.77f) sink[threadIdx.x] = r;
}

template <int KIND, int COUNT, bool SPREAD>
static void run(const char *name, int sms, double *sink) {
    const int iters = 60000, T = 384;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    probe<KIND, COUNT, SPREAD, T><<<sms, T>>>(iters / 10, sink, 3);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        probe<KIND, COUNT, SPREAD, T><<<sms, T>>>(iters, sink, 3);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    const double tf = 2.0 * sms * T * 64.0 * iters / (best * 1e-3) / 1e12;
    // cycles per iteration per SMSP at 1.965 GHz: 3 warps per SMSP, each doing one iteration
    const double cyc = best * 1e-3 * 1.965e9 / iters / 3.0;
    printf("%-34s %8.3f ms %7.2f TFLOP/s  %6.1f cyc/warp-iter (64 DFMA = 128)  extra %.1f cyc -> %.2f cyc per extra instr\n", name, best, tf, cyc,
           cyc - 128.0, COUNT ? (cyc - 128.0) / COUNT : 0.0);
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    double *sink; CK(cudaMalloc(&sink, 8192));
    run<0, 0, false>("64 DFMA", sms, sink);
    run<1, 8, false>("+ 8 LDS.128 clustered", sms, sink);   run<1, 8, true>("+ 8 LDS.128 spread", sms, sink);
    run<1, 16, false>("+16 LDS.128 clustered", sms, sink);  run<1, 16, true>("+16 LDS.128 spread", sms, sink);
    run<2, 16, false>("+16 LDS.64 clustered", sms, sink);   run<2, 16, true>("+16 LDS.64 spread", sms, sink);
    run<3, 16, false>("+16 LDS.32 clustered", sms, sink);   run<3, 16, true>("+16 LDS.32 spread", sms, sink);
    run<4, 16, false>("+16 int ops clustered", sms, sink);  run<4, 16, true>("+16 int ops spread", sms, sink);
    run<4, 32, false>("+32 int ops clustered", sms, sink);  run<4, 32, true>("+32 int ops spread", sms, sink);
    run<5, 16, false>("+16 const operands clustered", sms, sink); run<5, 16, true>("+16 const operands spread", sms, sink);
    run<6, 16, false>("+16 FADD clustered", sms, sink);     run<6, 16, true>("+16 FADD spread", sms, sink);
    run<6, 32, true>("+32 FADD spread", sms, sink);
    return 0;
}
