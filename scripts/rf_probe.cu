// rf_probe.cu -- does register-file read bandwidth bound an FP64 stream on B200?
// 64 FP64 instructions per iteration over 16 independent accumulators; the MODE decides how many DISTINCT vector-register
// source operands an instruction reads (operands repeated from the previous instruction can come from the operand reuse cache):
//   0  a = fma(a, x, y)        one distinct source (x, y shared by every instruction)
//   1  a = fma(b_i, c_j, a)    three distinct sources
//   2  a = fma(b_i, x, a)      two distinct sources
//   3  a = a * b_i             DMUL, two distinct sources
//   4  a = a + b_i             DADD, two distinct sources
//   5  a = fma(b_i, c_j, a) in PAIRS that share b_i (complex-multiply pattern: re/im of one product share one factor)
//   7  like 0 with an integer add after every second DFMA (does the reuse cache survive an instruction of another pipe?)
//   8  like 1 with 8 accumulators
//   9  like 4, and every fourth DADD is followed by a 16-byte shared-memory load (same address in all lanes) that refills two b's
//  10  like 9 with a different address per lane (conflict-free, four wavefronts per load)
//  11  like 9 with an 8-byte load (one b) after every second DADD
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int MODE, int Q>
__device__ __forceinline__ void lds_body(double (&a)[16], double (&b)[16], unsigned saddr) {
    if constexpr (Q < 64) {
        constexpr int i = Q & 15, j = (Q * 5 + 3) & 15;
        asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(a[i]) : "d"(b[j]));
        if constexpr (MODE != 11 && (Q & 3) == 3)
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(b[(Q >> 1) & 14]), "=d"(b[((Q >> 1) & 14) + 1]) : "r"(saddr), "n"((Q >> 2) * 16));
        if constexpr (MODE == 11 && (Q & 1) == 1)
            asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(b[(Q >> 1) & 15]) : "r"(saddr), "n"((Q >> 1) * 8));
        lds_body<MODE, Q + 1>(a, b, saddr);
    }
}

template <int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) probe(int iters, double *sink, double seed) {
    double a[16], b[16], c[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { a[k] = k * 1e-3 + seed; b[k] = 1.0 + 1e-9 * (threadIdx.x + k) + seed; c[k] = 1e-12 * (k + 1) + seed; }
    const double x = 1.0 + 1e-9 * threadIdx.x + seed, y = 1e-12 + seed;
    unsigned iv = threadIdx.x;
    __shared__ double2 img[1024];
    for (int e = threadIdx.x; e < 1024; e += THREADS) img[e] = make_double2(1e-9 * e, 1e-10 * e);
    __syncthreads();
    unsigned sbase = (unsigned)__cvta_generic_to_shared(img) + ((MODE == 10) ? (threadIdx.x & 31) * 16u : 0u);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        const unsigned saddr = sbase + ((unsigned)(it & 7) << 9);
        if constexpr (MODE >= 9) lds_body<MODE, 0>(a, b, saddr);
#pragma unroll
        for (int q = 0; q < (MODE >= 9 ? 0 : 64); ++q) {
            const int i = q & 15, j = (q * 5 + 3) & 15, k = (q * 7 + 1) & 15;
            if (MODE == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(x), "d"(y));
            if (MODE == 1) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a[i]) : "d"(b[j]), "d"(c[k]));
            if (MODE == 2) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a[i]) : "d"(b[j]), "d"(x));
            if (MODE == 3) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(a[i]) : "d"(b[j]));
            if (MODE == 4) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(a[i]) : "d"(b[j]));
            if (MODE == 5) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a[i]) : "d"(b[(q >> 1) & 15]), "d"(c[k]));
            if (MODE == 7) {
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(x), "d"(y));
                if (q & 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(iv) : "r"(q + 1));
            }
            if (MODE == 8) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a[q & 7]) : "d"(b[j]), "d"(c[k]));
        }
    }
    double r = 0.0;
#pragma unroll
    for (int k = 0; k < 16; ++k) r += a[k];
    if (r == 123.456 || iv == 0x7fffffffu) sink[threadIdx.x] = r;
}

template <int MODE, int THREADS>
static void run(const char *name, int sms, double *sink) {
    const int iters = 40000;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    probe<MODE, THREADS><<<sms, THREADS>>>(iters / 10, sink, 0.0);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        probe<MODE, THREADS><<<sms, THREADS>>>(iters, sink, 0.0);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    const int warps = THREADS / 128;
    const double cyc = best * 1e-3 * 1.965e9 / iters / warps;   // cycles per 64 instructions per SMSP-warp slot
    printf("%-52s %d warps/SMSP %8.3f ms  %6.2f cycles per FP64 warp instruction per SMSP (2.00 = pipe peak)  pipe %.1f %%\n", name, warps, best,
           cyc / 64.0, 200.0 * 64.0 / cyc);
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    double *sink; CK(cudaMalloc(&sink, 8192));
#define BOTH(M, name) run<M, 256>(name, sms, sink); run<M, 384>(name, sms, sink); run<M, 512>(name, sms, sink);
    BOTH(0, "DFMA a=fma(a,x,y): 1 distinct source");
    BOTH(2, "DFMA a=fma(b_i,x,a): 2 distinct sources");
    BOTH(1, "DFMA a=fma(b_i,c_j,a): 3 distinct sources");
    BOTH(5, "DFMA pairs sharing one factor (cmul pattern)");
    BOTH(3, "DMUL a=a*b_i: 2 distinct sources");
    BOTH(4, "DADD a=a+b_i: 2 distinct sources");
    run<0, 768>("DFMA 1 distinct source", sms, sink); run<0, 1024>("DFMA 1 distinct source", sms, sink);
    run<1, 768>("DFMA 3 distinct sources", sms, sink);
    run<7, 256>("DFMA a=fma(a,x,y) with an IADD between every two", sms, sink); run<7, 384>("DFMA a=fma(a,x,y) with an IADD between every two", sms, sink);
    run<8, 256>("DFMA a=fma(b_i,c_j,a), 8 accumulators (RAW distance 8)", sms, sink);
    BOTH(9, "DADD + one LDS.128 (broadcast) per 4 DADD");
    BOTH(10, "DADD + one LDS.128 (per-lane address) per 4 DADD");
    BOTH(11, "DADD + one LDS.64 (broadcast) per 2 DADD");
    return 0;
}
