// api.cu -- C ABI of libbossperm.so (see include/bossperm.h): context, staging, entry points.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "bp_common.cuh"

// kernels implemented in the other translation units
int bp_k1_launch(bp_context *h, const double *dA, int N, uint64_t lo, uint64_t hi, double *d_out_dd, double *d_exchange_out);
int bp_effective_matrix_launch(bp_context *h, const double *dU, int m, const int32_t *d_s, const int32_t *d_t,
                               int N, double *dA);
int bp_fp64_peak_launch(bp_context *h, int iters, double *d_sink);
int bp_k2_launch(bp_context *h, const double *dU, int m, const unsigned char *dS, const unsigned char *dT,
                 long long B, double *d_out, const unsigned char *hS);
int bp_k2_small_host(bp_context *h, const double *U, int m, const unsigned char *S, const unsigned char *T, long long B, double *out);
#define BP_K2_SMALL_MAX 256   // (= K2_HOST_PREP_MAX of guan_kernel.cu)

static char g_global_err[512] = "no error";

int bp_fail(bp_context *h, int code, const char *fmt, ...) {
    char *dst = h ? h->err : g_global_err;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
    return code;
}

int bp_reserve(bp_context *h, int slot, size_t bytes) {
    if (bytes <= h->d_cap[slot]) return BP_OK;
    size_t cap = bytes < 4096 ? 4096 : bytes + bytes / 2;
    if (h->d_buf[slot]) {
        BP_CUDA(h, cudaStreamSynchronize(h->stream));
        BP_CUDA(h, cudaFree(h->d_buf[slot]));
        h->d_buf[slot] = nullptr;
        h->d_cap[slot] = 0;
    }
    cudaError_t e = cudaMalloc(&h->d_buf[slot], cap);
    if (e != cudaSuccess) return bp_fail(h, BP_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", cap, cudaGetErrorString(e));
    h->d_cap[slot] = cap;
    return BP_OK;
}

int bp_reserve_pinned(bp_context *h, size_t bytes) {
    if (bytes <= h->h_cap) return BP_OK;
    size_t cap = bytes < 4096 ? 4096 : bytes + bytes / 2;
    if (h->h_pin) {
        BP_CUDA(h, cudaStreamSynchronize(h->stream));
        BP_CUDA(h, cudaFreeHost(h->h_pin));
        h->h_pin = nullptr;
        h->h_cap = 0;
    }
    cudaError_t e = cudaMallocHost(&h->h_pin, cap);
    if (e != cudaSuccess) return bp_fail(h, BP_ERR_NOMEM, "cudaMallocHost(%zu) failed: %s", cap, cudaGetErrorString(e));
    h->h_cap = cap;
    return BP_OK;
}

extern "C" {

int bp_abi_version(void) { return BP_ABI_VERSION; }

const char *bp_last_error(bp_handle h) { return h ? h->err : g_global_err; }

static int bp_create_impl(int device, void *stream, bool own, bp_handle *out) {
    if (!out) return bp_fail(nullptr, BP_ERR_INVALID, "bp_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return bp_fail(nullptr, BP_ERR_CUDA, "bp_create: no CUDA device (%s); libbossperm has no CPU fallback",
                       e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= count) return bp_fail(nullptr, BP_ERR_INVALID, "bp_create: device %d of %d", device, count);
    bp_context *h = new (std::nothrow) bp_context();
    if (!h) return bp_fail(nullptr, BP_ERR_NOMEM, "bp_create: out of host memory");
    h->device = device;
    cudaDeviceProp prop;
    bp_device_guard guard(device);
    if ((e = guard.err) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        delete h;
        return bp_fail(nullptr, BP_ERR_CUDA, "bp_create: %s", cudaGetErrorString(e));
    }
    h->sm_count = prop.multiProcessorCount;
    h->cc_major = prop.major;
    h->cc_minor = prop.minor;
    cudaDeviceGetAttribute(&h->clock_khz, cudaDevAttrClockRate, device);
    if (prop.major < 10) {
        int major = prop.major;
        delete h;
        return bp_fail(nullptr, BP_ERR_CUDA, "bp_create: device compute capability %d.x; this library is built for sm_100a only", major);
    }
    if (own) {
        if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) {
            delete h;
            return bp_fail(nullptr, BP_ERR_CUDA, "bp_create: cudaStreamCreate: %s", cudaGetErrorString(e));
        }
        h->own_stream = true;
    } else {
        h->stream = (cudaStream_t)stream;
    }
    if ((e = cudaEventCreate(&h->ev0)) != cudaSuccess || (e = cudaEventCreate(&h->ev1)) != cudaSuccess) {
        if (h->ev0) cudaEventDestroy(h->ev0);
        if (h->own_stream) cudaStreamDestroy(h->stream);
        delete h;
        return bp_fail(nullptr, BP_ERR_CUDA, "bp_create: cudaEventCreate: %s", cudaGetErrorString(e));
    }
    snprintf(h->err, sizeof(h->err), "no error");
    *out = h;
    return BP_OK;
}

int bp_create(int device, bp_handle *out) { return bp_create_impl(device, nullptr, true, out); }
int bp_create_on_stream(int device, void *cuda_stream, bp_handle *out) { return bp_create_impl(device, cuda_stream, false, out); }

static void exchange_release(bp_context *h);

int bp_destroy(bp_handle h) {
    if (!h) return BP_OK;
    bp_device_guard guard(h->device);
    cudaStreamSynchronize(h->stream);
    for (int i = 0; i < 8; ++i)
        if (h->d_buf[i]) cudaFree(h->d_buf[i]);
    if (h->h_pin) cudaFreeHost(h->h_pin);
    if (h->d_counter) cudaFree(h->d_counter);
    exchange_release(h);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
    return BP_OK;
}

int bp_synchronize(bp_handle h) {
    if (!h) return bp_fail(nullptr, BP_ERR_INVALID, "NULL handle");
    BP_ON_DEVICE(h);
    BP_CUDA(h, cudaStreamSynchronize(h->stream));
    return BP_OK;
}

int bp_device_info(bp_handle h, int *sm_count, int *cc_major, int *cc_minor, int *clock_khz) {
    if (!h) return bp_fail(nullptr, BP_ERR_INVALID, "NULL handle");
    if (sm_count) *sm_count = h->sm_count;
    if (cc_major) *cc_major = h->cc_major;
    if (cc_minor) *cc_minor = h->cc_minor;
    if (clock_khz) *clock_khz = h->clock_khz;
    return BP_OK;
}

int64_t bp_launch_count(bp_handle h) { return h ? h->launches : 0; }

int bp_timer_start(bp_handle h) {
    if (!h) return bp_fail(nullptr, BP_ERR_INVALID, "NULL handle");
    BP_ON_DEVICE(h);
    BP_CUDA(h, cudaEventRecord(h->ev0, h->stream));
    return BP_OK;
}

int bp_timer_stop(bp_handle h, float *elapsed_ms) {
    if (!h || !elapsed_ms) return bp_fail(h, BP_ERR_INVALID, "bp_timer_stop: NULL argument");
    BP_ON_DEVICE(h);
    BP_CUDA(h, cudaEventRecord(h->ev1, h->stream));
    BP_CUDA(h, cudaEventSynchronize(h->ev1));
    BP_CUDA(h, cudaEventElapsedTime(elapsed_ms, h->ev0, h->ev1));
    return BP_OK;
}

int bp_fp64_peak(bp_handle h, double target_ms, double *tflops) {
    if (!h || !tflops) return bp_fail(h, BP_ERR_INVALID, "bp_fp64_peak: NULL argument");
    BP_ON_DEVICE(h);
    int rc = bp_reserve(h, BP_SLOT_MISC, sizeof(double) * 1024);
    if (rc) return rc;
    double *sink = (double *)h->d_buf[BP_SLOT_MISC];
    // fixed-size launches (~target_ms each on a B200), one warm-up, best of three
    int iters = (int)(50000.0 * target_ms / 52.0);
    if (iters < 1000) iters = 1000;
    float ms = 0.f;
    double best = 0.0;
    rc = bp_fp64_peak_launch(h, iters / 4, sink);
    if (rc) return rc;
    for (int round = 0; round < 3; ++round) {
        BP_CUDA(h, cudaEventRecord(h->ev0, h->stream));
        rc = bp_fp64_peak_launch(h, iters, sink);
        if (rc) return rc;
        BP_CUDA(h, cudaEventRecord(h->ev1, h->stream));
        BP_CUDA(h, cudaEventSynchronize(h->ev1));
        BP_CUDA(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        // per launch: sm_count * 8 blocks * 256 threads * 64 DFMA per iteration
        const double flops = 2.0 * (double)h->sm_count * 8.0 * 256.0 * 64.0 * (double)iters;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    *tflops = best;
    return BP_OK;
}

// ---- K1 --------------------------------------------------------------------------------------
static int glynn_range_host(bp_handle h, const double *A, int N, uint64_t lo, uint64_t hi, double out_dd[4]) {
    BP_ON_DEVICE(h);
    const size_t bytes = sizeof(double) * 2 * (size_t)N * N;
    int rc = bp_reserve(h, BP_SLOT_MATRIX, bytes);
    if (rc) return rc;
    if ((rc = bp_reserve(h, BP_SLOT_OUT, sizeof(double) * 4))) return rc;
    if ((rc = bp_reserve_pinned(h, bytes + 64))) return rc;
    memcpy(h->h_pin, A, bytes);
    BP_CUDA(h, cudaMemcpyAsync(h->d_buf[BP_SLOT_MATRIX], h->h_pin, bytes, cudaMemcpyHostToDevice, h->stream));
    rc = bp_k1_launch(h, (const double *)h->d_buf[BP_SLOT_MATRIX], N, lo, hi, (double *)h->d_buf[BP_SLOT_OUT], nullptr);
    if (rc) return rc;
    double *res = (double *)((char *)h->h_pin + ((bytes + 31) / 32) * 32);
    BP_CUDA(h, cudaMemcpyAsync(res, h->d_buf[BP_SLOT_OUT], sizeof(double) * 4, cudaMemcpyDeviceToHost, h->stream));
    BP_CUDA(h, cudaStreamSynchronize(h->stream));
    for (int q = 0; q < 4; ++q) out_dd[q] = res[q];
    return BP_OK;
}

int bp_glynn_matrix_range(bp_handle h, const double *A, int N, uint64_t step_lo, uint64_t step_hi, double out_dd[4]) {
    if (!h || !A || !out_dd) return bp_fail(h, BP_ERR_INVALID, "bp_glynn_matrix_range: NULL argument");
    if (N < 1 || N > BP_MAX_N) return bp_fail(h, BP_ERR_UNSUPPORTED, "bp_glynn_matrix_range: N=%d outside [1, %d]", N, BP_MAX_N);
    return glynn_range_host(h, A, N, step_lo, step_hi, out_dd);
}

int bp_glynn_matrix_range_dev(bp_handle h, const double *dA, int N, uint64_t step_lo, uint64_t step_hi, double *d_out_dd) {
    if (!h || !dA || !d_out_dd) return bp_fail(h, BP_ERR_INVALID, "bp_glynn_matrix_range_dev: NULL argument");
    BP_ON_DEVICE(h);
    return bp_k1_launch(h, dA, N, step_lo, step_hi, d_out_dd, nullptr);
}

int bp_glynn_set_resident(bp_handle h, const double *dA) {
    if (!h) return bp_fail(nullptr, BP_ERR_INVALID, "NULL handle");
    h->resident_A = dA;
    h->resident_gen++;      // same pointer again: new contents -> the constant-bank image is rebuilt once
    return BP_OK;
}

// ---- K1 partial exchange over peer memory ----------------------------------------------------
static void exchange_release(bp_context *h) {
    for (int r = 0; r < BP_MAX_PEERS; ++r) {
        if (h->xchg_mapped[r] && h->xchg_peer[r]) cudaIpcCloseMemHandle(h->xchg_peer[r]);
        h->xchg_peer[r] = nullptr;
        h->xchg_mapped[r] = false;
    }
    if (h->xchg_local) cudaFree(h->xchg_local);
    h->xchg_local = nullptr;
    h->xchg_world = 0; h->xchg_rank = 0; h->xchg_seq = 0;
}

int bp_exchange_create(bp_handle h, int world, int rank, unsigned char ipc_handle_out[64]) {
    if (!h || !ipc_handle_out) return bp_fail(h, BP_ERR_INVALID, "bp_exchange_create: NULL argument");
    if (world < 1 || world > BP_MAX_PEERS || rank < 0 || rank >= world)
        return bp_fail(h, BP_ERR_INVALID, "bp_exchange_create: rank %d of %d (at most %d ranks)", rank, world, BP_MAX_PEERS);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    BP_ON_DEVICE(h);
    BP_CUDA(h, cudaStreamSynchronize(h->stream));
    exchange_release(h);
    const size_t bytes = sizeof(double) * 8 * 2 * (size_t)world;      // two halves of `world` slots of 8 doubles
    cudaError_t e = cudaMalloc((void **)&h->xchg_local, bytes);
    if (e != cudaSuccess) return bp_fail(h, BP_ERR_NOMEM, "bp_exchange_create: cudaMalloc: %s", cudaGetErrorString(e));
    BP_CUDA(h, cudaMemset(h->xchg_local, 0, bytes));                   // call numbers start at 1
    cudaIpcMemHandle_t ipc;
    if ((e = cudaIpcGetMemHandle(&ipc, h->xchg_local)) != cudaSuccess) {
        exchange_release(h);
        return bp_fail(h, BP_ERR_CUDA, "bp_exchange_create: cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    }
    memcpy(ipc_handle_out, &ipc, 64);
    h->xchg_world = world; h->xchg_rank = rank; h->xchg_seq = 0;
    h->xchg_peer[rank] = h->xchg_local;
    return BP_OK;
}

int bp_exchange_connect(bp_handle h, const unsigned char *ipc_handles) {
    if (!h || !ipc_handles) return bp_fail(h, BP_ERR_INVALID, "bp_exchange_connect: NULL argument");
    if (h->xchg_world < 1) return bp_fail(h, BP_ERR_INVALID, "bp_exchange_connect: bp_exchange_create comes first");
    BP_ON_DEVICE(h);
    for (int r = 0; r < h->xchg_world; ++r) {
        if (r == h->xchg_rank || h->xchg_peer[r]) continue;
        cudaIpcMemHandle_t ipc;
        memcpy(&ipc, ipc_handles + 64 * (size_t)r, 64);
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, ipc, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            return bp_fail(h, BP_ERR_CUDA, "bp_exchange_connect: cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
        }
        h->xchg_peer[r] = (double *)p;
        h->xchg_mapped[r] = true;
    }
    return BP_OK;
}

int bp_exchange_destroy(bp_handle h) {
    if (!h) return bp_fail(nullptr, BP_ERR_INVALID, "NULL handle");
    BP_ON_DEVICE(h);
    BP_CUDA(h, cudaStreamSynchronize(h->stream));
    exchange_release(h);
    return BP_OK;
}

int bp_glynn_matrix_range_exchange(bp_handle h, const double *dA, int N, uint64_t step_lo, uint64_t step_hi, double *d_out_all) {
    if (!h || !dA || !d_out_all) return bp_fail(h, BP_ERR_INVALID, "bp_glynn_matrix_range_exchange: NULL argument");
    BP_ON_DEVICE(h);
    return bp_k1_launch(h, dA, N, step_lo, step_hi, nullptr, d_out_all);
}

int bp_glynn_matrix_range_exchange_host(bp_handle h, const double *A, int N, uint64_t step_lo, uint64_t step_hi, double *out_all) {
    if (!h || !A || !out_all) return bp_fail(h, BP_ERR_INVALID, "bp_glynn_matrix_range_exchange_host: NULL argument");
    if (N < 1 || N > BP_MAX_N) return bp_fail(h, BP_ERR_UNSUPPORTED, "bp_glynn_matrix_range_exchange_host: N=%d outside [1, %d]", N, BP_MAX_N);
    if (h->xchg_world < 1) return bp_fail(h, BP_ERR_INVALID, "bp_glynn_matrix_range_exchange_host: call bp_exchange_create / bp_exchange_connect first");
    BP_ON_DEVICE(h);
    const size_t bytes = sizeof(double) * 2 * (size_t)N * N, out_bytes = sizeof(double) * 4 * (size_t)h->xchg_world;
    int rc = bp_reserve(h, BP_SLOT_MATRIX, bytes);
    if (rc) return rc;
    if ((rc = bp_reserve(h, BP_SLOT_OUT, out_bytes))) return rc;
    if ((rc = bp_reserve_pinned(h, bytes + 64 + out_bytes))) return rc;
    memcpy(h->h_pin, A, bytes);
    BP_CUDA(h, cudaMemcpyAsync(h->d_buf[BP_SLOT_MATRIX], h->h_pin, bytes, cudaMemcpyHostToDevice, h->stream));
    rc = bp_k1_launch(h, (const double *)h->d_buf[BP_SLOT_MATRIX], N, step_lo, step_hi, nullptr, (double *)h->d_buf[BP_SLOT_OUT]);
    if (rc) return rc;
    double *res = (double *)((char *)h->h_pin + ((bytes + 31) / 32) * 32);
    BP_CUDA(h, cudaMemcpyAsync(res, h->d_buf[BP_SLOT_OUT], out_bytes, cudaMemcpyDeviceToHost, h->stream));
    BP_CUDA(h, cudaStreamSynchronize(h->stream));
    memcpy(out_all, res, out_bytes);
    return BP_OK;
}

int bp_glynn_matrix(bp_handle h, const double *A, int N, double out[2]) {
    if (!h || !out) return bp_fail(h, BP_ERR_INVALID, "bp_glynn_matrix: NULL argument");
    if (N == 0) { out[0] = 1.0; out[1] = 0.0; return BP_OK; }   // glynn_gray_permanent_calculator.py:52-53
    if (!A) return bp_fail(h, BP_ERR_INVALID, "bp_glynn_matrix: A is NULL");
    if (N < 0 || N > BP_MAX_N) return bp_fail(h, BP_ERR_UNSUPPORTED, "bp_glynn_matrix: N=%d outside [0, %d]", N, BP_MAX_N);
    double p[4];
    int rc = glynn_range_host(h, A, N, 0, 1ull << (N - 1), p);
    if (rc) return rc;
    const double scale = ldexp(1.0, -(N - 1));
    out[0] = (p[0] + p[1]) * scale;
    out[1] = (p[2] + p[3]) * scale;
    return BP_OK;
}

int bp_glynn_single(bp_handle h, const double *U, int m, const int32_t *s, const int32_t *t, double out[2]) {
    if (!h || !U || !s || !t || !out) return bp_fail(h, BP_ERR_INVALID, "bp_glynn_single: NULL argument");
    if (m < 1 || m > BP_MAX_MODES) return bp_fail(h, BP_ERR_UNSUPPORTED, "bp_glynn_single: m=%d outside [1, %d]", m, BP_MAX_MODES);
    long ns = 0, nt = 0;
    for (int i = 0; i < m; ++i) {
        if (s[i] < 0 || t[i] < 0) return bp_fail(h, BP_ERR_INVALID, "bp_glynn_single: negative occupation");
        ns += s[i];
        nt += t[i];
    }
    if (ns == 0 || nt == 0) { out[0] = 1.0; out[1] = 0.0; return BP_OK; }   // :52-53 (empty effective matrix)
    if (ns != nt) return bp_fail(h, BP_ERR_SHAPE, "bp_glynn_single: %ld input vs %ld output particles", ns, nt);
    if (ns > BP_MAX_N) return bp_fail(h, BP_ERR_UNSUPPORTED, "bp_glynn_single: n=%ld > %d", ns, BP_MAX_N);
    const int N = (int)ns;
    BP_ON_DEVICE(h);
    const size_t ub = sizeof(double) * 2 * (size_t)m * m, sb = sizeof(int32_t) * (size_t)m;
    int rc;
    if ((rc = bp_reserve(h, BP_SLOT_AUX, ub))) return rc;
    if ((rc = bp_reserve(h, BP_SLOT_STATE, 2 * sb))) return rc;
    if ((rc = bp_reserve(h, BP_SLOT_MATRIX, sizeof(double) * 2 * (size_t)N * N))) return rc;
    if ((rc = bp_reserve(h, BP_SLOT_OUT, sizeof(double) * 4))) return rc;
    if ((rc = bp_reserve_pinned(h, ub + 2 * sb + 64))) return rc;
    char *pin = (char *)h->h_pin;
    memcpy(pin, U, ub);
    memcpy(pin + ub, s, sb);
    memcpy(pin + ub + sb, t, sb);
    BP_CUDA(h, cudaMemcpyAsync(h->d_buf[BP_SLOT_AUX], pin, ub, cudaMemcpyHostToDevice, h->stream));
    BP_CUDA(h, cudaMemcpyAsync(h->d_buf[BP_SLOT_STATE], pin + ub, 2 * sb, cudaMemcpyHostToDevice, h->stream));
    const int32_t *d_s = (const int32_t *)h->d_buf[BP_SLOT_STATE];
    rc = bp_effective_matrix_launch(h, (const double *)h->d_buf[BP_SLOT_AUX], m, d_s, d_s + m, N, (double *)h->d_buf[BP_SLOT_MATRIX]);
    if (rc) return rc;
    rc = bp_k1_launch(h, (const double *)h->d_buf[BP_SLOT_MATRIX], N, 0, 1ull << (N - 1), (double *)h->d_buf[BP_SLOT_OUT], nullptr);
    if (rc) return rc;
    double *res = (double *)(pin + ((ub + 2 * sb + 31) / 32) * 32);
    BP_CUDA(h, cudaMemcpyAsync(res, h->d_buf[BP_SLOT_OUT], sizeof(double) * 4, cudaMemcpyDeviceToHost, h->stream));
    BP_CUDA(h, cudaStreamSynchronize(h->stream));
    const double scale = ldexp(1.0, -(N - 1));
    out[0] = (res[0] + res[1]) * scale;
    out[1] = (res[2] + res[3]) * scale;
    return BP_OK;
}


// ---- K2 --------------------------------------------------------------------------------------
int bp_perm_batched_dev(bp_handle h, const double *dU, int m, const uint8_t *dS, const uint8_t *dT, int64_t B,
                        int formula, double *d_out) {
    if (!h || !dU || !dS || !dT || !d_out) return bp_fail(h, BP_ERR_INVALID, "bp_perm_batched_dev: NULL argument");
    if (m < 1 || m > BP_MAX_MODES) return bp_fail(h, BP_ERR_UNSUPPORTED, "bp_perm_batched_dev: m=%d outside [1, %d]", m, BP_MAX_MODES);
    if (formula < BP_FORMULA_RYSER || formula > BP_FORMULA_GLYNN) return bp_fail(h, BP_ERR_INVALID, "bp_perm_batched_dev: formula %d", formula);
    if (B < 0) return bp_fail(h, BP_ERR_INVALID, "bp_perm_batched_dev: B=%lld", (long long)B);
    BP_ON_DEVICE(h);
    return bp_k2_launch(h, dU, m, dS, dT, (long long)B, d_out, nullptr);
}

int bp_perm_batched(bp_handle h, const double *U, int m, const uint8_t *S, const uint8_t *T, int64_t B, int formula,
                    double *out) {
    if (!h || !U || !S || !T || !out) return bp_fail(h, BP_ERR_INVALID, "bp_perm_batched: NULL argument");
    if (m < 1 || m > BP_MAX_MODES) return bp_fail(h, BP_ERR_UNSUPPORTED, "bp_perm_batched: m=%d outside [1, %d]", m, BP_MAX_MODES);
    if (formula < BP_FORMULA_RYSER || formula > BP_FORMULA_GLYNN) return bp_fail(h, BP_ERR_INVALID, "bp_perm_batched: formula %d", formula);
    if (B < 0) return bp_fail(h, BP_ERR_INVALID, "bp_perm_batched: B=%lld", (long long)B);
    if (B == 0) return BP_OK;
    for (int64_t b = 0; b < B; ++b) {   // the reference raises before computing (bs_permanent_calculator_base.py:179-180)
        long ns = 0, nt = 0;
        for (int v = 0; v < m; ++v) { ns += S[b * m + v]; nt += T[b * m + v]; }
        if (ns != nt) return bp_fail(h, BP_ERR_SHAPE, "bp_perm_batched: item %lld has %ld input vs %ld output particles", (long long)b, ns, nt);
        if (ns > BP_MAX_N) return bp_fail(h, BP_ERR_UNSUPPORTED, "bp_perm_batched: item %lld has n=%ld > %d", (long long)b, ns, BP_MAX_N);
    }
    BP_ON_DEVICE(h);
    if (B <= BP_K2_SMALL_MAX) return bp_k2_small_host(h, U, m, S, T, (long long)B, out);
    const size_t ub = sizeof(double) * 2 * (size_t)m * m, sb = (size_t)B * m, ob = sizeof(double) * 2 * (size_t)B;
    int rc;
    if ((rc = bp_reserve(h, BP_SLOT_AUX, ub))) return rc;
    if ((rc = bp_reserve(h, BP_SLOT_STATE, 2 * sb + 32))) return rc;
    if ((rc = bp_reserve(h, BP_SLOT_OUT, ob))) return rc;
    // U, S, T are read by the DMA engine straight from the caller's (pageable) buffers
    BP_CUDA(h, cudaMemcpyAsync(h->d_buf[BP_SLOT_AUX], U, ub, cudaMemcpyHostToDevice, h->stream));
    unsigned char *dS = (unsigned char *)h->d_buf[BP_SLOT_STATE], *dT = dS + ((sb + 15) / 16) * 16;
    BP_CUDA(h, cudaMemcpyAsync(dS, S, sb, cudaMemcpyHostToDevice, h->stream));
    BP_CUDA(h, cudaMemcpyAsync(dT, T, sb, cudaMemcpyHostToDevice, h->stream));
    rc = bp_k2_launch(h, (const double *)h->d_buf[BP_SLOT_AUX], m, dS, dT, (long long)B, (double *)h->d_buf[BP_SLOT_OUT], S);
    if (rc) return rc;
    BP_CUDA(h, cudaMemcpyAsync(out, h->d_buf[BP_SLOT_OUT], ob, cudaMemcpyDeviceToHost, h->stream));
    BP_CUDA(h, cudaStreamSynchronize(h->stream));
    return BP_OK;
}

}  // extern "C"
