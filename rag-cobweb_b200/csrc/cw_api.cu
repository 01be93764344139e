// cw_api.cu -- error reporting, version, and the FFMA microbenchmark of the C ABI
// (include/cobweb_b200.h).
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/cobweb_b200.h"

static thread_local char g_err[512] = "";

void cw_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cw_check_cuda(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return 0;
    cw_set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
    return CW_E_CUDA;
}

extern "C" int cw_version(void) { return CW_VERSION; }
extern "C" const char *cw_last_error(void) { return g_err; }

// Independent FFMA chains: 16 accumulators per thread, 4 FFMAs each per inner step.  Used by
// bench.py to measure the FP32-FMA issue peak that bounds the dense scoring kernel.
__global__ void ffma_peak_kernel(int iters, float *sink) {
    float a[16];
    float x = 1.0f + 1e-7f * threadIdx.x, y = 1e-9f * blockIdx.x;
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = (float)i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], x, y);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
    if (s == 12345.678f) sink[0] = s;
}

// Same, with the packed fp32x2 FMA (FFMA2) the scoring kernel uses: 16 independent 64-bit
// accumulators per thread; mode 0 = all operands vary per chain, mode 1 = scalar-broadcast first
// operand like the kernel's x operand.
__global__ void ffma2_peak_kernel(int iters, float *sink) {
    unsigned long long a[16];
    float xf = 1.0f + 1e-7f * threadIdx.x, yf = 1e-9f * blockIdx.x;
    unsigned long long x, y;
    asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(xf), "f"(xf));
    asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(yf), "f"(yf));
#pragma unroll
    for (int i = 0; i < 16; i++) asm("mov.b64 %0, {%1, %2};" : "=l"(a[i]) : "f"((float)i), "f"((float)i + 0.5f));
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(x), "l"(y));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a[i]));
        s += lo + hi;
    }
    if (s == 12345.678f) sink[0] = s;
}

extern "C" int cw_ffma2_peak(int blocks, int threads, int iters, float *sink, void *stream) {
    ffma2_peak_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, sink);
    return cw_check_cuda(cudaGetLastError(), "cw_ffma2_peak");
}

extern "C" int cw_ffma_peak(int blocks, int threads, int iters, float *sink, void *stream) {
    ffma_peak_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, sink);
    return cw_check_cuda(cudaGetLastError(), "cw_ffma_peak");
}
