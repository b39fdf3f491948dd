// fp32_peak.cu — FP32 FMA roofline microbenchmark (the denominator of roofline.frac).
// MEASURED_PEAKS.json has no FP32 number, so the library measures its own:
// independent FMA chains on every SM, enough warps to fill all four schedulers.
//   variant 0: scalar FFMA, 16 independent chains per thread
//   variant 1: packed FFMA2 (fma.rn.f32x2), 8 independent 2-wide chains per thread
//   variant 2: an early instruction-mix probe (9 FFMA + min3/max per pair, no memory).  NOT a peak: ptxas
//              hoists part of the loop-invariant FFMAs, so the nominal flop count over-states what ran;
//              bench.py ignores it
//   variant 3: scalar FFMA, every chain with its own multiplier and addend registers
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/tracer_cuda.h"

namespace {

constexpr int CHAINS = 16;
constexpr int INNER = 256;

__global__ void __launch_bounds__(256) ffma_kernel(float *out, int outer, float a, float b) {
    float x[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) x[i] = (float)(threadIdx.x + i);
    for (int o = 0; o < outer; ++o) {
#pragma unroll
        for (int k = 0; k < INNER; ++k) {
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) x[i] = fmaf(x[i], a, b);
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += x[i];
    if (s == 123.456f) out[0] = s;
}

__global__ void __launch_bounds__(256) ffma2_kernel(float *out, int outer, float a, float b) {
    float2 x[CHAINS / 2];
    const float2 aa = make_float2(a, a * 1.0001f), bb = make_float2(b, b * 0.999f);
#pragma unroll
    for (int i = 0; i < CHAINS / 2; ++i) x[i] = make_float2((float)(threadIdx.x + i), (float)(threadIdx.x - i));
    for (int o = 0; o < outer; ++o) {
#pragma unroll
        for (int k = 0; k < INNER; ++k) {
#pragma unroll
            for (int i = 0; i < CHAINS / 2; ++i) x[i] = __ffma2_rn(x[i], aa, bb);
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS / 2; ++i) s += x[i].x + x[i].y;
    if (s == 123.456f) out[0] = s;
}

// scalar FFMA with distinct multiplier/addend registers per chain (no shared operands)
__global__ void __launch_bounds__(256) ffma_distinct_kernel(float *out, int outer, float a, float b) {
    float x[CHAINS], m[CHAINS], c[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) x[i] = (float)(threadIdx.x + i), m[i] = a + 1e-8f * i, c[i] = b * (i + 1);
    for (int o = 0; o < outer; ++o) {
#pragma unroll
        for (int k = 0; k < INNER; ++k) {
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) x[i] = fmaf(x[i], m[i], c[i]);
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += x[i];
    if (s == 123.456f) out[0] = s;
}

// the sweep's per-pair instruction mix, registers only: 8 rays x (9 FFMA + 2 FMNMX + 1 FMNMX)
__global__ void __launch_bounds__(256) mix_kernel(float *out, int outer, float a, float b) {
    float ex[8], ey[8], ez[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) ex[r] = a * (threadIdx.x + r), ey[r] = b * (threadIdx.x - r), ez[r] = a * b * r;
    float acc = -1.f;
    float4 rb = make_float4(a, b, a, b), rc = make_float4(b, a, b, a), rd = make_float4(a, a, b, b);
    for (int o = 0; o < outer; ++o) {
#pragma unroll 8
        for (int k = 0; k < INNER; ++k) {
            float M = -1.f;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float x = fmaf(ex[r], rb.x, fmaf(ey[r], rb.y, fmaf(ez[r], rb.z, rb.w)));
                const float y = fmaf(ex[r], rc.x, fmaf(ey[r], rc.y, fmaf(ez[r], rc.z, rc.w)));
                const float z = fmaf(ex[r], rd.x, fmaf(ey[r], rd.y, fmaf(ez[r], rd.z, rd.w)));
                M = fmaxf(M, fminf(fminf(x, y), z));
            }
            acc = fmaxf(acc, M);
            rb.x += 1e-7f, rc.y += 1e-7f, rd.z += 1e-7f; // keep the compiler from hoisting
        }
    }
    if (acc == 123.456f) out[0] = acc;
}

}  // namespace

extern "C" int tracer_cuda_fp32_peak(int32_t variant, int32_t iters, double *tflops_out, double *ms_out) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return TRACER_ERR_NO_DEVICE;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    float *out = nullptr;
    if (cudaMalloc(&out, 64) != cudaSuccess) return TRACER_ERR_NOMEM;
    const int grid = sms * 8, block = 256, outer = 64;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    auto launch = [&]() {
        if (variant == 0)
            ffma_kernel<<<grid, block>>>(out, outer, 1.0000001f, 1e-9f);
        else if (variant == 1)
            ffma2_kernel<<<grid, block>>>(out, outer, 1.0000001f, 1e-9f);
        else if (variant == 2)
            mix_kernel<<<grid, block>>>(out, outer, 1.0000001f, 1e-9f);
        else
            ffma_distinct_kernel<<<grid, block>>>(out, outer, 1.0000001f, 1e-9f);
    };
    for (int i = 0; i < 3; ++i) launch();
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) launch();
    cudaEventRecord(e1);
    cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0), cudaEventDestroy(e1);
    cudaFree(out);
    if (e != cudaSuccess) return TRACER_ERR_CUDA;
    double fma_per_thread = (double)outer * INNER * (variant == 2 ? 72.0 : (double)CHAINS);
    const double flops = 2.0 * fma_per_thread * (double)grid * block * iters;
    if (tflops_out) *tflops_out = flops / (ms * 1e-3) / 1e12;
    if (ms_out) *ms_out = ms / iters;
    return TRACER_OK;
}
