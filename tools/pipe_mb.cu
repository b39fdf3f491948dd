// pipe_mb.cu — instruction-throughput probes on B200: what does one FMNMX3 / LOP3 / SHF cost next to FFMA?
#include <cstdio>
#include <cuda_runtime.h>

#define REPS 4096
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float *out, float a, float b, int n) {
    float x[12];
    unsigned u[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) x[i] = a * (threadIdx.x + i), u[i] = threadIdx.x * 2654435761u + i;
    float m0 = a, m1 = b, m2 = a + b, m3 = a - b;
    unsigned l0 = threadIdx.x, l1 = threadIdx.x + 1;
    for (int it = 0; it < n; ++it) {
#pragma unroll 16
        for (int k = 0; k < 16; ++k) {
            // 12 independent FFMA
            if (MODE != 3 && MODE != 4) {
#pragma unroll
                for (int i = 0; i < 12; ++i) x[i] = fmaf(x[i], a, b);
            }
            if (MODE == 1 || MODE == 3) { // + 2 FMNMX3 per 12 FFMA (the sweep's ratio)
                m0 = fminf(fminf(m0, x[0]), x[1]);
                m1 = fmaxf(fmaxf(m1, x[2]), x[3]);
                if (MODE == 3) {
                    m2 = fminf(fminf(m2, x[4]), x[5]);
                    m3 = fmaxf(fmaxf(m3, x[6]), x[7]);
                    m0 = fminf(fminf(m0, x[8]), x[9]);
                    m1 = fmaxf(fmaxf(m1, x[10]), x[11]);
                    m2 = fminf(fminf(m2, x[0]), x[2]);
                    m3 = fmaxf(fmaxf(m3, x[1]), x[3]);
                }
            }
            if (MODE == 2 || MODE == 4) { // + 2 LOP3 per 12 FFMA
                l0 = l0 | __float_as_uint(x[0]) | __float_as_uint(x[1]);
                l1 = l1 & (__float_as_uint(x[2]) | __float_as_uint(x[3]));
                if (MODE == 4) {
                    l0 = l0 ^ (u[4] | u[5]);
                    l1 = l1 & (u[6] | u[7]);
                    l0 = l0 | u[8] | u[9];
                    l1 = l1 & (u[10] | u[11]);
                    l0 = l0 ^ (u[0] & u[2]);
                    l1 = l1 | (u[1] & u[3]);
                }
            }
            if (MODE == 5) { // + 2 FMNMX (2-input)
                m0 = fminf(m0, x[0]);
                m1 = fmaxf(m1, x[2]);
            }
        }
    }
    float s = m0 + m1 + m2 + m3 + __uint_as_float(l0 ^ l1);
#pragma unroll
    for (int i = 0; i < 12; ++i) s += x[i];
    if (s == 123.456f) out[0] = s;
}

template <int MODE>
void run(const char *name, double ops_per_inner, int sms) {
    float *out;
    cudaMalloc(&out, 64);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    const int n = 256;
    k<MODE><<<sms, 512>>>(out, 1.0000001f, 1e-9f, n);
    cudaEventRecord(e0);
    k<MODE><<<sms, 512>>>(out, 1.0000001f, 1e-9f, n);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int khz;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double warp_inst = (double)n * 16 * ops_per_inner * 16 /*warps*/ / 4 /*smsp*/;
    const double cycles = ms * 1e-3 * khz * 1e3;
    printf("%-40s %8.3f ms  %6.3f warp-inst/cycle/SMSP (at %d MHz nominal)\n", name, ms, warp_inst / cycles, khz / 1000);
}

int main() {
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    run<0>("12 FFMA", 12, sms);
    run<1>("12 FFMA + 2 FMNMX3", 14, sms);
    run<5>("12 FFMA + 2 FMNMX", 14, sms);
    run<2>("12 FFMA + 2 LOP3", 14, sms);
    run<3>("8 FMNMX3 only", 8, sms);
    run<4>("8 LOP3 only", 8, sms);
    return 0;
}
