// sweep_mb4.cu — round-2 microbenchmark of the SPAN form of the filter loop.
//
// For the rays of one thread (same q, or q-sorted with a shared qbar/qdelta) the three affine edge rows
// p*A + (q*B + C) >= 0 are three half-lines in p: two lower bounds and one upper bound, or one and two.  Dividing
// each row by |A| once per (origin, triangle) in FP64 turns the per-pair work into a two-sided span test:
//     per thread and triangle:  a_i = q*Bl_i + Cl_i (i = 1,2)   b_i = q*Bu_i + Cu_i   ax = min(a1,a2)  ay = min(b1,b2)
//     per pair:                 x = sat(p*S + ax)   y = sat(ay - p*S)   acc += x*y        (2 FMA-pipe ops + 1/2 packed FFMA2)
// against 3 FFMA.SAT + FMUL2/2 + FFMA2/2 per pair of the three-row form (tools/sweep_mb3.cu).  32-byte rows.
// Variants: how the two saturating ops are written (FADD.SAT on a pre-scaled p, FFMA.SAT with an immediate scale,
// FFMA.SAT with the scale in a register), rays per thread, accumulator groups.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

constexpr int TILE = 256, BATCH = 16;
enum { MODE_SHAREDQ = 1, MODE_QBAR = 2 };
enum { F_ADD = 0, F_FMAI = 1, F_FMAR = 2 };

template <int R, int MODE, int FORM, int GROUP, int NACC2>
__device__ __forceinline__ unsigned eval_span(const float4 *__restrict__ tp, const float (&rp)[R], float q, float qdelta, float S) {
    unsigned cand = 0;
#pragma unroll
    for (int g = 0; g < BATCH / GROUP; ++g) {
        float2 acc[NACC2];
#pragma unroll
        for (int a = 0; a < NACC2; ++a) acc[a] = make_float2(0.f, 0.f);
#pragma unroll
        for (int kk = 0; kk < GROUP; ++kk) {
            const int k = g * GROUP + kk;
            const float4 lo = tp[2 * k], hi = tp[2 * k + 1];
            float a1 = fmaf(q, lo.x, lo.y), a2 = fmaf(q, lo.z, lo.w), b1 = fmaf(q, hi.x, hi.y), b2 = fmaf(q, hi.z, hi.w);
            if (MODE == MODE_QBAR) {
                a1 = fmaf(fabsf(lo.x), qdelta, a1), a2 = fmaf(fabsf(lo.z), qdelta, a2);
                b1 = fmaf(fabsf(hi.x), qdelta, b1), b2 = fmaf(fabsf(hi.z), qdelta, b2);
            }
            const float ax = fminf(a1, a2), ay = fminf(b1, b2);
#pragma unroll
            for (int r = 0; r < R; r += 2) {
                float2 x, y;
                if (FORM == F_ADD) {
                    x = make_float2(__saturatef(rp[r] + ax), __saturatef(rp[r + 1] + ax));
                    y = make_float2(__saturatef(ay - rp[r]), __saturatef(ay - rp[r + 1]));
                } else if (FORM == F_FMAI) {
                    x = make_float2(__saturatef(fmaf(rp[r], 65536.f, ax)), __saturatef(fmaf(rp[r + 1], 65536.f, ax)));
                    y = make_float2(__saturatef(fmaf(rp[r], -65536.f, ay)), __saturatef(fmaf(rp[r + 1], -65536.f, ay)));
                } else {
                    x = make_float2(__saturatef(fmaf(rp[r], S, ax)), __saturatef(fmaf(rp[r + 1], S, ax)));
                    y = make_float2(__saturatef(fmaf(rp[r], -S, ay)), __saturatef(fmaf(rp[r + 1], -S, ay)));
                }
                acc[(r / 2) % NACC2] = __ffma2_rn(x, y, acc[(r / 2) % NACC2]);
            }
        }
        float s = 0.f;
#pragma unroll
        for (int a = 0; a < NACC2; ++a) s += acc[a].x + acc[a].y;
        if (s >= 1.f) cand |= 1u << g;
    }
    return cand;
}

template <int R, int MODE, int FORM, int NTH, int MINBLK, int GROUP, int NACC2>
__global__ void __launch_bounds__(NTH, MINBLK) k_span(const float4 *tile_g, int reps, float *out, float seed, float S) {
    __shared__ __align__(16) float4 tile[TILE * 2];
    for (int i = threadIdx.x; i < TILE * 2; i += blockDim.x) tile[i] = tile_g[i];
    __syncthreads();
    float rp[R];
#pragma unroll
    for (int r = 0; r < R; ++r) rp[r] = seed * (threadIdx.x + 1) * (r + 1);
    const float q = seed * (threadIdx.x + 7) * 3, qdelta = 1e-6f * (1 + (threadIdx.x & 3));
    int hits = 0;
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll 1
        for (int b0 = 0; b0 < TILE; b0 += BATCH) {
            const unsigned c = eval_span<R, MODE, FORM, GROUP, NACC2>(tile + 2 * b0, rp, q, qdelta, S);
            if (c) {
                hits += 1;
                asm volatile("" ::: "memory");
            }
        }
    }
    if (hits == 123456789) out[0] = hits;
}

static int g_sms = 0;
template <typename K, typename F>
void run(const char *tag, K kern, int nth, int minblk, F launch, int R, double flops_per_pair) {
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nth, 0);
    cudaFuncAttributes fa{};
    cudaFuncGetAttributes(&fa, kern);
    const int blocks = occ < minblk ? occ : minblk;
    const int reps = 300;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i) launch(g_sms * blocks, reps);
    cudaEventRecord(e0);
    const int iters = 4;
    for (int i = 0; i < iters; ++i) launch(g_sms * blocks, reps);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    const double pairs = (double)g_sms * blocks * nth * R * TILE * reps;
    const double tp = pairs * iters / (ms * 1e-3) / 1e12;
    const double cyc = 148.0 * 4 * 1.965e9 / (tp * 1e12) * 32 * R; // cycles per warp-triangle at 1965 MHz
    printf("%-38s R=%-2d %4dthr x%d (%3d regs) %8.3f ms  %6.3f Tpairs/s  %5.1f cyc/warp-tri  %5.2f cyc/pair  %6.2f TFLOP/s FMA-pipe  %s\n", tag,
           R, nth, blocks, fa.numRegs, ms / iters, tp, cyc, cyc / R, tp * flops_per_pair, err == cudaSuccess ? "" : cudaGetErrorString(err));
    fflush(stdout);
}

// executed FP32 flops per pair: FADD = 1, FFMA = 2; per thread and triangle 4 (8 in QBAR) FFMA
#define SPAN(R, MODE, FORM, NTH, MB, G, NA)                                                                                                \
    run("SPAN " #MODE " " #FORM " g" #G " acc2x" #NA, k_span<R, MODE, FORM, NTH, MB, G, NA>, NTH, MB,                                      \
        [&](int grid, int reps) { k_span<R, MODE, FORM, NTH, MB, G, NA><<<grid, NTH>>>(tile_g, reps, out, 1e-3f, 65536.f); }, R,           \
        (MODE == MODE_SHAREDQ ? 8.0 / R : 16.0 / R) + (FORM == F_ADD ? 4.0 : 6.0))

int main() {
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, 0);
    std::vector<float> h(TILE * 8);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (i & 1) ? -1e4f - 3.f * (float)(i % 97) : 0.5f + 0.01f * (float)(i % 13); // never a candidate
    float4 *tile_g;
    float *out;
    cudaMalloc(&tile_g, h.size() * 4);
    cudaMemcpy(tile_g, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&out, 64);
    printf("-- span form, shared q\n");
    SPAN(8, MODE_SHAREDQ, F_ADD, 256, 3, 4, 2);
    SPAN(8, MODE_SHAREDQ, F_FMAI, 256, 3, 4, 2);
    SPAN(8, MODE_SHAREDQ, F_FMAR, 256, 3, 4, 2);
    SPAN(12, MODE_SHAREDQ, F_ADD, 256, 3, 4, 3);
    SPAN(12, MODE_SHAREDQ, F_FMAI, 256, 3, 4, 3);
    SPAN(12, MODE_SHAREDQ, F_FMAR, 256, 3, 4, 3);
    SPAN(12, MODE_SHAREDQ, F_ADD, 256, 2, 4, 3);
    SPAN(12, MODE_SHAREDQ, F_FMAI, 256, 2, 4, 3);
    SPAN(16, MODE_SHAREDQ, F_ADD, 256, 3, 4, 4);
    SPAN(16, MODE_SHAREDQ, F_FMAI, 256, 3, 4, 4);
    SPAN(16, MODE_SHAREDQ, F_FMAR, 256, 3, 4, 4);
    SPAN(16, MODE_SHAREDQ, F_ADD, 256, 2, 4, 4);
    SPAN(16, MODE_SHAREDQ, F_FMAI, 256, 2, 4, 4);
    SPAN(16, MODE_SHAREDQ, F_FMAI, 256, 3, 4, 2);
    SPAN(16, MODE_SHAREDQ, F_FMAI, 256, 3, 8, 4);
    SPAN(16, MODE_SHAREDQ, F_FMAI, 256, 3, 2, 4);
    SPAN(16, MODE_SHAREDQ, F_FMAI, 512, 1, 4, 4);
    SPAN(16, MODE_SHAREDQ, F_FMAI, 128, 6, 4, 4);
    SPAN(20, MODE_SHAREDQ, F_FMAI, 256, 3, 4, 5);
    SPAN(20, MODE_SHAREDQ, F_ADD, 256, 3, 4, 5);
    SPAN(24, MODE_SHAREDQ, F_FMAI, 256, 2, 4, 4);
    SPAN(24, MODE_SHAREDQ, F_ADD, 256, 2, 4, 4);
    SPAN(24, MODE_SHAREDQ, F_FMAI, 256, 2, 2, 3);
    SPAN(32, MODE_SHAREDQ, F_FMAI, 256, 2, 2, 4);
    SPAN(32, MODE_SHAREDQ, F_ADD, 256, 2, 2, 4);
    printf("-- span form, any-hit (qbar + |B| qdelta)\n");
    SPAN(8, MODE_QBAR, F_ADD, 256, 3, 4, 2);
    SPAN(8, MODE_QBAR, F_FMAI, 256, 3, 4, 2);
    SPAN(12, MODE_QBAR, F_ADD, 256, 3, 4, 3);
    SPAN(12, MODE_QBAR, F_FMAI, 256, 3, 4, 3);
    SPAN(16, MODE_QBAR, F_ADD, 256, 3, 4, 4);
    SPAN(16, MODE_QBAR, F_FMAI, 256, 3, 4, 4);
    SPAN(16, MODE_QBAR, F_FMAI, 256, 2, 4, 4);
    SPAN(4, MODE_QBAR, F_FMAI, 256, 3, 8, 2);
    SPAN(2, MODE_QBAR, F_FMAI, 256, 3, 8, 1);
    return 0;
}
