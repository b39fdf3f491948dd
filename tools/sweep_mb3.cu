// sweep_mb3.cu — round-2 microbenchmark: sign logic in the FMA pipe.  Saturating edge rows (fma.sat) + product
// accumulate (FMUL + FFMA) against the LOP3 sign test of sweep::eval_batch, at several occupancies.
// Rows are pre-scaled so that a true candidate saturates to exactly 1 on all three edges (product 1, accumulators >= 1).
#define SWEEP_NO_STRICT
#include "../esctp1raytracer_b200/csrc/sweep.cuh"
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
using namespace sweep;

// GROUP triangles share NACC accumulators; returns bit g set = group g may hold a candidate
template <int R, int MODE, int GROUP, int NACC, int UNROLL>
__device__ __forceinline__ unsigned eval_batch_sat(const float4 *__restrict__ tp, const float (&rp)[R], const float (&rq)[R], float qbar, float qdelta) {
    unsigned cand = 0;
#pragma unroll
    for (int g = 0; g < BATCH / GROUP; ++g) {
        float acc[NACC];
#pragma unroll
        for (int a = 0; a < NACC; ++a) acc[a] = 0.f;
#pragma unroll UNROLL
        for (int kk = 0; kk < GROUP; ++kk) {
            const int k = g * GROUP + kk;
            const float4 rb = tp[3 * k], rc = tp[3 * k + 1], rd = tp[3 * k + 2];
            if (MODE == MODE_OWNQ) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const float x = __saturatef(fmaf(rp[r], rb.x, fmaf(rq[r], rb.y, rb.z)));
                    const float y = __saturatef(fmaf(rp[r], rc.x, fmaf(rq[r], rc.y, rc.z)));
                    const float z = __saturatef(fmaf(rp[r], rd.x, fmaf(rq[r], rd.y, rd.z)));
                    acc[r % NACC] = fmaf(__fmul_rn(x, y), z, acc[r % NACC]);
                }
            } else {
                const float qx = MODE == MODE_QBAR ? qterm_qbar(rb, qbar, qdelta) : fmaf(rq[0], rb.y, rb.z);
                const float qy = MODE == MODE_QBAR ? qterm_qbar(rc, qbar, qdelta) : fmaf(rq[0], rc.y, rc.z);
                const float qz = MODE == MODE_QBAR ? qterm_qbar(rd, qbar, qdelta) : fmaf(rq[0], rd.y, rd.z);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const float x = __saturatef(fmaf(rp[r], rb.x, qx));
                    const float y = __saturatef(fmaf(rp[r], rc.x, qy));
                    const float z = __saturatef(fmaf(rp[r], rd.x, qz));
                    acc[r % NACC] = fmaf(__fmul_rn(x, y), z, acc[r % NACC]);
                }
            }
        }
        float s = acc[0];
#pragma unroll
        for (int a = 1; a < NACC; ++a) s += acc[a];
        if (s >= 1.f) cand |= 1u << g;
    }
    return cand;
}

// as eval_batch_sat, but the conjunction of two rays at a time in packed form: FMUL2 + FFMA2 (the saturating row
// evaluations stay scalar: fma.sat has no f32x2 form)
template <int R, int MODE, int GROUP, int NACC2>
__device__ __forceinline__ unsigned eval_batch_sat2(const float4 *__restrict__ tp, const float (&rp)[R], const float (&rq)[R], float qbar, float qdelta) {
    static_assert(R % 2 == 0, "ray pairs");
    unsigned cand = 0;
#pragma unroll
    for (int g = 0; g < BATCH / GROUP; ++g) {
        float2 acc[NACC2];
#pragma unroll
        for (int a = 0; a < NACC2; ++a) acc[a] = make_float2(0.f, 0.f);
#pragma unroll
        for (int kk = 0; kk < GROUP; ++kk) {
            const int k = g * GROUP + kk;
            const float4 rb = tp[3 * k], rc = tp[3 * k + 1], rd = tp[3 * k + 2];
            const float qx = MODE == MODE_QBAR ? qterm_qbar(rb, qbar, qdelta) : fmaf(rq[0], rb.y, rb.z);
            const float qy = MODE == MODE_QBAR ? qterm_qbar(rc, qbar, qdelta) : fmaf(rq[0], rc.y, rc.z);
            const float qz = MODE == MODE_QBAR ? qterm_qbar(rd, qbar, qdelta) : fmaf(rq[0], rd.y, rd.z);
#pragma unroll
            for (int r = 0; r < R; r += 2) {
                const float2 x = make_float2(__saturatef(fmaf(rp[r], rb.x, qx)), __saturatef(fmaf(rp[r + 1], rb.x, qx)));
                const float2 y = make_float2(__saturatef(fmaf(rp[r], rc.x, qy)), __saturatef(fmaf(rp[r + 1], rc.x, qy)));
                const float2 z = make_float2(__saturatef(fmaf(rp[r], rd.x, qz)), __saturatef(fmaf(rp[r + 1], rd.x, qz)));
                acc[(r / 2) % NACC2] = __ffma2_rn(__fmul2_rn(x, y), z, acc[(r / 2) % NACC2]);
            }
        }
        float s = 0.f;
#pragma unroll
        for (int a = 0; a < NACC2; ++a) s += acc[a].x + acc[a].y;
        if (s >= 1.f) cand |= 1u << g;
    }
    return cand;
}

template <int R, int MODE, int NTH, int MINBLK, int GROUP, int NACC2>
__global__ void __launch_bounds__(NTH, MINBLK) k_sat2(const float4 *tile_g, int reps, float *out, float seed) {
    __shared__ __align__(16) float4 tile[TILE * 3];
    for (int i = threadIdx.x; i < TILE * 3; i += blockDim.x) tile[i] = tile_g[i];
    __syncthreads();
    float rp[R], rq[R];
#pragma unroll
    for (int r = 0; r < R; ++r) rp[r] = seed * (threadIdx.x + 1) * (r + 1), rq[r] = seed * (threadIdx.x + 7) * 3;
    const float qbar = rq[0], qdelta = 1e-6f * (1 + (threadIdx.x & 3));
    int hits = 0;
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll 1
        for (int b0 = 0; b0 < TILE; b0 += BATCH) {
            const unsigned c = eval_batch_sat2<R, MODE, GROUP, NACC2>(tile + 3 * b0, rp, rq, qbar, qdelta);
            if (c) {
                hits += 1;
                asm volatile("" ::: "memory");
            }
        }
    }
    if (hits == 123456789) out[0] = hits;
}

template <int R, int MODE, int NTH, int MINBLK, int GROUP, int NACC, int UNROLL>
__global__ void __launch_bounds__(NTH, MINBLK) k_sat(const float4 *tile_g, int reps, float *out, float seed) {
    __shared__ __align__(16) float4 tile[TILE * 3];
    for (int i = threadIdx.x; i < TILE * 3; i += blockDim.x) tile[i] = tile_g[i];
    __syncthreads();
    float rp[R], rq[R];
#pragma unroll
    for (int r = 0; r < R; ++r) rp[r] = seed * (threadIdx.x + 1) * (r + 1), rq[r] = seed * (threadIdx.x + 7) * (MODE == MODE_OWNQ ? r + 3 : 3);
    const float qbar = rq[0], qdelta = 1e-6f * (1 + (threadIdx.x & 3));
    int hits = 0;
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll 1
        for (int b0 = 0; b0 < TILE; b0 += BATCH) {
            const unsigned c = eval_batch_sat<R, MODE, GROUP, NACC, UNROLL>(tile + 3 * b0, rp, rq, qbar, qdelta);
            if (c) {
                hits += 1;
                asm volatile("" ::: "memory");
            }
        }
    }
    if (hits == 123456789) out[0] = hits;
}

template <int R, int MODE, int NTH, int MINBLK, int UNROLL>
__global__ void __launch_bounds__(NTH, MINBLK) k_lop(const float4 *tile_g, int reps, float *out, float seed) {
    __shared__ __align__(16) float4 tile[TILE * 3];
    for (int i = threadIdx.x; i < TILE * 3; i += blockDim.x) tile[i] = tile_g[i];
    __syncthreads();
    float rp[R], rq[R];
#pragma unroll
    for (int r = 0; r < R; ++r) rp[r] = seed * (threadIdx.x + 1) * (r + 1), rq[r] = seed * (threadIdx.x + 7) * (MODE == MODE_OWNQ ? r + 3 : 3);
    const float qbar = rq[0], qdelta = 1e-6f * (1 + (threadIdx.x & 3));
    int hits = 0;
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll 1
        for (int b0 = 0; b0 < TILE; b0 += BATCH) {
            const unsigned neg = eval_batch_lop3<R, MODE, UNROLL>(tile + 3 * b0, rp, rq, qbar, qdelta);
            if (~neg & 0xffffu) {
                hits += 1;
                asm volatile("" ::: "memory");
            }
        }
    }
    if (hits == 123456789) out[0] = hits;
}

static int g_sms = 0;
template <typename K, typename F>
void run(const char *tag, K kern, int nth, int minblk, F launch, int R, double fma_flops_per_pair) {
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
    printf("%-34s R=%-2d %4dthr x%d (%3d regs) %8.3f ms  %6.3f Tpairs/s  %5.1f cyc/warp-tri  %6.2f TFLOP/s FMA-pipe  %s\n", tag, R, nth, blocks, fa.numRegs,
           ms / iters, tp, cyc, tp * fma_flops_per_pair, err == cudaSuccess ? "" : cudaGetErrorString(err));
    fflush(stdout);
}

#define SAT(R, MODE, NTH, MB, G, NA, U)                                                                                                   \
    run("SAT " #MODE " g" #G " acc" #NA " u" #U, k_sat<R, MODE, NTH, MB, G, NA, U>, NTH, MB,                                               \
        [&](int grid, int reps) { k_sat<R, MODE, NTH, MB, G, NA, U><<<grid, NTH>>>(tile_g, reps, out, 1e-3f); }, R,                        \
        (MODE == MODE_SHAREDQ ? 2.0 * 3 / R : MODE == MODE_QBAR ? 2.0 * 6 / R : 6.0) + 9.0)
#define LOP(R, MODE, NTH, MB, U)                                                                                                          \
    run("LOP3 " #MODE " u" #U, k_lop<R, MODE, NTH, MB, U>, NTH, MB,                                                                        \
        [&](int grid, int reps) { k_lop<R, MODE, NTH, MB, U><<<grid, NTH>>>(tile_g, reps, out, 1e-3f); }, R,                               \
        (MODE == MODE_SHAREDQ ? 2.0 * 3 / R : MODE == MODE_QBAR ? 2.0 * 6 / R : 6.0) + 6.0)

#define SAT2(R, MODE, NTH, MB, G, NA)                                                                                                    \
    run("SAT+FMUL2/FFMA2 " #MODE " g" #G " acc2x" #NA, k_sat2<R, MODE, NTH, MB, G, NA>, NTH, MB,                                          \
        [&](int grid, int reps) { k_sat2<R, MODE, NTH, MB, G, NA><<<grid, NTH>>>(tile_g, reps, out, 1e-3f); }, R,                          \
        (MODE == MODE_SHAREDQ ? 2.0 * 3 / R : 2.0 * 6 / R) + 9.0)

int main() {
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, 0);
    std::vector<float> h(TILE * 12);
    for (size_t i = 0; i < h.size(); ++i) h[i] = -0.5f - 0.001f * (float)(i % 97); // never a candidate
    float4 *tile_g;
    float *out;
    cudaMalloc(&tile_g, h.size() * 4);
    cudaMemcpy(tile_g, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&out, 64);
    printf("-- packed conjunction (FMUL2 + FFMA2 per ray pair)\n");
    SAT(8, MODE_SHAREDQ, 256, 3, 4, 4, 4);
    SAT2(8, MODE_SHAREDQ, 256, 3, 4, 1);
    SAT2(8, MODE_SHAREDQ, 256, 3, 4, 2);
    SAT2(8, MODE_SHAREDQ, 256, 3, 4, 4);
    SAT2(8, MODE_SHAREDQ, 256, 2, 4, 2);
    SAT2(8, MODE_SHAREDQ, 256, 3, 8, 2);
    SAT2(8, MODE_SHAREDQ, 256, 3, 2, 2);
    SAT2(12, MODE_SHAREDQ, 256, 2, 4, 3);
    SAT2(16, MODE_SHAREDQ, 256, 2, 4, 4);
    SAT2(8, MODE_QBAR, 256, 3, 4, 2);
    SAT2(8, MODE_QBAR, 256, 3, 8, 2);
    SAT2(12, MODE_SHAREDQ, 256, 3, 4, 3);
    SAT2(12, MODE_SHAREDQ, 256, 3, 8, 3);
    SAT2(12, MODE_SHAREDQ, 256, 3, 8, 2);
    SAT2(12, MODE_SHAREDQ, 256, 2, 8, 3);
    SAT2(12, MODE_SHAREDQ, 256, 3, 16, 3);
    SAT2(10, MODE_SHAREDQ, 256, 3, 8, 5);
    SAT2(16, MODE_SHAREDQ, 256, 2, 8, 4);
    SAT2(16, MODE_SHAREDQ, 256, 3, 8, 2);
    SAT2(12, MODE_QBAR, 256, 3, 4, 3);
    SAT2(12, MODE_QBAR, 256, 3, 8, 3);
    SAT2(12, MODE_QBAR, 256, 2, 8, 3);
    SAT2(16, MODE_QBAR, 256, 2, 8, 4);
    SAT2(4, MODE_QBAR, 256, 3, 8, 2);
    SAT2(2, MODE_QBAR, 256, 3, 8, 1);
    printf("-- round-1 form and the scalar saturating form\n");
    LOP(8, MODE_SHAREDQ, 512, 2, 4);
    LOP(8, MODE_SHAREDQ, 512, 1, 4);
    SAT(8, MODE_SHAREDQ, 512, 1, 4, 2, 4);
    SAT(8, MODE_SHAREDQ, 512, 2, 4, 2, 4);
    SAT(8, MODE_SHAREDQ, 512, 2, 4, 1, 4);
    SAT(8, MODE_SHAREDQ, 512, 2, 4, 4, 4);
    SAT(8, MODE_SHAREDQ, 512, 2, 4, 8, 4);
    SAT(8, MODE_SHAREDQ, 512, 2, 4, 2, 2);
    SAT(8, MODE_SHAREDQ, 512, 2, 4, 2, 1);
    SAT(8, MODE_SHAREDQ, 512, 2, 8, 2, 4);
    SAT(8, MODE_SHAREDQ, 512, 2, 8, 4, 4);
    SAT(8, MODE_SHAREDQ, 512, 2, 16, 2, 4);
    SAT(8, MODE_SHAREDQ, 512, 2, 16, 4, 4);
    SAT(8, MODE_SHAREDQ, 512, 2, 2, 2, 2);
    SAT(8, MODE_SHAREDQ, 256, 3, 4, 2, 4);
    SAT(8, MODE_SHAREDQ, 256, 4, 4, 2, 4);
    SAT(8, MODE_SHAREDQ, 384, 2, 4, 2, 4);
    SAT(8, MODE_SHAREDQ, 1024, 1, 4, 2, 4);
    SAT(4, MODE_SHAREDQ, 512, 2, 4, 2, 4);
    SAT(6, MODE_SHAREDQ, 512, 2, 4, 2, 4);
    SAT(10, MODE_SHAREDQ, 512, 2, 4, 2, 4);
    SAT(12, MODE_SHAREDQ, 512, 2, 4, 2, 2);
    SAT(12, MODE_SHAREDQ, 512, 2, 4, 4, 2);
    SAT(16, MODE_SHAREDQ, 512, 2, 4, 4, 2);
    SAT(16, MODE_SHAREDQ, 256, 3, 4, 4, 2);
    SAT(16, MODE_SHAREDQ, 384, 2, 4, 4, 2);
    printf("-- any-hit (QBAR)\n");
    LOP(8, MODE_QBAR, 512, 2, 4);
    SAT(8, MODE_QBAR, 512, 2, 4, 2, 4);
    SAT(8, MODE_QBAR, 512, 1, 4, 2, 4);
    SAT(8, MODE_QBAR, 512, 2, 8, 2, 4);
    SAT(8, MODE_QBAR, 256, 3, 4, 2, 4);
    SAT(12, MODE_QBAR, 512, 2, 4, 4, 2);
    SAT(4, MODE_QBAR, 512, 2, 4, 2, 4);
    SAT(2, MODE_QBAR, 512, 2, 4, 2, 4);
    printf("-- own q per ray\n");
    LOP(8, MODE_OWNQ, 512, 2, 4);
    SAT(8, MODE_OWNQ, 512, 2, 4, 2, 4);
    SAT(8, MODE_OWNQ, 512, 1, 4, 2, 4);
    SAT(8, MODE_OWNQ, 512, 2, 8, 2, 4);
    return 0;
}
