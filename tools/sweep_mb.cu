// sweep_mb.cu — development microbenchmark of the sweep's inner loop (no TMA, no strict path):
// the tile sits in shared memory and is swept repeatedly, so only the issue/pipe behaviour of
// each formulation is measured.  Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a
//        -lineinfo tools/sweep_mb.cu -o tools/sweep_mb.bin ; run on the GPU box.
#define SWEEP_NO_STRICT
#include "../esctp1raytracer_b200/csrc/sweep.cuh"
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

constexpr int TILE = 256;

__device__ __forceinline__ float2 edge_min_tp(const float4 *q, float2 ex, float2 ey, float2 ez) {
    const float2 X = __ffma2_rn(ex, make_float2(q[0].x, q[0].y),
                                __ffma2_rn(ey, make_float2(q[0].z, q[0].w), __ffma2_rn(ez, make_float2(q[1].x, q[1].y), make_float2(q[1].z, q[1].w))));
    const float2 Y = __ffma2_rn(ex, make_float2(q[2].x, q[2].y),
                                __ffma2_rn(ey, make_float2(q[2].z, q[2].w), __ffma2_rn(ez, make_float2(q[3].x, q[3].y), make_float2(q[3].z, q[3].w))));
    const float2 Z = __ffma2_rn(ex, make_float2(q[4].x, q[4].y),
                                __ffma2_rn(ey, make_float2(q[4].z, q[4].w), __ffma2_rn(ez, make_float2(q[5].x, q[5].y), make_float2(q[5].z, q[5].w))));
    return make_float2(fminf(fminf(X.x, Y.x), Z.x), fminf(fminf(X.y, Y.y), Z.y));
}

// V0: triangle-pair FFMA2, ray direction broadcast (production v2)
template <int R, int UNROLL>
__global__ void __launch_bounds__(512, 1) k_tripair(const float4 *tile_g, int reps, float *out, float seed) {
    __shared__ __align__(16) float4 tile[TILE / 2 * 6];
    for (int i = threadIdx.x; i < TILE / 2 * 6; i += blockDim.x) tile[i] = tile_g[i];
    __syncthreads();
    float2 ex[R], ey[R], ez[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const float a = seed * (threadIdx.x + 1) * (r + 1), b = seed * (threadIdx.x + 7) * (r + 3), c = -1.f - seed * r;
        ex[r] = make_float2(a, a), ey[r] = make_float2(b, b), ez[r] = make_float2(c, c);
    }
    int hits = 0;
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll UNROLL
        for (int pi = 0; pi < TILE / 2; ++pi) {
            float4 q[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) q[j] = tile[6 * pi + j];
            float M = -1.f;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float2 m = edge_min_tp(q, ex[r], ey[r], ez[r]);
                M = fmaxf(fmaxf(M, m.x), m.y);
            }
            if (M >= 0.f) {
                hits += 1;
                asm volatile("" ::: "memory");
            }
        }
    }
    if (hits == 123456789) out[0] = hits;
}

// V1: ray-pair FFMA2: coefficients broadcast, two rays per packed lane
template <int R2, int UNROLL>  // R2 = ray pairs per thread
__global__ void __launch_bounds__(512, 1) k_raypair(const float4 *tile_g, int reps, float *out, float seed) {
    __shared__ __align__(16) float4 tile[TILE * 3];
    for (int i = threadIdx.x; i < TILE * 3; i += blockDim.x) tile[i] = tile_g[i];
    __syncthreads();
    float2 ex[R2], ey[R2], ez[R2];
#pragma unroll
    for (int r = 0; r < R2; ++r) {
        const float a = seed * (threadIdx.x + 1) * (r + 1), b = seed * (threadIdx.x + 7) * (r + 3);
        ex[r] = make_float2(a, a * 1.01f), ey[r] = make_float2(b, b * 0.99f), ez[r] = make_float2(-1.f - seed * r, -1.01f);
    }
    int hits = 0;
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll UNROLL
        for (int i = 0; i < TILE; ++i) {
            const float4 rb = tile[3 * i], rc = tile[3 * i + 1], rd = tile[3 * i + 2];
            float M = -1.f;
#pragma unroll
            for (int r = 0; r < R2; ++r) {
                const float2 X = __ffma2_rn(ex[r], make_float2(rb.x, rb.x), __ffma2_rn(ey[r], make_float2(rb.y, rb.y), __ffma2_rn(ez[r], make_float2(rb.z, rb.z), make_float2(rb.w, rb.w))));
                const float2 Y = __ffma2_rn(ex[r], make_float2(rc.x, rc.x), __ffma2_rn(ey[r], make_float2(rc.y, rc.y), __ffma2_rn(ez[r], make_float2(rc.z, rc.z), make_float2(rc.w, rc.w))));
                const float2 Z = __ffma2_rn(ex[r], make_float2(rd.x, rd.x), __ffma2_rn(ey[r], make_float2(rd.y, rd.y), __ffma2_rn(ez[r], make_float2(rd.z, rd.z), make_float2(rd.w, rd.w))));
                M = fmaxf(fmaxf(M, fminf(fminf(X.x, Y.x), Z.x)), fminf(fminf(X.y, Y.y), Z.y));
            }
            if (M >= 0.f) {
                hits += 1;
                asm volatile("" ::: "memory");
            }
        }
    }
    if (hits == 123456789) out[0] = hits;
}

// V2: scalar FFMA (v1 formulation)
template <int R, int UNROLL, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_scalar(const float4 *tile_g, int reps, float *out, float seed) {
    __shared__ __align__(16) float4 tile[TILE * 3];
    for (int i = threadIdx.x; i < TILE * 3; i += blockDim.x) tile[i] = tile_g[i];
    __syncthreads();
    float ex[R], ey[R], ez[R];
#pragma unroll
    for (int r = 0; r < R; ++r) ex[r] = seed * (threadIdx.x + 1) * (r + 1), ey[r] = seed * (threadIdx.x + 7) * (r + 3), ez[r] = -1.f - seed * r;
    int hits = 0;
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll UNROLL
        for (int i = 0; i < TILE; ++i) {
            const float4 rb = tile[3 * i], rc = tile[3 * i + 1], rd = tile[3 * i + 2];
            float M = -1.f;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float x = fmaf(ex[r], rb.x, fmaf(ey[r], rb.y, fmaf(ez[r], rb.z, rb.w)));
                const float y = fmaf(ex[r], rc.x, fmaf(ey[r], rc.y, fmaf(ez[r], rc.z, rc.w)));
                const float z = fmaf(ex[r], rd.x, fmaf(ey[r], rd.y, fmaf(ez[r], rd.z, rd.w)));
                M = fmaxf(M, fminf(fminf(x, y), z));
            }
            if (M >= 0.f) {
                hits += 1;
                asm volatile("" ::: "memory");
            }
        }
    }
    if (hits == 123456789) out[0] = hits;
}

// V3: triangle-pair FFMA2 but sign test through integer OR instead of FMNMX (LOP3 on the ALU pipe)
template <int R, int UNROLL>
__global__ void __launch_bounds__(512, 1) k_tripair_lop(const float4 *tile_g, int reps, float *out, float seed) {
    __shared__ __align__(16) float4 tile[TILE / 2 * 6];
    for (int i = threadIdx.x; i < TILE / 2 * 6; i += blockDim.x) tile[i] = tile_g[i];
    __syncthreads();
    float2 ex[R], ey[R], ez[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const float a = seed * (threadIdx.x + 1) * (r + 1), b = seed * (threadIdx.x + 7) * (r + 3), c = -1.f - seed * r;
        ex[r] = make_float2(a, a), ey[r] = make_float2(b, b), ez[r] = make_float2(c, c);
    }
    int hits = 0;
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll UNROLL
        for (int pi = 0; pi < TILE / 2; ++pi) {
            float4 q[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) q[j] = tile[6 * pi + j];
            unsigned A = 0xffffffffu; // sign bit of A stays set unless some ray has all three >= 0 for some triangle
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float2 X = __ffma2_rn(ex[r], make_float2(q[0].x, q[0].y), __ffma2_rn(ey[r], make_float2(q[0].z, q[0].w), __ffma2_rn(ez[r], make_float2(q[1].x, q[1].y), make_float2(q[1].z, q[1].w))));
                const float2 Y = __ffma2_rn(ex[r], make_float2(q[2].x, q[2].y), __ffma2_rn(ey[r], make_float2(q[2].z, q[2].w), __ffma2_rn(ez[r], make_float2(q[3].x, q[3].y), make_float2(q[3].z, q[3].w))));
                const float2 Z = __ffma2_rn(ex[r], make_float2(q[4].x, q[4].y), __ffma2_rn(ey[r], make_float2(q[4].z, q[4].w), __ffma2_rn(ez[r], make_float2(q[5].x, q[5].y), make_float2(q[5].z, q[5].w))));
                const unsigned s0 = __float_as_uint(X.x) | __float_as_uint(Y.x) | __float_as_uint(Z.x);
                const unsigned s1 = __float_as_uint(X.y) | __float_as_uint(Y.y) | __float_as_uint(Z.y);
                A &= s0 & s1;
            }
            if ((int)A >= 0) {
                hits += 1;
                asm volatile("" ::: "memory");
            }
        }
    }
    if (hits == 123456789) out[0] = hits;
}

// V4: triangle-pair layout, NF2 of the three rows as packed FFMA2 and the rest as scalar FFMA;
//     per-pair branch replaced by a sign-bit shift register tested once per batch of BATCH pairs.
template <int R, int NF2, int BATCH, int UNROLL>
__global__ void __launch_bounds__(512, 1) k_mixed(const float4 *tile_g, int reps, float *out, float seed) {
    __shared__ __align__(16) float4 tile[TILE / 2 * 6];
    for (int i = threadIdx.x; i < TILE / 2 * 6; i += blockDim.x) tile[i] = tile_g[i];
    __syncthreads();
    float ex[R], ey[R], ez[R];
#pragma unroll
    for (int r = 0; r < R; ++r) ex[r] = seed * (threadIdx.x + 1) * (r + 1), ey[r] = seed * (threadIdx.x + 7) * (r + 3), ez[r] = -1.f - seed * r;
    int hits = 0;
    for (int rep = 0; rep < reps; ++rep) {
        for (int b = 0; b < TILE / 2; b += BATCH) {
            unsigned neg = 0xffffffffu;
#pragma unroll UNROLL
            for (int k = 0; k < BATCH; ++k) {
                const float4 *q = &tile[6 * (b + k)];
                float M = -1.f;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    float m0[3], m1[3];
#pragma unroll
                    for (int row = 0; row < 3; ++row) {
                        const float4 qa = q[2 * row], qb = q[2 * row + 1];
                        if (row < NF2) {
                            const float2 X = __ffma2_rn(make_float2(ex[r], ex[r]), make_float2(qa.x, qa.y),
                                                        __ffma2_rn(make_float2(ey[r], ey[r]), make_float2(qa.z, qa.w),
                                                                   __ffma2_rn(make_float2(ez[r], ez[r]), make_float2(qb.x, qb.y), make_float2(qb.z, qb.w))));
                            m0[row] = X.x, m1[row] = X.y;
                        } else {
                            m0[row] = fmaf(ex[r], qa.x, fmaf(ey[r], qa.z, fmaf(ez[r], qb.x, qb.z)));
                            m1[row] = fmaf(ex[r], qa.y, fmaf(ey[r], qa.w, fmaf(ez[r], qb.y, qb.w)));
                        }
                    }
                    M = fmaxf(fmaxf(M, fminf(fminf(m0[0], m0[1]), m0[2])), fminf(fminf(m1[0], m1[1]), m1[2]));
                }
                neg = __funnelshift_l(__float_as_uint(M), neg, 1); // shift in the sign bit: 0 = candidate
            }
            if (~neg & ((BATCH >= 32) ? 0xffffffffu : ((1u << BATCH) - 1u))) {
                hits += 1;
                asm volatile("" ::: "memory");
            }
        }
    }
    if (hits == 123456789) out[0] = hits;
}

// V6: 2-D rows (A,B,C per edge: 6 FFMA per test), sign test via LOP3 (SIGN=1) or FMNMX3 (SIGN=0); 3-D rows when DIM3
template <int R, int SIGN, int DIM3, int BATCH, int UNROLL>
__global__ void __launch_bounds__(512, 1) k_2d(const float4 *tile_g, int reps, float *out, float seed) {
    __shared__ __align__(16) float4 tile[TILE * 3];
    for (int i = threadIdx.x; i < TILE * 3; i += blockDim.x) tile[i] = tile_g[i];
    __syncthreads();
    float px[R], py[R], pz[R];
#pragma unroll
    for (int r = 0; r < R; ++r) px[r] = seed * (threadIdx.x + 1) * (r + 1), py[r] = seed * (threadIdx.x + 7) * (r + 3), pz[r] = -1.f - seed * r;
    int hits = 0;
    for (int rep = 0; rep < reps; ++rep) {
        for (int b = 0; b < TILE; b += BATCH) {
            unsigned neg = 0xffffffffu;
#pragma unroll UNROLL
            for (int k = 0; k < BATCH; ++k) {
                const float4 rb = tile[3 * (b + k)], rc = tile[3 * (b + k) + 1], rd = tile[3 * (b + k) + 2];
                float M = -1.f;
                unsigned A = 0xffffffffu;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    float x, y, z;
                    if (DIM3) {
                        x = fmaf(px[r], rb.x, fmaf(py[r], rb.y, fmaf(pz[r], rb.z, rb.w)));
                        y = fmaf(px[r], rc.x, fmaf(py[r], rc.y, fmaf(pz[r], rc.z, rc.w)));
                        z = fmaf(px[r], rd.x, fmaf(py[r], rd.y, fmaf(pz[r], rd.z, rd.w)));
                    } else {
                        x = fmaf(px[r], rb.x, fmaf(py[r], rb.y, rb.z));
                        y = fmaf(px[r], rc.x, fmaf(py[r], rc.y, rc.z));
                        z = fmaf(px[r], rd.x, fmaf(py[r], rd.y, rd.z));
                    }
                    if (SIGN)
                        A &= __float_as_uint(x) | __float_as_uint(y) | __float_as_uint(z);
                    else
                        M = fmaxf(M, fminf(fminf(x, y), z));
                }
                neg = __funnelshift_l(SIGN ? A : __float_as_uint(M), neg, 1);
            }
            if (~neg & ((BATCH >= 32) ? 0xffffffffu : ((1u << BATCH) - 1u))) {
                hits += 1;
                asm volatile("" ::: "memory");
            }
        }
    }
    if (hits == 123456789) out[0] = hits;
}


// V7: 2-D rows, saturating rows + product accumulate: everything in the FMA pipe (no LOP3).
// rows are pre-scaled so that a true candidate saturates to exactly 1: ind = x'*y'*z', acc += ind over the batch
template <int R, int NACC, int BATCH, int UNROLL>
__global__ void __launch_bounds__(512, 1) k_sat(const float4 *tile_g, int reps, float *out, float seed) {
    __shared__ __align__(16) float4 tile[TILE * 3];
    for (int i = threadIdx.x; i < TILE * 3; i += blockDim.x) tile[i] = tile_g[i];
    __syncthreads();
    float px[R], py[R];
#pragma unroll
    for (int r = 0; r < R; ++r) px[r] = seed * (threadIdx.x + 1) * (r + 1), py[r] = seed * (threadIdx.x + 7) * (r + 3);
    int hits = 0;
    for (int rep = 0; rep < reps; ++rep) {
        for (int b = 0; b < TILE; b += BATCH) {
            float acc[NACC];
#pragma unroll
            for (int a = 0; a < NACC; ++a) acc[a] = 0.f;
#pragma unroll UNROLL
            for (int k = 0; k < BATCH; ++k) {
                const float4 rb = tile[3 * (b + k)], rc = tile[3 * (b + k) + 1], rd = tile[3 * (b + k) + 2];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const float x = __saturatef(fmaf(px[r], rb.x, fmaf(py[r], rb.y, rb.z)));
                    const float y = __saturatef(fmaf(px[r], rc.x, fmaf(py[r], rc.y, rc.z)));
                    const float z = __saturatef(fmaf(px[r], rd.x, fmaf(py[r], rd.y, rd.z)));
                    acc[r % NACC] = fmaf(x * y, z, acc[r % NACC]);
                }
            }
            float s = acc[0];
#pragma unroll
            for (int a = 1; a < NACC; ++a) s += acc[a];
            if (s >= 1.f) {
                hits += 1;
                asm volatile("" ::: "memory");
            }
        }
    }
    if (hits == 123456789) out[0] = hits;
}

// V6q: as V6 but the R rays of a thread share q (py), sign test via LOP3 (SIGN=1) or FMNMX3 (SIGN=0); 3-D rows when DIM3
template <int R, int SIGN, int DIM3, int BATCH, int UNROLL>
__global__ void __launch_bounds__(512, 1) k_2d_sq(const float4 *tile_g, int reps, float *out, float seed) {
    __shared__ __align__(16) float4 tile[TILE * 3];
    for (int i = threadIdx.x; i < TILE * 3; i += blockDim.x) tile[i] = tile_g[i];
    __syncthreads();
    float px[R], py[R], pz[R];
#pragma unroll
    for (int r = 0; r < R; ++r) px[r] = seed * (threadIdx.x + 1) * (r + 1), py[r] = seed * (threadIdx.x + 7) * 3, pz[r] = -1.f - seed * r;
    int hits = 0;
    for (int rep = 0; rep < reps; ++rep) {
        for (int b = 0; b < TILE; b += BATCH) {
            unsigned neg = 0xffffffffu;
#pragma unroll UNROLL
            for (int k = 0; k < BATCH; ++k) {
                const float4 rb = tile[3 * (b + k)], rc = tile[3 * (b + k) + 1], rd = tile[3 * (b + k) + 2];
                float M = -1.f;
                unsigned A = 0xffffffffu;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    float x, y, z;
                    if (DIM3) {
                        x = fmaf(px[r], rb.x, fmaf(py[r], rb.y, fmaf(pz[r], rb.z, rb.w)));
                        y = fmaf(px[r], rc.x, fmaf(py[r], rc.y, fmaf(pz[r], rc.z, rc.w)));
                        z = fmaf(px[r], rd.x, fmaf(py[r], rd.y, fmaf(pz[r], rd.z, rd.w)));
                    } else {
                        x = fmaf(px[r], rb.x, fmaf(py[r], rb.y, rb.z));
                        y = fmaf(px[r], rc.x, fmaf(py[r], rc.y, rc.z));
                        z = fmaf(px[r], rd.x, fmaf(py[r], rd.y, rd.z));
                    }
                    if (SIGN)
                        A &= __float_as_uint(x) | __float_as_uint(y) | __float_as_uint(z);
                    else
                        M = fmaxf(M, fminf(fminf(x, y), z));
                }
                neg = __funnelshift_l(SIGN ? A : __float_as_uint(M), neg, 1);
            }
            if (~neg & ((BATCH >= 32) ? 0xffffffffu : ((1u << BATCH) - 1u))) {
                hits += 1;
                asm volatile("" ::: "memory");
            }
        }
    }
    if (hits == 123456789) out[0] = hits;
}



// V7q: as V7 with shared q: everything in the FMA pipe (no LOP3).
// rows are pre-scaled so that a true candidate saturates to exactly 1: ind = x'*y'*z', acc += ind over the batch
template <int R, int NACC, int BATCH, int UNROLL>
__global__ void __launch_bounds__(512, 1) k_sat_sq(const float4 *tile_g, int reps, float *out, float seed) {
    __shared__ __align__(16) float4 tile[TILE * 3];
    for (int i = threadIdx.x; i < TILE * 3; i += blockDim.x) tile[i] = tile_g[i];
    __syncthreads();
    float px[R], py[R];
#pragma unroll
    for (int r = 0; r < R; ++r) px[r] = seed * (threadIdx.x + 1) * (r + 1), py[r] = seed * (threadIdx.x + 7) * 3;
    int hits = 0;
    for (int rep = 0; rep < reps; ++rep) {
        for (int b = 0; b < TILE; b += BATCH) {
            float acc[NACC];
#pragma unroll
            for (int a = 0; a < NACC; ++a) acc[a] = 0.f;
#pragma unroll UNROLL
            for (int k = 0; k < BATCH; ++k) {
                const float4 rb = tile[3 * (b + k)], rc = tile[3 * (b + k) + 1], rd = tile[3 * (b + k) + 2];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const float x = __saturatef(fmaf(px[r], rb.x, fmaf(py[r], rb.y, rb.z)));
                    const float y = __saturatef(fmaf(px[r], rc.x, fmaf(py[r], rc.y, rc.z)));
                    const float z = __saturatef(fmaf(px[r], rd.x, fmaf(py[r], rd.y, rd.z)));
                    acc[r % NACC] = fmaf(x * y, z, acc[r % NACC]);
                }
            }
            float s = acc[0];
#pragma unroll
            for (int a = 1; a < NACC; ++a) s += acc[a];
            if (s >= 1.f) {
                hits += 1;
                asm volatile("" ::: "memory");
            }
        }
    }
    if (hits == 123456789) out[0] = hits;
}


// VP: the production data path (TMA tile stream + mbarrier + per-tile barrier) without the strict path
template <int R>
__global__ void __launch_bounds__(sweep::THREADS, 1) k_prod(const float4 *table, int n_tiles, int n_blocks, int *work, float *out, float seed) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    sweep::Smem<R> &sm = *reinterpret_cast<sweep::Smem<R> *>(smem_raw);
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < sweep::STAGES; ++s) sweep::mbar_init(&sm.full_bar[s], 1), sm.consumed[s] = 0;
        sweep::fence_barrier_init();
    }
    __syncthreads();
    unsigned gtile = 0, n_strict = 0, n_swept = 0, n_miss = 0;
    for (;;) {
        if (tid == 0) sm.blk = atomicAdd(work, 1);
        __syncthreads();
        const int blk = sm.blk;
        if (blk >= n_blocks) break;
        float ex[R], ey[R];
#pragma unroll
        for (int r = 0; r < R; ++r) ex[r] = seed * (tid + 1) * (r + 1 + blk), ey[r] = seed * (tid + 7) * (r + 3);
        unsigned done = 0;
        sweep::sweep_table<R, false, false, false>(sm, table, 0, n_tiles, n_tiles * 256, nullptr, sweep::RaySrc{}, ex, ey, 0.f, 0.f, 0xffu >> (8 - R), done, gtile,
                                            n_strict, n_swept, n_miss);
        __syncthreads();
    }
    if (n_strict == 123456789u) out[0] = n_strict;
}

template <typename F>
double run(const char *name, F launch, double pairs_per_launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i) launch();
    cudaEventRecord(e0);
    const int iters = 5;
    for (int i = 0; i < iters; ++i) launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    const double tf = 18.0 * pairs_per_launch * iters / (ms * 1e-3) / 1e12;
    printf("%-34s %8.3f ms/launch  %7.2f TFLOP/s-equiv  %s\n", name, ms / iters, tf, err == cudaSuccess ? "" : cudaGetErrorString(err));
    return tf;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    std::vector<float> h(TILE * 12);
    for (size_t i = 0; i < h.size(); ++i) h[i] = -0.5f - 0.001f * (float)(i % 97); // never a candidate
    float4 *tile_g;
    float *out;
    cudaMalloc(&tile_g, h.size() * 4);
    cudaMemcpy(tile_g, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&out, 64);
    const int reps = 400;
    const float seed = 1e-3f;
#define PAIRS(threads, R) ((double)sms * (threads) * (R) * TILE * reps)
    run("tripair R=4 u2 512thr", [&] { k_tripair<4, 2><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 4));
    run("tripair R=8 u2 512thr", [&] { k_tripair<8, 2><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 8));
    run("tripair R=8 u1 512thr", [&] { k_tripair<8, 1><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 8));
    run("tripair R=8 u4 512thr", [&] { k_tripair<8, 4><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 8));
    run("tripair R=6 u2 512thr", [&] { k_tripair<6, 2><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 6));
    run("tripair R=8 u2 256thr", [&] { k_tripair<8, 2><<<sms, 256>>>(tile_g, reps, out, seed); }, PAIRS(256, 8));
    run("tripair R=8 u2 384thr", [&] { k_tripair<8, 2><<<sms, 384>>>(tile_g, reps, out, seed); }, PAIRS(384, 8));
    run("tripair-lop R=8 u2 512thr", [&] { k_tripair_lop<8, 2><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 8));
    run("tripair-lop R=4 u2 512thr", [&] { k_tripair_lop<4, 2><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 4));
    run("raypair R=8 (4 pairs) u2", [&] { k_raypair<4, 2><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 8));
    run("raypair R=8 (4 pairs) u4", [&] { k_raypair<4, 4><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 8));
    run("raypair R=4 (2 pairs) u4", [&] { k_raypair<2, 4><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 4));
    run("raypair R=16 (8 pairs) u2", [&] { k_raypair<8, 2><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 16));
    run("scalar R=8 u2 512thr", [&] { k_scalar<8, 2, 512><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 8));
    run("scalar R=8 u4 256thr", [&] { k_scalar<8, 4, 256><<<sms, 256>>>(tile_g, reps, out, seed); }, PAIRS(256, 8));
    run("scalar R=4 u4 512thr", [&] { k_scalar<4, 4, 512><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 4));
    run("mixed R=8 NF2=3 batch8 u4", [&] { k_mixed<8, 3, 8, 4><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 8));
    run("mixed R=8 NF2=3 batch16 u4", [&] { k_mixed<8, 3, 16, 4><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 8));
    run("mixed R=8 NF2=3 batch8 u8", [&] { k_mixed<8, 3, 8, 8><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 8));
    run("mixed R=8 NF2=2 batch8 u4", [&] { k_mixed<8, 2, 8, 4><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 8));
    run("mixed R=8 NF2=1 batch8 u4", [&] { k_mixed<8, 1, 8, 4><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 8));
    run("mixed R=8 NF2=0 batch8 u4", [&] { k_mixed<8, 0, 8, 4><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 8));
    run("mixed R=4 NF2=3 batch8 u4", [&] { k_mixed<4, 3, 8, 4><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 4));
    run("mixed R=4 NF2=2 batch8 u4", [&] { k_mixed<4, 2, 8, 4><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 4));
    run("mixed R=4 NF2=1 batch8 u4", [&] { k_mixed<4, 1, 8, 4><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 4));
    run("mixed R=4 NF2=1 batch8 u8", [&] { k_mixed<4, 1, 8, 8><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 4));
    run("mixed R=6 NF2=1 batch8 u4", [&] { k_mixed<6, 1, 8, 4><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 6));
    run("mixed R=6 NF2=2 batch8 u4", [&] { k_mixed<6, 2, 8, 4><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 6));
    printf("-- k_2d: TFLOP/s-equiv column counts 18 flop/test for every variant (so 2-D rows show >1x speed, not flops)\n");
    run("3D rows FMNMX3 R=8 b16 u4", [&] { k_2d<8, 0, 1, 16, 4><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 8));
    run("3D rows LOP3   R=8 b16 u4", [&] { k_2d<8, 1, 1, 16, 4><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 8));
    run("2D rows FMNMX3 R=8 b16 u4", [&] { k_2d<8, 0, 0, 16, 4><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 8));
    run("2D rows LOP3   R=8 b16 u4", [&] { k_2d<8, 1, 0, 16, 4><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 8));
    run("2D rows LOP3   R=12 b16 u4", [&] { k_2d<12, 1, 0, 16, 4><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 12));
    run("2D rows LOP3   R=16 b16 u2", [&] { k_2d<16, 1, 0, 16, 2><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 16));
    run("2D rows LOP3   R=8 b16 u8", [&] { k_2d<8, 1, 0, 16, 8><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 8));
    run("2D rows LOP3   R=4 b16 u4", [&] { k_2d<4, 1, 0, 16, 4><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 4));
    run("2D rows FMNMX3 R=16 b16 u2", [&] { k_2d<16, 0, 0, 16, 2><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 16));
    printf("-- R / unroll sweep of the 2-D LOP3 loop (18-flop-equiv)\n");
#define SW(R, U) run("2D LOP3 R=" #R " b16 u" #U, [&] { k_2d<R, 1, 0, 16, U><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, R));
    SW(5, 4) SW(6, 4) SW(7, 4) SW(9, 4) SW(10, 4) SW(6, 2) SW(10, 2) SW(8, 2) SW(8, 1) SW(6, 8) SW(10, 1)
    run("2D LOP3 R=8 b32 u4", [&] { k_2d<8, 1, 0, 32, 4><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 8));
    run("2D LOP3 R=8 b8 u4", [&] { k_2d<8, 1, 0, 8, 4><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 8));
    run("2D LOP3 R=8 b16 u4 256thr", [&] { k_2d<8, 1, 0, 16, 4><<<sms, 256>>>(tile_g, reps, out, seed); }, PAIRS(256, 8));
    run("2D LOP3 R=8 b16 u4 384thr", [&] { k_2d<8, 1, 0, 16, 4><<<sms, 384>>>(tile_g, reps, out, seed); }, PAIRS(384, 8));
    printf("-- k_sat: saturating rows + product accumulate (8 FMA-pipe ops / pair, no LOP3); same 18-flop-equiv column\n");
#define ST(R, N, B, U) run("SAT R=" #R " acc" #N " b" #B " u" #U, [&] { k_sat<R, N, B, U><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, R));
    ST(8, 1, 16, 4) ST(8, 2, 16, 4) ST(8, 4, 16, 4) ST(8, 8, 16, 4) ST(8, 2, 16, 2) ST(8, 2, 32, 4) ST(8, 2, 16, 8) ST(10, 2, 16, 4) ST(12, 2, 16, 2) ST(6, 2, 16, 4) ST(4, 2, 16, 4)
    run("SAT R=8 acc2 b16 u4 256thr", [&] { k_sat<8, 2, 16, 4><<<sms, 256>>>(tile_g, reps, out, seed); }, PAIRS(256, 8));
    run("SAT R=8 acc2 b16 u4 384thr", [&] { k_sat<8, 2, 16, 4><<<sms, 384>>>(tile_g, reps, out, seed); }, PAIRS(384, 8));
    printf("-- shared-q variants (rays of a thread share q): same 18-flop-equiv column, i.e. a pair-rate scale\n");
    run("SQ 2D LOP3 R=8 b16 u4", [&] { k_2d_sq<8, 1, 0, 16, 4><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 8));
    run("SQ 2D LOP3 R=8 b16 u2", [&] { k_2d_sq<8, 1, 0, 16, 2><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 8));
    run("SQ 2D LOP3 R=12 b16 u2", [&] { k_2d_sq<12, 1, 0, 16, 2><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 12));
    run("SQ 2D LOP3 R=16 b16 u2", [&] { k_2d_sq<16, 1, 0, 16, 2><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 16));
    run("SQ 2D FMNMX3 R=8 b16 u4", [&] { k_2d_sq<8, 0, 0, 16, 4><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 8));
    run("SQ SAT R=8 acc2 b16 u4", [&] { k_sat_sq<8, 2, 16, 4><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 8));
    run("SQ SAT R=8 acc4 b16 u4", [&] { k_sat_sq<8, 4, 16, 4><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 8));
    run("SQ SAT R=8 acc4 b16 u2", [&] { k_sat_sq<8, 4, 16, 2><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 8));
    run("SQ SAT R=16 acc4 b16 u2", [&] { k_sat_sq<16, 4, 16, 2><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 16));
    run("SQ SAT R=12 acc4 b16 u2", [&] { k_sat_sq<12, 4, 16, 2><<<sms, 512>>>(tile_g, reps, out, seed); }, PAIRS(512, 12));
    {   // production data path
        const int n_tiles = 400;
        std::vector<float> ht((size_t)n_tiles * TILE * 12, 0.f);
        for (size_t i = 0; i < ht.size(); ++i) ht[i] = ((i % 4) == 2) ? -1.f : 0.0001f * (float)(i % 89); // C = -1: never candidate
        float4 *table;
        int *work;
        cudaMalloc(&table, ht.size() * 4);
        cudaMemcpy(table, ht.data(), ht.size() * 4, cudaMemcpyHostToDevice);
        cudaMalloc(&work, 4);
        auto prod = [&](auto kern, int R, int blocks_per_sm) {
            const size_t smem = R == 8 ? sizeof(sweep::Smem<8>) : R == 4 ? sizeof(sweep::Smem<4>) : sizeof(sweep::Smem<2>);
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            const int n_blocks = sms * blocks_per_sm;
            char name[64];
            snprintf(name, sizeof name, "PROD path R=%d blocks/SM=%d", R, blocks_per_sm);
            run(name, [&] { cudaMemset(work, 0, 4); kern<<<sms, sweep::THREADS, smem>>>(table, n_tiles, n_blocks, work, out, seed); },
                (double)n_blocks * sweep::THREADS * R * TILE * n_tiles);
        };
        prod(k_prod<4>, 4, 1);
        prod(k_prod<8>, 8, 1);
        prod(k_prod<4>, 4, 4);
        prod(k_prod<2>, 2, 2);
    }
    return 0;
}
