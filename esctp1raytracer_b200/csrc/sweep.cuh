// sweep.cuh — the O(rays x triangles) sweeps (closest hit and first-occluder),
// B200-native formulation.
//
// What the reference does per (ray, triangle) pair: full Moller-Trumbore with
// a double-precision divide (ray_triangle.h:7-57, ~52 flop).  What this kernel
// does per pair: 9 FMA (issued as 4.5 packed FFMA2).  The saving comes from two
// observations.
//
//  1. Every sweep is a bundle of rays through ONE common point.  Primary rays
//     all start at the eye (camera.h:31-34).  A shadow ray runs from the hit
//     point to the sampled light point, and the reference's light "sample" is
//     always exactly a light vertex (main.cpp:749-754: v0=v1=v2), so all
//     shadow rays toward the same light vertex lie on lines through that
//     vertex.  For a fixed point O and triangle (v0,v1,v2), with a=v0-O,
//     e1=v1-v0, e2=v2-v0, the three Moller-Trumbore numerators are LINEAR in
//     the line direction d:   u' = d.(a x e2)   v' = d.(e1 x a)
//     w' = det-u'-v' = d.(e2 x e1 - a x e2 - e1 x a),  and the side of the
//     plane O lies on, s = sign(e2.(e1 x a)), is a per-triangle constant.
//     A line through O can only hit the triangle if s*u', s*v', s*w' >= 0.
//     The three vectors (pre-multiplied by s) plus a safety margin K are
//     tabulated once per (O, triangle): 48 bytes.
//
//  2. The test above is used as a CONSERVATIVE FILTER only: K bounds every
//     rounding difference between this evaluation and the reference's own
//     float/double evaluation (see build_origin_table), so a pair the
//     reference would accept is never filtered out.  Pairs that pass (a few
//     per ray out of N) are re-evaluated in the reference's exact arithmetic
//     (strict_math.cuh) in index order, so closest-hit ties, the `first
//     occluder in order` rule and all the chaotic self-shadow decisions come
//     out bit-identical to the serial path.
//
// Mapping to the SM:
//  * rows stream HBM/L2 -> shared memory in 12 KB tiles via TMA 1-D bulk copies
//    (cp.async.bulk + mbarrier complete_tx, UBLKCP in SASS), STAGES deep;
//  * two consecutive triangles are interleaved in one 96-byte record so that a
//    packed fma.rn.f32x2 (FFMA2) evaluates the same edge function of BOTH
//    triangles for one ray: half the issue slots of scalar FFMA, which is what
//    bounds this loop (the FP32 pipe itself is then the limiter);
//  * every lane of every warp reads the SAME record at the same time, so the six
//    LDS.128 per triangle pair are pure broadcasts, amortised over R rays per
//    thread held in registers as (d,d) pairs;
//  * 512 threads (16 warps, 4 per scheduler) per CTA, one CTA per SM.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "strict_math.cuh"

namespace sweep {

constexpr int TILE = 256;          // triangles per shared-memory stage
constexpr int PAIRS = TILE / 2;    // 96-byte records per stage (12 KB)
constexpr int STAGES = 4;          // TMA pipeline depth
constexpr int THREADS = 512;       // threads per CTA
constexpr float CK = 64.f;         // safety factor of the filter margins (units of FLT_EPSILON)
constexpr uint32_t TILE_BYTES = PAIRS * 6 * sizeof(float4);

// ---- mbarrier / TMA bulk-copy primitives (sm_90+; sm_100a here) -------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk copy global -> shared, completion signalled on an mbarrier
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ float4 lds128_opaque(const float4 *p) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
    return v;
}

// ---- shared memory of one sweep CTA -------------------------------------------
template <int R>
struct __align__(128) Smem {
    float4 tile[STAGES][PAIRS * 6];
    // per-ray state touched only on the (rare) strict path; SoA over threads => conflict-free
    float ox[R][THREADS], oy[R][THREADS], oz[R][THREADS];
    float dx[R][THREADS], dy[R][THREADS], dz[R][THREADS];
    float t[R][THREADS], v[R][THREADS];
    int tri[R][THREADS];
    uint64_t full_bar[STAGES];
    int blk, seg, scan[THREADS / 32], base_out;
};

struct Counters {
    unsigned long long tests_primary, tests_shadow, strict_evals, tests_shadow_ref, n_hits, filter_misses;
};

// One 96-byte record = two triangles t0,t1 interleaved:
//   q[0] = (Bx0,Bx1,By0,By1)  q[1] = (Bz0,Bz1,K0,K1)   u'-row
//   q[2], q[3] likewise for the v'-row (C), q[4], q[5] for the w'-row (D)
// Returns min over the three edge functions, per triangle: m.x for t0, m.y for t1.
__device__ __forceinline__ float2 edge_min(const float4 (&q)[6], float2 ex, float2 ey, float2 ez) {
    const float2 X = __ffma2_rn(ex, make_float2(q[0].x, q[0].y),
                                __ffma2_rn(ey, make_float2(q[0].z, q[0].w),
                                           __ffma2_rn(ez, make_float2(q[1].x, q[1].y), make_float2(q[1].z, q[1].w))));
    const float2 Y = __ffma2_rn(ex, make_float2(q[2].x, q[2].y),
                                __ffma2_rn(ey, make_float2(q[2].z, q[2].w),
                                           __ffma2_rn(ez, make_float2(q[3].x, q[3].y), make_float2(q[3].z, q[3].w))));
    const float2 Z = __ffma2_rn(ex, make_float2(q[4].x, q[4].y),
                                __ffma2_rn(ey, make_float2(q[4].z, q[4].w),
                                           __ffma2_rn(ez, make_float2(q[5].x, q[5].y), make_float2(q[5].z, q[5].w))));
    return make_float2(fminf(fminf(X.x, Y.x), Z.x), fminf(fminf(X.y, Y.y), Z.y));
}

// ---- strict path: the reference's own test on the surviving pairs ---------------
// mask0/mask1: rays (bit r) that are candidates for triangle tri0 / tri0+1.
// CLOSEST: cpp_intersect semantics (main.cpp:176-192) — keep going, lower index wins ties.
// ANYHIT : occlusion() semantics (main.cpp:314-329) — the first accepted face in order ends
//          the ray; it leaves t = t2 behind (the multi-light carry).  Returns newly-done rays.
template <int R, bool ANYHIT>
__device__ __noinline__ unsigned strict_pair(Smem<R> &sm, int tid, unsigned mask0, unsigned mask1, int tri0,
                                             const float *__restrict__ tri_verts, unsigned &n_strict) {
    unsigned newly = 0;
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
        unsigned mask = (which ? mask1 : mask0) & ~newly;
        if (!mask) continue;
        const int tri = tri0 + which;
        const float *p = tri_verts + 9 * (size_t)tri;
        const strict::f3 v0 = strict::mk(__ldg(p), __ldg(p + 1), __ldg(p + 2));
        const strict::f3 v1 = strict::mk(__ldg(p + 3), __ldg(p + 4), __ldg(p + 5));
        const strict::f3 v2 = strict::mk(__ldg(p + 6), __ldg(p + 7), __ldg(p + 8));
        while (mask) {
            const int r = __ffs(mask) - 1;
            mask &= mask - 1;
            const strict::f3 o = strict::mk(sm.ox[r][tid], sm.oy[r][tid], sm.oz[r][tid]);
            const strict::f3 d = strict::mk(sm.dx[r][tid], sm.dy[r][tid], sm.dz[r][tid]);
            float t = sm.t[r][tid], v = sm.v[r][tid];
            ++n_strict;
            if (strict::intersect_triangle(o, d, v0, v1, v2, t, v)) {
                sm.t[r][tid] = t;
                sm.v[r][tid] = v;
                sm.tri[r][tid] = tri;
                if (ANYHIT) newly |= 1u << r;
            }
        }
    }
    return newly;
}

// ---- the sweep over tiles [tile_lo, tile_hi) of one origin table, for the R rays of each thread ---
// e2x/e2y/e2z: filter directions as (d,d) pairs (unit-ish vectors along the line through the table's origin)
// valid: bit r set = ray r exists; done: bit r set = ray r needs no more tests
// gtile: running tile counter of this CTA (mbarrier phase bookkeeping across ray blocks)
template <int R, bool ANYHIT, bool EXHAUSTIVE>
__device__ __forceinline__ void sweep_table(Smem<R> &sm, const float4 *__restrict__ table, int tile_lo, int tile_hi,
                                            int n_tris, const float *__restrict__ tri_verts, const float2 (&e2x)[R],
                                            const float2 (&e2y)[R], const float2 (&e2z)[R], unsigned valid,
                                            unsigned &done, unsigned &gtile, unsigned &n_strict,
                                            unsigned &n_tiles_swept, unsigned &n_miss) {
    const int tid = threadIdx.x;
    const int n_tiles = tile_hi - tile_lo;
    const float4 *__restrict__ src = table + (size_t)tile_lo * PAIRS * 6;
    int last_issued = (n_tiles < STAGES ? n_tiles : STAGES) - 1;
    if (tid == 0) {
        for (int i = 0; i <= last_issued; ++i) {
            const unsigned g = gtile + i;
            mbar_expect_tx(&sm.full_bar[g % STAGES], TILE_BYTES);
            tma_load_1d(sm.tile[g % STAGES], src + (size_t)i * PAIRS * 6, TILE_BYTES, &sm.full_bar[g % STAGES]);
        }
    }
    bool stop = false;
    int it = 0;
    for (; it < n_tiles; ++it) {
        const unsigned g = gtile + it;
        const int s = g % STAGES;
        mbar_wait(&sm.full_bar[s], (g / STAGES) & 1u);
        if (!stop) {
            ++n_tiles_swept;
            const float4 *__restrict__ tp = sm.tile[s];
#pragma unroll 2
            for (int pi = 0; pi < PAIRS; ++pi) {
                float4 q[6];
#pragma unroll
                for (int j = 0; j < 6; ++j) q[j] = tp[6 * pi + j];
                float M = -1.f;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const float2 m = edge_min(q, e2x[r], e2y[r], e2z[r]);
                    M = fmaxf(fmaxf(M, m.x), m.y);
                }
                if (EXHAUSTIVE || M >= 0.f) {
                    // rare: rebuild the per-ray candidate masks, then the reference's own arithmetic.
                    // The record is re-read through an opaque load so that the compiler cannot merge
                    // this with the hot evaluation above and keep 2R extra values live / compute the
                    // masks unconditionally.
                    float4 qq[6];
#pragma unroll
                    for (int j = 0; j < 6; ++j) qq[j] = lds128_opaque(&tp[6 * pi + j]);
                    unsigned mask0 = 0, mask1 = 0;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const float2 m = edge_min(qq, e2x[r], e2y[r], e2z[r]);
                        mask0 |= (m.x >= 0.f ? 1u : 0u) << r;
                        mask1 |= (m.y >= 0.f ? 1u : 0u) << r;
                    }
                    const int tri0 = (tile_lo + it) * TILE + 2 * pi;
                    const unsigned live = valid & ~done;
                    if (EXHAUSTIVE) {
                        // validation mode: strict-test every pair, count accepts the filter would have lost
                        const unsigned l0 = tri0 < n_tris ? live : 0u, l1 = tri0 + 1 < n_tris ? live : 0u;
                        if (l0 | l1) {
                            int before[R];
#pragma unroll
                            for (int r = 0; r < R; ++r) before[r] = sm.tri[r][tid];
                            const unsigned nw = strict_pair<R, ANYHIT>(sm, tid, l0, l1, tri0, tri_verts, n_strict);
                            if (ANYHIT) done |= nw;
#pragma unroll
                            for (int r = 0; r < R; ++r) {
                                const int now = sm.tri[r][tid];
                                if (now != before[r] && !(((now == tri0 ? mask0 : mask1) >> r) & 1u)) ++n_miss;
                            }
                        }
                    } else {
                        mask0 &= live, mask1 &= live;
                        if (mask0 | mask1) {
                            const unsigned nw = strict_pair<R, ANYHIT>(sm, tid, mask0, mask1, tri0, tri_verts, n_strict);
                            if (ANYHIT) done |= nw;
                        }
                    }
                }
            }
        }
        // everyone is done with stage s (also: have all rays of the CTA found their occluder?)
        const int all_done = __syncthreads_and((done | ~valid) == 0xffffffffu);
        if (ANYHIT && all_done) stop = true;
        if (!stop && it + STAGES < n_tiles) {
            last_issued = it + STAGES;
            if (tid == 0) {
                mbar_expect_tx(&sm.full_bar[s], TILE_BYTES);
                tma_load_1d(sm.tile[s], src + (size_t)last_issued * PAIRS * 6, TILE_BYTES, &sm.full_bar[s]);
            }
        }
        if (stop && it >= last_issued) {
            ++it;
            break;
        }
    }
    gtile += it; // every issued tile has been waited for
}

}  // namespace sweep
