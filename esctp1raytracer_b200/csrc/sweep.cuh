// sweep.cuh — the O(rays x triangles) sweeps (closest hit and first-occluder),
// B200-native formulation.
//
// What the reference does per (ray, triangle) pair: full Moller-Trumbore with
// a double-precision divide (ray_triangle.h:7-57, ~52 flop).  What this kernel
// does per pair: 9 FFMA.  The saving comes from two observations.
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
//     The three vectors (pre-multiplied by s) are tabulated once per (O,
//     triangle): 48 bytes = three float4 rows whose .w holds a safety margin K.
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
// Data movement: the 48-byte rows stream HBM/L2 -> shared memory in tiles via
// TMA 1-D bulk copies (cp.async.bulk + mbarrier complete_tx), STAGES deep;
// every lane of every warp tests the SAME triangle at the same time, so the
// three LDS.128 per triangle are pure broadcasts, amortised over R rays per
// thread held in registers.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "strict_math.cuh"

namespace sweep {

constexpr int TILE = 256;    // triangles per shared-memory stage (12 KB)
constexpr int STAGES = 4;    // TMA pipeline depth
constexpr int THREADS = 256; // threads per CTA
constexpr float CK = 64.f;   // safety factor of the filter margins (units of FLT_EPSILON)

// ---- mbarrier / TMA bulk-copy primitives (sm_90+; sm_100a here) -------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
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
// 1-D bulk copy global -> shared, completion signalled on an mbarrier (UBLKCP in SASS)
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- shared memory of one sweep CTA -------------------------------------------
template <int R>
struct __align__(128) Smem {
    float4 tile[STAGES][TILE * 3];
    // per-ray state touched only on the (rare) strict path; SoA over threads => conflict-free
    float ox[R][THREADS], oy[R][THREADS], oz[R][THREADS];
    float dx[R][THREADS], dy[R][THREADS], dz[R][THREADS];
    float t[R][THREADS], v[R][THREADS];
    int tri[R][THREADS];
    uint64_t full_bar[STAGES];
    int blk;
    int seg;
};

struct Counters {
    unsigned long long tests_primary, tests_shadow, strict_evals, tests_shadow_ref, n_hits, filter_misses;
};

// ---- strict path, closest hit: cpp_intersect semantics (main.cpp:176-192) ------
template <int R>
__device__ __noinline__ void strict_closest(Smem<R> &sm, int tid, unsigned mask, int tri, const float *__restrict__ tri_verts,
                                            unsigned &n_strict) {
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
        }
    }
}

// ---- strict path, any hit: occlusion() semantics (main.cpp:314-329) -------------
// returns the mask of rays that found their first in-order occluder here
template <int R>
__device__ __noinline__ unsigned strict_anyhit(Smem<R> &sm, int tid, unsigned mask, int tri,
                                               const float *__restrict__ tri_verts, unsigned &n_strict) {
    const float *p = tri_verts + 9 * (size_t)tri;
    const strict::f3 v0 = strict::mk(__ldg(p), __ldg(p + 1), __ldg(p + 2));
    const strict::f3 v1 = strict::mk(__ldg(p + 3), __ldg(p + 4), __ldg(p + 5));
    const strict::f3 v2 = strict::mk(__ldg(p + 6), __ldg(p + 7), __ldg(p + 8));
    unsigned newly = 0;
    while (mask) {
        const int r = __ffs(mask) - 1;
        mask &= mask - 1;
        const strict::f3 o = strict::mk(sm.ox[r][tid], sm.oy[r][tid], sm.oz[r][tid]);
        const strict::f3 d = strict::mk(sm.dx[r][tid], sm.dy[r][tid], sm.dz[r][tid]);
        float t = sm.t[r][tid], v;
        ++n_strict;
        if (strict::intersect_triangle(o, d, v0, v1, v2, t, v)) {
            sm.t[r][tid] = t; // occlusion() leaves t = t2 behind (main.cpp:320-324): the multi-light carry
            sm.tri[r][tid] = tri;
            newly |= 1u << r;
        }
    }
    return newly;
}

// ---- the sweep over all tiles of one origin table, for the R rays of each thread ---
// ex/ey/ez: filter directions (unit-ish vectors along the line through the table's origin)
// valid:    bit r set = ray r exists; done: bit r set = ray r needs no more tests
// gtile:    running tile counter of this CTA (mbarrier phase bookkeeping across ray blocks)
template <int R, bool ANYHIT, bool EXHAUSTIVE>
__device__ __forceinline__ void sweep_table(Smem<R> &sm, const float4 *__restrict__ table, int n_tiles, int n_tris,
                                            const float *__restrict__ tri_verts, const float (&ex)[R],
                                            const float (&ey)[R], const float (&ez)[R], unsigned valid, unsigned &done,
                                            unsigned &gtile, unsigned &n_strict, unsigned &n_tiles_swept,
                                            unsigned &n_miss) {
    const int tid = threadIdx.x;
    constexpr uint32_t TILE_BYTES = TILE * 3 * sizeof(float4);
    int last_issued = (n_tiles < STAGES ? n_tiles : STAGES) - 1;
    if (tid == 0) {
        for (int i = 0; i <= last_issued; ++i) {
            const unsigned g = gtile + i;
            mbar_expect_tx(&sm.full_bar[g % STAGES], TILE_BYTES);
            tma_load_1d(sm.tile[g % STAGES], table + (size_t)i * TILE * 3, TILE_BYTES, &sm.full_bar[g % STAGES]);
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
            for (int i = 0; i < TILE; ++i) {
                const float4 rb = tp[3 * i], rc = tp[3 * i + 1], rd = tp[3 * i + 2];
                float mr[R];
                float M = -1.f;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const float x = fmaf(ex[r], rb.x, fmaf(ey[r], rb.y, fmaf(ez[r], rb.z, rb.w)));
                    const float y = fmaf(ex[r], rc.x, fmaf(ey[r], rc.y, fmaf(ez[r], rc.z, rc.w)));
                    const float z = fmaf(ex[r], rd.x, fmaf(ey[r], rd.y, fmaf(ez[r], rd.z, rd.w)));
                    mr[r] = fminf(fminf(x, y), z);
                    M = fmaxf(M, mr[r]);
                }
                if (EXHAUSTIVE || M >= 0.f) {
                    unsigned mask = 0;
#pragma unroll
                    for (int r = 0; r < R; ++r) mask |= (mr[r] >= 0.f ? 1u : 0u) << r;
                    const int tri = it * TILE + i;
                    if (EXHAUSTIVE) {
                        // validation mode: strict-test every pair, count pairs the filter would have lost
                        const unsigned todo = tri < n_tris ? (valid & ~done) : 0u; // skip padding rows
                        if (todo) {
                            if (ANYHIT) {
                                const unsigned nw = strict_anyhit<R>(sm, tid, todo, tri, tri_verts, n_strict);
                                n_miss += __popc(nw & ~mask);
                                done |= nw;
                            } else {
                                int before[R];
#pragma unroll
                                for (int r = 0; r < R; ++r) before[r] = sm.tri[r][tid];
                                strict_closest<R>(sm, tid, todo, tri, tri_verts, n_strict);
#pragma unroll
                                for (int r = 0; r < R; ++r)
                                    if (sm.tri[r][tid] != before[r] && !((mask >> r) & 1u)) ++n_miss;
                            }
                        }
                    } else {
                        mask &= valid & ~done;
                        if (mask) {
                            if (ANYHIT)
                                done |= strict_anyhit<R>(sm, tid, mask, tri, tri_verts, n_strict);
                            else
                                strict_closest<R>(sm, tid, mask, tri, tri_verts, n_strict);
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
                tma_load_1d(sm.tile[s], table + (size_t)last_issued * TILE * 3, TILE_BYTES, &sm.full_bar[s]);
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
