// sweep.cuh — the O(rays x triangles) sweeps (closest hit and first-occluder),
// B200-native formulation.
//
// What the reference does per (ray, triangle) pair: full Moller-Trumbore with
// a double-precision divide (ray_triangle.h:7-57, ~52 flop).  What this kernel
// does per pair: at most 6 FFMA + 1.5 LOP3 (3.4-3.75 FFMA when the rays of a
// thread share the q-term of each row, see 3.).  The saving comes from these
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
//     tabulated once per (O, triangle).
//
//  1b. Directions through one point have two degrees of freedom.  Each ray group
//     uses a parametrisation d' = p*U + q*V + W (the image plane for primary
//     rays: U,V,W = horizontal, vertical, llc-origin and (p,q) = the reference's
//     own (s,t); a cube face around a light vertex for shadow rays), so each
//     edge function becomes AFFINE in (p,q):  u'(p,q) = p*(U.B) + q*(V.B) + (W.B + K|d'|max)
//     = 2 FFMA.  The table row is 3 x (A,B,C,-) = 48 bytes per (O, group, triangle).
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
//  3. The R rays of a thread can share the q-term B*q + C of every row.  Closest-hit
//     ray blocks are screen tiles in which a thread holds R consecutive pixels of one
//     image row: same q, the compiler merges the R identical inner FFMAs (kernels.cuh,
//     primary_kernel SHAREDQ).  Shadow-ray lists are sorted by q, a thread takes R
//     consecutive rays and uses qbar*B + |B|*qdelta + C with qdelta >= max|q_r - qbar|,
//     which is >= every ray's own term, so the test stays a necessary condition
//     (edge_sign_qbar; sweep_table QBAR).
//
// Mapping to the SM (choices measured with tools/sweep_mb.cu):
//  * rows stream HBM/L2 -> shared memory in 12 KB tiles via TMA 1-D bulk copies
//    (cp.async.bulk + mbarrier complete_tx, UBLKCP in SASS), STAGES deep;
//  * every lane of every warp reads the SAME row at the same time, so the three
//    LDS.128 per triangle are pure broadcasts, amortised over R rays per thread
//    held in registers (R = 8: 27-48 FFMA per 3 loads);
//  * "all three >= 0" is tested on the sign bits: (u'|v'|w') as integers (one LOP3),
//    AND-ed over the R rays; FMNMX3 measured ~2 issue slots next to FFMA, LOP3 ~1.3;
//  * the inner loop has NO per-triangle branch: the resulting sign bit is shifted
//    into a bit register and tested once per batch of BATCH triangles; candidates
//    (a few per ray per sweep) are then re-evaluated in index order.  This keeps
//    the hot loop a straight FFMA/LOP3/LDS stream;
//  * scalar FFMA, not packed fma.rn.f32x2: with honest operands the packed form
//    measured no faster in this loop on B200;
//  * closest-hit sweeps recycle a tile stage without a block-wide barrier (per-warp
//    arrival count in shared memory, the last warp issues the refill); any-hit sweeps
//    keep __syncthreads_and, which also carries the "every ray occluded" early exit;
//  * 512 threads (16 warps, 4 per scheduler) per CTA, one CTA per SM.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "strict_math.cuh"

namespace sweep {

constexpr int TILE = 256;          // triangles per shared-memory stage (48-byte rows: 12 KB)
constexpr int BATCH = 16;          // triangles between candidate checks
constexpr int STAGES = 4;          // TMA pipeline depth
constexpr int THREADS = 512;       // threads per CTA
constexpr float CK = 64.f;         // safety factor of the filter margins (units of FLT_EPSILON)
constexpr uint32_t TILE_BYTES = TILE * 3 * sizeof(float4);

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
    float4 tile[STAGES][TILE * 3];
    // per-ray state touched only on the (rare) strict path; SoA over threads => conflict-free
    float ox[R][THREADS], oy[R][THREADS], oz[R][THREADS];
    float dx[R][THREADS], dy[R][THREADS], dz[R][THREADS];
    float t[R][THREADS], v[R][THREADS];
    int tri[R][THREADS];
    uint64_t full_bar[STAGES];
    int blk, seg, scan[THREADS / 32], base_out;
    int consumed[STAGES]; // closest-hit sweeps: warps that have finished the tile in this stage
};

// Where the strict path finds a ray.  Closest-hit sweeps compute their rays in the item prologue and keep them
// in the shared-memory slots (src.ro == nullptr).  Any-hit sweeps would have to GATHER origin, direction and
// length of 512*R rays per work item although only a few rays per item ever reach the strict path, so they
// leave them in the pixel state and the strict path fetches them on demand through the pixel index, which is
// parked in the (otherwise unused) tri slot as -2 - k until the ray finds its occluder.
struct RaySrc {
    const float *ro, *rd, *rt; // [3][n], [3][n], [n]
    int n;
};

struct Counters {
    unsigned long long tests_primary, tests_shadow, strict_evals, tests_shadow_ref, n_hits, filter_misses;
    unsigned long long cull_l0, cull_l1, cull_tiles_any, cull_tiles_fallback; // bundle-cull diagnostics
    unsigned long long cull_overflow; // bundle-cull: a candidate buffer was too small (the frame is rejected)
};

// One 48-byte row triple per triangle: rb = (A,B,C,-) of s*u', rc of s*v', rd of s*w'.
// Sign word of the three edge functions at the ray parameters (p,q): sign bit clear <=> all three >= 0.
__device__ __forceinline__ unsigned edge_sign(const float4 rb, const float4 rc, const float4 rd, float p, float q) {
    const float x = fmaf(p, rb.x, fmaf(q, rb.y, rb.z));
    const float y = fmaf(p, rc.x, fmaf(q, rc.y, rc.z));
    const float z = fmaf(p, rd.x, fmaf(q, rd.y, rd.z));
    return __float_as_uint(x) | __float_as_uint(y) | __float_as_uint(z);
}

// QBAR form: the R rays of a thread are evaluated with ONE q-term per edge row.  With qbar the thread's mean q and
// qdelta >= max |q_r - qbar| the row value p*A + (qbar*B + |B|*qdelta + C) is >= the ray's own p*A + q_r*B + C, so
// the test stays a necessary condition while the inner term costs 2 FFMA per thread instead of 1 per ray
// (shadow rays sorted by q: qdelta ~ 1e-6, far inside the margin K already in C).
__device__ __forceinline__ unsigned edge_sign_qbar(const float4 rb, const float4 rc, const float4 rd, float p, float qbar, float qdelta) {
    const float x = fmaf(p, rb.x, fmaf(fabsf(rb.y), qdelta, fmaf(qbar, rb.y, rb.z)));
    const float y = fmaf(p, rc.x, fmaf(fabsf(rc.y), qdelta, fmaf(qbar, rc.y, rc.z)));
    const float z = fmaf(p, rd.x, fmaf(fabsf(rd.y), qdelta, fmaf(qbar, rd.y, rd.z)));
    return __float_as_uint(x) | __float_as_uint(y) | __float_as_uint(z);
}

// ---- strict path: the reference's own test on the surviving pairs ---------------
// mask: rays (bit r) that are candidates for triangle tri.
// CLOSEST: cpp_intersect semantics (main.cpp:176-192) — keep going, lower index wins ties.
// ANYHIT : occlusion() semantics (main.cpp:314-329) — the first accepted face in order ends
//          the ray; it leaves t = t2 behind (the multi-light carry).  Returns newly-done rays.
template <int R, bool ANYHIT, int RS> // RS: rays per thread the slot arrays are laid out for (>= R)
__device__ __noinline__ unsigned strict_tri(Smem<RS> &sm, int tid, unsigned mask, int tri,
                                            const float *__restrict__ tri_verts, const RaySrc src, unsigned &n_strict) {
    unsigned newly = 0;
    const float *p = tri_verts + 9 * (size_t)tri;
    const strict::f3 v0 = strict::mk(__ldg(p), __ldg(p + 1), __ldg(p + 2));
    const strict::f3 v1 = strict::mk(__ldg(p + 3), __ldg(p + 4), __ldg(p + 5));
    const strict::f3 v2 = strict::mk(__ldg(p + 6), __ldg(p + 7), __ldg(p + 8));
    while (mask) {
        const int r = __ffs(mask) - 1;
        mask &= mask - 1;
        strict::f3 o, d;
        float t, v;
        if (ANYHIT && src.ro) { // on demand from the pixel state; the ray has no occluder yet, so t is its initial length
            const int k = -2 - sm.tri[r][tid];
            o = strict::mk(src.ro[k], src.ro[src.n + k], src.ro[2 * (size_t)src.n + k]);
            d = strict::mk(src.rd[k], src.rd[src.n + k], src.rd[2 * (size_t)src.n + k]);
            t = src.rt[k], v = 0.f;
        } else {
            o = strict::mk(sm.ox[r][tid], sm.oy[r][tid], sm.oz[r][tid]);
            d = strict::mk(sm.dx[r][tid], sm.dy[r][tid], sm.dz[r][tid]);
            t = sm.t[r][tid], v = sm.v[r][tid];
        }
        ++n_strict;
        if (strict::intersect_triangle(o, d, v0, v1, v2, t, v)) {
            sm.t[r][tid] = t;
            sm.v[r][tid] = v;
            sm.tri[r][tid] = tri;
            if (ANYHIT) newly |= 1u << r;
        }
    }
    return newly;
}

// ---- the sweep over tiles [tile_lo, tile_hi) of one origin table, for the R rays of each thread ---
// rp/rq: the rays' parameters in the table's direction parametrisation
// valid: bit r set = ray r exists; done: bit r set = ray r needs no more tests
// gtile: running tile counter of this CTA (mbarrier phase bookkeeping across ray blocks)
template <int R, bool ANYHIT, bool EXHAUSTIVE, bool QBAR, int RS>
__device__ __forceinline__ void sweep_table(Smem<RS> &sm, const float4 *__restrict__ table, int tile_lo, int tile_hi,
                                            int n_tris, const float *__restrict__ tri_verts, const RaySrc rsrc, const float (&rp)[R],
                                            const float (&rq)[R], float qbar, float qdelta, unsigned valid, unsigned &done,
                                            unsigned &gtile, unsigned &n_strict, unsigned &n_tiles_swept, unsigned &n_miss) {
    const int tid = threadIdx.x;
    const int n_tiles = tile_hi - tile_lo;
    const float4 *__restrict__ src = table + (size_t)tile_lo * TILE * 3;
    int last_issued = (n_tiles < STAGES ? n_tiles : STAGES) - 1;
    if (tid == 0) {
        for (int i = 0; i <= last_issued; ++i) {
            const unsigned g = gtile + i;
            mbar_expect_tx(&sm.full_bar[g % STAGES], TILE_BYTES);
            tma_load_1d(sm.tile[g % STAGES], src + (size_t)i * TILE * 3, TILE_BYTES, &sm.full_bar[g % STAGES]);
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
#pragma unroll 1
            for (int b0 = 0; b0 < TILE; b0 += BATCH) {
                // hot loop: straight-line, no branch per triangle.  neg collects, per triangle, the AND
                // over rays of sign(u'|v'|w'): 0 = some ray may hit that triangle.
                unsigned neg = 0xffffffffu;
#pragma unroll 4
                for (int k = 0; k < BATCH; ++k) {
                    const float4 rb = tp[3 * (b0 + k)], rc = tp[3 * (b0 + k) + 1], rd = tp[3 * (b0 + k) + 2];
                    unsigned A = 0xffffffffu;
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        A &= QBAR ? edge_sign_qbar(rb, rc, rd, rp[r], qbar, qdelta) : edge_sign(rb, rc, rd, rp[r], rq[r]);
                    neg = __funnelshift_l(A, neg, 1);
                }
#ifdef SWEEP_NO_STRICT // development microbenchmark only (tools/sweep_mb.cu): timing without the strict path
                if (~neg & 0xffffu) ++n_strict;
#else
                unsigned cand = EXHAUSTIVE ? 0xffffu : (~neg & 0xffffu); // bit (BATCH-1-k) = triangle b0+k
                while (cand) {
                    // rare: rebuild the per-ray candidate mask of this triangle, then the reference's own
                    // arithmetic.  The row is re-read through an opaque load so that the compiler cannot
                    // merge this with the hot evaluation above.
                    const int hb = 31 - __clz(cand);
                    cand &= ~(1u << hb);
                    const int k = BATCH - 1 - hb;
                    const int tri = (tile_lo + it) * TILE + b0 + k;
                    const float4 rb = lds128_opaque(&tp[3 * (b0 + k)]), rc = lds128_opaque(&tp[3 * (b0 + k) + 1]),
                                 rd = lds128_opaque(&tp[3 * (b0 + k) + 2]);
                    unsigned mask = 0;
#pragma unroll
                    for (int r = 0; r < R; ++r) mask |= ((edge_sign(rb, rc, rd, rp[r], rq[r]) >> 31) ^ 1u) << r;
                    // (QBAR sweeps hand the rays' exact q in rq as well: the candidate masks are as tight as ever)
                    const unsigned live = valid & ~done;
                    if (EXHAUSTIVE) {
                        // validation mode: strict-test every pair, count accepts the filter would have lost
                        if (tri < n_tris && live) {
                            int before[R];
#pragma unroll
                            for (int r = 0; r < R; ++r) before[r] = sm.tri[r][tid];
                            const unsigned nw = strict_tri<R, ANYHIT, RS>(sm, tid, live, tri, tri_verts, rsrc, n_strict);
                            if (ANYHIT) done |= nw;
#pragma unroll
                            for (int r = 0; r < R; ++r)
                                if (sm.tri[r][tid] != before[r] && !((mask >> r) & 1u)) ++n_miss;
                        }
                    } else {
                        mask &= live;
                        if (mask) {
                            const unsigned nw = strict_tri<R, ANYHIT, RS>(sm, tid, mask, tri, tri_verts, rsrc, n_strict);
                            if (ANYHIT) done |= nw;
                        }
                    }
                }
#endif
            }
        }
        if (!ANYHIT) {
            // Closest hit never stops early, so no block-wide barrier is needed per tile: every warp counts
            // itself out of stage s, and the LAST one refills it.  Warps run up to STAGES-1 tiles apart, which
            // absorbs the skew of the (rare, long) strict evaluations instead of stalling 15 warps behind one.
            __syncwarp();
            if ((tid & 31) == 0) {
                __threadfence_block(); // this warp's reads of the stage are done before the count is visible
                if (atomicAdd(&sm.consumed[s], 1) == THREADS / 32 - 1) {
                    sm.consumed[s] = 0;
                    __threadfence_block();
                    if (it + STAGES < n_tiles) {
                        mbar_expect_tx(&sm.full_bar[s], TILE_BYTES);
                        tma_load_1d(sm.tile[s], src + (size_t)(it + STAGES) * TILE * 3, TILE_BYTES, &sm.full_bar[s]);
                    }
                }
            }
            continue;
        }
        // everyone is done with stage s (also: have all rays of the CTA found their occluder?)
        const int all_done = __syncthreads_and((done | ~valid) == 0xffffffffu);
        if (all_done) stop = true;
        if (!stop && it + STAGES < n_tiles) {
            last_issued = it + STAGES;
            if (tid == 0) {
                mbar_expect_tx(&sm.full_bar[s], TILE_BYTES);
                tma_load_1d(sm.tile[s], src + (size_t)last_issued * TILE * 3, TILE_BYTES, &sm.full_bar[s]);
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
