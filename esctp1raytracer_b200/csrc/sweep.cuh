// sweep.cuh — the O(rays x triangles) sweeps (closest hit and first-occluder),
// B200-native formulation.
//
// What the reference does per (ray, triangle) pair: full Moller-Trumbore with
// a double-precision divide (ray_triangle.h:7-57, ~52 flop).  What this kernel
// does per pair: two saturating adds and half a packed multiply-add (4.), plus
// 4-8 FFMA per thread and triangle.  The saving comes from these observations.
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
//
//  1b. Directions through one point have two degrees of freedom.  Each ray group
//     uses a parametrisation d' = p*U + q*V + W (the image plane for primary
//     rays: U,V,W = horizontal, vertical, llc-origin and (p,q) = the reference's
//     own (s,t); a cube face around a light vertex for shadow rays), so each
//     edge function becomes AFFINE in (p,q):  u'(p,q) = p*(U.B) + q*(V.B) + (W.B + K|d'|max).
//
//  2. The test above is used as a CONSERVATIVE FILTER only: K bounds every
//     rounding difference between this evaluation and the reference's own
//     float/double evaluation (see build_origin_table), so a pair the
//     reference would accept is never filtered out.  Pairs that pass (a few
//     per ray out of N) are re-evaluated in the reference's exact arithmetic
//     (strict_math.cuh), so closest-hit ties, the `first occluder in order`
//     rule and all the chaotic self-shadow decisions come out bit-identical
//     to the serial path.  The strict results are MERGED IN GLOBAL MEMORY with
//     one 64-bit atomicMin per accepted pair:
//        closest hit  key = t bits << 32 | index   (cpp_intersect keeps the lexicographic
//                                                   minimum of (t, index): t2 >= t rejects)
//        occluder     key = index << 32  | t2 bits (occlusion() returns the lowest index)
//     so the sweep keeps NO per-ray state in shared memory: a CTA's shared
//     memory is the table-tile pipeline only.
//
//  3. The R rays of a thread share q.  Closest-hit ray blocks are screen tiles in which a
//     thread holds R consecutive pixels of one image row: same q exactly (MODE_SHAREDQ).
//     Shadow-ray lists are sorted by q, a thread takes R consecutive rays and evaluates the
//     q-dependent terms at qbar with |B|*qdelta added, qdelta >= max|q_r - qbar|, which is
//     >= every ray's own term, so the test stays a necessary condition (MODE_QBAR; also used
//     for jittered primary samples, whose q lie inside one stratum).
//
//  4. SPAN form.  For rays that share q, the three affine rows are half-lines in p: with
//     beta = -B/A, gamma = -C/A (divided once per (origin, triangle) in FP64) row i says
//     p >= q*beta_i + gamma_i where A_i > 0 and p <= q*beta_i + gamma_i where A_i < 0.  A
//     triangle has two lower bounds and one upper bound or the reverse, so the table row is
//     two lower + two upper bounds = 32 bytes, the thread evaluates the four at its q and keeps
//     the binding ones (4 FFMA + 2 FMNMX), and each PAIR is a two-sided span test (below).
//     MODE_OWNQ (every ray its own q) keeps the 48-byte three-row table: 3 x (A,B,C,-).
//
// Mapping to the SM (choices measured with tools/sweep_mb2-4.cu):
//  * rows stream HBM/L2 -> shared memory in tiles of 256 triangles via TMA 1-D bulk copies
//    (cp.async.bulk + mbarrier complete_tx, UBLKCP in SASS), STAGES deep;
//  * every lane of every warp reads the SAME row at the same time, so the two LDS.128 per
//    triangle are pure broadcasts, amortised over R rays per thread held in registers
//    (R = 32 closest hit, 16 any-hit);
//  * the conjunction is decided IN THE FMA PIPE: bounds carry the +1 of a saturating test, so
//    x = sat(S*p + ax) and y = sat(ay - S*p) are both exactly 1 for a pair the reference could
//    accept, and acc += x*y runs on ray pairs as one packed FFMA2: 2 FFMA.SAT-slot instructions
//    + half an FFMA2 per pair, no integer op.  History on this part (cycles per pair, closest hit):
//    round 1 sign bits through LOP3 on the half-rate ALU pipe 7.9, three saturating rows with a
//    product accumulate 6.8, span form 4.0;
//  * the inner loop has NO per-triangle branch: an accumulator group of 4 triangles is
//    tested once; candidates (a few per ray per sweep) are then re-evaluated one by one;
//  * NO block-wide barrier per tile in either sweep: every warp counts itself out of a
//    tile stage (shared-memory atomicAdd) and the last one issues the refill, so warps
//    run up to STAGES-1 tiles apart.  An any-hit warp whose rays all have their occluder
//    stops evaluating (it only keeps the stage counts going);
//  * SWEEP_NT threads per CTA, SWEEP_MINB CTAs per SM (launch bounds cap the registers).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "strict_math.cuh"

#ifndef SWEEP_NT
#define SWEEP_NT 128
#endif
#ifndef SWEEP_MINB
#define SWEEP_MINB 4
#endif
#ifndef SWEEP_STAGES
#define SWEEP_STAGES 4
#endif

namespace sweep {

constexpr int TILE = 256;          // triangles per shared-memory stage (48-byte rows: 12 KB)
constexpr int BATCH = 16;          // triangles between candidate checks
constexpr int STAGES = SWEEP_STAGES; // TMA pipeline depth
constexpr int THREADS = 512;       // threads per CTA of the bundle-cull kernels (cull.cuh)
constexpr int NT = SWEEP_NT;       // threads per CTA of the default sweeps
constexpr int MINB = SWEEP_MINB;   // CTAs per SM of the default sweeps
constexpr float CK = 64.f;         // safety factor of the filter margins (units of FLT_EPSILON)
constexpr uint32_t TILE_BYTES = TILE * 3 * sizeof(float4); // 48-byte rows (MODE_OWNQ, bundle-cull mode, tools/)
constexpr float SPAN_S = 65536.f;  // span rows: row units per unit of p (a power of two: p * SPAN_S is exact)
constexpr float SPAN_OPEN = 1e30f; // span rows: an unused bound

// MODE_OWNQ sweeps the 48-byte three-row table (every ray its own q: jittered primary rays); MODE_SHAREDQ and MODE_QBAR
// sweep the 32-byte SPAN table (4. below)
enum { MODE_OWNQ = 0, MODE_SHAREDQ = 1, MODE_QBAR = 2 };
template <int MODE>
struct Rows {
    static constexpr int N = MODE == MODE_OWNQ ? 3 : 2;                // float4 per triangle
    static constexpr uint32_t BYTES = TILE * N * sizeof(float4);      // per staged tile
};

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

// ---- shared memory of one sweep CTA: the table-tile pipeline and nothing else -------------------
template <int ROWS> // float4 per triangle of the table the kernel sweeps: 2 (span rows) or 3 (three-row table)
struct __align__(128) SmemT {
    float4 tile[STAGES][TILE * ROWS];
    uint64_t full_bar[STAGES];
    int consumed[STAGES]; // warps that have finished the tile in this stage
    int blk, seg, slice;  // the work item the CTA is on
};
using Smem = SmemT<3>;

struct Counters {
    unsigned long long tests_primary, tests_shadow, strict_evals, tests_shadow_ref, n_hits, filter_misses;
    unsigned long long cull_l0, cull_l1, cull_tiles_any, cull_tiles_fallback; // bundle-cull diagnostics
    unsigned long long cull_overflow; // bundle-cull: a candidate buffer was too small (the frame is rejected)
    unsigned long long pipeline_errors; // validation mode: staged tiles that differed from their source (must be 0)
    // shadow_light_kernel time split, SM cycles summed over CTAs (thread 0's clock; TRACER_SHADOW_DIAG prints them):
    // inside work items, waiting at grid barriers, compaction passes, everything else (work-queue pops, prefix sums)
    unsigned long long cyc_items, cyc_barrier, cyc_compact, cyc_total, n_items, n_runs;
};

template <int ROWS>
__device__ __forceinline__ void smem_init(SmemT<ROWS> &sm) {
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&sm.full_bar[s], 1), sm.consumed[s] = 0;
        fence_barrier_init();
    }
    __syncthreads();
}

// One 48-byte row triple per triangle: rb = (A,B,C,-) of s*u', rc of s*v', rd of s*w', in SATURATING FORM
// (kernels.cuh, build_origin_table): a pair the reference could accept evaluates to >= 1 on all three rows.
// Per-ray test used by the (rare) candidate path: bit clear = some row < 1.
__host__ __device__ __forceinline__ bool edge_pass(const float4 rb, const float4 rc, const float4 rd, float p, float q) {
    const float x = fmaf(p, rb.x, fmaf(q, rb.y, rb.z));
    const float y = fmaf(p, rc.x, fmaf(q, rc.y, rc.z));
    const float z = fmaf(p, rd.x, fmaf(q, rd.y, rd.z));
    return fminf(fminf(x, y), z) >= 1.f;
}
// Sign word of the three edge functions at the ray parameters (p,q): sign bit clear <=> all three >= 0 (bundle-cull
// mode and the LOP3 form of the hot loop; only looser than the saturating test on the same rows).
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
__device__ __forceinline__ float qterm_qbar(const float4 row, float qbar, float qdelta) {
    return fmaf(fabsf(row.y), qdelta, fmaf(qbar, row.y, row.z));
}

// ---- three-row form (MODE_OWNQ; also tools/sweep_mb2,3.cu): BATCH triangles of a staged tile against the R rays of this thread ---
// Everything runs in the FMA pipe (measured on B200, tools/sweep_mb3.cu: a scalar FMA-pipe instruction costs ~1.3
// issue cycles, a LOP3/SHF/FMNMX on the half-rate ALU pipe ~2, and the two do not overlap in this loop):
//   x' = sat(p*A_u + q_u)   y' = sat(p*A_v + q_v)   z' = sat(p*A_w + q_w)      3 FFMA.SAT per pair
//   acc += (x' * y') * z'                                                        1 FMUL + 1 FFMA per pair (packed: below)
// A candidate pair contributes exactly 1, any other pair something in [0,1), so after GROUP triangles
// "sum of the accumulators >= 1" is a necessary condition for the group to hold a candidate.  Straight-line FFMA /
// FMUL / LDS.128 stream, no branch per triangle.  Returns bit g set = triangles [g*GROUP, (g+1)*GROUP) may hold one.
// The conjunction runs on ray PAIRS in packed form (FMUL2 + FFMA2, fma.rn.f32x2): on this part a packed instruction
// costs ~2.2 issue cycles for two FMAs against ~2.6 for two scalar ones; the saturating row evaluations stay scalar
// (fma.sat has no f32x2 form).  Measured (tools/sweep_mb3.cu, shared q): scalar conjunction 5.05 Tpairs/s at R = 8,
// packed 5.41 at R = 8 and 5.71 at R = 12.
template <int R>
struct Batch {
    static_assert(R % 2 == 0, "rays are evaluated in pairs");
    static constexpr int GROUP = R >= 12 ? 4 : 8;                   // triangles per accumulator group
    static constexpr int NACC = R >= 16 ? 4 : R >= 12 ? 3 : (R >= 4 ? 2 : 1); // independent packed accumulators (FFMA2 chains) per group
    static_assert(R <= 32, "ray masks are 32 bits wide");
    static constexpr int NGROUPS = BATCH / GROUP;
};

template <int R, int MODE>
__device__ __forceinline__ unsigned eval_batch(const float4 *__restrict__ tp, const float (&rp)[R], const float (&rq)[R], float qbar,
                                               float qdelta) {
    constexpr int GROUP = Batch<R>::GROUP, NA = Batch<R>::NACC;
    unsigned cand = 0;
#pragma unroll
    for (int g = 0; g < BATCH / GROUP; ++g) {
        float2 acc[NA];
#pragma unroll
        for (int a = 0; a < NA; ++a) acc[a] = make_float2(0.f, 0.f);
#pragma unroll
        for (int kk = 0; kk < GROUP; ++kk) {
            const int k = g * GROUP + kk;
            const float4 rb = tp[3 * k], rc = tp[3 * k + 1], rd = tp[3 * k + 2];
            if (MODE == MODE_OWNQ) {
#pragma unroll
                for (int r = 0; r < R; r += 2) {
                    const float2 x = make_float2(__saturatef(fmaf(rp[r], rb.x, fmaf(rq[r], rb.y, rb.z))),
                                                 __saturatef(fmaf(rp[r + 1], rb.x, fmaf(rq[r + 1], rb.y, rb.z))));
                    const float2 y = make_float2(__saturatef(fmaf(rp[r], rc.x, fmaf(rq[r], rc.y, rc.z))),
                                                 __saturatef(fmaf(rp[r + 1], rc.x, fmaf(rq[r + 1], rc.y, rc.z))));
                    const float2 z = make_float2(__saturatef(fmaf(rp[r], rd.x, fmaf(rq[r], rd.y, rd.z))),
                                                 __saturatef(fmaf(rp[r + 1], rd.x, fmaf(rq[r + 1], rd.y, rd.z))));
                    acc[(r / 2) % NA] = __ffma2_rn(__fmul2_rn(x, y), z, acc[(r / 2) % NA]);
                }
            } else {
                const float qx = MODE == MODE_QBAR ? qterm_qbar(rb, qbar, qdelta) : fmaf(rq[0], rb.y, rb.z);
                const float qy = MODE == MODE_QBAR ? qterm_qbar(rc, qbar, qdelta) : fmaf(rq[0], rc.y, rc.z);
                const float qz = MODE == MODE_QBAR ? qterm_qbar(rd, qbar, qdelta) : fmaf(rq[0], rd.y, rd.z);
#pragma unroll
                for (int r = 0; r < R; r += 2) {
                    const float2 x = make_float2(__saturatef(fmaf(rp[r], rb.x, qx)), __saturatef(fmaf(rp[r + 1], rb.x, qx)));
                    const float2 y = make_float2(__saturatef(fmaf(rp[r], rc.x, qy)), __saturatef(fmaf(rp[r + 1], rc.x, qy)));
                    const float2 z = make_float2(__saturatef(fmaf(rp[r], rd.x, qz)), __saturatef(fmaf(rp[r + 1], rd.x, qz)));
                    acc[(r / 2) % NA] = __ffma2_rn(__fmul2_rn(x, y), z, acc[(r / 2) % NA]);
                }
            }
        }
        float sum = 0.f;
#pragma unroll
        for (int a = 0; a < NA; ++a) sum += acc[a].x + acc[a].y;
        if (sum >= 1.f) cand |= 1u << g;
    }
    return cand;
}

// ---- SPAN form -------------------------------------------------------------------------------------------------------
// For rays that share q the three affine rows p*A_i + (q*B_i + C_i) >= 0 are half-lines in p: p >= c_i(q) where A_i > 0,
// p <= c_i(q) where A_i < 0, with c_i(q) = q*beta_i + gamma_i, beta = -B/A, gamma = -C/A (divided once per (origin,
// triangle) in FP64, kernels.cuh: build_origin_table).  A triangle in front of the parametrisation plane has two lower
// bounds and one upper bound or the reverse, so the 32-byte span row holds two of each, already in row units and with the
// +1 of the saturating test folded in:
//     lo = (Bl1, Cl1, Bl2, Cl2):  a_i = q*Bl_i + Cl_i = 1 - S*(c_i(q) - margin)      ax = min(a_1, a_2)
//     hi = (Bu1, Cu1, Bu2, Cu2):  b_i = q*Bu_i + Cu_i = 1 + S*(c_i(q) + margin)      ay = min(b_1, b_2)
// and a pair is a candidate iff  x = sat(S*p + ax)  and  y = sat(ay - S*p)  are both exactly 1.  Per thread and triangle:
// 4 FFMA + 2 FMNMX (8 FFMA in MODE_QBAR, which adds |B|*qdelta to every bound); per PAIR: 2 FADD.SAT + half a packed
// FFMA2 (acc += x*y), against 3 FFMA.SAT + FMUL2/2 + FFMA2/2 of the three-row form.  Measured (tools/sweep_mb4.cu, B200):
// 8.9 Tpairs/s at 16 rays per thread, 9.3 at 24, 9.5 at 32, against 5.7 for the three-row loop at 12.
__host__ __device__ __forceinline__ void span_terms(const float4 lo, const float4 hi, float q, float &ax, float &ay) {
    ax = fminf(fmaf(q, lo.x, lo.y), fmaf(q, lo.z, lo.w));
    ay = fminf(fmaf(q, hi.x, hi.y), fmaf(q, hi.z, hi.w));
}
__host__ __device__ __forceinline__ void span_terms_qbar(const float4 lo, const float4 hi, float qbar, float qdelta, float &ax, float &ay) {
    ax = fminf(fmaf(fabsf(lo.x), qdelta, fmaf(qbar, lo.x, lo.y)), fmaf(fabsf(lo.z), qdelta, fmaf(qbar, lo.z, lo.w)));
    ay = fminf(fmaf(fabsf(hi.x), qdelta, fmaf(qbar, hi.x, hi.y)), fmaf(fabsf(hi.z), qdelta, fmaf(qbar, hi.z, hi.w)));
}
// per-ray test of the (rare) candidate path: ps = p * SPAN_S, the ray's own q
__host__ __device__ __forceinline__ bool span_pass(const float4 lo, const float4 hi, float ps, float q) {
    float ax, ay;
    span_terms(lo, hi, q, ax, ay);
    return fminf(ps + ax, ay - ps) >= 1.f;
}

// BATCH triangles of a staged span tile against the R rays of this thread (ps[r] = p_r * SPAN_S).  A candidate pair
// contributes exactly 1 to its accumulator, any other pair something in [0,1).  Returns bit g set = triangles
// [g*GROUP, (g+1)*GROUP) may hold a candidate.
template <int R, int MODE>
__device__ __forceinline__ unsigned eval_span(const float4 *__restrict__ tp, const float (&ps)[R], float q, float qdelta) {
    constexpr int GROUP = Batch<R>::GROUP, NA = Batch<R>::NACC;
    unsigned cand = 0;
#pragma unroll
    for (int g = 0; g < BATCH / GROUP; ++g) {
        float2 acc[NA];
#pragma unroll
        for (int a = 0; a < NA; ++a) acc[a] = make_float2(0.f, 0.f);
#pragma unroll
        for (int kk = 0; kk < GROUP; ++kk) {
            const int k = g * GROUP + kk;
            float ax, ay;
            if (MODE == MODE_QBAR) span_terms_qbar(tp[2 * k], tp[2 * k + 1], q, qdelta, ax, ay);
            else span_terms(tp[2 * k], tp[2 * k + 1], q, ax, ay);
#pragma unroll
            for (int r = 0; r < R; r += 2) {
                const float2 x = make_float2(__saturatef(ps[r] + ax), __saturatef(ps[r + 1] + ax));
                const float2 y = make_float2(__saturatef(ay - ps[r]), __saturatef(ay - ps[r + 1]));
                acc[(r / 2) % NA] = __ffma2_rn(x, y, acc[(r / 2) % NA]);
            }
        }
        float sum = 0.f;
#pragma unroll
        for (int a = 0; a < NA; ++a) sum += acc[a].x + acc[a].y;
        if (sum >= 1.f) cand |= 1u << g;
    }
    return cand;
}

// The LOP3 form of the same loop (round 1; kept for tools/sweep_mb3.cu): sign bits OR-ed per ray, AND-ed over rays,
// shifted into a bit register.  Returns bit (BATCH-1-k) CLEAR = some ray of this thread may hit triangle k.
template <int R, int MODE, int UNROLL>
__device__ __forceinline__ unsigned eval_batch_lop3(const float4 *__restrict__ tp, const float (&rp)[R], const float (&rq)[R], float qbar,
                                                    float qdelta) {
    unsigned neg = 0xffffffffu;
#pragma unroll UNROLL
    for (int k = 0; k < BATCH; ++k) {
        const float4 rb = tp[3 * k], rc = tp[3 * k + 1], rd = tp[3 * k + 2];
        unsigned A = 0xffffffffu;
        if (MODE == MODE_OWNQ) {
#pragma unroll
            for (int r = 0; r < R; ++r) A &= edge_sign(rb, rc, rd, rp[r], rq[r]);
        } else {
            const float qx = MODE == MODE_QBAR ? qterm_qbar(rb, qbar, qdelta) : fmaf(rq[0], rb.y, rb.z);
            const float qy = MODE == MODE_QBAR ? qterm_qbar(rc, qbar, qdelta) : fmaf(rq[0], rc.y, rc.z);
            const float qz = MODE == MODE_QBAR ? qterm_qbar(rd, qbar, qdelta) : fmaf(rq[0], rd.y, rd.z);
#pragma unroll
            for (int r = 0; r < R; ++r)
                A &= __float_as_uint(fmaf(rp[r], rb.x, qx)) | __float_as_uint(fmaf(rp[r], rc.x, qy)) |
                     __float_as_uint(fmaf(rp[r], rd.x, qz));
        }
        neg = __funnelshift_l(A, neg, 1);
    }
    return neg;
}

// ---- the sweep over tiles [tile_lo, tile_hi) of one origin table, for the R rays of each thread ---
// rp/rq: the rays' parameters in the table's direction parametrisation (rq is read by MODE_OWNQ and by the candidate path
//        of the any-hit sweeps; MODE_SHAREDQ takes the common q from rq[0], MODE_QBAR from qbar/qdelta)
// valid: bit r set = ray r exists; done: bit r set = ray r needs no more tests
// gtile: running tile counter of this CTA (mbarrier phase bookkeeping across work items)
// strict(mask, tri, filt) -> newly done rays: the reference's own test on the surviving pairs (kernels.cuh);
//        mask = rays to test, filt = rays the filter passed (differs from mask only in EXHAUSTIVE validation mode)
template <int R, int MODE, bool ANYHIT, bool EXHAUSTIVE, class Strict>
__device__ __forceinline__ void sweep_table(SmemT<Rows<MODE>::N> &sm, const float4 *__restrict__ table, int tile_lo, int tile_hi, int n_tris,
                                            const float (&rp)[R], const float (&rq)[R], float qbar, float qdelta, unsigned valid,
                                            unsigned &done, unsigned &gtile, unsigned &n_tiles_swept, Strict &&strict,
                                            unsigned *n_pipe_err = nullptr) {
    constexpr int ROWS = Rows<MODE>::N;
    constexpr uint32_t BYTES = Rows<MODE>::BYTES;
    constexpr bool SPAN = MODE != MODE_OWNQ;
    const int tid = threadIdx.x;
    const int n_tiles = tile_hi - tile_lo;
    const float4 *__restrict__ src = table + (size_t)tile_lo * TILE * ROWS;
    if (tid == 0) {
        const int first = n_tiles < STAGES ? n_tiles : STAGES;
        for (int i = 0; i < first; ++i) {
            const unsigned g = gtile + i;
            mbar_expect_tx(&sm.full_bar[g % STAGES], BYTES);
            tma_load_1d(sm.tile[g % STAGES], src + (size_t)i * TILE * ROWS, BYTES, &sm.full_bar[g % STAGES]);
        }
    }
    float ps[R]; // span form: the rays' p in row units (exact: SPAN_S is a power of two)
#pragma unroll
    for (int r = 0; r < R; ++r) ps[r] = rp[r] * SPAN_S;
    const float qhot = MODE == MODE_QBAR ? qbar : rq[0];
    for (int it = 0; it < n_tiles; ++it) {
        const unsigned g = gtile + it;
        const int s = g % STAGES;
        mbar_wait(&sm.full_bar[s], (g / STAGES) & 1u);
        // Validation mode doubles as the pipeline's own race check (compute-sanitizer is not available on this pool):
        // every warp compares the staged tile with its source in global memory when the tile arrives AND after it has
        // used it — a copy that had not landed, or a refill issued while a warp was still reading, shows up here.
        auto tile_differs = [&]() {
            unsigned bad = 0;
            const uint4 *a = reinterpret_cast<const uint4 *>(sm.tile[s]);
            const uint4 *b = reinterpret_cast<const uint4 *>(src + (size_t)it * TILE * ROWS);
            for (int i = tid & 31; i < TILE * ROWS; i += 32) {
                const uint4 x = a[i], y = __ldcg(&b[i]);
                bad |= (x.x ^ y.x) | (x.y ^ y.y) | (x.z ^ y.z) | (x.w ^ y.w);
            }
            return __any_sync(0xffffffffu, bad != 0);
        };
        if (EXHAUSTIVE && n_pipe_err && tile_differs() && (tid & 31) == 0) ++*n_pipe_err;
        // any-hit: a warp whose rays all have their occluder has nothing left to evaluate
        const bool idle = ANYHIT && __all_sync(0xffffffffu, (done | ~valid) == 0xffffffffu);
        if (!idle) {
            ++n_tiles_swept;
            const float4 *__restrict__ tp = sm.tile[s];
#pragma unroll 1
            for (int b0 = 0; b0 < TILE; b0 += BATCH) {
                unsigned cand;
                if (SPAN) cand = eval_span<R, MODE>(tp + ROWS * b0, ps, qhot, qdelta);
                else cand = eval_batch<R, MODE>(tp + ROWS * b0, rp, rq, qbar, qdelta);
#ifdef SWEEP_NO_STRICT // development microbenchmark only (tools/sweep_mb2.cu): timing without the strict path
                if (cand) ++done;
#else
                constexpr int GROUP = Batch<R>::GROUP;
                // validation mode: every group goes to the strict path; the hot loop's own verdict is kept so that an
                // accepted pair whose group it would have skipped counts as a filter miss too
                const unsigned cand_hot = cand;
                if (EXHAUSTIVE) cand = (1u << (BATCH / GROUP)) - 1u;
                while (cand) {
                    // rare: a group of GROUP triangles may hold a candidate.  Rebuild the per-ray candidate mask of each
                    // of its triangles (each ray's own q, exact threshold), then the reference's own arithmetic.  The rows
                    // are re-read through opaque loads so that the compiler cannot merge this with the hot evaluation.
                    const int gq = __ffs(cand) - 1; // ascending triangle order
                    cand &= cand - 1;
#pragma unroll 1
                    for (int kk = 0; kk < GROUP; ++kk) {
                        const int k = b0 + gq * GROUP + kk;
                        const int tri = (tile_lo + it) * TILE + k;
                        unsigned filt = 0;
                        if (SPAN) {
                            const float4 lo = lds128_opaque(&tp[2 * k]), hi = lds128_opaque(&tp[2 * k + 1]);
                            if (MODE == MODE_QBAR && ANYHIT) {
                                // shadow rays: each ray's own q (the widening of the shared-q bounds, ~10 % of a pixel over
                                // 16 consecutive rays of a q-sorted list, would cost ~5 % more strict evaluations)
#pragma unroll
                                for (int r = 0; r < R; ++r) filt |= (unsigned)span_pass(lo, hi, ps[r], rq[r]) << r;
                            } else {
                                // closest hit: the hot loop's own test, ray by ray — exact with a common q; with jittered samples
                                // (mean q + |B| * qdelta, 1/8 pixel) it saves keeping 32 q values in registers during the sweep
                                // for ~2 % more strict evaluations
                                float ax, ay;
                                if (MODE == MODE_QBAR) span_terms_qbar(lo, hi, qhot, qdelta, ax, ay);
                                else span_terms(lo, hi, qhot, ax, ay);
#pragma unroll
                                for (int r = 0; r < R; ++r) filt |= (unsigned)(fminf(ps[r] + ax, ay - ps[r]) >= 1.f) << r;
                            }
                        } else {
                            const float4 rb = lds128_opaque(&tp[3 * k]), rc = lds128_opaque(&tp[3 * k + 1]), rd = lds128_opaque(&tp[3 * k + 2]);
#pragma unroll
                            for (int r = 0; r < R; ++r) filt |= (unsigned)edge_pass(rb, rc, rd, rp[r], rq[r]) << r;
                        }
                        if (EXHAUSTIVE && !((cand_hot >> gq) & 1u)) filt = 0;
                        const unsigned live = valid & ~done;
                        const unsigned mask = EXHAUSTIVE ? live : (filt & live);
                        if (mask && tri < n_tris) {
                            const unsigned nw = strict(mask, tri, filt);
                            if (ANYHIT) done |= nw;
                        }
                    }
                }
#endif
            }
        }
        if (EXHAUSTIVE && n_pipe_err && tile_differs() && (tid & 31) == 0) ++*n_pipe_err;
        // No block-wide barrier per tile: every warp counts itself out of stage s, and the LAST one refills it.
        // Warps run up to STAGES-1 tiles apart, which absorbs the skew of the (rare, long) strict evaluations
        // instead of stalling every warp behind one.
        __syncwarp();
        if ((tid & 31) == 0) {
            __threadfence_block(); // this warp's reads of the stage are done before the count is visible
            if (atomicAdd(&sm.consumed[s], 1) == (int)(blockDim.x >> 5) - 1) {
                sm.consumed[s] = 0;
                __threadfence_block();
                if (it + STAGES < n_tiles) {
                    mbar_expect_tx(&sm.full_bar[s], BYTES);
                    tma_load_1d(sm.tile[s], src + (size_t)(it + STAGES) * TILE * ROWS, BYTES, &sm.full_bar[s]);
                }
            }
        }
    }
    gtile += n_tiles; // every issued tile has been waited for
}

}  // namespace sweep
