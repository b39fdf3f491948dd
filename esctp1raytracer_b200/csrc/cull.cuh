// cull.cuh — OPTIONAL mode (opts.bundle_cull): hierarchical evaluation of the same conservative filter.
//
// The default sweeps (sweep.cuh) evaluate the three affine edge functions for every (ray, triangle)
// pair: brute force, FP32-issue bound, the formulation BASELINE.json's north star prescribes.  Because
// the edge functions are affine in the ray parameters (p,q), their maximum over an axis-aligned box of
// (p,q) is attained at a corner:  max = A*pc + |A|*ph + B*qc + |B|*qh + C.  If that maximum is negative
// for one of the three edges, NO ray inside the box can pass the filter.  So a bundle of rays that is
// compact in (p,q) — a screen tile of primary rays, a cell of the light's cube face for shadow rays —
// can reject a triangle for all of its rays with 12 FFMA, evaluated with one triangle per LANE instead
// of one triangle per warp.  Every bundle still considers every triangle (there is no acceleration
// structure and no build step), the survivors go through exactly the per-ray filter and the strict
// path of the default mode, in index order, so results are bit-identical to the default mode
// (tests/test_gpu_parity.py::test_bundle_cull_*).  What changes is the bound: the sweep becomes a
// stream of the 48-byte rows through L2/shared memory (TMA), not FP32 issue.
//
// Levels: CTA box (all 512*R rays of the work item) tested by one thread per triangle of the staged
// tile; the few survivors are then tested against each warp's box (warp-uniform), then per ray.
// Shadow rays are first ordered by (light vertex, cube face, Morton code of (p,q)) with a radix sort so
// that consecutive rays form compact bundles.
#pragma once
#include "sweep.cuh"

namespace cull {

constexpr int CTA_WALK = 24; // CTA-box survivors per tile above which each warp culls the tile against its own box instead

struct Box {
    float pc, ph, qc, qh;
};

// sign word of the three edge-function maxima over the box: sign bit clear <=> all three >= 0
__device__ __forceinline__ unsigned box_sign(const float4 rb, const float4 rc, const float4 rd, const Box b) {
    const float x = fmaf(rb.x, b.pc, fmaf(fabsf(rb.x), b.ph, fmaf(rb.y, b.qc, fmaf(fabsf(rb.y), b.qh, rb.z))));
    const float y = fmaf(rc.x, b.pc, fmaf(fabsf(rc.x), b.ph, fmaf(rc.y, b.qc, fmaf(fabsf(rc.y), b.qh, rc.z))));
    const float z = fmaf(rd.x, b.pc, fmaf(fabsf(rd.x), b.ph, fmaf(rd.y, b.qc, fmaf(fabsf(rd.y), b.qh, rd.z))));
    return __float_as_uint(x) | __float_as_uint(y) | __float_as_uint(z);
}

// (min,max) ranges -> centre / half extent, slightly enlarged: the box evaluation rounds differently
// from the per-ray evaluation, and must never be the stricter of the two
__device__ __forceinline__ Box make_box(float pmin, float pmax, float qmin, float qmax) {
    Box b;
    b.pc = 0.5f * (pmin + pmax), b.qc = 0.5f * (qmin + qmax);
    b.ph = 0.5f * (pmax - pmin) * 1.0001f + 1e-6f * (fabsf(b.pc) + 1.f);
    b.qh = 0.5f * (qmax - qmin) * 1.0001f + 1e-6f * (fabsf(b.qc) + 1.f);
    return b;
}

// warp box and CTA box of the rays held by this thread block (rp/rq of invalid rays must be duplicates
// of valid ones).  scratch: 4 * THREADS/32 floats of shared memory.
template <int R>
__device__ __forceinline__ void bundle_boxes(const float (&rp)[R], const float (&rq)[R], float *scratch, Box &warp_box,
                                             Box &cta_box) {
    float pmin = rp[0], pmax = rp[0], qmin = rq[0], qmax = rq[0];
#pragma unroll
    for (int r = 1; r < R; ++r) {
        pmin = fminf(pmin, rp[r]), pmax = fmaxf(pmax, rp[r]);
        qmin = fminf(qmin, rq[r]), qmax = fmaxf(qmax, rq[r]);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        pmin = fminf(pmin, __shfl_xor_sync(0xffffffffu, pmin, o)), pmax = fmaxf(pmax, __shfl_xor_sync(0xffffffffu, pmax, o));
        qmin = fminf(qmin, __shfl_xor_sync(0xffffffffu, qmin, o)), qmax = fmaxf(qmax, __shfl_xor_sync(0xffffffffu, qmax, o));
    }
    warp_box = make_box(pmin, pmax, qmin, qmax);
    const int w = threadIdx.x >> 5, nw = sweep::THREADS / 32;
    if ((threadIdx.x & 31) == 0) scratch[w] = pmin, scratch[nw + w] = pmax, scratch[2 * nw + w] = qmin, scratch[3 * nw + w] = qmax;
    __syncthreads();
    for (int i = 0; i < nw; ++i) {
        pmin = fminf(pmin, scratch[i]), pmax = fmaxf(pmax, scratch[nw + i]);
        qmin = fminf(qmin, scratch[2 * nw + i]), qmax = fmaxf(qmax, scratch[3 * nw + i]);
    }
    cta_box = make_box(pmin, pmax, qmin, qmax);
    __syncthreads();
}

// ---- the culled sweep only EMITS the surviving (ray, triangle) pairs (wavefront organisation) ----------
// In bundle-cull mode a tile of 256 triangles costs a few hundred cycles, so an in-line strict evaluation
// (L2 round trip for the vertices + FP64 divide, one or two lanes active) would dominate and stall the
// whole CTA at the tile barrier.  Instead the sweep appends ray<<32|triangle to a global buffer; the
// buffer is sorted (ray major, triangle minor = the reference's iteration order per ray) and one thread
// per ray then walks its candidates in order with the strict arithmetic (kernels.cuh: strict_*_from_candidates).
constexpr int CTILE = 512;  // triangles per stage in emit mode: one per thread at level 0
constexpr int CSTAGES = 4;
constexpr uint32_t CTILE_BYTES = CTILE * 3 * sizeof(float4);

struct __align__(128) EmitSmem {
    float4 tile[CSTAGES][CTILE * 3];
    uint64_t full_bar[CSTAGES];
    unsigned cmask[CTILE / 32];
    float scratch[4 * sweep::THREADS / 32];
    int blk, seg, slice;
};

// Candidate output.  Each warp owns a private chunk of the global buffer (one atomicAdd per CHUNK entries,
// so no atomic latency inside the tile loop: a warp stalled on an atomic would stall its whole CTA at the next
// tile barrier); the unused tail of a warp's last chunk is filled with the all-ones sentinel, which sorts last.
constexpr unsigned CHUNK = 1024;
constexpr unsigned long long SENTINEL = 0xffffffffffffffffull;
struct Emitter {
    unsigned long long *buf, *count;
    unsigned long long cap;
};
struct WarpChunk { // warp-uniform
    unsigned long long base;
    unsigned used;
};
__device__ __forceinline__ void chunk_close(const Emitter em, const WarpChunk wc) {
    const int lane = threadIdx.x & 31;
    if (wc.used >= CHUNK) return;
    for (unsigned i = wc.used + lane; i < CHUNK; i += 32)
        if (wc.base + i < em.cap) em.buf[wc.base + i] = SENTINEL;
}

// tile_lo/tile_hi in units of CTILE triangles.  One barrier per tile in the common case (no survivor of the
// CTA box in the tile): __syncthreads_or both publishes "any survivor" and proves that every thread is done
// with the previous tile's stage, which thread 0 then refills.
template <int R>
__device__ __forceinline__ void sweep_cull_emit(EmitSmem &sm, const float4 *__restrict__ table, int tile_lo, int tile_hi,
                                                const float (&rp)[R], const float (&rq)[R], unsigned valid,
                                                const int (&ray_id)[R], unsigned &gtile, const Box cta_box, const Box warp_box,
                                                const Emitter em, WarpChunk &wc, sweep::Counters *diag) {
    using namespace sweep;
    const int tid = threadIdx.x, lane = tid & 31;
    const int n_tiles = tile_hi - tile_lo;
    const float4 *__restrict__ src = table + (size_t)tile_lo * CTILE * 3;
    unsigned d_l0 = 0, d_l1 = 0, d_any = 0, d_fb = 0;
    if (tid == 0) {
        for (int i = 0; i < (n_tiles < CSTAGES ? n_tiles : CSTAGES); ++i) {
            const unsigned g = gtile + i;
            mbar_expect_tx(&sm.full_bar[g % CSTAGES], CTILE_BYTES);
            tma_load_1d(sm.tile[g % CSTAGES], src + (size_t)i * CTILE * 3, CTILE_BYTES, &sm.full_bar[g % CSTAGES]);
        }
    }
    for (int it = 0; it < n_tiles; ++it) {
        const unsigned g = gtile + it;
        const int s = g % CSTAGES;
        mbar_wait(&sm.full_bar[s], (g / CSTAGES) & 1u);
        const float4 *__restrict__ tp = sm.tile[s];
        // level 0: one triangle per thread against the CTA box
        const unsigned pass = (box_sign(tp[3 * tid], tp[3 * tid + 1], tp[3 * tid + 2], cta_box) >> 31) ^ 1u;
        const int any = __syncthreads_or((int)pass);
        if (tid == 0 && it >= 1 && it - 1 + CSTAGES < n_tiles) { // the stage of the previous tile is free now
            const unsigned gp = g - 1;
            mbar_expect_tx(&sm.full_bar[gp % CSTAGES], CTILE_BYTES);
            tma_load_1d(sm.tile[gp % CSTAGES], src + (size_t)(it - 1 + CSTAGES) * CTILE * 3, CTILE_BYTES, &sm.full_bar[gp % CSTAGES]);
        }
        if (!any) continue;
        const unsigned m0 = __ballot_sync(0xffffffffu, pass);
        if (lane == 0) sm.cmask[tid >> 5] = m0;
        __syncthreads();
        int n_surv = 0;
#pragma unroll
        for (int w = 0; w < CTILE / 32; ++w) n_surv += __popc(sm.cmask[w]);
        d_l0 += n_surv, ++d_any, d_fb += n_surv > CTA_WALK;
#pragma unroll 1
        for (int w = 0; w < CTILE / 32; ++w) {
            unsigned mm = sm.cmask[w];
            if (!mm) continue;
            if (n_surv > CTA_WALK) { // level 1 lane-parallel: one triangle per lane against this warp's box
                const int k = w * 32 + lane;
                const unsigned p1 = ((mm >> lane) & 1u) & ((box_sign(tp[3 * k], tp[3 * k + 1], tp[3 * k + 2], warp_box) >> 31) ^ 1u);
                mm = __ballot_sync(0xffffffffu, p1);
            }
            while (mm) { // survivors, in index order (warp-uniform loop)
                const int k = w * 32 + __ffs(mm) - 1;
                mm &= mm - 1;
                const float4 rb = tp[3 * k], rc = tp[3 * k + 1], rd = tp[3 * k + 2];
                if (n_surv <= CTA_WALK && (box_sign(rb, rc, rd, warp_box) >> 31)) continue; // level 1, warp-uniform
                unsigned mask = 0;                                                          // level 2: per-ray filter
                ++d_l1;
#pragma unroll
                for (int r = 0; r < R; ++r) mask |= ((edge_sign(rb, rc, rd, rp[r], rq[r]) >> 31) ^ 1u) << r;
                mask &= valid;
                if (__ballot_sync(0xffffffffu, mask != 0) == 0) continue;
                const int mine = __popc(mask); // warp-aggregated append
                int incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int y = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += y;
                }
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                if (wc.used + (unsigned)total > CHUNK) { // rare: take a fresh chunk (the only atomic)
                    chunk_close(em, wc);
                    unsigned long long nb = 0;
                    if (lane == 0) nb = atomicAdd(em.count, (unsigned long long)CHUNK);
                    wc.base = __shfl_sync(0xffffffffu, nb, 0), wc.used = 0;
                }
                unsigned long long base = wc.base + wc.used + (unsigned long long)(incl - mine);
                wc.used += (unsigned)total;
                const unsigned tri = (unsigned)((tile_lo + it) * CTILE + k);
                while (mask) {
                    const int r = __ffs(mask) - 1;
                    mask &= mask - 1;
                    if (base < em.cap) em.buf[base] = ((unsigned long long)(unsigned)ray_id[r] << 32) | tri;
                    ++base;
                }
            }
        }
    }
    gtile += n_tiles;
    if (diag && lane == 0) {
        if (tid == 0) atomicAdd(&diag->cull_l0, (unsigned long long)d_l0), atomicAdd(&diag->cull_tiles_any, (unsigned long long)d_any),
            atomicAdd(&diag->cull_tiles_fallback, (unsigned long long)d_fb);
        atomicAdd(&diag->cull_l1, (unsigned long long)d_l1);
    }
    __syncthreads(); // all stages consumed before the next item's prologue refills them
}

}  // namespace cull
