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
    float lmax; // longest ray of the bundle (distance from the rays' common point O to the far end); FLT_MAX: unbounded
};

// sign word of the three edge-function maxima over the box: sign bit clear <=> all three >= 0
__device__ __forceinline__ unsigned box_sign(const float4 rb, const float4 rc, const float4 rd, const Box b) {
    const float x = fmaf(rb.x, b.pc, fmaf(fabsf(rb.x), b.ph, fmaf(rb.y, b.qc, fmaf(fabsf(rb.y), b.qh, rb.z))));
    const float y = fmaf(rc.x, b.pc, fmaf(fabsf(rc.x), b.ph, fmaf(rc.y, b.qc, fmaf(fabsf(rc.y), b.qh, rc.z))));
    const float z = fmaf(rd.x, b.pc, fmaf(fabsf(rd.x), b.ph, fmaf(rd.y, b.qc, fmaf(fabsf(rd.y), b.qh, rd.z))));
    // rb.w: lower bound of the distance from O to any point of the triangle (0 in tables without one): a triangle
    // that lies wholly beyond the far end of every ray of the bundle cannot be hit by any of them
    return __float_as_uint(x) | __float_as_uint(y) | __float_as_uint(z) | (rb.w > b.lmax ? 0x80000000u : 0u);
}

// (min,max) ranges -> centre / half extent, slightly enlarged: the box evaluation rounds differently
// from the per-ray evaluation, and must never be the stricter of the two
__device__ __forceinline__ Box make_box(float pmin, float pmax, float qmin, float qmax, float lmax) {
    Box b;
    b.lmax = lmax;
    b.pc = 0.5f * (pmin + pmax), b.qc = 0.5f * (qmin + qmax);
    b.ph = 0.5f * (pmax - pmin) * 1.0001f + 1e-6f * (fabsf(b.pc) + 1.f);
    b.qh = 0.5f * (qmax - qmin) * 1.0001f + 1e-6f * (fabsf(b.qc) + 1.f);
    return b;
}

// box of the R rays held by one thread
template <int R>
__device__ __forceinline__ Box lane_box_of(const float (&rp)[R], const float (&rq)[R], const float (&rl)[R]) {
    float pmin = rp[0], pmax = rp[0], qmin = rq[0], qmax = rq[0], lmax = rl[0];
#pragma unroll
    for (int r = 1; r < R; ++r) {
        pmin = fminf(pmin, rp[r]), pmax = fmaxf(pmax, rp[r]);
        qmin = fminf(qmin, rq[r]), qmax = fmaxf(qmax, rq[r]);
        lmax = fmaxf(lmax, rl[r]);
    }
    return make_box(pmin, pmax, qmin, qmax, lmax);
}

// warp box and CTA box of the rays held by this thread block (rp/rq/rl of invalid rays must be duplicates
// of valid ones).  rl: far-end distance of each ray.  scratch: 5 * THREADS/32 floats of shared memory.
template <int R>
__device__ __forceinline__ void bundle_boxes(const float (&rp)[R], const float (&rq)[R], const float (&rl)[R], float *scratch,
                                             Box &warp_box, Box &cta_box) {
    float pmin = rp[0], pmax = rp[0], qmin = rq[0], qmax = rq[0], lmax = rl[0];
#pragma unroll
    for (int r = 1; r < R; ++r) {
        pmin = fminf(pmin, rp[r]), pmax = fmaxf(pmax, rp[r]);
        qmin = fminf(qmin, rq[r]), qmax = fmaxf(qmax, rq[r]);
        lmax = fmaxf(lmax, rl[r]);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        pmin = fminf(pmin, __shfl_xor_sync(0xffffffffu, pmin, o)), pmax = fmaxf(pmax, __shfl_xor_sync(0xffffffffu, pmax, o));
        qmin = fminf(qmin, __shfl_xor_sync(0xffffffffu, qmin, o)), qmax = fmaxf(qmax, __shfl_xor_sync(0xffffffffu, qmax, o));
        lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    }
    warp_box = make_box(pmin, pmax, qmin, qmax, lmax);
    const int w = threadIdx.x >> 5, nw = sweep::THREADS / 32;
    if ((threadIdx.x & 31) == 0)
        scratch[w] = pmin, scratch[nw + w] = pmax, scratch[2 * nw + w] = qmin, scratch[3 * nw + w] = qmax, scratch[4 * nw + w] = lmax;
    __syncthreads();
    for (int i = 0; i < nw; ++i) {
        pmin = fminf(pmin, scratch[i]), pmax = fmaxf(pmax, scratch[nw + i]);
        qmin = fminf(qmin, scratch[2 * nw + i]), qmax = fmaxf(qmax, scratch[3 * nw + i]);
        lmax = fmaxf(lmax, scratch[4 * nw + i]);
    }
    cta_box = make_box(pmin, pmax, qmin, qmax, lmax);
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
    float scratch[5 * sweep::THREADS / 32];
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

// warp-aggregated append of (ray, triangle) pairs: bit r of `mask` = this lane's ray r is a candidate for `tri`.
// Called by all 32 lanes (mask may be 0 in some of them).
template <int R>
__device__ __forceinline__ void emit_pairs(const Emitter em, WarpChunk &wc, unsigned mask, const int (&ray_id)[R], unsigned tri) {
    const int lane = threadIdx.x & 31;
    const int mine = __popc(mask);
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
    while (mask) {
        const int r = __ffs(mask) - 1;
        mask &= mask - 1;
        if (base < em.cap) em.buf[base] = ((unsigned long long)(unsigned)ray_id[r] << 32) | tri;
        ++base;
    }
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
                for (int r = 0; r < R; ++r) mask |= (unsigned)sweep::edge_pass(rb, rc, rd, rp[r], rq[r]) << r;
                mask &= valid;
                if (__ballot_sync(0xffffffffu, mask != 0) == 0) continue;
                emit_pairs<R>(em, wc, mask, ray_id, (unsigned)((tile_lo + it) * CTILE + k));
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

// ---- two-phase organisation of the same hierarchy ("block lists") ---------------------------------------
// The streaming sweep above spends most of its time on per-tile bookkeeping (mbarrier wait, CTA-wide vote,
// stage refill) although the level-0 test itself is ~15 instructions per triangle.  The two-phase form
// separates the levels into dense kernels:
//   phase A  cull_l0_kernel: EVERY (ray block, triangle) pair — still no acceleration structure, no build
//            step — with one triangle per thread held in registers and the block boxes broadcast from
//            shared memory; survivors are tested against the block's 16 warp boxes at once and appended as
//            (block*16+warp)<<tri_bits|triangle, then radix-sorted;
//   phase B  walk_block_list: each warp of the ray block walks the block's (short) survivor list 32 triangles
//            at a time against its own warp box, then per ray, and emits ray<<32|triangle candidates.
// The candidates then take the same sort + strict path as before, so results stay bit-identical.
struct BlockBoxes {
    Box cta;
    Box warp[sweep::THREADS / 32];
};

struct L0Params {
    const float4 *tables;  // group j -> tables + ((j / nface) * 6 + j % nface) * table_stride; nface == 0: one table
    const float4 *allcand; // group with j % nface == nface - 1
    size_t table_stride;
    int n_groups, nface, n_tris;
    int tri_bits;       // keys are block << tri_bits | triangle: fewer radix passes than a 32-bit split
    const int *blk_off; // [n_groups + 1] ray blocks of each group (device)
    const BlockBoxes *boxes;
    unsigned long long *keys, *count;
    unsigned long long cap;
    sweep::Counters *diag;
};

__device__ __forceinline__ const float4 *group_table(const float4 *tables, const float4 *allcand, size_t stride, int nface, int j) {
    if (nface == 0) return tables;
    const int face = j % nface;
    return face == nface - 1 ? allcand : tables + (size_t)((j / nface) * 6 + face) * stride;
}

constexpr int L0_THREADS = 256;
constexpr int L0_STAGE = 256; // keys staged per warp in shared memory between flushes (one global atomic per flush)
__global__ void __launch_bounds__(L0_THREADS) cull_l0_kernel(const L0Params p) {
    __shared__ Box sbox[L0_THREADS];
    __shared__ unsigned long long stage[L0_THREADS / 32][L0_STAGE];
    const int j = blockIdx.y;
    const int b_lo = p.blk_off[j], b_hi = p.blk_off[j + 1];
    if (b_lo >= b_hi) return;
    const float4 *__restrict__ tab = group_table(p.tables, p.allcand, p.table_stride, p.nface, j);
    const int tri = blockIdx.x * L0_THREADS + threadIdx.x, lane = threadIdx.x & 31;
    unsigned long long *__restrict__ st = stage[threadIdx.x >> 5];
    const bool live = tri < p.n_tris;
    float4 rb = make_float4(0.f, 0.f, -1.f, 0.f), rc = rb, rd = rb;
    if (live) rb = tab[3 * (size_t)tri], rc = tab[3 * (size_t)tri + 1], rd = tab[3 * (size_t)tri + 2];
    unsigned fill = 0, n_pass = 0; // warp-uniform
    auto flush = [&]() {
        __syncwarp();
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(p.count, (unsigned long long)fill);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (unsigned k = lane; k < fill; k += 32)
            if (base + k < p.cap) p.keys[base + k] = st[k];
        __syncwarp();
        n_pass += fill, fill = 0;
    };
    for (int b0 = b_lo; b0 < b_hi; b0 += L0_THREADS) {
        __syncthreads();
        if (b0 + (int)threadIdx.x < b_hi) sbox[threadIdx.x] = p.boxes[b0 + threadIdx.x].cta;
        __syncthreads();
        const int nb = min(L0_THREADS, b_hi - b0);
#pragma unroll 2
        for (int i = 0; i < nb; ++i) {
            const bool pass = live && !(box_sign(rb, rc, rd, sbox[i]) >> 31);
            unsigned m = __ballot_sync(0xffffffffu, pass);
            while (m) { // (block, triangle) pairs that pass the block box, two per step: each half-warp tests one
                        // pair's row against the block's 16 warp boxes
                const int s0 = __ffs(m) - 1;
                m &= m - 1;
                int s1 = -1;
                if (m) s1 = __ffs(m) - 1, m &= m - 1;
                const int wl = lane & 15, mysrc = (lane >> 4) ? s1 : s0, srcl = mysrc < 0 ? 0 : mysrc;
                float4 sb, sc, sd;
                sb.x = __shfl_sync(0xffffffffu, rb.x, srcl), sb.y = __shfl_sync(0xffffffffu, rb.y, srcl);
                sb.z = __shfl_sync(0xffffffffu, rb.z, srcl), sb.w = __shfl_sync(0xffffffffu, rb.w, srcl);
                sc.x = __shfl_sync(0xffffffffu, rc.x, srcl), sc.y = __shfl_sync(0xffffffffu, rc.y, srcl);
                sc.z = __shfl_sync(0xffffffffu, rc.z, srcl), sc.w = 0.f;
                sd.x = __shfl_sync(0xffffffffu, rd.x, srcl), sd.y = __shfl_sync(0xffffffffu, rd.y, srcl);
                sd.z = __shfl_sync(0xffffffffu, rd.z, srcl), sd.w = 0.f;
                static_assert(sweep::THREADS / 32 == 16, "one half-warp per ray block's warp boxes");
                bool wp = false;
                if (mysrc >= 0) wp = !(box_sign(sb, sc, sd, p.boxes[b0 + i].warp[wl]) >> 31);
                const unsigned wm = __ballot_sync(0xffffffffu, wp);
                if (wm) {
                    if (wp)
                        st[fill + __popc(wm & ((1u << lane) - 1u))] =
                            ((unsigned long long)(unsigned)((b0 + i) * 16 + wl) << p.tri_bits) | (unsigned)(tri - lane + srcl);
                    fill += __popc(wm);
                    if (fill > L0_STAGE - 32) flush();
                }
            }
        }
    }
    if (fill) flush();
    if (p.diag && lane == 0 && n_pass) atomicAdd(&p.diag->cull_l0, (unsigned long long)n_pass);
}

// first index i in [0, n) with keys[i] >= key
__device__ __forceinline__ unsigned long long lower_bound_key(const unsigned long long *__restrict__ keys, unsigned long long n,
                                                              unsigned long long key) {
    unsigned long long lo = 0, hi = n;
    while (lo < hi) {
        const unsigned long long mid = (lo + hi) >> 1;
        if (keys[mid] < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// phase B works on the sorted keys one WARP at a time (kernels.cuh: cull2_body): 32 keys are staged per step
// in a per-warp slice of shared memory (each lane gathers the row of one key), then walked in order with the
// row broadcast from shared memory: lane box, per-ray filter, candidates appended through the warp's chunk.
// No block-wide barrier: the survivors are spread very unevenly over the warps of a ray block.
struct WarpListSmem {
    float4 row[sweep::THREADS / 32][32 * 3];
    unsigned long long key[sweep::THREADS / 32][32];
};

}  // namespace cull
