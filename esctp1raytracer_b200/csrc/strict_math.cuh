// strict_math.cuh — the reference's arithmetic, bit for bit, on the device.
//
// The reference serial path is built for baseline x86-64: every float
// operation rounds once (SSE scalar, no FMA contraction).  Its image is
// numerically chaotic (shadow rays start ON the surface, src/main.cpp:757-758),
// so every decision and every value that reaches the framebuffer is evaluated
// here with correctly-rounded intrinsics, which nvcc never contracts:
//   a*b -> __fmul_rn   a+b -> __fadd_rn   a-b -> __fsub_rn
//   a/b -> __fdiv_rn   sqrtf -> __fsqrt_rn
//   double det / inv_det of ray_triangle.h:21-26 -> real FP64 (__ddiv_rn, __dmul_rn)
// The O(N) sweeps run a cheap conservative filter instead (sweep.cuh) and
// only the surviving (ray, triangle) pairs come through these functions.
#pragma once
#include <cfloat>
#include <cuda_runtime.h>

namespace strict {

#define TRC_EPS 1.1920928955078125e-07f /* std::numeric_limits<float>::epsilon() */

struct f3 {
    float x, y, z;
};

__device__ __forceinline__ f3 mk(float x, float y, float z) {
    f3 r;
    r.x = x, r.y = y, r.z = z;
    return r;
}
__device__ __forceinline__ f3 ld(const float *p) { return mk(p[0], p[1], p[2]); }

// vec.h:95-101 — sum starts at 0 and accumulates left to right
__device__ __forceinline__ float dot(f3 a, f3 b) {
    float s = __fadd_rn(0.f, __fmul_rn(a.x, b.x));
    s = __fadd_rn(s, __fmul_rn(a.y, b.y));
    s = __fadd_rn(s, __fmul_rn(a.z, b.z));
    return s;
}
// vec.h:103-109
__device__ __forceinline__ f3 cross(f3 a, f3 b) {
    return mk(__fsub_rn(__fmul_rn(a.y, b.z), __fmul_rn(a.z, b.y)), __fsub_rn(__fmul_rn(a.z, b.x), __fmul_rn(a.x, b.z)),
              __fsub_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}
__device__ __forceinline__ f3 add(f3 a, f3 b) { return mk(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z)); }
__device__ __forceinline__ f3 sub(f3 a, f3 b) { return mk(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)); }
__device__ __forceinline__ f3 mul(f3 a, float s) { return mk(__fmul_rn(a.x, s), __fmul_rn(a.y, s), __fmul_rn(a.z, s)); }
__device__ __forceinline__ f3 div(f3 a, float s) { return mk(__fdiv_rn(a.x, s), __fdiv_rn(a.y, s), __fdiv_rn(a.z, s)); }
// vec.h:135-139
__device__ __forceinline__ float length(f3 a) { return __fsqrt_rn(dot(a, a)); }
__device__ __forceinline__ f3 normalize(f3 a) { return div(a, __fsqrt_rn(dot(a, a))); }

// tracer::intersect_triangle, src/scene/ray_triangle.h:7-57.
// On accept: t <- t2, v <- v2 (the caller's u aliases v, src/main.cpp:307/310,
// so only v survives) and returns true.
__device__ __forceinline__ bool intersect_triangle(f3 orig, f3 dir, f3 vert0, f3 vert1, f3 vert2, float &t, float &v) {
    const f3 edge1 = sub(vert1, vert0);
    const f3 edge2 = sub(vert2, vert0);
    const f3 pvec = cross(dir, edge2);
    const double det = (double)dot(edge1, pvec);
    if (det > -(double)TRC_EPS && det < (double)TRC_EPS) return false;
    const double inv_det = __ddiv_rn(1.0, det);
    const f3 tvec = sub(orig, vert0);
    const float u2 = __double2float_rn(__dmul_rn((double)dot(tvec, pvec), inv_det));
    if (u2 < TRC_EPS || u2 > 1.0f) return false;
    const f3 qvec = cross(tvec, edge1);
    const float v2 = __double2float_rn(__dmul_rn((double)dot(dir, qvec), inv_det));
    if (v2 < TRC_EPS || __fadd_rn(u2, v2) > 1.0f) return false;
    const float t2 = __double2float_rn(__dmul_rn((double)dot(edge2, qvec), inv_det));
    if (t2 < TRC_EPS) return false;
    if (t2 >= t) return false;
    t = t2;
    v = v2;
    return true;
}

// Extension with no reference code (src/intersect.h is empty): analytic
// ray-sphere for unit-length dir; same epsilon rules as the triangle test.
// Mirrors oracle/restated.c:sphere_test op for op.
__device__ __forceinline__ bool intersect_sphere(f3 orig, f3 dir, float4 cr, float &t) {
    const f3 oc = sub(orig, mk(cr.x, cr.y, cr.z));
    const float b = dot(oc, dir);
    const float c = __fsub_rn(dot(oc, oc), __fmul_rn(cr.w, cr.w));
    const float disc = __fsub_rn(__fmul_rn(b, b), c);
    if (!(disc >= 0.f)) return false;
    const float sq = __fsqrt_rn(disc);
    float t2 = __fsub_rn(-b, sq);
    if (t2 < TRC_EPS) t2 = __fadd_rn(-b, sq);
    if (t2 < TRC_EPS) return false;
    if (t2 >= t) return false;
    t = t2;
    return true;
}

}  // namespace strict
