// strict_host_shim.cpp — test infrastructure: csrc/strict_math.cuh (the device's restatement of the reference's
// arithmetic) compiled for the HOST by plain g++, with the correctly-rounded CUDA intrinsics mapped to the IEEE
// operations they are defined as (-ffp-contract=off: no FMA contraction, SSE scalar: every operation rounds once).
// tests/test_strict_host.py compares it, pair by pair and bit by bit, with the pinned oracle (oracle/restated.c), so an
// edit of strict_math.cuh that changes an operation or its order is caught on the CPU, before any GPU run.
#include <cmath>
#include <cstdint>
#include <cuda_runtime.h> // plain g++: __device__ / __forceinline__ expand to nothing special, float4 etc. are defined

static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return std::sqrt(a); }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline float __double2float_rn(double a) { return (float)a; }

#include "../esctp1raytracer_b200/csrc/strict_math.cuh"

extern "C" {
// n pairs: orig[3n], dir[3n], tri[9n], t_in[n] -> hit[n], t_out[n], v_out[n] (t, v unchanged on a miss; v enters as 0)
void strict_host_intersect_triangles(int n, const float *orig, const float *dir, const float *tri, const float *t_in,
                                     uint8_t *hit, float *t_out, float *v_out) {
    for (int i = 0; i < n; ++i) {
        float t = t_in[i], v = 0.f;
        const float *p = tri + 9 * (size_t)i;
        hit[i] = strict::intersect_triangle(strict::ld(orig + 3 * (size_t)i), strict::ld(dir + 3 * (size_t)i), strict::ld(p),
                                            strict::ld(p + 3), strict::ld(p + 6), t, v);
        t_out[i] = t, v_out[i] = v;
    }
}
// extension (no reference code): cr = centre, radius
void strict_host_intersect_spheres(int n, const float *orig, const float *dir, const float *cr, const float *t_in, uint8_t *hit,
                                   float *t_out) {
    for (int i = 0; i < n; ++i) {
        float t = t_in[i];
        const float *c = cr + 4 * (size_t)i;
        hit[i] = strict::intersect_sphere(strict::ld(orig + 3 * (size_t)i), strict::ld(dir + 3 * (size_t)i), make_float4(c[0], c[1], c[2], c[3]), t);
        t_out[i] = t;
    }
}
// vec.h helpers on n vectors: dot, cross, normalize, length
void strict_host_vec(int n, const float *a, const float *b, float *dot, float *cross, float *norm, float *len) {
    for (int i = 0; i < n; ++i) {
        const strict::f3 x = strict::ld(a + 3 * (size_t)i), y = strict::ld(b + 3 * (size_t)i);
        dot[i] = strict::dot(x, y);
        const strict::f3 c = strict::cross(x, y), nn = strict::normalize(x);
        cross[3 * i] = c.x, cross[3 * i + 1] = c.y, cross[3 * i + 2] = c.z;
        norm[3 * i] = nn.x, norm[3 * i + 1] = nn.y, norm[3 * i + 2] = nn.z;
        len[i] = strict::length(x);
    }
}
}
