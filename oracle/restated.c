/* ORACLE / TEST INFRASTRUCTURE ONLY — see restated.h for the rules and the
 * parity status (PINNED against oracle/_ref and tests/golden).
 *
 * Plain-C restatement of the reference's serial per-pixel path, in the exact
 * operation order of SURVEY.md Appendix A.  Build with -ffp-contract=off and
 * no -ffast-math: every float op below must round exactly once, like the
 * reference built for baseline x86-64 (SSE scalar, no FMA).
 */
#include "restated.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define EPS FLT_EPSILON /* std::numeric_limits<float>::epsilon(), ray_triangle.h:23 */

typedef struct {
    float x, y, z;
} v3;

/* ---- src/math/vec.h ------------------------------------------------------ */
static inline v3 V(float x, float y, float z) {
    v3 r = {x, y, z};
    return r;
}
static inline v3 ld3(const float *p) { return V(p[0], p[1], p[2]); }
/* vec.h:95-101  sum starts at 0 and accumulates left to right */
static inline float dot3(v3 a, v3 b) {
    float s = 0;
    s += a.x * b.x;
    s += a.y * b.y;
    s += a.z * b.z;
    return s;
}
/* vec.h:103-109 */
static inline v3 cross3(v3 a, v3 b) { return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline v3 add3(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }  /* vec.h:111-113 */
static inline v3 sub3(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }  /* vec.h:115-117 */
static inline v3 div3(v3 a, float s) { return V(a.x / s, a.y / s, a.z / s); }     /* vec.h:119-125 */
static inline v3 mul3(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }     /* vec.h:127-133 */
static inline v3 norm3(v3 a) { return div3(a, sqrtf(dot3(a, a))); }               /* vec.h:135-137 */
static inline float len3(v3 a) { return sqrtf(dot3(a, a)); }                      /* vec.h:139 */

float rst_dot(const float a[3], const float b[3]) { return dot3(ld3(a), ld3(b)); }
void rst_cross(const float a[3], const float b[3], float out[3]) {
    v3 r = cross3(ld3(a), ld3(b));
    out[0] = r.x;
    out[1] = r.y;
    out[2] = r.z;
}

/* ---- src/scene/camera.h:16-29 -------------------------------------------- */
void rst_camera(const float eye[3], const float look[3], const float vup_[3], float vfov, float aspect, float out[12]) {
    float theta = (float)(vfov * M_PI / 180); /* float * double / int, narrowed */
    float half_height = (float)tan(theta / 2); /* = the value the optimised reference build folds camera.h:20 to (see csrc/tracer_host.cpp) */
    float half_width = aspect * half_height;
    v3 origin = ld3(eye);
    v3 w = norm3(sub3(ld3(eye), ld3(look)));
    v3 u = norm3(cross3(ld3(vup_), w));
    v3 v = cross3(w, u);
    v3 llc = sub3(sub3(sub3(origin, mul3(u, half_width)), mul3(v, half_height)), w);
    v3 hor = mul3(mul3(u, 2.f), half_width);
    v3 ver = mul3(mul3(v, 2.f), half_height);
    out[0] = origin.x, out[1] = origin.y, out[2] = origin.z;
    out[3] = llc.x, out[4] = llc.y, out[5] = llc.z;
    out[6] = hor.x, out[7] = hor.y, out[8] = hor.z;
    out[9] = ver.x, out[10] = ver.y, out[11] = ver.z;
}

/* ---- src/scene/ray_triangle.h:7-57 --------------------------------------- */
static inline int tri_test(v3 orig, v3 dir, v3 vert0, v3 vert1, v3 vert2, float *t, float *u, float *v) {
    v3 edge1 = sub3(vert1, vert0);
    v3 edge2 = sub3(vert2, vert0);
    v3 pvec = cross3(dir, edge2);
    double det = dot3(edge1, pvec);
    if (det > -EPS && det < EPS) return 0;
    double inv_det = 1.0f / det;
    v3 tvec = sub3(orig, vert0);
    float u2 = (float)(dot3(tvec, pvec) * inv_det);
    if (u2 < EPS || u2 > 1.0f) return 0;
    v3 qvec = cross3(tvec, edge1);
    float v2 = (float)(dot3(dir, qvec) * inv_det);
    if (v2 < EPS || u2 + v2 > 1.0f) return 0;
    float t2 = (float)(dot3(edge2, qvec) * inv_det);
    if (t2 < EPS) return 0;
    if (t2 >= *t) return 0;
    *t = t2;
    *u = u2;
    *v = v2;
    return 1;
}

int rst_intersect_triangle(const float orig[3], const float dir[3], const float v0[3], const float v1[3],
                           const float v2[3], float *t, float *u, float *v) {
    return tri_test(ld3(orig), ld3(dir), ld3(v0), ld3(v1), ld3(v2), t, u, v);
}

/* ---- extension, parity unpinned: analytic ray-sphere (dir is unit length) -- */
static inline int sphere_test(v3 orig, v3 dir, const float *cr, float *t) {
    v3 oc = sub3(orig, ld3(cr));
    float b = dot3(oc, dir);
    float c = dot3(oc, oc) - cr[3] * cr[3];
    float disc = b * b - c;
    if (!(disc >= 0.f)) return 0;
    float sq = sqrtf(disc);
    float t2 = -b - sq;
    if (t2 < EPS) t2 = -b + sq;
    if (t2 < EPS) return 0;
    if (t2 >= *t) return 0;
    *t = t2;
    return 1;
}

static inline int n_tris_of(const rst_scene *sc) { return sc->geom_tri_offset[sc->n_geoms]; }

/* ---- closest hit: cpp_intersect, main.cpp:176-192, called as (t, v, v) ---- */
static int closest_hit_scalar(const rst_scene *sc, v3 o, v3 d, float *t, float *v, int64_t *tests) {
    int best = -1;
    const int n = n_tris_of(sc);
    float alias_uv = *v; /* u and v alias the caller's v (main.cpp:307/310) */
    for (int i = 0; i < n; ++i) {
        const float *p = sc->tri_verts + 9 * (size_t)i;
        if (tri_test(o, d, ld3(p), ld3(p + 3), ld3(p + 6), t, &alias_uv, &alias_uv)) best = i;
    }
    *v = alias_uv;
    *tests += n;
    for (int s = 0; s < sc->n_spheres; ++s)
        if (sphere_test(o, d, sc->sphere_cr + 4 * s, t)) best = n + s;
    return best;
}

/* ---- any hit: occlusion, main.cpp:314-329 (first accepted face in order) -- */
static int first_occluder_scalar(const rst_scene *sc, v3 o, v3 d, float *t, int64_t *tests) {
    const int n = n_tris_of(sc);
    float u, v;
    for (int i = 0; i < n; ++i) {
        const float *p = sc->tri_verts + 9 * (size_t)i;
        if (tri_test(o, d, ld3(p), ld3(p + 3), ld3(p + 6), t, &u, &v)) {
            *tests += i + 1;
            return i;
        }
    }
    *tests += n;
    for (int s = 0; s < sc->n_spheres; ++s)
        if (sphere_test(o, d, sc->sphere_cr + 4 * s, t)) return n + s;
    return -1;
}


/* ---- the same tests 8 triangles at a time (AVX2) -------------------------
 * A CPU comparator for boxes without ispc: the reference's per-lane arithmetic, in the same
 * order and precision (float products and sums, double det / inv_det / scaling, ordered
 * compares so NaNs take the same branches), applied to 8 consecutive triangles per step;
 * lanes that pass are then resolved one by one in index order, so closest-hit ties and the
 * first-occluder rule come out exactly as in the scalar loops above.  Pinned bit-for-bit
 * against them (tests/test_oracle_pin.py).  Enabled per process with rst_set_simd(1). */
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define RST_AVX2 __attribute__((target("avx2")))
typedef struct {
    int n, n_pad;
    float *a[9]; /* v0.xyz, e1.xyz, e2.xyz, each n_pad floats */
} soa_tris;
static int g_simd = 0;
static soa_tris g_soa; /* of the scene being rendered (one render at a time when SIMD is on) */

int rst_set_simd(int on) {
    g_simd = on && __builtin_cpu_supports("avx2");
    return g_simd;
}

static void soa_build(const rst_scene *sc) {
    const int n = n_tris_of(sc), n_pad = (n + 7) & ~7;
    g_soa.n = n, g_soa.n_pad = n_pad;
    for (int c = 0; c < 9; ++c) g_soa.a[c] = (float *)aligned_alloc(32, sizeof(float) * (size_t)(n_pad ? n_pad : 8));
    for (int i = 0; i < n_pad; ++i) {
        v3 v0 = V(0, 0, 0), e1 = V(0, 0, 0), e2 = V(0, 0, 0); /* padding: det == 0, rejected */
        if (i < n) {
            const float *p = sc->tri_verts + 9 * (size_t)i;
            v0 = ld3(p), e1 = sub3(ld3(p + 3), v0), e2 = sub3(ld3(p + 6), v0);
        }
        g_soa.a[0][i] = v0.x, g_soa.a[1][i] = v0.y, g_soa.a[2][i] = v0.z;
        g_soa.a[3][i] = e1.x, g_soa.a[4][i] = e1.y, g_soa.a[5][i] = e1.z;
        g_soa.a[6][i] = e2.x, g_soa.a[7][i] = e2.y, g_soa.a[8][i] = e2.z;
    }
}
static void soa_free(void) {
    for (int c = 0; c < 9; ++c) free(g_soa.a[c]), g_soa.a[c] = NULL;
}

/* (float)((double)x * inv) for 8 lanes, inv given as two 4-lane halves */
RST_AVX2 static inline __m256 scale8(__m256 x, __m256d inv_lo, __m256d inv_hi) {
    const __m256d lo = _mm256_mul_pd(_mm256_cvtps_pd(_mm256_castps256_ps128(x)), inv_lo);
    const __m256d hi = _mm256_mul_pd(_mm256_cvtps_pd(_mm256_extractf128_ps(x, 1)), inv_hi);
    return _mm256_set_m128(_mm256_cvtpd_ps(hi), _mm256_cvtpd_ps(lo));
}
RST_AVX2 static inline __m256 dot8(__m256 ax, __m256 ay, __m256 az, __m256 bx, __m256 by, __m256 bz) {
    /* vec.h:95-101: ((0 + ax*bx) + ay*by) + az*bz; 0 + x is exact except for x = -0, which it turns into +0 */
    __m256 r = _mm256_add_ps(_mm256_setzero_ps(), _mm256_mul_ps(ax, bx));
    r = _mm256_add_ps(r, _mm256_mul_ps(ay, by));
    return _mm256_add_ps(r, _mm256_mul_ps(az, bz));
}
/* tests triangles i..i+7 against t_now; returns the lane mask of accepts and their t2 / v2 */
RST_AVX2 static inline int tri_test8(v3 o, v3 d, int i, float t_now, float t2_out[8], float v2_out[8]) {
    const __m256 v0x = _mm256_load_ps(g_soa.a[0] + i), v0y = _mm256_load_ps(g_soa.a[1] + i), v0z = _mm256_load_ps(g_soa.a[2] + i);
    const __m256 e1x = _mm256_load_ps(g_soa.a[3] + i), e1y = _mm256_load_ps(g_soa.a[4] + i), e1z = _mm256_load_ps(g_soa.a[5] + i);
    const __m256 e2x = _mm256_load_ps(g_soa.a[6] + i), e2y = _mm256_load_ps(g_soa.a[7] + i), e2z = _mm256_load_ps(g_soa.a[8] + i);
    const __m256 dx = _mm256_set1_ps(d.x), dy = _mm256_set1_ps(d.y), dz = _mm256_set1_ps(d.z);
    const __m256 eps = _mm256_set1_ps(EPS), one = _mm256_set1_ps(1.0f);
    /* pvec = cross(dir, edge2) */
    const __m256 px = _mm256_sub_ps(_mm256_mul_ps(dy, e2z), _mm256_mul_ps(dz, e2y));
    const __m256 py = _mm256_sub_ps(_mm256_mul_ps(dz, e2x), _mm256_mul_ps(dx, e2z));
    const __m256 pz = _mm256_sub_ps(_mm256_mul_ps(dx, e2y), _mm256_mul_ps(dy, e2x));
    const __m256 detf = dot8(e1x, e1y, e1z, px, py, pz);
    /* det > -EPS && det < EPS (as doubles of floats: same truth values as the float compares) */
    __m256 rej = _mm256_and_ps(_mm256_cmp_ps(detf, _mm256_sub_ps(_mm256_setzero_ps(), eps), _CMP_GT_OQ), _mm256_cmp_ps(detf, eps, _CMP_LT_OQ));
    const __m256d inv_lo = _mm256_div_pd(_mm256_set1_pd(1.0), _mm256_cvtps_pd(_mm256_castps256_ps128(detf)));
    const __m256d inv_hi = _mm256_div_pd(_mm256_set1_pd(1.0), _mm256_cvtps_pd(_mm256_extractf128_ps(detf, 1)));
    const __m256 tx = _mm256_sub_ps(_mm256_set1_ps(o.x), v0x), ty = _mm256_sub_ps(_mm256_set1_ps(o.y), v0y), tz = _mm256_sub_ps(_mm256_set1_ps(o.z), v0z);
    const __m256 u2 = scale8(dot8(tx, ty, tz, px, py, pz), inv_lo, inv_hi);
    rej = _mm256_or_ps(rej, _mm256_or_ps(_mm256_cmp_ps(u2, eps, _CMP_LT_OQ), _mm256_cmp_ps(u2, one, _CMP_GT_OQ)));
    /* qvec = cross(tvec, edge1) */
    const __m256 qx = _mm256_sub_ps(_mm256_mul_ps(ty, e1z), _mm256_mul_ps(tz, e1y));
    const __m256 qy = _mm256_sub_ps(_mm256_mul_ps(tz, e1x), _mm256_mul_ps(tx, e1z));
    const __m256 qz = _mm256_sub_ps(_mm256_mul_ps(tx, e1y), _mm256_mul_ps(ty, e1x));
    const __m256 v2 = scale8(dot8(dx, dy, dz, qx, qy, qz), inv_lo, inv_hi);
    rej = _mm256_or_ps(rej, _mm256_or_ps(_mm256_cmp_ps(v2, eps, _CMP_LT_OQ), _mm256_cmp_ps(_mm256_add_ps(u2, v2), one, _CMP_GT_OQ)));
    const __m256 t2 = scale8(dot8(e2x, e2y, e2z, qx, qy, qz), inv_lo, inv_hi);
    rej = _mm256_or_ps(rej, _mm256_or_ps(_mm256_cmp_ps(t2, eps, _CMP_LT_OQ), _mm256_cmp_ps(t2, _mm256_set1_ps(t_now), _CMP_GE_OQ)));
    _mm256_storeu_ps(t2_out, t2);
    _mm256_storeu_ps(v2_out, v2);
    return ~_mm256_movemask_ps(rej) & 0xff;
}

RST_AVX2 static int closest_hit_simd(const rst_scene *sc, v3 o, v3 d, float *t, float *v, int64_t *tests) {
    int best = -1;
    float alias_uv = *v, t2[8], v2[8];
    for (int i = 0; i < g_soa.n_pad; i += 8) {
        int m = tri_test8(o, d, i, *t, t2, v2);
        while (m) { /* in index order, against the running t (ray_triangle.h:49) */
            const int l = __builtin_ctz(m);
            m &= m - 1;
            if (t2[l] >= *t) continue;
            *t = t2[l], alias_uv = v2[l], best = i + l;
        }
    }
    *v = alias_uv;
    *tests += g_soa.n;
    for (int s = 0; s < sc->n_spheres; ++s)
        if (sphere_test(o, d, sc->sphere_cr + 4 * s, t)) best = g_soa.n + s;
    return best;
}

RST_AVX2 static int first_occluder_simd(const rst_scene *sc, v3 o, v3 d, float *t, int64_t *tests) {
    float t2[8], v2[8];
    for (int i = 0; i < g_soa.n_pad; i += 8) {
        const int m = tri_test8(o, d, i, *t, t2, v2);
        if (m) {
            const int l = __builtin_ctz(m);
            *t = t2[l];
            *tests += i + l + 1;
            return i + l;
        }
    }
    *tests += g_soa.n;
    for (int s = 0; s < sc->n_spheres; ++s)
        if (sphere_test(o, d, sc->sphere_cr + 4 * s, t)) return g_soa.n + s;
    return -1;
}
#else
static int g_simd = 0;
int rst_set_simd(int on) { (void)on; return 0; }
static void soa_build(const rst_scene *sc) { (void)sc; }
static void soa_free(void) {}
#define closest_hit_simd closest_hit_scalar
#define first_occluder_simd first_occluder_scalar
#endif

static int closest_hit(const rst_scene *sc, v3 o, v3 d, float *t, float *v, int64_t *tests) {
    return g_simd ? closest_hit_simd(sc, o, d, t, v, tests) : closest_hit_scalar(sc, o, d, t, v, tests);
}
static int first_occluder(const rst_scene *sc, v3 o, v3 d, float *t, int64_t *tests) {
    return g_simd ? first_occluder_simd(sc, o, d, t, tests) : first_occluder_scalar(sc, o, d, t, tests);
}

static int geom_of_tri(const rst_scene *sc, int tri) {
    int lo = 0, hi = sc->n_geoms; /* largest g with offset[g] <= tri */
    while (hi - lo > 1) {
        int mid = (lo + hi) / 2;
        if (sc->geom_tri_offset[mid] <= tri)
            lo = mid;
        else
            hi = mid;
    }
    return lo;
}

typedef struct {
    int tri;
    float t, v;
} hit_rec;

/* camera.h:31-34 for image-plane parameters (s, t) */
static inline v3 dir_st(const float cam[12], float s, float tt) {
    v3 origin = ld3(cam), llc = ld3(cam + 3), hor = ld3(cam + 6), ver = ld3(cam + 9);
    return norm3(sub3(add3(add3(llc, mul3(hor, s)), mul3(ver, tt)), origin));
}

static inline v3 primary_dir(const float cam[12], int W, int H, int w, int h) {
    /* main.cpp:709-710 */
    float s = (float)w / (W - 1);
    float tt = (float)h / (H - 1);
    return dir_st(cam, s, tt);
}

static hit_rec trace_dir(const rst_scene *sc, const float cam[12], v3 dir, int64_t *tests) {
    hit_rec r;
    r.t = FLT_MAX; /* main.cpp:715-717 */
    r.v = 0;
    r.tri = closest_hit(sc, ld3(cam), dir, &r.t, &r.v, tests);
    return r;
}

static hit_rec trace_primary(const rst_scene *sc, const float cam[12], int W, int H, int w, int h, int64_t *tests) {
    return trace_dir(sc, cam, primary_dir(cam, W, H, w, h), tests);
}

/* ---- shading block, main.cpp:723-789 ------------------------------------- */
static void shade_dir(const rst_scene *sc, const float cam[12], v3 dir, hit_rec hr, const int32_t *faceid, float rgb[3],
                      int32_t *occ_tri, int64_t *tests);
static void shade(const rst_scene *sc, const float cam[12], int W, int H, int w, int h, hit_rec hr,
                  const int32_t *faceid, float rgb[3], int32_t *occ_tri, int64_t *tests) {
    shade_dir(sc, cam, primary_dir(cam, W, H, w, h), hr, faceid, rgb, occ_tri, tests);
}
static void shade_dir(const rst_scene *sc, const float cam[12], v3 dir, hit_rec hr, const int32_t *faceid, float rgb[3],
                      int32_t *occ_tri, int64_t *tests) {
    rgb[0] = rgb[1] = rgb[2] = 0.f; /* vec3 ctor zero-fills, vec.h:44 */
    const int L = sc->n_lights;
    if (hr.tri < 0) {
        for (int l = 0; l < L && occ_tri; ++l) occ_tri[l] = -2;
        return;
    }
    const int n = n_tris_of(sc);
    v3 origin = ld3(cam);
    float t = hr.t;
    const float u = 0.f, v = hr.v;
    v3 N;
    const float *mat;
    if (hr.tri < n) {
        const float *p = sc->tri_verts + 9 * (size_t)hr.tri;
        int g = geom_of_tri(sc, hr.tri);
        N = norm3(cross3(sub3(ld3(p + 3), ld3(p)), sub3(ld3(p + 6), ld3(p)))); /* main.cpp:728-731 */
        if (sc->geom_has_normals && sc->geom_has_normals[g]) {                   /* main.cpp:733-738 */
            const float *q = sc->tri_normals + 9 * (size_t)hr.tri;
            v3 N0 = ld3(q), N1 = ld3(q + 3), N2 = ld3(q + 6);
            N = norm3(add3(add3(mul3(N1, u), mul3(N2, v)), mul3(N0, (1 - u - v))));
        }
        mat = sc->geom_material + 13 * (size_t)g;
    } else { /* extension: sphere normal */
        const float *cr = sc->sphere_cr + 4 * (size_t)(hr.tri - n);
        N = norm3(sub3(add3(origin, mul3(dir, t)), ld3(cr)));
        mat = sc->sphere_material + 13 * (size_t)(hr.tri - n);
    }
    v3 ka = ld3(mat), kd = ld3(mat + 3), ks = ld3(mat + 6), ke = ld3(mat + 9);
    float Ns = mat[12];
    for (int l = 0; l < L; ++l) {
        int lg = sc->light_geom[l];
        int fid = faceid[l];
        /* light.vertex[faceID]: the faceID-th de-indexed vertex (main.cpp:749-751) */
        v3 v0 = ld3(sc->tri_verts + 9 * (size_t)sc->geom_tri_offset[lg] + 3 * (size_t)fid);
        v3 zero = sub3(v0, v0);
        /* P = v0 + ((v1-v0)*r1 + (v2-v0)*r2), r in [0,1): zero vectors (main.cpp:753-754) */
        v3 P = add3(v0, add3(mul3(zero, 0.5f), mul3(zero, 0.5f)));
        v3 hit = add3(origin, mul3(dir, (t - EPS))); /* main.cpp:757-758 */
        v3 Lv = sub3(P, hit);
        float len = len3(Lv);
        t = len - EPS; /* main.cpp:764: clobbers the primary t */
        Lv = norm3(Lv);
        v3 c = div3(add3(mul3(ka, 0.5f), ke), (float)L); /* main.cpp:769-770 */
        int occ = first_occluder(sc, hit, Lv, &t, tests);
        if (occ_tri) occ_tri[l] = occ;
        if (occ >= 0) continue; /* main.cpp:772-773 */
        float d = dot3(N, Lv);
        if (d <= 0) continue; /* main.cpp:775-778 */
        v3 Hh = norm3(mul3(add3(N, Lv), 2.f));
        c = add3(c, div3(add3(mul3(kd, d), mul3(ks, powf(dot3(N, Hh), Ns))), (float)L));
        rgb[0] += c.x;
        rgb[1] += c.y;
        rgb[2] += c.z;
    }
}

/* main.cpp:679-684; u8 cannot hold the reference's negative / INT_MIN prints:
 * values are clamped into [0,255] (documented deviation for non-finite input). */
static uint8_t quantise(float x) {
    x = (x > 1.f) ? 1.f : x;
    float y = x * 255;
    if (!(y >= 0.f)) return 0;
    int q = (int)y;
    return (uint8_t)(q > 255 ? 255 : q);
}

/* ---- std::mt19937 (32-bit Mersenne Twister, the C++11 parameters) -------- */
typedef struct {
    uint32_t mt[624];
    int idx;
} mt19937;
static void mt_seed(mt19937 *g, uint32_t seed) {
    g->mt[0] = seed;
    for (int i = 1; i < 624; ++i) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
    g->idx = 624;
}
static uint32_t mt_next(mt19937 *g) {
    if (g->idx >= 624) {
        for (int i = 0; i < 624; ++i) {
            uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
            g->mt[i] = g->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        g->idx = 0;
    }
    uint32_t y = g->mt[g->idx++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}
/* libstdc++ uniform_int_distribution<int>(0, F-1) over a 32-bit engine:
 * Lemire's nearly-divisionless method (bits/uniform_int_dist.h, _S_nd) */
static int mt_uniform_int(mt19937 *g, uint32_t range /* = F */) {
    uint64_t product = (uint64_t)mt_next(g) * (uint64_t)range;
    uint32_t low = (uint32_t)product;
    if (low < range) {
        uint32_t threshold = (0u - range) % range;
        while (low < threshold) {
            product = (uint64_t)mt_next(g) * (uint64_t)range;
            low = (uint32_t)product;
        }
    }
    return (int)(product >> 32);
}

void rst_replay_faceids(const rst_scene *sc, int W, int H, uint32_t seed, const uint8_t *hit, int32_t *faceid) {
    mt19937 g;
    mt_seed(&g, seed);
    const int L = sc->n_lights;
    for (int h = H - 1; h >= 0; --h) { /* main.cpp:628, 704 */
        for (int w = 0; w < W; ++w) {
            size_t i = (size_t)h * W + w;
            for (int l = 0; l < L; ++l) {
                int fid = -1;
                if (hit[i]) {
                    int lg = sc->light_geom[l];
                    uint32_t F = (uint32_t)(sc->geom_tri_offset[lg + 1] - sc->geom_tri_offset[lg]);
                    fid = mt_uniform_int(&g, F); /* main.cpp:743-748 */
                    (void)mt_next(&g);           /* uniform_real<float>: one draw each (main.cpp:753-754) */
                    (void)mt_next(&g);
                }
                faceid[i * L + l] = fid;
            }
        }
    }
}

/* ---- extension, parity unpinned: counter-based RNG and n x n stratified jitter ------------
 * Restates the device's definitions (esctp1raytracer_b200/csrc/kernels.cuh: mix32, hash_faceid,
 * Bands::pixel_st) so that the multi-sample path has a CPU checker; there is no reference code. */
static uint32_t mix32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}
static int hash_faceid(uint32_t seed, uint32_t image_index, uint32_t light, uint32_t F) {
    uint32_t h = mix32((seed ^ 0x9e3779b9u) + image_index);
    h = mix32(h ^ (light * 0x85ebca6bu + 0xc2b2ae35u));
    return (int)(((uint64_t)h * F) >> 32);
}
static void pixel_st(int W, int H, int w, int h, int spp_n, int sample, uint32_t seed, float *s, float *t) {
    float fw = (float)w, fh = (float)h;
    if (spp_n > 1) {
        uint32_t h1 = mix32(mix32((seed ^ 0x51ed270bu) + (uint32_t)(h * W + w)) ^ ((uint32_t)sample * 0x9e3779b1u + 0x7f4a7c15u));
        uint32_t h2 = mix32(h1 + 0x632be5abu);
        float u1 = (float)(h1 >> 8) * 5.9604644775390625e-08f;
        float u2 = (float)(h2 >> 8) * 5.9604644775390625e-08f;
        float jx = ((float)(sample % spp_n) + u1) / (float)spp_n;
        float jy = ((float)(sample / spp_n) + u2) / (float)spp_n;
        fw = fw + (jx - 0.5f);
        fh = fh + (jy - 0.5f);
    }
    *s = fw / (float)(W - 1);
    *t = fh / (float)(H - 1);
}

typedef struct {
    const rst_scene *sc;
    const float *cam;
    int W, H, spp_n, tid, n_threads;
    uint32_t seed;
    float *rgb;
    uint8_t *rgb8;
} spp_job;

static uint8_t quantise(float x);

static void *spp_worker(void *arg) {
    spp_job *j = (spp_job *)arg;
    const int L = j->sc->n_lights, S = j->spp_n * j->spp_n;
    int64_t tests = 0;
    for (int h = j->tid; h < j->H; h += j->n_threads) {
        for (int w = 0; w < j->W; ++w) {
            float total[3] = {0, 0, 0};
            for (int smp = 0; smp < S; ++smp) {
                float s, t, rgb[3];
                int32_t fid[64];
                pixel_st(j->W, j->H, w, h, j->spp_n, smp, j->seed, &s, &t);
                v3 dir = dir_st(j->cam, s, t);
                hit_rec hr = trace_dir(j->sc, j->cam, dir, &tests);
                const uint32_t seed_s = j->seed + (uint32_t)smp * 0x9e3779b1u;
                for (int l = 0; l < L; ++l) {
                    int lg = j->sc->light_geom[l];
                    uint32_t F = (uint32_t)(j->sc->geom_tri_offset[lg + 1] - j->sc->geom_tri_offset[lg]);
                    fid[l] = hash_faceid(seed_s, (uint32_t)(h * j->W + w), (uint32_t)l, F);
                }
                shade_dir(j->sc, j->cam, dir, hr, fid, rgb, NULL, &tests);
                for (int c = 0; c < 3; ++c) total[c] = smp == 0 ? rgb[c] : total[c] + rgb[c];
            }
            size_t i = (size_t)h * j->W + w, k = (size_t)(j->H - 1 - h) * j->W + w;
            for (int c = 0; c < 3; ++c) {
                float v = S > 1 ? total[c] / (float)S : total[c];
                if (j->rgb) j->rgb[3 * i + c] = v;
                if (j->rgb8) j->rgb8[3 * k + c] = quantise(v);
            }
        }
    }
    return NULL;
}

int rst_render_spp(const rst_scene *sc, const float cam[12], int W, int H, uint32_t seed, int spp_n, int n_threads, float *rgb,
                   uint8_t *rgb8) {
    if (sc->n_lights > 64 || spp_n < 1) return -1;
    if (n_threads < 1) n_threads = 1;
    if (g_simd) soa_build(sc);
    spp_job *js = (spp_job *)malloc(sizeof(spp_job) * n_threads);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
    for (int k = 0; k < n_threads; ++k) {
        spp_job j = {sc, cam, W, H, spp_n, k, n_threads, seed, rgb, rgb8};
        js[k] = j;
        pthread_create(&th[k], NULL, spp_worker, &js[k]);
    }
    for (int k = 0; k < n_threads; ++k) pthread_join(th[k], NULL);
    free(js);
    free(th);
    if (g_simd) soa_free();
    return 0;
}

/* ---- drivers --------------------------------------------------------------- */
typedef struct {
    const rst_scene *sc;
    const float *cam;
    int W, H, phase, tid, n_threads, h_lo, h_hi;
    hit_rec *hits;
    const int32_t *faceid;
    rst_outputs *out;
    int64_t tests[2];
    /* pixel-list mode */
    int n;
    const int32_t *pw, *ph;
} job;

static void *frame_worker(void *arg) {
    job *j = (job *)arg;
    const int L = j->sc->n_lights;
    for (int h = j->h_lo + j->tid; h < j->h_hi; h += j->n_threads) {
        for (int w = 0; w < j->W; ++w) {
            size_t i = (size_t)h * j->W + w;
            if (j->phase == 0) {
                j->hits[i] = trace_primary(j->sc, j->cam, j->W, j->H, w, h, &j->tests[0]);
            } else {
                float rgb[3];
                int32_t occ[64];
                shade(j->sc, j->cam, j->W, j->H, w, h, j->hits[i], j->faceid + i * L, rgb, occ, &j->tests[1]);
                rst_outputs *o = j->out;
                if (o->rgb) memcpy(o->rgb + 3 * i, rgb, sizeof rgb);
                if (o->occ_tri)
                    for (int l = 0; l < L; ++l) o->occ_tri[i * L + l] = occ[l];
                if (o->rgb8) {
                    size_t k = (size_t)(j->H - 1 - h) * j->W + w;
                    for (int c = 0; c < 3; ++c) o->rgb8[3 * k + c] = quantise(rgb[c]);
                }
            }
        }
    }
    return NULL;
}

static void run_jobs(job *tmpl, int n_threads, void *(*fn)(void *), int64_t *tests_acc) {
    if (n_threads < 1) n_threads = 1;
    job *js = (job *)malloc(sizeof(job) * n_threads);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
    for (int k = 0; k < n_threads; ++k) {
        js[k] = *tmpl;
        js[k].tid = k;
        js[k].n_threads = n_threads;
        js[k].tests[0] = js[k].tests[1] = 0;
        if (n_threads > 1) pthread_create(&th[k], NULL, fn, &js[k]);
    }
    if (n_threads == 1) fn(&js[0]);
    for (int k = 0; k < n_threads; ++k) {
        if (n_threads > 1) pthread_join(th[k], NULL);
        tests_acc[0] += js[k].tests[0];
        tests_acc[1] += js[k].tests[1];
    }
    free(js);
    free(th);
}

int rst_render(const rst_scene *sc, const float cam[12], int W, int H, uint32_t seed, const int32_t *faceid_in,
               int n_threads, rst_outputs *out) {
    if (sc->n_lights > 64) return -1;
    const size_t P = (size_t)W * H;
    const int L = sc->n_lights;
    hit_rec *hits = (hit_rec *)malloc(sizeof(hit_rec) * P);
    int64_t tests[2] = {0, 0};
    job j;
    memset(&j, 0, sizeof j);
    j.sc = sc, j.cam = cam, j.W = W, j.H = H, j.hits = hits, j.out = out;
    j.h_lo = 0, j.h_hi = H;
    j.phase = 0;
    if (g_simd) soa_build(sc);
    run_jobs(&j, n_threads, frame_worker, tests);
    int32_t *faceid = NULL;
    if (!faceid_in) {
        uint8_t *mask = (uint8_t *)malloc(P);
        for (size_t i = 0; i < P; ++i) mask[i] = hits[i].tri >= 0;
        faceid = (int32_t *)malloc(sizeof(int32_t) * P * (L > 0 ? L : 1));
        rst_replay_faceids(sc, W, H, seed, mask, faceid);
        free(mask);
        faceid_in = faceid;
    }
    for (size_t i = 0; i < P; ++i) {
        if (out->tri) out->tri[i] = hits[i].tri;
        if (out->t) out->t[i] = hits[i].t;
        if (out->v) out->v[i] = hits[i].v;
    }
    if (out->faceid) memcpy(out->faceid, faceid_in, sizeof(int32_t) * P * L);
    j.phase = 1;
    j.faceid = faceid_in;
    run_jobs(&j, n_threads, frame_worker, tests);
    if (g_simd) soa_free();
    if (out->n_tests) out->n_tests[0] = tests[0], out->n_tests[1] = tests[1];
    free(hits);
    free(faceid);
    return 0;
}

static void *pixel_worker(void *arg) {
    job *j = (job *)arg;
    const int L = j->sc->n_lights;
    rst_outputs *o = j->out;
    for (int k = j->tid; k < j->n; k += j->n_threads) {
        int w = j->pw[k], h = j->ph[k];
        hit_rec hr = trace_primary(j->sc, j->cam, j->W, j->H, w, h, &j->tests[0]);
        float rgb[3];
        int32_t occ[64];
        shade(j->sc, j->cam, j->W, j->H, w, h, hr, j->faceid + (size_t)k * L, rgb, occ, &j->tests[1]);
        if (o->tri) o->tri[k] = hr.tri;
        if (o->t) o->t[k] = hr.t;
        if (o->v) o->v[k] = hr.v;
        if (o->rgb) memcpy(o->rgb + 3 * (size_t)k, rgb, sizeof rgb);
        if (o->occ_tri)
            for (int l = 0; l < L; ++l) o->occ_tri[(size_t)k * L + l] = occ[l];
        if (o->rgb8)
            for (int c = 0; c < 3; ++c) o->rgb8[3 * (size_t)k + c] = quantise(rgb[c]);
    }
    return NULL;
}

int rst_render_pixels(const rst_scene *sc, const float cam[12], int W, int H, int n, const int32_t *pw,
                      const int32_t *ph, const int32_t *faceids, int n_threads, rst_outputs *out) {
    if (sc->n_lights > 64) return -1;
    int64_t tests[2] = {0, 0};
    job j;
    memset(&j, 0, sizeof j);
    j.sc = sc, j.cam = cam, j.W = W, j.H = H, j.out = out;
    j.n = n, j.pw = pw, j.ph = ph, j.faceid = faceids;
    if (g_simd) soa_build(sc);
    run_jobs(&j, n_threads, pixel_worker, tests);
    if (g_simd) soa_free();
    if (out->n_tests) out->n_tests[0] = tests[0], out->n_tests[1] = tests[1];
    return 0;
}
