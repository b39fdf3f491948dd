/* ORACLE / TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's
 * per-pixel render path in plain C.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load the library built from restated.c.  The
 * product never links, loads or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_pin.py checks this restatement
 * bit-for-bit (all three float channels of every pixel, hit ids, t, replayed
 * faceIDs) against the unmodified reference compiled into oracle/_ref (see
 * oracle/ref_harness.cpp) on the shipped Cornell models and on multi-light
 * synthetic scenes, and against the fixtures committed in tests/golden/.
 * The sphere / multi-sample extensions (no reference code exists,
 * src/intersect.h:1-3 is an empty comment) are "parity unpinned".
 */
#ifndef RESTATED_H
#define RESTATED_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int32_t n_geoms;
    const int32_t *geom_tri_offset;  /* [n_geoms+1] first triangle of each geometry        */
    const float *tri_verts;          /* [n_tris*9]  v0 v1 v2, geometry-major / face-minor  */
    const float *tri_normals;        /* [n_tris*9]  or NULL                                 */
    const int32_t *geom_has_normals; /* [n_geoms]   or NULL                                 */
    const float *geom_material;      /* [n_geoms*13] ka kd ks ke Ns                         */
    int32_t n_lights;
    const int32_t *light_geom;       /* [n_lights]  geometry index, light_sources order     */
    /* extension (parity unpinned): analytic spheres tested after all triangles */
    int32_t n_spheres;
    const float *sphere_cr;          /* [n_spheres*4] centre xyz, radius */
    const float *sphere_material;    /* [n_spheres*13] */
} rst_scene;

typedef struct {
    /* all indexed by image index h*W+w unless noted; any pointer may be NULL */
    int32_t *tri;        /* flat triangle index of the closest hit, -1 miss; spheres: n_tris + s */
    float *t;            /* closest-hit t (FLT_MAX on miss) */
    float *v;            /* closest-hit v (the reference's u stays 0, main.cpp:307/310) */
    int32_t *faceid;     /* [P*L] faceID used per (pixel, light), -1 when no shadow ray */
    int32_t *occ_tri;    /* [P*L] first in-order occluder (flat index), -1 unoccluded, -2 no shadow ray */
    float *rgb;          /* [P*3] float accumulator */
    uint8_t *rgb8;       /* [P*3] packed u8 in PPM row order (row 0 = h=H-1) */
    int64_t *n_tests;    /* [2] ray-triangle tests executed: primary, shadow */
} rst_outputs;

/* camera.h:16-29 -> origin, lower_left_corner, horizontal, vertical */
void rst_camera(const float eye[3], const float look[3], const float vup[3], float vfov, float aspect, float out[12]);

/* std::mt19937 + the libstdc++ draws scan_row makes (main.cpp:743-754), in scan
 * order h=H-1..0, w=0..W-1, for pixels with hit[h*W+w] != 0.  faceid: [P*L]. */
void rst_replay_faceids(const rst_scene *sc, int W, int H, uint32_t seed, const uint8_t *hit, int32_t *faceid);

/* Full frame.  faceid_in == NULL -> faceIDs come from rst_replay_faceids(seed).
 * rows [h_lo, h_hi) only are rendered when h_lo < h_hi (others untouched; the
 * RNG replay still walks the whole frame).  n_threads splits rows. */
int rst_render(const rst_scene *sc, const float cam[12], int W, int H, uint32_t seed, const int32_t *faceid_in,
               int n_threads, rst_outputs *out);

/* Pixel subset (pw[k], ph[k]); faceids [n*L] required.  Outputs indexed by k. */
int rst_render_pixels(const rst_scene *sc, const float cam[12], int W, int H, int n, const int32_t *pw,
                      const int32_t *ph, const int32_t *faceids, int n_threads, rst_outputs *out);

/* extension, parity unpinned (no reference code): spp_n x spp_n stratified jittered samples per pixel,
 * counter-based faceIDs; rgb [P*3] image index order (mean over samples), rgb8 PPM order */
int rst_render_spp(const rst_scene *sc, const float cam[12], int W, int H, uint32_t seed, int spp_n, int n_threads, float *rgb,
                   uint8_t *rgb8);

/* SURVEY 8f-4 CPU comparator: run the triangle loops 8 triangles at a time (AVX2), bit-identical to the scalar
 * loops by construction and pinned against them.  Process-wide switch; returns the mode in force (0 when the CPU
 * has no AVX2).  One render at a time while it is on. */
int rst_set_simd(int on);

/* single-call known-answer entry: Moller-Trumbore exactly as ray_triangle.h:7-57 */
int rst_intersect_triangle(const float orig[3], const float dir[3], const float v0[3], const float v1[3],
                           const float v2[3], float *t, float *u, float *v);

/* vec.h identities for the KATs of src/ispc/test.ispc:24-37 */
float rst_dot(const float a[3], const float b[3]);
void rst_cross(const float a[3], const float b[3], float out[3]);

#ifdef __cplusplus
}
#endif
#endif
