/* ORACLE / TEST INFRASTRUCTURE ONLY — never included by the product.
 *
 * Stand-in for the header the ispc compiler would generate from the
 * reference's src/ispc/trace.ispc (no `ispc` binary exists in this image).
 * The reference includes it unconditionally (src/main.cpp:25,
 * src/simplify/flatten_iscp.h:3), so the serial path cannot be compiled
 * without it.  Layouts restate src/ispc/ispc_helpers.h:16-29, 52-56, 59-65;
 * the entry point restates the export at src/ispc/trace.ispc:86-92.
 * The harness defines ispc::trace as an empty function: the ISPC mode is
 * never exercised by the oracle (it is broken upstream).
 */
#pragma once
#include <stdint.h>
#ifdef __cplusplus
namespace ispc {
#endif
struct ispc_triangle {
    float vertices[3][3];
    float normals[3][3];
    int32_t prim_id;
    int32_t geom_id;
    int32_t has_normals;
    int32_t is_light;
    float ka[3];
    float kd[3];
    float ks[3];
    float ke[3];
    float Ns;
};
struct ispc_light {
    int32_t geom_id;
    int32_t *light_faces;
    int32_t num_light_faces;
};
struct ispc_cam {
    float lookfrom[3];
    float lookat[3];
    float vup[3];
    float vfov;
    float aspect;
};
#ifdef __cplusplus
extern "C" {
#endif
extern void trace(int32_t image_width, int32_t image_height, struct ispc_cam &cam,
                  int32_t num_triangles, struct ispc_triangle *triangles,
                  int32_t num_lights, struct ispc_light *lights,
                  int32_t num_light_triangles, struct ispc_triangle *light_triangles,
                  float *return_image, int32_t debug, int32_t test);
#ifdef __cplusplus
}
}
#endif
