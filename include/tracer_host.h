/* tracer_host.h — host-side steps either side of the render path:
 * the OBJ/MTL -> flat-scene loader and the PPM writer.  Pure CPU, no CUDA; they exist so
 * that a user of the reference finds its whole `-m model.obj ... -o out.ppm` flow here.
 */
#ifndef TRACER_HOST_H
#define TRACER_HOST_H

#include <stdint.h>

#include "tracer_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tracer_scene_host tracer_scene_host; /* owns the arrays a tracer_scene_flat points to */

/* model::loadobj (src/scene/sceneloader.cpp:14-106) over tinyobjloader v1.0.5's LoadObj
 * (src/scene/tiny_obj_loader.h:1415-1721, triangulate = true), rewritten:
 *  - one geometry per OBJ shape; shapes split on `g` and `o` only; a `usemtl` inside a shape
 *    keeps accumulating into the same shape (tiny_obj_loader.h:1519-1546) and the shape takes
 *    the material of its FIRST face (sceneloader.cpp:52);
 *  - polygons are fan-triangulated (tiny_obj_loader.h:881-925); vertices are de-indexed
 *    (sceneloader.cpp:78-82); normals are normalised (sceneloader.cpp:84-89);
 *  - a geometry is a light when dot(ke,ke) > 0 (sceneloader.cpp:63-64, 101-103);
 *  - numbers go through the same decimal->double->float conversion as tinyobj's tryParseDouble
 *    (tiny_obj_loader.h:465-580), so vertex floats are bit-identical;
 *  - ANY loader warning is fatal, as in the reference (sceneloader.cpp:27-30): missing .mtl,
 *    a material defining both `d` and `Tr` (tiny_obj_loader.h:1105-1128) ...
 * Returns 0 and *out, or TRACER_ERR_INVALID with the message in tracer_host_last_error(). */
int tracer_scene_load_obj(const char *obj_path, tracer_scene_host **out);
const tracer_scene_flat *tracer_scene_host_flat(const tracer_scene_host *scene);

/* The reference's optional flatten pass, flatten_scene (src/simplify/flatten.cpp:50-82): every face of
 * every geometry copied into one flat triangle array (geometry major, face minor, each triangle keeping
 * its geom_id / prim_id, flatten.cpp:62-63) and sorted by the x coordinate of its FIRST vertex
 * (comparator leftMostTriangle, flatten.cpp:20-27).  Triangle count is unchanged; what changes is the
 * iteration order, i.e. which triangle wins a tie in t and which occluder is "first in order".
 * std::sort (flatten.cpp:78) leaves equal keys in unspecified order; this helper sorts stably.
 * The result is an ordinary flat scene: maximal runs of one original geometry become geometries (same
 * material), and each light geometry is appended once more at the end in its original face order so
 * that light.vertex[faceID] (src/main.cpp:749) is unchanged — the copies cannot alter any hit or
 * occlusion decision (see host_io.cpp).  tracer_scene_host_origin() maps a triangle of the sorted
 * scene back to (geom_id, prim_id) of the input. */
int tracer_scene_flatten_sorted(const tracer_scene_flat *in, tracer_scene_host **out);
int tracer_scene_host_origin(const tracer_scene_host *scene, const int32_t **geom, const int32_t **prim);
void tracer_scene_host_free(tracer_scene_host *scene);
const char *tracer_host_last_error(void);

/* The reference's writer (src/main.cpp:658-689) emits ASCII P3: "P3\nW H\n255\n" then one
 * "r g b\n" line per pixel, rows top to bottom.  rgb is packed u8 in that row order (what
 * tracer_cuda_render returns).  binary != 0 writes P6 instead (same pixels, 3 bytes each). */
int tracer_write_ppm(const char *path, const uint8_t *rgb, int32_t width, int32_t height, int32_t binary);

#ifdef __cplusplus
}
#endif
#endif
