/* tracer_cuda.h — C ABI of the B200 (sm_100a) renderer for the per-pixel hot
 * path of pg42819/EscTp1RayTracer.
 *
 * Position in the reference: this library is called exactly where the
 * reference calls its ISPC accelerator,
 *     ispc::trace(W, H, ispc_cam, nTri, tris, nLights, lights, nLightFaces,
 *                 lightFaces, flat_image, debug, test)        src/main.cpp:619-624
 * (export declared at src/ispc/trace.ispc:86-92), after the scene has been
 * flattened to plain arrays (src/simplify/flatten_iscp.cpp:35-111) and the
 * camera marshalled (flatten_iscp.cpp:117-128).  It computes what the serial
 * path computes in scan_row (src/main.cpp:698-791) followed by the quantiser
 * of the PPM writer (src/main.cpp:679-684).
 *
 * Plain pointers and sizes only; no C++ or torch types.  All functions return
 * 0 on success or a negative tracer_status; the message is available from
 * tracer_cuda_last_error().  There is NO CPU fallback: without a CUDA device
 * (or without the sm_100a kernels) every compute entry point fails with
 * TRACER_ERR_NO_DEVICE / TRACER_ERR_CUDA.
 */
#ifndef TRACER_CUDA_H
#define TRACER_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRACER_CUDA_ABI_VERSION 1

typedef enum tracer_status {
    TRACER_OK = 0,
    TRACER_ERR_INVALID = -1,   /* bad argument (NULL, negative size, bad index) */
    TRACER_ERR_NO_DEVICE = -2, /* no CUDA device / tracer_cuda_init not called  */
    TRACER_ERR_CUDA = -3,      /* a CUDA runtime call or kernel failed          */
    TRACER_ERR_NOMEM = -4,
    TRACER_ERR_STATE = -5      /* handle destroyed / wrong call order           */
} tracer_status;

/* Flat scene, structure of arrays, HOST pointers, caller owned.
 * Replaces ispc_triangle[] / ispc_light[] (src/ispc/ispc_helpers.h:16-29,
 * 52-56) and tracer::scene (src/scene/scene.h:9-44).
 * Triangle order MUST be the reference's iteration order — geometry major,
 * face minor (src/main.cpp:179-180) — because ties keep the lower index
 * (ray_triangle.h:49) and occlusion() returns the first face in order
 * (src/main.cpp:317-325).  The reference's ISPC flatten sorts by centroid x
 * (flatten_iscp.cpp:110); do NOT sort for serial-path parity. */
typedef struct tracer_scene_flat {
    int32_t n_geoms;
    const int32_t *geom_tri_offset;  /* [n_geoms+1] first triangle of geometry g; last = n_tris */
    const float *tri_verts;          /* [n_tris*9]  v0 v1 v2 (de-indexed, sceneloader.cpp:78-82) */
    const float *tri_normals;        /* [n_tris*9]  per-corner normals, or NULL                  */
    const int32_t *geom_has_normals; /* [n_geoms]   (!normals.empty(), main.cpp:733) or NULL     */
    const float *geom_material;      /* [n_geoms*13] ka[3] kd[3] ks[3] ke[3] Ns (scene.h:11-18)  */
    int32_t n_lights;
    const int32_t *light_geom;       /* [n_lights]  geometry index, in light_sources order       */
    /* Extension with no reference behaviour (src/intersect.h is empty):
     * analytic spheres, tested after all triangles, object index n_tris+s. */
    int32_t n_spheres;
    const float *sphere_cr;          /* [n_spheres*4]  centre xyz, radius */
    const float *sphere_material;    /* [n_spheres*13] */
} tracer_scene_flat;

/* The four vectors tracer::camera precomputes (src/scene/camera.h:16-29);
 * replaces ispc_cam (ispc_helpers.h:59-65). */
typedef struct tracer_camera {
    float origin[3];
    float lower_left_corner[3];
    float horizontal[3];
    float vertical[3];
} tracer_camera;

/* Host helper: the reference camera constructor, same arithmetic
 * (camera.h:16-29; main.cpp:548-551 uses vfov=60, vup=(0,1,0), aspect=W/H). */
void tracer_camera_lookat(const float eye[3], const float look[3], const float vup[3], float vfov_deg,
                          float aspect, tracer_camera *out);

typedef enum tracer_rng_mode {
    /* counter-based hash of (seed, pixel, light): independent of GPU count and
     * of scan order.  Production default. */
    TRACER_RNG_HASH = 0,
    /* std::mt19937(seed) consumed exactly as scan_row does (main.cpp:743-754,
     * three draws per hit pixel per light, scan order h=H-1..0, w=0..W-1,
     * libstdc++ uniform_int_distribution): reproduces the seeded serial path.
     * The draws are sequential over the whole frame's hit mask, so this mode
     * needs the whole frame in one call (band_count <= 1). */
    TRACER_RNG_MT19937 = 1,
    /* faceID per (pixel, light) supplied by the caller in opts->faceid. */
    TRACER_RNG_EXPLICIT = 2
} tracer_rng_mode;

typedef struct tracer_render_opts {
    uint32_t struct_size;  /* = sizeof(tracer_render_opts) */
    int32_t rng_mode;      /* tracer_rng_mode */
    uint32_t seed;
    const int32_t *faceid; /* HOST [W*H*n_lights], image index h*W+w; TRACER_RNG_EXPLICIT only */

    /* Row-band partition for multi-GPU (1 process per GPU).  The frame's PPM
     * rows (row 0 = h=H-1, main.cpp:662) are cut into bands of band_rows
     * rows; this call renders the bands b with b % band_count == band_index,
     * packed back to back in the output.  band_count<=1 renders everything. */
    int32_t band_rows;
    int32_t band_index;
    int32_t band_count;

    int32_t rgb_out_is_device; /* rgb_out is a device pointer (stays in HBM, e.g. for an NCCL gather) */
    void *cuda_stream;         /* cudaStream_t to launch on; NULL = the library's own (non-blocking) stream.  To order the
                                  render with work on the legacy default stream pass cudaStreamLegacy (0x1), not 0 */

    int32_t exhaustive_strict; /* debug: bypass the conservative filter, strict-test every pair */
    int32_t samples_per_pixel; /* extension (parity unpinned): 0/1 = reference; n*n stratified jitter */
    int32_t rays_per_thread;   /* tuning: 0 = auto, else 2, 4, 8, 16, 24 or 32 rays per thread in the closest-hit sweep
                                  (2, 4, 8 or 12 with jittered samples) */
    int32_t shadow_chunks;     /* tuning: 0 = auto; triangle chunks between shadow-ray compactions */
    int32_t bundle_cull;       /* OPTIONAL mode: hierarchical (bundle box -> warp box -> ray) evaluation of the same
                                  conservative filter; identical results.  1 = two-phase (dense block-box pass, then
                                  per-block survivor lists), 2 = single streaming sweep (also the fall-back of 1),
                                  3 = auto: two-phase unless the scene is so small that the default sweeps are faster */

    /* optional debug outputs, HOST pointers, indexed like the output rows
     * (local pixel k = local_row*W + w), any may be NULL */
    int32_t *out_tri;     /* closest-hit flat triangle index (n_tris+s for spheres), -1 miss */
    float *out_t;         /* closest-hit t */
    float *out_v;         /* closest-hit v */
    int32_t *out_occ_tri; /* [n_px*n_lights] first in-order occluder, -1 none, -2 no shadow ray */
    float *out_rgb;       /* [n_px*3] float accumulator before quantisation */
} tracer_render_opts;

typedef struct tracer_frame_stats {
    double ms_total;        /* CUDA events around all launches of the frame            */
    double ms_primary;      /* closest-hit sweep                                       */
    double ms_shadow;       /* any-hit sweeps, all lights                              */
    double ms_other;        /* table build, ray set-up, shading, quantise              */
    int64_t n_pixels;       /* pixels rendered by this call                            */
    int64_t n_primary_rays;
    int64_t n_shadow_rays;
    int64_t tests_primary;  /* ray-triangle pairs swept (filter evaluations), primary  */
    int64_t tests_shadow;   /* same, shadow sweeps                                     */
    int64_t tests_shadow_ref; /* sum over shadow rays of (first occluder index+1, else N): the reference's count */
    int64_t strict_evals;   /* pairs re-evaluated in reference arithmetic              */
    int64_t filter_misses;  /* exhaustive_strict only: strict accepts the filter would have lost (must be 0) */
    int32_t kernel_launches;
    int32_t n_sms;
    double flop_primary;    /* FP32 flops the closest-hit sweep executes per swept pair, all in the FMA pipe (FFMA = 2,
                               FADD = FMUL = 1).  Span form (no jitter: the R rays of a thread share q): per pair two
                               saturating adds and one multiply-add, per thread and triangle 4 FFMA: 4 + 8/R.  Jittered
                               samples (three-row form, every ray its own q): 6 FFMA + FMUL + FFMA = 15.
                               0 in bundle-cull mode                                                              */
    double flop_shadow;     /* same for the any-hit sweeps (span form, R = 16 q-sorted rays share one q-term per bound:
                               8 FFMA per thread and triangle): 4 + 16/R                                          */
    /* of which the multiply-adds that evaluate the bounds / edge rows (the rest is each pair's conjunction) */
    double flop_primary_edges;
    double flop_shadow_edges;
    int64_t pipeline_errors; /* exhaustive_strict only: the sweeps' own race check — staged table tiles (TMA pipeline,
                                per-warp stage recycling) that differed from their source when a warp started or
                                finished using them (must be 0) */
} tracer_frame_stats;

typedef struct tracer_device_info {
    char name[128];
    int32_t sm_count;
    int32_t cc_major, cc_minor;
    int32_t clock_khz; /* cudaDevAttrClockRate */
    int64_t total_mem;
    int64_t l2_bytes;
} tracer_device_info;

typedef struct tracer_scene_dev tracer_scene_dev; /* opaque: scene resident in HBM */

/* ---- lifetime ------------------------------------------------------------ */
int tracer_cuda_abi_version(void);
int tracer_cuda_init(int device_ordinal); /* cudaSetDevice + stream; makes that GPU the current context */
void tracer_cuda_shutdown(void);
const char *tracer_cuda_last_error(void);
int tracer_cuda_device_info(tracer_device_info *out);

/* ---- the drop-in call: host buffers in, packed 8-bit RGB out --------------
 * rgb_out: caller-owned, n_rows_of_this_call * W * 3 bytes, PPM row order,
 * fully overwritten (unlike ispc::trace, which += into an un-zeroed malloc,
 * trace.ispc:262-265 / main.cpp:593).  Uploads the scene, renders, downloads. */
int tracer_cuda_render(const tracer_scene_flat *scene, const tracer_camera *cam, int32_t width, int32_t height,
                       const tracer_render_opts *opts, uint8_t *rgb_out);

/* ---- resident-scene variant (scene stays in HBM across frames) ------------ */
int tracer_cuda_scene_create(const tracer_scene_flat *scene, tracer_scene_dev **out);
void tracer_cuda_scene_destroy(tracer_scene_dev *scene);
int tracer_cuda_render_scene(tracer_scene_dev *scene, const tracer_camera *cam, int32_t width, int32_t height,
                             const tracer_render_opts *opts, uint8_t *rgb_out);
int tracer_cuda_last_stats(tracer_scene_dev *scene, tracer_frame_stats *out);

/* ---- all GPUs of one box behind one call ------------------------------------
 * The reference has no multi-device path (its closest analogue is thread-per-row,
 * src/main.cpp:629-643).  One process; a context, stream and scene replica per GPU; one
 * host thread per GPU while a frame renders; one NCCL communicator set (ncclCommInitAll).
 * The frame's PPM rows are cut into bands of opts->band_rows (default 8) rows, band b is
 * rendered by GPU b % n_gpus, every GPU quantises its bands straight into its send
 * buffer and the packed bands are gathered on GPU 0 with grouped ncclSend/ncclRecv over
 * NVLink.  Results are byte-identical to a single-GPU frame (TRACER_RNG_HASH and
 * TRACER_RNG_EXPLICIT; TRACER_RNG_MT19937 needs the whole frame on one GPU).
 * After tracer_cuda_init_multi(n > 1) the drop-in tracer_cuda_render() uses all n GPUs. */
typedef struct tracer_scene_multi tracer_scene_multi; /* opaque: one scene replica per GPU */
int tracer_cuda_init_multi(int n_gpus);               /* devices 0 .. n_gpus-1 */
int tracer_cuda_multi_gpu_count(void);                /* 0 before tracer_cuda_init_multi */
int tracer_cuda_render_multi(const tracer_scene_flat *scene, const tracer_camera *cam, int32_t width, int32_t height,
                             const tracer_render_opts *opts, uint8_t *rgb_out);
int tracer_cuda_scene_create_multi(const tracer_scene_flat *scene, tracer_scene_multi **out);
void tracer_cuda_scene_destroy_multi(tracer_scene_multi *scene);
int tracer_cuda_render_scene_multi(tracer_scene_multi *scene, const tracer_camera *cam, int32_t width, int32_t height,
                                   const tracer_render_opts *opts, uint8_t *rgb_out); /* rgb_out: host, or GPU 0 if rgb_out_is_device */
int tracer_cuda_last_stats_multi(tracer_scene_multi *scene, tracer_frame_stats *out); /* counts summed, times of the slowest GPU */

/* number of PPM rows the band selection (band_rows, band_index, band_count) covers */
int32_t tracer_band_row_count(int32_t height, int32_t band_rows, int32_t band_index, int32_t band_count);

/* Rank 0 after the gather: scatter band-packed rank buffers (each padded to
 * rows_per_rank_padded rows) into one PPM-ordered frame.  Device pointers. */
int tracer_cuda_assemble_bands(const uint8_t *gathered_dev, uint8_t *frame_dev, int32_t width, int32_t height,
                               int32_t band_rows, int32_t band_count, int32_t rows_per_rank_padded, void *cuda_stream);

/* ---- host utilities -------------------------------------------------------- */
/* std::mt19937 replay of the reference's draws (main.cpp:743-754): hit is the
 * full-frame mask in image index order; faceid [W*H*n_lights]. */
int tracer_mt19937_faceids(const tracer_scene_flat *scene, int32_t width, int32_t height, uint32_t seed,
                           const uint8_t *hit, int32_t *faceid);

/* ---- measurement: FP32 peak microbenchmark ---------------------------------
 * variant 0: scalar FFMA chains; 1: packed FFMA2 (fma.rn.f32x2) chains;
 * returns achieved TFLOP/s (2 flop per FMA) over `iters` launches timed with
 * CUDA events. */
int tracer_cuda_fp32_peak(int32_t variant, int32_t iters, double *tflops_out, double *ms_out);

#ifdef __cplusplus
}
#endif
#endif /* TRACER_CUDA_H */
