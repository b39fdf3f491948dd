"""Host-side mirror of the reference's interface for the render path.

Reference                                   here
------------------------------------------  -----------------------------------
tracer::scene (src/scene/scene.h:9-44)      Scene  (flat SoA, reference order)
tracer::camera(lookfrom, lookat, vup,       Camera(lookfrom, lookat, vup, vfov,
  vfov, aspect)  (src/scene/camera.h:16)      aspect)
ispc::trace(W, H, cam, tris, lights, ...,   Renderer.trace(scene, cam, W, H) ->
  image)  (src/ispc/trace.ispc:86-92,         packed u8 RGB rows in PPM order
  called at src/main.cpp:619-624)
scan_row + PPM quantiser                    (what trace() computes)
  (src/main.cpp:698-791, 679-684)

Everything computes on the GPU through the C ABI in include/tracer_cuda.h;
there is no CPU path here.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import TracerError

RNG_HASH, RNG_MT19937, RNG_EXPLICIT = 0, 1, 2


def _ptr(a, ctype):
    return None if a is None else a.ctypes.data_as(C.POINTER(ctype))


@dataclass
class Scene:
    """Flat scene in the reference's iteration order: geometry major, face minor
    (src/main.cpp:179-180).  ``geom_material`` rows are ka[3] kd[3] ks[3] ke[3] Ns
    (src/scene/scene.h:11-18); ``light_geom`` is ``light_sources`` (scene.h:38)."""

    geom_tri_offset: np.ndarray
    tri_verts: np.ndarray
    geom_material: np.ndarray
    light_geom: np.ndarray
    tri_normals: np.ndarray | None = None
    geom_has_normals: np.ndarray | None = None
    sphere_cr: np.ndarray = field(default_factory=lambda: np.zeros((0, 4), np.float32))
    sphere_material: np.ndarray = field(default_factory=lambda: np.zeros((0, 13), np.float32))

    def __post_init__(self):
        self.geom_tri_offset = np.ascontiguousarray(self.geom_tri_offset, np.int32)
        self.tri_verts = np.ascontiguousarray(self.tri_verts, np.float32).reshape(-1, 3, 3)
        self.geom_material = np.ascontiguousarray(self.geom_material, np.float32).reshape(-1, 13)
        self.light_geom = np.ascontiguousarray(self.light_geom, np.int32).reshape(-1)
        if self.tri_normals is not None:
            self.tri_normals = np.ascontiguousarray(self.tri_normals, np.float32).reshape(-1, 3, 3)
        if self.geom_has_normals is None:
            self.geom_has_normals = np.zeros(self.n_geoms, np.int32)
        self.geom_has_normals = np.ascontiguousarray(self.geom_has_normals, np.int32)
        self.sphere_cr = np.ascontiguousarray(self.sphere_cr, np.float32).reshape(-1, 4)
        self.sphere_material = np.ascontiguousarray(self.sphere_material, np.float32).reshape(-1, 13)
        if len(self.geom_tri_offset) < 1 or self.geom_tri_offset[-1] != len(self.tri_verts):
            raise ValueError("geom_tri_offset[-1] must equal the triangle count")
        if len(self.geom_material) != self.n_geoms:
            raise ValueError("one material row per geometry")
        if self.geom_has_normals.any() and (self.tri_normals is None or len(self.tri_normals) != self.n_tris):
            raise ValueError("geom_has_normals set but tri_normals missing")

    @property
    def n_geoms(self):
        return len(self.geom_tri_offset) - 1

    @property
    def n_tris(self):
        return int(self.geom_tri_offset[-1])

    @property
    def n_lights(self):
        return len(self.light_geom)

    @property
    def faces_per_light(self):
        o = self.geom_tri_offset
        return np.array([o[g + 1] - o[g] for g in self.light_geom], np.int64)

    @staticmethod
    def load_obj(path: str) -> "Scene":
        """model::loadobj (src/scene/sceneloader.cpp:14-106): OBJ/MTL -> scene, through the library's own
        C++ loader (csrc/host_io.cpp).  Raises TracerError where the reference throws (e.g. any
        loader warning, sceneloader.cpp:27-30)."""
        lib = _lib.load()
        h = C.c_void_p()
        if lib.tracer_scene_load_obj(path.encode(), C.byref(h)) != 0:
            raise TracerError(lib.tracer_host_last_error().decode())
        try:
            return Scene._from_host(lib, h)
        finally:
            lib.tracer_scene_host_free(h)

    @staticmethod
    def _from_host(lib, h) -> "Scene":
        f = lib.tracer_scene_host_flat(h).contents
        G = f.n_geoms
        off = np.ctypeslib.as_array(f.geom_tri_offset, (G + 1,)).copy()
        N = int(off[-1])
        grab = lambda ptr, shape, dt: (np.ctypeslib.as_array(ptr, shape).astype(dt).copy() if int(np.prod(shape)) and ptr else np.zeros(shape, dt))
        hasn = grab(f.geom_has_normals, (G,), np.int32)
        return Scene(off, grab(f.tri_verts, (N, 3, 3), np.float32), grab(f.geom_material, (G, 13), np.float32),
                     grab(f.light_geom, (f.n_lights,), np.int32),
                     tri_normals=grab(f.tri_normals, (N, 3, 3), np.float32) if hasn.any() else None, geom_has_normals=hasn)

    def flatten_sorted(self):
        """flatten_scene (src/simplify/flatten.cpp:50-82): flat triangle array sorted by vertices[0].x, through the
        library's host helper (tracer_scene_flatten_sorted).  Returns (Scene in sorted iteration order,
        origin_geom [N'], origin_prim [N']) — the (geom_id, prim_id) every sorted triangle keeps (flatten.cpp:62-63)."""
        lib = _lib.load()
        h = C.c_void_p()
        cs = self.c_struct()
        if lib.tracer_scene_flatten_sorted(C.byref(cs), C.byref(h)) != 0:
            raise TracerError(lib.tracer_host_last_error().decode())
        try:
            sc = Scene._from_host(lib, h)
            pg, pp = C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)()
            _lib.check(lib.tracer_scene_host_origin(h, C.byref(pg), C.byref(pp)))
            return sc, np.ctypeslib.as_array(pg, (sc.n_tris,)).copy(), np.ctypeslib.as_array(pp, (sc.n_tris,)).copy()
        finally:
            lib.tracer_scene_host_free(h)

    def c_struct(self) -> _lib.SceneFlat:
        s = _lib.SceneFlat()
        s.n_geoms = self.n_geoms
        s.geom_tri_offset = _ptr(self.geom_tri_offset, C.c_int32)
        s.tri_verts = _ptr(self.tri_verts, C.c_float)
        s.tri_normals = _ptr(self.tri_normals, C.c_float)
        s.geom_has_normals = _ptr(self.geom_has_normals, C.c_int32)
        s.geom_material = _ptr(self.geom_material, C.c_float)
        s.n_lights = self.n_lights
        s.light_geom = _ptr(self.light_geom, C.c_int32)
        s.n_spheres = len(self.sphere_cr)
        s.sphere_cr = _ptr(self.sphere_cr, C.c_float)
        s.sphere_material = _ptr(self.sphere_material, C.c_float)
        return s


class Camera:
    """tracer::camera (src/scene/camera.h:16-29): same constructor arguments, same
    arithmetic (computed by the C library's host helper)."""

    def __init__(self, lookfrom, lookat, vup=(0.0, 1.0, 0.0), vfov=60.0, aspect=4.0 / 3.0):
        lib = _lib.load()
        self.c = _lib.CameraC()
        e = (C.c_float * 3)(*[float(x) for x in lookfrom])
        l = (C.c_float * 3)(*[float(x) for x in lookat])
        u = (C.c_float * 3)(*[float(x) for x in vup])
        lib.tracer_camera_lookat(e, l, u, C.c_float(vfov), C.c_float(float(np.float32(aspect))), C.byref(self.c))

    @staticmethod
    def for_frame(lookfrom, lookat, width, height):
        """The camera main() builds: vfov 60, vup (0,1,0), aspect = float(W)/H (main.cpp:548-551)."""
        return Camera(lookfrom, lookat, (0, 1, 0), 60.0, np.float32(width) / np.float32(height))

    def as_array(self):
        return np.array(list(self.c.origin) + list(self.c.lower_left_corner) + list(self.c.horizontal)
                        + list(self.c.vertical), np.float32)


def hash_faceids(seed: int, width: int, height: int, faces_per_light) -> np.ndarray:
    """numpy mirror of the device's counter-based faceID (kernels.cuh:hash_faceid):
    [H*W, L] in image index order h*W+w."""

    def mix(x):
        x = x.astype(np.uint32)
        x ^= x >> np.uint32(16)
        x = (x * np.uint32(0x7FEB352D)).astype(np.uint32)
        x ^= x >> np.uint32(15)
        x = (x * np.uint32(0x846CA68B)).astype(np.uint32)
        x ^= x >> np.uint32(16)
        return x

    with np.errstate(over="ignore"):
        idx = np.arange(width * height, dtype=np.uint32)
        base = mix(np.uint32((seed ^ 0x9E3779B9) & 0xFFFFFFFF) + idx)
        out = np.zeros((width * height, len(faces_per_light)), np.int32)
        for l, F in enumerate(faces_per_light):
            k = np.uint32((l * 0x85EBCA6B + 0xC2B2AE35) & 0xFFFFFFFF)
            h = mix(base ^ k)
            out[:, l] = ((h.astype(np.uint64) * np.uint64(F)) >> np.uint64(32)).astype(np.int32)
    return out


def band_row_count(height, band_rows, band_index, band_count) -> int:
    return int(_lib.load().tracer_band_row_count(height, band_rows, band_index, band_count))


@dataclass
class Frame:
    rgb8: np.ndarray  # [rows, W, 3] uint8, PPM row order (row 0 = h=H-1)
    stats: dict
    tri: np.ndarray | None = None
    t: np.ndarray | None = None
    v: np.ndarray | None = None
    occ_tri: np.ndarray | None = None
    rgb: np.ndarray | None = None


class ResidentScene:
    """A scene uploaded to HBM (tracer_cuda_scene_create)."""

    def __init__(self, renderer, scene: Scene):
        self.renderer = renderer
        self.scene = scene
        self.handle = C.c_void_p()
        cs = scene.c_struct()
        _lib.check(renderer.lib.tracer_cuda_scene_create(C.byref(cs), C.byref(self.handle)))

    def close(self):
        if self.handle:
            self.renderer.lib.tracer_cuda_scene_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiRenderer:
    """All GPUs of the box behind one call, one process (tracer_cuda_init_multi): interleaved 8-row bands, scene
    replicated, packed bands gathered on GPU 0 with NCCL inside the library.  The multi-process form (one rank per
    GPU, torch.distributed) lives in dist.py; both produce the bytes a single GPU produces."""

    def __init__(self, n_gpus: int):
        self.lib = _lib.load()
        _lib.check(self.lib.tracer_cuda_init_multi(n_gpus))
        self.n_gpus = n_gpus

    def upload(self, scene: Scene):
        h = C.c_void_p()
        cs = scene.c_struct()
        _lib.check(self.lib.tracer_cuda_scene_create_multi(C.byref(cs), C.byref(h)))
        return _MultiScene(self, scene, h)

    def trace(self, scene, camera: "Camera", width: int, height: int, *, rng_mode=RNG_HASH, seed=1, faceid=None,
              samples_per_pixel=0, bundle_cull=False, band_rows=0, out_device_ptr=None) -> "Frame":
        """``out_device_ptr``: leave the assembled frame in HBM on GPU 0 at that address instead of copying it to the host."""
        sc = scene.scene if isinstance(scene, _MultiScene) else scene
        o = _lib.RenderOpts()
        o.struct_size = C.sizeof(_lib.RenderOpts)
        o.rng_mode, o.seed = rng_mode, seed & 0xFFFFFFFF
        keep = None
        if rng_mode == RNG_EXPLICIT:
            keep = np.ascontiguousarray(faceid, np.int32)
            o.faceid = _ptr(keep, C.c_int32)
        o.samples_per_pixel, o.bundle_cull, o.band_rows = samples_per_pixel, int(bundle_cull), band_rows
        if out_device_ptr is not None:
            out, dst, o.rgb_out_is_device = None, C.c_void_p(out_device_ptr), 1
        else:
            out = np.empty((height, width, 3), np.uint8)
            dst = out.ctypes.data_as(C.c_void_p)
        stats = {}
        if isinstance(scene, _MultiScene):
            _lib.check(self.lib.tracer_cuda_render_scene_multi(scene.handle, C.byref(camera.c), width, height, C.byref(o), dst))
            st = _lib.FrameStats()
            _lib.check(self.lib.tracer_cuda_last_stats_multi(scene.handle, C.byref(st)))
            stats = st.asdict()
        else:
            cs = sc.c_struct()
            _lib.check(self.lib.tracer_cuda_render_multi(C.byref(cs), C.byref(camera.c), width, height, C.byref(o), dst))
        return Frame(rgb8=out, stats=stats)


class _MultiScene:
    def __init__(self, renderer, scene, handle):
        self.renderer, self.scene, self.handle = renderer, scene, handle

    def close(self):
        if self.handle:
            self.renderer.lib.tracer_cuda_scene_destroy_multi(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Renderer:
    """One CUDA device (tracer_cuda_init); see MultiRenderer for all GPUs of the box in one process."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        _lib.check(self.lib.tracer_cuda_init(device))
        self.device = device

    def device_info(self) -> dict:
        d = _lib.DeviceInfo()
        _lib.check(self.lib.tracer_cuda_device_info(C.byref(d)))
        return dict(name=d.name.decode(), sm_count=d.sm_count, cc=(d.cc_major, d.cc_minor), clock_khz=d.clock_khz,
                    total_mem=d.total_mem, l2_bytes=d.l2_bytes)

    def upload(self, scene: Scene) -> ResidentScene:
        return ResidentScene(self, scene)

    def _opts(self, scene, width, height, rng_mode, seed, faceid, bands, exhaustive, debug, n_px, keep, tuning):
        o = _lib.RenderOpts()
        o.struct_size = C.sizeof(_lib.RenderOpts)
        o.rng_mode, o.seed = rng_mode, seed & 0xFFFFFFFF
        if rng_mode == RNG_EXPLICIT:
            f = np.ascontiguousarray(faceid, np.int32)
            if f.size != width * height * scene.n_lights:
                raise ValueError("faceid must have W*H*n_lights entries (image index order)")
            keep.append(f)
            o.faceid = _ptr(f, C.c_int32)
        if bands is not None:
            o.band_rows, o.band_index, o.band_count = bands
        o.exhaustive_strict = int(exhaustive)
        o.rays_per_thread, o.shadow_chunks, o.samples_per_pixel, o.bundle_cull = tuning
        dbg = {}
        if debug:
            L = scene.n_lights
            dbg = dict(tri=np.zeros(n_px, np.int32), t=np.zeros(n_px, np.float32), v=np.zeros(n_px, np.float32),
                       occ_tri=np.full((n_px, L), -2, np.int32), rgb=np.zeros((n_px, 3), np.float32))
            o.out_tri, o.out_t, o.out_v = _ptr(dbg["tri"], C.c_int32), _ptr(dbg["t"], C.c_float), _ptr(dbg["v"], C.c_float)
            o.out_occ_tri, o.out_rgb = _ptr(dbg["occ_tri"], C.c_int32), _ptr(dbg["rgb"], C.c_float)
        return o, dbg

    def trace(self, scene, camera: Camera, width: int, height: int, *, rng_mode=RNG_HASH, seed=1, faceid=None,
              bands=None, exhaustive_strict=False, debug=False, out_device_ptr=None, stream=None,
              rays_per_thread=0, shadow_chunks=0, samples_per_pixel=0, bundle_cull=False, out=None) -> Frame:
        """Render; ``scene`` is a Scene (one-shot: upload + render, the drop-in call) or a
        ResidentScene.  ``bands`` = (band_rows, band_index, band_count).  With
        ``out_device_ptr`` the packed rows are left in HBM at that address; ``out`` is an optional
        caller-owned uint8 [rows, W, 3] host array (e.g. pinned) to receive them."""
        sc = scene.scene if isinstance(scene, ResidentScene) else scene
        rows = height if bands is None else band_row_count(height, *bands)
        n_px = rows * width
        keep = []
        o, dbg = self._opts(sc, width, height, rng_mode, seed, faceid, bands, exhaustive_strict, debug, n_px, keep,
                            (rays_per_thread, shadow_chunks, samples_per_pixel, int(bundle_cull)))
        if out_device_ptr is not None:
            out = None
            o.rgb_out_is_device = 1
            dst = C.c_void_p(out_device_ptr)
        else:
            if out is None:
                out = np.zeros((rows, width, 3), np.uint8)
            assert out.dtype == np.uint8 and out.size == rows * width * 3 and out.flags.c_contiguous
            dst = out.ctypes.data_as(C.c_void_p)
        if stream is not None:
            o.cuda_stream = C.c_void_p(stream)
        stats = {}
        if isinstance(scene, ResidentScene):
            _lib.check(self.lib.tracer_cuda_render_scene(scene.handle, C.byref(camera.c), width, height, C.byref(o), dst))
            st = _lib.FrameStats()
            _lib.check(self.lib.tracer_cuda_last_stats(scene.handle, C.byref(st)))
            stats = st.asdict()
        else:
            cs = sc.c_struct()
            _lib.check(self.lib.tracer_cuda_render(C.byref(cs), C.byref(camera.c), width, height, C.byref(o), dst))
        return Frame(rgb8=out, stats=stats, **dbg)

    def assemble_bands(self, gathered_ptr, frame_ptr, width, height, band_rows, band_count, rows_pad, stream=None):
        _lib.check(self.lib.tracer_cuda_assemble_bands(C.c_void_p(gathered_ptr), C.c_void_p(frame_ptr), width, height,
                                                       band_rows, band_count, rows_pad,
                                                       C.c_void_p(stream) if stream else None))

    def fp32_peak(self, variant=0, iters=10):
        tf, ms = C.c_double(), C.c_double()
        _lib.check(self.lib.tracer_cuda_fp32_peak(variant, iters, C.byref(tf), C.byref(ms)))
        return tf.value, ms.value


def write_ppm(path: str, rgb8: np.ndarray, binary: bool = False) -> None:
    """The reference's PPM writer (src/main.cpp:658-689): ASCII P3, rows top to bottom; P6 when binary."""
    a = np.ascontiguousarray(rgb8, np.uint8)
    h, w = a.shape[0], a.shape[1]
    lib = _lib.load()
    if lib.tracer_write_ppm(path.encode(), a.ctypes.data_as(C.c_void_p), w, h, int(binary)) != 0:
        raise TracerError(lib.tracer_host_last_error().decode())


def mt19937_faceids(scene: Scene, width: int, height: int, seed: int, hit_mask) -> np.ndarray:
    """std::mt19937 replay of scan_row's draws (host; tracer_mt19937_faceids)."""
    lib = _lib.load()
    hit = np.ascontiguousarray(hit_mask, np.uint8).reshape(-1)
    out = np.full((width * height, scene.n_lights), -1, np.int32)
    cs = scene.c_struct()
    _lib.check(lib.tracer_mt19937_faceids(C.byref(cs), width, height, seed & 0xFFFFFFFF, _ptr(hit, C.c_uint8),
                                          _ptr(out, C.c_int32)))
    return out
