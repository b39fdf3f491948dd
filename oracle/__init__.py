"""ORACLE / TEST INFRASTRUCTURE ONLY.

ctypes bindings for the two CPU checkers:

* ``Restated``  -> oracle/librestated.so  (our plain-C restatement, restated.c)
* ``RefOracle`` -> oracle/_ref/libref_oracle.so (the unmodified reference serial
  path compiled from /root/reference by oracle/Makefile; present only when it
  was built in the authoring container — the .so travels to the GPU box)

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.  Nothing under
``esctp1raytracer_b200/`` imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"
RESTATED_SO = os.path.join(HERE, "librestated.so")
REF_SO = os.path.join(HERE, "_ref", "libref_oracle.so")
DROPIN_BIN = os.path.join(HERE, "_ref", "ESCViewer2021_cuda")  # reference main.cpp + INTEGRATION.md section 3 patch
REF_RELEASE_SO = os.path.join(HERE, "_ref_release", "libref_oracle.so")  # -O3 -ffast-math: timing only, never parity

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(verbose: bool = False) -> None:
    """Compile the restatement, and the reference harness when /root/reference exists."""
    have_ref = os.path.exists(os.path.join(REF_ROOT, "src", "main.cpp"))
    targets = ["restated"] + (["ref", "ref_release"] if have_ref else [])
    if have_ref and os.path.exists(os.path.join(HERE, "..", "esctp1raytracer_b200", "libtracer_cuda.so")):
        targets.append("dropin")  # the reference's main.cpp + the INTEGRATION.md binding, linked with the product library
    r = subprocess.run(["make", "-C", HERE] + targets, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout, r.stderr)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed")


def _ptr(a, ctype):
    return None if a is None else a.ctypes.data_as(C.POINTER(ctype))


@dataclass
class FlatScene:
    """Flat scene in reference iteration order (geometry major, face minor)."""

    geom_tri_offset: np.ndarray  # int32 [G+1]
    tri_verts: np.ndarray  # float32 [N,3,3]
    tri_normals: np.ndarray | None  # float32 [N,3,3]
    geom_has_normals: np.ndarray  # int32 [G]
    geom_material: np.ndarray  # float32 [G,13]
    light_geom: np.ndarray  # int32 [L]
    sphere_cr: np.ndarray = field(default_factory=lambda: np.zeros((0, 4), np.float32))
    sphere_material: np.ndarray = field(default_factory=lambda: np.zeros((0, 13), np.float32))

    def __post_init__(self):
        self.geom_tri_offset = np.ascontiguousarray(self.geom_tri_offset, np.int32)
        self.tri_verts = np.ascontiguousarray(self.tri_verts, np.float32).reshape(-1, 3, 3)
        if self.tri_normals is not None:
            self.tri_normals = np.ascontiguousarray(self.tri_normals, np.float32).reshape(-1, 3, 3)
        self.geom_has_normals = np.ascontiguousarray(self.geom_has_normals, np.int32)
        self.geom_material = np.ascontiguousarray(self.geom_material, np.float32).reshape(-1, 13)
        self.light_geom = np.ascontiguousarray(self.light_geom, np.int32)
        self.sphere_cr = np.ascontiguousarray(self.sphere_cr, np.float32).reshape(-1, 4)
        self.sphere_material = np.ascontiguousarray(self.sphere_material, np.float32).reshape(-1, 13)

    @property
    def n_geoms(self):
        return len(self.geom_tri_offset) - 1

    @property
    def n_tris(self):
        return int(self.geom_tri_offset[-1])

    @property
    def n_lights(self):
        return len(self.light_geom)

    def save(self, path):
        np.savez_compressed(
            path,
            geom_tri_offset=self.geom_tri_offset,
            tri_verts=self.tri_verts,
            tri_normals=self.tri_normals if self.tri_normals is not None else np.zeros((0, 3, 3), np.float32),
            geom_has_normals=self.geom_has_normals,
            geom_material=self.geom_material,
            light_geom=self.light_geom,
            sphere_cr=self.sphere_cr,
            sphere_material=self.sphere_material,
        )

    @staticmethod
    def load(path):
        z = np.load(path)
        tn = z["tri_normals"]
        return FlatScene(
            z["geom_tri_offset"],
            z["tri_verts"],
            tn if len(tn) else None,
            z["geom_has_normals"],
            z["geom_material"],
            z["light_geom"],
            z["sphere_cr"] if "sphere_cr" in z else np.zeros((0, 4), np.float32),
            z["sphere_material"] if "sphere_material" in z else np.zeros((0, 13), np.float32),
        )


class _RstScene(C.Structure):
    _fields_ = [
        ("n_geoms", C.c_int32),
        ("geom_tri_offset", C.POINTER(C.c_int32)),
        ("tri_verts", C.POINTER(C.c_float)),
        ("tri_normals", C.POINTER(C.c_float)),
        ("geom_has_normals", C.POINTER(C.c_int32)),
        ("geom_material", C.POINTER(C.c_float)),
        ("n_lights", C.c_int32),
        ("light_geom", C.POINTER(C.c_int32)),
        ("n_spheres", C.c_int32),
        ("sphere_cr", C.POINTER(C.c_float)),
        ("sphere_material", C.POINTER(C.c_float)),
    ]


class _RstOutputs(C.Structure):
    _fields_ = [
        ("tri", C.POINTER(C.c_int32)),
        ("t", C.POINTER(C.c_float)),
        ("v", C.POINTER(C.c_float)),
        ("faceid", C.POINTER(C.c_int32)),
        ("occ_tri", C.POINTER(C.c_int32)),
        ("rgb", C.POINTER(C.c_float)),
        ("rgb8", C.POINTER(C.c_uint8)),
        ("n_tests", C.POINTER(C.c_int64)),
    ]


@dataclass
class OracleFrame:
    tri: np.ndarray  # [P] flat triangle index (image index h*W+w), -1 miss
    t: np.ndarray
    v: np.ndarray
    faceid: np.ndarray  # [P,L]
    occ_tri: np.ndarray  # [P,L]
    rgb: np.ndarray  # [P,3] float32
    rgb8: np.ndarray  # [H,W,3] uint8 PPM row order (or [n,3] for pixel lists)
    n_tests: np.ndarray  # [2] primary, shadow


class Restated:
    """The plain-C restatement (oracle/restated.c)."""

    def __init__(self):
        if not os.path.exists(RESTATED_SO):
            build()
        self.lib = C.CDLL(RESTATED_SO)
        L = self.lib
        L.rst_camera.argtypes = [_f32p, _f32p, _f32p, C.c_float, C.c_float, _f32p]
        L.rst_camera.restype = None
        L.rst_render.restype = C.c_int
        L.rst_render_pixels.restype = C.c_int
        L.rst_intersect_triangle.restype = C.c_int
        L.rst_dot.restype = C.c_float
        L.rst_dot.argtypes = [_f32p, _f32p]
        L.rst_cross.argtypes = [_f32p, _f32p, _f32p]

    def set_simd(self, on: bool) -> bool:
        """SURVEY 8f-4 CPU comparator: 8 triangles per step (AVX2), bit-identical to the scalar loops.
        Process-wide; returns the mode in force (False when the CPU has no AVX2)."""
        self.lib.rst_set_simd.restype = C.c_int
        return bool(self.lib.rst_set_simd(int(bool(on))))

    def camera(self, eye, look, W, H, vup=(0, 1, 0), vfov=60.0):
        out = np.zeros(12, np.float32)
        aspect = np.float32(W) / np.float32(H)
        self.lib.rst_camera(
            np.asarray(eye, np.float32), np.asarray(look, np.float32), np.asarray(vup, np.float32),
            C.c_float(vfov), C.c_float(float(aspect)), out,
        )
        return out

    @staticmethod
    def _scene(fs: FlatScene):
        s = _RstScene()
        s.n_geoms = fs.n_geoms
        s.geom_tri_offset = _ptr(fs.geom_tri_offset, C.c_int32)
        s.tri_verts = _ptr(fs.tri_verts, C.c_float)
        s.tri_normals = _ptr(fs.tri_normals, C.c_float)
        s.geom_has_normals = _ptr(fs.geom_has_normals, C.c_int32)
        s.geom_material = _ptr(fs.geom_material, C.c_float)
        s.n_lights = fs.n_lights
        s.light_geom = _ptr(fs.light_geom, C.c_int32)
        s.n_spheres = len(fs.sphere_cr)
        s.sphere_cr = _ptr(fs.sphere_cr, C.c_float)
        s.sphere_material = _ptr(fs.sphere_material, C.c_float)
        return s

    @staticmethod
    def _outputs(n, L, shape8):
        f = OracleFrame(
            tri=np.full(n, -1, np.int32), t=np.zeros(n, np.float32), v=np.zeros(n, np.float32),
            faceid=np.full((n, L), -1, np.int32),
            occ_tri=np.full((n, L), -2, np.int32), rgb=np.zeros((n, 3), np.float32),
            rgb8=np.zeros(shape8, np.uint8), n_tests=np.zeros(2, np.int64),
        )
        o = _RstOutputs()
        o.tri = _ptr(f.tri, C.c_int32)
        o.t = _ptr(f.t, C.c_float)
        o.v = _ptr(f.v, C.c_float)
        o.faceid = _ptr(f.faceid, C.c_int32)
        o.occ_tri = _ptr(f.occ_tri, C.c_int32)
        o.rgb = _ptr(f.rgb, C.c_float)
        o.rgb8 = _ptr(f.rgb8, C.c_uint8)
        o.n_tests = _ptr(f.n_tests, C.c_int64)
        return f, o

    def render(self, fs: FlatScene, cam12, W, H, seed=1, faceid=None, n_threads=None) -> OracleFrame:
        n_threads = n_threads or os.cpu_count() or 1
        s = self._scene(fs)
        f, o = self._outputs(W * H, fs.n_lights, (H, W, 3))
        cam12 = np.ascontiguousarray(cam12, np.float32)
        fid = None if faceid is None else np.ascontiguousarray(faceid, np.int32)
        rc = self.lib.rst_render(C.byref(s), _ptr(cam12, C.c_float), W, H, C.c_uint32(seed), _ptr(fid, C.c_int32),
                                 n_threads, C.byref(o))
        assert rc == 0
        return f

    def render_spp(self, fs: FlatScene, cam12, W, H, seed, spp_n, n_threads=None):
        """extension (parity unpinned): -> (rgb [P,3] float32 image order, rgb8 [H,W,3] PPM order)"""
        n_threads = n_threads or os.cpu_count() or 1
        s = self._scene(fs)
        cam12 = np.ascontiguousarray(cam12, np.float32)
        rgb = np.zeros((W * H, 3), np.float32)
        rgb8 = np.zeros((H, W, 3), np.uint8)
        self.lib.rst_render_spp.restype = C.c_int
        rc = self.lib.rst_render_spp(C.byref(s), _ptr(cam12, C.c_float), W, H, C.c_uint32(seed), spp_n, n_threads,
                                     _ptr(rgb, C.c_float), _ptr(rgb8, C.c_uint8))
        assert rc == 0
        return rgb, rgb8

    def render_pixels(self, fs: FlatScene, cam12, W, H, pw, ph, faceids, n_threads=None) -> OracleFrame:
        n_threads = n_threads or os.cpu_count() or 1
        s = self._scene(fs)
        pw = np.ascontiguousarray(pw, np.int32)
        ph = np.ascontiguousarray(ph, np.int32)
        n = len(pw)
        f, o = self._outputs(n, fs.n_lights, (n, 3))
        o.faceid = None
        cam12 = np.ascontiguousarray(cam12, np.float32)
        fid = np.ascontiguousarray(faceids, np.int32).reshape(n, fs.n_lights)
        f.faceid = fid
        rc = self.lib.rst_render_pixels(C.byref(s), _ptr(cam12, C.c_float), W, H, n, _ptr(pw, C.c_int32),
                                        _ptr(ph, C.c_int32), _ptr(fid, C.c_int32), n_threads, C.byref(o))
        assert rc == 0
        return f

    def replay_faceids(self, fs: FlatScene, W, H, seed, hit_mask):
        s = self._scene(fs)
        hit = np.ascontiguousarray(hit_mask, np.uint8).reshape(-1)
        out = np.full((W * H, fs.n_lights), -1, np.int32)
        self.lib.rst_replay_faceids(C.byref(s), W, H, C.c_uint32(seed), _ptr(hit, C.c_uint8), _ptr(out, C.c_int32))
        return out

    def intersect_triangle(self, orig, d, v0, v1, v2, t):
        tt, u, v = C.c_float(t), C.c_float(0), C.c_float(0)
        a = [np.asarray(x, np.float32) for x in (orig, d, v0, v1, v2)]
        hit = self.lib.rst_intersect_triangle(*[_ptr(x, C.c_float) for x in a], C.byref(tt), C.byref(u), C.byref(v))
        return bool(hit), tt.value, u.value, v.value

    def dot(self, a, b):
        return float(self.lib.rst_dot(np.asarray(a, np.float32), np.asarray(b, np.float32)))

    def cross(self, a, b):
        out = np.zeros(3, np.float32)
        self.lib.rst_cross(np.asarray(a, np.float32), np.asarray(b, np.float32), out)
        return out


def ref_available() -> bool:
    return os.path.exists(REF_SO)


def ref_release_available() -> bool:
    return os.path.exists(REF_RELEASE_SO)


class RefOracle:
    """The unmodified reference serial path (oracle/ref_harness.cpp).  ``release=True`` loads the build with the
    reference's own Release flags (-O3 -ffast-math, cmake/gcc.cmake:16): a TIMING comparator, not a parity oracle."""

    def __init__(self, release: bool = False):
        so = REF_RELEASE_SO if release else REF_SO
        if not os.path.exists(so):
            build()
        if not os.path.exists(so):
            raise FileNotFoundError(so)
        self.release = release
        self.lib = C.CDLL(so)
        L = self.lib
        L.ref_load_obj.restype = C.c_void_p
        L.ref_load_obj.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.ref_from_flat.restype = C.c_void_p
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_counts.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 3
        L.ref_render_pixels.restype = C.c_double
        L.ref_time_rows.restype = C.c_double

    # -- scenes ---------------------------------------------------------------
    def load_obj(self, path):
        err = C.create_string_buffer(1024)
        h = self.lib.ref_load_obj(path.encode(), err, 1024)
        if not h:
            raise RuntimeError(err.value.decode())
        return h

    def from_flat(self, fs: FlatScene):
        assert len(fs.sphere_cr) == 0, "the reference has no spheres"
        tn = fs.tri_normals if fs.tri_normals is not None else np.zeros_like(fs.tri_verts)
        h = self.lib.ref_from_flat(
            C.c_int(fs.n_geoms), _ptr(fs.geom_tri_offset, C.c_int32), _ptr(fs.tri_verts, C.c_float),
            _ptr(tn, C.c_float), _ptr(fs.geom_has_normals, C.c_int32), _ptr(fs.geom_material, C.c_float),
            C.c_int(fs.n_lights), _ptr(fs.light_geom, C.c_int32),
        )
        return C.c_void_p(h)

    def free(self, h):
        self.lib.ref_free(h if isinstance(h, C.c_void_p) else C.c_void_p(h))

    def dump(self, h) -> FlatScene:
        h = h if isinstance(h, C.c_void_p) else C.c_void_p(h)
        g, n, l = C.c_int(), C.c_int(), C.c_int()
        self.lib.ref_counts(h, C.byref(g), C.byref(n), C.byref(l))
        G, N, L = g.value, n.value, l.value
        off = np.zeros(G + 1, np.int32)
        tv = np.zeros((N, 3, 3), np.float32)
        tn = np.zeros((N, 3, 3), np.float32)
        hn = np.zeros(G, np.int32)
        mat = np.zeros((G, 13), np.float32)
        lg = np.zeros(L, np.int32)
        self.lib.ref_dump(h, _ptr(off, C.c_int32), _ptr(tv, C.c_float), _ptr(tn, C.c_float), _ptr(hn, C.c_int32),
                          _ptr(mat, C.c_float), _ptr(lg, C.c_int32))
        return FlatScene(off, tv, tn if hn.any() else None, hn, mat, lg)

    # -- camera / frames --------------------------------------------------------
    def camera(self, eye, look, W, H):
        out = np.zeros(12, np.float32)
        e, l = np.asarray(eye, np.float32), np.asarray(look, np.float32)
        self.lib.ref_camera(_ptr(e, C.c_float), _ptr(l, C.c_float), C.c_int(W), C.c_int(H), _ptr(out, C.c_float))
        return out

    def render_frame(self, h, W, H, eye, look, seed=1):
        """-> (OracleFrame-like dict, replay_exact)"""
        h = h if isinstance(h, C.c_void_p) else C.c_void_p(h)
        g, n, l = C.c_int(), C.c_int(), C.c_int()
        self.lib.ref_counts(h, C.byref(g), C.byref(n), C.byref(l))
        L, P = l.value, W * H
        img = np.zeros((P, 3), np.float32)
        geom = np.zeros(P, np.int32)
        prim = np.zeros(P, np.int32)
        t = np.zeros(P, np.float32)
        v = np.zeros(P, np.float32)
        fid = np.zeros((P, L), np.int32)
        e, lk = np.asarray(eye, np.float32), np.asarray(look, np.float32)
        ok = self.lib.ref_render_frame(h, C.c_int(W), C.c_int(H), _ptr(e, C.c_float), _ptr(lk, C.c_float),
                                       C.c_uint(seed), _ptr(img, C.c_float), _ptr(geom, C.c_int32),
                                       _ptr(prim, C.c_int32), _ptr(t, C.c_float), _ptr(v, C.c_float),
                                       _ptr(fid, C.c_int32))
        q = np.zeros((H, W, 3), np.int32)
        self.lib.ref_quantise(_ptr(img, C.c_float), C.c_int(W), C.c_int(H), _ptr(q, C.c_int32))
        return dict(rgb=img, geom=geom, prim=prim, t=t, v=v, faceid=fid, q=q), bool(ok)

    def render_pixels(self, h, W, H, eye, look, pw, ph, faceids, n_threads=1):
        h = h if isinstance(h, C.c_void_p) else C.c_void_p(h)
        pw = np.ascontiguousarray(pw, np.int32)
        ph = np.ascontiguousarray(ph, np.int32)
        n = len(pw)
        fid = np.ascontiguousarray(faceids, np.int32)
        rgb = np.zeros((n, 3), np.float32)
        geom = np.zeros(n, np.int32)
        prim = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32)
        e, lk = np.asarray(eye, np.float32), np.asarray(look, np.float32)
        secs = self.lib.ref_render_pixels(h, C.c_int(W), C.c_int(H), _ptr(e, C.c_float), _ptr(lk, C.c_float),
                                          C.c_int(n), _ptr(pw, C.c_int32), _ptr(ph, C.c_int32),
                                          _ptr(fid, C.c_int32), _ptr(rgb, C.c_float), _ptr(geom, C.c_int32),
                                          _ptr(prim, C.c_int32), _ptr(t, C.c_float), C.c_int(n_threads))
        return dict(rgb=rgb, geom=geom, prim=prim, t=t, seconds=secs)

    def time_rows(self, h, W, H, eye, look, seed, h_lo, h_hi, threaded):
        h = h if isinstance(h, C.c_void_p) else C.c_void_p(h)
        e, lk = np.asarray(eye, np.float32), np.asarray(look, np.float32)
        return self.lib.ref_time_rows(h, C.c_int(W), C.c_int(H), _ptr(e, C.c_float), _ptr(lk, C.c_float),
                                      C.c_uint(seed), C.c_int(h_lo), C.c_int(h_hi), C.c_int(threaded))
