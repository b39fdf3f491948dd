"""CPU checks of the COMPLETE filter-row construction (csrc/kernels.cuh: origin_rows = the body of build_origin_table,
margins K / K2, sign-ambiguous rows, span rows) through the library's host build of the SAME source
(tracer__origin_rows, tracer__filter_pass: no GPU needed), against the reference's own intersection test
(src/scene/ray_triangle.h:7-57 as restated in oracle/restated.c, pinned to the reference bit for bit).

The sweeps are only correct if the filter is a NECESSARY condition: every (ray, triangle) pair the reference's float /
double test accepts must pass the three-row test, the span test with the ray's own q, and the span test in the hot
loop's shared-q form.  On the GPU this is what `exhaustive_strict` + `filter_misses == 0` assert; here the same
property is checked on the CPU, with rays built exactly as the device builds them:

* primary rays: (p, q) = the reference's (s, t) (src/main.cpp:709-710), direction by camera::get_ray
  (src/scene/camera.h:31-34) in float arithmetic, table on the image plane (tracer_cuda.cu, "eye table");
* shadow rays: origin = a hit point, direction = normalize(light vertex - hit) (src/main.cpp:757-766), (p, q) on the cube
  face of the dominant axis of (hit - light vertex) (kernels.cuh: light_step_kernel), table of that face (face_param).

Rays are aimed at the interior, the edges and the vertices of each triangle (and a few ulps either side), which is where
a too-tight margin would lose an accepted pair.
"""
import ctypes as C

import numpy as np
import pytest

from esctp1raytracer_b200 import _lib

F32 = np.float32
F32P = C.POINTER(C.c_float)
EPS = F32(1.1920929e-7)
FLT_MAX = F32(3.4028235e38)


def fptr(a):
    return a.ctypes.data_as(F32P)


@pytest.fixture(scope="module")
def flt(restated):
    lib = _lib.load()
    lib.tracer__origin_rows.argtypes = [F32P, C.POINTER(C.c_double), F32P]
    lib.tracer__origin_rows.restype = None
    lib.tracer__filter_pass.argtypes = [F32P, C.c_int, F32P, F32P, C.c_float, C.c_float, C.POINTER(C.c_ubyte)]
    lib.tracer__filter_pass.restype = None
    rl = restated.lib  # (shared with other tests: no argtypes are set on it here, pointers are passed explicitly)

    class Flt:
        @staticmethod
        def rows(tri, o, U, V, W, dmax, lmax):
            tri = np.ascontiguousarray(tri, F32).reshape(9)
            tp = np.ascontiguousarray(np.concatenate([o, U, V, W, [dmax, lmax]]), np.float64)
            out = np.zeros(20, F32)
            lib.tracer__origin_rows(fptr(tri), tp.ctypes.data_as(C.POINTER(C.c_double)), fptr(out))
            return out

        @staticmethod
        def passes(rows20, p, q, qbar=0.0, qdelta=0.0):
            p, q = np.ascontiguousarray(p, F32), np.ascontiguousarray(q, F32)
            out = np.zeros(len(p), np.uint8)
            lib.tracer__filter_pass(fptr(rows20), len(p), fptr(p), fptr(q), float(qbar), float(qdelta),
                                    out.ctypes.data_as(C.POINTER(C.c_ubyte)))
            return out

        @staticmethod
        def camera(eye, look, aspect):
            from esctp1raytracer_b200 import Camera

            out = np.asarray(Camera(eye, look, (0, 1, 0), 60.0, aspect).as_array(), F32).reshape(12)  # tracer_camera_lookat
            return out[0:3].copy(), out[3:6].copy(), out[6:9].copy(), out[9:12].copy()

        @staticmethod
        def accepts(orig, dirs, tri, t0):
            """the reference's test, ray by ray: -> bool [n].  t0: the ray's t on entry (scalar or [n])"""
            tri = np.ascontiguousarray(tri, F32).reshape(3, 3)
            orig = np.ascontiguousarray(np.broadcast_to(orig, dirs.shape), F32)
            dirs = np.ascontiguousarray(dirs, F32)
            t0 = np.ascontiguousarray(np.broadcast_to(t0, (len(dirs),)), F32)
            out = np.zeros(len(dirs), bool)
            t, u, v = (np.zeros(1, F32) for _ in range(3))
            v0, v1, v2 = (np.ascontiguousarray(tri[i]) for i in range(3))
            for i in range(len(dirs)):
                t[0] = t0[i]
                out[i] = rl.rst_intersect_triangle(fptr(orig[i]), fptr(dirs[i]), fptr(v0), fptr(v1), fptr(v2), fptr(t), fptr(u), fptr(v))
            return out

    return Flt


def f32_dot(a, b):
    """vec.h:95-101: the sum starts at 0 and adds x, y, z in order, every step rounded to float"""
    s = F32(0)
    for k in range(3):
        s = F32(s + F32(a[..., k] * b[..., k]))
    return s


def f32_normalize(d):
    """vec.h:135-139: d / sqrtf(dot(d, d)), three divides"""
    n = np.sqrt(f32_dot(d, d)).astype(F32)
    return (d / n[..., None]).astype(F32)


def aim_points(rng, tri, n):
    """points of the triangle's plane: interior, on / just off the edges, at / around the vertices"""
    v0, e1, e2 = tri[0].astype(np.float64), (tri[1] - tri[0]).astype(np.float64), (tri[2] - tri[0]).astype(np.float64)
    kind = rng.integers(0, 4, n)
    u, v = rng.uniform(0, 1, n), rng.uniform(0, 1, n)
    fold = u + v > 1
    u[fold], v[fold] = 1 - u[fold], 1 - v[fold]
    off = rng.choice([0.0, 1e-7, -1e-7, 1e-6, -1e-6, 1e-5, -1e-5, 1e-3, -1e-3], n)
    e = kind == 1  # on the edges u = 0, v = 0, u + v = 1
    which = rng.integers(0, 3, n)
    u = np.where(e & (which == 0), off, u)
    v = np.where(e & (which == 1), off, v)
    v = np.where(e & (which == 2), 1 - u + off, v)
    c = kind == 2  # around the vertices
    cu, cv = rng.choice([0.0, 1.0], n), rng.choice([0.0, 1.0], n)
    cv = np.where(cu == 1, 0.0, cv)
    u = np.where(c, cu + off, u)
    v = np.where(c, cv + rng.choice([0.0, 1e-7, -1e-7, 1e-5, -1e-5], n), v)
    return v0 + u[:, None] * e1 + v[:, None] * e2


def random_triangle(rng, centre, scale, needle=False):
    v = centre + rng.normal(size=(3, 3)) * scale
    if needle:  # long and thin
        v[2] = v[0] + (v[1] - v[0]) * rng.uniform(0.3, 0.7) + rng.normal(size=3) * scale * 1e-3
    return v.astype(F32)


def qbar_form(rng, q):
    """a shared q-term as shadow_item (kernels.cuh) builds it: the mid q of rays up to a pixel-ish spread apart"""
    spread = F32(rng.choice([0.0, 1e-7, 1e-6, 1e-4])) * F32(rng.uniform(0, 1))
    other = F32(q + F32(rng.choice([-1.0, 1.0])) * spread)
    qmin, qmax = min(F32(q), other), max(F32(q), other)
    qbar = F32(0.5) * F32(qmin + qmax)
    qdelta = F32(F32(max(F32(qmax - qbar), F32(qbar - qmin))) * F32(1.0001) + F32(2.4e-7) * F32(abs(qbar) + F32(1)))
    return qbar, qdelta


def test_eye_table_rows_never_lose_a_primary_ray_the_reference_accepts(flt):
    rng = np.random.default_rng(11)
    n_acc = n_pairs = n_pass = n_ambiguous = 0
    for it in range(260):
        eye = np.array([0, 1, 3], F32) + rng.normal(size=3).astype(F32) * F32(0.3)
        look = np.array([0, 1, 0], F32) + rng.normal(size=3).astype(F32) * F32(0.2)
        aspect = F32(rng.choice([3840 / 2160, 1024 / 768, 1.0]))
        o, llc, hor, ver = flt.camera(eye, look, aspect)
        U, V, Wv = hor.astype(np.float64), ver.astype(np.float64), llc.astype(np.float64) - o.astype(np.float64)
        dmax = max(np.linalg.norm(Wv + a * U + b * V) for a in (0, 1) for b in (0, 1))
        # a triangle somewhere in the view: towards a random point of the image plane, at a random distance
        s, t = rng.uniform(0.05, 0.95, 2)
        dist = 10 ** rng.uniform(-0.5, 1.3)
        centre = o + (Wv + s * U + t * V) / np.linalg.norm(Wv + s * U + t * V) * dist
        scale = dist * 10 ** rng.uniform(-3.5, -0.3)
        tri = random_triangle(rng, centre, scale, needle=it % 5 == 0)
        if it % 13 == 0:  # the eye (nearly) in the triangle's plane: sign-ambiguous rows
            nrm = np.cross(tri[1] - tri[0], tri[2] - tri[0]).astype(np.float64)
            nrm /= max(np.linalg.norm(nrm), 1e-30)
            shift = np.dot(o - tri[0], nrm) * (1 - rng.choice([0.0, 1e-7, 1e-5]))
            tri = (tri + (nrm * shift)[None, :]).astype(F32)
            n_ambiguous += 1
        rows = flt.rows(tri, o, U, V, Wv, dmax, 0.0)
        # (p, q) of rays through the aim points: p*U + q*V + W = lambda * (X - o)
        X = aim_points(rng, tri, 260)
        M = np.stack([np.broadcast_to(U, X.shape), np.broadcast_to(V, X.shape), -(X - o)], axis=2)
        try:
            sol = np.linalg.solve(M, np.broadcast_to(-Wv, X.shape)[..., None])[..., 0]
        except np.linalg.LinAlgError:
            continue
        p = sol[:, 0].astype(F32)
        q = sol[:, 1].astype(F32)
        keep = (sol[:, 2] > 0) & (p >= 0) & (p <= 1) & (q >= 0) & (q <= 1)
        p, q = p[keep], q[keep]
        if len(p) == 0:
            continue
        # nudge by a few ulps as well: the grid values w/(W-1) are arbitrary floats in [0, 1]
        p = np.clip(np.nextafter(p, F32(rng.choice([-1, 2]))), 0, 1).astype(F32)
        # camera::get_ray in float (camera.h:31-34): normalize(((llc + hor*s) + ver*t) - origin)
        d = ((llc[None, :] + (hor[None, :] * p[:, None]).astype(F32)).astype(F32) + (ver[None, :] * q[:, None]).astype(F32)).astype(F32)
        d = f32_normalize((d - o[None, :]).astype(F32))
        acc = flt.accepts(o, d, tri, FLT_MAX)
        res = flt.passes(rows, p, q)
        assert np.all((res[acc] & 3) == 3), (it, tri, o, p[acc][(res[acc] & 3) != 3], q[acc][(res[acc] & 3) != 3])
        for i in np.nonzero(acc)[0][:6]:  # the hot loop's shared-q form for jittered samples
            qbar, qdelta = qbar_form(rng, q[i])
            assert flt.passes(rows, p[i:i + 1], q[i:i + 1], qbar, qdelta)[0] & 4, (it, tri, p[i], q[i], qbar, qdelta)
        n_acc += int(acc.sum())
        n_pairs += len(p)
        # selectivity on rays that are nowhere near the triangle
        pr, qr = rng.uniform(0, 1, 64).astype(F32), rng.uniform(0, 1, 64).astype(F32)
        n_pass += int(((flt.passes(rows, pr, qr) & 2) != 0).sum())
    assert n_acc > 8000 and n_pairs > 2 * n_acc / 2 and n_ambiguous >= 15
    assert n_pass < 0.2 * 260 * 64  # the filter still filters (big triangles do cover part of the frame)


def shadow_setup(hit, lv):
    """main.cpp:757-766 and light_step_kernel: -> dir, t, face, (p, q), len   (all float32, per ray)"""
    Lv = (lv[None, :] - hit).astype(F32)
    ln = np.sqrt(f32_dot(Lv, Lv)).astype(F32)
    t = (ln - EPS).astype(F32)
    d = f32_normalize(Lv)
    fd = (hit - lv[None, :]).astype(F32)
    a = np.abs(fd)
    c = np.where((a[:, 0] >= a[:, 1]) & (a[:, 0] >= a[:, 2]), 0, np.where(a[:, 1] >= a[:, 2], 1, 2))
    idx = np.arange(len(hit))
    dc = fd[idx, c]
    inv = (F32(1) / np.abs(dc)).astype(F32)
    p = (fd[idx, (c + 1) % 3] * inv).astype(F32)
    q = (fd[idx, (c + 2) % 3] * inv).astype(F32)
    face = 2 * c + (dc < 0)
    return d, t, face, p, q, ln


def face_param(lv, f):
    """tracer_cuda.cu: face_param"""
    c, a, b = f // 2, (f // 2 + 1) % 3, (f // 2 + 2) % 3
    U, V, W = np.zeros(3), np.zeros(3), np.zeros(3)
    U[a], V[b], W[c] = 1.0, 1.0, (-1.0 if f & 1 else 1.0)
    return lv.astype(np.float64), U, V, W, np.sqrt(3.0)


def test_light_tables_never_lose_a_shadow_ray_the_reference_accepts(flt):
    rng = np.random.default_rng(12)
    n_acc = n_rays = n_own = 0
    for it in range(260):
        lv = (np.array([0, 1.9, 0]) + rng.normal(size=3) * 0.5).astype(F32)  # the light vertex: light.vertex[faceID]
        dist = 10 ** rng.uniform(-1.0, 0.8)
        dirn = rng.normal(size=3)
        dirn /= np.linalg.norm(dirn)
        centre = lv + dirn * dist
        scale = dist * 10 ** rng.uniform(-3.0, -0.2)
        tri = random_triangle(rng, centre, scale, needle=it % 5 == 0)
        if it % 11 == 0:  # the light's own face: the light vertex IS a vertex of the triangle
            tri[rng.integers(0, 3)] = lv
            n_own += 1
        elif it % 11 == 1:  # the light vertex in the triangle's plane
            nrm = np.cross(tri[1] - tri[0], tri[2] - tri[0]).astype(np.float64)
            nrm /= max(np.linalg.norm(nrm), 1e-30)
            tri = (tri + (nrm * np.dot(lv - tri[0], nrm))[None, :]).astype(F32)
        # hit points behind the triangle as seen from the light (the segment hit -> light crosses the aim point), and
        # a few in front of it (t2 >= t must reject, whatever the filter says)
        X = aim_points(rng, tri, 240)
        alpha = np.where(rng.uniform(size=len(X)) < 0.9, 1 + 10 ** rng.uniform(-4, 0.5, len(X)), rng.uniform(0.2, 0.999, len(X)))
        hit = (lv[None, :] + (X - lv[None, :]) * alpha[:, None]).astype(F32)
        ok = np.linalg.norm(hit - lv[None, :], axis=1) > 1e-6
        hit = hit[ok]
        d, t, face, p, q, ln = shadow_setup(hit, lv)
        lmax = float(ln.max()) * 1.5  # the host passes (k+1)*diag*1.5 >= every ray's length
        acc = flt.accepts(hit, d, tri, t)
        for f in np.unique(face):
            sel = np.nonzero(face == f)[0]
            rows = flt.rows(tri, *face_param(lv, int(f)), lmax)
            res = flt.passes(rows, p[sel], q[sel])
            bad = sel[acc[sel] & ((res & 3) != 3)]
            assert len(bad) == 0, (it, f, tri, lv, hit[bad], p[bad], q[bad])
            for i in sel[acc[sel]][:5]:  # the hot loop's form: mean q of the thread's rays + |B| * spread
                qbar, qdelta = qbar_form(rng, q[i])
                assert flt.passes(rows, p[i:i + 1], q[i:i + 1], qbar, qdelta)[0] & 4, (it, f, tri, lv, hit[i], qbar, qdelta)
        n_acc += int(acc.sum())
        n_rays += len(hit)
    assert n_acc > 6000 and n_rays > n_acc and n_own >= 20


def test_padding_and_degenerate_triangles(flt):
    o, U, V, W = np.zeros(3), np.array([1.0, 0, 0]), np.array([0, 1.0, 0]), np.array([-0.5, -0.5, -1.0])
    p, q = np.linspace(0, 1, 33, dtype=F32), np.full(33, 0.5, F32)
    # zero-area triangles (a point, a segment): the reference's det is 0 -> never accepted; the rows must simply be finite
    # or "always candidate", never NaN-poisoned into rejecting a neighbour
    for tri in (np.zeros((3, 3)), np.array([[0, 0, -2], [1, 0, -2], [2, 0, -2]]), np.array([[0, 0, -2]] * 3)):
        rows = flt.rows(tri, o, U, V, W, 1.3, 0.0)
        assert np.all(np.isfinite(rows))
        flt.passes(rows, p, q)
    # coordinates near the float range: non-finite intermediates make the triangle a candidate, the strict path decides
    rows = flt.rows(np.array([[1e30, 0, -1e30], [0, 1e30, -1e30], [-1e30, -1e30, -1e30]]), o, U, V, W, 1.3, 0.0)
    assert np.all(np.isfinite(rows))
    assert np.all(flt.passes(rows, p, q) & 2)
