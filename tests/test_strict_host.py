"""csrc/strict_math.cuh — the device's restatement of the reference's arithmetic (src/scene/ray_triangle.h:7-57,
src/math/vec.h:95-139) — compiled for the HOST (tests/strict_host_shim.cpp: plain g++, the correctly-rounded CUDA intrinsics
mapped to the IEEE operations they are defined as) and compared bit for bit with the pinned oracle (oracle/restated.c,
itself bit-equal to the unmodified reference: tests/test_oracle_pin.py).

The GPU parity tests establish the same thing on the device; this test catches an edit of strict_math.cuh that changes
an operation or its order on the CPU, before any GPU run, and covers the epsilon edges of the test (u, v, u+v, t2 within a
few ulps of their bounds, det within FLT_EPSILON of zero) far more densely than rendered frames do.
"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F32 = np.float32
F32P = C.POINTER(C.c_float)
U8P = C.POINTER(C.c_ubyte)
CUDA_INC = "/usr/local/cuda/include"


def fptr(a):
    return a.ctypes.data_as(F32P)


@pytest.fixture(scope="module")
def strict(tmp_path_factory):
    if not shutil.which("g++") or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA headers")
    so = tmp_path_factory.mktemp("strict") / "libstrict_host.so"
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-I" + CUDA_INC,
                        os.path.join(ROOT, "tests", "strict_host_shim.cpp"), "-o", str(so)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return C.CDLL(str(so))


def oracle_pairs(restated, orig, dirs, tris, t_in):
    """rst_intersect_triangle pair by pair -> hit, t, v  (v enters as 0; u is a separate variable there)"""
    rl = restated.lib
    n = len(orig)
    hit, t_out, v_out = np.zeros(n, np.uint8), np.zeros(n, F32), np.zeros(n, F32)
    t, u, v = (np.zeros(1, F32) for _ in range(3))
    for i in range(n):
        t[0], u[0], v[0] = t_in[i], 0, 0
        tri = tris[i]
        hit[i] = rl.rst_intersect_triangle(fptr(orig[i]), fptr(dirs[i]), fptr(tri[0]), fptr(tri[1]), fptr(tri[2]), fptr(t), fptr(u), fptr(v))
        t_out[i], v_out[i] = t[0], v[0]
    return hit, t_out, v_out


def make_pairs(rng, n):
    """rays aimed at the interior / edges / vertices / plane of random triangles, from in front, behind and inside the plane"""
    tris = np.zeros((n, 3, 3), F32)
    orig = np.zeros((n, 3), F32)
    dirs = np.zeros((n, 3), F32)
    t_in = np.zeros(n, F32)
    for i in range(n):
        scale = 10 ** rng.uniform(-3, 1)
        tri = (rng.normal(size=3) * 2 + rng.normal(size=(3, 3)) * scale).astype(F32)
        kind = rng.integers(0, 6)
        if kind == 0:  # needle
            tri[2] = (tri[0] + (tri[1] - tri[0]) * rng.uniform(0.2, 0.8) + rng.normal(size=3) * scale * 1e-4).astype(F32)
        u, v = rng.uniform(0, 1, 2)
        if u + v > 1:
            u, v = 1 - u, 1 - v
        off = rng.choice([0.0, 1e-7, -1e-7, 1.2e-7, 1e-6, -1e-6, 1e-4])
        if kind == 1:
            u = off
        elif kind == 2:
            v = off
        elif kind == 3:
            v = 1 - u + off
        elif kind == 4:
            u, v = rng.choice([0.0, 1.0]) + off, off
        X = tri[0] + u * (tri[1] - tri[0]).astype(np.float64) + v * (tri[2] - tri[0]).astype(np.float64)
        o = (X + rng.normal(size=3) * 10 ** rng.uniform(-7, 1)).astype(F32)
        if kind == 5:  # origin in the triangle's plane: det ~ 0
            nrm = np.cross(tri[1] - tri[0], tri[2] - tri[0]).astype(np.float64)
            nrm /= max(np.linalg.norm(nrm), 1e-30)
            o = (o - nrm * np.dot(o - tri[0], nrm)).astype(F32)
        d = (X - o).astype(F32)
        ln = np.sqrt(F32(F32(F32(d[0] * d[0]) + F32(d[1] * d[1])) + F32(d[2] * d[2])))
        d = (d / ln).astype(F32) if ln > 0 else np.array([0, 0, 1], F32)
        tris[i], orig[i], dirs[i] = tri, o, d
        # t on entry: far away, exactly at / just around the hit distance (the `t2 >= t` tie rule, ray_triangle.h:49), tiny
        t_in[i] = rng.choice([F32(3.4028235e38), ln, np.nextafter(ln, F32(0)), np.nextafter(ln, F32(np.inf)), F32(ln * F32(0.5)), F32(1.1920929e-7)])
    return orig, dirs, tris, t_in


def test_strict_triangle_test_equals_the_oracle_bit_for_bit(strict, restated):
    rng = np.random.default_rng(21)
    n = 120_000
    orig, dirs, tris, t_in = make_pairs(rng, n)
    hit, t_out, v_out = np.zeros(n, np.uint8), np.zeros(n, F32), np.zeros(n, F32)
    strict.strict_host_intersect_triangles(n, fptr(orig), fptr(dirs), fptr(np.ascontiguousarray(tris.reshape(n, 9))), fptr(t_in),
                                           hit.ctypes.data_as(U8P), fptr(t_out), fptr(v_out))
    o_hit, o_t, o_v = oracle_pairs(restated, orig, dirs, tris, t_in)
    assert np.array_equal(hit, o_hit)
    assert np.array_equal(t_out.view(np.uint32), o_t.view(np.uint32))
    assert np.array_equal(v_out.view(np.uint32), o_v.view(np.uint32))
    frac = hit.mean()
    assert 0.15 < frac < 0.85, frac  # both outcomes are well represented ...
    # ... and so are decisions that hang on the last bits: pairs whose verdict flips when t on entry moves by one ulp
    with np.errstate(over="ignore"):
        t_up = np.nextafter(t_in, F32(np.inf))
    hit2 = np.zeros(n, np.uint8)
    strict.strict_host_intersect_triangles(n, fptr(orig), fptr(dirs), fptr(np.ascontiguousarray(tris.reshape(n, 9))), fptr(t_up),
                                           hit2.ctypes.data_as(U8P), fptr(t_out), fptr(v_out))
    assert (hit != hit2).sum() > 500


def test_strict_vector_helpers_equal_the_oracle(strict, restated):
    rng = np.random.default_rng(22)
    n = 20_000
    a = (rng.normal(size=(n, 3)) * 10 ** rng.uniform(-3, 3, (n, 1))).astype(F32)
    b = (rng.normal(size=(n, 3)) * 10 ** rng.uniform(-3, 3, (n, 1))).astype(F32)
    dot, cross, norm, ln = np.zeros(n, F32), np.zeros((n, 3), F32), np.zeros((n, 3), F32), np.zeros(n, F32)
    strict.strict_host_vec(n, fptr(a), fptr(b), fptr(dot), fptr(cross), fptr(norm), fptr(ln))
    rl = restated.lib
    out = np.zeros(3, F32)
    for i in range(0, n, 7):
        assert F32(rl.rst_dot(a[i], b[i])).view(np.uint32) == dot[i].view(np.uint32)  # vec.h:95-101 (ndpointer argtypes)
        rl.rst_cross(a[i], b[i], out)                                                                # vec.h:103-109
        assert np.array_equal(out.view(np.uint32), cross[i].view(np.uint32))
    # vec.h:135-139: length = sqrtf(dot(a, a)), normalize = a / length, in float
    d = F32(0)
    for k in range(3):
        d = (d + (a[:, k] * a[:, k]).astype(F32)).astype(F32)
    assert np.array_equal(np.sqrt(d).astype(F32).view(np.uint32), ln.view(np.uint32))
    assert np.array_equal((a / np.sqrt(d).astype(F32)[:, None]).astype(F32).view(np.uint32), norm.view(np.uint32))
    # the four identities of src/ispc/test.ispc:24-37, the reference's only known-answer vectors
    x, y = np.array([[1, 0, 1]], F32), np.array([[1, 2, 3]], F32)
    strict.strict_host_vec(1, fptr(x), fptr(y), fptr(dot), fptr(cross), fptr(norm), fptr(ln))
    assert dot[0] == 4 and list(cross[0]) == [-2, -2, 2]
