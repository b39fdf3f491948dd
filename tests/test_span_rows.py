"""CPU checks of the SPAN rows (csrc/kernels.cuh: span_rows; csrc/sweep.cuh: span_terms / span_pass), through the
library's internal host build of the SAME source (tracer__span_rows, tracer__span_pass: no GPU needed).

The default sweeps replace the three affine edge rows p*A_i + q*B_i + C_i >= 0 of a (origin, triangle) by two lower and two
upper bounds of p and test each pair with two saturating adds.  That must stay a NECESSARY condition: whenever the three
exact rows hold at a float (p, q), the float evaluation of the span row must pass — per ray with its own q, and in the hot
loop's shared-q form (mean q + |B| * spread) for every ray within the spread.
"""
import ctypes as C

import numpy as np
import pytest

from esctp1raytracer_b200 import _lib

F32P = C.POINTER(C.c_float)


@pytest.fixture(scope="module")
def span():
    lib = _lib.load()
    lib.tracer__span_rows.argtypes = [C.POINTER(C.c_double), F32P]
    lib.tracer__span_rows.restype = None
    lib.tracer__span_pass.argtypes = [F32P, C.c_float, C.c_float, C.c_float]
    lib.tracer__span_pass.restype = C.c_int

    class Span:
        @staticmethod
        def rows(r):
            r = np.ascontiguousarray(r, np.float64).reshape(9)
            out = np.zeros(8, np.float32)
            lib.tracer__span_rows(r.ctypes.data_as(C.POINTER(C.c_double)), out.ctypes.data_as(F32P))
            return out

        @staticmethod
        def passes(row8, p, q, qdelta=-1.0):
            return bool(lib.tracer__span_pass(row8.ctypes.data_as(F32P), float(p), float(q), float(qdelta)))

    return Span


def exact(rows, p, q):
    """the three rows in exact (here: float64 on float32 inputs, error ~1e-16 relative) arithmetic"""
    return all(float(p) * A + float(q) * B + Cc >= 0 for A, B, Cc in rows)


def triangle_rows(rng, plane_w=(-0.5, -0.5, 1.0), margin=0.0):
    """edge rows of a random triangle seen from the origin through the plane d' = p*ex + q*ey + W, (p,q) in [0,1]^2"""
    ctr = rng.normal(size=3) * 1.5 + np.array([0, 0, rng.uniform(-1, 4)])
    v = ctr + rng.normal(size=(3, 3)) * 10 ** rng.uniform(-3, 0.7)
    a, e1, e2 = v[0], v[1] - v[0], v[2] - v[0]
    Bv, Cv = np.cross(a, e2), np.cross(e1, a)
    Dv = np.cross(e2, e1) - Bv - Cv
    s = 1.0 if np.dot(e2, Cv) > 0 else -1.0
    U, V, W = np.array([1.0, 0, 0]), np.array([0, 1.0, 0]), np.array(plane_w)
    return [(U @ (s * x), V @ (s * x), W @ (s * x) + margin * np.linalg.norm(x)) for x in (Bv, Cv, Dv)]


def test_span_rows_never_lose_a_pair_the_exact_rows_accept(span):
    rng = np.random.default_rng(1)
    n_true = n_pass = n_same_side = 0
    for _ in range(1500):
        rows = triangle_rows(rng)
        sg = [r[0] > 0 for r in rows]
        n_same_side += all(sg) or not any(sg)
        row8 = span.rows(rows)
        for p, q in rng.uniform(0, 1, size=(60, 2)).astype(np.float32):
            ok = exact(rows, p, q)
            sp = span.passes(row8, p, q)
            n_true += ok
            n_pass += sp
            assert sp or not ok, (rows, p, q)
    assert n_same_side > 5      # cones that reach around the plane (three bounds on one side) were exercised
    assert n_true > 500         # ... and so were real hits
    assert n_pass < 1.02 * n_true + 20  # the filter still filters: (almost) nothing but the exact passes gets through


def test_span_rows_on_the_edges(span):
    """points on each edge line, nudged by a few ulps either way: the evaluation margin M must cover the rounding"""
    rng = np.random.default_rng(2)
    n = 0
    for _ in range(800):
        rows = triangle_rows(rng, plane_w=(-0.5, -0.5, rng.uniform(0.5, 2.0)))
        row8 = span.rows(rows)
        for A, B, Cc in rows:
            if abs(A) < 1e-12:
                continue
            for _ in range(8):
                q = np.float32(rng.uniform(0, 1))
                p0 = -(float(q) * B + Cc) / A
                if not 0 <= p0 <= 1:
                    continue
                for k in range(-3, 4):
                    p = np.float32(p0)
                    for _k in range(abs(k)):
                        p = np.nextafter(p, np.float32(2 if k > 0 else -2))
                    if exact(rows, p, q):
                        n += 1
                        assert span.passes(row8, p, q), (rows, p, q)
    assert n > 1000


def test_span_rows_shared_q_form_covers_every_ray_within_the_spread(span):
    """hot loop of the shadow sweeps / jittered primary rays: bounds evaluated at qbar and widened by |B| * qdelta"""
    rng = np.random.default_rng(3)
    n = 0
    for _ in range(1200):
        rows = triangle_rows(rng)
        row8 = span.rows(rows)
        for _ in range(30):
            qbar = np.float32(rng.uniform(0.05, 0.95))
            spread = np.float32(10 ** rng.uniform(-7, -2))
            q = np.float32(qbar + rng.uniform(-1, 1) * spread)
            # the kernels' qdelta: >= |q - qbar| with room for its own roundings (kernels.cuh: shadow_item)
            qdelta = np.float32(abs(np.float32(q - qbar)) * np.float32(1.0001) + np.float32(2.4e-7) * (abs(qbar) + 1))
            p = np.float32(rng.uniform(0, 1))
            if exact(rows, p, q):
                n += 1
                assert span.passes(row8, p, qbar, qdelta), (rows, p, q, qbar, qdelta)
    assert n > 300


def test_span_rows_special_rows(span):
    always = [(0, 0, 1)] * 3
    never = [(0, 0, -1)] * 3
    pts = [(0.0, 0.0), (1.0, 1.0), (0.3, 0.9), (-1.0, 1.0), (1.0625, -1.0625)]
    r = span.rows(always)
    assert all(span.passes(r, p, q) for p, q in pts)
    r = span.rows(never)
    assert not any(span.passes(r, p, q) for p, q in pts)
    # all-zero rows (a zero-size triangle with zero margin) say 0 >= 0: no bound at all
    r = span.rows([(0, 0, 0)] * 3)
    assert all(span.passes(r, p, q) for p, q in pts)
    # non-finite rows: always candidate (the strict path decides)
    for bad in (np.nan, np.inf, -np.inf, 1e300):
        r = span.rows([(bad, 1, 1), (1, 0, 0), (-1, 0, 1)])
        assert all(span.passes(r, p, q) for p, q in pts)
    # rows that do not depend on p at all: q >= 0.25 and q <= 0.75, p in [0.1, 0.9] from the third + an always-true row
    rows = [(0, 1, -0.25), (0, -1, 0.75), (1, 0, -0.1)]
    r = span.rows(rows)
    for p, q in [(0.5, 0.5), (0.1, 0.25), (0.5, 0.75), (1.0, 0.3)]:
        assert span.passes(r, p, q) or not exact(rows, p, q)
    assert not span.passes(r, 0.5, 0.2) and not span.passes(r, 0.5, 0.8) and not span.passes(r, 0.05, 0.5)
    # three lower bounds (cone wraps around the plane): one is dropped, the test only loosens
    rows = [(1, 0.5, -0.2), (1, -0.5, -0.1), (1, 0.1, -0.3)]
    r = span.rows(rows)
    rng = np.random.default_rng(4)
    for p, q in rng.uniform(-1, 1, size=(400, 2)).astype(np.float32):
        assert span.passes(r, p, q) or not exact(rows, p, q)
    # huge coefficients stay finite and conservative
    rows = [(1e-7, 1e12, -3e11), (-1e20, 1.0, 5e19), (3.0, -2.0, 1.0)]
    r = span.rows(rows)
    assert np.all(np.isfinite(r))
    for p, q in rng.uniform(0, 1, size=(400, 2)).astype(np.float32):
        assert span.passes(r, p, q) or not exact(rows, p, q)
