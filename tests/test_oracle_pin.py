"""Pin the oracle: the plain-C restatement must equal the reference bit for bit.

* against the committed golden fixtures (frames rendered by the reference's own
  scan_row through oracle/ref_harness.cpp — see tests/golden/make_golden.py);
* against oracle/_ref live, when it is present, on more scenes / sizes;
* the four vector identities of src/ispc/test.ispc:24-37 (the reference's only
  known-answer vectors) and the epsilon edges of ray_triangle.h:23-50.
"""
import os

import numpy as np
import pytest
from conftest import bits, golden_names, load_golden, to_flat


@pytest.mark.parametrize("name", golden_names())
def test_restated_equals_reference_golden(restated, name):
    fs, fr = load_golden(name)
    W, H = fr["W"], fr["H"]
    cam = restated.camera(fr["eye"], fr["look"], W, H)
    assert np.array_equal(bits(cam), bits(fr["cam"]))  # camera.h:16-29
    o = restated.render(fs, cam, W, H, seed=fr["seed"])  # own mt19937 replay
    assert np.array_equal(o.tri, fr["tri"])
    assert np.array_equal(bits(o.t), bits(fr["t"]))
    assert np.array_equal(bits(o.v), bits(fr["v"]))
    assert np.array_equal(o.faceid, fr["faceid"])
    assert np.array_equal(bits(o.rgb), bits(fr["rgb"]))  # all three float channels, every pixel
    assert np.array_equal(o.rgb8, fr["q"])
    # explicit faceids give the same frame
    o2 = restated.render(fs, cam, W, H, faceid=fr["faceid"])
    assert np.array_equal(bits(o2.rgb), bits(fr["rgb"]))


def test_restated_pixel_subset_matches_frame(restated):
    fs, fr = load_golden("cornell_original_3lights")
    W, H = fr["W"], fr["H"]
    rng = np.random.default_rng(0)
    idx = rng.choice(W * H, 200, replace=False)
    ph, pw = idx // W, idx % W
    o = restated.render_pixels(fs, fr["cam"], W, H, pw, ph, fr["faceid"][idx])
    assert np.array_equal(bits(o.rgb), bits(fr["rgb"][idx]))
    assert np.array_equal(o.tri, fr["tri"][idx])


LIVE = [
    ("cornell/CornellBox-Original.obj", (0, 1, 2), (0, 1, 0), 200, 150, 1),
    ("cornell/CornellBox-Mirror.obj", (0, 1, 2), (0, 1, 0), 80, 60, 2),
    ("cornell/CornellBox-Empty-RG.obj", (0.2, 0.7, 2.2), (0, 1, 0), 80, 60, 3),
    ("cornell/CornellBox-Empty-Squashed.obj", (0, 1, 2), (0, 1, 0), 80, 60, 4),
    ("cornell_box.obj", (0, 1, 3), (0, 1, 0), 120, 90, 5),
    ("cornell/CornellBox-Sphere.obj", (0.5, 1.2, 1.8), (0, 0.8, 0), 48, 36, 6),
    ("cornell/water.obj", (0, 1, 2), (0, 1, 0), 32, 24, 7),  # no light: black image
]


@pytest.mark.parametrize("model,eye,look,W,H,seed", LIVE)
def test_restated_equals_reference_live(restated, ref_oracle, model, eye, look, W, H, seed):
    path = "/root/reference/src/models/" + model
    if not os.path.exists(path):
        pytest.skip("reference models not present")
    h = ref_oracle.load_obj(path)
    try:
        fs = ref_oracle.dump(h)
        fr, exact = ref_oracle.render_frame(h, W, H, eye, look, seed=seed)
        assert exact, "mt19937 replay must end in scan_row's generator state"
        cam = restated.camera(eye, look, W, H)
        assert np.array_equal(bits(cam), bits(ref_oracle.camera(eye, look, W, H)))
        o = restated.render(fs, cam, W, H, seed=seed)
        assert np.array_equal(o.faceid, fr["faceid"])
        assert np.array_equal(bits(o.rgb), bits(fr["rgb"]))
        assert np.array_equal(bits(o.t), bits(fr["t"])) and np.array_equal(bits(o.v), bits(fr["v"]))
        assert np.array_equal(o.rgb8.astype(np.int32), fr["q"])
    finally:
        ref_oracle.free(h)


def test_restated_equals_reference_synthetic_multilight(restated, ref_oracle):
    """4 lights, normals on some geometries, specular: the t carry between lights (main.cpp:764)."""
    from esctp1raytracer_b200 import scenes

    s = scenes.soup_scene(3000, 12, 4, seed=3, edge=(0.05, 0.2), specular=True, with_normals=True)
    fs = to_flat(s)
    W, H, eye, look = 64, 48, (0, 1, 3), (0, 1, 0)
    h = ref_oracle.from_flat(fs)
    try:
        fr, exact = ref_oracle.render_frame(h, W, H, eye, look, seed=12)
        assert exact
        o = restated.render(fs, restated.camera(eye, look, W, H), W, H, seed=12)
        assert np.array_equal(o.faceid, fr["faceid"])
        assert np.array_equal(bits(o.rgb), bits(fr["rgb"]))
        # glue-driven pixel subset of the harness == scan_row
        idx = np.arange(0, W * H, 7)
        sub = ref_oracle.render_pixels(h, W, H, eye, look, idx % W, idx // W, fr["faceid"][idx])
        assert np.array_equal(bits(sub["rgb"]), bits(fr["rgb"][idx]))
    finally:
        ref_oracle.free(h)


def test_vector_kats(restated):
    # src/ispc/test.ispc:24-37: a=[1 0 1], b=[1 2 3]
    a, b = (1, 0, 1), (1, 2, 3)
    assert restated.dot(a, b) == 4.0
    assert np.array_equal(restated.cross(a, b), np.array([-2, -2, 2], np.float32))


def test_intersect_triangle_edges(restated):
    v0, v1, v2 = (0, 0, 0), (1, 0, 0), (0, 1, 0)
    d = (0, 0, -1)
    FMAX = np.finfo(np.float32).max
    eps = np.finfo(np.float32).eps
    hit, t, u, v = restated.intersect_triangle((0.25, 0.25, 1), d, v0, v1, v2, FMAX)
    assert hit and t == 1.0 and u == 0.25 and v == 0.25
    # two-sided: from below as well (no culling, ray_triangle.h:23-25)
    assert restated.intersect_triangle((0.25, 0.25, -1), (0, 0, 1), v0, v1, v2, FMAX)[0]
    # u exactly 0 is rejected (u2 < eps), u = 1 - v accepted (u2+v2 > 1 rejects only above 1)
    assert not restated.intersect_triangle((0.0, 0.25, 1), d, v0, v1, v2, FMAX)[0]
    assert restated.intersect_triangle((0.75, 0.25, 1), d, v0, v1, v2, FMAX)[0]
    assert not restated.intersect_triangle((0.75, 0.2500001, 1), d, v0, v1, v2, FMAX)[0]
    # t2 >= t rejects ties (first index wins, ray_triangle.h:49); t2 < eps rejects
    assert not restated.intersect_triangle((0.25, 0.25, 1), d, v0, v1, v2, 1.0)[0]
    assert not restated.intersect_triangle((0.25, 0.25, eps / 2), d, v0, v1, v2, FMAX)[0]
    # parallel ray: |det| < eps
    assert not restated.intersect_triangle((0.25, 0.25, 1), (1, 0, 0), v0, v1, v2, FMAX)[0]


def test_replay_consumes_three_draws_per_hit_light(restated):
    fs, fr = load_golden("cornell_box_ks")  # about half the pixels miss
    fid = restated.replay_faceids(fs, fr["W"], fr["H"], fr["seed"], fr["tri"] >= 0)
    assert np.array_equal(fid, fr["faceid"])
    assert (fid[fr["tri"] < 0] == -1).all()


def test_spp_extension_reduces_to_single_sample(restated):
    """rst_render_spp (unpinned extension) with 1 sample == the pinned path fed with the hash faceIDs"""
    from esctp1raytracer_b200 import Camera, hash_faceids, scenes

    s = scenes.box_scene()
    W, H, seed = 48, 36, 5
    cam = Camera.for_frame((0, 1, 2.9), (0, 1, 0), W, H).as_array()
    rgb, rgb8 = restated.render_spp(to_flat(s), cam, W, H, seed, 1)
    o = restated.render(to_flat(s), cam, W, H, faceid=hash_faceids(seed, W, H, s.faces_per_light))
    assert np.array_equal(bits(rgb), bits(o.rgb)) and np.array_equal(rgb8, o.rgb8)
    rgb4, _ = restated.render_spp(to_flat(s), cam, W, H, seed, 2)
    assert not np.array_equal(rgb4, rgb) and abs(float(rgb4.mean()) - float(rgb.mean())) < 0.05


def test_simd_comparator_is_bit_identical_to_the_scalar_restatement(restated):
    """the 8-triangles-per-step AVX2 form of the restatement (the CPU comparator for boxes without ispc)
    must reproduce the scalar loops bit for bit: ids, t, v, occluders, float accumulator, u8 — on the reference's
    golden models (index-order ties, self-shadow chaos) and on a multi-light soup with normals, specular and spheres."""
    from esctp1raytracer_b200 import Camera, hash_faceids, scenes

    if not restated.set_simd(True):
        restated.set_simd(False)
        pytest.skip("no AVX2 on this CPU")
    try:
        cases = []
        for name in golden_names():
            fs, fr = load_golden(name)
            cases.append((fs, fr["cam"], fr["W"], fr["H"], fr["faceid"]))
        s = scenes.soup_scene(5003, 17, 3, seed=11, with_normals=True, specular=True, n_spheres=7)
        W, H = 40, 30
        cases.append((to_flat(s), Camera.for_frame((0, 1, 3), (0, 1, 0), W, H).as_array(), W, H, hash_faceids(9, W, H, s.faces_per_light)))
        for fs, cam, W, H, fid in cases:
            restated.set_simd(True)
            a = restated.render(fs, cam, W, H, faceid=fid)
            restated.set_simd(False)
            b = restated.render(fs, cam, W, H, faceid=fid)
            assert np.array_equal(a.tri, b.tri) and np.array_equal(bits(a.t), bits(b.t)) and np.array_equal(bits(a.v), bits(b.v))
            assert np.array_equal(a.occ_tri, b.occ_tri) and np.array_equal(bits(a.rgb), bits(b.rgb)) and np.array_equal(a.rgb8, b.rgb8)
            assert np.array_equal(a.n_tests, b.n_tests)
    finally:
        restated.set_simd(False)
