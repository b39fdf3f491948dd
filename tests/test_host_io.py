"""§8f rows 1-2: the OBJ/MTL loader and the PPM writer (host side, C++ in csrc/host_io.cpp).

Loader parity: the flat scene must equal, bit for bit, the dump of the reference's own
model::loadobj (through oracle/_ref) on every shipped model, and must fail where the reference
throws.  PPM writer: byte-equal to the format of src/main.cpp:661-685."""
import os

import numpy as np
import pytest
from conftest import bits

from esctp1raytracer_b200 import Scene, TracerError, write_ppm

MODELS = "/root/reference/src/models/"
LOADABLE = ["cornell_box.obj", "cornell/CornellBox-Original.obj", "cornell/CornellBox-Mirror.obj",
            "cornell/CornellBox-Empty-CO.obj", "cornell/CornellBox-Empty-RG.obj", "cornell/CornellBox-Empty-White.obj",
            "cornell/CornellBox-Empty-Squashed.obj", "cornell/CornellBox-Sphere.obj", "cornell/CornellBox-Water.obj",
            "cornell/water.obj"]
MUST_FAIL = ["cornell/CornellBox-Glossy.obj", "cornell/CornellBox-Glossy-Floor.obj"]


@pytest.mark.parametrize("model", LOADABLE)
def test_loader_equals_reference_loader(ref_oracle, model):
    path = MODELS + model
    if not os.path.exists(path):
        pytest.skip("reference models not present")
    h = ref_oracle.load_obj(path)
    try:
        want = ref_oracle.dump(h)
    finally:
        ref_oracle.free(h)
    got = Scene.load_obj(path)
    assert np.array_equal(got.geom_tri_offset, want.geom_tri_offset)
    assert np.array_equal(bits(got.tri_verts), bits(want.tri_verts))
    assert np.array_equal(got.geom_has_normals, want.geom_has_normals)
    assert np.array_equal(bits(got.geom_material), bits(want.geom_material))
    assert np.array_equal(got.light_geom, want.light_geom)
    if want.tri_normals is not None:
        assert np.array_equal(bits(got.tri_normals), bits(want.tri_normals))


@pytest.mark.parametrize("model", MUST_FAIL)
def test_loader_fails_where_reference_throws(ref_oracle, model):
    path = MODELS + model
    if not os.path.exists(path):
        pytest.skip("reference models not present")
    with pytest.raises(RuntimeError):
        ref_oracle.load_obj(path)  # warnings are fatal (sceneloader.cpp:27-30)
    with pytest.raises(TracerError):
        Scene.load_obj(path)


def test_loader_semantics_on_a_handwritten_obj(tmp_path):
    """usemtl inside a shape keeps the FIRST material; g/o split shapes; polygons fan out; negative and
    i//k indices; exponent and odd number spellings; CRLF."""
    (tmp_path / "m.mtl").write_text("newmtl a\nKa 0.1 0.2 0.3\nKd 1 0.5 0.25\nNs 7\n\nnewmtl b\nKd 0 1 0\nKe 2 2 2\n")
    (tmp_path / "s.obj").write_text(
        "mtllib m.mtl\r\n"
        "v 0 0 0\r\nv 1 0 0\r\nv 1 1 0\r\nv 0 1 0\r\nv 1e-1 2.5E+1 -3.25e0\r\nv +4 .5 7junk\r\n"
        "vn 0 0 2\r\n"
        "g first\r\nusemtl a\r\nf 1 2 3 4\r\nusemtl b\r\nf -6//1 -5//1 -4//1\r\n"
        "o second\r\nusemtl b\r\nf 1/1/1 3/1/1 5/1/1\r\n")
    s = Scene.load_obj(str(tmp_path / "s.obj"))
    assert list(s.geom_tri_offset) == [0, 3, 4]           # quad -> 2 triangles + 1, then 1
    assert np.allclose(s.geom_material[0][:6], [0.1, 0.2, 0.3, 1, 0.5, 0.25]) and s.geom_material[0][12] == 7
    assert list(s.light_geom) == [1]                        # only "second" uses the emissive material first
    assert np.array_equal(s.tri_verts[1], np.array([[0, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32))  # fan (0,2,3)
    assert np.array_equal(s.tri_verts[2], np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0]], np.float32))  # -6,-5,-4 of 6
    assert np.array_equal(s.tri_verts[3][2], np.array([0.1, 25.0, -3.25], np.float32))
    assert list(s.geom_has_normals) == [1, 1] or list(s.geom_has_normals) == [1, 1]
    assert np.array_equal(s.tri_normals[3][0], np.array([0, 0, 1], np.float32))  # normalised
    with pytest.raises(TracerError):
        (tmp_path / "t.obj").write_text("mtllib missing.mtl\nv 0 0 0\n")
        Scene.load_obj(str(tmp_path / "t.obj"))


def test_ppm_writer_matches_reference_format(tmp_path):
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(7, 5, 3), dtype=np.uint8)
    p3, p6 = tmp_path / "a.ppm", tmp_path / "b.ppm"
    write_ppm(str(p3), img)
    want = "P3\n5 7\n255\n" + "".join(f"{r} {g} {b}\n" for r, g, b in img.reshape(-1, 3))  # main.cpp:661-685
    assert p3.read_text() == want
    write_ppm(str(p6), img, binary=True)
    assert p6.read_bytes() == b"P6\n5 7\n255\n" + img.tobytes()


def test_bench_model_workloads_load_the_reference_models():
    """bench.py --workload c1|c3 run on the reference's own models: the flat scenes its loader produced, stored
    with the golden frames (the OBJ files themselves stay in the reference tree)."""
    import argparse
    import importlib
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    bench = importlib.import_module("bench")
    a = argparse.Namespace(tris=0, width=0, height=0)
    sc, W, H = bench.make_scene("c1", a)
    assert (sc.n_tris, W, H, sc.n_lights) == (36, 1024, 768, 1) and bench.EYE == (0.0, 1.0, 2.0)
    sc, W, H = bench.make_scene("c3", a)
    assert (sc.n_tris, W, H, sc.n_lights) == (7088, 1920, 1080, 1) and sc.tri_normals is not None
    assert "reference model" in bench.workload_config("c3", sc, W, H, 1, "brute")["workload"]


def test_loader_fuzz_against_the_reference_loader():
    """tests/fuzz_loader.py: random OBJ/MTL text (polygons, every index form, negative indices, g / o / usemtl / s in random
    places, CRLF, tabs, ~45 odd number spellings, shuffled and duplicated materials, d + Tr, several mtllib names) through
    model::loadobj (oracle/_ref, in a forked child: it crashes on input outside its contract) and through ours: equal bit for
    bit, or refused by both; where the reference crashes or indexes obj_materials[-1], ours refuses."""
    import subprocess
    import sys

    if not os.path.exists(MODELS):
        pytest.skip("reference tree not present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "fuzz_loader.py"), "250", "5"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and " 0 mismatches" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
    equal = int(r.stdout.split(" files: ")[1].split(" equal")[0])
    assert equal >= 80  # a good share of the files is loadable, i.e. the bit-for-bit comparison is exercised
