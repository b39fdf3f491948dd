"""Generate the committed golden fixtures from the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference and oracle/_ref):
    python tests/golden/make_golden.py
Each fixture holds the flat-scene dump the reference's own loader produced
(model::loadobj, src/scene/sceneloader.cpp:14-106) and one seeded frame rendered by
the reference's own scan_row (src/main.cpp:698-791) through oracle/ref_harness.cpp:
per-pixel hit ids / t / v from intersect(), the replayed faceIDs, the float
accumulator and the quantised PPM values.  The GPU box has no /root/reference;
tests there read these files.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import FlatScene, RefOracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
M = "/root/reference/src/models/"

CASES = [
    # name, model, eye, look, W, H, seed, extra light geoms appended to light_sources
    ("cornell_original", "cornell/CornellBox-Original.obj", (0, 1, 2), (0, 1, 0), 128, 96, 7, []),
    ("cornell_box_ks", "cornell_box.obj", (0.3, 1, 1.5), (0, 0.8, 0), 96, 72, 11, []),
    ("cornell_sphere", "cornell/CornellBox-Sphere.obj", (0, 1, 2), (0, 1, 0), 64, 48, 5, []),
    ("cornell_water", "cornell/CornellBox-Water.obj", (0, 1, 2), (0, 1, 0), 48, 36, 3, []),
    # 3 lights: the light itself plus two ordinary geometries used as lights
    # (exercises the t carry between lights, src/main.cpp:764/772)
    ("cornell_original_3lights", "cornell/CornellBox-Original.obj", (0, 1, 2), (0, 1, 0), 96, 72, 9, [1, 5]),
    ("cornell_empty_co", "cornell/CornellBox-Empty-CO.obj", (0, 1, 2.5), (0, 1, 0), 64, 48, 2, []),
]


# BASELINE.json configs at their configured sizes: C1/C2 = the reference's default run
# (scripts/run.sh:28-30, 1024x768, src/main.cpp:427), C3 = the two largest loadable meshes at 1920x1080 in geometry
# order AND in the flatten+sort order of src/simplify/flatten.cpp:50-82 (our host helper builds the sorted scene, the
# unmodified reference renders it).  The per-pixel float arrays are too big to commit whole: tri / q / faceid are kept
# in full, t / v / rgb as SHA-256 digests of their bytes plus every 97th pixel for diagnostics.
BIG_CASES = [
    # name, model, eye, look, W, H, seed, sorted
    ("c1_cornell_original_1024x768", "cornell/CornellBox-Original.obj", (0, 1, 2), (0, 1, 0), 1024, 768, 1, False),
    ("c3_sphere_1080p", "cornell/CornellBox-Sphere.obj", (0, 1, 2), (0, 1, 0), 1920, 1080, 5, False),
    ("c3_sphere_1080p_sorted", "cornell/CornellBox-Sphere.obj", (0, 1, 2), (0, 1, 0), 1920, 1080, 5, True),
    ("c3_water_1080p", "cornell/CornellBox-Water.obj", (0, 1, 2), (0, 1, 0), 1920, 1080, 3, False),
    ("c3_water_1080p_sorted", "cornell/CornellBox-Water.obj", (0, 1, 2), (0, 1, 0), 1920, 1080, 3, True),
]
STRIDE = 97


def sha(a):
    import hashlib

    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def big(only=None):
    from esctp1raytracer_b200 import Scene

    ref = RefOracle()
    for name, model, eye, look, W, H, seed, srt in BIG_CASES:
        if only and name != only:
            continue
        h = ref.load_obj(M + model)
        fs = ref.dump(h)
        origin = None
        if srt:
            sc = Scene(fs.geom_tri_offset, fs.tri_verts, fs.geom_material, fs.light_geom, tri_normals=fs.tri_normals,
                       geom_has_normals=fs.geom_has_normals)
            s2, og, op = sc.flatten_sorted()
            ref.free(h)
            fs = FlatScene(s2.geom_tri_offset, s2.tri_verts, s2.tri_normals, s2.geom_has_normals, s2.geom_material, s2.light_geom)
            h = ref.from_flat(fs)
            origin = np.stack([og, op], 1).astype(np.int32)
        fr, exact = ref.render_frame(h, W, H, eye, look, seed=seed)
        assert exact, "mt19937 replay diverged from scan_row's generator"
        cam = ref.camera(eye, look, W, H)
        tri = np.where(fr["geom"] >= 0, fs.geom_tri_offset[np.maximum(fr["geom"], 0)] + fr["prim"], -1).astype(np.int32)
        assert fr["q"].min() >= 0 and fr["q"].max() <= 255
        assert fr["faceid"].min() >= -1 and fr["faceid"].max() < 127
        np.savez_compressed(
            os.path.join(HERE, "full", name + ".npz"),
            geom_tri_offset=fs.geom_tri_offset, tri_verts=fs.tri_verts,
            tri_normals=fs.tri_normals if fs.tri_normals is not None else np.zeros((0, 3, 3), np.float32),
            geom_has_normals=fs.geom_has_normals, geom_material=fs.geom_material, light_geom=fs.light_geom,
            eye=np.array(eye, np.float32), look=np.array(look, np.float32), W=W, H=H, seed=seed, cam=cam,
            tri=tri, faceid=fr["faceid"].astype(np.int8), q=fr["q"].astype(np.uint8),
            sha_t=sha(fr["t"]), sha_v=sha(fr["v"]), sha_rgb=sha(fr["rgb"]), stride=STRIDE,
            t_s=fr["t"][::STRIDE], v_s=fr["v"][::STRIDE], rgb_s=fr["rgb"][::STRIDE],
            origin=origin if origin is not None else np.zeros((0, 2), np.int32), sorted=int(srt),
        )
        print(name, "N", fs.n_tris, "G", fs.n_geoms, "L", fs.n_lights, "hit", (tri >= 0).mean(), "mean q", fr["q"].mean(), flush=True)
        ref.free(h)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "big":
        return big(sys.argv[2] if len(sys.argv) > 2 else None)
    ref = RefOracle()
    for name, model, eye, look, W, H, seed, extra in CASES:
        h = ref.load_obj(M + model)
        fs = ref.dump(h)
        if extra:
            fs.light_geom = np.concatenate([fs.light_geom, np.array(extra, np.int32)])
            ref.free(h)
            h = ref.from_flat(fs)
            fs2 = ref.dump(h)
            assert np.array_equal(fs2.tri_verts, fs.tri_verts) and np.array_equal(fs2.light_geom, fs.light_geom)
        fr, exact = ref.render_frame(h, W, H, eye, look, seed=seed)
        assert exact, "mt19937 replay diverged from scan_row's generator"
        cam = ref.camera(eye, look, W, H)
        tri = np.where(fr["geom"] >= 0, fs.geom_tri_offset[np.maximum(fr["geom"], 0)] + fr["prim"], -1).astype(np.int32)
        assert fr["q"].min() >= 0 and fr["q"].max() <= 255
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            geom_tri_offset=fs.geom_tri_offset, tri_verts=fs.tri_verts,
            tri_normals=fs.tri_normals if fs.tri_normals is not None else np.zeros((0, 3, 3), np.float32),
            geom_has_normals=fs.geom_has_normals, geom_material=fs.geom_material, light_geom=fs.light_geom,
            eye=np.array(eye, np.float32), look=np.array(look, np.float32), W=W, H=H, seed=seed, cam=cam,
            tri=tri, t=fr["t"], v=fr["v"], faceid=fr["faceid"], rgb=fr["rgb"], q=fr["q"].astype(np.uint8),
        )
        print(name, "N", fs.n_tris, "L", fs.n_lights, "hit", (tri >= 0).mean(), "mean q", fr["q"].mean())
        ref.free(h)


if __name__ == "__main__":
    main()
