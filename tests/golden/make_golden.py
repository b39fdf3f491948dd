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


def main():
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
