import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


def load_golden(name):
    """-> (FlatScene for the oracle, dict of the reference's frame)"""
    from oracle import FlatScene

    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    tn = z["tri_normals"]
    fs = FlatScene(z["geom_tri_offset"], z["tri_verts"], tn if len(tn) else None, z["geom_has_normals"],
                   z["geom_material"], z["light_geom"])
    fr = {k: z[k] for k in ("eye", "look", "cam", "tri", "t", "v", "faceid", "rgb", "q")}
    fr.update(W=int(z["W"]), H=int(z["H"]), seed=int(z["seed"]))
    return fs, fr


def to_scene(fs):
    """oracle FlatScene -> product Scene (same arrays)"""
    from esctp1raytracer_b200 import Scene

    return Scene(fs.geom_tri_offset, fs.tri_verts, fs.geom_material, fs.light_geom, tri_normals=fs.tri_normals,
                 geom_has_normals=fs.geom_has_normals, sphere_cr=fs.sphere_cr, sphere_material=fs.sphere_material)


def to_flat(s):
    from oracle import FlatScene

    return FlatScene(s.geom_tri_offset, s.tri_verts, s.tri_normals, s.geom_has_normals, s.geom_material, s.light_geom,
                     s.sphere_cr, s.sphere_material)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="session")
def restated():
    from oracle import Restated, build

    build()
    return Restated()


@pytest.fixture(scope="session")
def ref_oracle():
    from oracle import REF_ROOT, RefOracle, build, ref_available

    if os.path.exists(os.path.join(REF_ROOT, "src", "main.cpp")):
        build()
    if not ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    return RefOracle()


def write_obj(s, obj_path, mtl_name="scene.mtl"):
    """Write a flat scene as OBJ + MTL (one `g` + `usemtl` per geometry, floats via repr so that they parse back to the
    same bits): what the reference's loader (and ours) turns back into the same flat scene."""
    mtl_path = os.path.join(os.path.dirname(str(obj_path)), mtl_name)
    with open(mtl_path, "w") as f:
        for g in range(s.n_geoms):
            m = [float(x) for x in s.geom_material[g]]
            f.write(f"newmtl m{g}\nKa {m[0]!r} {m[1]!r} {m[2]!r}\nKd {m[3]!r} {m[4]!r} {m[5]!r}\nKs {m[6]!r} {m[7]!r} {m[8]!r}\n"
                    f"Ke {m[9]!r} {m[10]!r} {m[11]!r}\nNs {m[12]!r}\n")
    with open(obj_path, "w") as f:
        f.write(f"mtllib {mtl_name}\n")
        for v in s.tri_verts.reshape(-1, 3).tolist():
            f.write(f"v {v[0]!r} {v[1]!r} {v[2]!r}\n")
        for g in range(s.n_geoms):
            f.write(f"g geom{g}\nusemtl m{g}\n")
            for t in range(s.geom_tri_offset[g], s.geom_tri_offset[g + 1]):
                f.write(f"f {3 * t + 1} {3 * t + 2} {3 * t + 3}\n")
