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
