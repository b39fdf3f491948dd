"""Compile-proof of the drop-in: the reference's OWN main.cpp with the INTEGRATION.md section 3
binding applied (oracle/apply_cuda_patch.py -> oracle/_ref/ESCViewer2021_cuda, built in the authoring container and
shipped like the other built files) renders `-m model.obj --cuda --seed s` through libtracer_cuda.so, and its PPM is the
PPM the unmodified serial path wrote for the same seed (tests/golden/full/c1_cornell_original_1024x768: the reference's
default run, 1024x768 hard-wired at src/main.cpp:427, eye 0,1,2 as in scripts/run.sh:28-30)."""
import os
import subprocess

import numpy as np
import pytest
from conftest import GOLDEN, ROOT, write_obj

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "oracle", "_ref", "ESCViewer2021_cuda")


def test_reference_main_with_cuda_flag_writes_the_serial_paths_ppm(tmp_path):
    from esctp1raytracer_b200 import Scene

    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/ESCViewer2021_cuda not built (needs /root/reference at build time)")
    z = np.load(os.path.join(GOLDEN, "full", "c1_cornell_original_1024x768.npz"))
    sc = Scene(z["geom_tri_offset"], z["tri_verts"], z["geom_material"], z["light_geom"], geom_has_normals=z["geom_has_normals"])
    obj = tmp_path / "CornellBox-Original.obj"
    write_obj(sc, obj)
    out = tmp_path / "outputcuda.ppm"
    r = subprocess.run([BIN, "-m", str(obj), "-v", "0,1,2", "-l", "0,1,0", "--cuda", "--seed", str(int(z["seed"])), "-o", str(out)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Duration" in r.stderr and "Rendered image in" in r.stdout  # the reference's own prints (main.cpp:650-654, 688)
    tok = out.read_text().split()
    W, H = int(z["W"]), int(z["H"])
    assert tok[:4] == ["P3", str(W), str(H), "255"]
    got = np.array(tok[4:], dtype=np.int64).reshape(H, W, 3)
    assert np.array_equal(got, z["q"]), "the patched reference binary's PPM differs from the serial path's"
    # and the unpatched path of the same binary still is the reference's serial renderer (no --cuda: CPU, random seed):
    # it must at least hit the same pixels (ids do not depend on the RNG), i.e. write a non-black frame of the same size
    out2 = tmp_path / "output.ppm"
    r2 = subprocess.run([BIN, "-m", str(obj), "-v", "0,1,2", "-l", "0,1,0", "-o", str(out2)], capture_output=True, text=True, timeout=600)
    assert r2.returncode == 0
    tok2 = out2.read_text().split()
    assert tok2[:4] == tok[:4]
    serial = np.array(tok2[4:], dtype=np.int64).reshape(H, W, 3)
    # different random light samples, same scene: most pixels agree exactly (each pixel picks one of 2 light vertices)
    assert (serial == got).all(axis=2).mean() > 0.4
