"""One-off check: the UNMODIFIED reference serial path (oracle/ref_harness.cpp over src/main.cpp,
sceneloader.cpp, ...) and the plain-C restatement (oracle/restated.c, scalar and AVX2 loops, threaded) under
AddressSanitizer + UndefinedBehaviorSanitizer.  Test infrastructure; needs /root/reference.

    make -C oracle ref_asan CXX=/usr/bin/g++ CC=/usr/bin/gcc
    LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libstdc++.so)" ASAN_OPTIONS=detect_leaks=0 \
        python tests/oracle_sanitize.py

(libstdc++ is preloaded so that ASan can intercept __cxa_throw — the reference's loader throws on its two
unloadable models; leak detection is off because the interpreter itself never frees everything)
"""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

SAN = os.path.join(ROOT, "oracle", "_ref_asan")
oracle.REF_SO = os.path.join(SAN, "libref_oracle.so")
oracle.RESTATED_SO = os.path.join(SAN, "librestated.so")
MODELS = "/root/reference/src/models"

ref, rst = oracle.RefOracle(), oracle.Restated()
n_ok = n_refused = 0
for path in sorted(glob.glob(os.path.join(MODELS, "**", "*.obj"), recursive=True)):
    try:
        h = ref.load_obj(path)
    except RuntimeError:
        n_refused += 1  # warnings are fatal in the reference (sceneloader.cpp:27-30)
        continue
    fs = ref.dump(h)
    W, H = (96, 72) if fs.n_tris < 3000 else (40, 30)
    eye, look = (0.0, 1.0, 2.0), (0.0, 1.0, 0.0)
    fr, replay_exact = ref.render_frame(h, W, H, eye, look, seed=3)          # scan_row, intersect, occlusion, RNG replay
    assert replay_exact
    pw, ph = np.arange(W, dtype=np.int32), np.full(W, H // 2, np.int32)
    ref.render_pixels(h, W, H, eye, look, pw, ph, fr["faceid"].reshape(H, W, -1)[H // 2], n_threads=4)
    ref.time_rows(h, W, H, eye, look, 3, H // 2, H // 2 + 4, 1)               # the reference's thread-per-row scheme
    ref.free(h)
    # the reference's own camera from the SAME build: camera.h:20 calls tan(float); an optimised build folds it at compile
    # time (correctly rounded, what oracle/_ref and the product use), an instrumented or -O0 build calls glibc's tanf at
    # run time, which returns the neighbouring float for vfov = 60 — the frames then differ in most pixels
    cam = ref.camera(eye, look, W, H)
    for simd in (False, True):
        rst.set_simd(simd)
        o = rst.render(fs, cam, W, H, seed=3, faceid=fr["faceid"], n_threads=4)
        assert np.array_equal(o.rgb.view(np.uint32), fr["rgb"].view(np.uint32)), path  # bit-equal under the sanitizers too
    n_ok += 1
    print(f"clean: {os.path.relpath(path, MODELS)}  ({fs.n_tris} triangles, {W}x{H})", flush=True)
print(f"oracle_sanitize: {n_ok} models rendered by the reference and the restatement under ASan+UBSan, {n_refused} refused "
      f"by the reference's loader, no sanitizer report")
