"""One-off fuzz of the oracle pin: random synthetic scenes (triangle count, geometry count, 1-4 lights, edge lengths, specular,
normals, camera, frame size, seed) rendered by the UNMODIFIED reference (oracle/_ref: scan_row) and by the restatement (scalar and AVX2
loops): faceIDs, every float channel of every pixel, t, v and the camera must be bit-equal.  Needs /root/reference (test infrastructure).
    python tests/fuzz_oracle.py [n_scenes]
"""
import sys, numpy as np
import os; ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tests'))
import oracle
from conftest import to_flat, bits
from esctp1raytracer_b200 import scenes
ref, rst = oracle.RefOracle(), oracle.Restated()
rng=np.random.default_rng(5)
bad=0
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 24):
    n=int(rng.integers(200,6000)); L=int(rng.integers(1,5)); g=int(rng.integers(L+3,24))
    lo=10**rng.uniform(-2.5,-0.8); hi=lo*rng.uniform(1.5,6)
    s=scenes.soup_scene(n,g,L,seed=int(rng.integers(1,10**6)),edge=(lo,hi),specular=bool(rng.integers(0,2)),with_normals=bool(rng.integers(0,2)))
    fs=to_flat(s)
    W,H=int(rng.integers(24,90)),int(rng.integers(18,70))
    eye=(float(rng.normal()*0.3),1+float(rng.normal()*0.3),3+float(rng.normal()*0.3)); look=(0,1,0)
    seed=int(rng.integers(1,10**6))
    h=ref.from_flat(fs)
    fr,exact=ref.render_frame(h,W,H,eye,look,seed=seed)
    cam=rst.camera(eye,look,W,H)
    for simd in (False,True):
        rst.set_simd(simd)
        o=rst.render(fs,cam,W,H,seed=seed)
        ok = exact and np.array_equal(o.faceid,fr['faceid']) and np.array_equal(bits(o.rgb),bits(fr['rgb'])) and np.array_equal(bits(o.t),bits(fr['t'])) and np.array_equal(bits(o.v),bits(fr['v'])) and np.array_equal(bits(cam), bits(ref.camera(eye,look,W,H)))
        if not ok: bad+=1; print('MISMATCH', it, simd, n,g,L,W,H)
    ref.free(h)
    print(it, n, g, L, W, H, 'ok', flush=True)
print('bad', bad)
