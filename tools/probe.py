"""Quick on-GPU probe: FP32 peak microbenchmarks + sweep throughput on soup scenes."""
import json
import os
import sys
import time


sys.path.insert(0, ".")
from esctp1raytracer_b200 import RNG_HASH, Camera, Renderer, scenes

r = Renderer(0)
print(json.dumps(r.device_info()))
argv = sys.argv[1:]
if argv and argv[0] == "nopeak":
    argv = argv[1:]
else:
    for v in (0, 1, 2, 3):
        tf, ms = r.fp32_peak(v, 10)
        print(f"fp32_peak variant {v}: {tf:.2f} TFLOP/s  ({ms:.3f} ms/launch)")
cases = [(200_000, 1920, 1080, 4), (1_000_000, 960, 540, 4)]
if argv:
    cases = [tuple(int(x) for x in a.split(",")) for a in argv]
for n, W, H, L in cases:
    s = scenes.soup_scene(n, max(10, n // 1000), L, seed=42)
    look = (0, 1, 6) if os.environ.get('PROBE_AWAY') else (0, 1, 0)
    cam = Camera.for_frame((0, 1, 3), look, W, H)
    rs = r.upload(s)
    bands = tuple(int(x) for x in os.environ['PROBE_BANDS'].split(',')) if os.environ.get('PROBE_BANDS') else None  # e.g. 8,0,8
    for it in range(3):
        t0 = time.time()
        out = r.trace(rs, cam, W, H, rng_mode=RNG_HASH, seed=1, bands=bands, rays_per_thread=int(os.environ.get('TRACER_RAYS', '0')),
                      shadow_chunks=int(os.environ.get('TRACER_CHUNKS', '0')), bundle_cull=bool(os.environ.get('TRACER_CULL')))
        wall = time.time() - t0
    st = out.stats
    fp, fs = st.get("flop_primary") or 0.0, st.get("flop_shadow") or 0.0  # flops the formulation needs per pair (0 in cull mode)
    prim = st["tests_primary"] * fp / (st["ms_primary"] * 1e-3) / 1e12
    shad = st["tests_shadow"] * fs / max(st["ms_shadow"], 1e-9) / 1e-3 / 1e12
    shad_ref = st["tests_shadow_ref"] * fs / max(st["ms_shadow"], 1e-9) / 1e-3 / 1e12
    rays = st["n_primary_rays"] + st["n_shadow_rays"]
    print(json.dumps(dict(R=os.environ.get("TRACER_RAYS", "auto"), chunks=os.environ.get("TRACER_CHUNKS", "auto"), n_tris=n, W=W, H=H, L=L, wall_s=round(wall, 3), **{k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()},
                          primary_tflops=round(prim, 2), shadow_tflops_swept=round(shad, 2), shadow_tflops_ref=round(shad_ref, 2),
                          mrays_s=round(rays / st["ms_total"] / 1e3, 3), hit_frac=round(st["n_shadow_rays"] / max(1, L) / st["n_pixels"], 3),
                          strict_per_ray=round(st["strict_evals"] / rays, 2))))
    rs.close()

# ---- one-shot (host buffers) path timing breakdown --------------------------------------
if os.environ.get("PROBE_E2E"):
    s = scenes.soup_scene(100_000, 100, 4, seed=42)
    W, H = 1280, 720
    cam = Camera.for_frame((0, 1, 3), (0, 1, 0), W, H)
    for it in range(3):
        t0 = time.time(); rs = r.upload(s); t1 = time.time()
        out = r.trace(rs, cam, W, H, rng_mode=RNG_HASH, seed=1); t2 = time.time()
        out = r.trace(rs, cam, W, H, rng_mode=RNG_HASH, seed=1); t3 = time.time()
        rs.close(); t4 = time.time()
        print(f"e2e breakdown: upload {t1-t0:.4f}s first-trace {t2-t1:.4f}s second-trace {t3-t2:.4f}s close {t4-t3:.4f}s (gpu ms_total {out.stats['ms_total']:.1f})")
