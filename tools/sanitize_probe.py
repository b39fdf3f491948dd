"""Small frames through every sweep variant, for compute-sanitizer (racecheck / memcheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_probe.py
Default mode (TMA tile pipeline with per-warp stage recycling, persistent cooperative shadow kernel with grid barriers
and in-kernel compaction), jittered samples (own q per ray), exhaustive-strict validation mode, and both forms of the
optional bundle-cull mode.  Prints one line per mode and a final OK; the frames must all be identical."""
import sys

import numpy as np

sys.path.insert(0, ".")
from esctp1raytracer_b200 import RNG_HASH, Camera, Renderer, scenes

r = Renderer(0)
s = scenes.soup_scene(20000, 20, 2, seed=3, n_spheres=7, edge=(0.03, 0.2))
W, H = 128, 80
cam = Camera.for_frame((0, 1, 3), (0, 1, 0), W, H)
rs = r.upload(s)
ref = None
for name, kw in (("default", {}), ("default chunks=24", dict(shadow_chunks=24)), ("exhaustive", dict(exhaustive_strict=True)),
                 ("cull two-phase", dict(bundle_cull=1)), ("cull streaming", dict(bundle_cull=2)), ("4 spp", dict(samples_per_pixel=4)),
                 ("R=4", dict(rays_per_thread=4)), ("R=2", dict(rays_per_thread=2))):
    out = r.trace(rs, cam, W, H, rng_mode=RNG_HASH, seed=5, **kw)
    if name != "4 spp":
        ref = out.rgb8 if ref is None else ref
        assert np.array_equal(out.rgb8, ref), name
    print(name, "launches", out.stats["kernel_launches"], "strict", out.stats["strict_evals"], "filter misses", out.stats["filter_misses"],
          "pipeline errors", out.stats["pipeline_errors"], flush=True)
    assert out.stats["filter_misses"] == 0 and out.stats["pipeline_errors"] == 0
print("OK")
