"""BASELINE.json configs[4] as configured: 7680x4320, 16 spp stratified jittered primary rays, 1 light, triangle-count
sweep, on all ranks of a torchrun launch (extension: parity unpinned, oracle = the restatement at small sizes,
tests/test_gpu_fullsize.py::test_16spp_jittered_parity).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/c5_sweep.py \
        [--brute 10000,100000,1000000,4000000] [--cull 10000,30000,...] [--width 7680 --height 4320 --spp 16]

One JSON line per (mode, N) on rank 0: ms/frame (max over ranks, CUDA events around the frame incl. the band gather),
Mrays/s, and for the default (brute-force) mode the FP32 rate of the sweeps: with jitter the 32 rays of a thread (one sample
of 32 pixels of an image row) share a mean q within their stratum, so the closest-hit sweep runs the span form with
8 FFMA per thread and triangle: 4 + 16/32 = 4.5 flops per pair (round 2 before the span form: three rows, own q, 15)."""
import argparse
import hashlib
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esctp1raytracer_b200 import Camera, Renderer, scenes  # noqa: E402
from esctp1raytracer_b200 import dist as tdist  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--brute", default="10000,100000,1000000,4000000")
ap.add_argument("--cull", default="10000,30000,100000,300000,1000000,2000000,4000000")
ap.add_argument("--width", type=int, default=7680)
ap.add_argument("--height", type=int, default=4320)
ap.add_argument("--spp", type=int, default=16)
ap.add_argument("--steps", type=int, default=1)
args = ap.parse_args()
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
r = Renderer(lr)
W, H = args.width, args.height
cam = Camera.for_frame((0, 1, 3), (0, 1, 0), W, H)
peak = max(r.fp32_peak(v, 5)[0] for v in (0, 1, 3))
for mode, sizes in (("cull", args.cull), ("brute", args.brute)):
    for n in [int(x) for x in sizes.split(",") if x]:
        s = scenes.soup_scene(n, min(1000, max(8, n // 100)), 1, seed=42)
        rs = r.upload(s)
        kw = dict(rank=rank, world=world, seed=42, bundle_cull=(mode == "cull"), samples_per_pixel=args.spp)
        # warm-up at 1 spp (1/16 of the cost): builds the filter tables, sizes the workspace, loads the kernels
        tdist.render_frame(r, rs, cam, W, H, **dict(kw, samples_per_pixel=1))
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        acc = {}
        for _ in range(args.steps):
            frame, st = tdist.render_frame(r, rs, cam, W, H, **kw)
            for k, v in st.items():
                acc[k] = acc.get(k, 0) + v
        e1.record()
        torch.cuda.synchronize()
        keys = ["n_primary_rays", "n_shadow_rays", "tests_primary", "tests_shadow", "tests_shadow_ref", "strict_evals", "ms_primary", "ms_shadow"]
        t = torch.tensor([e0.elapsed_time(e1)] + [float(acc[k]) for k in keys], dtype=torch.float64, device="cuda")
        mx = t.clone()
        if world > 1:
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        if rank == 0:
            tot = {k: float(t[i + 1]) / args.steps for i, k in enumerate(keys)}
            ms = float(mx[0]) / args.steps
            rays = tot["n_primary_rays"] + tot["n_shadow_rays"]
            sweep_s = (float(mx[7]) + float(mx[8])) / args.steps * 1e-3
            line = dict(config="c5", mode=mode, n_tris=n, width=W, height=H, spp=args.spp, n_gpus=world, ms_per_frame=ms,
                        mrays_s=rays / ms / 1e3, rays=rays, strict_per_ray=tot["strict_evals"] / rays,
                        frame_sha256=hashlib.sha256(frame.cpu().numpy().tobytes()).hexdigest())
            if mode == "brute":
                fp, fs = float(st["flop_primary"]), float(st["flop_shadow"])
                alg = fp * tot["tests_primary"] + fs * tot["tests_shadow_ref"]
                line.update(flop_per_pair_primary=fp, flop_per_pair_shadow=fs, flop_per_pair_primary_edge_rows=float(st["flop_primary_edges"]),
                            pairs_primary=tot["tests_primary"], pairs_shadow_ref=tot["tests_shadow_ref"], sweep_ms=sweep_s * 1e3,
                            tflops_per_gpu=alg / world / sweep_s / 1e12, fp32_peak_measured=peak, frac=alg / world / sweep_s / 1e12 / peak,
                            tpairs_per_s_per_gpu=(tot["tests_primary"] + tot["tests_shadow"]) / world / sweep_s / 1e12)
            print(json.dumps(line), flush=True)
        rs.close()
if world > 1:
    dist.destroy_process_group()
