"""Multi-GPU frame check (development tool, not a test: needs N GPUs).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/check_dist.py

Renders the C4 frame (NTRIS triangles, default 1M) twice in each mode (default, bundle-cull two-phase, bundle-cull
streaming) through dist.render_frame on N ranks and compares every assembled frame, row by row, with the frame one
GPU renders alone.  Expected output: 0 differing rows everywhere."""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esctp1raytracer_b200 import Camera, Renderer, scenes
from esctp1raytracer_b200 import dist as tdist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
r = Renderer(lr)
W, H = 3840, 2160
n = int(os.environ.get("NTRIS", "1000000"))
s = scenes.soup_scene(n, 1000, 4, n_spheres=1000, seed=42)
cam = Camera.for_frame((0, 1, 3), (0, 1, 0), W, H)
rs = r.upload(s)
frames = {}
for it in range(2):
    for name, bc in (("default", 0), ("cull", 1), ("stream", 2)):
        f, st = tdist.render_frame(r, rs, cam, W, H, rank=rank, world=world, seed=42, bundle_cull=bc)
        torch.cuda.synchronize()
        if rank == 0:
            frames[(name, it)] = f.cpu().numpy()
if rank == 0:
    full, _ = tdist.render_frame(r, rs, cam, W, H, rank=0, world=1, seed=42, bundle_cull=1)
    full = full.cpu().numpy()
    for k, v in frames.items():
        d = (v != full).any(axis=(1, 2))
        print(k, "rows differing from the 1-GPU frame:", int(d.sum()), np.nonzero(d)[0][:12].tolist(), flush=True)
dist.destroy_process_group()
