"""N-GPU frame == 1-GPU frame, on hardware.  Skipped on boxes with one GPU; the driver's
scaling run and `gpurun --gpus N` exercise them.  Two paths: the library's own (one process, tracer_cuda_init_multi,
NCCL send/recv inside the C ABI, also reachable as `tracer_cli --gpus N`) and the one-process-per-GPU path bench.py
uses under torchrun (dist.render_frame), here run through torch.multiprocessing."""
import os
import subprocess
import sys

import numpy as np
import pytest
from conftest import ROOT

pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch

    return torch.cuda.device_count()


def test_native_multi_gpu_equals_single_gpu():
    from esctp1raytracer_b200 import RNG_EXPLICIT, RNG_HASH, Camera, MultiRenderer, Renderer, hash_faceids, scenes

    n = _n_gpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    s = scenes.soup_scene(30000, 30, 3, seed=9, n_spheres=5)
    W, H = 200, 121  # 16 bands of 8 rows, the last one ragged
    cam = Camera.for_frame((0, 1, 3), (0, 1, 0), W, H)
    one = Renderer(0)
    want = one.trace(one.upload(s), cam, W, H, rng_mode=RNG_HASH, seed=4).rgb8
    want4 = one.trace(one.upload(s), cam, W, H, rng_mode=RNG_HASH, seed=4, samples_per_pixel=4).rgb8
    for k in sorted({2, n}):
        m = MultiRenderer(k)
        rs = m.upload(s)
        a = m.trace(rs, cam, W, H, rng_mode=RNG_HASH, seed=4)
        assert np.array_equal(a.rgb8, want), f"{k}-GPU frame differs"
        assert a.stats["n_primary_rays"] == W * H and a.stats["tests_primary"] == W * H * s.n_tris
        assert np.array_equal(m.trace(rs, cam, W, H, rng_mode=RNG_HASH, seed=4, bundle_cull=True).rgb8, want)
        assert np.array_equal(m.trace(rs, cam, W, H, rng_mode=RNG_HASH, seed=4, samples_per_pixel=4).rgb8, want4)
        assert np.array_equal(m.trace(rs, cam, W, H, rng_mode=RNG_HASH, seed=4, band_rows=5).rgb8, want)
        fid = hash_faceids(4, W, H, s.faces_per_light)
        assert np.array_equal(m.trace(rs, cam, W, H, rng_mode=RNG_EXPLICIT, faceid=fid).rgb8, want)
        rs.close()
        assert np.array_equal(m.trace(s, cam, W, H, rng_mode=RNG_HASH, seed=4).rgb8, want)  # the drop-in call


def test_cli_gpus_flag_writes_the_same_ppm(tmp_path):
    from esctp1raytracer_b200 import scenes

    n = _n_gpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    s = scenes.box_scene()
    from conftest import write_obj

    obj = tmp_path / "box.obj"
    write_obj(s, obj)
    cli = os.path.join(ROOT, "esctp1raytracer_b200", "tracer_cli.bin")
    outs = []
    for k in (1, n):
        out = tmp_path / f"out{k}.ppm"
        r = subprocess.run([cli, "-m", str(obj), "-v", "0,1,2.9", "-l", "0,1,0", "-w", "160,100", "--rng", "hash", "--seed", "5",
                            "--gpus", str(k), "-o", str(out)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        outs.append(out.read_bytes())
    assert outs[0] == outs[1]


def _rank_main(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from esctp1raytracer_b200 import RNG_HASH, Camera, Renderer, scenes
    from esctp1raytracer_b200 import dist as tdist

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    r = Renderer(rank)
    s = scenes.soup_scene(30000, 30, 3, seed=9, n_spheres=5)
    W, H = 200, 121
    cam = Camera.for_frame((0, 1, 3), (0, 1, 0), W, H)
    rs = r.upload(s)
    frame, _ = tdist.render_frame(r, rs, cam, W, H, rank=rank, world=world, seed=4)
    frame_c, _ = tdist.render_frame(r, rs, cam, W, H, rank=rank, world=world, seed=4, bundle_cull=True)
    if rank == 0:
        want = r.trace(rs, cam, W, H, rng_mode=RNG_HASH, seed=4).rgb8
        q.put((bool(np.array_equal(frame.cpu().numpy(), want)), bool(np.array_equal(frame_c.cpu().numpy(), want))))
    dist.destroy_process_group()


def test_torch_distributed_bands_equal_single_gpu():
    import torch.multiprocessing as mp

    n = _n_gpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 400
    procs = [ctx.Process(target=_rank_main, args=(r, n, port, q)) for r in range(n)]
    for p in procs:
        p.start()
    ok = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
    assert ok == (True, True)
