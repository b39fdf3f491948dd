#!/usr/bin/env python
"""bench.py — the reference's headline metric on BASELINE.json's config.

    python bench.py --gpus N --steps K --warmup W            (ours; torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W

Metric: Mrays/s (primary + shadow rays of one frame / frame time).  A "step" is one frame
of the hot path over synthetic input.  Default workload = BASELINE.json configs[3]:
synthetic 1M-triangle + 1k-sphere scene at 3840x2160, 4 lights ("c4").  Other workloads
(--workload c1|c3|small) are development conveniences, not bench lines.

`value`  : scene resident in HBM, frame left in HBM on rank 0 (CUDA events, max over ranks).
`e2e`    : the drop-in C-ABI call with HOST buffers each step — scene upload (pinned host ->
           HBM), render, gather, device -> host read of the packed frame.
`roofline`: FP32-FMA bound (no tensor cores on this path).  achieved = the flops the sweeps
           EXECUTE per ray-triangle pair (shadow 12 = 6 FFMA; primary 6.75: the q-term of each
           edge row is shared by the 8 rays of a thread) x ALGORITHMIC pairs (P*N primary + the
           reference's own in-order count for shadow rays) / sweep-kernel time; peak = our own FFMA microbenchmark measured in the same process
           (MEASURED_PEAKS.json has no FP32 number); nominal peak printed beside it.
`cpu_baseline`: oracle/_ref (the unmodified reference compiled from /root/reference; kind
           "reference") or the plain-C port, on a bounded pixel sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_PAIR = 12.0      # 6 FFMA (2-D affine edge rows) per (ray, triangle) pair when every ray evaluates its own rows.
                          # The sweeps share the q-term of each row among the 8 rays of a thread: primary 2*(3+3*8)/8 = 6.75
                          # (same image row: exact), shadow 2*(6+3*8)/8 = 7.5 (q-sorted rays: mean q + |B|*spread); the library
                          # reports the figures it ran (stats flop_primary / flop_shadow)
FLOP_PER_PAIR_REF = 46.0  # Moller-Trumbore with precomputed edges, reported beside it
EYE, LOOK = (0.0, 1.0, 3.0), (0.0, 1.0, 0.0)

WORKLOADS = {
    # name: (n_tris, n_geoms, n_lights, n_spheres, W, H)
    "c4": (1_000_000, 1000, 4, 1000, 3840, 2160),
    "c3": (7_088, 9, 1, 0, 1920, 1080),       # CornellBox-Water as loaded by the reference (tests/golden/cornell_water.npz)
    "c1": (36, 7, 1, 0, 1024, 768),           # CornellBox-Original, the reference's default run (tests/golden/cornell_original.npz)
    "small": (100_000, 100, 4, 100, 1280, 720),
    # BASELINE.json configs[4]: 7680x4320, 16 spp jittered primary rays, 1 light, triangle-count sweep via --tris
    # (extension, parity unpinned; not a bench line: run with --workload c5 --tris N [--mode cull])
    "c5": (1_000_000, 1000, 1, 0, 7680, 4320),
}
WORKLOAD_SPP = {"c5": 16}
GOLDEN_MODELS = {"c1": "cornell_original.npz", "c3": "cornell_water.npz"}


def make_scene(name, args):
    from esctp1raytracer_b200 import scenes

    n_tris, n_geoms, n_lights, n_spheres, W, H = WORKLOADS[name]
    n_tris = args.tris or n_tris
    W, H = args.width or W, args.height or H
    if name in GOLDEN_MODELS and not args.tris:
        global EYE, LOOK
        sc, EYE, LOOK = scenes.from_npz(os.path.join(ROOT, "tests", "golden", GOLDEN_MODELS[name]))
        return sc, W, H
    return scenes.soup_scene(n_tris, min(n_geoms, max(8, n_tris // 100)), n_lights, n_spheres=n_spheres, seed=42), W, H


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])), mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [x for x in sm if x > 0.5 * (max(mx) if mx else 1)] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_sample(scene, W, H, n_pixels, threads, seed=42):
    """Time the reference's own CPU path on a pixel sample of this workload."""
    from esctp1raytracer_b200 import hash_faceids
    from oracle import FlatScene, RefOracle, Restated, ref_available

    rng = np.random.default_rng(7)
    idx = rng.choice(W * H, size=n_pixels, replace=False)
    ph, pw = (idx // W).astype(np.int32), (idx % W).astype(np.int32)
    fid = hash_faceids(seed, W, H, scene.faces_per_light)[idx]
    fs = FlatScene(scene.geom_tri_offset, scene.tri_verts, scene.tri_normals, scene.geom_has_normals,
                   scene.geom_material, scene.light_geom)  # the reference has no spheres
    L = scene.n_lights
    if ref_available():
        ref = RefOracle()
        h = ref.from_flat(fs)
        out = ref.render_pixels(h, W, H, EYE, LOOK, pw, ph, fid, n_threads=threads)
        ref.free(h)
        secs, hits, kind = out["seconds"], int((out["geom"] >= 0).sum()), "reference"
    else:
        from esctp1raytracer_b200 import Camera

        rst = Restated()
        t0 = time.time()
        o = rst.render_pixels(fs, Camera.for_frame(EYE, LOOK, W, H).as_array(), W, H, pw, ph, fid, n_threads=threads)
        secs, hits, kind = time.time() - t0, int((o.tri >= 0).sum()), "port"
    rays = n_pixels + hits * L
    # the same algorithm as a SIMD CPU comparator (oracle/restated.c, 8 triangles per AVX2 step, pinned
    # bit-identical to the scalar restatement), timed on the same pixels: what a vectorised CPU build would reach
    simd = None
    try:
        from esctp1raytracer_b200 import Camera

        rst = Restated()
        if rst.set_simd(True):
            cam12 = Camera.for_frame(EYE, LOOK, W, H).as_array()
            t0 = time.time()
            o = rst.render_pixels(fs, cam12, W, H, pw, ph, fid, n_threads=threads)
            dt = time.time() - t0
            simd = dict(value=(n_pixels + int((o.tri >= 0).sum()) * L) / dt / 1e6, unit="Mrays/s", cores=threads, kind="port",
                        seconds=dt, simd="AVX2, 8 triangles per step", sample="same pixels as cpu_baseline")
        rst.set_simd(False)
    except Exception as e:  # the comparator is optional
        simd = {"unavailable": str(e)[:200]}
    return dict(value=rays / secs / 1e6, unit="Mrays/s", cores=threads, kind=kind, seconds=secs, simd_comparator=simd,
                sample=f"{n_pixels} random pixels of the {W}x{H} frame (all {scene.n_tris} triangles, {L} lights; spheres omitted: "
                       f"the reference has none), {'reference intersect()/occlusion() via oracle/_ref' if kind == 'reference' else 'oracle/restated.c'}, "
                       f"{threads} threads")


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation on the host cores, same metric/config."""
    if rank != 0:
        return
    scene, W, H = make_scene(args.workload, args)
    threads = os.cpu_count() or 1
    per_px = scene.n_tris * 3.0 / 55e6  # ~55 M tests/s/core
    n_px = int(max(threads, min(W * H, (20.0 * threads / max(1, args.steps + args.warmup)) / max(per_px, 1e-6))))
    for _ in range(args.warmup):
        cpu_sample(scene, W, H, max(threads, n_px // 4), threads)
    vals, secs = [], 0.0
    for _ in range(args.steps):
        c = cpu_sample(scene, W, H, n_px, threads)
        vals.append(c["value"]), secs
        secs += c["seconds"]
    v = float(np.mean(vals))
    c["value"] = v
    print(json.dumps({
        "impl": "reference", "metric": "Mrays/s (primary+shadow)", "value": v, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, scene, W, H), "cpu_baseline": c,
        "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(name, scene, W, H):
    kind = f"reference model {GOLDEN_MODELS[name][:-4]}" if name in GOLDEN_MODELS and scene.n_tris == WORKLOADS[name][0] else "synthetic"
    return {"workload": f"{name}: {kind} {scene.n_tris}-triangle + {len(scene.sphere_cr)}-sphere scene at {W}x{H}, "
                        f"{scene.n_lights} lights, brute force over all objects",
            "n_tris": scene.n_tris, "n_spheres": int(len(scene.sphere_cr)), "width": W, "height": H,
            "n_lights": scene.n_lights, "spp": 1, "l2_policy": "inputs larger than L2 are not needed: the per-frame working "
            "set is re-streamed every step and the ray workspace (>=400 MB at 4K) exceeds L2; no cached outputs",
            "rng": "counter-based hash (seeded)", "partition": "interleaved 8-row bands, scene replicated"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--tris", type=int, default=0)
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cull", action="store_true", help="skip the extra measurement of the optional bundle-cull mode")
    ap.add_argument("--mode", default="brute", choices=["brute", "cull"],
                    help="which mode `value` measures; the default is the north star's brute-force formulation")
    ap.add_argument("--spp", type=int, default=0, help="samples per pixel (extension); default: the workload's")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank)

    import torch
    import torch.distributed as dist

    from esctp1raytracer_b200 import Camera, Renderer
    from esctp1raytracer_b200 import dist as tdist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    renderer = Renderer(local_rank)
    scene, W, H = make_scene(args.workload, args)
    cam = Camera.for_frame(EYE, LOOK, W, H)
    seed = 42
    spp = args.spp or WORKLOAD_SPP.get(args.workload, 0)
    main_cull = args.mode == "cull"
    import functools
    tdist_render = functools.partial(tdist.render_frame, samples_per_pixel=spp)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- FP32 peak of this GPU, measured now ----------------------------------------------
    # variants 0/3 scalar FFMA chains, 1 packed FFMA2 chains.  (Variant 2, an instruction-mix kernel, is not a
    # peak: the compiler hoists part of its FFMAs, so its flop count over-states what was executed.)
    peaks = {v: renderer.fp32_peak(v, 5)[0] for v in (0, 1, 3)}
    peak_tflops = max(peaks.values())
    info = renderer.device_info()
    nominal = info["sm_count"] * 128 * 2 * info["clock_khz"] * 1e3 / 1e12

    # ---- value: resident scene, frame stays in HBM ------------------------------------------
    rs = renderer.upload(scene)
    for _ in range(args.warmup):
        tdist_render(renderer, rs, cam, W, H, rank=rank, world=world, seed=seed, bundle_cull=main_cull)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    acc = {}
    barrier()
    ev0.record()
    for _ in range(args.steps):
        frame, st = tdist_render(renderer, rs, cam, W, H, rank=rank, world=world, seed=seed, bundle_cull=main_cull)
        for k, v in st.items():
            acc[k] = acc.get(k, 0) + v
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    flop_primary = float(st.get("flop_primary") or FLOP_PER_PAIR)  # executed flops per primary pair (same on every rank)
    flop_shadow = float(st.get("flop_shadow") or FLOP_PER_PAIR)
    keys = ["n_primary_rays", "n_shadow_rays", "tests_primary", "tests_shadow", "tests_shadow_ref", "strict_evals",
            "kernel_launches", "ms_primary", "ms_shadow", "ms_total"]
    t = torch.tensor([ms] + [float(acc.get(k, 0)) for k in keys], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms = float(mx[0])
        sweep_ms_max = float(mx[8] + mx[9])
    else:
        sweep_ms_max = float(t[8] + t[9])
    tot = {k: float(t[i + 1]) for i, k in enumerate(keys)}
    rays = tot["n_primary_rays"] + tot["n_shadow_rays"]
    value = rays / (ms * 1e-3) / 1e6
    launches = int(tot["kernel_launches"]) + (args.steps if world > 1 else 0)

    # ---- extra: the OPTIONAL bundle-cull mode (same frame, hierarchical evaluation of the same filter) ------
    cull_extra = None
    if not args.no_cull and not main_cull:
        frame_ref, _ = tdist_render(renderer, rs, cam, W, H, rank=rank, world=world, seed=seed)
        for _ in range(2):
            frame_c, _ = tdist_render(renderer, rs, cam, W, H, rank=rank, world=world, seed=seed, bundle_cull=True)
        same = bool(torch.equal(frame_ref, frame_c)) if rank == 0 else True
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(args.steps):
            tdist_render(renderer, rs, cam, W, H, rank=rank, world=world, seed=seed, bundle_cull=True)
        c1.record()
        barrier()
        cms = torch.tensor([c0.elapsed_time(c1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(cms, op=dist.ReduceOp.MAX)
        cull_extra = {"value": rays / (float(cms[0]) * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": float(cms[0]) / args.steps,
                      "frame_identical_to_default_mode": same,
                      "note": "opts.bundle_cull (two-phase): every (ray block, triangle) pair against the block box, survivors against "
                              "the 16 warp boxes -> sorted per-(block, warp) lists -> lane box -> per-ray filter -> strict pairs merged "
                              "by atomicMin; no acceleration structure, results bit-identical.  "
                              "Reported beside the headline, which stays on the brute-force per-ray formulation of the north star."}

    # ---- e2e: the drop-in call with host buffers --------------------------------------------
    e2e = None
    if not args.no_e2e:
        h2d = int(scene.tri_verts.nbytes + (scene.tri_normals.nbytes if scene.tri_normals is not None else 0)
                  + scene.geom_material.nbytes + scene.geom_tri_offset.nbytes + scene.sphere_cr.nbytes
                  + scene.sphere_material.nbytes + 48)
        host_frame = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory() if rank == 0 else None

        def e2e_step():
            ta = time.perf_counter()
            r2 = renderer.upload(scene)  # host -> HBM every step
            tb = time.perf_counter()
            fr, _ = tdist_render(renderer, r2, cam, W, H, rank=rank, world=world, seed=seed, bundle_cull=main_cull)
            tc = time.perf_counter()
            if rank == 0:
                host_frame.copy_(fr, non_blocking=False)  # HBM -> host read of the result
            td = time.perf_counter()
            r2.close()
            if os.environ.get("BENCH_DEBUG") and rank == 0:
                print(f"e2e_step: upload {tb-ta:.3f} render {tc-tb:.3f} readback {td-tc:.3f} close {time.perf_counter()-td:.3f}", file=sys.stderr)

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": rays / float(dt[0]) / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": W * H * 3, "ms_per_step": float(dt[0]) / args.steps * 1e3,
               "includes": "scene upload + filter-table build + render + band gather + frame read-back"}

    if rank == 0:
        alg_pairs = tot["tests_primary"] + tot["tests_shadow_ref"]
        swept_pairs = tot["tests_primary"] + tot["tests_shadow"]
        sweep_s = sweep_ms_max * 1e-3
        alg_flop = flop_primary * tot["tests_primary"] + flop_shadow * tot["tests_shadow_ref"]
        swept_flop = flop_primary * tot["tests_primary"] + flop_shadow * tot["tests_shadow"]
        achieved = alg_flop / world / sweep_s / 1e12  # per GPU
        hbm_peak = None
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            pass
        prim_bytes = 48.0 * scene.n_tris * (1 + scene.n_lights * 2) + 36.0 * scene.n_tris
        hbm_gbs = (prim_bytes + 3.0 * W * H / world) * args.steps / (ms * 1e-3) / 1e9
        traffic = traffic_detail = None  # dram bytes per launch of the dominant kernel, from the committed ncu --set full capture
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
            traffic = tj["traffic_bytes_per_launch"]
            traffic_detail = {"algorithmic_bytes_per_launch": tj["algorithmic_bytes_per_launch"], "kernel": tj["kernel"],
                              "capture_config": tj["config_of_capture"], "source": tj["capture"], "note": tj.get("algorithmic_note")}
        except Exception:
            pass
        cpu = None
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            per_px = scene.n_tris * 3.0 / 55e6
            cpu = cpu_sample(scene, W, H, int(max(threads, min(W * H, 15.0 * threads / max(per_px, 1e-6)))), threads)
        line = {
            "metric": "Mrays/s (primary+shadow)", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": dict(workload_config(args.workload, scene, W, H), spp=max(1, spp), mode=args.mode),
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "roofline": {
                "bound": "fp32", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved / peak_tflops,
                "traffic": traffic, "traffic_detail": traffic_detail,
                "peak_source": "own FFMA/FFMA2 microbenchmark in this process (tracer_cuda_fp32_peak); MEASURED_PEAKS.json has no FP32 figure",
                "peak_nominal": nominal, "frac_of_nominal": achieved / nominal, "peak_variants_tflops": peaks,
                "flop_per_pair": {"primary": flop_primary, "shadow": flop_shadow,
                                  "note": "flops the sweeps execute per (ray, triangle) pair.  Three 2-D affine edge rows = 6 FFMA = 12 when "
                                          "each ray evaluates its own; both sweeps compute the q-term of each row once per thread: closest hit "
                                          "(3 + 3*8) FFMA / 8 pairs = 6.75 (8 rays of one image row share q exactly), shadow (6 + 3*8) / 8 = 7.5 "
                                          "(8 consecutive rays of the q-sorted list: mean q plus |B| * spread, still a necessary condition)"},
                "algorithmic_pairs_per_step": alg_pairs / args.steps,
                "swept_pairs_per_step": swept_pairs / args.steps, "sweep_ms_per_step": sweep_ms_max / args.steps,
                "executed_tflops": swept_flop / world / sweep_s / 1e12,
                "primary_ms_per_step": tot["ms_primary"] / world / args.steps, "shadow_ms_per_step": tot["ms_shadow"] / world / args.steps,
                "primary_tflops": flop_primary * tot["tests_primary"] / (tot["ms_primary"] * 1e-3) / 1e12 if tot["ms_primary"] else None,
                "shadow_tflops": flop_shadow * tot["tests_shadow"] / (tot["ms_shadow"] * 1e-3) / 1e12 if tot["ms_shadow"] else None,
                "pair_rate_tpairs_s": swept_pairs / world / sweep_s / 1e12,
                "reference_formulation_tflops": FLOP_PER_PAIR_REF * alg_pairs / world / sweep_s / 1e12,
                "pair_formulation_view": {
                    "flop_per_pair": FLOP_PER_PAIR, "tflops": FLOP_PER_PAIR * alg_pairs / world / sweep_s / 1e12,
                    "frac_of_measured_peak": FLOP_PER_PAIR * alg_pairs / world / sweep_s / 1e12 / peak_tflops,
                    "note": "NOT roofline.frac: what the rate would be called if every pair were charged the 12 flop of an "
                            "independent evaluation of its three edge rows; the closest-hit sweep avoids 5.25 of them per pair"},
                "ceiling_note": "every instruction mix tried issues at IPC ~0.8 per scheduler.  One q per ray: 6 FFMA + 1.5 LOP3 + 0.5 LDS/SHF = 8 instr "
                                "per pair (0.58 of nominal FMA peak).  Shared q-terms: 3.4-3.75 FFMA + 1.1-1.5 LOP3 + 0.5 = 5-5.75 instr per pair: 1.3-1.4x "
                                "faster per pair, but only ~2/3 of the instructions are FFMA, so the executed-flop fraction is lower "
                                "(tools/sweep_mb.cu, DESIGN.md 4)",
                "hbm": {"achieved_gbs": hbm_gbs, "peak_gbs": hbm_peak, "frac": (hbm_gbs / hbm_peak) if hbm_peak else None,
                        "streams": "filter tables (48 B/triangle/origin) + vertices (36 B/triangle) + framebuffer (3 B/pixel)"},
            },
            "cpu_baseline": cpu, "optional_bundle_cull_mode": cull_extra,
            "rays_per_step": rays / args.steps, "strict_evals_per_step": tot["strict_evals"] / args.steps,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
