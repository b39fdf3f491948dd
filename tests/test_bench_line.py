"""bench.py's own arm, dry-run on the CPU: the GPU-facing calls (Renderer, dist.render_frame, torch.cuda events / pinned
memory) are replaced by stand-ins that return plausible frame statistics, everything else — argument handling, the timed
loop, the roofline arithmetic, the JSON line — is bench.py's real code.  Asserts that the line carries every key the
measurement contract names and that the derived numbers follow from the statistics.  (The numbers themselves are only
meaningful on the B200: tests/test_gpu_*.py and the committed profiles/ hold those.)
"""
import io
import json
import os
import sys
import time
from contextlib import redirect_stdout

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

PAIRS_PRIMARY, PAIRS_SHADOW, PAIRS_SHADOW_REF = 6_144_000, 2_000_000, 1_900_000
MS_PRIMARY, MS_SHADOW = 3.0, 1.0


class FakeEvent:
    def __init__(self, enable_timing=True):
        self.t = None

    def record(self):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


class FakeFrame:
    def __init__(self, stats, rgb8):
        self.stats, self.rgb8 = stats, rgb8


def make_stats(W, H, L):
    hit = W * H // 2
    return {"n_pixels": W * H, "n_primary_rays": W * H, "n_shadow_rays": hit * L, "tests_primary": PAIRS_PRIMARY,
            "tests_shadow": PAIRS_SHADOW, "tests_shadow_ref": PAIRS_SHADOW_REF, "strict_evals": 12345, "filter_misses": 0,
            "pipeline_errors": 0, "kernel_launches": 21, "ms_primary": MS_PRIMARY, "ms_shadow": MS_SHADOW, "ms_other": 0.1,
            "ms_total": MS_PRIMARY + MS_SHADOW + 0.1, "flop_primary": 4.25, "flop_shadow": 5.0, "flop_primary_edges": 0.25,
            "flop_shadow_edges": 1.0, "n_sms": 148}


class FakeResident:
    def close(self):
        pass


class FakeRenderer:
    def __init__(self, device=0):
        self.device = device

    def fp32_peak(self, variant, iters):
        return {0: 55.0, 1: 68.0, 3: 47.0}[variant], 1.0

    def device_info(self):
        return {"sm_count": 148, "clock_khz": 1_965_000, "name": "stand-in"}

    def upload(self, scene):
        return FakeResident()

    def trace(self, scene, cam, W, H, out=None, **kw):
        if out is not None:
            out[...] = 7
        return FakeFrame(make_stats(W, H, 4), out)


@pytest.fixture
def dry_bench(monkeypatch):
    import torch

    import bench
    import esctp1raytracer_b200 as pkg
    from esctp1raytracer_b200 import dist as tdist

    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    real_tensor, real_empty = torch.tensor, torch.empty
    monkeypatch.setattr(torch, "tensor", lambda *a, **k: real_tensor(*a, **{x: y for x, y in k.items() if x != "device"}))
    monkeypatch.setattr(torch, "empty", lambda *a, **k: real_empty(*a, **{x: y for x, y in k.items() if x != "device"}))
    monkeypatch.setattr(pkg, "Renderer", FakeRenderer)

    def render_frame(renderer, rs, cam, W, H, **kw):
        time.sleep(0.002)
        return torch.full((H, W, 3), 7, dtype=torch.uint8), make_stats(W, H, 4)

    monkeypatch.setattr(tdist, "render_frame", render_frame)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)

    def run(*argv):
        monkeypatch.setattr(sys, "argv", ["bench.py", *argv])
        buf = io.StringIO()
        with redirect_stdout(buf):
            bench.main()
        lines = [l for l in buf.getvalue().splitlines() if l.startswith("{")]
        assert len(lines) == 1, buf.getvalue()
        return json.loads(lines[0])

    return run


def test_bench_line_has_the_contract_keys_and_consistent_arithmetic(dry_bench):
    d = dry_bench("--workload", "small", "--tris", "3000", "--width", "64", "--height", "48", "--steps", "4", "--warmup", "3",
                  "--no-cpu-baseline")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline", "frame_sha256"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 4 and d["warmup"] == 3 and d["higher_is_better"] is True
    assert d["dtype"] == "f32" and d["vs_baseline"] is None and d["scaling"] == "strong" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"] and d["config"]["mode"] == "brute"
    assert "l2_policy" in d["config"]
    assert d["gpu_launches"] == 21 * 4
    rays = 64 * 48 + 64 * 48 // 2 * 4
    assert d["rays_per_step"] == rays
    assert d["value"] == pytest.approx(rays / (d["ms_per_step"] * 1e-3) / 1e6)
    e = d["e2e"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in e, k
    assert e["d2h_bytes_per_step"] == 64 * 48 * 3 and e["h2d_bytes_per_step"] > 3000 * 36
    assert "tracer_cuda_render" in e["path"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["peak"] == 68.0 and r["unit"] == "TFLOP/s"
    sweep_s = (MS_PRIMARY + MS_SHADOW) * 1e-3
    want = (4.25 * PAIRS_PRIMARY + 5.0 * PAIRS_SHADOW_REF) / sweep_s / 1e12  # algorithmic pairs, per launch-time of both sweeps
    assert r["achieved"] == pytest.approx(want) and r["frac"] == pytest.approx(want / 68.0)
    assert r["primary_tflops"] == pytest.approx(4.25 * PAIRS_PRIMARY / (MS_PRIMARY * 1e-3) / 1e12)
    assert r["shadow_tflops"] == pytest.approx(5.0 * PAIRS_SHADOW / (MS_SHADOW * 1e-3) / 1e12)  # swept pairs here
    dk = r["dominant_kernel"]
    assert dk["kernel"].startswith("trk::primary_kernel")
    assert dk["achieved"] == pytest.approx(r["primary_tflops"]) and dk["frac"] == pytest.approx(r["primary_tflops"] / 68.0)
    assert dk["share_of_sweep_time"] == pytest.approx(MS_PRIMARY / (MS_PRIMARY + MS_SHADOW))
    assert r["as_issued"]["flop_per_pair"] == {"primary": 6.25, "shadow": 7.0}
    assert r["fma_pipe"]["lane_ops_per_pair"] == {"primary": 3.125, "shadow": 3.5}
    assert r["mix_ceiling_frac"] == pytest.approx((4.25 * PAIRS_PRIMARY + 5.0 * PAIRS_SHADOW_REF) /
                                                  (2 * (3.125 * PAIRS_PRIMARY + 3.5 * PAIRS_SHADOW_REF)))
    assert d["clocks"]["reasons"] in (["nvidia-smi unavailable"], []) or isinstance(d["clocks"]["reasons"], list)
    oc = d["optional_bundle_cull_mode"]
    assert oc["frame_identical_to_default_mode"] is True


def test_bench_refuses_to_run_without_a_gpu(monkeypatch):
    import bench

    monkeypatch.setattr(sys, "argv", ["bench.py", "--no-cpu-baseline"])
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    with pytest.raises(SystemExit) as ei:
        bench.main()
    assert "no CPU fallback" in str(ei.value)


@pytest.mark.parametrize("argv", [
    ("--workload", "c1", "--steps", "3", "--warmup", "3", "--no-cpu-baseline"),
    ("--workload", "c3", "--steps", "2", "--warmup", "3", "--no-cpu-baseline", "--no-e2e", "--no-cull"),
    ("--workload", "small", "--tris", "2000", "--width", "32", "--height", "24", "--mode", "cull", "--no-cpu-baseline"),
    ("--workload", "c5", "--tris", "2000", "--width", "32", "--height", "24", "--no-cpu-baseline", "--no-cull"),
])
def test_bench_flag_combinations_produce_one_line(dry_bench, argv):
    d = dry_bench(*argv)
    assert d["metric"].startswith("Mrays/s") and d["value"] > 0 and d["roofline"]["frac"] > 0
    if "--no-e2e" in argv:
        assert d["e2e"] is None
    if "c5" in argv:
        assert d["config"]["spp"] == 16
