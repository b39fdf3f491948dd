"""Stand-ins for the GPU-facing calls of bench.py's own arm (used by tests/test_bench_line.py in-process and by
tests/bench_dryrun_worker.py under torchrun + gloo): Renderer, dist.render_frame, torch.cuda events / pinned memory.
Everything else — argument handling, the timed loop, reductions over ranks, roofline arithmetic, the JSON line — stays
bench.py's real code."""
import time

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


def make_stats(W, H, L, share=1):
    """statistics of one rank's share (1/share of the frame)"""
    hit = W * H // 2
    return {"n_pixels": W * H // share, "n_primary_rays": W * H // share, "n_shadow_rays": hit * L // share,
            "tests_primary": PAIRS_PRIMARY // share, "tests_shadow": PAIRS_SHADOW // share, "tests_shadow_ref": PAIRS_SHADOW_REF // share,
            "strict_evals": 12345, "filter_misses": 0, "pipeline_errors": 0, "kernel_launches": 21, "ms_primary": MS_PRIMARY / share,
            "ms_shadow": MS_SHADOW / share, "ms_other": 0.1, "ms_total": (MS_PRIMARY + MS_SHADOW) / share + 0.1, "flop_primary": 4.25,
            "flop_shadow": 5.0, "flop_primary_edges": 0.25, "flop_shadow_edges": 1.0, "n_sms": 148}


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


def install(setattr_):
    """setattr_(obj, name, value): monkeypatch.setattr in a test, plain setattr in the worker"""
    import torch

    import esctp1raytracer_b200 as pkg
    from esctp1raytracer_b200 import dist as tdist

    setattr_(torch.cuda, "is_available", lambda: True)
    setattr_(torch.cuda, "set_device", lambda d: None)
    setattr_(torch.cuda, "synchronize", lambda *a: None)
    setattr_(torch.cuda, "Event", FakeEvent)
    setattr_(torch.Tensor, "pin_memory", lambda self: self)
    real_tensor, real_empty = torch.tensor, torch.empty
    setattr_(torch, "tensor", lambda *a, **k: real_tensor(*a, **{x: y for x, y in k.items() if x != "device"}))
    setattr_(torch, "empty", lambda *a, **k: real_empty(*a, **{x: y for x, y in k.items() if x != "device"}))
    setattr_(pkg, "Renderer", FakeRenderer)

    def render_frame(renderer, rs, cam, W, H, rank=0, world=1, **kw):
        time.sleep(0.002)
        frame = torch.full((H, W, 3), 7, dtype=torch.uint8) if rank == 0 else None
        return frame, make_stats(W, H, 4, share=world)

    setattr_(tdist, "render_frame", render_frame)
