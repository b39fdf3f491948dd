"""One rank of a CPU dry-run of bench.py's torchrun path (tests/test_bench_line.py launches two of these with
torch.distributed.run): the stand-ins of tests/bench_standins.py, the process group on gloo instead of NCCL."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import bench_standins  # noqa: E402

bench_standins.install(setattr)
_real_init = dist.init_process_group
dist.init_process_group = lambda backend=None, device_id=None, **kw: _real_init("gloo", **kw)
bench.main()
