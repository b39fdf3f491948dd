"""World-size-2 (and 3) CPU test of the multi-GPU host logic: band partition, equal-size
padding, the gather to rank 0 and the row mapping used to reassemble (gloo backend)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from esctp1raytracer_b200 import band_row_count
from esctp1raytracer_b200.dist import band_rows_of_rank, gather_bands, padded_rows


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _frame(W, H):
    pr = np.arange(H, dtype=np.int64)[:, None, None]
    w = np.arange(W, dtype=np.int64)[None, :, None]
    c = np.arange(3, dtype=np.int64)[None, None, :]
    return ((pr * 7 + w * 3 + c * 11) % 251).astype(np.uint8)


def _worker(rank, world, port, W, H, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = _frame(W, H)
        rows = band_rows_of_rank(H, B, rank, world)
        assert len(rows) == band_row_count(H, B, rank, world)  # host mirror == C library
        pad = padded_rows(H, B, world)
        local = torch.zeros((pad, W, 3), dtype=torch.uint8)
        local[: len(rows)] = torch.from_numpy(full[rows])  # what the renderer would have produced
        g = gather_bands(local, rank, world, 0)
        if rank == 0:
            out = np.zeros_like(full)
            for r in range(world):
                rr = band_rows_of_rank(H, B, r, world)
                out[rr] = g[r, : len(rr)].numpy()
            q.put(bool(np.array_equal(out, full)))
        else:
            assert g is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,W,H,B", [(2, 33, 77, 8), (3, 16, 50, 4), (2, 8, 5, 8)])
def test_band_gather_reassembles_frame(world, W, H, B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, W, H, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_partition_covers_every_row_once():
    for H, B, n in [(2160, 8, 8), (77, 8, 2), (10, 4, 3), (5, 8, 2), (4320, 16, 4)]:
        rows = np.concatenate([band_rows_of_rank(H, B, r, n) for r in range(n)])
        assert sorted(rows.tolist()) == list(range(H))
        assert padded_rows(H, B, n) == max(band_row_count(H, B, r, n) for r in range(n))
