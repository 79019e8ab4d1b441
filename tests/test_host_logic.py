"""CPU: host-side logic of the mirror -- window grids, sharding, and the N>1 gather over gloo."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from instageo_b200 import ops
from instageo_b200.model import infer_utils as IU
from oracle import preprocess as OP


def test_window_grid_matches_oracle():
    for (h, w, c, s, e) in [(512, 512, 224, 224, False), (512, 512, 224, 112, False), (3660, 3660, 224, 112, True),
                            (300, 340, 64, 48, True), (224, 224, 224, 224, True)]:
        assert ops.window_grid(h, w, c, s, e) == OP.window_grid(h, w, c, s, e)


def test_partition_covers_everything():
    for n in (0, 1, 7, 289, 1024, 100000):
        for ws in (1, 2, 3, 8):
            parts = [IU.partition(n, ws, r) for r in range(ws)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(ws - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def test_windows_for_rows_halo():
    ys = ops.window_origins(3660, 224, 112, True)
    for ws in (2, 4, 8):
        covered = set()
        for r in range(ws):
            y0, y1 = IU.stripe_rows(3660, ws, r)
            lo, hi = IU.windows_for_rows(ys, 224, y0, y1)
            # every window row that touches the stripe is included, none that does not
            for i, t in enumerate(ys):
                assert (lo <= i < hi) == (t < y1 and t + 224 > y0)
            covered |= set(range(lo, hi))
        assert covered == set(range(len(ys)))


def _gather_worker(rank, world, port, height, width, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = (torch.arange(height * width, dtype=torch.int64).reshape(height, width) % 120).to(torch.int8)
    y0, y1 = IU.stripe_rows(height, world, rank)
    got = IU.gather_stripes(full[y0:y1].clone(), height, world)
    chips = (torch.arange(7 * 4 * 4).reshape(7, 4, 4) % 100).to(torch.int8)
    lo, hi = IU.partition(7, world, rank)
    got2 = IU.gather_chip_masks(chips[lo:hi].clone(), 7, world)
    q.put((rank, bool(torch.equal(got, full)), bool(torch.equal(got2, chips))))
    dist.destroy_process_group()


def test_gather_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, 37, 5, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(30)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] for r in res)
