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


def _exchange_worker(rank, world, port, height, win, stride, nx, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from instageo_b200 import ops
    ys = ops.window_origins(height, win, stride, True)
    ny = len(ys)
    full = torch.arange(ny * nx * 2 * 3 * 3, dtype=torch.float32).reshape(ny * nx, 2, 3, 3)   # "logits" of every window
    lo, hi = IU.partition(ny, world, rank)
    got = IU.exchange_window_rows(full[lo * nx:hi * nx].clone(), ys, win, height, nx, rank, world)
    a, b = IU.windows_for_rows(ys, win, *IU.stripe_rows(height, world, rank))
    q.put((rank, bool(torch.equal(got, full[a * nx:b * nx])), (a, b), (lo, hi)))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,height,stride", [(2, 3660, 112), (3, 3660, 224), (4, 1000, 112)])
def test_window_row_exchange_gloo(world, height, stride):
    """every rank ends up with exactly the window rows that cover its output stripe, whoever computed them"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() + world * 7 + stride) % 300
    procs = [ctx.Process(target=_exchange_worker, args=(r, world, port, height, 224, stride, 3, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(30)
    assert sorted(r[0] for r in res) == list(range(world))
    assert all(r[1] for r in res), res
    # the exchange is not vacuous: some rank needs rows it does not own
    assert any(r[2][0] < r[3][0] or r[2][1] > r[3][1] for r in res)


@pytest.mark.parametrize("H,W,win", [(3660, 3660, 224), (448, 672, 224), (500, 224, 224), (224, 700, 224), (70, 53, 16)])
def test_scatter_window_masks_equals_per_window_or(H, W, win):
    ty, tx = ops.window_origins(H, win, win, True), ops.window_origins(W, win, win, True)
    g = torch.Generator().manual_seed(H + W)
    pix = torch.rand((H, W), generator=g) < 0.3                      # the mask is a function of the pixel
    m = torch.stack([pix[t:t + win, l:l + win] for t in ty for l in tx])
    want = torch.zeros((H, W), dtype=torch.bool)
    for i, (t, l) in enumerate((t, l) for t in ty for l in tx):
        want[t:t + win, l:l + win] |= m[i]
    got = IU.scatter_window_masks(m, len(ty), len(tx), H, W, win)
    assert torch.equal(got, want) and torch.equal(got, pix)


@pytest.mark.parametrize("height,stride", [(3660, 112), (3660, 224), (1000, 112), (224, 224), (700, 160), (5000, 64)])
def test_exchange_plan_is_consistent(height, stride):
    """for every world size 1..8: each rank's needed rows = local part + disjoint received parts, every send has the
    matching receive on the peer, nobody sends a row it does not own, and every window row is computed exactly once"""
    ys = ops.window_origins(height, 224, stride, True)
    for world in range(1, 9):
        plan = IU.exchange_plan(ys, 224, height, world)
        owned = []
        for r, (need, own, local, sends, recvs) in enumerate(plan):
            owned += list(range(*own))
            got = set(range(*local)) if local else set()
            for q, (lo, hi) in recvs.items():
                rows = set(range(lo, hi))
                assert not (rows & got) and plan[q][3][r] == (lo, hi)          # disjoint, and q sends exactly that to r
                assert plan[q][1][0] <= lo and hi <= plan[q][1][1]             # q owns what it sends
                got |= rows
            assert got == set(range(*need)), (world, r)
            for q, rng in sends.items():
                assert plan[q][4][r] == rng
        assert sorted(owned) == list(range(len(ys)))
        if world == 1:
            assert plan[0][3] == {} and plan[0][4] == {}


@pytest.mark.parametrize("height,stride", [(3660, 112), (3660, 224), (1000, 112), (224, 224), (700, 160)])
def test_tile_engine_window_plan(height, stride):
    """The engine's plan on FLAT window indices: equal output stripes (one all-gather slot each), window shares balanced
    to one window, and for every world size each rank's needed windows = its local part + disjoint received parts
    whose sends exist on the peer -- every window computed exactly once."""
    win = 224
    ys = ops.window_origins(height, win, stride, True)
    nx = len(ys)
    n_win = len(ys) * nx
    for world in range(1, 9):
        stripes = IU.even_stripes(height, world)
        rows = -(-height // world)
        assert stripes[0][0] == 0 and max(b for _, b in stripes) == height
        assert all(a == min(height, r * rows) and b - a <= rows for r, (a, b) in enumerate(stripes))
        assert all(stripes[i][1] == stripes[i + 1][0] for i in range(world - 1))
        need = [tuple(v * nx for v in IU.windows_for_rows(ys, win, a, b)) if b > a else (0, 0) for a, b in stripes]
        own = [IU.partition(n_win, world, q) for q in range(world)]
        sizes = [b - a for a, b in own]
        assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n_win
        plan = IU.interval_plan(need, own)
        for r, (local, sends, recvs) in enumerate(plan):
            got = set(range(*local)) if local else set()
            for q, (lo, hi) in recvs.items():
                units = set(range(lo, hi))
                assert not (units & got) and plan[q][1][r] == (lo, hi)
                assert own[q][0] <= lo and hi <= own[q][1]
                got |= units
            assert got == set(range(*need[r])), (world, r)
            for q, rng in sends.items():
                assert plan[q][2][r] == rng and own[r][0] <= rng[0] and rng[1] <= own[r][1]


def _flat_exchange_worker(rank, world, port, height, stride, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    win = 224
    ys = ops.window_origins(height, win, stride, True)
    nx = len(ys)
    n_win = len(ys) * nx
    full = torch.arange(n_win * 4, dtype=torch.float32).reshape(n_win, 2, 2)       # "logits" of every window
    stripes = IU.even_stripes(height, world)
    need = [tuple(v * nx for v in IU.windows_for_rows(ys, win, a, b)) if b > a else (0, 0) for a, b in stripes]
    own = [IU.partition(n_win, world, r) for r in range(world)]
    _, sends, recvs = IU.interval_plan(need, own)[rank]
    u0, u1 = min(need[rank][0], own[rank][0]), max(need[rank][1], own[rank][1])
    buf = torch.full((u1 - u0, 2, 2), -1.0)
    buf[own[rank][0] - u0:own[rank][1] - u0] = full[own[rank][0]:own[rank][1]]       # what the model wrote in place
    for req in IU._exchange(buf, u0, buf, u0, sends, recvs, world):
        req.wait()
    n0, n1 = need[rank]
    q.put((rank, bool(torch.equal(buf[n0 - u0:n1 - u0], full[n0:n1])), len(sends) + len(recvs)))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,height,stride", [(2, 3660, 224), (3, 1000, 112)])
def test_flat_window_exchange_gloo(world, height, stride):
    """in-place exchange on one union buffer (own windows written by the model, needed ones received): every rank ends
    up with exactly the windows that cover its output stripe"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() + world * 11 + stride) % 90
    procs = [ctx.Process(target=_flat_exchange_worker, args=(r, world, port, height, stride, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(30)
    assert sorted(r[0] for r in res) == list(range(world))
    assert all(r[1] for r in res), res
    assert any(r[2] > 0 for r in res)
