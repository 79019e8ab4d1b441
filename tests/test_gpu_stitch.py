"""GPU parity: kernel 5 (overlap-averaging stitch) vs oracle/stitch.py -- bit-identical averages and
class maps on identical window logits -- plus size-independent properties at the full 3660^2 tile."""
import numpy as np
import pytest
import torch

from oracle import stitch as OS

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["auto", "tma", "direct"])
def stitch_path(request, monkeypatch):
    """Both kernels behind ig_stitch (TMA-staged tiles / direct loads) must agree with the oracle on every input."""
    if request.param != "auto":
        monkeypatch.setenv("IG_STITCH_PATH", request.param)
    return request.param


@pytest.mark.parametrize("H,W,win,stride,nc", [(300, 340, 64, 32, 2), (500, 470, 224, 112, 13), (256, 256, 64, 64, 3),
                                               (130, 257, 64, 40, 1), (64, 64, 64, 64, 2), (333, 517, 224, 75, 17),
                                               (90, 1100, 32, 24, 5)])
def test_bit_identical_to_oracle(cuda_dev, stitch_path, H, W, win, stride, nc):
    from instageo_b200 import ops
    rng = np.random.default_rng(H + stride)
    ys, xs = ops.window_origins(H, win, stride, True), ops.window_origins(W, win, stride, True)
    org = [(t, l) for t in ys for l in xs]
    lg = rng.standard_normal((len(org), nc, win, win)).astype(np.float32)
    lg[0, :, :4, :4] = 0.5  # exact ties -> first max must win
    nd = rng.random((H, W)) < 0.03
    avg, cls = OS.stitch(lg, org, H, W, nd)
    out = ops.stitch(torch.from_numpy(lg).to(cuda_dev), ys, xs, H, W, nodata_px=torch.from_numpy(nd).to(cuda_dev),
                     want_avg=True, want_hist=True)
    assert np.array_equal(out["avg"].cpu().numpy(), avg)
    assert np.array_equal(out["class_map"].cpu().numpy(), cls)
    hist = out["hist"].cpu().numpy()
    assert [int(h) for h in hist] == [int((cls == k).sum()) for k in range(nc)] + [int((cls == -1).sum())]


def test_stripes_equal_full_and_halo(cuda_dev, stitch_path):
    """Row stripes computed from only the windows that touch them == the single-pass result."""
    from instageo_b200 import ops
    from instageo_b200.model import infer_utils as IU
    H, W, win, stride, nc = 700, 300, 224, 112, 2
    ys, xs = ops.window_origins(H, win, stride, True), ops.window_origins(W, win, stride, True)
    lg = torch.randn(len(ys) * len(xs), nc, win, win, device=cuda_dev)
    full = ops.stitch(lg, ys, xs, H, W)["class_map"]
    for ws in (2, 3, 8):
        parts = []
        for r in range(ws):
            y0, y1 = IU.stripe_rows(H, ws, r)
            lo, hi = IU.windows_for_rows(ys, win, y0, y1)
            sub = lg[lo * len(xs): hi * len(xs)].contiguous()
            parts.append(ops.stitch(sub, ys, xs, H, W, y0=y0, y1=y1, win_base=lo * len(xs))["class_map"])
        assert torch.equal(torch.cat(parts), full)
    assert ops.stitch(lg, ys, xs, H, W, y0=10, y1=10)["class_map"].shape == (0, W)
    # the nodata map of a stripe holds the stripe's own rows (what ig_nodata_map writes); the whole map is accepted too
    nd = torch.rand((H, W), device=cuda_dev) < 0.1
    full_nd = ops.stitch(lg, ys, xs, H, W, nodata_px=nd, nodata_class=-7)["class_map"]
    assert bool((full_nd[nd] == -7).all()) and torch.equal(full_nd[~nd], full[~nd])
    for (y0, y1) in ((0, 233), (233, 467), (467, 700), (5, 6)):
        lo, hi = IU.windows_for_rows(ys, win, y0, y1)
        sub = lg[lo * len(xs): hi * len(xs)].contiguous()
        a = ops.stitch(sub, ys, xs, H, W, y0=y0, y1=y1, win_base=lo * len(xs), nodata_px=nd[y0:y1].contiguous(),
                       nodata_class=-7)["class_map"]
        b = ops.stitch(sub, ys, xs, H, W, y0=y0, y1=y1, win_base=lo * len(xs), nodata_px=nd, nodata_class=-7)["class_map"]
        assert torch.equal(a, full_nd[y0:y1]) and torch.equal(b, full_nd[y0:y1])


def test_full_tile_properties(cuda_dev):
    """3660^2 tile (BASELINE config 4): properties that need no CPU oracle at this size."""
    from instageo_b200 import ops
    H = W = 3660
    win, nc = 224, 2
    for stride in (224, 112):
        ys, xs = ops.window_origins(H, win, stride, True), ops.window_origins(W, win, stride, True)
        assert len(ys) * len(xs) == (289 if stride == 224 else 1024)
        # every window carries the same per-class constants -> the average is that constant everywhere
        lg = torch.empty(len(ys) * len(xs), nc, win, win, device=cuda_dev)
        lg[:, 0] = 0.25
        lg[:, 1] = -1.5
        out = ops.stitch(lg, ys, xs, H, W, want_avg=True, want_hist=True)
        assert bool((out["avg"][0] == 0.25).all()) and bool((out["avg"][1] == -1.5).all())
        assert bool((out["class_map"] == 0).all()) and out["hist"].tolist() == [H * W, 0, 0]
        # linearity in the window logits: stitch(a) + stitch(b) == stitch(a + b) up to f32 rounding
        a, b = torch.randn_like(lg), torch.randn_like(lg)
        sa = ops.stitch(a, ys, xs, H, W, want_avg=True)["avg"]
        sb = ops.stitch(b, ys, xs, H, W, want_avg=True)["avg"]
        sab = ops.stitch(a + b, ys, xs, H, W, want_avg=True)["avg"]
        assert float((sa + sb - sab).abs().max()) < 1e-5
        if stride == 224:  # interior of non-overlapped windows is a plain mosaic
            assert torch.equal(sa[:, :224, :224], a[0])
