"""GPU parity: sliding-window inference over a raster tile (windows -> kernel 1 -> model -> kernel 5)
vs the same pipeline assembled from the oracles, and single-GPU == sharded stripes, bit for bit."""
import numpy as np
import pytest
import torch

from conftest import FLOOD_MEAN, FLOOD_STD
from oracle import preprocess as OP
from oracle import prithvi as P
from oracle import stitch as OS

pytestmark = pytest.mark.gpu


def test_tile_vs_oracle_and_sharding(cuda_dev):
    from instageo_b200.model import PrithviSeg, sliding_window_inference
    from instageo_b200.model import infer_utils as IU
    variant, T, nc, depth = "prithvi_eo_tiny", 1, 2, 1
    sd = P.make_state_dict(variant, T, nc, depth=depth, seed=21, stress=True)
    m = PrithviSeg(temporal_step=T, num_classes=nc, load_pretrained_weights=False, variant=variant, depth=depth)
    m.load_state_dict(sd)
    m.to(cuda_dev).eval()
    rng = np.random.default_rng(4)
    H, W, win, stride = 500, 460, 224, 112
    tile = rng.integers(0, 10001, size=(6, H, W)).astype(np.int16)
    tile[:, :40, :60] = -9999  # nodata wedge
    kw = dict(window_size=(win, win), stride=stride, batch_size=5, mean=FLOOD_MEAN, std=FLOOD_STD,
              constant_multiplier=1.0, no_data_value=-9999)
    got = sliding_window_inference(tile, m, **kw)
    assert got.dtype == np.int8 and got.shape == (H, W)
    # oracle pipeline on the same windows
    org = OS.tile_windows(H, W, win, stride)
    assert len(org) == 16
    xs = np.stack([OP.normalize(tile[:, t:t + win, l:l + win] * 1.0, FLOOD_MEAN, FLOOD_STD, T) for t, l in org])
    logits = P.prithvi_seg_forward(torch.from_numpy(xs), sd, 4, T).numpy()
    nd = OS.tile_nodata_px(tile, None, 1.0, -9999)
    avg, cls = OS.stitch(logits, org, H, W, nd)
    assert (got[nd] == -1).all() and (got[:40, :60] == -1).all()
    top2 = np.sort(avg, axis=0)[-2:]
    safe = ((top2[1] - top2[0]) > 4e-2) & ~nd
    assert safe.mean() > 0.5 and (got == cls)[safe].all()
    # sharded: each simulated rank computes only its stripe from only the windows that touch it
    for ws in (2, 4):
        parts = []
        for r in range(ws):
            y0, y1 = IU.stripe_rows(H, ws, r)
            parts.append(sliding_window_inference(tile, m, rows=(y0, y1), **kw))
        assert np.array_equal(np.concatenate(parts), got)
    # stride == window: plain mosaic of per-window argmax (the reference's non-overlapping case)
    got2 = sliding_window_inference(tile[:, :448, :448], m, window_size=(win, win), stride=224, batch_size=4,
                                    mean=FLOOD_MEAN, std=FLOOD_STD)
    x00 = torch.from_numpy(OP.normalize(tile[:, :224, :224] * 1.0, FLOOD_MEAN, FLOOD_STD, T))[None].to(cuda_dev)
    assert np.array_equal(got2[:224, :224], m.predict(x00)[0].cpu().numpy())


def test_sharded_tile_is_bit_identical_across_gpu_counts():
    """end-to-end tile, 1 vs N GPUs -- needs a multi-GPU box (skipped on one GPU; the same script,
    tools/check_tile_sharded.py, was run by hand at 2 and 8 GPUs: PASS)."""
    import os
    import subprocess
    import sys
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU: the N-GPU comparison needs torchrun over >= 2 devices")
    n = 2 if n < 4 else (4 if n < 8 else 8)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(root, "tools", "check_tile_sharded.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "SHARDED TILE CHECK PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
