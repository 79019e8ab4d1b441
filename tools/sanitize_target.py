"""Small-shape pass over every kernel family for compute-sanitizer (tools/sanitize.sh): preprocess, nodata map,
tcgen05 GEMM (all epilogues through a tiny PrithviSeg forward), attention, LayerNorm, stitch (both paths), chip mask,
metrics.  Shapes are as small as the kernels allow: racecheck slows a launch by 10-100x."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import instageo_b200  # noqa: E402,F401
from instageo_b200 import ops  # noqa: E402
from instageo_b200.model import PrithviSeg  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
mean, std = [0.14, 0.13, 0.12, 0.31, 0.20, 0.12], [0.04, 0.04, 0.05, 0.08, 0.06, 0.05]
rng = np.random.default_rng(0)

if what in ("all", "pre"):
    raw = torch.from_numpy(rng.integers(0, 10001, size=(2, 6, 224, 224)).astype(np.int16)).to(dev)
    spec = ops.PreprocessSpec(mean, std, 1, None, 1e-4, -9999, dev)
    ops.preprocess(raw, spec, want_f32=True, want_patches=True, want_mask_elem=True, want_mask_px=True)
    tile = torch.from_numpy(rng.integers(0, 10001, size=(6, 300, 333)).astype(np.int16)).to(dev)
    spec1 = ops.PreprocessSpec(mean, std, 1, None, 1.0, -9999, dev)
    ops.nodata_map(tile, spec1, 7, 290)
    print("pre ok")
if what in ("all", "ops"):
    a = torch.randn(300, 256, device=dev).bfloat16()
    w = torch.randn(512, 256, device=dev).bfloat16()
    ops.linear(a, w, torch.randn(512, device=dev), act=1)
    ops.linear(a, w, torch.randn(512, device=dev), resid=torch.randn(300, 512, device=dev), out_dtype=torch.float32)
    ops.layernorm(torch.randn(77, 256, device=dev), torch.ones(256, device=dev), torch.zeros(256, device=dev))
    qkv = torch.randn(2 * 197, 3 * 128, device=dev).bfloat16()
    ops.attention(qkv, 2, 197, 2)
    print("ops ok")
if what in ("all", "model"):
    torch.manual_seed(0)
    m = PrithviSeg(temporal_step=1, num_classes=2, load_pretrained_weights=False, variant="prithvi_eo_tiny", depth=1,
                   embed_dims=[256, 64, 32, 16, 16]).to(dev).eval()
    x = torch.randn(1, 6, 1, 224, 224, device=dev)
    m(x)
    m.predict(x)
    print("model ok")
if what in ("all", "stitch"):
    for nc, path in ((2, "direct"), (5, "tma")):
        os.environ["IG_STITCH_PATH"] = path
        H, W, win, stride = 150, 200, 64, 40
        ys, xs = ops.window_origins(H, win, stride, True), ops.window_origins(W, win, stride, True)
        lg = torch.randn(len(ys) * len(xs), nc, win, win, device=dev)
        nd = torch.rand(H, W, device=dev) < 0.1
        ops.stitch(lg, ys, xs, H, W, nodata_px=nd, want_avg=True, want_hist=True)
    print("stitch ok")
if what in ("all", "aux"):
    from instageo_b200.data import create_chip
    from instageo_b200.model.metrics import RunningAUC, RunningConfusionMatrix, segmentation_eval_update
    chip = rng.integers(0, 12000, size=(6, 64, 72)).astype(np.int16)
    fm = rng.integers(0, 255, size=(1, 64, 72)).astype(np.uint8)
    create_chip(chip, fm, rng.integers(0, 2, size=(64, 72)).astype(np.int8))
    cm, auc = RunningConfusionMatrix(3, device=dev), RunningAUC(3, n_bins=64, device=dev)
    segmentation_eval_update(torch.randn(2, 3, 16, 16, device=dev), torch.randint(0, 3, (2, 16, 16), device=dev), cm, auc)
    print("aux ok")
torch.cuda.synchronize()
print("SANITIZE TARGET DONE", what)
