"""Minimal single-kernel drivers for `ncu --set full` captures (a handful of launches each).

    python tools/ncu_target.py <stitch2|stitch13|stitch2o|pre_prod|pre_parity|attn|ln|gemm_qkv|gemm_proj|gemm_fc2>
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import instageo_b200  # noqa: E402,F401
from instageo_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
MEAN = [494.905781, 815.239594, 924.335066, 2968.881459, 2634.621962, 1739.579917]
STD = [284.925432, 357.84876, 575.566823, 896.601013, 951.900334, 921.407808]


def main(what, iters=4):
    if what.startswith("stitch"):
        nc = 13 if "13" in what else 2
        stride = 112 if what.endswith("o") else 224
        H = W = 3660
        ys = ops.window_origins(H, 224, stride, True)
        xs = ops.window_origins(W, 224, stride, True)
        lg = torch.randn(len(ys) * len(xs), nc, 224, 224, device=dev)
        nd = torch.zeros(H, W, dtype=torch.bool, device=dev)
        fn = lambda: ops.stitch(lg, ys, xs, H, W, nodata_px=nd)  # noqa: E731
    elif what.startswith("pre"):
        raw = torch.randint(0, 10000, (1024, 18, 224, 224), dtype=torch.int16, device=dev)
        spec = ops.PreprocessSpec(MEAN, STD, 3, constant_multiplier=1.0, no_data_value=-9999, device=dev)
        kw = dict(want_f32=True, want_mask_elem=True) if what == "pre_parity" else \
            dict(want_f32=False, want_patches=True, want_mask_px=True)
        fn = lambda: ops.preprocess(raw, spec, **kw)  # noqa: E731
    elif what == "attn":
        qkv = torch.randn(64 * 589, 3 * 768, device=dev).bfloat16()
        fn = lambda: ops.attention(qkv, 64, 589, 12)  # noqa: E731
    elif what.startswith("segm"):   # fused eval step: confusion (+ ROC histograms for "segm_auc")
        from instageo_b200.model.metrics import RunningAUC, RunningConfusionMatrix, segmentation_eval_update
        lg = torch.randn(64, 13, 224, 224, device=dev) * 3
        lab = torch.randint(0, 13, (64, 224, 224), device=dev)
        cm, auc = RunningConfusionMatrix(13, -100, device=dev), RunningAUC(13, device=dev)
        fn = lambda: segmentation_eval_update(lg, lab, cm, auc if what == "segm_auc" else None)  # noqa: E731
    elif what == "chipmask":
        from instageo_b200.data import create_chip
        tile = torch.randint(-100, 10200, (18, 3660, 3660), device=dev, dtype=torch.int16)
        fm = (torch.rand((3, 3660, 3660), device=dev) < 0.2).to(torch.uint8) * 2
        seg = torch.randint(-1, 5, (3660, 3660), device=dev, dtype=torch.int8)
        fn = lambda: create_chip(tile, fm, seg, "each")  # noqa: E731
    elif what == "ln":
        x = torch.randn(64 * 589, 768, device=dev)
        g, b = torch.randn(768, device=dev), torch.randn(768, device=dev)
        fn = lambda: ops.layernorm(x, g, b)  # noqa: E731
    elif what.startswith("gemm"):
        M = 64 * 589
        N, K = {"gemm_qkv": (2304, 768), "gemm_proj": (768, 768), "gemm_fc1": (3072, 768), "gemm_fc2": (768, 3072)}[what]
        a = torch.randn(M, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) * 0.03).bfloat16()
        bias = torch.randn(N, device=dev)
        if what in ("gemm_proj", "gemm_fc2"):
            r = torch.randn(M, N, device=dev)
            fn = lambda: ops.linear(a, w, bias, resid=r)  # noqa: E731
        else:
            fn = lambda: ops.linear(a, w, bias, act=int(what == "gemm_fc1"))  # noqa: E731
    else:
        raise SystemExit(f"unknown target {what}")
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()


if __name__ == "__main__":
    main(sys.argv[1])
