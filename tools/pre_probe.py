"""Kernel 1 throughput (CUDA events): chip batches as the bench step launches them (64 chips, T = 3, production mode), a
large batch, and WINDOW mode over the 3660 x 3660 x 6 tile (unaligned 224-px windows, stride 112 / 224; each raw
pixel is read by up to four windows, so algorithmic bytes count every window's read).  Also the nodata-map kernel."""
import json
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
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6500.0


def timeit(fn, iters=20, flush=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        if flush is not None:
            flush.zero_()           # > L2: the next launch reads its input from DRAM
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = {}
for n in (64, 1024):
    raw = torch.randint(0, 10001, (n, 18, 224, 224), dtype=torch.int16, device=dev)
    spec = ops.PreprocessSpec(MEAN, STD, 3, None, 1.0, None, dev)
    out = torch.empty((n * 3 * 196, 1536), dtype=torch.bfloat16, device=dev)
    ms = timeit(lambda: ops.preprocess(raw, spec, want_f32=False, want_patches=True, out_patches=out), flush=flush)
    b = n * 18 * 224 * 224 * 4
    res[f"chips_{n}_T3_production"] = dict(us=ms * 1e3, gbs=b / ms / 1e6, frac=b / ms / 1e6 / peak)
tile = torch.randint(0, 10001, (6, 3660, 3660), dtype=torch.int16, device=dev)
tile[:, :300, :300] = -9999
spec1 = ops.PreprocessSpec([m for m in MEAN], STD, 1, None, 1.0, -9999, dev)
for stride in (224, 112):
    ys = ops.window_origins(3660, 224, stride, True)
    wins = torch.tensor([(0, t, l) for t in ys for l in ys], dtype=torch.int32, device=dev)
    for nwin in (256, len(wins)):
        w = wins[:nwin].contiguous()
        out = torch.empty((nwin * 196, 1536), dtype=torch.bfloat16, device=dev)
        ms = timeit(lambda: ops.preprocess(tile.unsqueeze(0), spec1, windows=w, win=224, want_f32=False, want_patches=True,
                                           out_patches=out), flush=flush)
        b = nwin * 6 * 224 * 224 * 4
        res[f"tile_windows_stride{stride}_n{nwin}"] = dict(us=ms * 1e3, gbs=b / ms / 1e6, frac=b / ms / 1e6 / peak)
nd = torch.empty((3660, 3660), dtype=torch.uint8, device=dev)
ms = timeit(lambda: ops.nodata_map(tile, spec1, out=nd), flush=flush)
b = 3660 * 3660 * 13
res["nodata_map_3660"] = dict(us=ms * 1e3, gbs=b / ms / 1e6, frac=b / ms / 1e6 / peak)
for k, v in res.items():
    print(f"{k:36s} {v['us']:8.1f} us  {v['gbs']:7.0f} GB/s = {v['frac']:.2f} of the measured HBM peak ({peak:.0f} GB/s)")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "pre_probe.json"), "w"), indent=1)
