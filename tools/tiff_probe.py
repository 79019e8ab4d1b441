"""Device-side TIFF kernels (csrc/tiffio.cu) timed alone with CUDA events: the predictor-2 undo + byte swap + chunky -> planar
unpack of a 3660 x 3660 x 6 16-bit raster (strips of 256 rows and 512 x 512 tiles), and the forward differencing of
batched predictions.  Algorithmic bytes: 2 read + 2 written per 16-bit sample; 1 + 1 per int8 pixel."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import instageo_b200  # noqa: E402,F401
from instageo_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3   # us


H = W = 3660
spp = 6
for name, bw, bh, planar in (("chunky strips of 256 rows", W, 256, 0), ("chunky 512 x 512 tiles", 512, 512, 0),
                             ("planar strips of 256 rows", W, 256, 1)):
    nbx, nby = -(-W // bw), -(-H // bh)
    n_samples = nbx * nby * bw * bh * spp        # padded blocks as stored
    blocks = torch.randint(-32768, 32767, (n_samples,), dtype=torch.int16, device=dev)
    out = torch.empty((spp, H, W), dtype=torch.int16, device=dev)
    for pred, swap in ((2, 0), (2, 1), (1, 0)):
        us = timed(lambda: _lib.call("ig_tiff_unpack16", dev, blocks.data_ptr(), out.data_ptr(), W, H, spp, bw, bh, planar, pred, swap))
        gb = (n_samples * 2 + spp * H * W * 2) / 1e9
        print(f"ig_tiff_unpack16 {H}x{W}x{spp} {name:26s} predictor {pred} byteswap {swap}: {us:8.1f} us  {gb / us * 1e6:7.1f} GB/s")
for shape, dt in (((64, 224, 224), torch.int8), ((1, 3660, 3660), torch.int8), ((64, 224, 224), torch.int16)):
    n, h, w = shape
    src = torch.randint(-100, 100, shape, dtype=dt, device=dev)
    dst = torch.empty_like(src)
    us = timed(lambda: _lib.call("ig_tiff_predict", dev, src.data_ptr(), dst.data_ptr(), src.element_size(), n * h, w))
    gb = 2 * src.numel() * src.element_size() / 1e9
    print(f"ig_tiff_predict  {shape} {str(dt):12s}: {us:8.1f} us  {gb / us * 1e6:7.1f} GB/s")
