import sys, os, torch
sys.path.insert(0, os.getcwd())
import instageo_b200
from instageo_b200 import ops
dev = torch.device("cuda:0")
def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for B in (8, 16, 32, 64, 128, 256):
    qkv = torch.randn(B * 589, 3 * 768, device=dev).bfloat16()
    ms = timeit(lambda: ops.attention(qkv, B, 589, 12))
    mb = qkv.numel() * 2 / 1e6
    print(f"B={B:4d} qkv {mb:7.1f} MB  {ms*1e3:8.1f} us  {ms*1e3/B:6.2f} us/chip  qkv-bytes/time {mb/ms/1e3:6.2f} TB/s")
