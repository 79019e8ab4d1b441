"""TMA tile-load latency probe (library built by `tools/attn_ablate.sh prof`)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import instageo_b200  # noqa
from instageo_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
rows, D3 = 64 * 589, 3 * 768
qkv = torch.randn(rows, D3, device=dev).bfloat16()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
n = 64
for ctas in (1, 148, 296):
    out = torch.zeros(ctas * 2 * n, dtype=torch.int64, device=dev)
    flush.zero_(); torch.cuda.synchronize()
    lib.ig_debug_tma_latency.argtypes = None
    rc = lib.ig_debug_tma_latency(_lib.C.c_void_p(qkv.data_ptr()), rows, D3, _lib.C.c_void_p(out.data_ptr()), n, 576, 768, ctas)
    assert rc == 0
    o = out.cpu().view(ctas, 2 * n).double()
    cold, warm = o[:, 1:n], o[:, n + 1:]
    print(f"ctas {ctas:4d}: cold median {cold.median():.0f} p90 {cold.quantile(0.9):.0f} max {cold.max():.0f} | "
          f"L2-warm median {warm.median():.0f} p90 {warm.quantile(0.9):.0f} max {warm.max():.0f}  (clk, one 64x64 bf16 tile per load)")
