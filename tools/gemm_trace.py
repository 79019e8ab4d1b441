"""Hand-off timeline of one GEMM kernel instantiation (debug library built with -DIG_GEMM_TRACE=<epi>):
  tools/variants.sh build gemm_tc.cu trace6:"-DIG_GEMM_TRACE=6"
  INSTAGEO_B200_LIB=$PWD/instageo-e2e-geospatial-ml_b200/libig_trace6.so python tools/gemm_trace.py [T nc B [first_tile ntiles]]
Prints, for the leader CTA of pair 0, the SM-clock time of every role event of a few consecutive tiles:
P = producer found the stage free and issued its loads, M0 = MMA warp got the accumulator slot, M1 = stage data landed,
M2 = tile's UMMAs issued + committed, E0 = epilogue warp 0 ready, E1 = accumulator complete, E2 = slot handed back,
E3 = tile stored."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import instageo_b200  # noqa: E402,F401
from instageo_b200 import _lib  # noqa: E402
from instageo_b200.model import PrithviSeg  # noqa: E402

lib = ctypes.CDLL(os.environ["INSTAGEO_B200_LIB"])
lib.ig_debug_gemm_trace.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
dev = torch.device("cuda:0")
torch.manual_seed(0)
if len(sys.argv) > 1 and sys.argv[1] == "linear":
    # one plain GEMM instead of a forward: linear M N K resid(0|1) [first_tile ntiles]  (EPI 2 with resid, else EPI 0)
    from instageo_b200 import ops  # noqa: E402
    M, N, K, with_resid = (int(v) for v in sys.argv[2:6])
    first, ntiles = (int(v) for v in sys.argv[6:8]) if len(sys.argv) > 7 else (2, 2)
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    x = torch.randn(M, N, device=dev)

    def run():
        return ops.linear(a, w, None, resid=x) if with_resid else ops.linear(a, w, None)
else:
    T, nc, B = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (1, 2, 145)
    first, ntiles = (int(v) for v in sys.argv[4:6]) if len(sys.argv) > 5 else (20, 4)
    m = PrithviSeg(temporal_step=T, num_classes=nc, load_pretrained_weights=False).to(dev).eval()
    patches = torch.randn(B * T * 196, 1536, device=dev).bfloat16()

    def run():
        return m.forward_patches(patches, want_logits=True, want_argmax=True)
for _ in range(3):
    run()
torch.cuda.synchronize()
lib.ig_debug_gemm_trace(None, 0, 1)
run()
torch.cuda.synchronize()
buf = np.zeros(3 << 13, dtype=np.uint64)
lib.ig_debug_gemm_trace(buf.ctypes.data, buf.size, 1)
ev = [(int(v) >> 24, (int(v) >> 8) & 0xffff, int(v) & 0xff) for v in buf if v]
n = len(ev)
names = {0x10: "P ", 0x20: "M0", 0x21: "M1", 0x22: "M2", 0x23: "M3", 0x30: "E0", 0x31: "E1", 0x32: "E2", 0x33: "E3"}
tiles = sorted({t for _, t, _ in ev})
print(f"{n} events, {len(tiles)} tiles of pair 0 (tile ids {tiles[:4]} ...)")
sel = tiles[first:first + ntiles]
t0 = min(c for c, t, _ in ev if t == sel[0])
for c, t, tag in sorted(e for e in ev if e[1] in sel):
    print(f"  +{c - t0:7d} clk  tile {t:6d}  {names.get(tag, hex(tag))}")
# per-role period
for tag in (0x22, 0x31, 0x33):
    ts = sorted(c for c, _, g in ev if g == tag)
    d = np.diff(ts[5:-5])
    if len(d):
        print(f"{names[tag]} period: median {np.median(d):.0f} clk, mean {d.mean():.0f}")
# waits
def gaps(a, b):
    ta = {t: c for c, t, g in ev if g == a}
    tb = {t: c for c, t, g in ev if g == b}
    g = [tb[t] - ta[t] for t in tiles[5:-5] if t in ta and t in tb]
    return np.median(g) if g else float("nan")
print(f"E0->E1 (epilogue waits for the accumulator) median {gaps(0x30, 0x31):.0f} clk; E1->E2 (TMEM read + math) {gaps(0x31, 0x32):.0f}; "
      f"E2->E3 (exchange + stores) {gaps(0x32, 0x33):.0f}; M0->M2 (MMA warp per tile) {gaps(0x20, 0x22):.0f}")
