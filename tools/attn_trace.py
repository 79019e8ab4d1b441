"""Event timeline of attention CTA 0 (library built with PROFMODE=3 tools/attn_ablate.sh prof)."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import instageo_b200  # noqa
from instageo_b200 import _lib, ops
lib = _lib.load()
dev = torch.device("cuda:0")
B, N, H = 64, 589, 12
qkv = torch.randn(B * N, 3 * H * 64, device=dev).bfloat16()
buf = (ctypes.c_longlong * (3 * 4096 * 3))()
cnt = (ctypes.c_int * 3)()
for _ in range(3):
    ops.attention(qkv, B, N, H)
lib.ig_attention_trace(buf, cnt)
ops.attention(qkv, B, N, H)
lib.ig_attention_trace(buf, cnt)
names = {1: "prod K issue", 2: "prod V issue", 10: "mma QK waits ok", 11: "mma QK issued", 12: "mma V,P ready", 13: "mma PV issued", 14: "mma V ready",
         20: "soft S ready", 21: "soft exps start", 22: "soft P published", 23: "soft last PV done", 24: "soft item stored"}
ev = []
for r in range(3):
    for i in range(min(cnt[r], 4096)):
        o = (r * 4096 + i) * 3
        ev.append((buf[o + 2], r, buf[o], buf[o + 1]))
ev.sort()
t0 = ev[0][0]
lo, hi = int(sys.argv[1]) if len(sys.argv) > 1 else 14, int(sys.argv[2]) if len(sys.argv) > 2 else 36
for t, r, e, b in ev:
    if lo <= b <= hi:
        wq = e // 100
        print(f"{t - t0:9d}  {'  ' * r * 8}{'    ' * max(0, wq - 2)}{names[e % 100]} {b}" + (f" w{wq}" if wq else ""))
