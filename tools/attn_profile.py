"""In-kernel phase breakdown of the attention kernel (library built with -DATTN_PROFILE by tools/attn_ablate.sh prof)."""
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
buf = (ctypes.c_ulonglong * 32)()
for _ in range(3):
    ops.attention(qkv, B, N, H)
lib.ig_attention_profile(buf)
ops.attention(qkv, B, N, H)
lib.ig_attention_profile(buf)
v = list(buf)
ctas = v[15]
nb = (N + 63) // 64
names = ["softmax: wait S_j", "softmax: TMEM load S_j", "softmax: mask+max(+rescale)", "softmax: wait P free", "softmax: exps + P stores",
         "softmax: fences + arrive", "", "", "warp0: issue QK_g+2", "warp1: wait P_g", "warp1: issue PV_g"]
print(f"CTAs {ctas}, KV blocks per CTA {nb}, mean CTA lifetime {v[14]/ctas:.0f} clk")
items = -(-(64 * 12 * 5) // ctas)
print(f"  per item (~{items} per CTA): wait last PV {v[7]/ctas/items:.0f}, O epilogue {v[13]/ctas/items:.0f}; per block: "
      f"QK wait Q/K {v[11]/ctas/items/nb:.0f}, QK wait S free {v[6]/ctas/items/nb:.0f}, PV wait V {v[12]/ctas/items/nb:.0f}")
print(f"  warp0 per block: wait V buffer free {v[16]/ctas/items/nb:.0f}, issue V load {v[17]/ctas/items/nb:.0f}, V load issue->landed (mode 4 only) {v[18]/ctas/items/nb:.0f}")
print(f"  V loads: {v[20]} of {v[21]} not landed when PV_g wanted them; for those, issue -> landed = {v[19]/max(1,v[20]):.0f} clk")
print("  late V loads by block index within the item:", [int(x) for x in v[22:32]])
nb *= items
for i, n in enumerate(names):
    if n:
        print(f"  {n:45s} {v[i]/ctas/nb:8.1f} clk per KV block")
