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
buf = (ctypes.c_ulonglong * 16)()
for _ in range(3):
    ops.attention(qkv, B, N, H)
lib.ig_attention_profile(buf)
ops.attention(qkv, B, N, H)
lib.ig_attention_profile(buf)
v = list(buf)
ctas = v[15]
nb = (N + 63) // 64
names = ["softmax: wait S_j", "softmax: TMEM load S_j", "softmax: mask+max(+rescale)", "softmax: wait P free", "softmax: exps + P stores",
         "softmax: fences + arrive", "", "", "mma: issue QK_j+2 (incl. waits K, S free)", "mma: wait V_j, P_j", "mma: issue PV_j"]
print(f"CTAs {ctas}, KV blocks per CTA {nb}, mean CTA lifetime {v[14]/ctas:.0f} clk")
print(f"  per CTA: kernel start -> softmax loop {v[12]/ctas:.0f} clk, wait S_0 {v[6]/ctas:.0f}, wait last PV {v[7]/ctas:.0f}, "
      f"O epilogue {v[13]/ctas:.0f}, softmax loop total {sum(v[0:6])/ctas:.0f}")
for i, n in enumerate(names):
    if n:
        print(f"  {n:45s} {v[i]/ctas/nb:8.1f} clk per KV block")
