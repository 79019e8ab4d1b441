"""Per-launch CUDA-event times of one PrithviSeg forward (kernel by kernel, eager): python tools/forward_launches.py
[variant T nc B].  Head launches are labelled convT<i> / conv<i> / final."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import instageo_b200  # noqa: E402,F401
from instageo_b200 import _lib  # noqa: E402
from instageo_b200.model import PrithviSeg  # noqa: E402
from instageo_b200.model.model import flops_per_chip  # noqa: E402

variant = sys.argv[1] if len(sys.argv) > 1 else "prithvi_eo_v1_100"
T, nc, B = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (1, 2, 145)
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = PrithviSeg(temporal_step=T, num_classes=nc, load_pretrained_weights=False, variant=variant).to(dev).eval()
rows = T * 196
patches = torch.randn(B * rows, 1536, device=dev).bfloat16()
for _ in range(3):
    m.forward_patches(patches, want_logits=True, want_argmax=True)
torch.cuda.synchronize()
_lib.profile_enable(True)
_lib.profile_launches()
reps = 3
for _ in range(reps):
    m.forward_patches(patches, want_logits=True, want_argmax=True)
torch.cuda.synchronize()
rec = _lib.profile_launches()
_lib.profile_enable(False)
n = len(rec) // reps
L = len(m.prithvi_encoder.blocks)
names = ["cls", "patch_embed"]
for i in range(L):
    names += [f"b{i}.ln1", f"b{i}.qkv", f"b{i}.attn", f"b{i}.proj", f"b{i}.ln2", f"b{i}.fc1", f"b{i}.fc2"]
names += ["norm"]
for i in range(4):
    names += [f"convT{i}", f"conv{i}" if i < 3 else "final"]
D = m.prithvi_encoder.embed_dim
dims = m.embed_dims
hw = 14
hflops = {}
for i in range(4):
    hflops[f"convT{i}"] = 2 * hw * hw * 9 * dims[i] * dims[i + 1] * B
    hw *= 2
    hflops[f"conv{i}" if i < 3 else "final"] = 2 * hw * hw * 9 * dims[i + 1] ** 2 * B
tot = 0.0
agg = {}
for j in range(n):
    ms = sum(rec[r * n + j][1] for r in range(reps)) / reps
    nm = names[j] if j < len(names) else f"#{j}"
    tot += ms
    key = nm.split(".")[-1] if nm.startswith("b") else nm
    agg[key] = agg.get(key, 0.0) + ms
for k, v in agg.items():
    extra = f"  {hflops[k] / v / 1e9:8.1f} TFLOP/s" if k in hflops else ""
    print(f"{k:12s} {v * 1e3:9.1f} us{extra}")
print(f"total {tot:.3f} ms for B={B} ({B / tot * 1e3:.0f} chips/s kernel-only), {n} launches")
