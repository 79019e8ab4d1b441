import os, sys, torch
sys.path.insert(0, '/root/repo')
import instageo_b200
from instageo_b200 import ops
import torch.nn.functional as F
dev = torch.device('cuda:0')
for (B, N, H) in ((64, 589, 12), (256, 197, 12), (128, 589, 16)):
    qkv = torch.randn(B * N, 3 * H * 64, device=dev).bfloat16()
    for _ in range(5): out = ops.attention(qkv, B, N, H)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30): out = ops.attention(qkv, B, N, H)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 30 * 1e3
    q, k, v = qkv[: 2 * N].float().reshape(2, N, 3, H, 64).permute(2, 0, 3, 1, 4).unbind(0)
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(2 * N, H * 64)
    err = (out[: 2 * N].float() - ref).abs().max().item()
    print(f"PAIR={os.environ.get('IG_ATTN_PAIR','0')} B={B} N={N} H={H}: {us:8.1f} us   {4*B*H*N*N*64/us/1e6:7.1f} TFLOP/s   max-abs err {err:.2e}")
