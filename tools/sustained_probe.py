"""Sustained (power-capped) throughput of the encoder GEMM shapes: this library's tcgen05 kernel against cuBLAS (torch.matmul)
on the SAME shapes, each in a ~1.5 s back-to-back loop, interleaved so that both see the same thermal / power state, with the
SM clock sampled during each loop.  Puts the bench line's roofline fraction (denominator: a large square cuBLAS GEMM under the
cap) next to what cuBLAS itself sustains on these K = 768 / 3072 shapes."""
import json
import os
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import instageo_b200  # noqa: E402,F401
from instageo_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
M = 64 * 589


def clock():
    try:
        o = subprocess.run(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits"],
                           capture_output=True, text=True, timeout=5).stdout.strip().split(",")
        return int(o[0]), float(o[1])
    except Exception:
        return None, None


def loop(fn, seconds=1.5):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    n, t0 = 0, time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    clk = None
    while time.perf_counter() - t0 < seconds:
        for _ in range(50):
            fn()
        n += 50
        if clk is None and time.perf_counter() - t0 > seconds / 2:
            clk = clock()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, clk


res = {}
for name, N, K in (("qkv", 2304, 768), ("fc1", 3072, 768), ("fc2", 768, 3072), ("proj", 768, 768), ("square8k", 8192, 8192)):
    m = 8192 if name == "square8k" else M
    a = torch.randn(m, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.03).bfloat16()
    bias = torch.randn(N, device=dev)
    out = torch.empty(m, N, device=dev, dtype=torch.bfloat16)
    fl = 2.0 * m * N * K
    for rep in range(2):
        ms_c, clk_c = loop(lambda: torch.matmul(a, w.t(), out=out))
        ms_o, clk_o = loop(lambda: ops.linear(a, w, bias))
        res[f"{name}_{rep}"] = dict(cublas_tflops=fl / ms_c / 1e9, ours_tflops=fl / ms_o / 1e9, cublas_clk_w=clk_c, ours_clk_w=clk_o)
        print(f"{name:9s} rep {rep}: cuBLAS {fl / ms_c / 1e9:7.1f} TFLOP/s (SM MHz, W: {clk_c})   ours {fl / ms_o / 1e9:7.1f} TFLOP/s ({clk_o})"
              f"   ours / cuBLAS = {ms_c / ms_o:.3f}")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "sustained_probe.json"), "w"), indent=1)
