"""GPU bring-up probe: runs one kernel family against a torch reference and prints diagnostics.

    python tools/gpu_probe.py <gemm|ln|attn|perf|bench|aux|all>

The probes that compare against the oracle (pre, stitch, model) live in tests/probe_parity.py: only tests/,
smoke() and bench.py's CPU legs may import oracle/.

Each family is meant to be run in its own process (a device trap poisons the CUDA context).
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import instageo_b200  # noqa: E402
from instageo_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
RES = {}


def report(name, got, ref, tol):
    got, ref = got.float().cpu(), ref.float().cpu()
    err = (got - ref).abs()
    rel = err.max().item() / max(ref.abs().max().item(), 1e-12)
    bad = (err > tol).float().mean().item()
    RES[name] = dict(max_abs=err.max().item(), rel=rel, frac_bad=bad, ref_absmax=ref.abs().max().item(),
                     nan=bool(torch.isnan(got).any()))
    print(f"[{name}] max_abs={err.max().item():.4e} rel={rel:.3e} frac>{tol}={bad:.4f} nan={RES[name]['nan']}")
    if bad > 0:
        idx = (err > tol).nonzero()[:8].tolist()
        print("   first bad idx:", idx)
        rows = (err > tol).any(dim=-1).nonzero().flatten()
        print("   bad rows (first 16):", rows[:16].tolist(), "count", rows.numel(), "of", got.shape[0])
        cols = (err > tol).any(dim=0).nonzero().flatten() if got.dim() == 2 else []
        if len(cols):
            print("   bad cols (first 16):", cols[:16].tolist(), "count", len(cols))


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def probe_gemm():
    g = torch.Generator(device="cpu").manual_seed(0)
    for (M, N, K) in [(128, 64, 64), (128, 256, 64), (300, 256, 128), (1000, 768, 768), (4096, 2304, 768), (589 * 4, 3072, 768)]:
        a = (torch.randn(M, K, generator=g) * 0.5).to(dev).bfloat16()
        w = (torch.randn(N, K, generator=g) * 0.05).to(dev).bfloat16()
        b = torch.randn(N, generator=g).to(dev)
        ref = a.float() @ w.float().t() + b
        out = ops.linear(a, w, b, out_dtype=torch.float32)
        torch.cuda.synchronize()
        report(f"gemm_f32_{M}x{N}x{K}", out, ref, 2e-3 * max(1.0, ref.abs().max().item()))
        out = ops.linear(a, w, b, act=1)
        report(f"gemm_gelu_{M}x{N}x{K}", out, torch.nn.functional.gelu(ref), 2e-2 * max(1.0, ref.abs().max().item()))
        r = torch.randn(M, N, generator=g).to(dev)
        r0 = r.clone()
        out = ops.linear(a, w, b, resid=r)
        report(f"gemm_resid_{M}x{N}x{K}", out, ref + r0, 2e-3 * max(1.0, ref.abs().max().item()))
    M, N, K = 37696, 2304, 768
    a = torch.randn(M, K, device=dev).bfloat16()
    w = torch.randn(N, K, device=dev).bfloat16()
    b = torch.randn(N, device=dev)
    ms = timeit(lambda: ops.linear(a, w, b))
    print(f"[gemm perf] {M}x{N}x{K}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s")
    RES["gemm_tflops_qkv"] = 2 * M * N * K / ms / 1e9
    ms2 = timeit(lambda: torch.matmul(a, w.t()))
    print(f"[cublas ref] {ms2:.3f} ms {2*M*N*K/ms2/1e9:.1f} TFLOP/s")
    RES["cublas_tflops_qkv"] = 2 * M * N * K / ms2 / 1e9


def probe_ln():
    for D in (256, 768, 1024):
        x = torch.randn(1000, D, device=dev) * 3 + 1
        g_, b_ = torch.randn(D, device=dev), torch.randn(D, device=dev)
        ref = torch.nn.functional.layer_norm(x, (D,), g_, b_, 1e-5)
        report(f"ln_{D}", ops.layernorm(x, g_, b_), ref, 3e-2)


def probe_attn():
    for (B, N, H) in [(1, 128, 1), (2, 197, 4), (2, 589, 12), (3, 64, 2)]:
        D = H * 64
        qkv = (torch.randn(B * N, 3 * D, device=dev)).bfloat16()
        out = ops.attention(qkv, B, N, H)
        torch.cuda.synchronize()
        q, k, v = qkv.float().reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4).unbind(0)
        ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * N, D)
        report(f"attn_B{B}_N{N}_H{H}", out, ref, 2e-2)
    B, N, H = 64, 589, 12
    qkv = torch.randn(B * N, 3 * H * 64, device=dev).bfloat16()
    ms = timeit(lambda: ops.attention(qkv, B, N, H))
    fl = 4 * N * N * 64 * H * B
    print(f"[attn perf] B{B} N{N} H{H}: {ms:.3f} ms {fl/ms/1e9:.1f} TFLOP/s")
    RES["attn_tflops"] = fl / ms / 1e9


def probe_perf():
    """Per-op device time at the bench shapes (V1-100M T=3 B=64 and V2-300M T=3 B=128)."""
    for tag, M, D, H in [("v1_b64", 64 * 589, 768, 12), ("v2_b128", 128 * 589, 1024, 16)]:
        x32 = torch.randn(M, D, device=dev)
        xn = torch.randn(M, D, device=dev).bfloat16()
        hid = torch.randn(M, 4 * D, device=dev).bfloat16()
        wq = (torch.randn(3 * D, D, device=dev) * 0.02).bfloat16()
        wp = (torch.randn(D, D, device=dev) * 0.02).bfloat16()
        w1 = (torch.randn(4 * D, D, device=dev) * 0.02).bfloat16()
        w2 = (torch.randn(D, 4 * D, device=dev) * 0.02).bfloat16()
        bq, bp, b1 = (torch.randn(n, device=dev) for n in (3 * D, D, 4 * D))
        gam, bet = torch.ones(D, device=dev), torch.zeros(D, device=dev)
        qkv = torch.randn(M, 3 * D, device=dev).bfloat16()
        ops_ = [
            ("qkv", lambda: ops.linear(xn, wq, bq), 2 * M * D * 3 * D),
            ("proj+resid", lambda: ops.linear(xn, wp, bp, resid=x32), 2 * M * D * D),
            ("fc1+gelu", lambda: ops.linear(xn, w1, b1, act=1), 2 * M * D * 4 * D),
            ("fc2+resid", lambda: ops.linear(hid, w2, bp, resid=x32), 2 * M * D * 4 * D),
            ("attention", lambda: ops.attention(qkv, M // 589, 589, H), 4 * 589 * 589 * D * (M // 589)),
            ("layernorm", lambda: ops.layernorm(x32, gam, bet), 0),
        ]
        tot = 0.0
        for name, fn, fl in ops_:
            ms = timeit(fn, iters=10)
            tot += ms * (2 if name == "layernorm" else 1)
            print(f"[perf {tag}] {name:11s} {ms*1e3:8.1f} us  {fl/ms/1e9:8.1f} TFLOP/s")
            RES[f"perf_{tag}_{name}"] = dict(us=ms * 1e3, tflops=fl / ms / 1e9)
        print(f"[perf {tag}] block total {tot*1e3:.1f} us")


def probe_bench():
    from instageo_b200.model import PrithviSeg
    for (variant, T, nc, B) in [("prithvi_eo_v1_100", 3, 13, 64)]:
        m = PrithviSeg(temporal_step=T, num_classes=nc, load_pretrained_weights=False, variant=variant).to(dev).eval()
        x = torch.randn(B, 6, T, 224, 224, device=dev)
        ms = timeit(lambda: m.predict(x), iters=5)
        print(f"[bench {variant} T{T} B{B}] {ms:.2f} ms/step {B/ms*1e3:.1f} chips/s")
        RES[f"bench_{variant}_T{T}_B{B}"] = B / ms * 1e3


def probe_aux():
    """§8(f) kernels: eval metrics and tile-scale chip masking -- time and achieved HBM bandwidth."""
    from instageo_b200.data import create_chip
    from instageo_b200.model.metrics import RunningAUC, RunningConfusionMatrix, segmentation_eval_update
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    g = torch.Generator(device="cuda").manual_seed(1042)
    for nc, B, sharp in ((2, 64, 3.0), (13, 64, 3.0), (13, 64, 30.0)):  # sharp = 30: confident model, scores pile up in bins 0 / 1023
        logits = torch.randn(B, nc, 224, 224, generator=g, device=dev) * sharp
        labels = torch.randint(0, nc, (B, 224, 224), generator=g, device=dev)
        labels[torch.rand(B, 224, 224, generator=g, device=dev) < 0.1] = -100
        pred = logits.argmax(1).to(torch.int8)
        cm, auc = RunningConfusionMatrix(nc, -100, device=dev), RunningAUC(nc, device=dev)
        npx = B * 224 * 224
        for name, fn, nbytes in (
                ("confusion_from_int8_map", lambda: cm.update(labels, pred), npx * 9),
                ("eval_step_confusion", lambda: segmentation_eval_update(logits, labels, cm, None), npx * (4 * nc + 8)),
                ("eval_step_confusion_auc", lambda: segmentation_eval_update(logits, labels, cm, auc), npx * (4 * nc + 8))):
            ms = timeit(fn)
            RES[f"aux_{name}_nc{nc}_x{sharp:g}"] = dict(ms=ms, gbs=nbytes / ms / 1e6, frac_hbm=nbytes / ms / 1e6 / peak)
            print(f"[{name} nc={nc} B={B} x{sharp:g}] {ms*1e3:.1f} us  {nbytes/ms/1e6:.0f} GB/s = {nbytes/ms/1e6/peak:.2f} of HBM peak")
        # what the reference does per step with the same tensors (device part only; then 3 D2H copies + numpy)
        def ref_step():
            keep = labels.ne(-100).reshape(-1).nonzero().squeeze()
            p_ = torch.argmax(logits, 1).reshape(-1)[keep]
            q_ = torch.softmax(logits, 1).permute(0, 2, 3, 1).reshape(-1, nc)[keep]
            return p_.cpu(), q_.cpu(), labels.reshape(-1)[keep].cpu()
        ms = timeit(ref_step, iters=5)
        RES[f"aux_reference_device_part_nc{nc}"] = dict(ms=ms)
        print(f"[reference _shared_step gather + D2H, nc={nc}] {ms:.2f} ms (before its numpy bincount / add.at)")
    for T in (1, 3):
        tile = torch.randint(-100, 10200, (6 * T, 3660, 3660), generator=g, device=dev, dtype=torch.int16)
        fm = (torch.rand((T, 3660, 3660), generator=g, device=dev) < 0.2).to(torch.uint8) * 2
        seg = torch.randint(-1, 5, (3660, 3660), generator=g, device=dev, dtype=torch.int8)
        nbytes = 3660 * 3660 * (6 * T * 4 + T + 2)
        ms = timeit(lambda: create_chip(tile, fm, seg, "each"))
        RES[f"aux_chip_mask_T{T}"] = dict(ms=ms, gbs=nbytes / ms / 1e6, frac_hbm=nbytes / ms / 1e6 / peak)
        print(f"[chip_mask 3660^2 T={T}] {ms*1e3:.1f} us  {nbytes/ms/1e6:.0f} GB/s = {nbytes/ms/1e6/peak:.2f} of HBM peak")


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    fams = dict(gemm=probe_gemm, ln=probe_ln, attn=probe_attn, perf=probe_perf, bench=probe_bench, aux=probe_aux)
    t0 = time.time()
    try:
        for k, f in fams.items():
            if which in (k, "all"):
                f()
    finally:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"probe_{which}.json"), "w") as fh:
            json.dump(RES, fh, indent=1, default=str)
        print(f"probe {which} done in {time.time()-t0:.1f}s")
