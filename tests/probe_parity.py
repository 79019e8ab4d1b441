"""GPU bring-up probes that compare a kernel family against the oracle and print diagnostics (not collected by pytest).

    python tests/probe_parity.py <pre|stitch|model|all>
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from gpu_probe import *  # noqa: E402,F401,F403  (report, timeit, dev, RES, ops, torch, np, json, time)
from gpu_probe import RES, dev, ops, report, timeit  # noqa: E402,F401


def probe_pre():
    from oracle import preprocess as OP
    mean = [0.14245495, 0.13921481, 0.12434631, 0.31420089, 0.20743526, 0.12046503]
    std = [0.04036231, 0.04186983, 0.05267646, 0.0822221, 0.06834774, 0.05294205]
    for T, cm, nd in [(1, 1.0, -9999), (3, 1e-4, -9999), (3, 1.0, 0)]:
        raw = OP.synth_chips(3, T, seed=7, nodata=nd)
        spec = ops.PreprocessSpec(mean, std, T, constant_multiplier=cm, no_data_value=nd, device=dev)
        out = ops.preprocess(torch.from_numpy(raw).to(dev), spec, want_f32=True, want_patches=True,
                             want_mask_elem=True, want_mask_px=True)
        ref = [OP.preprocess_chip(r, None, cm, mean, std, T, nd) for r in raw]
        rx = np.stack([r[0] for r in ref]); rm = np.stack([r[1] for r in ref])
        eq = np.array_equal(out["f32"].cpu().numpy(), rx)
        meq = np.array_equal(out["mask_elem"].cpu().numpy(), rm)
        peq = np.array_equal(out["mask_px"].cpu().numpy(), rm.any(axis=1))
        print(f"[pre T={T} cm={cm} nd={nd}] f32 bit-exact={eq} mask_elem={meq} mask_px={peq} masked={rm.mean():.4f}")
        RES[f"pre_T{T}_cm{cm}"] = dict(f32=eq, mask_elem=meq, mask_px=peq)
    n, T = 2048, 3
    raw = torch.randint(0, 10000, (n, 18, 224, 224), dtype=torch.int16, device=dev)
    spec = ops.PreprocessSpec(mean, std, T, constant_multiplier=1.0, no_data_value=-9999, device=dev)
    for tag, kw, bpc in [("parity", dict(want_f32=True, want_mask_elem=True), 18 * 224 * 224 * 7),
                         ("prod", dict(want_f32=False, want_patches=True, want_mask_px=True), 18 * 224 * 224 * 4 + 224 * 224)]:
        ms = timeit(lambda: ops.preprocess(raw, spec, **kw), iters=5)
        print(f"[pre perf {tag}] {n} chips T=3: {ms:.3f} ms  {n*bpc/ms/1e6:.1f} GB/s  {n/ms*1e3:.0f} chips/s")
        RES[f"pre_gbs_{tag}"] = n * bpc / ms / 1e6


def probe_stitch():
    from oracle import stitch as OS
    rng = np.random.default_rng(3)
    for (H, W, win, stride, nc) in [(300, 340, 64, 32, 2), (500, 500, 224, 112, 13), (256, 256, 64, 64, 3)]:
        ys = ops.window_origins(H, win, stride, True); xs = ops.window_origins(W, win, stride, True)
        org = [(t, l) for t in ys for l in xs]
        lg = rng.standard_normal((len(org), nc, win, win)).astype(np.float32)
        nd = rng.random((H, W)) < 0.05
        avg, cls = OS.stitch(lg, org, H, W, nd)
        out = ops.stitch(torch.from_numpy(lg).to(dev), ys, xs, H, W, nodata_px=torch.from_numpy(nd).to(dev),
                         want_avg=True, want_hist=True)
        a_eq = np.array_equal(out["avg"].cpu().numpy(), avg)
        c_eq = np.array_equal(out["class_map"].cpu().numpy(), cls)
        hist = out["hist"].cpu().numpy()
        h_ref = [int((cls == k).sum()) for k in range(nc)] + [int((cls == -1).sum())]
        print(f"[stitch {H}x{W} win{win} s{stride} nc{nc}] avg bit-exact={a_eq} cls={c_eq} hist={list(hist)==h_ref}")
        RES[f"stitch_{H}_{stride}_{nc}"] = dict(avg=a_eq, cls=c_eq, hist=list(map(int, hist)) == h_ref)
    from instageo_b200 import _lib
    H = W = 3660; win = 224
    for nc in (2, 13):
        for stride in (224, 112):
            ys = ops.window_origins(H, win, stride, True); xs = ops.window_origins(W, win, stride, True)
            lg = torch.randn(len(ys) * len(xs), nc, win, win, device=dev)
            nd = torch.zeros(H, W, dtype=torch.bool, device=dev)
            for _ in range(3):
                ops.stitch(lg, ys, xs, H, W, nodata_px=nd)
            torch.cuda.synchronize()
            _lib.profile_enable(True); _lib.profile_report()
            for _ in range(10):
                ops.stitch(lg, ys, xs, H, W, nodata_px=nd)
            torch.cuda.synchronize()
            ms, n = _lib.profile_report()["stitch"]
            _lib.profile_enable(False)
            ms /= n
            by = lg.numel() * 4 + 2 * H * W
            print(f"[stitch perf nc{nc} stride {stride}] kernel {ms*1e3:.1f} us  {by/ms/1e6:.1f} GB/s ({by/1e6:.0f} MB)")
            RES[f"stitch_gbs_nc{nc}_{stride}"] = by / ms / 1e6
            del lg

def probe_model():
    from instageo_b200.model import PrithviSeg
    from oracle import prithvi as P
    for (variant, T, nc, depth, B) in [("prithvi_eo_tiny", 1, 2, 0, 2), ("prithvi_eo_tiny", 1, 2, 1, 2),
                                       ("prithvi_eo_tiny", 3, 13, 2, 2), ("prithvi_eo_v1_100", 1, 2, 2, 2)]:
        tag = f"{variant}_T{T}_nc{nc}_d{depth}"
        sd = P.make_state_dict(variant, T, nc, depth=depth, seed=5, stress=True)
        m = PrithviSeg(temporal_step=T, num_classes=nc, load_pretrained_weights=False, variant=variant, depth=depth)
        m.load_state_dict(sd, strict=True)
        m.to(dev).eval()
        x = torch.randn(B, 6, T, 224, 224, generator=torch.Generator().manual_seed(1))
        taps = {}
        ref = P.prithvi_seg_forward(x, sd, P.VARIANTS[variant][2], T, taps=taps)
        try:
            out, feat = m(x.to(dev), return_features=True)
            torch.cuda.synchronize()
        except Exception as e:
            print(f"[{tag}] FAILED: {e}")
            RES[tag] = str(e)
            raise
        D = P.VARIANTS[variant][0]
        N = 1 + T * 196
        xres = m.debug_tap("x", B, (B, N, D))
        report(f"{tag}_x", xres, taps[f"block{depth-1}"] if depth > 0 else taps["embed"], 5e-2)
        report(f"{tag}_feat", feat, P.tokens_to_image(taps["tokens"], T), 5e-2)
        dims = P.head_dims(D, T)
        hw = 14
        for i in range(4):
            hw *= 2
            report(f"{tag}_convt{i}", m.debug_tap(f"convt{i}", B, (B, dims[i + 1], hw, hw)), taps[f"convt{i}"], 5e-2)
            if i < 3:
                report(f"{tag}_stage{i}", m.debug_tap(f"stage{i}", B, (B, dims[i + 1], hw, hw)), taps[f"stage{i}"], 5e-2)
        report(f"{tag}_logits", out, ref, 2e-2)
        am = m.predict(x.to(dev)).cpu()
        top2 = ref.topk(2, dim=1).values
        eps = (out.cpu() - ref).abs().max().item()
        safe = (top2[:, 0] - top2[:, 1]) > 2 * eps
        agree = (am.long() == ref.argmax(1))[safe].float().mean().item()
        print(f"[{tag}] argmax agree outside ties {agree:.5f} (safe frac {safe.float().mean().item():.3f})")
        RES[f"{tag}_argmax"] = dict(agree=agree, safe=safe.float().mean().item())
        del m



if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    for k, f in dict(pre=probe_pre, stitch=probe_stitch, model=probe_model).items():
        if which in (k, "all"):
            f()
