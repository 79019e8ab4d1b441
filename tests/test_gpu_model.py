"""GPU parity: PrithviSeg forward (kernels 2-4) through the drop-in class vs the fp32 oracle and the
golden vectors frozen from the reference.  Bars (BASELINE.json north_star): logits <= 2e-2 max-abs
(bf16 engine vs fp32 reference); argmax bit-identical outside ties (margin <= 2*eps)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import prithvi as P
from oracle.gen_golden import MODEL_CASES, model_input

pytestmark = pytest.mark.gpu
TOL = 2e-2


def _build(variant, T, nc, depth, dev, seed=5, stress=True):
    from instageo_b200.model import PrithviSeg
    sd = P.make_state_dict(variant, T, nc, depth=depth, seed=seed, stress=stress)
    m = PrithviSeg(temporal_step=T, num_classes=nc, load_pretrained_weights=False, variant=variant, depth=depth)
    m.load_state_dict(sd, strict=True)
    return m.to(dev).eval(), sd


def _argmax_check(am, ref, eps):
    top2 = ref.topk(2, dim=1).values
    safe = (top2[:, 0] - top2[:, 1]) > 2 * eps
    assert safe.float().mean().item() > 0.5, "margin rule excluded most pixels: test is vacuous"
    assert bool((am.cpu().long() == ref.argmax(1))[safe].all())
    return 1.0 - safe.float().mean().item()


@pytest.mark.parametrize("name", sorted(MODEL_CASES))
def test_golden_from_reference(cuda_dev, name):
    variant, T, nc, depth, wseed, stress, iseed, batch = MODEL_CASES[name]
    g = np.load(os.path.join(GOLDEN, f"model_{name}.npz"))
    m, _ = _build(variant, T, nc, depth, cuda_dev, seed=wseed, stress=stress)
    x = model_input(iseed, batch, T).to(cuda_dev)
    y, feat = m(x, return_features=True)
    assert y.shape == (batch, nc, 224, 224) and y.dtype == torch.float32
    eps = float(np.abs(y[:, :, ::4, ::4].cpu().numpy() - g["logits_sub"]).max())
    assert eps < TOL
    assert float(np.abs(feat[:, ::16].cpu().numpy() - g["feat_sub"]).max()) < 5e-2
    am = m.predict(x).cpu().numpy()
    assert am.dtype == np.int8 and am.shape == (batch, 224, 224)
    assert (am == g["argmax"]).mean() > 0.97


@pytest.mark.parametrize("variant,T,nc,depth,B,stress", [
    ("prithvi_eo_tiny", 1, 2, 0, 2, True), ("prithvi_eo_tiny", 3, 13, 2, 3, True),
    ("prithvi_eo_v1_100", 1, 2, 2, 2, True), ("prithvi_eo_v1_100", 3, 13, 1, 1, False),
    ("prithvi_eo_v2_300", 3, 13, 1, 1, True), ("prithvi_eo_v1_100", 1, 1, 1, 2, True)])
def test_forward_vs_oracle(cuda_dev, variant, T, nc, depth, B, stress):
    m, sd = _build(variant, T, nc, depth, cuda_dev, stress=stress)
    x = torch.randn(B, 6, T, 224, 224, generator=torch.Generator().manual_seed(B))
    taps = {}
    ref = P.prithvi_seg_forward(x, sd, P.VARIANTS[variant][2], T, taps=taps)
    y = m(x.to(cuda_dev))
    eps = (y.cpu() - ref).abs().max().item()
    assert eps < TOL, f"logits max-abs {eps}"
    # stage-by-stage taps (bf16 activations, f32 accumulation): relative error of every head map
    D = P.VARIANTS[variant][0]
    dims, hw = P.head_dims(D, T), 14
    for i in range(4):
        hw *= 2
        for nm in ([f"convt{i}", f"stage{i}"] if i < 3 else [f"convt{i}"]):
            got = m.debug_tap(nm, B, (B, dims[i + 1], hw, hw)).cpu()
            rel = (got - taps[nm]).abs().max().item() / max(taps[nm].abs().max().item(), 1e-6)
            assert rel < 2e-2, f"{nm}: rel err {rel}"
    if nc > 1:
        _argmax_check(m.predict(x.to(cuda_dev)), ref, eps)
        # fused argmax == argmax of the engine's own logits, bit for bit (first max wins)
        assert torch.equal(m.predict(x.to(cuda_dev)).long(), torch.argmax(y, dim=1))
    else:
        assert y.shape == (B, 1, 224, 224)


def test_interface_behaviour(cuda_dev):
    m, sd = _build("prithvi_eo_tiny", 1, 2, 1, cuda_dev)
    x = torch.randn(2, 6, 224, 224, device=cuda_dev)  # 4-D input accepted when T == 1 (pritvhi.py:507-509)
    y4 = m(x)
    assert torch.equal(y4, m(x.unsqueeze(2)))
    y1 = m(x[:1])
    assert torch.equal(y1, y4[:1])  # batch independence / determinism
    with pytest.raises(ValueError):
        m(torch.randn(1, 6, 1, 112, 112, device=cuda_dev))
    # weights edited in place are re-packed (version counter), like a fine-tuned checkpoint load
    with torch.no_grad():
        m.segmentation_head[5].bias.add_(torch.tensor([3.0, -3.0], device=cuda_dev))
    assert bool((m.predict(x) == 0).all())
    sd2 = P.make_state_dict("prithvi_eo_tiny", 1, 2, depth=1, seed=99, stress=True)
    m.load_state_dict(sd2)
    ref = P.prithvi_seg_forward(x.cpu(), sd2, 4, 1)
    assert (m(x).cpu() - ref).abs().max().item() < TOL
    assert m.launches_per_forward() == 3 + 7 * 1 + 1 + 8


def test_fused_preprocess_to_model(cuda_dev):
    """raw int16 -> kernel 1 tubelet rows -> model == raw -> f32 tensor -> model (same bf16 operands)."""
    from instageo_b200 import ops
    from oracle import preprocess as OP
    from conftest import FLOOD_MEAN, FLOOD_STD
    m, sd = _build("prithvi_eo_tiny", 3, 13, 1, cuda_dev)
    raw = torch.from_numpy(OP.synth_chips(2, 3, seed=3)).to(cuda_dev)
    spec = ops.PreprocessSpec(FLOOD_MEAN, FLOOD_STD, 3, None, 1e-4, None, cuda_dev)
    pre = ops.preprocess(raw, spec, want_f32=True, want_patches=True)
    la, aa = m.forward_patches(pre["patches"], want_logits=True, want_argmax=True)
    lb = m(pre["f32"])
    assert torch.equal(la, lb) and torch.equal(aa.long(), lb.argmax(1))


def test_chip_inference_loop(cuda_dev, tmp_path):
    from instageo_b200.model import chip_inference
    from instageo_b200.model.dataloader import InstaGeoChipDataset, make_preprocess_func
    from oracle import preprocess as OP
    from conftest import FLOOD_MEAN, FLOOD_STD
    m, sd = _build("prithvi_eo_tiny", 1, 2, 1, cuda_dev)
    raw = OP.synth_chips(5, 1, seed=8)
    ds = InstaGeoChipDataset(list(raw), [f"chip_{i}.tif" for i in range(5)],
                             make_preprocess_func(FLOOD_MEAN, FLOOD_STD, 1, 224), -9999, 1e-4)

    def collate(batch):  # infer_collate_fn, instageo/model/pipeline_utils.py:92-104
        return (torch.stack([a[0][0] for a in batch], 0), [a[0][1] for a in batch]), [a[1] for a in batch]

    dl = torch.utils.data.DataLoader(ds, batch_size=2, collate_fn=collate, num_workers=0)
    got = {}
    info = chip_inference(dl, None, m, device="gpu", writer=lambda p, f, o: got.__setitem__(f, p))
    assert info == {"predictions": 5} and len(got) == 5
    x = torch.from_numpy(np.stack([OP.preprocess_chip(r, None, 1e-4, FLOOD_MEAN, FLOOD_STD, 1, None)[0] for r in raw]))
    ref = P.prithvi_seg_forward(x, sd, 4, 1)
    for i in range(5):
        p = got[f"chip_{i}.tif"]
        assert p.dtype == np.int8 and p.shape == (224, 224)
        top2 = ref[i].topk(2, dim=0).values
        safe = ((top2[0] - top2[1]) > 1e-2).numpy()
        assert (p == ref[i].argmax(0).numpy())[safe].all()


@pytest.mark.parametrize("T,nc", [(1, 2), (3, 13)])
def test_predict_step_probability_and_regression_head(cuda_dev, T, nc):
    """predict_step variants: softmax(logits)[:, 1] fused into the head epilogue
    (segmentation.py:202-213) and the 1-channel regression head's forward().squeeze(1) (regression.py:338-339)."""
    from oracle import metrics as OM
    m, sd = _build("prithvi_eo_tiny", T, nc, 2, cuda_dev)
    x = model_input(11, 2, T)
    ref = P.prithvi_seg_forward(x, sd, P.VARIANTS["prithvi_eo_tiny"][2], T)
    want = OM.positive_probability(ref)
    got = m.predict_proba(x.to(cuda_dev)).cpu()
    assert got.shape == want.shape and got.dtype == torch.float32
    # |d softmax / d logit| <= 1/4 per logit, so a 2e-2 logit budget gives at most 1e-2 on a probability
    assert (got - want).abs().max().item() < 1e-2
    # the same epilogue without the logits round trip: identical to softmax of the engine's own logits to 2 ulp
    own = torch.softmax(m(x.to(cuda_dev)), dim=1)[:, 1].cpu()
    assert (got - own).abs().max().item() < 1e-6
    if T == 1:
        r, sdr = _build("prithvi_eo_tiny", 1, 1, 2, cuda_dev)
        yr = r(x.to(cuda_dev))
        assert yr.shape == (2, 1, 224, 224)
        refr = P.prithvi_seg_forward(x, sdr, P.VARIANTS["prithvi_eo_tiny"][2], 1)
        assert (yr.cpu() - refr).abs().max().item() < TOL
        with pytest.raises(RuntimeError, match="classification head"):
            r.predict_proba(x.to(cuda_dev))


@pytest.mark.parametrize("variant,T,nc,B", [("prithvi_eo_v1_100", 1, 2, 2), ("prithvi_eo_v1_100", 3, 13, 1),
                                            ("prithvi_eo_v2_300", 3, 13, 1)])
def test_full_depth_configs_vs_oracle(cuda_dev, variant, T, nc, B):
    """BASELINE.json configs[0..2] at FULL depth (12 / 12 / 24 blocks, default and stress init): the bf16 engine against
    the fp32 CPU oracle within the north-star budget (2e-2 max-abs on logits), argmax identical outside ties."""
    for stress in (False, True):
        m, sd = _build(variant, T, nc, -1, cuda_dev, seed=3, stress=stress)
        x = torch.randn(B, 6, T, 224, 224, generator=torch.Generator().manual_seed(7 + T))
        ref = P.prithvi_seg_forward(x, sd, P.VARIANTS[variant][2], T)
        y = m(x.to(cuda_dev)).cpu()
        eps = (y - ref).abs().max().item()
        scale = ref.abs().max().item()
        assert eps < TOL, f"{variant} T={T} stress={stress}: logits max-abs {eps} (|logits| up to {scale})"
        excluded = _argmax_check(m.predict(x.to(cuda_dev)), ref, eps)
        print(f"{variant} T={T} nc={nc} stress={stress}: max-abs {eps:.3e} of {scale:.2f}, tie-excluded pixels {excluded:.4f}")
        del m
        torch.cuda.empty_cache()


@pytest.mark.parametrize("variant,T,nc,depth,B", [("prithvi_eo_tiny", 1, 2, 4, 2), ("prithvi_eo_v1_100", 3, 13, 3, 1),
                                                  ("prithvi_eo_v2_300", 1, 2, 2, 2)])
def test_per_block_taps_vs_oracle(cuda_dev, variant, T, nc, depth, B):
    """SURVEY.md Appendix F 'patch-embed, per-block': the residual stream after the patch embed, after every block
    and after the final LayerNorm against the fp32 oracle's taps (bf16 GEMM operands, f32 accumulation / residual:
    tolerance 1e-2 of the tap's own scale, written here)."""
    m, sd = _build(variant, T, nc, depth, cuda_dev, stress=True)
    x = torch.randn(B, 6, T, 224, 224, generator=torch.Generator().manual_seed(40 + B))
    taps = {}
    ref = P.prithvi_seg_forward(x, sd, P.VARIANTS[variant][2], T, taps=taps)
    m.enable_taps(True)
    y = m(x.to(cuda_dev))
    assert not m.graph_status()["last_forward_was_graph"]
    D, N = P.VARIANTS[variant][0], 1 + T * 196
    worst = 0.0
    for name in ["embed"] + [f"block{i}" for i in range(depth)] + ["tokens"]:
        got = m.debug_tap(name, B, (B, N, D)).cpu()
        want = taps[name]
        rel = (got - want).abs().max().item() / max(want.abs().max().item(), 1e-6)
        worst = max(worst, rel)
        assert rel < 1e-2, f"{name}: max-abs error {rel:.3e} of the tap's scale"
    print(f"{variant} T={T}: worst per-block relative error {worst:.3e}")
    assert (y.cpu() - ref).abs().max().item() < TOL
    with pytest.raises(Exception):
        m.debug_tap(f"block{depth}", B, (B, N, D))
    m.enable_taps(False)
    y2 = m(x.to(cuda_dev))
    assert torch.equal(y, y2)  # the taps do not change the result


def test_graph_replay_follows_caller_pointers(cuda_dev):
    """The production entry replays a CUDA graph; x / logits / argmax pointers that change between calls are patched
    into the graph, and the result is bit-identical to the kernel-by-kernel path (the f32 entry)."""
    from instageo_b200 import ops
    from oracle import preprocess as OP
    from conftest import FLOOD_MEAN, FLOOD_STD
    m, sd = _build("prithvi_eo_tiny", 1, 2, 2, cuda_dev)
    spec = ops.PreprocessSpec(FLOOD_MEAN, FLOOD_STD, 1, None, 1e-4, None, cuda_dev)
    raws = [torch.from_numpy(OP.synth_chips(3, 1, seed=s)).to(cuda_dev) for s in (1, 2)]
    pres = [ops.preprocess(r, spec, want_f32=True, want_patches=True) for r in raws]
    keep = []
    for rep in range(3):
        for pre in pres:
            la, aa = m.forward_patches(pre["patches"], want_logits=True, want_argmax=True)
            st = m.graph_status()
            assert st["enabled"], st["note"]
            assert st["last_forward_was_graph"] and st["kernels_in_graph"] == m.launches_per_forward() - 1
            keep.append((la, aa, pre))  # outputs stay alive: every call gets NEW output addresses
    for la, aa, pre in keep:
        lb = m(pre["f32"])
        assert torch.equal(la, lb) and torch.equal(aa.long(), lb.argmax(1))
    # argmax-only graph (another flag set) next to the first one, and a second batch size on the same workspace
    a1 = m.forward_patches(pres[0]["patches"], want_logits=False, want_argmax=True)[1]
    assert torch.equal(a1, keep[0][1])
    rows = 196
    a2 = m.forward_patches(pres[1]["patches"][:2 * rows], want_logits=False, want_argmax=True)[1]
    assert torch.equal(a2, keep[1][1][:2])
    a3 = m.forward_patches(pres[0]["patches"], want_logits=False, want_argmax=True)[1]
    assert torch.equal(a3, keep[0][1])


def test_timed_configuration_multiwave_vs_oracle(cuda_dev):
    """BASELINE.json configs[1] as bench.py times it: V1-100M, T=3, 13 classes, FULL depth, through
    preprocess -> tubelet rows -> forward_patches(argmax) at B = 64 (M = 37 696 token rows: several waves of the
    persistent GEMMs and of the attention kernel).  The first 16 chips are also run as their own batch (still
    multi-wave) and compared with the fp32 oracle; the B = 64 result must contain them bit for bit."""
    from instageo_b200 import ops
    from oracle import preprocess as OP
    from conftest import CROP_MEAN, CROP_STD
    variant, T, nc = "prithvi_eo_v1_100", 3, 13
    m, sd = _build(variant, T, nc, -1, cuda_dev, seed=0, stress=True)
    raw = OP.synth_chips(64, T, seed=1042, nodata=None)
    spec = ops.PreprocessSpec(CROP_MEAN, CROP_STD, T, None, 1.0, None, cuda_dev)
    d_raw = torch.from_numpy(raw).to(cuda_dev)
    pre64 = ops.preprocess(d_raw, spec, want_f32=False, want_patches=True)
    l64, a64 = m.forward_patches(pre64["patches"], want_logits=True, want_argmax=True)
    pre16 = ops.preprocess(d_raw[:16], spec, want_f32=False, want_patches=True)
    l16, a16 = m.forward_patches(pre16["patches"], want_logits=True, want_argmax=True)
    assert torch.equal(l64[:16], l16) and torch.equal(a64[:16], a16)
    assert torch.equal(a64.long(), l64.argmax(1))
    n_ref = 4  # oracle cost: ~1 s per chip on the host cores
    xr = np.stack([OP.preprocess_chip(r, None, 1.0, CROP_MEAN, CROP_STD, T, None)[0] for r in raw[:n_ref]])
    ref = P.prithvi_seg_forward(torch.from_numpy(xr), sd, P.VARIANTS[variant][2], T)
    eps = (l64[:n_ref].cpu() - ref).abs().max().item()
    assert eps < TOL, f"B=64 logits max-abs {eps}"
    # the last chips of the batch sit in the last, partially filled wave: check them as well
    xr2 = np.stack([OP.preprocess_chip(r, None, 1.0, CROP_MEAN, CROP_STD, T, None)[0] for r in raw[62:]])
    ref2 = P.prithvi_seg_forward(torch.from_numpy(xr2), sd, P.VARIANTS[variant][2], T)
    eps2 = (l64[62:].cpu() - ref2).abs().max().item()
    assert eps2 < TOL, f"B=64 (last chips) logits max-abs {eps2}"
    excluded = _argmax_check(a64[:n_ref], ref, max(eps, eps2))
    print(f"B=64 timed configuration: max-abs {max(eps, eps2):.3e}, tie-excluded {excluded:.4f}")


def test_chip_pipeline_unpinned_sources_and_short_last_batch(cuda_dev):
    """ChipPipeline.run with pageable (non-pinned) host batches -- the staging buffer of a slot must not be
    overwritten while its previous H2D copy is in flight (tiny model: the copies are the slow part) -- and with a
    last batch shorter than ``batch``; every mask equals the unpipelined path's."""
    from instageo_b200 import ops
    from instageo_b200.model.infer_utils import ChipPipeline
    from oracle import preprocess as OP
    from conftest import FLOOD_MEAN, FLOOD_STD
    m, sd = _build("prithvi_eo_tiny", 1, 2, 0, cuda_dev)
    spec = ops.PreprocessSpec(FLOOD_MEAN, FLOOD_STD, 1, None, 1e-4, None, cuda_dev)
    B = 8
    sizes = [B] * 9 + [3, 1]
    batches = [OP.synth_chips(n, 1, seed=100 + i) for i, n in enumerate(sizes)]   # numpy: pageable memory
    want = []
    for raw in batches:
        pre = ops.preprocess(torch.from_numpy(raw).to(cuda_dev), spec, want_f32=False, want_patches=True)
        want.append(m.forward_patches(pre["patches"], want_logits=False, want_argmax=True)[1].cpu().numpy())
    pipe = ChipPipeline(m, spec, B, cuda_dev)
    got = []
    n = pipe.run(batches, consume=lambda a: got.append(a.copy()))
    assert n == sum(sizes) and [g.shape[0] for g in got] == sizes
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    # pinned sources take the direct path; same masks
    got2 = []
    pipe.run([torch.from_numpy(b).pin_memory() for b in batches], consume=lambda a: got2.append(a.copy()))
    assert all(np.array_equal(g, w) for g, w in zip(got2, want))
    with pytest.raises(ValueError):
        pipe.run([batches[0][:, :5]])
