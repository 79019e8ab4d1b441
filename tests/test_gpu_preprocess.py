"""GPU parity: kernel 1 (fused normalise / mask / window crop) through the C ABI vs the oracle
and vs the golden vectors frozen from the reference.  Bar: BIT-exact (0 ulp) values and masks."""
import os

import numpy as np
import pytest
import torch

from conftest import CROP_MEAN, CROP_STD, FLOOD_MEAN, FLOOD_STD, GOLDEN
from oracle import preprocess as OP

pytestmark = pytest.mark.gpu


def _run(raw, dev, mean, std, T, bands=None, cm=1.0, nd=None, **kw):
    from instageo_b200 import ops
    from instageo_b200.model.dataloader import _to_device
    spec = ops.PreprocessSpec(mean, std, T, bands, cm, nd, dev)
    x = _to_device(raw, dev)
    return ops.preprocess(x, spec, win=kw.pop("win", raw.shape[-1]), **kw)


def test_golden_from_reference(cuda_dev):
    g = np.load(os.path.join(GOLDEN, "preprocess.npz"))
    for tag, cm in (("cm1", 1.0), ("cm1e4", 1e-4)):
        out = _run(g["a_raw"][None], cuda_dev, FLOOD_MEAN, FLOOD_STD, 3, cm=cm, nd=-9999, want_f32=True,
                   want_mask_elem=True)
        assert np.array_equal(out["f32"][0].cpu().numpy(), g[f"a_{tag}_out"])
        assert np.array_equal(out["mask_elem"][0].cpu().numpy(), g[f"a_{tag}_mask"])
    out = _run(g["b_raw"][None], cuda_dev, CROP_MEAN, CROP_STD, 1, bands=g["b_bands"].tolist(), nd=0, want_f32=True,
               want_mask_elem=True)
    assert np.array_equal(out["f32"][0].cpu().numpy(), g["b_out"])
    assert np.array_equal(out["mask_elem"][0].cpu().numpy(), g["b_mask"])
    # process_test golden through the drop-in function (float64 input, like the reference call)
    from instageo_b200.model.dataloader import process_test
    imgs, labels = process_test(g["c_raw"] * 1e-4, np.zeros((80, 80), np.float32), FLOOD_MEAN, FLOOD_STD,
                                temporal_size=1, img_size=80, crop_size=32, stride=16)
    assert np.array_equal(imgs.cpu().numpy(), g["c_out"]) and labels.shape == (16, 32, 32)


@pytest.mark.parametrize("T,cm,nd,dtype", [(1, 1.0, -9999, np.int16), (3, 1e-4, -9999, np.int16), (3, 1.0, 0, np.uint16),
                                           (3, 0.001, None, np.int16), (2, 1e-4, 0, np.uint16)])
def test_chips_bit_exact(cuda_dev, T, cm, nd, dtype):
    raw = OP.synth_chips(3, T, seed=11 + T, nodata=nd if (nd is None or nd >= 0 or dtype == np.int16) else 0, dtype=dtype)
    out = _run(raw, cuda_dev, FLOOD_MEAN, FLOOD_STD, T, cm=cm, nd=nd, want_f32=True, want_patches=True,
               want_mask_elem=True, want_mask_px=True)
    ref = [OP.preprocess_chip(r, None, cm, FLOOD_MEAN, FLOOD_STD, T, nd) for r in raw]
    rx, rm = np.stack([r[0] for r in ref]), np.stack([r[1] for r in ref])
    got = out["f32"].cpu().numpy()
    assert got.dtype == np.float32 and got.shape == (3, 6, T, 224, 224)
    assert np.array_equal(got, rx), f"max ulp-ish diff {np.abs(got - rx).max()}"
    assert np.array_equal(out["mask_elem"].cpu().numpy(), rm)
    assert np.array_equal(out["mask_px"].cpu().numpy(), rm.any(axis=1))
    # tubelet rows = bf16(round-to-nearest-even) of the very same values, in Conv3d weight order
    p = out["patches"].float().cpu().reshape(3, T, 14, 14, 6, 16, 16)
    want = torch.from_numpy(rx).bfloat16().float().reshape(3, 6, T, 14, 16, 14, 16).permute(0, 2, 3, 5, 1, 4, 6)
    assert torch.equal(p, want)


def test_windows_unaligned_and_float64(cuda_dev):
    """Windows of a 3660-wide tile start at 2/8-byte aligned addresses only (row pitch 7320 B)."""
    rng = np.random.default_rng(2)
    tile = rng.integers(-100, 10001, size=(6, 301, 333)).astype(np.int16)
    from instageo_b200 import ops
    wins = [(0, 0, 0), (0, 5, 3), (0, 77, 109), (0, 301 - 64, 333 - 64), (0, 1, 2)]
    spec = ops.PreprocessSpec(FLOOD_MEAN, FLOOD_STD, 1, None, 1e-4, None, cuda_dev)
    wt = torch.tensor(wins, dtype=torch.int32, device=cuda_dev)
    got = ops.preprocess(torch.from_numpy(tile).to(cuda_dev)[None], spec, windows=wt, win=64)["f32"].cpu().numpy()
    for i, (_, t, l) in enumerate(wins):
        want = OP.normalize(tile[:, t:t + 64, l:l + 64] * 1e-4, FLOOD_MEAN, FLOOD_STD, 1)
        assert np.array_equal(got[i], want)
    # the reference hands the already multiplied float64 array to process_and_augment
    from instageo_b200.model.dataloader import normalize_and_convert_to_tensor, process_and_augment
    arr = tile[:, :64, :64] * 1e-4
    t1, _ = process_and_augment(arr, None, FLOOD_MEAN, FLOOD_STD, temporal_size=1, im_size=64, crop=False)
    assert t1.shape == (6, 1, 64, 64) and np.array_equal(t1.cpu().numpy(), OP.normalize(arr, FLOOD_MEAN, FLOOD_STD, 1))
    t2, _ = normalize_and_convert_to_tensor(list(arr.astype(np.float32)), None, FLOOD_MEAN, FLOOD_STD, 1)
    assert np.array_equal(t2.cpu().numpy(), t1.cpu().numpy())
    # reference shape pins (tests/model_tests/test_dataloader.py:42-65)
    x2 = rng.random((6, 224, 224))
    t3, lab = process_and_augment(x2, np.zeros((224, 224)), [0.5] * 3, [0.2] * 3, temporal_size=2)
    assert t3.shape == (3, 2, 224, 224) and lab.shape == (224, 224)
    # crop=True on a larger chip draws a RandomCrop origin from torch's RNG
    torch.manual_seed(1042)
    big = rng.integers(0, 10000, size=(6, 256, 256)).astype(np.float64)
    t4, _ = process_and_augment(big, None, CROP_MEAN, CROP_STD, temporal_size=1, im_size=224, crop=True)
    torch.manual_seed(1042)
    i = int(torch.randint(0, 33, (1,)).item()); j = int(torch.randint(0, 33, (1,)).item())
    assert np.array_equal(t4.cpu().numpy(), OP.normalize(big[:, i:i + 224, j:j + 224], CROP_MEAN, CROP_STD, 1))


@pytest.mark.parametrize("strategy", ["each", "any"])
def test_fmask_cloud_masking(cuda_dev, strategy):
    rng = np.random.default_rng(5)
    T = 3
    raw = rng.integers(1, 10001, size=(2, 18, 64, 64)).astype(np.uint16)
    fm = np.zeros((2, T, 64, 64), np.uint8)
    for bit in (1, 2, 3, 5, 6):  # 6 is not a decode position used below
        fm |= ((rng.random(fm.shape) < 0.05).astype(np.uint8) << bit)
    bits = sum(1 << p for p in OP.HLS_FMASK_POS.values())
    from instageo_b200 import ops
    spec = ops.PreprocessSpec(CROP_MEAN, CROP_STD, T, None, 1.0, 0, cuda_dev)
    out = ops.preprocess(torch.from_numpy(raw.view(np.int16)).to(cuda_dev).view(torch.uint16), spec, win=64,
                         want_f32=True, want_mask_px=True, fmask=torch.from_numpy(fm).to(cuda_dev), fmask_bits=bits,
                         masking_strategy=strategy)
    for i in range(2):
        masked = OP.apply_fmask(raw[i].astype(np.int64), fm[i], 0, strategy)
        want, m = OP.preprocess_chip(masked, None, 1.0, CROP_MEAN, CROP_STD, T, 0)
        assert np.array_equal(out["f32"][i].cpu().numpy(), want)
        assert np.array_equal(out["mask_px"][i].cpu().numpy(), m.any(axis=0))


def test_edge_cases(cuda_dev):
    from instageo_b200 import _lib, ops
    spec = ops.PreprocessSpec(FLOOD_MEAN, FLOOD_STD, 1, device=cuda_dev)
    empty = torch.zeros((0, 6, 224, 224), dtype=torch.int16, device=cuda_dev)
    assert ops.preprocess(empty, spec)["f32"].shape == (0, 6, 1, 224, 224)
    with pytest.raises(_lib.IgError):
        ops.preprocess(torch.zeros((1, 6, 100, 100), dtype=torch.int16, device=cuda_dev), spec, win=100)
    with pytest.raises(IndexError):
        ops.preprocess(torch.zeros((1, 4, 224, 224), dtype=torch.int16, device=cuda_dev), spec)
    # int16 extremes survive the f64 product
    raw = torch.tensor([-32768, 32767, -9999, 0] * 4, dtype=torch.int16, device=cuda_dev).repeat(6 * 16 * 1).reshape(1, 6, 16, 16)
    o = ops.preprocess(raw, ops.PreprocessSpec(FLOOD_MEAN, FLOOD_STD, 1, None, 1e-4, None, cuda_dev), win=16)["f32"]
    want = OP.normalize(raw[0].cpu().numpy() * 1e-4, FLOOD_MEAN, FLOOD_STD, 1)
    assert np.array_equal(o[0].cpu().numpy(), want)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.int16, np.uint16])
@pytest.mark.parametrize("cm", [1.0, 1e-4])
def test_every_raw_value_divides_exactly(cuda_dev, dtype, cm):
    """The fast path replaces the IEEE division by an FMA sequence on a precomputed reciprocal:
    check all 65536 raw values x 6 bands against the oracle's true division, for the two config
    statistics and for adversarial divisors (tiny, huge, near powers of two, all-ones significand)."""
    from instageo_b200 import ops
    vals = np.arange(65536, dtype=np.uint16).view(dtype).reshape(256, 256)
    raw = np.broadcast_to(vals, (1, 6, 256, 256)).copy()
    rng = np.random.default_rng(5)
    stat_sets = [(FLOOD_MEAN, FLOOD_STD), (CROP_MEAN, CROP_STD)]
    for _ in range(3):
        stat_sets.append((list(rng.uniform(-3000, 3000, 6)), list(np.exp(rng.uniform(-9, 9, 6)))))
    ones = float(np.frombuffer(np.uint32(0x3fffffff).tobytes(), dtype=np.float32)[0])  # 1.9999999
    stat_sets.append(([0.0, 1.0, -1.0, 0.5, 1e-3, 123.456], [1.0, ones, 3.0, 1.0000001, 0.99999994, 7.0]))
    stat_sets.append(([0.0, 1e-20, 5.0, -7.0, 0.25, 1e3], [1e-25, 1e25, 3e-19, 2e18, 1.5e-18, 0.1]))  # outside the FMA path's safe range
    for mean, std in stat_sets:
        spec = ops.PreprocessSpec(mean, std, 1, None, cm, -9999, cuda_dev)
        torch_raw = torch.from_numpy(raw.view(np.int16)).to(cuda_dev)
        if dtype == np.uint16:
            torch_raw = torch_raw.view(torch.uint16)
        out = ops.preprocess(torch_raw, spec, win=256, want_f32=True, want_mask_elem=True)
        want, mask = OP.preprocess_chip(raw[0], None, cm, mean, std, 1, -9999)
        assert np.array_equal(out["f32"][0].cpu().numpy(), want), (mean, std)
        assert np.array_equal(out["mask_elem"][0].cpu().numpy(), mask)
