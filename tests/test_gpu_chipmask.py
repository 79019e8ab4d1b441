"""GPU parity of the chip-creation masking kernel (csrc/chipmask.cu) against oracle/preprocess.py
(apply_fmask / mask_segmentation_map / create_chip, themselves pinned by the reference's known-answer
tests in tests/test_oracle.py).  Integer work: bit-exact."""
import numpy as np
import pytest
import torch

from oracle import preprocess as OP

pytestmark = pytest.mark.gpu


def _case(seed, T, H, W, dtype=np.int16):
    rng = np.random.default_rng(seed)
    chip = rng.integers(-200, 10400, size=(T * 6, H, W)).astype(dtype) if dtype == np.int16 else \
        rng.integers(0, 10400, size=(T * 6, H, W)).astype(dtype)
    hole = rng.random((T, H, W)) < 0.03
    for t in range(T):
        chip[t * 6:(t + 1) * 6][np.broadcast_to(hole[t], (6, H, W))] = 0 if dtype == np.uint16 else -9999
    fmask = np.zeros((T, H, W), dtype=np.uint8)
    for bit in (0, 1, 2, 3, 4, 5, 6, 7):
        fmask |= ((rng.random((T, H, W)) < 0.06).astype(np.uint8) << bit)
    seg = rng.integers(-1, 5, size=(H, W)).astype(np.int8)
    return chip, fmask, seg


@pytest.mark.parametrize("T,H,W,dtype", [(1, 224, 224, np.int16), (3, 224, 224, np.int16), (3, 96, 100, np.uint16),
                                         (2, 37, 53, np.int16), (1, 1, 1, np.int16), (3, 366, 366, np.int16)])
@pytest.mark.parametrize("strategy", ["each", "any"])
def test_create_chip_matches_oracle(cuda_dev, T, H, W, dtype, strategy):
    from instageo_b200.data import create_chip
    chip, fmask, seg = _case(T * 1000 + H, T, H, W, dtype)
    for types in (("cloud", "near_cloud_or_shadow", "cloud_shadow", "water"), ("cloud",), ("water", "bogus")):
        out, seg_out, counts = create_chip(torch.from_numpy(chip).to(cuda_dev), fmask, seg, strategy, types)
        want, want_seg, n_valid, n_lab = OP.create_chip(chip, fmask, seg, strategy, types)
        assert out.dtype == torch.uint16 and seg_out.dtype == torch.int8
        assert np.array_equal(out.cpu().numpy(), want) and np.array_equal(seg_out.cpu().numpy(), want_seg)
        assert counts.tolist() == [n_valid, n_lab]


def test_apply_mask_and_label_mask_alone(cuda_dev):
    from instageo_b200.data import apply_mask, decode_fmask_value, mask_segmentation_map
    chip, fmask, seg = _case(5, 3, 64, 72)
    for strategy in ("each", "any"):
        got = apply_mask(chip, fmask, 0, masking_strategy=strategy)
        assert got.dtype == torch.int16  # no clip: negative reflectances survive, as in the reference
        assert np.array_equal(got.cpu().numpy(), OP.apply_fmask(chip, fmask, 0, strategy).astype(np.int16))
        got = mask_segmentation_map(chip, seg, -9999, strategy)
        assert np.array_equal(got.cpu().numpy(), OP.mask_segmentation_map(chip, seg, -9999, strategy))
    with pytest.raises(ValueError, match="Invalid masking strategy"):
        apply_mask(chip, fmask, 0, masking_strategy="some")
    # known answers of the reference: decode_fmask_value(100, 0..7), tests/data_tests/test_hls_utils.py:145-159
    v = torch.tensor([100], device=cuda_dev)
    assert [int(decode_fmask_value(v, p)) for p in range(8)] == [0, 0, 1, 0, 0, 1, 1, 0]
    # known answers of the reference: tests/data_tests/test_create_chips.py:91-139 (chip [3 bands, 1 x 4 px])
    c = np.array([[1, 2, 3, 4], [1, 3, -9, 7], [6, 7, 3, 9]], dtype=np.int16).reshape(3, 1, 4)
    s = np.array([[1, -1, 1, 2]], dtype=np.int8)
    each = mask_segmentation_map(c, s, -9, "each")
    assert each.cpu().tolist() == [[1, -1, 1, 2]]
    assert mask_segmentation_map(c, each, -9, "any").cpu().tolist() == [[1, -1, -1, 2]]
    c = np.array([[1, 2, 3, 4], [-9, -9, -9, -9], [6, 7, 3, 9]], dtype=np.int16).reshape(3, 1, 4)
    assert mask_segmentation_map(c, s, -9).cpu().tolist() == [[-1, -1, -1, -1]]


@pytest.mark.parametrize("dtype,no_data,clip", [(np.int16, -9999, (0, 10000)), (np.int16, 0, None), (np.int16, -9999, None),
                                                (np.uint16, 0, (0, 65535)), (np.uint16, 65535, (100, 60000)),
                                                (np.int16, 40000, (0, 65535)), (np.int16, 7, (5, 9)),
                                                (np.uint16, 70000, (0, 10000)), (np.int16, -5, (40000, 50000))])
def test_fill_value_and_clip_corner_cases(cuda_dev, dtype, no_data, clip):
    """full-range inputs, fill values outside the clip range / the element type: the packed 16-bit path and the
    scalar path it falls back to must both follow the integer arithmetic of the oracle."""
    from instageo_b200.data import create_chip
    rng = np.random.default_rng(abs(no_data) + 1)
    info = np.iinfo(dtype)
    chip = rng.integers(info.min, info.max + 1, size=(12, 40, 48)).astype(dtype)
    chip[rng.random(chip.shape) < 0.05] = np.clip(no_data, info.min, info.max)
    fmask = (rng.random((2, 40, 48)) < 0.3).astype(np.uint8) * 8
    seg = rng.integers(-1, 3, size=(40, 48)).astype(np.int8)
    for strategy in ("each", "any"):
        out, seg_out, counts = create_chip(chip, fmask, seg, strategy, no_data_value=no_data, clip=clip)
        want, want_seg, n_valid, n_lab = OP.create_chip(chip, fmask, seg, strategy, no_data_value=no_data, clip=clip)
        got = out.cpu().numpy()
        assert np.array_equal(got.astype(np.uint16), want.astype(np.uint16))  # compare 16-bit patterns
        assert np.array_equal(seg_out.cpu().numpy(), want_seg) and counts.tolist() == [n_valid, n_lab]


def test_whole_tile_properties(cuda_dev):
    """3660 x 3660 x 6 tile (BASELINE configs[3] geometry): idempotence and count consistency."""
    from instageo_b200.data import create_chip
    g = torch.Generator(device="cuda").manual_seed(1042)
    tile = torch.randint(-100, 10200, (6, 3660, 3660), generator=g, device=cuda_dev, dtype=torch.int16)
    fm = (torch.rand((1, 3660, 3660), generator=g, device=cuda_dev) < 0.2).to(torch.uint8) * 2  # cloud bit
    out, _, counts = create_chip(tile, fm, None, "each")
    o = out.view(torch.int16)
    assert int(o.max()) <= 10000 and int(o.min()) >= 0
    cloudy = (fm[0] & 2).bool()
    assert int(o[:, cloudy].abs().sum()) == 0
    assert torch.equal(o[:, ~cloudy], tile[:, ~cloudy].clamp(0, 10000))
    assert int(counts[0]) == int((o != 0).sum())
    again, _, counts2 = create_chip(o, fm, None, "each")   # idempotent
    assert torch.equal(again.view(torch.int16), o) and int(counts2[0]) == int(counts[0])


def test_label_maps_are_not_silently_truncated(cuda_dev):
    """The reference keeps the label map's dtype (float32 regression targets, hls_utils.py:398); the device kernel carries
    int8 class labels: anything a cast would change is refused, int16 / int64 labels in range are taken."""
    import numpy as np
    import torch
    from instageo_b200.data import create_chip
    chip = np.random.default_rng(0).integers(0, 9000, size=(6, 32, 40)).astype(np.int16)
    fm = np.zeros((1, 32, 40), np.uint8)
    with pytest.raises(TypeError, match="integer class labels"):
        create_chip(chip, fm, np.zeros((32, 40), np.float32))
    with pytest.raises(ValueError, match="int8"):
        create_chip(chip, fm, np.full((32, 40), 300, np.int32))
    seg = np.random.default_rng(1).integers(-1, 100, size=(32, 40))
    out = create_chip(chip, fm, seg.astype(np.int64))[1]
    assert out.dtype == torch.int8 and np.array_equal(out.cpu().numpy(), seg.astype(np.int8))
