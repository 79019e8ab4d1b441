"""Freeze golden vectors from the REAL reference modules (dev container only).

    python -m oracle.gen_golden          # writes tests/golden/*.npz

Inputs are regenerated from seeds by the tests; only the reference's OUTPUTS (and the
small raw inputs of the preprocessing cases) are stored, so fixtures stay small.
Reference entry points exercised:
  instageo.model.dataloader.process_and_augment / process_test / crop_array
  instageo.model.model.PrithviSeg (random init replaced by oracle.prithvi.make_state_dict)
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import preprocess as PP
from . import prithvi as P
from . import refstub

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

FLOOD_MEAN = [0.14245495, 0.13921481, 0.12434631, 0.31420089, 0.20743526, 0.12046503]
FLOOD_STD = [0.04036231, 0.04186983, 0.05267646, 0.0822221, 0.06834774, 0.05294205]
CROP_MEAN = [494.905781, 815.239594, 924.335066, 2968.881459, 2634.621962, 1739.579917]
CROP_STD = [284.925432, 357.84876, 575.566823, 896.601013, 951.900334, 921.407808]

MODEL_CASES = {
    # name: (variant, temporal, classes, depth, weight seed, stress, input seed, batch)
    "tiny_t1": ("prithvi_eo_tiny", 1, 2, 2, 11, True, 21, 1),
    "tiny_t3": ("prithvi_eo_tiny", 3, 13, 1, 12, True, 22, 1),
}


def model_input(seed: int, batch: int, temporal: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, 6, temporal, 224, 224, generator=g)


def gen_preprocess(dl):
    out = {}
    rng = np.random.default_rng(7)
    # (a) int16, T=3, nodata -9999, cm in {1, 1e-4}
    raw = PP.synth_chips(1, 3, seed=101, size=32)[0]
    out["a_raw"] = raw
    for tag, cm in (("cm1", 1.0), ("cm1e4", 1e-4)):
        arr = raw * cm
        t, _ = dl.process_and_augment(arr, None, FLOOD_MEAN, FLOOD_STD, temporal_size=3,
                                      im_size=32, crop=False)
        out[f"a_{tag}_out"] = t.numpy()
        out[f"a_{tag}_mask"] = arr == -9999  # dataloader.py:899
    # (b) uint16, T=1, 6-of-13 band gather, crop means (raw DN scale)
    raw_b = rng.integers(0, 10001, size=(13, 32, 32)).astype(np.uint16)
    raw_b[:, 3:6, 3:6] = 0
    bands = [1, 2, 3, 8, 11, 12]
    arr = raw_b[bands, ...] * 1.0
    t, _ = dl.process_and_augment(arr, None, CROP_MEAN, CROP_STD, temporal_size=1, im_size=32,
                                  crop=False)
    out["b_raw"], out["b_bands"], out["b_out"], out["b_mask"] = raw_b, np.array(bands), t.numpy(), arr == 0
    # (c) process_test grid: 80x80, crop 32, stride 16 -> 4x4 windows
    raw_c = rng.integers(0, 10001, size=(6, 80, 80)).astype(np.int16)
    lab = np.zeros((80, 80), dtype=np.float32)
    imgs, _ = dl.process_test(raw_c * 1e-4, lab, FLOOD_MEAN, FLOOD_STD, temporal_size=1,
                              img_size=80, crop_size=32, stride=16)
    out["c_raw"], out["c_out"] = raw_c, imgs.numpy()
    np.savez_compressed(os.path.join(GOLDEN, "preprocess.npz"), **out)
    print("preprocess.npz", {k: v.shape for k, v in out.items()})


def gen_model():
    for name, (variant, T, nc, depth, wseed, stress, iseed, batch) in MODEL_CASES.items():
        ref = refstub.reference_prithvi_seg(temporal_step=T, num_classes=nc, variant=variant,
                                            depth=depth)
        sd = P.make_state_dict(variant, T, nc, depth=depth, seed=wseed, stress=stress)
        ref.load_state_dict(sd, strict=True)
        x = model_input(iseed, batch, T)
        with torch.no_grad():
            y, feat = ref(x, return_features=True)
        out = {
            "logits_sub": y[:, :, ::4, ::4].numpy(),
            "logits_sum": np.array(y.double().sum().item()),
            "logits_absmax": np.array(y.abs().max().item()),
            "argmax": torch.argmax(y, dim=1).numpy().astype(np.int8),
            "feat_sub": feat[:, ::16].numpy(),
        }
        np.savez_compressed(os.path.join(GOLDEN, f"model_{name}.npz"), **out)
        print(name, {k: v.shape for k, v in out.items()})


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    gen_preprocess(refstub.reference_dataloader())
    gen_model()


if __name__ == "__main__":
    main()
