"""Dev container only: the oracle must be BIT-identical to the real reference modules
(imported with the stub recipe of oracle/refstub.py).  Skipped where /root/reference is absent."""
import numpy as np
import pytest
import torch

from conftest import FLOOD_MEAN, FLOOD_STD
from oracle import preprocess as OP
from oracle import prithvi as P
from oracle import refstub

pytestmark = pytest.mark.skipif(not refstub.available(), reason="reference tree not present (GPU box)")


@pytest.mark.parametrize("variant,T,nc,depth", [("prithvi_eo_tiny", 1, 2, -1), ("prithvi_eo_v1_100", 3, 13, 1)])
def test_forward_bit_identical(variant, T, nc, depth):
    ref = refstub.reference_prithvi_seg(temporal_step=T, num_classes=nc, variant=variant, depth=depth)
    sd = P.make_state_dict(variant, T, nc, depth=depth, seed=3, stress=True)
    assert set(sd) == set(ref.state_dict())
    ref.load_state_dict(sd, strict=True)
    x = torch.randn(1, 6, T, 224, 224, generator=torch.Generator().manual_seed(9))
    with torch.no_grad():
        y_ref, f_ref = ref(x, return_features=True)
    y, f = P.prithvi_seg_forward(x, sd, P.VARIANTS[variant][2], T, return_features=True)
    assert torch.equal(y, y_ref) and torch.equal(f, f_ref)


def test_preprocess_bit_identical():
    dl = refstub.reference_dataloader()
    raw = OP.synth_chips(1, 3, seed=5, size=64)[0]
    for cm in (1.0, 1e-4, 0.001):
        arr = raw * cm
        t_ref, _ = dl.process_and_augment(arr, None, FLOOD_MEAN, FLOOD_STD, temporal_size=3, im_size=64, crop=False)
        assert np.array_equal(t_ref.numpy(), OP.normalize(arr, FLOOD_MEAN, FLOOD_STD, 3))
    assert np.array_equal(dl.crop_array(raw, 3, 5, 20, 30), OP.crop_array(raw, 3, 5, 20, 30))
