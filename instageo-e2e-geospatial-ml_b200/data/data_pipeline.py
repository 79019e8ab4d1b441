"""Chip-creation masking on the device.

Mirrors the array semantics of ``instageo/data/data_pipeline.py`` (``apply_mask`` :229-267,
``mask_segmentation_map`` :66-98), ``instageo/data/hls_utils.py`` (``decode_fmask_value`` :77-86 and the
clip / cast / "anything left?" chain of ``HLSRasterPipeline`` :359-403) on plain arrays instead of
``xarray.DataArray`` (rioxarray / GDAL IO is out of scope): inputs are CUDA tensors or NumPy arrays of
a whole chip OR a whole tile -- the kernel is pixel-wise, so a 3660 x 3660 x 18-band HLS tile is one launch.
One pass of ``csrc/chipmask.cu``; there is no CPU path.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from .. import _lib

# instageo/data/settings.py (MASK_DECODING_POS["HLS"]): Fmask bit positions per mask type
MASK_DECODING_POS = {"HLS": {"cloud": 1, "near_cloud_or_shadow": 2, "cloud_shadow": 3, "water": 5}}
_CHIP_DTYPES = {torch.int16: _lib.IG_I16, torch.uint16: _lib.IG_U16}


def _dev(a, device=None) -> torch.Tensor:
    t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
    if not t.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("chip masking runs on the GPU: instageo_b200 has no CPU path")
        t = t.to(device or "cuda")
    return t.contiguous()


def _bits(data_source: str, mask_types: Sequence[str]) -> int:
    bits = 0
    for name in mask_types:
        pos = MASK_DECODING_POS[data_source].get(name, None)
        if pos:  # the reference skips unknown types AND position 0 (``if pos:``, data_pipeline.py:254)
            bits |= 1 << pos
    return bits


def _strategy(name: str) -> int:
    if name == "each":
        return _lib.IG_MASK_EACH
    if name == "any":
        return _lib.IG_MASK_ANY
    raise ValueError(f"Invalid masking strategy: {name}")


def decode_fmask_value(value: torch.Tensor, position: int) -> torch.Tensor:
    """Bit ``position`` of an HLS v2.0 Fmask byte (hls_utils.py:77-86); integer tensor arithmetic, any device."""
    quotient = torch.div(value, 2 ** position, rounding_mode="floor")
    return quotient - torch.div(quotient, 2, rounding_mode="floor") * 2


def create_chip(chip, fmask=None, seg_map=None, masking_strategy: str = "each",
                mask_types: Sequence[str] = tuple(MASK_DECODING_POS["HLS"].keys()), data_source: str = "HLS",
                no_data_value: int = 0, clip: Optional[tuple] = (0, 10000), seg_no_data_value: int = -1):
    """Fused form of hls_utils.py:359-403.  chip [T*C, H, W] int16|uint16, fmask [T, H, W] uint8 or None,
    seg_map [H, W] integer or None.  Returns (chip uint16 (input dtype when ``clip`` is None), seg int8 or
    None, counts int64 CUDA tensor [2] = (chip elements != no_data_value, label pixels != seg_no_data_value))
    -- the two counts the reference tests against 0 to skip all-cloud chips and empty labels."""
    x = _dev(chip)
    if x.dim() != 3 or x.dtype not in _CHIP_DTYPES:
        raise TypeError("chip must be int16 or uint16 [bands, H, W]")
    nb, H, W = x.shape
    fm, steps, bits = None, 1, 0
    if fmask is not None:
        fm = _dev(fmask, x.device)
        if fm.dtype != torch.uint8:
            fm = fm.to(torch.uint8)
        if fm.dim() != 3 or tuple(fm.shape[1:]) != (H, W) or nb % fm.shape[0]:
            raise ValueError(f"fmask must be [T, {H}, {W}] with T dividing {nb} bands, got {tuple(fm.shape)}")
        steps, bits = fm.shape[0], _bits(data_source, mask_types)
    seg_in = seg_out = None
    if seg_map is not None:
        seg_in = _dev(seg_map, x.device).reshape(H, W)
        # The reference keeps the label map's dtype (float32 for regression targets, hls_utils.py:398); the device
        # kernel carries int8 class labels only.  Refuse anything a cast to int8 would change instead of truncating.
        if seg_in.is_floating_point() or seg_in.dtype == torch.bool:
            raise TypeError("create_chip: seg_map must hold integer class labels (float regression targets are not "
                            "supported by the device path; mask them with the reference's xarray chain)")
        if seg_in.dtype != torch.int8:
            if seg_in.numel() and (int(seg_in.max()) > 127 or int(seg_in.min()) < -128):
                raise ValueError("create_chip: seg_map labels must fit int8 ([-128, 127])")
            seg_in = seg_in.to(torch.int8)
        seg_in = seg_in.contiguous()
        seg_out = torch.empty_like(seg_in)
    out = torch.empty((nb, H, W), dtype=torch.uint16 if clip is not None else x.dtype, device=x.device)
    counts = torch.zeros(2, dtype=torch.int64, device=x.device)
    lo, hi = (int(clip[0]), int(clip[1])) if clip is not None else (1, 0)
    st = _strategy(masking_strategy)
    _lib.call("ig_chip_mask", x.device,
              x.data_ptr(), _CHIP_DTYPES[x.dtype], nb, H, W, _lib.ptr(fm), steps, bits, st, int(no_data_value), lo, hi,
              out.data_ptr(), _lib.ptr(seg_in), st, int(seg_no_data_value), _lib.ptr(seg_out), counts.data_ptr())
    return out, seg_out, counts


def apply_mask(chip, mask, no_data_value: int, mask_decoder=None, data_source: str = "HLS",
               masking_strategy: str = "each",
               mask_types: Sequence[str] = tuple(MASK_DECODING_POS["HLS"].keys())):
    """data_pipeline.py:229-267 on arrays: pixels whose Fmask has any of the ``mask_types`` bits set become
    ``no_data_value`` (per timestep for "each", for every band if any timestep is hit for "any").
    ``mask_decoder`` is accepted for signature parity; the device kernel always decodes HLS Fmask bits."""
    return create_chip(chip, mask, None, masking_strategy, mask_types, data_source, no_data_value, clip=None)[0]


def mask_segmentation_map(chip, seg_map, chip_no_data_value: int, masking_strategy: str = "any"):
    """data_pipeline.py:66-98 on arrays: label pixels whose chip is no-data in every band ("each") / in at
    least one band ("any") become ``NO_DATA_VALUES.SEG_MAP`` = -1."""
    return create_chip(chip, None, seg_map, masking_strategy, (), "HLS", chip_no_data_value, clip=None)[1]
