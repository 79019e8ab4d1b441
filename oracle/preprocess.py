"""Numpy oracle for the raster -> chip preprocessing step (TEST INFRASTRUCTURE ONLY).

Follows, step by step, what the reference does on the host for one chip:

* ``get_raster_data`` / ``process_data``   instageo/model/dataloader.py:672-750
* ``InstaGeoDataset.__getitem__`` mask      instageo/model/dataloader.py:895-900
* ``process_and_augment`` (no crop, no aug) instageo/model/dataloader.py:527-585
* ``normalize_and_convert_to_tensor``       instageo/model/dataloader.py:495-524
* ``process_test`` / ``crop_array`` grid    instageo/model/dataloader.py:588-669
* ``decode_fmask_value``                    instageo/data/hls_utils.py:77-86
* ``apply_mask`` each/any                   instageo/data/data_pipeline.py:229-267
* ``mask_segmentation_map`` each/any        instageo/data/data_pipeline.py:66-98

Closed form of the normalisation chain: the reference multiplies
the integer raster by a Python float (-> float64, dataloader.py:741), pushes each
band through ``PIL.Image.fromarray`` (float64 -> mode "F" float32, :567), then
``ToTensor().float()`` (:516) and ``transforms.Normalize`` which does an IEEE
float32 subtract followed by a true float32 divide (:519).
"""
from __future__ import annotations

import numpy as np

HLS_FMASK_POS = {"cloud": 1, "near_cloud_or_shadow": 2, "cloud_shadow": 3, "water": 5}


def scale_raw(raw: np.ndarray, bands, constant_multiplier: float) -> np.ndarray:
    """``data[bands] * constant_multiplier`` -> float64 (dataloader.py:702-703, :741)."""
    data = raw[list(bands), ...] if bands is not None else raw
    return data * float(constant_multiplier)


def nodata_mask(arr_x: np.ndarray, no_data_value) -> np.ndarray:
    """``arr_x == no_data_value`` on the already-multiplied array (dataloader.py:899).

    ``no_data_value=None`` compares against ``None`` -> all False.
    """
    if no_data_value is None:
        return np.zeros(arr_x.shape, dtype=bool)
    return arr_x == no_data_value


def normalize(arr_x: np.ndarray, mean, std, temporal_size: int = 1) -> np.ndarray:
    """[T*C,H,W] float64 -> [C,T,H,W] float32 (dataloader.py:515-521)."""
    x32 = arr_x.astype(np.float32)  # PIL mode "F" + ToTensor().float()
    tc, h, w = x32.shape
    c = tc // temporal_size
    x32 = x32.reshape(temporal_size, c, h, w)
    m = np.asarray(mean, dtype=np.float32).reshape(1, c, 1, 1)
    s = np.asarray(std, dtype=np.float32).reshape(1, c, 1, 1)
    out = (x32 - m) / s  # float32 sub, float32 true divide
    return np.ascontiguousarray(out.transpose(1, 0, 2, 3))


def preprocess_chip(raw, bands, constant_multiplier, mean, std, temporal_size, no_data_value):
    """raw integer chip [nb,H,W] -> (tensor [C,T,H,W] f32, mask [T*C,H,W] bool)."""
    arr_x = scale_raw(raw, bands, constant_multiplier)
    return normalize(arr_x, mean, std, temporal_size), nodata_mask(arr_x, no_data_value)


# ----------------------------------------------------------------------------- windows
def window_origins(size: int, crop: int, stride: int, edge: bool) -> list[int]:
    """Window origins along one axis.

    ``edge=False`` is exactly ``range(0, size-crop+1, stride)`` (dataloader.py:658-659).
    ``edge=True`` appends the edge-aligned origin ``size-crop`` when the stride does
    not land on it (OUR rule, SURVEY.md A.6 -- the reference drops the remainder).
    """
    o = list(range(0, size - crop + 1, stride))
    if edge and (size - crop) % stride != 0:
        o.append(size - crop)
    return o


def window_grid(height: int, width: int, crop: int, stride: int, edge: bool = False):
    """Row-major (top outer, left inner) list of (top, left) (dataloader.py:658-664)."""
    return [(t, l) for t in window_origins(height, crop, stride, edge)
            for l in window_origins(width, crop, stride, edge)]


def crop_array(arr: np.ndarray, left: int, top: int, right: int, bottom: int) -> np.ndarray:
    """dataloader.py:588-615."""
    if arr.ndim == 2:
        return arr[top:bottom, left:right]
    if arr.ndim == 3:
        return arr[:, top:bottom, left:right]
    if arr.ndim == 4:
        return arr[:, :, top:bottom, left:right]
    raise ValueError("Input array must be a 2D, 3D or 4D array")


def process_test(arr_x, mean, std, temporal_size=1, img_size=512, crop_size=224, stride=224):
    """[T*C,S,S] float64 -> [n,C,T,crop,crop] f32 (dataloader.py:618-669, image part)."""
    outs = []
    for top, left in window_grid(img_size, img_size, crop_size, stride, edge=False):
        outs.append(normalize(crop_array(arr_x, left, top, left + crop_size, top + crop_size),
                              mean, std, temporal_size))
    return np.stack(outs)


# ----------------------------------------------------------------------------- cloud masks
def decode_fmask_value(value, position: int):
    """hls_utils.py:77-86 -- bit ``position`` by floor-division arithmetic."""
    quotient = value // (2 ** position)
    return quotient - ((quotient // 2) * 2)


def apply_fmask(chip: np.ndarray, fmask: np.ndarray, no_data_value, strategy="each",
                mask_types=tuple(HLS_FMASK_POS.keys())) -> np.ndarray:
    """chip [T*C,H,W], fmask [T,H,W] -> chip with masked pixels = no_data_value.

    data_pipeline.py:229-267.  NB the reference tests ``if pos:`` so a position of 0
    would be skipped; none of the HLS positions is 0.
    """
    chip = chip.copy()
    for name in mask_types:
        pos = HLS_FMASK_POS.get(name)
        if not pos:
            continue
        dec = decode_fmask_value(fmask.astype(np.int64), pos)
        if strategy == "each":
            dec = dec.repeat(chip.shape[0] // fmask.shape[0], axis=0)
        elif strategy == "any":
            dec = dec.any(axis=0)
        else:
            raise ValueError(strategy)
        chip = np.where(dec == 0, chip, no_data_value)
    return chip


def mask_segmentation_map(chip: np.ndarray, seg_map: np.ndarray, chip_no_data_value,
                          strategy="any", seg_no_data_value=-1) -> np.ndarray:
    """data_pipeline.py:66-98 (``NoDataValues.SEG_MAP`` = -1)."""
    if strategy == "each":
        valid = (chip != chip_no_data_value).any(axis=0)
    elif strategy == "any":
        valid = (chip != chip_no_data_value).all(axis=0)
    else:
        raise ValueError(strategy)
    return np.where(valid, seg_map, seg_no_data_value)


def create_chip(chip: np.ndarray, fmask, seg_map=None, strategy="each", mask_types=tuple(HLS_FMASK_POS.keys()),
                no_data_value=0, clip=(0, 10000), seg_no_data_value=-1):
    """The array arithmetic of ``HLSRasterPipeline`` chip creation, hls_utils.py:359-403, on plain arrays:
    apply_mask (fill 0) -> clip(0, 10000) -> "all cloud?" count -> mask_segmentation_map on the clipped chip
    -> "empty label?" count -> uint16 chip / int8 label map.
    Returns (chip uint16, seg int8 or None, n_valid_chip_elements, n_valid_label_pixels or None)."""
    out = chip.astype(np.int64)
    if fmask is not None:
        out = apply_fmask(out, fmask, no_data_value, strategy, mask_types)
    if clip is not None:
        out = np.clip(out, clip[0], clip[1])
    n_valid = int((out != no_data_value).sum())
    if seg_map is None:
        return out.astype(np.uint16), None, n_valid, None
    seg = mask_segmentation_map(out, seg_map, no_data_value, strategy, seg_no_data_value)
    return out.astype(np.uint16), seg.astype(np.int8), n_valid, int((seg != seg_no_data_value).sum())


# ----------------------------------------------------------------------------- synthetic inputs
def synth_chips(n: int, temporal: int, seed: int = 1042, size: int = 224, nodata=-9999,
                dtype=np.int16, n_src_bands: int | None = None) -> np.ndarray:
    """Seeded synthetic raw chips [n, nb, size, size] (SURVEY.md §8d value law).

    Reflectance DN uniform in [0, 10000]; ~2% of pixels of a timestep set to nodata in
    all 6 bands of that timestep, plus a 10x10 all-band nodata block.
    """
    rng = np.random.default_rng(seed)
    nb = n_src_bands if n_src_bands is not None else temporal * 6
    raw = rng.integers(0, 10001, size=(n, nb, size, size), dtype=np.int64)
    if nodata is not None:
        hole = rng.random((n, temporal, size, size)) < 0.02
        for t in range(temporal):
            lo, hi = t * 6, min(nb, t * 6 + 6)
            raw[:, lo:hi][np.broadcast_to(hole[:, t:t + 1], raw[:, lo:hi].shape)] = nodata
        raw[:, :, 5:15, 7:17] = nodata
    return raw.astype(dtype)
