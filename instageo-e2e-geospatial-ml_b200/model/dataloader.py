"""Chip dataloader's normalise-and-mask step on the GPU, behind InstaGeo's function names.

Mirrors instageo/model/dataloader.py -- ``normalize_and_convert_to_tensor`` (:495-524),
``process_and_augment`` (:527-585), ``crop_array`` (:588-615), ``process_test`` (:618-669),
``process_data`` arithmetic (:741) and the ``InstaGeoDataset`` nodata mask (:895-900) -- with
the same argument names, meaning and error behaviour.  The reference functions receive the
ALREADY multiplied float64 array; these drop-ins accept that (``IG_F64`` input) and also the
raw integer raster plus ``constant_multiplier`` (the fused, fast form).  All arithmetic runs
in ``ig_preprocess`` (sm_100a); results are CUDA tensors (use ``num_workers=0``).
"""
from __future__ import annotations

import os

from functools import partial
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import ops


def _to_device(arr, device) -> torch.Tensor:
    if isinstance(arr, torch.Tensor):
        return arr.to(device)
    arr = np.ascontiguousarray(arr)
    if arr.dtype == np.uint16:
        return torch.from_numpy(arr.view(np.int16)).to(device).view(torch.uint16)
    return torch.from_numpy(arr).to(device)


def _stack_images(ims) -> np.ndarray:
    """List of PIL images / 2-D arrays -> [T*C, H, W] float32 (what ``ToTensor().float()`` sees)."""
    if isinstance(ims, (np.ndarray, torch.Tensor)):
        return ims
    return np.stack([np.asarray(im, dtype=np.float32) for im in ims])


def crop_array(arr: np.ndarray, left: int, top: int, right: int, bottom: int) -> np.ndarray:
    """Crop a 2-D/3-D/4-D array (dataloader.py:588-615)."""
    if len(arr.shape) == 2:
        return arr[top:bottom, left:right]
    elif len(arr.shape) == 3:
        return arr[:, top:bottom, left:right]
    elif len(arr.shape) == 4:
        return arr[:, :, top:bottom, left:right]
    raise ValueError("Input array must be a 2D, 3D or 4D array")


def normalize_and_convert_to_tensor(ims, label, mean: List[float], std: List[float], temporal_size: int = 1,
                                    device="cuda") -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """[T*C,H,W] -> normalised [C,T,H,W] float32 CUDA tensor (dataloader.py:495-524)."""
    x = _to_device(_stack_images(ims), device)
    tc, h, w = x.shape
    if h != w or h % 16:
        raise ValueError(f"the CUDA preprocessing kernel needs square chips with side % 16 == 0, got {h}x{w}")
    spec = ops.PreprocessSpec(mean, std, temporal_size, device=device)
    out = ops.preprocess(x.unsqueeze(0), spec, win=h, want_f32=True)["f32"][0]
    if label is not None:
        label = torch.from_numpy(np.array(label)).squeeze()
    return out, label


def process_and_augment(x: np.ndarray, y: Optional[np.ndarray], mean: List[float], std: List[float],
                        temporal_size: int = 1, im_size: int = 224, crop: bool = True,
                        label_no_data_value: int = -1, chip_no_data_value: int = 0,
                        max_pixel_value: float = 10000.0,
                        augmentations: Optional[List[Dict[str, Any]]] = None, device="cuda"
                        ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Inference form of dataloader.py:527-585 (``augmentations`` must be None).

    ``crop=True`` draws the RandomCrop origin from torch's global RNG exactly like
    torchvision's ``RandomCrop.get_params`` (no draw when the chip already has ``im_size``).
    """
    if augmentations:
        raise NotImplementedError("training-time augmentations are outside the inference hot path")
    h, w = x.shape[-2:]
    top = left = 0
    if crop and (h, w) != (im_size, im_size):
        if h < im_size or w < im_size:
            raise ValueError(f"Required crop size {(im_size, im_size)} is larger than input image size {(h, w)}")
        top = int(torch.randint(0, h - im_size + 1, size=(1,)).item())
        left = int(torch.randint(0, w - im_size + 1, size=(1,)).item())
        size = im_size
    else:
        size = h
        if h != w:
            raise ValueError("without cropping the chip must be square")
    xd = _to_device(x, device)
    spec = ops.PreprocessSpec(mean, std, temporal_size, device=device)
    win = torch.tensor([[0, top, left]], dtype=torch.int32, device=device)
    out = ops.preprocess(xd.unsqueeze(0), spec, windows=win, win=size, want_f32=True)["f32"][0]
    label = None
    if y is not None:
        lab = np.asarray(y).astype(np.float32).squeeze()
        label = torch.from_numpy(np.ascontiguousarray(lab[top:top + size, left:left + size])).squeeze()
    return out, label


def process_test(x: np.ndarray, y: np.ndarray, mean: List[float], std: List[float], temporal_size: int = 1,
                 img_size: int = 512, crop_size: int = 224, stride: int = 224, device="cuda"
                 ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Strided crop grid, top outer / left inner (dataloader.py:618-669) -> [n,C,T,crop,crop]."""
    grid = ops.window_grid(img_size, img_size, crop_size, stride, edge=False)
    xd = _to_device(x, device)
    spec = ops.PreprocessSpec(mean, std, temporal_size, device=device)
    win = torch.tensor([[0, t, l] for t, l in grid], dtype=torch.int32, device=device)
    imgs = ops.preprocess(xd.unsqueeze(0), spec, windows=win, win=crop_size, want_f32=True)["f32"]
    lab = np.asarray(y).astype(np.float32)
    labels = torch.stack([torch.from_numpy(np.ascontiguousarray(
        crop_array(lab, l, t, l + crop_size, t + crop_size))).squeeze() for t, l in grid])
    return imgs, labels


def process_raw_chips(raw, mean: Sequence[float], std: Sequence[float], temporal_size: int = 1,
                      bands: Optional[Sequence[int]] = None, constant_multiplier: float = 1.0,
                      no_data_value: Optional[float] = None, device="cuda", spec: ops.PreprocessSpec | None = None,
                      want_f32: bool = True, want_patches: bool = False, want_mask: bool = True):
    """Fused form of ``process_data`` (:733-741) + mask (:899) + ``process_and_augment(crop=False)``
    for a batch of raw integer chips [n, bands, 224, 224]: one kernel, one pass over HBM.

    Returns dict: ``f32`` [n,C,T,H,W], ``patches`` (bf16 tubelet rows), ``mask_elem`` [n,T*C,H,W]
    (the reference's ``arr_x == no_data_value``), ``mask_px`` [n,H,W].
    """
    xd = _to_device(raw, device)
    if spec is None:
        spec = ops.PreprocessSpec(mean, std, temporal_size, bands, constant_multiplier, no_data_value, device)
    return ops.preprocess(xd, spec, win=xd.shape[-1], want_f32=want_f32, want_patches=want_patches,
                          want_mask_elem=want_mask, want_mask_px=want_mask)


def get_raster_data(fname, is_label: bool = True, bands: Optional[List[int]] = None,
                    no_data_value: Optional[int] = -9999, mask_cloud: bool = True, water_mask: bool = False,
                    device=None):
    """instageo/model/dataloader.py:672-704 for a single GeoTIFF: ``rasterio.open(fname).read()`` (all bands,
    [bands, H, W], file dtype) and the band gather for non-label rasters.  rasterio when importable, else the
    in-repo TIFF codec (``instageo_b200.data.geotiff``: strips / tiles, none / Deflate / LZW, predictor 2).  The
    multi-file dict form of the reference (``open_mf_tiff_dataset``, xarray) is out of scope.

    ``device``: return a CUDA tensor instead -- 16-bit rasters are inflated on host threads into pinned memory and
    unpacked (predictor, byte order, de-interleave) by ``ig_tiff_unpack16`` on the GPU, ready for kernel 1."""
    if isinstance(fname, dict):
        raise NotImplementedError("multi-file tile dictionaries (open_mf_tiff_dataset / xarray) are out of scope")
    if device is not None:
        from ..data.geotiff import read_geotiff_device
        data = read_geotiff_device(fname, device)[0]
        if (not is_label) and bands:
            data = data[list(bands)]
        return data
    try:
        import rasterio  # type: ignore
        with rasterio.open(fname) as src:
            data = src.read()
    except ImportError:
        from ..data.geotiff import read_geotiff
        data = read_geotiff(fname)[0]
    if (not is_label) and bands:
        data = data[bands, ...]
    return data


class InstaGeoChipDataset(torch.utils.data.Dataset):
    """Counterpart of ``InstaGeoDataset`` (dataloader.py:832-906): ``chips`` holds decoded rasters
    [bands, H, W] or GeoTIFF paths (read on access with ``get_raster_data``); ``__getitem__`` returns
    ``((tensor, label), name, nodata_mask)`` like the reference does with ``include_filenames=True``."""

    def __init__(self, chips: Sequence[np.ndarray], names: Sequence[str], preprocess_func, no_data_value,
                 constant_multiplier: float = 1.0, bands: Optional[List[int]] = None,
                 include_filenames: bool = True):
        self.chips, self.names = chips, list(names)
        self.preprocess_func = preprocess_func
        self.no_data_value = no_data_value
        self.constant_multiplier = constant_multiplier
        self.bands = bands
        self.include_filenames = include_filenames

    def __len__(self) -> int:
        return len(self.chips)

    def __getitem__(self, i: int):
        data = self.chips[i]
        if isinstance(data, (str, os.PathLike)):
            data = get_raster_data(data, is_label=False)
        if self.bands:
            data = data[self.bands, ...]
        arr_x = data * self.constant_multiplier
        if self.include_filenames:
            return self.preprocess_func(arr_x, None), self.names[i], arr_x == self.no_data_value
        return self.preprocess_func(arr_x, None)


def make_preprocess_func(mean, std, temporal_size=1, im_size=224, **kw):
    """``partial(process_and_augment, ...)`` as built in instageo/model/run.py:225-232."""
    return partial(process_and_augment, mean=mean, std=std, temporal_size=temporal_size, im_size=im_size,
                   augmentations=None, **kw)
