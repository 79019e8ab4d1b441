"""Numpy oracle for the sliding-window overlap-averaging stitch (TEST INFRASTRUCTURE ONLY).

PARITY UNPINNED: the reference snapshot has no stitch.  What exists is
the strided crop grid ``process_test`` (instageo/model/dataloader.py:618-669) -- reused
for the window order -- the per-chip ``argmax -> int8`` (instageo/model/infer_utils.py:
96-101) and the nodata comparison (instageo/model/dataloader.py:899).  The averaging rule
below is OUR specification, frozen here before any kernel.
"""
from __future__ import annotations

import numpy as np

from .preprocess import window_grid


def stitch(win_logits: np.ndarray, origins, height: int, width: int,
           nodata_px: np.ndarray | None = None, nodata_class: int = -1):
    """Accumulate window logits in window order, divide by the cover count, argmax.

    win_logits [n, nc, win, win] float32; origins = [(top, left)] in row-major order.
    Returns (avg [nc,H,W] f32, class_map [H,W] int8).  Uncovered pixels (cnt == 0) get
    avg 0 and ``nodata_class``.
    """
    n, nc, win, _ = win_logits.shape
    acc = np.zeros((nc, height, width), dtype=np.float32)
    cnt = np.zeros((height, width), dtype=np.float32)
    for i, (t, l) in enumerate(origins):
        acc[:, t:t + win, l:l + win] += win_logits[i]  # float32 adds, window order
        cnt[t:t + win, l:l + win] += 1.0
    covered = cnt > 0
    avg = np.where(covered, acc / np.where(covered, cnt, 1.0), 0.0).astype(np.float32)
    cls = np.argmax(avg, axis=0).astype(np.int8)  # first max wins, like torch.argmax
    cls[~covered] = nodata_class
    if nodata_px is not None:
        cls[nodata_px] = nodata_class
    return avg, cls


def tile_nodata_px(raw_tile: np.ndarray, bands, constant_multiplier, no_data_value):
    """Pixel is nodata when ANY selected band/timestep equals no_data_value (A.6)."""
    if no_data_value is None:
        return np.zeros(raw_tile.shape[1:], dtype=bool)
    data = raw_tile[list(bands), ...] if bands is not None else raw_tile
    return ((data * float(constant_multiplier)) == no_data_value).any(axis=0)


def tile_windows(height, width, win, stride):
    """Windows of the tile path: reference order + edge-aligned extras."""
    return window_grid(height, width, win, stride, edge=True)
