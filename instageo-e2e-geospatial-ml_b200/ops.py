"""Thin torch-tensor wrappers over the C ABI kernels (tensor ownership only; no math here)."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib

_RAW_DTYPES = {torch.int16: _lib.IG_I16, torch.uint16: _lib.IG_U16, torch.float32: _lib.IG_F32,
               torch.float64: _lib.IG_F64}


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: instageo_b200 has no CPU path")


def window_origins(size: int, crop: int, stride: int, edge: bool) -> list[int]:
    """``range(0, size-crop+1, stride)`` (instageo/model/dataloader.py:658-659) plus, with
    ``edge=True``, the edge-aligned origin ``size-crop``."""
    o = list(range(0, size - crop + 1, stride))
    if edge and (size - crop) % stride != 0:
        o.append(size - crop)
    return o


def window_grid(height: int, width: int, crop: int, stride: int, edge: bool = False):
    """Row-major (top outer, left inner) window list, like ``process_test``."""
    return [(t, l) for t in window_origins(height, crop, stride, edge)
            for l in window_origins(width, crop, stride, edge)]


class PreprocessSpec:
    """Per-dataset constants of the normalise-and-mask step, resident on the device.

    Mirrors ``cfg.dataloader.{bands, mean, std, no_data_value, constant_multiplier,
    temporal_dim}`` (instageo/model/configs/config.yaml:34-60).
    """

    def __init__(self, mean: Sequence[float], std: Sequence[float], temporal_size: int = 1,
                 bands: Optional[Sequence[int]] = None, constant_multiplier: float = 1.0,
                 no_data_value: Optional[float] = None, device="cuda"):
        self.C = len(mean)
        self.T = int(temporal_size)
        if len(std) != self.C:
            raise ValueError("mean and std must have the same length")
        tc = self.T * self.C
        bands = list(range(tc)) if bands is None else list(bands)
        if len(bands) != tc:
            raise ValueError(f"bands must list temporal_size*len(mean) = {tc} source bands, got {len(bands)}")
        self.bands = bands
        self.cm = float(constant_multiplier)
        self.no_data_value = no_data_value
        self.device = torch.device(device)
        # transforms.Normalize builds float32 tensors from the Python floats (dataloader.py:515)
        self.mean = torch.tensor(np.asarray(mean, dtype=np.float32), device=self.device)
        self.std = torch.tensor(np.asarray(std, dtype=np.float32), device=self.device)
        self.band_idx = torch.tensor(bands, dtype=torch.int32, device=self.device)


def preprocess(raw: torch.Tensor, spec: PreprocessSpec, windows: Optional[torch.Tensor] = None, win: int = 224,
               want_f32: bool = True, want_patches: bool = False, want_mask_elem: bool = False,
               want_mask_px: bool = False, fmask: Optional[torch.Tensor] = None, fmask_bits: int = 0,
               masking_strategy: str = "each", out_patches: Optional[torch.Tensor] = None):
    """Kernel 1.  raw [n_img, nb, H, W] (int16 | uint16 | f32 | f64, CUDA, last dim contiguous).

    windows: int32 CUDA [n_win, 3] rows (image, top, left) or None for whole ``win x win`` images.
    Returns dict with the requested outputs: ``f32`` [n,C,T,win,win], ``patches``
    [n*T*(win/16)^2, C*256] bf16, ``mask_elem`` [n,T*C,win,win] bool, ``mask_px`` [n,win,win] bool.
    """
    lib = _lib.load()
    _require_cuda(raw, "raw")
    if raw.dim() == 3:
        raw = raw.unsqueeze(0)
    if raw.dim() != 4 or raw.stride(3) != 1:
        raise ValueError("raw must be [n_img, bands, H, W] with a contiguous last dimension")
    if raw.dtype not in _RAW_DTYPES:
        raise TypeError(f"unsupported raw dtype {raw.dtype}")
    n_img, nb, H, W = raw.shape
    if max(spec.bands) >= nb or min(spec.bands) < 0:
        raise IndexError(f"band index out of range for a raster with {nb} bands")
    if windows is not None:
        _require_cuda(windows, "windows")
        windows = windows.to(torch.int32).contiguous()
        n_win = windows.shape[0]
    else:
        n_win = n_img
    dev = raw.device
    C_, T = spec.C, spec.T
    out = {}
    f32 = torch.empty((n_win, C_, T, win, win), dtype=torch.float32, device=dev) if want_f32 else None
    patches = None
    if want_patches:
        rows = n_win * T * (win // 16) ** 2
        if out_patches is not None:   # caller-owned tubelet-row buffer (a stable address keeps CUDA graphs unpatched)
            if out_patches.dtype != torch.bfloat16 or not out_patches.is_contiguous() or out_patches.numel() < rows * C_ * 256:
                raise ValueError("out_patches must be a contiguous bf16 buffer of at least n_win*T*(win/16)^2 x C*256")
            patches = out_patches.view(-1)[: rows * C_ * 256].view(rows, C_ * 256)
        else:
            patches = torch.empty((rows, C_ * 256), dtype=torch.bfloat16, device=dev)
    m_el = torch.empty((n_win, T * C_, win, win), dtype=torch.uint8, device=dev) if want_mask_elem else None
    m_px = torch.empty((n_win, win, win), dtype=torch.uint8, device=dev) if want_mask_px else None
    if fmask is not None:
        _require_cuda(fmask, "fmask")
        fmask = fmask.to(torch.uint8).contiguous()
        if tuple(fmask.shape) != (n_img, T, H, W):
            raise ValueError(f"fmask must be [n_img, T, H, W] = {(n_img, T, H, W)}, got {tuple(fmask.shape)}")
    has_nd = spec.no_data_value is not None
    with _lib.on_device(raw):
        _lib.check(lib.ig_preprocess(
            raw.data_ptr(), _RAW_DTYPES[raw.dtype], n_img, nb, H, W, raw.stride(0), raw.stride(1), raw.stride(2),
            spec.band_idx.data_ptr(), T, C_, _lib.ptr(windows), n_win, win, spec.cm, spec.mean.data_ptr(),
            spec.std.data_ptr(), int(has_nd), float(spec.no_data_value) if has_nd else 0.0, _lib.ptr(fmask),
            int(fmask_bits) if fmask is not None else 0,
            _lib.IG_MASK_ANY if masking_strategy == "any" else _lib.IG_MASK_EACH, _lib.ptr(f32), _lib.ptr(patches),
            _lib.ptr(m_el), _lib.ptr(m_px), _lib.current_stream(dev)))
    if f32 is not None:
        out["f32"] = f32
    if patches is not None:
        out["patches"] = patches
    if m_el is not None:
        out["mask_elem"] = m_el.view(torch.bool)
    if m_px is not None:
        out["mask_px"] = m_px.view(torch.bool)
    return out


def nodata_map(raw: torch.Tensor, spec: PreprocessSpec, y0: int = 0, y1: Optional[int] = None,
               fmask: Optional[torch.Tensor] = None, fmask_bits: int = 0, masking_strategy: str = "each",
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[y1-y0, W] bool: tile pixel is nodata in ANY selected band / timestep (the element test of kernel 1,
    dataloader.py:741, :899), for rows [y0, y1) of one raster ``raw`` [bands, H, W].  ``fmask`` [T, H, W] uint8."""
    lib = _lib.load()
    _require_cuda(raw, "raw")
    if raw.dim() != 3 or raw.stride(2) != 1:
        raise ValueError("raw must be [bands, H, W] with a contiguous last dimension")
    if raw.dtype not in _RAW_DTYPES:
        raise TypeError(f"unsupported raw dtype {raw.dtype}")
    nb, H, W = raw.shape
    y1 = H if y1 is None else y1
    if max(spec.bands) >= nb or min(spec.bands) < 0:
        raise IndexError(f"band index out of range for a raster with {nb} bands")
    if out is None:
        out = torch.empty((y1 - y0, W), dtype=torch.uint8, device=raw.device)
    elif out.dtype != torch.uint8 or not out.is_contiguous() or tuple(out.shape) != (y1 - y0, W):
        raise ValueError("out must be a contiguous uint8 [y1-y0, W] tensor")
    if fmask is not None:
        _require_cuda(fmask, "fmask")
        fmask = fmask.to(torch.uint8).contiguous()
        if tuple(fmask.shape) != (spec.T, H, W):
            raise ValueError(f"fmask must be [T, H, W] = {(spec.T, H, W)}, got {tuple(fmask.shape)}")
    has_nd = spec.no_data_value is not None
    _lib.call("ig_nodata_map", raw.device, raw.data_ptr(), _RAW_DTYPES[raw.dtype], nb, H, W, raw.stride(0),
              raw.stride(1), spec.band_idx.data_ptr(), spec.T, spec.C, spec.cm, int(has_nd),
              float(spec.no_data_value) if has_nd else 0.0, _lib.ptr(fmask),
              int(fmask_bits) if fmask is not None else 0,
              _lib.IG_MASK_ANY if masking_strategy == "any" else _lib.IG_MASK_EACH, y0, y1, out.data_ptr())
    return out.view(torch.bool)


def stitch(win_logits: torch.Tensor, ys: Sequence[int], xs: Sequence[int], height: int, width: int,
           y0: int = 0, y1: Optional[int] = None, win_base: int = 0, nodata_px: Optional[torch.Tensor] = None,
           nodata_class: int = -1, want_avg: bool = False, want_hist: bool = False,
           origins: Optional[tuple] = None, out: Optional[torch.Tensor] = None):
    """Kernel 5.  win_logits [n_win, nc, win, win] f32 CUDA holding windows
    ``win_base .. win_base+n_win`` of the row-major (ys x xs) grid.  Returns dict with
    ``class_map`` int8 [y1-y0, W] and optionally ``avg`` f32 [nc, y1-y0, W], ``hist`` int64 [nc+1].
    ``nodata_px`` [y1-y0, W] (the stripe's own rows; [H, W] is accepted and sliced).  ``origins`` = (ys, xs) as int32
    CUDA tensors and ``out`` = the int8 result buffer, when the caller keeps them across tiles."""
    lib = _lib.load()
    _require_cuda(win_logits, "win_logits")
    win_logits = win_logits.contiguous()
    if win_logits.dtype != torch.float32 or win_logits.dim() != 4:
        raise TypeError("win_logits must be float32 [n_win, nc, win, win]")
    n_win, nc, win, _ = win_logits.shape
    y1 = height if y1 is None else y1
    dev = win_logits.device
    if origins is not None:
        ys_t, xs_t = origins
    else:
        ys_t = torch.tensor(list(ys), dtype=torch.int32, device=dev)
        xs_t = torch.tensor(list(xs), dtype=torch.int32, device=dev)
    if out is not None:
        if out.dtype != torch.int8 or not out.is_contiguous() or tuple(out.shape) != (y1 - y0, width):
            raise ValueError("out must be a contiguous int8 [y1-y0, W] tensor")
        cls = out
    else:
        cls = torch.empty((y1 - y0, width), dtype=torch.int8, device=dev)
    avg = torch.empty((nc, y1 - y0, width), dtype=torch.float32, device=dev) if want_avg else None
    hist = torch.zeros((nc + 1,), dtype=torch.int64, device=dev) if want_hist else None
    if nodata_px is not None:
        _require_cuda(nodata_px, "nodata_px")
        if tuple(nodata_px.shape) == (height, width) and (y0, y1) != (0, height):
            nodata_px = nodata_px[y0:y1]
        nodata_px = nodata_px.contiguous().view(torch.uint8)
        if tuple(nodata_px.shape) != (y1 - y0, width):
            raise ValueError("nodata_px must be [y1-y0, W] (or the whole [H, W] map)")
    with _lib.on_device(win_logits):
        _lib.check(lib.ig_stitch(win_logits.data_ptr(), n_win, win_base, nc, win, ys_t.data_ptr(), len(ys),
                                 xs_t.data_ptr(), len(xs), height, width, y0, y1, _lib.ptr(nodata_px), nodata_class,
                                 _lib.ptr(avg), cls.data_ptr(), _lib.ptr(hist), _lib.current_stream(dev)))
    out = {"class_map": cls}
    if avg is not None:
        out["avg"] = avg
    if hist is not None:
        out["hist"] = hist
    return out


def linear(a: torch.Tensor, w: torch.Tensor, bias=None, resid=None, act: int = 0, out_dtype=torch.bfloat16):
    """tcgen05 GEMM building block: act(a @ w.T + bias) (+ resid).  a [M,K], w [N,K] bf16."""
    lib = _lib.load()
    _require_cuda(a, "a")
    M, K = a.shape
    N = w.shape[0]
    out = resid if resid is not None else torch.empty((M, N), dtype=out_dtype, device=a.device)
    with _lib.on_device(a):
        _lib.check(lib.ig_linear(a.data_ptr(), w.data_ptr(), _lib.ptr(bias), _lib.ptr(resid), out.data_ptr(),
                                 _lib.IG_BF16 if out.dtype == torch.bfloat16 else _lib.IG_F32, M, N, K, act,
                                 _lib.current_stream(a.device)))
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    _require_cuda(x, "x")
    M, D = x.shape
    out = torch.empty((M, D), dtype=torch.bfloat16, device=x.device)
    with _lib.on_device(x):
        _lib.check(lib.ig_layernorm(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), out.data_ptr(), M, D,
                                    _lib.current_stream(x.device)))
    return out


def attention(qkv: torch.Tensor, batch: int, ntok: int, heads: int) -> torch.Tensor:
    lib = _lib.load()
    _require_cuda(qkv, "qkv")
    out = torch.empty((batch * ntok, heads * 64), dtype=torch.bfloat16, device=qkv.device)
    with _lib.on_device(qkv):
        _lib.check(lib.ig_attention(qkv.data_ptr(), out.data_ptr(), batch, ntok, heads,
                                    _lib.current_stream(qkv.device)))
    return out
