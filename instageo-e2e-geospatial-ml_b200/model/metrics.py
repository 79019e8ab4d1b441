"""Device-resident mirrors of the reference's streaming metrics.

Same class names, constructor arguments, method names and result dictionaries as
``instageo/model/metrics.py`` (RunningConfusionMatrix :63-172, RunningAUC :180-281,
RunningRegressionMetrics :289-433), but ``update`` accumulates into uint64 / float64 counters that
live on the GPU (kernels in ``csrc/metrics.cu``): the per-step D2H copies and NumPy passes of
``PrithviSegmentationModule._shared_step`` (instageo/model/segmentation.py:117-156) disappear, and
only k*k + 2*k*n_bins integers are read back when a metric is asked for.

``update`` accepts CUDA tensors (no copy) or NumPy arrays / CPU tensors (uploaded).  There is no CPU
path: without the CUDA library every ``update`` raises.  ``segmentation_eval_update`` is the fused
form of the whole eval step: logits + labels in, confusion matrix and ROC histograms updated by one
kernel (argmax and softmax never materialised).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .. import _lib

__all__ = ["RunningConfusionMatrix", "RunningAUC", "RunningRegressionMetrics", "segmentation_eval_update",
           "confusion_summary", "auc_from_histograms"]

_LABEL_DTYPES = {torch.int64: _lib.IG_I64, torch.int32: _lib.IG_I32, torch.uint8: _lib.IG_U8, torch.int8: _lib.IG_I8}


def _device(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("instageo_b200 metrics accumulate on the GPU: no CUDA device (there is no CPU path)")
    return torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())


def _to_dev(a, device: torch.device) -> torch.Tensor:
    t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
    return t.to(device, non_blocking=True)


def _labels(a, device) -> torch.Tensor:
    t = _to_dev(a, device).reshape(-1)
    if t.dtype not in _LABEL_DTYPES:
        if t.dtype.is_floating_point or t.dtype == torch.bool:
            raise TypeError(f"labels must be an integer array, got {t.dtype}")
        t = t.long()
    return t.contiguous()


def _safe_div(num: np.ndarray, den: np.ndarray) -> np.ndarray:
    den = np.asarray(den, dtype=float)
    out = np.zeros_like(den)
    np.divide(num, den, out=out, where=den != 0)
    return out


def confusion_summary(mat: np.ndarray, total: int) -> dict:
    """Everything the reference derives from a confusion matrix (metrics.py:110-166), from the k x k integers read
    back from the device: per-class vectors under ``*_per_class`` and their macro means, plus accuracy."""
    hits = np.diag(mat)
    predicted, actual = mat.sum(axis=0), mat.sum(axis=1)            # column sums = tp + fp, row sums = tp + fn
    per = {"precision": _safe_div(hits, predicted), "recall": _safe_div(hits, actual),
           "jaccard": _safe_div(hits, predicted + actual - hits)}
    per["f1"] = _safe_div(2 * per["precision"] * per["recall"], per["precision"] + per["recall"])
    out = {"accuracy": float("nan") if total == 0 else hits.sum() / total}
    for name in ("precision", "recall", "f1", "jaccard"):
        out[name] = per[name].mean()
        out[name + "_per_class"] = per[name]
    return out


def auc_from_histograms(pos: np.ndarray, neg: np.ndarray) -> np.ndarray:
    """Per-class one-vs-rest ROC-AUC from the score histograms (metrics.py:246-259): a positive outranks every
    negative in a lower bin and ties half of those in its own bin.  Accumulated bin by bin in float, like the
    reference's loop; NaN for a class without positives or without negatives."""
    aucs = []
    for p_row, n_row in zip(pos.tolist(), neg.tolist()):
        n_pos, n_neg = sum(p_row), sum(n_row)
        if n_pos == 0 or n_neg == 0:
            aucs.append(float("nan"))
            continue
        area, below = 0.0, 0
        for p_cnt, n_cnt in zip(p_row, n_row):
            area += p_cnt * below
            area += 0.5 * p_cnt * n_cnt
            below += n_cnt
        aucs.append(area / (n_pos * n_neg))
    return np.array(aucs)


class RunningConfusionMatrix:
    """Streaming confusion matrix for single-label classification (reference: metrics.py:63-172)."""

    def __init__(self, num_classes: int, ignore_index: Optional[int] = None, device=None) -> None:
        self.num_classes = int(num_classes)
        self.ignore_index = ignore_index
        self.device = _device(device)
        self._mat = torch.zeros(self.num_classes * self.num_classes, dtype=torch.int64, device=self.device)
        self._cnt = torch.zeros(2, dtype=torch.int64, device=self.device)  # valid samples, out-of-range samples

    # -- accumulation -------------------------------------------------------------------------
    def update(self, y_true, y_pred) -> None:
        """y_true: integer labels, y_pred: predicted classes (any shape, same number of elements)."""
        t = _labels(y_true, self.device)
        p = _to_dev(y_pred, self.device).reshape(-1)
        if t.shape != p.shape:
            raise ValueError("y_true and y_pred shapes differ.")
        if p.dtype != torch.int8:
            # No host round trip for the range check (the reference raises inside update; here an out-of-range
            # prediction stays out of range after the clamp -- num_classes <= 127 -- and lands in the kernel's
            # out-of-range counter, which _sync() turns into the same ValueError at the first read-back).
            if self.num_classes > 127:
                raise ValueError("non-int8 predictions need num_classes <= 127")
            p = p.clamp(-1, 127).to(torch.int8)
        p = p.contiguous()
        has_ign = self.ignore_index is not None
        _lib.call("ig_confusion_update", self.device,
                  p.data_ptr(), t.data_ptr(), _LABEL_DTYPES[t.dtype], t.numel(), self.num_classes, int(has_ign),
                  int(self.ignore_index) if has_ign else 0, self._mat.data_ptr(), self._cnt.data_ptr())

    def _sync(self):
        cnt = self._cnt.cpu().numpy()
        if cnt[1] != 0:
            # the reference fails inside update (np.bincount on a negative index / reshape of a longer
            # bincount, metrics.py:103-105); the device path reports it at the first read-back
            raise ValueError(f"{int(cnt[1])} samples had a label or prediction outside [0, {self.num_classes})")
        return self._mat.cpu().numpy().reshape(self.num_classes, self.num_classes), int(cnt[0])

    @property
    def matrix(self) -> np.ndarray:
        return self._sync()[0]

    @property
    def total(self) -> int:
        return self._sync()[1]

    # -- derived metrics (host arithmetic on k x k integers) --------------------------------------
    def _summary(self) -> dict:
        return confusion_summary(*self._sync())

    def accuracy(self) -> float:
        return self._summary()["accuracy"]

    def precision(self) -> np.ndarray:
        return self._summary()["precision_per_class"]

    def recall(self) -> np.ndarray:
        return self._summary()["recall_per_class"]

    def f1(self) -> np.ndarray:
        return self._summary()["f1_per_class"]

    def jaccard(self) -> np.ndarray:
        return self._summary()["jaccard_per_class"]

    def compute(self, include_per_class: bool = True) -> dict:
        full = self._summary()
        out = {k: full[k] for k in ("accuracy", "precision", "recall", "f1", "jaccard")}
        if include_per_class:
            out.update({k: v.tolist() for k, v in full.items() if k.endswith("_per_class")})
        return out

    def reset(self) -> None:
        self._mat.zero_()
        self._cnt.zero_()


class RunningAUC:
    """Histogram-based streaming one-vs-rest ROC-AUC (reference: metrics.py:180-281)."""

    def __init__(self, num_classes: int, n_bins: int = 1024, min_score: float = 0.0, max_score: float = 1.0,
                 device=None) -> None:
        self.num_classes = int(num_classes)
        self.n_bins = int(n_bins)
        self.min_score = min_score
        self.max_score = max_score
        self.device = _device(device)
        self._pos = torch.zeros((self.num_classes, self.n_bins), dtype=torch.int64, device=self.device)
        self._neg = torch.zeros_like(self._pos)

    def update(self, y_true, y_score) -> None:
        """y_score: probabilities [n, num_classes] (float32 or float64), or [n] positive-class
        probabilities when num_classes == 2."""
        t = _labels(y_true, self.device)
        s = _to_dev(y_score, self.device)
        if s.dtype not in (torch.float32, torch.float64):
            s = s.double()
        if s.dim() == 1:
            if self.num_classes != 2:
                raise ValueError("For 1-D y_score, num_classes must be 2.")
            s = torch.stack([1 - s, s], dim=1)
        if t.shape[0] != s.shape[0]:
            raise ValueError("y_true and y_score length mismatch.")
        if s.dim() != 2 or s.shape[1] != self.num_classes:
            raise ValueError("Second dim of y_score must equal num_classes.")
        s = s.contiguous()
        _lib.call("ig_auc_update", self.device,
                  s.data_ptr(), _lib.IG_F32 if s.dtype == torch.float32 else _lib.IG_F64, s.shape[0], self.num_classes,
                  t.data_ptr(), _LABEL_DTYPES[t.dtype], self.n_bins, float(self.min_score), float(self.max_score),
                  self._pos.data_ptr(), self._neg.data_ptr())

    @property
    def pos_hist(self) -> np.ndarray:
        return self._pos.cpu().numpy()

    @property
    def neg_hist(self) -> np.ndarray:
        return self._neg.cpu().numpy()

    @property
    def n_pos(self) -> np.ndarray:
        return self.pos_hist.sum(axis=1)

    @property
    def n_neg(self) -> np.ndarray:
        return self.neg_hist.sum(axis=1)

    def _auc_one_class(self, c: int) -> float:
        return float(auc_from_histograms(self.pos_hist[c:c + 1], self.neg_hist[c:c + 1])[0])

    def score(self, include_per_class: bool = True) -> dict:
        per_class = auc_from_histograms(self.pos_hist, self.neg_hist)
        macro = np.nanmean(per_class)
        if include_per_class:
            return {"roc_auc_macro": macro, "roc_auc_per_class": per_class.tolist()}
        return {"roc_auc_macro": macro}

    def reset(self) -> None:
        self._pos.zero_()
        self._neg.zero_()


def segmentation_eval_update(logits: torch.Tensor, labels: torch.Tensor, confusion: Optional[RunningConfusionMatrix],
                             auc: Optional[RunningAUC] = None, ignore_index: Optional[int] = -100) -> None:
    """The metric half of ``_shared_step`` (segmentation.py:117-156) as ONE kernel: first-max argmax and
    float32 softmax of ``logits`` [B, nc, H, W], pixels whose label equals ``ignore_index`` dropped,
    ``confusion`` and (test mode) ``auc`` updated in place on the device."""
    if not logits.is_cuda:
        raise RuntimeError("logits must be a CUDA tensor: instageo_b200 has no CPU path")
    if logits.dtype != torch.float32 or logits.dim() != 4:
        raise TypeError("logits must be float32 [B, nc, H, W]")
    B, nc, H, W = logits.shape
    logits = logits.contiguous()
    lab = _labels(labels, logits.device)
    if lab.numel() != B * H * W:
        raise ValueError("labels must have one entry per pixel")
    for obj in (confusion, auc):
        if obj is not None and obj.num_classes != nc:
            raise ValueError("metric object and logits disagree on the number of classes")
    if confusion is None and auc is None:
        return
    cnt = confusion._cnt if confusion is not None else torch.zeros(2, dtype=torch.int64, device=logits.device)
    has_ign = ignore_index is not None
    _lib.call("ig_seg_metrics_update", logits.device,
              logits.data_ptr(), B, nc, H * W, lab.data_ptr(), _LABEL_DTYPES[lab.dtype], int(has_ign),
              int(ignore_index) if has_ign else 0, _lib.ptr(confusion._mat if confusion is not None else None),
              cnt.data_ptr(), auc.n_bins if auc is not None else 2, float(auc.min_score) if auc is not None else 0.0,
              float(auc.max_score) if auc is not None else 1.0, _lib.ptr(auc._pos if auc is not None else None),
              _lib.ptr(auc._neg if auc is not None else None))


class RunningRegressionMetrics:
    """Streaming regression statistics (reference: metrics.py:289-433); sums carried in float64 on the device."""

    def __init__(self, ee_bias: float = 0.05, ee_coef: float = 0.15, include_ee: bool = False, device=None) -> None:
        self.ee_bias = ee_bias
        self.ee_coef = ee_coef
        self.include_ee = include_ee
        self.device = _device(device)
        self._sums = torch.zeros(7, dtype=torch.float64, device=self.device)
        self._counts = torch.zeros(2, dtype=torch.int64, device=self.device)

    def reset(self) -> None:
        self._sums.zero_()
        self._counts.zero_()

    def update(self, y_true, y_pred, ignore_value: Optional[float] = None) -> None:
        x = _to_dev(y_true, self.device).reshape(-1).float().contiguous()
        y = _to_dev(y_pred, self.device).reshape(-1).float().contiguous()
        if x.shape != y.shape:
            raise ValueError("y_true and y_pred shapes differ.")
        has_ign = ignore_value is not None
        _lib.call("ig_regression_update", self.device,
                  x.data_ptr(), y.data_ptr(), x.numel(), int(has_ign), float(ignore_value) if has_ign else 0.0,
                  float(self.ee_bias), float(self.ee_coef), self._sums.data_ptr(), self._counts.data_ptr())

    def _state(self) -> dict:
        s, c = self._sums.cpu().numpy(), self._counts.cpu().numpy()
        return {"n": int(c[0]), "sum_x": s[0], "sum_y": s[1], "sum_xy": s[2], "sum_x2": s[3], "sum_y2": s[4],
                "sum_abs_error": s[5], "sum_squared_error": s[6], "within_ee_count": int(c[1])}

    n = property(lambda self: self._state()["n"])

    def mae(self) -> float:
        s = self._state()
        return float("nan") if s["n"] == 0 else s["sum_abs_error"] / s["n"]

    def rmse(self) -> float:
        s = self._state()
        return float("nan") if s["n"] == 0 else np.sqrt(s["sum_squared_error"] / s["n"])

    def r2_score(self) -> float:
        s = self._state()
        if s["n"] < 2:
            return float("nan")
        x_mean = s["sum_x"] / s["n"]
        ss_tot = s["sum_x2"] - s["n"] * x_mean * x_mean
        return float("nan") if ss_tot == 0 else 1 - (s["sum_squared_error"] / ss_tot)

    def pearson_corrcoef(self) -> float:
        s = self._state()
        if s["n"] < 2:
            return float("nan")
        n = s["n"]
        x_mean, y_mean = s["sum_x"] / n, s["sum_y"] / n
        cov_xy = s["sum_xy"] - n * x_mean * y_mean
        std_x = np.sqrt(s["sum_x2"] - n * x_mean * x_mean)
        std_y = np.sqrt(s["sum_y2"] - n * y_mean * y_mean)
        return float("nan") if std_x == 0 or std_y == 0 else cov_xy / (std_x * std_y)

    def ee_percentage(self) -> float:
        s = self._state()
        return float("nan") if s["n"] == 0 else (s["within_ee_count"] / s["n"]) * 100

    def compute(self) -> dict:
        return {"mae": self.mae(), "rmse": self.rmse(), "r2_score": self.r2_score(),
                "pearson_corrcoef": self.pearson_corrcoef(),
                "ee_percentage": self.ee_percentage() if self.include_ee else None,
                "ee_bias": self.ee_bias, "ee_coef": self.ee_coef}
