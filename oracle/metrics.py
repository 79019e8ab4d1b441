"""ORACLE (test infrastructure only): CPU restatement of the reference's streaming eval metrics.

Follows instageo/model/metrics.py (RunningConfusionMatrix :63-172, RunningAUC :180-281,
RunningRegressionMetrics :289-433) and the per-step glue that feeds them,
instageo/model/segmentation.py:107-156 (`_shared_step`) and :202-213 (`predict_step`).
Pinned by tests/test_oracle_metrics.py: against the LIVE reference module when /root/reference is
present (metrics.py only needs numpy), against scikit-learn like the reference's own
tests/model_tests/test_metrics.py:51-140, and against tests/golden/metrics.npz frozen from the
reference by oracle/gen_golden_metrics.py.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import numpy as np
import torch


# ------------------------------------------------------------------ confusion matrix (metrics.py:86-108)
def confusion_counts(y_true, y_pred, num_classes: int, ignore_index=None):
    """(matrix int64 [k,k] with rows = truth, cols = prediction, number of valid samples).
    Raises ValueError like np.bincount / reshape do for labels outside [0, k)."""
    t = np.asarray(y_true).ravel().astype(np.int64)
    p = np.asarray(y_pred).ravel().astype(np.int64)
    if t.shape != p.shape:
        raise ValueError("y_true and y_pred shapes differ.")
    if ignore_index is not None:
        keep = t != ignore_index
        t, p = t[keep], p[keep]
    k = num_classes
    mat = np.zeros((k, k), dtype=np.int64)
    if t.size == 0:
        return mat, 0
    flat = t * k + p
    if flat.min() < 0 or flat.max() >= k * k:
        raise ValueError("label outside [0, num_classes)")
    np.add.at(mat.reshape(-1), flat, 1)
    return mat, int(t.size)


def _sdiv(num, den):
    den = np.asarray(den, dtype=float)
    out = np.zeros_like(den)
    np.divide(num, den, out=out, where=den != 0)
    return out


def confusion_metrics(mat: np.ndarray, total: int, include_per_class: bool = True) -> dict:
    """metrics.py:110-166: macro accuracy / precision / recall / F1 / Jaccard (+ per-class lists)."""
    tp = np.diag(mat)
    fp = mat.sum(axis=0) - tp
    fn = mat.sum(axis=1) - tp
    prec, rec = _sdiv(tp, tp + fp), _sdiv(tp, tp + fn)
    f1 = _sdiv(2 * prec * rec, prec + rec)
    jac = _sdiv(tp, tp + fp + fn)
    out = {"accuracy": float("nan") if total == 0 else tp.sum() / total, "precision": prec.mean(),
           "recall": rec.mean(), "f1": f1.mean(), "jaccard": jac.mean()}
    if include_per_class:
        out.update(precision_per_class=prec.tolist(), recall_per_class=rec.tolist(), f1_per_class=f1.tolist(),
                   jaccard_per_class=jac.tolist())
    return out


# ------------------------------------------------------------------ ROC-AUC histograms (metrics.py:209-265)
def auc_bins(scores: np.ndarray, n_bins: int = 1024, lo: float = 0.0, hi: float = 1.0) -> np.ndarray:
    """Vectorised `RunningAUC._bin` (metrics.py:209-212).  The reference evaluates
    `int((s - lo) / (hi - lo) * (n_bins - 1))` on NumPy SCALARS of the score's dtype; with the pinned
    numpy 2.2.6 (uv.lock:3258) Python floats/ints are weak, so float32 scores are binned in float32
    arithmetic and float64 scores in float64.  Values clamped to lo / hi become Python floats (f64);
    the clamp results 0 and n_bins-1 are the same in either precision.  NaN -> bin 0 (Python's
    max(lo, nan) keeps lo)."""
    s = np.asarray(scores)
    dt = s.dtype if s.dtype in (np.float32, np.float64) else np.float64
    s = s.astype(dt, copy=False)
    x = (s - dt.type(lo)) / dt.type(hi - lo) * dt.type(n_bins - 1)
    b = np.where(s > lo, x, 0).astype(np.int64)           # NaN and s <= lo -> clamp to lo -> bin 0
    b = np.where(s >= hi, int((hi - lo) / (hi - lo) * (n_bins - 1)), b)
    return b.astype(np.int32)


def auc_hist(y_true, y_score, num_classes: int, n_bins: int = 1024, lo: float = 0.0, hi: float = 1.0):
    """(pos_hist, neg_hist int64 [k, n_bins], n_pos, n_neg int64 [k]) of one `RunningAUC.update`."""
    t = np.asarray(y_true).ravel()
    s = np.asarray(y_score)
    if s.ndim == 1:
        if num_classes != 2:
            raise ValueError("1-D y_score needs num_classes == 2")
        s = np.stack([1 - s, s], axis=1)
    if t.shape[0] != s.shape[0] or s.shape[1] != num_classes:
        raise ValueError("y_true / y_score shape mismatch")
    pos = np.zeros((num_classes, n_bins), np.int64)
    neg = np.zeros((num_classes, n_bins), np.int64)
    for c in range(num_classes):
        b = auc_bins(s[:, c], n_bins, lo, hi)
        is_pos = t == c
        pos[c] = np.bincount(b[is_pos], minlength=n_bins)
        neg[c] = np.bincount(b[~is_pos], minlength=n_bins)
    return pos, neg, pos.sum(1), neg.sum(1)


def auc_scores(pos, neg, n_pos, n_neg) -> dict:
    """metrics.py:246-275: per-class trapezoid over the bin histograms, nan-mean macro."""
    per = []
    for c in range(pos.shape[0]):
        if n_pos[c] == 0 or n_neg[c] == 0:
            per.append(float("nan"))
            continue
        cum_neg_before = np.concatenate([[0], np.cumsum(neg[c])[:-1]]).astype(float)
        a = float((pos[c] * cum_neg_before).sum() + 0.5 * (pos[c].astype(float) * neg[c]).sum())
        per.append(a / (float(n_pos[c]) * float(n_neg[c])))
    per = np.array(per)
    return {"roc_auc_macro": np.nanmean(per) if not np.isnan(per).all() else float("nan"),
            "roc_auc_per_class": per.tolist()}


# ------------------------------------------------------------------ the eval step (segmentation.py:117-156)
def segmentation_eval_step(logits: torch.Tensor, labels: torch.Tensor, ignore_index: int = -100):
    """What `_shared_step` hands to the metric objects: (labels int64 [n], preds int64 [n],
    probs float32 [n, nc]) over the non-ignored pixels, in pixel order."""
    labels = labels.long()
    keep = labels.ne(ignore_index).reshape(-1)
    preds = torch.argmax(logits, dim=1).reshape(-1)[keep]
    probs = torch.softmax(logits, dim=1).permute(0, 2, 3, 1).reshape(-1, logits.size(1))[keep]
    return labels.reshape(-1)[keep].numpy().astype(np.int64), preds.numpy().astype(np.int64), probs.numpy()


def positive_probability(logits: torch.Tensor) -> torch.Tensor:
    """`predict_step`, segmentation.py:211-213: softmax over classes, channel 1."""
    return torch.softmax(logits, dim=1)[:, 1, :, :]


# ------------------------------------------------------------------ regression sums (metrics.py:330-356)
def regression_sums(y_true, y_pred, ee_bias: float = 0.05, ee_coef: float = 0.15) -> dict:
    """The running sums of one update, accumulated in float64 (the reference sums float32 arrays in
    float32; the device path and this oracle both carry float64, compared within a tolerance)."""
    x = np.asarray(y_true, dtype=np.float64).ravel()
    y = np.asarray(y_pred, dtype=np.float64).ravel()
    err = np.abs(y - x)
    return {"n": int(x.size), "sum_x": x.sum(), "sum_y": y.sum(), "sum_xy": (x * y).sum(), "sum_x2": (x * x).sum(),
            "sum_y2": (y * y).sum(), "sum_abs_error": err.sum(), "sum_squared_error": (err * err).sum(),
            "within_ee_count": int(np.sum(np.abs(np.asarray(y_pred, np.float32).ravel()
                                                 - np.asarray(y_true, np.float32).ravel())
                                          <= (np.float32(ee_bias) + np.float32(ee_coef)
                                              * np.asarray(y_true, np.float32).ravel())))}


def regression_metrics(s: dict, include_ee: bool = False) -> dict:
    """metrics.py:358-433 from the running sums."""
    n = s["n"]
    nan = float("nan")
    mae = nan if n == 0 else s["sum_abs_error"] / n
    rmse = nan if n == 0 else float(np.sqrt(s["sum_squared_error"] / n))
    r2 = pear = nan
    if n >= 2:
        xm, ym = s["sum_x"] / n, s["sum_y"] / n
        ss_tot = s["sum_x2"] - n * xm * xm
        if ss_tot != 0:
            r2 = 1 - s["sum_squared_error"] / ss_tot
        sx, sy = np.sqrt(ss_tot), np.sqrt(s["sum_y2"] - n * ym * ym)
        if sx != 0 and sy != 0:
            pear = (s["sum_xy"] - n * xm * ym) / (sx * sy)
    return {"mae": mae, "rmse": rmse, "r2_score": r2, "pearson_corrcoef": pear,
            "ee_percentage": (nan if n == 0 else s["within_ee_count"] / n * 100) if include_ee else None}
