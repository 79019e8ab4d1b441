"""Freeze outputs of the REAL reference metric classes (instageo/model/metrics.py, loaded from
/root/reference by file path -- it only needs numpy) into tests/golden/metrics.npz.

    python -m oracle.gen_golden_metrics

Runs only in the dev container (the reference does not travel to the GPU box); the committed
fixture pins oracle/metrics.py on any machine.  NumPy here is 2.3 (NEP 50 scalar rules, like the
reference's pinned 2.2.6), so float32 scores are binned in float32.
"""
from __future__ import annotations

import importlib.util
import os

import numpy as np
import torch

REF = "/root/reference/instageo/model/metrics.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "metrics.npz")


def load_reference():
    spec = importlib.util.spec_from_file_location("_ref_metrics", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def synth_eval_batch(seed: int, batch: int, nc: int, size: int, ignore_index: int = -100, ignore_frac: float = 0.1):
    """Seeded logits [B, nc, S, S] f32 and int64 labels with ~ignore_frac ignored pixels."""
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(batch, nc, size, size, generator=g) * 3.0
    labels = torch.randint(0, nc, (batch, size, size), generator=g)
    labels[torch.rand(batch, size, size, generator=g) < ignore_frac] = ignore_index
    return logits, labels


def main():
    ref = load_reference()
    out = {}
    for tag, nc in (("nc2", 2), ("nc13", 13)):
        logits, labels = synth_eval_batch(1042 + nc, 2, nc, 32)
        keep = labels.ne(-100).reshape(-1)
        preds = torch.argmax(logits, 1).reshape(-1)[keep].numpy().astype(np.int64)
        probs = torch.softmax(logits, 1).permute(0, 2, 3, 1).reshape(-1, nc)[keep].numpy()
        lab = labels.reshape(-1)[keep].numpy().astype(np.int64)
        cm = ref.RunningConfusionMatrix(nc, -100)
        cm.update(labels.numpy(), torch.argmax(logits, 1).numpy())
        auc = ref.RunningAUC(nc)
        auc.update(lab, probs)
        m, a = cm.compute(), auc.score()
        out.update({f"{tag}_logits": logits.numpy(), f"{tag}_labels": labels.numpy(), f"{tag}_probs": probs,
                    f"{tag}_matrix": cm.matrix, f"{tag}_total": np.int64(cm.total),
                    f"{tag}_pos": auc.pos_hist, f"{tag}_neg": auc.neg_hist,
                    f"{tag}_scalars": np.array([m["accuracy"], m["precision"], m["recall"], m["f1"], m["jaccard"],
                                                a["roc_auc_macro"]]),
                    f"{tag}_auc_per_class": np.array(a["roc_auc_per_class"]),
                    f"{tag}_jaccard_per_class": np.array(m["jaccard_per_class"])})
    # float64 scores are binned in float64 by the reference (its own tests use them)
    rng = np.random.default_rng(7)
    s64 = rng.random((4000, 3))
    s64 /= s64.sum(1, keepdims=True)
    s64[:5, 0] = [0.0, 1.0, 1.5, -0.25, np.nan]
    t = rng.integers(0, 3, 4000)
    auc = ref.RunningAUC(3, n_bins=257)
    auc.update(t, s64)
    out.update(f64_scores=s64, f64_labels=t, f64_pos=auc.pos_hist, f64_neg=auc.neg_hist)
    # regression
    x = rng.random(5000).astype(np.float32) * 2
    y = (x + rng.normal(0, 0.2, 5000)).astype(np.float32)
    rm = ref.RunningRegressionMetrics(include_ee=True)
    rm.update(x[:3000], y[:3000])
    rm.update(x[3000:], y[3000:])
    r = rm.compute()
    out.update(reg_x=x, reg_y=y, reg_scalars=np.array([r["mae"], r["rmse"], r["r2_score"], r["pearson_corrcoef"],
                                                      r["ee_percentage"]], dtype=np.float64),
               reg_within=np.int64(rm.within_ee_count))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
