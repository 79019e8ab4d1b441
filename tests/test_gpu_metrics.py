"""GPU parity of the eval-metric kernels (csrc/metrics.cu) against oracle/metrics.py and the fixture frozen
from the reference (tests/golden/metrics.npz).  Counts are integers: bit-exact."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import metrics as OM

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "metrics.npz"))


def _batch(seed, batch, nc, size, ignore_frac=0.1):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(batch, nc, size, size, generator=g) * 3.0
    labels = torch.randint(0, nc, (batch, size, size), generator=g)
    labels[torch.rand(batch, size, size, generator=g) < ignore_frac] = -100
    return logits, labels


@pytest.mark.parametrize("tag,nc", [("nc2", 2), ("nc13", 13)])
def test_confusion_and_auc_match_reference_golden(cuda_dev, gold, tag, nc):
    from instageo_b200.model.metrics import RunningAUC, RunningConfusionMatrix
    logits, labels = torch.from_numpy(gold[f"{tag}_logits"]), torch.from_numpy(gold[f"{tag}_labels"])
    cm = RunningConfusionMatrix(nc, ignore_index=-100, device=cuda_dev)
    cm.update(labels.to(cuda_dev), torch.argmax(logits, 1).to(cuda_dev))
    assert np.array_equal(cm.matrix, gold[f"{tag}_matrix"]) and cm.total == int(gold[f"{tag}_total"])
    keep = labels.ne(-100).reshape(-1)
    auc = RunningAUC(nc, device=cuda_dev)
    auc.update(labels.reshape(-1)[keep].numpy(), gold[f"{tag}_probs"])     # numpy in, like the reference's callers
    assert np.array_equal(auc.pos_hist, gold[f"{tag}_pos"]) and np.array_equal(auc.neg_hist, gold[f"{tag}_neg"])
    m, a = cm.compute(), auc.score()
    got = np.array([m["accuracy"], m["precision"], m["recall"], m["f1"], m["jaccard"], a["roc_auc_macro"]])
    assert np.allclose(got, gold[f"{tag}_scalars"], rtol=1e-12, atol=0)
    assert np.allclose(a["roc_auc_per_class"], gold[f"{tag}_auc_per_class"], rtol=1e-12)


def test_auc_float64_scores_and_clamping(cuda_dev, gold):
    from instageo_b200.model.metrics import RunningAUC
    auc = RunningAUC(3, n_bins=257, device=cuda_dev)
    s, t = gold["f64_scores"], gold["f64_labels"]
    auc.update(t[:1500], s[:1500])
    auc.update(torch.from_numpy(t[1500:]).to(cuda_dev), torch.from_numpy(s[1500:]).to(cuda_dev))
    assert np.array_equal(auc.pos_hist, gold["f64_pos"]) and np.array_equal(auc.neg_hist, gold["f64_neg"])
    b = RunningAUC(2, n_bins=16, device=cuda_dev)  # 1-D positive-class scores, metrics.py:224-228
    p1 = np.array([0.0, 1.0, 0.3, 0.9], dtype=np.float32)
    b.update(np.array([0, 1, 1, 0]), p1)
    pos, neg, _, _ = OM.auc_hist(np.array([0, 1, 1, 0]), p1, 2, n_bins=16)
    assert np.array_equal(b.pos_hist, pos) and np.array_equal(b.neg_hist, neg)
    big = RunningAUC(3, n_bins=20000, device=cuda_dev)  # histograms too large for shared memory: global atomics
    big.update(t, s.astype(np.float32))
    pos, neg, _, _ = OM.auc_hist(t, s.astype(np.float32), 3, n_bins=20000)
    assert np.array_equal(big.pos_hist, pos) and np.array_equal(big.neg_hist, neg)


@pytest.mark.parametrize("nc,batch,size,ldt", [(2, 3, 224, torch.int64), (13, 2, 224, torch.int64), (5, 4, 36, torch.int32),
                                               (16, 1, 64, torch.uint8), (1, 2, 32, torch.int8)])
def test_fused_eval_step_matches_oracle(cuda_dev, nc, batch, size, ldt):
    """logits -> argmax + softmax + ignore mask + confusion + ROC histograms in one kernel."""
    from instageo_b200.model.metrics import RunningAUC, RunningConfusionMatrix, segmentation_eval_update
    ign = -100 if ldt in (torch.int64, torch.int32) else (255 if ldt == torch.uint8 else -1)
    logits, labels = _batch(nc * 31 + size, batch, nc, size)
    labels[labels == -100] = ign
    cm, auc = RunningConfusionMatrix(nc, ign, device=cuda_dev), RunningAUC(nc, device=cuda_dev)
    for _ in range(2):  # accumulation across steps
        segmentation_eval_update(logits.to(cuda_dev), labels.to(ldt).to(cuda_dev), cm, auc, ignore_index=ign)
    lab, preds, probs = OM.segmentation_eval_step(logits, labels, ign)
    mat, total = OM.confusion_counts(lab, preds, nc)
    assert np.array_equal(cm.matrix, 2 * mat) and cm.total == 2 * total     # argmax: bit-exact
    pos, neg, n_pos, n_neg = OM.auc_hist(lab, probs, nc)
    assert np.array_equal(auc.n_pos, 2 * n_pos) and np.array_equal(auc.n_neg, 2 * n_neg)
    # softmax on the device differs from torch's CPU softmax by <= 2 ulp (expf), so a score sitting on a bin
    # edge may move to the neighbouring bin: compare cumulative histograms, allowing that.  Tolerance: the number
    # of moved samples is below 1e-3 of all samples and no sample moves further than one bin.
    for got, want in ((auc.pos_hist, 2 * pos), (auc.neg_hist, 2 * neg)):
        moved = np.abs(np.cumsum(got, 1) - np.cumsum(want, 1))
        assert moved.sum() <= 1e-3 * max(1, want.sum())
        assert np.abs(got - want).sum() <= 2 * moved.sum()
    want_auc = OM.auc_scores(pos, neg, n_pos, n_neg)["roc_auc_per_class"]
    assert np.allclose(auc.score()["roc_auc_per_class"], want_auc, atol=1e-5, equal_nan=True)


def test_confusion_edge_cases(cuda_dev):
    from instageo_b200.model.metrics import RunningConfusionMatrix
    cm = RunningConfusionMatrix(4, device=cuda_dev)
    cm.update(np.array([], dtype=np.int64), np.array([], dtype=np.int64))                     # empty
    assert cm.total == 0 and np.isnan(cm.accuracy())
    rng = np.random.default_rng(3)
    for n in (1, 3, 5, 1027, 4099):                                                              # ragged tails
        t, p = rng.integers(0, 4, n), rng.integers(0, 4, n)
        cm.reset()
        cm.update(t, p)
        assert np.array_equal(cm.matrix, OM.confusion_counts(t, p, 4)[0]) and cm.total == n
    cm.reset()
    t = torch.randint(0, 4, (1001,), device=cuda_dev)
    cm.update(t[1:], t[1:].to(torch.int8))                                                      # misaligned views
    assert cm.total == 1000 and np.trace(cm.matrix) == 1000
    ig = RunningConfusionMatrix(4, ignore_index=-100, device=cuda_dev)
    ig.update(np.full(64, -100), np.zeros(64, dtype=np.int64))                                  # everything ignored
    assert ig.total == 0 and ig.matrix.sum() == 0
    bad = RunningConfusionMatrix(4, device=cuda_dev)
    bad.update(np.array([0, 5, 1, 2]), np.array([0, 0, 1, 2]))                                  # label >= k
    with pytest.raises(ValueError):
        bad.compute()
    with pytest.raises(ValueError):
        cm.update(np.zeros(3, dtype=np.int64), np.zeros(4, dtype=np.int64))
    big = RunningConfusionMatrix(40, device=cuda_dev)                                           # k up to 64
    t, p = rng.integers(0, 40, 10000), rng.integers(0, 40, 10000)
    big.update(t, p)
    assert np.array_equal(big.matrix, OM.confusion_counts(t, p, 40)[0])


def test_model_argmax_feeds_confusion(cuda_dev):
    """eval path end to end: fused-head int8 class map -> confusion matrix, nothing leaves the device."""
    from instageo_b200.model import PrithviSeg
    from instageo_b200.model.metrics import RunningConfusionMatrix
    from oracle import prithvi as P
    sd = P.make_state_dict("prithvi_eo_tiny", 1, 4, depth=1, seed=2, stress=True)
    model = PrithviSeg(temporal_step=1, num_classes=4, load_pretrained_weights=False, variant="prithvi_eo_tiny", depth=1)
    model.load_state_dict(sd, strict=True)
    model.to(cuda_dev).eval()
    x = torch.randn(2, 6, 1, 224, 224, generator=torch.Generator().manual_seed(0)).to(cuda_dev)
    pred = model.predict(x)
    labels = torch.randint(0, 4, (2, 224, 224), device=cuda_dev)
    cm = RunningConfusionMatrix(4, device=cuda_dev)
    cm.update(labels, pred)
    want, total = OM.confusion_counts(labels.cpu().numpy(), pred.cpu().numpy(), 4)
    assert np.array_equal(cm.matrix, want) and cm.total == total == 2 * 224 * 224


def test_regression_metrics(cuda_dev, gold):
    from instageo_b200.model.metrics import RunningRegressionMetrics
    rm = RunningRegressionMetrics(include_ee=True, device=cuda_dev)
    x, y = gold["reg_x"], gold["reg_y"]
    rm.update(x[:3000], y[:3000])
    rm.update(torch.from_numpy(x[3000:]).to(cuda_dev), torch.from_numpy(y[3000:]).to(cuda_dev))
    r = rm.compute()
    got = np.array([r["mae"], r["rmse"], r["r2_score"], r["pearson_corrcoef"], r["ee_percentage"]])
    assert np.allclose(got, gold["reg_scalars"], rtol=2e-5)  # reference sums in float32, device in float64
    s = OM.regression_sums(x, y)
    st = rm._state()
    assert st["n"] == s["n"] and st["within_ee_count"] == s["within_ee_count"] == int(gold["reg_within"])
    for k in ("sum_x", "sum_y", "sum_xy", "sum_x2", "sum_y2", "sum_abs_error", "sum_squared_error"):
        assert np.isclose(st[k], s[k], rtol=1e-12)       # float64 sums: only summation-order noise
    ig = RunningRegressionMetrics(device=cuda_dev)
    xi = x.copy()
    xi[::7] = -100.0
    ig.update(xi, y, ignore_value=-100.0)
    assert ig.n == int((xi != -100.0).sum())
    assert np.isnan(RunningRegressionMetrics(device=cuda_dev).mae())


def test_out_of_range_predictions_need_no_host_round_trip(cuda_dev):
    """non-int8 predictions are clamped on the device (an out-of-range value stays out of range and lands in the kernel's
    counter); the ValueError of the reference surfaces at the first read-back, not through a .max() on the host"""
    import numpy as np
    import torch
    from instageo_b200.model.metrics import RunningConfusionMatrix
    cm = RunningConfusionMatrix(3, device=cuda_dev)
    y = torch.tensor([0, 1, 2, 1], device=cuda_dev)
    cm.update(y, torch.tensor([0, 1, 2, 2], device=cuda_dev, dtype=torch.int64))
    assert cm.matrix.sum() == 4
    cm.update(y, torch.tensor([0, 1, 2, 700], device=cuda_dev, dtype=torch.int32))
    with pytest.raises(ValueError, match="outside"):
        cm.matrix
