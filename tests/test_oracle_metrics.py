"""CPU: oracle/metrics.py against the fixture frozen from the real reference (tests/golden/metrics.npz,
made by oracle/gen_golden_metrics.py), against scikit-learn like the reference's own
tests/model_tests/test_metrics.py:51-140, and against the live reference module when it is present."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import metrics as OM


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "metrics.npz"))


@pytest.mark.parametrize("tag,nc", [("nc2", 2), ("nc13", 13)])
def test_eval_step_matches_reference_golden(gold, tag, nc):
    logits, labels = torch.from_numpy(gold[f"{tag}_logits"]), torch.from_numpy(gold[f"{tag}_labels"])
    lab, preds, probs = OM.segmentation_eval_step(logits, labels, -100)
    assert np.array_equal(probs, gold[f"{tag}_probs"])
    mat, total = OM.confusion_counts(lab, preds, nc)
    assert np.array_equal(mat, gold[f"{tag}_matrix"]) and total == int(gold[f"{tag}_total"])
    # the ignore mask applied inside update gives the same matrix
    mat2, total2 = OM.confusion_counts(labels.numpy(), torch.argmax(logits, 1).numpy(), nc, ignore_index=-100)
    assert np.array_equal(mat2, mat) and total2 == total
    pos, neg, n_pos, n_neg = OM.auc_hist(lab, probs, nc)
    assert np.array_equal(pos, gold[f"{tag}_pos"]) and np.array_equal(neg, gold[f"{tag}_neg"])  # float32 binning
    m, a = OM.confusion_metrics(mat, total), OM.auc_scores(pos, neg, n_pos, n_neg)
    got = np.array([m["accuracy"], m["precision"], m["recall"], m["f1"], m["jaccard"], a["roc_auc_macro"]])
    assert np.allclose(got, gold[f"{tag}_scalars"], rtol=1e-12, atol=0)
    assert np.allclose(a["roc_auc_per_class"], gold[f"{tag}_auc_per_class"], rtol=1e-12)
    assert np.allclose(m["jaccard_per_class"], gold[f"{tag}_jaccard_per_class"], rtol=1e-12)


def test_float64_scores_and_clamping_match_reference_golden(gold):
    pos, neg, _, _ = OM.auc_hist(gold["f64_labels"], gold["f64_scores"], 3, n_bins=257)
    assert np.array_equal(pos, gold["f64_pos"]) and np.array_equal(neg, gold["f64_neg"])
    # clamp / NaN rules of RunningAUC._bin (metrics.py:209-212)
    b = OM.auc_bins(np.array([0.0, 1.0, 1.5, -0.25, np.nan, 0.5], dtype=np.float32), 1024)
    assert b.tolist() == [0, 1023, 1023, 0, 0, 511]


def test_regression_matches_reference_golden(gold):
    s = OM.regression_sums(gold["reg_x"], gold["reg_y"])
    r = OM.regression_metrics(s, include_ee=True)
    got = np.array([r["mae"], r["rmse"], r["r2_score"], r["pearson_corrcoef"], r["ee_percentage"]])
    # the reference accumulates float32 partial sums; float64 sums agree to float32 round-off
    assert np.allclose(got, gold["reg_scalars"], rtol=2e-5)
    assert s["within_ee_count"] == int(gold["reg_within"])


def test_confusion_against_sklearn():
    sk = pytest.importorskip("sklearn.metrics")
    rng = np.random.default_rng(0)
    y_true = rng.integers(0, 3, 5000)
    y_pred = np.where(rng.random(5000) < 0.7, y_true, rng.integers(0, 3, 5000))
    mat, total = OM.confusion_counts(y_true, y_pred, 3)
    m = OM.confusion_metrics(mat, total)
    assert np.isclose(m["accuracy"], sk.accuracy_score(y_true, y_pred))
    assert np.isclose(m["precision"], sk.precision_score(y_true, y_pred, average="macro", zero_division=0))
    assert np.isclose(m["recall"], sk.recall_score(y_true, y_pred, average="macro", zero_division=0))
    assert np.isclose(m["f1"], sk.f1_score(y_true, y_pred, average="macro", zero_division=0))
    assert np.isclose(m["jaccard"], sk.jaccard_score(y_true, y_pred, average="macro", zero_division=0))
    logits = rng.normal(size=(5000, 3)) + 2.0 * np.eye(3)[y_true]
    probs = np.exp(logits) / np.exp(logits).sum(1, keepdims=True)
    pos, neg, n_pos, n_neg = OM.auc_hist(y_true, probs, 3, n_bins=4096)
    auc = OM.auc_scores(pos, neg, n_pos, n_neg)["roc_auc_macro"]
    assert abs(auc - sk.roc_auc_score(y_true, probs, multi_class="ovr", average="macro")) < 5e-3


def test_edge_cases():
    mat, total = OM.confusion_counts(np.array([], dtype=np.int64), np.array([], dtype=np.int64), 4)
    assert total == 0 and mat.sum() == 0 and np.isnan(OM.confusion_metrics(mat, total)["accuracy"])
    mat, total = OM.confusion_counts(np.full(7, -100), np.zeros(7, dtype=np.int64), 4, ignore_index=-100)
    assert total == 0
    with pytest.raises(ValueError):
        OM.confusion_counts(np.array([0, 5]), np.array([0, 0]), 4)
    with pytest.raises(ValueError):
        OM.confusion_counts(np.array([0, 1]), np.array([0]), 4)
    a = OM.auc_scores(np.zeros((2, 8), np.int64), np.ones((2, 8), np.int64), np.zeros(2), np.full(2, 8))
    assert np.isnan(a["roc_auc_macro"])


@pytest.mark.skipif(not os.path.exists("/root/reference/instageo/model/metrics.py"), reason="reference not mounted")
def test_oracle_equals_live_reference():
    from oracle.gen_golden_metrics import load_reference, synth_eval_batch
    ref = load_reference()
    for nc, seed in ((2, 3), (5, 4)):
        logits, labels = synth_eval_batch(seed, 3, nc, 24)
        lab, preds, probs = OM.segmentation_eval_step(logits, labels)
        cm, auc = ref.RunningConfusionMatrix(nc), ref.RunningAUC(nc, n_bins=300)
        for lo in range(0, lab.size, 500):  # streamed in chunks like the reference's tests
            cm.update(lab[lo:lo + 500], preds[lo:lo + 500])
            auc.update(lab[lo:lo + 500], probs[lo:lo + 500])
        mat, total = OM.confusion_counts(lab, preds, nc)
        pos, neg, n_pos, n_neg = OM.auc_hist(lab, probs, nc, n_bins=300)
        assert np.array_equal(mat, cm.matrix) and total == cm.total
        assert np.array_equal(pos, auc.pos_hist) and np.array_equal(neg, auc.neg_hist)
        assert np.array_equal(n_pos, auc.n_pos) and np.array_equal(n_neg, auc.n_neg)
        want, got = cm.compute(), OM.confusion_metrics(mat, total)
        assert all(np.allclose(got[k], want[k], rtol=1e-12) for k in want)
        assert np.allclose(OM.auc_scores(pos, neg, n_pos, n_neg)["roc_auc_per_class"],
                           auc.score()["roc_auc_per_class"], rtol=1e-12, equal_nan=True)


def test_host_mirror_derivations_match_reference_golden(gold):
    """the mirror's host-side arithmetic (what it does with the integers read back from the device), on CPU: pure
    functions against the fixture frozen from the reference, and the class wrappers over a stubbed read-back"""
    from instageo_b200.model import metrics as MM
    for tag in ("nc2", "nc13"):
        mat, total = gold[f"{tag}_matrix"], int(gold[f"{tag}_total"])
        s = MM.confusion_summary(mat, total)
        got = np.array([s["accuracy"], s["precision"], s["recall"], s["f1"], s["jaccard"]])
        assert np.allclose(got, gold[f"{tag}_scalars"][:5], rtol=1e-12, atol=0)
        assert np.allclose(s["jaccard_per_class"], gold[f"{tag}_jaccard_per_class"], rtol=1e-12)
        per = MM.auc_from_histograms(gold[f"{tag}_pos"], gold[f"{tag}_neg"])
        assert np.allclose(per, gold[f"{tag}_auc_per_class"], rtol=1e-12)
        cm = object.__new__(MM.RunningConfusionMatrix)            # no device: stub the read-back
        cm._sync = lambda mat=mat, total=total: (mat, total)
        want = OM.confusion_metrics(mat, total)
        out = cm.compute()
        assert set(out) == set(want) and all(np.allclose(out[k], want[k], rtol=1e-12) for k in want)
        assert np.allclose(cm.precision(), want["precision_per_class"]) and np.allclose(cm.f1(), want["f1_per_class"])
        assert np.isclose(cm.accuracy(), want["accuracy"]) and np.allclose(cm.recall(), want["recall_per_class"])
        assert set(cm.compute(include_per_class=False)) == {"accuracy", "precision", "recall", "f1", "jaccard"}
        auc = object.__new__(MM.RunningAUC)
        auc.num_classes = mat.shape[0]
        saved = MM.RunningAUC.__dict__["pos_hist"], MM.RunningAUC.__dict__["neg_hist"]
        type(auc).pos_hist = property(lambda self, t=tag: gold[f"{t}_pos"])
        type(auc).neg_hist = property(lambda self, t=tag: gold[f"{t}_neg"])
        try:
            sc = auc.score()
            assert np.isclose(sc["roc_auc_macro"], gold[f"{tag}_scalars"][5], rtol=1e-12)
            assert np.isclose(auc._auc_one_class(1), gold[f"{tag}_auc_per_class"][1], rtol=1e-12)
        finally:
            MM.RunningAUC.pos_hist, MM.RunningAUC.neg_hist = saved
    empty = MM.confusion_summary(np.zeros((3, 3), dtype=np.int64), 0)
    assert np.isnan(empty["accuracy"]) and empty["f1"] == 0.0
    assert np.isnan(MM.auc_from_histograms(np.zeros((1, 4), np.int64), np.ones((1, 4), np.int64))[0])
