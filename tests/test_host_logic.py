"""CPU tests of the host-side pieces next to the path: text batch builder and AUC."""
import numpy as np
import pytest

from oracle import refport


def test_parse_libsvm_and_libffm(pkg):
    d = pkg.data
    rows, cols, vals, targets = d.parse(["1 3:1 7:0.5", "", "0 1:2"], "libsvm")
    assert rows.tolist() == [0, 0, 1] and cols.tolist() == [2, 6, 0]          # keys are 1-based (SampleParser.scala:37)
    assert vals.tolist() == [1.0, 0.5, 2.0] and targets.tolist() == [1.0, 0.0]
    rows, cols, vals, targets = d.parse(["1 0:3:1 1:7:1"], "libffm")
    assert cols.tolist() == [2, 6] and rows.tolist() == [0, 0]
    with pytest.raises(ValueError):
        d.parse([], "csv")
    # round trip of a synthetic batch
    _, feats = pkg.synth.make_feats(1, 0, 8, 5, 500)
    t = pkg.synth.make_targets(1, feats, 8, 5)
    r2, c2, _, t2 = d.parse(d.to_libsvm(feats, t, 5))
    assert np.array_equal(c2, feats) and np.array_equal(t2, t)
    assert np.array_equal(r2, np.repeat(np.arange(8), 5))


def test_auc_matches_oracle_and_sklearn(pkg):
    rng = np.random.default_rng(0)
    t = (rng.uniform(size=500) < 0.3).astype(np.float32)
    p = rng.uniform(size=500).astype(np.float32) * 0.5 + 0.3 * t
    a = pkg.metrics.auc(t, p)
    assert abs(a - refport.auc(t, p)) < 1e-12
    from sklearn.metrics import roc_auc_score
    assert abs(a - roc_auc_score(t, p)) < 1e-9     # no ties here
    assert np.isnan(pkg.metrics.auc([1, 1], [0.2, 0.3]))


def test_optimizer_oracle_textbook_forms():
    w, g = np.array([1.0, -2.0], np.float32), np.array([0.5, 0.25], np.float32)
    assert np.allclose(refport.optimizer_update("sgd", w, g, {}, 0.1), [0.95, -2.025])
    st = {}
    w1 = refport.optimizer_update("adam", w, g, st, 0.1, step=1)
    # first adam step moves every weight by ~lr against the gradient sign
    assert np.allclose(w - w1, 0.1, atol=1e-4)
    st = {}
    w1 = refport.optimizer_update("momentum", w, g, st, 0.1)
    w2 = refport.optimizer_update("momentum", w1, g, st, 0.1)
    assert np.allclose(w1 - w2, 0.1 * (0.9 * g + g))
