"""GPU parity: Internal<M>Model.forward / .backward through the C ABI vs the oracle.

Tolerance (north_star): forward outputs and gradients within 1e-5 relative in fp32, measured
against the largest magnitude of the tensor, plus the fp32 oracle's own distance to its fp64 twin
(the oracle is itself one fp32 rounding path among many)."""
import numpy as np
import pytest

from common import CONFIGS, assert_close, away_from_kinks, make_inputs, oracle_model, pkg_model

pytestmark = pytest.mark.gpu

SHAPES = [(1, 3, 4), (7, 5, 8), (64, 39, 16), (130, 39, 16), (33, 6, 12), (16, 4, 6)]


@pytest.mark.parametrize("mode", [1, 0], ids=["tc3xtf32", "simt"])
@pytest.mark.parametrize("name", list(CONFIGS))
@pytest.mark.parametrize("B,F,K", SHAPES)
def test_forward_backward_parity(gpu_pkg, name, B, F, K, mode):
    if name == "pnn" and F < 2:
        pytest.skip("PNN needs two fields")
    if mode == 0 and name in ("lr", "fm"):
        pytest.skip("no dense contraction in this model")
    index, w, bias, emb, mats, targets = make_inputs(name, 3 * B, F, K, seed=B * 1000 + F)
    o32, o64 = oracle_model(name, F, K), oracle_model(name, F, K, np.float64)
    index, w, emb, ids = away_from_kinks(o64, 3 * B, F, K, index, w, bias, emb, mats, keep=B)
    targets = targets[ids]
    m = pkg_model(gpu_pkg, name, F, K)
    m.setGemmMode(mode)
    assert m.getMatsSize() == o32.mats_size()
    # forward
    p_ref = o32.forward(B, index, w, bias, emb, mats)
    p_64 = o64.forward(B, index, w, bias, emb, mats)
    p = m.forward(B, index, w, bias, emb, mats)
    assert_close(p, p_ref, what=f"{name} preds", ref64=p_64)
    # backward: buffers come back holding gradients
    cp = lambda a: None if a is None else a.copy()
    gw, gb, ge, gm = cp(w), cp(bias), cp(emb), cp(mats)
    loss = m.backward(B, index, gw, gb, ge, gm, targets)
    rw, rb, re, rm = cp(w), cp(bias), cp(emb), cp(mats)
    rloss = o32.backward(B, index, rw, rb, re, rm, targets)
    dw, db, de, dm = (None if a is None else a.astype(np.float64) for a in (w, bias, emb, mats))
    dloss = o64.backward(B, index, dw, db, de, dm, targets)
    assert abs(loss - rloss) <= 1e-5 * abs(rloss) + 2 * abs(rloss - dloss)
    assert_close(gw, rw, what=f"{name} dweights", ref64=dw)
    # dbias = sum_b dlogit_b cancels heavily: its rounding error scales with sum |dlogit_b|
    assert_close(gb, rb, what=f"{name} dbias", ref64=db, atol=1e-6 * np.abs(rw).sum() / max(1, F))
    if emb is not None:
        assert_close(ge, re, what=f"{name} dembedding", ref64=de)
    if mats is not None:
        assert_close(gm, rm, what=f"{name} dmats", ref64=dm)
    m.close()


@pytest.mark.parametrize("name", ["lr", "fm", "deepfm"])
def test_unsorted_index(gpu_pkg, name):
    """Scatter accepts any index < batchSize (nn/Scatter.scala:17-36), not only the sorted COO rows."""
    B, F, K = 9, 5, 8
    index, w, bias, emb, mats, targets = make_inputs(name, B, F, K, seed=5)
    rng = np.random.default_rng(1)
    index = rng.integers(0, B, index.shape[0]).astype(np.int32)
    o32 = oracle_model(name, F, K)
    m = pkg_model(gpu_pkg, name, F, K)
    assert_close(m.forward(B, index, w, bias, emb, mats), o32.forward(B, index, w, bias, emb, mats),
                 what="preds")
    cp = lambda a: None if a is None else a.copy()
    gw, gb, ge, gm = cp(w), cp(bias), cp(emb), cp(mats)
    rw, rb, re, rm = cp(w), cp(bias), cp(emb), cp(mats)
    loss = m.backward(B, index, gw, gb, ge, gm, targets)
    rloss = o32.backward(B, index, rw, rb, re, rm, targets)
    assert abs(loss - rloss) <= 1e-5 * abs(rloss)
    assert_close(gw, rw, what="dweights")
    if emb is not None:
        assert_close(ge, re, what="dembedding")
    m.close()


def test_errors(gpu_pkg):
    """index >= batchSize -> IllegalArgumentException in the reference (nn/Scatter.scala:29-30);
    nnz != B*F breaks the Reshape to [B,F,K] (SecondOrderEncoder.scala:29)."""
    B, F, K = 4, 3, 4
    index, w, bias, emb, mats, targets = make_inputs("fm", B, F, K)
    m = pkg_model(gpu_pkg, "fm", F, K)
    bad = index.copy()
    bad[-1] = B
    with pytest.raises(ValueError, match="index should smaller"):
        m.forward(B, bad, w, bias, emb, None)
    with pytest.raises(ValueError, match="nnz"):
        m.forward(B, index[:-1], w[:-1], bias, emb[:-K], None)
    # the handle still works after an error
    o32 = oracle_model("fm", F, K)
    assert_close(m.forward(B, index, w, bias, emb, None), o32.forward(B, index, w, bias, emb, None))
    with pytest.raises(ValueError):
        gpu_pkg.make_model("xdeepfm", F, K, [8], [])
    m.close()


def test_saturated_labels_and_targets_threshold(gpu_pkg):
    """targets are thresholded `label > 0` (DeepFM.scala:106): -1 / 0 are negatives, 1 / 5 positives."""
    B, F, K = 8, 3, 4
    index, w, bias, emb, mats, _ = make_inputs("fm", B, F, K)
    targets = np.array([-1, 0, 1, 5, 0.5, -0.5, 0, 1], np.float32)
    o32 = oracle_model("fm", F, K)
    m = pkg_model(gpu_pkg, "fm", F, K)
    gw, gb, ge = w.copy(), bias.copy(), emb.copy()
    rw, rb, re = w.copy(), bias.copy(), emb.copy()
    loss = m.backward(B, index, gw, gb, ge, None, targets)
    rloss = o32.backward(B, index, rw, rb, re, None, targets)
    assert abs(loss - rloss) <= 1e-5 * abs(rloss)
    assert_close(ge, re, what="dembedding")
    m.close()


@pytest.mark.parametrize("mode", [1, 0], ids=["tc3xtf32", "simt"])
@pytest.mark.parametrize("kind,B,kw", [
    ("deepfm", 700, dict(fc_dims=[400, 400, 400])),
    ("xdeepfm", 96, dict(fc_dims=[400, 400], cin_dims=[200, 200, 200])),
    ("xdeepfm", 40, dict(fc_dims=[64], cin_dims=[100, 50])),
    ("dcn", 300, dict(fc_dims=[400, 400], cross_depth=6)),
    ("pnn", 300, dict(fc_dims=[400, 400, 400])),
])
def test_baseline_sized_models(gpu_pkg, kind, B, kw, mode):
    """The layer sizes BASELINE.json names (MLP 400-400-400, CIN 200-200-200, 6 cross layers) at a
    batch the numpy oracle finishes in seconds."""
    from oracle import refport
    F, K = 39, 16
    rng = np.random.default_rng(B)
    cand = 8 * B if kind == "xdeepfm" else 4 * B
    n = cand * F
    index = np.repeat(np.arange(cand, dtype=np.int32), F)
    w = rng.uniform(-0.05, 0.05, n).astype(np.float32)
    bias = np.array([0.1], np.float32)
    emb = rng.uniform(-0.3, 0.3, n * K).astype(np.float32)
    fc, cin, depth = kw.get("fc_dims", ()), kw.get("cin_dims", ()), kw.get("cross_depth", 0)
    mats = gpu_pkg.synth.init_mats(3, refport.mats_size(kind, F, K, fc, cin, depth))
    o32 = refport.Model(kind, F, K, fc, cin, depth)
    o64 = refport.Model(kind, F, K, fc, cin, depth, np.float64)
    # CIN has 16 x sum(cinDims) ReLU units per sample: a 1e-4 margin leaves no candidates; 2e-5 is still
    # 10x the measured arithmetic error of either side
    index, w, emb, _ = away_from_kinks(o64, cand, F, K, index, w, bias, emb, mats,
                                       margin=2e-5 if kind == "xdeepfm" else 1e-4, keep=B)
    targets = (rng.uniform(0, 1, B) < 0.4).astype(np.float32)
    m = gpu_pkg.make_model(kind, F, K, fc, cin, depth)
    m.setGemmMode(mode)
    p = m.forward(B, index, w, bias, emb, mats)
    assert_close(p, o32.forward(B, index, w, bias, emb, mats), what="preds",
                 ref64=o64.forward(B, index, w, bias, emb, mats))
    gw, gb, ge, gm = w.copy(), bias.copy(), emb.copy(), mats.copy()
    loss = m.backward(B, index, gw, gb, ge, gm, targets)
    rw, rb, re, rm = w.copy(), bias.copy(), emb.copy(), mats.copy()
    rloss = o32.backward(B, index, rw, rb, re, rm, targets)
    dw, db, de, dm = (a.astype(np.float64) for a in (w, bias, emb, mats))
    dloss = o64.backward(B, index, dw, db, de, dm, targets)
    assert abs(loss - rloss) <= 1e-5 * abs(rloss) + 2 * abs(rloss - dloss)
    assert_close(ge, re, what="dembedding", ref64=de)
    assert_close(gm, rm, what="dmats", ref64=dm)
    assert_close(gw, rw, what="dweights", ref64=dw)
    m.close()
