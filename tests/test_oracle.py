"""CPU tests that pin the oracle (oracle/refport.py).

The reference ships no tests or golden vectors and cannot run here, so the
oracle is pinned by (1) the committed fixtures produced by an independent fp64 autograd
implementation of the papers' formulas (tests/golden/make_golden.py), (2) that implementation run
live on other shapes, (3) hand-computed cases, (4) finite differences, (5) algebraic properties.
"""
import glob
import os

import numpy as np
import pytest

import independent
from common import CONFIGS, kind_of, make_inputs, oracle_model
from oracle import refport

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def _run_oracle(name, F, K, g, dtype):
    o = oracle_model(name, F, K, dtype)
    get = lambda k: g[k].astype(dtype) if k in g else None
    B = g["targets"].shape[0]
    w, b, e, m = get("weights"), get("bias"), get("embedding"), get("mats")
    pred = o.forward(B, g["index"], w, b, e, m)
    loss = o.backward(B, g["index"], w, b, e, m, g["targets"])
    return pred, loss, w, b, e, m


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_golden(path):
    base = os.path.basename(path)[:-4]
    name, B, F, K = base.split("_")
    F, K = int(F[1:]), int(K[1:])
    g = np.load(path)
    # fp64 twin: agreement to rounding noise of fp64
    pred, loss, gw, gb, ge, gm = _run_oracle(name, F, K, g, np.float64)
    np.testing.assert_allclose(pred, g["pred"], rtol=1e-11, atol=1e-13)
    assert abs(loss - float(g["loss"])) < 1e-11
    np.testing.assert_allclose(gw, g["gw"], rtol=1e-9, atol=1e-14)
    np.testing.assert_allclose(gb, g["gb"], rtol=1e-9, atol=1e-14)
    if ge is not None:
        np.testing.assert_allclose(ge, g["ge"], rtol=1e-8, atol=1e-12 * np.abs(g["ge"]).max() + 1e-13)
    if gm is not None:
        np.testing.assert_allclose(gm, g["gm"], rtol=1e-8, atol=1e-12 * np.abs(g["gm"]).max() + 1e-13)
    # fp32 (the reference's precision): within 1e-5 of the tensor scale
    pred, loss, gw, gb, ge, gm = _run_oracle(name, F, K, g, np.float32)
    for got, key in ((pred, "pred"), (gw, "gw"), (gb, "gb"), (ge, "ge"), (gm, "gm")):
        if got is None:
            continue
        want = g[key]
        assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max() + 1e-12, key
    assert abs(loss - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))


REF_DUMPS = sorted(d for d in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "reference", "*"))
                   if os.path.exists(os.path.join(d, "out_loss.bin")))


@pytest.mark.skipif(not REF_DUMPS, reason="no fixtures dumped from the real reference yet (tests/golden/README.md): "
                                          "parity stays UNPINNED by the reference")
@pytest.mark.parametrize("d", REF_DUMPS or [None], ids=[os.path.basename(d) for d in REF_DUMPS] or ["none"])
def test_oracle_matches_reference_dump(d):
    """The oracle (fp32) against outputs of yaochitc/recommendation-models itself, dumped on a JVM by
    scala/DumpFixtures.scala: forward preds, loss and the four in/out-aliased gradient buffers."""
    name, B, F, K = os.path.basename(d).split("_")
    B, F, K = int(B[1:]), int(F[1:]), int(K[1:])
    rd = lambda n, dt="<f4": np.fromfile(os.path.join(d, n), dt) if os.path.exists(os.path.join(d, n)) else None
    index = rd("in_index.bin", "<i4")
    w, b, e, m, t = (rd(f"in_{k}.bin") for k in ("weights", "bias", "embedding", "mats", "targets"))
    o = oracle_model(name, F, K)
    close = lambda got, want, what: np.testing.assert_array_less(
        np.abs(np.asarray(got, np.float64) - want).max(), 1e-5 * np.abs(want).max() + 1e-12, err_msg=what)
    close(o.forward(B, index, w.copy(), b.copy(), None if e is None else e.copy(), None if m is None else m.copy()),
          rd("out_pred.bin"), "pred")
    loss = o.backward(B, index, w, b, e, m, t)
    assert abs(loss - float(rd("out_loss.bin")[0])) <= 1e-5 * abs(loss)
    for got, n in ((w, "weights"), (b, "bias"), (e, "embedding"), (m, "mats")):
        if got is not None:
            close(got, rd(f"out_{n}.bin"), n)


@pytest.mark.parametrize("name", list(CONFIGS))
def test_oracle_matches_independent_live(name):
    B, F, K = 9, 7, 6
    cfg = CONFIGS[name]
    index, w, bias, emb, mats, targets = make_inputs(name, B, F, K, seed=99)
    r = independent.run(kind_of(name), B, F, K, index, w, bias, emb, mats, targets,
                        cfg.get("fc_dims", ()), cfg.get("cin_dims", ()), cfg.get("cross_depth", 0))
    o = oracle_model(name, F, K, np.float64)
    c = lambda a: None if a is None else a.astype(np.float64)
    gw, gb, ge, gm = c(w), c(bias), c(emb), c(mats)
    np.testing.assert_allclose(o.forward(B, index, gw, gb, ge, gm), r["pred"], rtol=1e-11)
    loss = o.backward(B, index, gw, gb, ge, gm, targets)
    assert abs(loss - r["loss"]) < 1e-11
    np.testing.assert_allclose(gw, r["gw"], rtol=1e-9, atol=1e-14)
    if ge is not None:
        np.testing.assert_allclose(ge, r["ge"], rtol=1e-8, atol=1e-13)
    if gm is not None:
        np.testing.assert_allclose(gm, r["gm"], rtol=1e-8, atol=1e-13)


def test_hand_computed_fm():
    """B=1, F=2, K=2: v1=(1,2), v2=(3,4), w=(0.5,-0.25), bias=0.1.
    S=(4,6); S^2=(16,36); Q=(10,20); second = 0.5*mean(6,16) = 5.5 (mean over K, SURVEY B-3);
    first = 0.25; logit = 5.85."""
    o = refport.Model("fm", 2, 2)
    p = o.forward(1, [0, 0], np.array([0.5, -0.25], np.float32), np.array([0.1], np.float32),
                  np.array([1, 2, 3, 4], np.float32))
    assert abs(p[0] - 1 / (1 + np.exp(-5.85))) < 1e-6
    # gradient: dlogit = p - 1 (target 1, B=1); dv = dlogit/K * (S - v)
    w = np.array([0.5, -0.25], np.float32); b = np.array([0.1], np.float32)
    e = np.array([1, 2, 3, 4], np.float32)
    o.backward(1, [0, 0], w, b, e, None, np.array([1.0], np.float32))
    dl = p[0] - 1.0
    np.testing.assert_allclose(e, dl / 2 * np.array([3, 4, 1, 2]), rtol=1e-5)
    np.testing.assert_allclose(w, [dl, dl], rtol=1e-5)
    np.testing.assert_allclose(b, [dl], rtol=1e-5)


def test_hand_computed_cin_single_layer():
    """One sample, F=2, K=1, one CIN unit: Z = [x0_0*x0_0, x0_0*x0_1, x0_1*x0_0, x0_1*x0_1]
    (i-major: CINEncoder.scala:152 MM(x0, x0^T) then Reshape to F*H)."""
    x = np.array([2.0, 3.0], np.float32)               # [B=1,F=2,K=1]
    fc = [1]
    # mats: DNN W(1x2), b(1); CIN W(1x4), b(1); W_out(1 x (1+1))
    mats = np.array([0, 0, 0,   1, 10, 100, 1000, 0.5,   1.0, 0.0], np.float32)
    out, _ = refport.cin_fwd(x, mats, 1, 2, 1, fc, [1])
    assert out[0, 0] == 4 + 60 + 600 + 9000 + 0.5


def test_hand_computed_cross_and_pnn():
    # DCN, D=2, depth 1: x1 = x0*(x0.w) + x0 + c
    x = np.array([1.0, 2.0], np.float32)
    mats = np.array([0.5, 0.25,  0.1,   0, 0, 0,   1, 1, 0], np.float32)  # w, c, DNN(1x2 + b), W_out(2+1)
    out, sv = refport.cross_fwd(x, mats, 1, 2, 1, 1, [1])
    s = 0.5 + 0.5
    np.testing.assert_allclose(sv["xs"][-1][0], [1 * s + 1 + 0.1, 2 * s + 2 + 0.1], rtol=1e-6)
    # PNN pairs are lexicographic i<j (ProductEncoder.scala:110-120)
    r, c = refport.pnn_pairs(4)
    assert list(zip(r, c)) == [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]


def test_mats_sizes_match_survey():
    """DeepFM 39x16, 400^3 -> 571 201 floats; xDeepFM 200^3 + 400^3 -> 3 996 600."""
    assert refport.mats_len("deepfm", 39, 16, [400, 400, 400]) == 571201
    assert refport.mats_len("xdeepfm", 39, 16, [400, 400, 400], [200, 200, 200]) == 3996600
    assert refport.mats_size("deepfm", 2, 2, [3]) == [4, 3, 3, 1, 3, 1, 1, 1]


@pytest.mark.parametrize("name", ["fm", "deepfm", "xdeepfm", "dcn", "pnn"])
def test_finite_differences(name):
    B, F, K = 3, 4, 3
    index, w, bias, emb, mats, targets = make_inputs(name, B, F, K, seed=3)
    o = oracle_model(name, F, K, np.float64)
    c = lambda a: None if a is None else a.astype(np.float64)

    def loss_of(w_, e_, m_):
        p = o.forward(B, index, w_, c(bias), e_, m_)
        t = (targets > 0).astype(np.float64)
        return float(-(t * np.log(p) + (1 - t) * np.log(1 - p)).mean())

    gw, gb, ge, gm = c(w), c(bias), c(emb), c(mats)
    o.backward(B, index, gw, gb, ge, gm, targets)
    rng = np.random.default_rng(0)
    h = 1e-6
    for buf, grad, which in ((emb, ge, 1), (mats, gm, 2)):
        if buf is None:
            continue
        for j in rng.choice(buf.size, size=min(12, buf.size), replace=False):
            args = [c(w), c(emb), c(mats)]
            args[which][j] += h
            up = loss_of(*args)
            args[which][j] -= 2 * h
            dn = loss_of(*args)
            fd = (up - dn) / (2 * h)
            assert abs(fd - grad[j]) <= 1e-6 + 1e-4 * abs(fd), (name, which, j, fd, grad[j])


def test_gather_scatter_adjoint_and_order():
    rng = np.random.default_rng(0)
    rows, K, n = 20, 4, 500
    feats = rng.integers(0, rows, n).astype(np.int32)
    T = rng.standard_normal((rows, K))
    g = rng.standard_normal((n, K))
    ids, G = refport.make_embedding_grad(g.reshape(-1), feats, K)
    # <gather(T), g> == <T, scatter_add(g)>
    lhs = (refport.make_embeddings(T, feats).reshape(n, K) * g).sum()
    full = np.zeros((rows, K)); full[ids] = G
    assert abs(lhs - (T * full).sum()) < 1e-9
    # sequential fp32 order: equals an explicit i-ascending loop bit for bit
    g32 = g.astype(np.float32)
    ids, G32 = refport.make_embedding_grad(g32.reshape(-1), feats, K)
    acc = {}
    for i in range(n):
        acc[feats[i]] = acc.get(feats[i], np.zeros(K, np.float32)) + g32[i]
    for u, row in zip(ids, G32):
        assert np.array_equal(row, acc[u])
    assert np.array_equal(ids, np.unique(feats))


def test_scatter_module_semantics():
    out = refport.scatter_update_output(np.array([1, 2, 3, 4], np.float32), [1, 0, 1, 1], 3)
    assert out.tolist() == [[2], [8], [0]]
    with pytest.raises(ValueError):
        refport.scatter_update_output(np.ones(2, np.float32), [0, 3], 3)
    gi = refport.scatter_update_grad_input(np.array([[10], [20], [30]], np.float32), np.array([2, 0]), 3)
    assert gi.tolist() == [[30], [10]]


def test_fm_field_permutation_invariance():
    B, F, K = 5, 6, 4
    rng = np.random.default_rng(1)
    e = rng.standard_normal((B, F, K))
    perm = rng.permutation(F)
    a = refport.second_order_fwd(e.reshape(-1), B, F, K)
    b = refport.second_order_fwd(e[:, perm].reshape(-1), B, F, K)
    np.testing.assert_allclose(a, b, rtol=1e-12)


def test_auc():
    assert refport.auc([0, 0, 1, 1], [0.1, 0.4, 0.35, 0.8]) == 0.75
    assert refport.auc([1, 0], [0.9, 0.1]) == 1.0
    assert np.isnan(refport.auc([1, 1], [0.9, 0.1]))


def test_synth_generator_properties(pkg):
    synth = pkg.synth
    B, F, rows = 512, 39, 39 * 4096
    index, feats = synth.make_feats(1234, 0, B, F, rows)
    assert index.tolist() == np.repeat(np.arange(B), F).tolist()
    off, voc = synth.field_layout(rows, F)
    f = feats.reshape(B, F)
    assert ((f >= off[None]) & (f < (off + voc)[None])).all()
    # power law: the hottest id of a field takes a visible share; many ids appear once
    _, counts = np.unique(feats, return_counts=True)
    assert counts.max() > B // 20 and (counts == 1).sum() > len(counts) // 2
    # deterministic & step-dependent
    assert np.array_equal(feats, synth.make_feats(1234, 0, B, F, rows)[1])
    assert not np.array_equal(feats, synth.make_feats(1234, 1, B, F, rows)[1])
    t = synth.make_targets(1234, feats, B, F)
    assert 0.2 < t.mean() < 0.8
