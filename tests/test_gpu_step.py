"""GPU parity of the resident step (lookup -> forward -> backward -> dedup scatter-add) against
the oracle's gather + Model.backward + make*Grad, on synthetic Criteo-shaped batches."""
import numpy as np
import pytest

from common import CONFIGS, assert_close, away_from_kinks, kind_of
from oracle import refport

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(CONFIGS))
@pytest.mark.parametrize("B", [1, 96, 257])
def test_resident_step(gpu_pkg, name, B):
    synth = gpu_pkg.synth
    cfg = CONFIGS[name]
    kind = kind_of(name)
    F, K, rows = 39, 16, 39 * 300
    model = gpu_pkg.make_model(kind, F, K, cfg.get("fc_dims", ()), cfg.get("cin_dims", ()), cfg.get("cross_depth", 0))
    table = gpu_pkg.EmbeddingTable(rows, K if kind != "lr" else 0)
    table.init_uniform(42, -0.3, 0.3)
    mats = synth.init_mats(7, model.getMatsSize())
    bias = np.array([0.1], np.float32)
    # 4B candidate samples; keep B whose ReLU pre-activations stay away from the kinks (common.py)
    _, cand = synth.make_feats(1234, 3, 4 * B, F, rows)
    o64 = refport.Model(kind, F, K, cfg.get("fc_dims", ()), cfg.get("cin_dims", ()), cfg.get("cross_depth", 0), np.float64)
    ce = synth.table_rows(42, cand, K, -0.3, 0.3).reshape(-1) if kind != "lr" else None
    cw = synth.wtable_rows(42, cand, -0.3, 0.3)
    cidx = np.repeat(np.arange(4 * B, dtype=np.int32), F)
    index, _, _, ids = away_from_kinks(o64, 4 * B, F, K, cidx, cw, bias, ce, mats if mats.size else None, keep=B)
    feats = np.ascontiguousarray(cand.reshape(4 * B, F)[ids].reshape(-1))
    targets = synth.make_targets(1234, feats, B, F)
    ps = gpu_pkg.ParRecModel(model, table)
    ps.setParams(bias, mats)
    # predict
    preds = ps.predict(feats, B)
    emb = synth.table_rows(42, feats, K, -0.3, 0.3).reshape(-1) if kind != "lr" else None
    w = synth.wtable_rows(42, feats, -0.3, 0.3)
    o = refport.Model(kind, F, K, cfg.get("fc_dims", ()), cfg.get("cin_dims", ()), cfg.get("cross_depth", 0))
    assert_close(preds, o.forward(B, index, w, bias, emb, mats), what="preds",
                 ref64=o64.forward(B, index, w, bias, emb, mats))
    # optimize
    loss = ps.optimize(feats, targets) / B
    res = ps.stepResults()
    cp = lambda a: None if a is None else a.copy()
    oe, ow, ob, om = cp(emb), cp(w), cp(bias), cp(mats)
    oloss = o.backward(B, index, ow, ob, oe, om if om is not None and om.size else None, targets)
    c64 = lambda a: None if a is None else a.astype(np.float64)
    de, dw, db, dm = c64(emb), c64(w), c64(bias), c64(mats)
    o64.backward(B, index, dw, db, de, dm if dm is not None and dm.size else None, targets)
    assert abs(loss - oloss) <= 1e-5 * abs(oloss)
    ids, gw = refport.make_weights_grad(ow, feats)
    _, gw64 = refport.make_weights_grad(dw, feats)
    assert np.array_equal(res["unique"], ids)
    assert_close(res["w_grad"], gw, what="w_grad", ref64=gw64)
    if kind != "lr":
        _, G = refport.make_embedding_grad(oe, feats, K)
        _, G64 = refport.make_embedding_grad(de, feats, K)
        assert_close(res["emb_grad"], G, what="emb_grad", ref64=G64)
    assert abs(res["bias_grad"] - ob[0]) <= 1e-5 * abs(ob[0]) + 2 * abs(ob[0] - db[0]) + 1e-6 * np.abs(ow).sum() / F
    if mats is not None and mats.size:
        assert_close(res["mats_grad"], om, what="mats_grad", ref64=dm)
    model.close()
    table.close()


def test_step_needs_params_and_valid_ids(gpu_pkg):
    F, K, rows, B = 5, 8, 100, 4
    model = gpu_pkg.make_model("fm", F, K)
    table = gpu_pkg.EmbeddingTable(rows, K)
    ps = gpu_pkg.ParRecModel(model, table)
    feats = np.zeros(B * F, np.int32)
    with pytest.raises(gpu_pkg.B200RecError):
        ps.optimize(feats, np.zeros(B, np.float32))
    ps.setParams(np.zeros(1, np.float32))
    feats[3] = rows
    with pytest.raises(ValueError):
        ps.optimize(feats, np.zeros(B, np.float32))
    model.close(); table.close()


@pytest.mark.parametrize("name", ["fm", "deepfm"])
def test_step_graph_replay(gpu_pkg, name):
    """The resident step is captured into a CUDA graph after one eager warm-up; replays with new ids
    must give the results of a fresh eager step (and of the oracle)."""
    synth = gpu_pkg.synth
    cfg = CONFIGS[name]
    F, K, rows, B = 39, 16, 39 * 300, 128
    model = gpu_pkg.make_model(name, F, K, cfg.get("fc_dims", ()))
    table = gpu_pkg.EmbeddingTable(rows, K)
    table.init_uniform(42, -0.3, 0.3)
    mats = synth.init_mats(7, model.getMatsSize())
    bias = np.array([0.1], np.float32)
    ps = gpu_pkg.ParRecModel(model, table)
    ps.setParams(bias, mats)
    o = refport.Model(name, F, K, cfg.get("fc_dims", ()))
    o64 = refport.Model(name, F, K, cfg.get("fc_dims", ()), dtype=np.float64)
    for step in range(4):     # step 0 eager, step 1 captured, steps 2-3 replayed
        _, cand = synth.make_feats(99, step, 4 * B, F, rows)
        ce = synth.table_rows(42, cand, K, -0.3, 0.3).reshape(-1)
        cw = synth.wtable_rows(42, cand, -0.3, 0.3)
        cidx = np.repeat(np.arange(4 * B, dtype=np.int32), F)
        index, _, _, ids = away_from_kinks(o64, 4 * B, F, K, cidx, cw, bias, ce, mats if mats.size else None, keep=B)
        feats = np.ascontiguousarray(cand.reshape(4 * B, F)[ids].reshape(-1))
        targets = synth.make_targets(99, feats, B, F)
        loss = ps.optimize(feats, targets) / B
        res = ps.stepResults()
        emb = synth.table_rows(42, feats, K, -0.3, 0.3).reshape(-1)
        w = synth.wtable_rows(42, feats, -0.3, 0.3)
        ob, om = bias.copy(), (mats.copy() if mats.size else None)
        oloss = o.backward(B, index, w, ob, emb, om, targets)
        assert abs(loss - oloss) <= 1e-5 * abs(oloss), step
        uids, G = refport.make_embedding_grad(emb, feats, K)
        assert np.array_equal(res["unique"], uids), step
        assert_close(res["emb_grad"], G, what=f"emb_grad step {step}", rtol=2e-5)
        if om is not None:
            assert_close(res["mats_grad"], om, what=f"mats_grad step {step}", rtol=2e-5)
    model.close(); table.close()


@pytest.mark.parametrize("name", ["fm", "deepfm", "xdeepfm"])
def test_fused_scatter_is_bit_identical(gpu_pkg, name):
    """b200rec_model_set_fused_scatter: the per-nnz gradient computed inside the segment reduce must give
    the same bits as the two-kernel path (same arithmetic, same non-zero order), with and without the
    per-nnz gradients kept."""
    synth = gpu_pkg.synth
    cfg = CONFIGS[name]
    F, K, rows, B = 39, 16, 39 * 40, 700          # small vocabulary: hot ids far beyond 64 rows, many 5..64
    res = []
    for fused, keep in ((0, 0), (1, 0), (1, 1)):
        model = gpu_pkg.make_model(name, F, K, cfg.get("fc_dims", ()), cfg.get("cin_dims", ()))
        gpu_pkg._lib.check(gpu_pkg.lib().b200rec_model_set_fused_scatter(model.handle, fused, keep))
        table = gpu_pkg.EmbeddingTable(rows, K)
        table.init_uniform(42, -0.3, 0.3)
        ps = gpu_pkg.ParRecModel(model, table)
        ps.setParams(np.array([0.1], np.float32), synth.init_mats(7, model.getMatsSize()))
        _, feats = synth.make_feats(5, 1, B, F, rows)
        targets = synth.make_targets(5, feats, B, F)
        for _ in range(3):                        # eager, captured, replayed
            ps.optimize(feats, targets)
        res.append(ps.stepResults())
        model.close(); table.close()
    for r in res[1:]:
        assert np.array_equal(r["unique"], res[0]["unique"])
        assert np.array_equal(r["emb_grad"], res[0]["emb_grad"])
        assert np.array_equal(r["w_grad"], res[0]["w_grad"])


@pytest.mark.parametrize("rows", [200_000, 5_000_000, 40_000_000, 100_000_000])
def test_resident_step_key_widths(gpu_pkg, rows):
    """The sort half sizes its passes from the table: 18, 23, 26 and 27 key bits (3-4 passes of 8-bit digits; with
    B200REC_SORT_DB9=1 the 9-bit digit kernels take the 18 / 26 / 27-bit cases).  Ids spread over the whole table
    so that every digit varies; distinct ids exact, summed gradients against the oracle."""
    synth = gpu_pkg.synth
    F, K, B = 39, 4, 512
    model = gpu_pkg.make_model("fm", F, K, (), (), 0)
    table = gpu_pkg.EmbeddingTable(rows, K)
    table.init_uniform(42, -0.3, 0.3)
    rng = np.random.default_rng(rows % 9973)
    feats = rng.integers(0, rows, B * F).astype(np.int32)
    feats[::7] = feats[0]                       # one hot id, and ...
    feats[1::11] = rows - 1                     # ... the largest key
    targets = (rng.random(B) < 0.5).astype(np.float32)
    bias = np.array([0.05], np.float32)
    ps = gpu_pkg.ParRecModel(model, table)
    ps.setParams(bias, np.zeros(0, np.float32))
    ps.optimize(feats, targets)
    res = ps.stepResults()
    emb = synth.table_rows(42, feats, K, -0.3, 0.3).reshape(-1)
    w = synth.wtable_rows(42, feats, -0.3, 0.3)
    index = np.repeat(np.arange(B, dtype=np.int32), F)
    o = refport.Model("fm", F, K, (), (), 0)
    o64 = refport.Model("fm", F, K, (), (), 0, np.float64)
    oe, ow, ob = emb.copy(), w.copy(), bias.copy()
    o.backward(B, index, ow, ob, oe, None, targets)
    de, dw, db = emb.astype(np.float64), w.astype(np.float64), bias.astype(np.float64)
    o64.backward(B, index, dw, db, de, None, targets)
    ids, gw = refport.make_weights_grad(ow, feats)
    _, gw64 = refport.make_weights_grad(dw, feats)
    assert np.array_equal(res["unique"], ids)
    assert_close(res["w_grad"], gw, what="w_grad", ref64=gw64)
    _, G = refport.make_embedding_grad(oe, feats, K)
    _, G64 = refport.make_embedding_grad(de, feats, K)
    assert_close(res["emb_grad"], G, what="emb_grad", ref64=G64)
