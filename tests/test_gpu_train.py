"""Optimizer step, predict + AUC, and the text batch builder on the GPU."""
import numpy as np
import pytest

from common import assert_close
from oracle import refport

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("optim", ["sgd", "momentum", "adagrad", "adam"])
def test_optimizer_steps_match_oracle(gpu_pkg, optim):
    synth = gpu_pkg.synth
    F, K, rows, B = 39, 16, 39 * 64, 96
    kind, fc = "deepfm", [32, 16]
    model = gpu_pkg.make_model(kind, F, K, fc)
    table = gpu_pkg.EmbeddingTable(rows, K)
    table.init_uniform(42, -0.3, 0.3)
    ids = np.arange(rows)
    E = synth.table_rows(42, ids, K, -0.3, 0.3)
    W = synth.wtable_rows(42, ids, -0.3, 0.3)
    mats = synth.init_mats(7, model.getMatsSize())
    bias = np.array([0.1], np.float32)
    ps = gpu_pkg.ParRecModel(model, table)
    ps.setParams(bias, mats)
    o = refport.Model(kind, F, K, fc)
    stE, stW, stM, stB = {}, {}, {}, {}
    lr = {"sgd": 0.05, "momentum": 0.05, "adagrad": 0.003, "adam": 0.01}[optim]
    for step in range(1, 4):
        index, feats = synth.make_feats(5, step, B, F, rows)
        targets = synth.make_targets(5, feats, B, F)
        ps.optimize(feats, targets)
        ps.applyOptimizer(optim, lr)
        # oracle: same step on host copies
        emb, w = E[feats].reshape(-1).copy(), W[feats].copy()
        gb, gm = bias.copy(), mats.copy()
        o.backward(B, index, w, gb, emb, gm, targets)
        u, G = refport.make_embedding_grad(emb, feats, K)
        _, gw = refport.make_weights_grad(w, feats)
        sub = lambda st: {k: v[u] for k, v in st.items()}
        se, sw = sub(stE), sub(stW)
        E[u] = refport.optimizer_update(optim, E[u], G, se, lr, step=step)
        W[u] = refport.optimizer_update(optim, W[u], gw, sw, lr, step=step)
        for st, s_new, shape in ((stE, se, E.shape), (stW, sw, W.shape)):
            for k, v in s_new.items():
                st.setdefault(k, np.zeros(shape, np.float32))[u] = v
        mats = refport.optimizer_update(optim, mats, gm, stM, lr, step=step)
        bias = refport.optimizer_update(optim, bias, gb, stB, lr, step=step)
    gE, gW = table.read(0, rows)
    gbias, gmats = ps.getParams()
    # three chained steps: the gradients carry 1e-5-level differences which the optimizers amplify
    # (adam / adagrad divide by sqrt(v)); 1e-3 of the parameter scale separates right from wrong
    assert_close(gE, E, rtol=1e-3, what="table")
    assert_close(gW, W, rtol=1e-3, what="weights")
    assert_close(gmats, mats, rtol=1e-3, what="mats")
    assert abs(gbias[0] - bias[0]) <= 1e-3 * abs(bias[0]) + 1e-5
    # rows that never occurred are untouched, bit for bit
    seen = np.zeros(rows, bool)
    for step in range(1, 4):
        seen[synth.make_feats(5, step, B, F, rows)[1]] = True
    assert np.array_equal(gE[~seen], synth.table_rows(42, ids[~seen], K, -0.3, 0.3))
    with pytest.raises(ValueError):
        ps.applyOptimizer("lbfgs", 0.1)
    model.close(); table.close()


def test_training_reduces_loss_and_auc_matches_oracle(gpu_pkg):
    synth = gpu_pkg.synth
    F, K, rows, B = 39, 16, 39 * 128, 1024
    model = gpu_pkg.make_model("deepfm", F, K, [32, 16])
    table = gpu_pkg.EmbeddingTable(rows, K)
    table.init_uniform(42)
    ps = gpu_pkg.ParRecModel(model, table)
    ps.setParams(np.zeros(1, np.float32), synth.init_mats(42, model.getMatsSize()))
    batches = []
    for s in range(4):
        _, feats = synth.make_feats(1234, s, B, F, rows)
        batches.append((feats, synth.make_targets(1234, feats, B, F)))
    losses = []
    for epoch in range(6):
        tot = 0.0
        for f, t in batches:
            tot += ps.optimize(f, t)
            ps.applyOptimizer("adam", 0.01)
        losses.append(tot / (4 * B))
    assert losses[-1] < losses[0] - 0.02, losses
    preds = np.concatenate([ps.predict(f, B) for f, _ in batches])
    targets = np.concatenate([t for _, t in batches])
    a = gpu_pkg.metrics.auc(targets, preds)
    assert abs(a - refport.auc(targets, preds)) < 1e-12
    assert a > 0.6, a
    model.close(); table.close()


def test_checkpoint_round_trip(gpu_pkg, tmp_path):
    """save -> fresh model + table -> load gives bit-identical predictions."""
    synth = gpu_pkg.synth
    F, K, rows, B = 39, 16, 39 * 64, 64
    fc = [32, 16]
    def fresh():
        m = gpu_pkg.make_model("deepfm", F, K, fc)
        t = gpu_pkg.EmbeddingTable(rows, K)
        return gpu_pkg.ParRecModel(m, t)
    ps = fresh()
    ps.table.init_uniform(42, -0.3, 0.3)
    ps.setParams(np.array([0.1], np.float32), synth.init_mats(7, ps.model.getMatsSize()))
    _, feats = synth.make_feats(5, 1, B, F, rows)
    targets = synth.make_targets(5, feats, B, F)
    ps.optimize(feats, targets)
    ps.applyOptimizer("adam", 0.01)          # move every kind of parameter off its initial value
    want = ps.predict(feats, B)
    path = str(tmp_path / "ckpt.npz")
    ps.save(path, chunk_rows=1000)           # several chunks
    ps2 = fresh()
    ps2.load(path, chunk_rows=700)
    got = ps2.predict(feats, B)
    assert np.array_equal(got, want)
    bad = gpu_pkg.ParRecModel(gpu_pkg.make_model("deepfm", F, K, [8]), gpu_pkg.EmbeddingTable(rows, K))
    with pytest.raises(ValueError):
        bad.load(path)


def test_staged_step_equals_host_step(gpu_pkg):
    """b200rec_stage_batch + b200rec_step_staged (input prefetch) give the losses and gradients of
    b200rec_step on the same batches; the staging depth and order are enforced."""
    synth = gpu_pkg.synth
    F, K, rows, B = 39, 16, 39 * 64, 128
    fc = [32, 16]
    def fresh():
        m = gpu_pkg.make_model("deepfm", F, K, fc)
        t = gpu_pkg.EmbeddingTable(rows, K)
        t.init_uniform(42, -0.3, 0.3)
        ps = gpu_pkg.ParRecModel(m, t)
        ps.setParams(np.array([0.1], np.float32), synth.init_mats(7, m.getMatsSize()))
        return ps
    batches = []
    for s_ in range(5):
        _, f = synth.make_feats(5, s_, B, F, rows)
        batches.append((f, synth.make_targets(5, f, B, F)))
    a, b = fresh(), fresh()
    want = []
    for f, t in batches:
        want.append(a.optimize(f, t))
        a.applyOptimizer("sgd", 0.05)
    with pytest.raises(RuntimeError):
        b.optimizeStaged()                       # nothing staged
    got = []
    b.stage(*batches[0])
    for i in range(len(batches)):
        if i + 1 < len(batches):
            b.stage(*batches[i + 1])             # copy of the next batch under this step
        got.append(b.optimizeStaged())
        b.applyOptimizer("sgd", 0.05)
    assert got == want
    ra, rb = a.stepResults(), b.stepResults()
    assert np.array_equal(ra["unique"], rb["unique"]) and np.array_equal(ra["emb_grad"], rb["emb_grad"])
    assert np.array_equal(ra["mats_grad"], rb["mats_grad"])
    # the same losses through the asynchronous pair, read one step late
    c = fresh()
    got_async = []
    c.stage(*batches[0])
    for i in range(len(batches)):
        if i + 1 < len(batches):
            c.stage(*batches[i + 1])
        c.optimizeStagedAsync()
        if i > 0:
            got_async.append(c.waitLoss())
        c.applyOptimizer("sgd", 0.05)
    got_async.append(c.waitLoss())
    assert got_async == want
    with pytest.raises(RuntimeError):
        c.waitLoss()                             # nothing in flight
    b.stage(*batches[0]); b.stage(*batches[1])
    with pytest.raises(RuntimeError):
        b.stage(*batches[2])                     # two staged already
