"""GPU regression tests for the review findings of round 1: a replayed CUDA graph after a workspace
reallocation, out-of-range ids in the optimizer kernels, the table status word."""
import ctypes as C

import numpy as np
import pytest

from common import assert_close
from oracle import refport

pytestmark = pytest.mark.gpu


def _oracle_step(synth, kind, fc, F, K, feats, targets, bias, mats, lo=-0.3, hi=0.3):
    B = targets.size
    index = np.repeat(np.arange(B, dtype=np.int32), F)
    emb = synth.table_rows(42, feats, K, lo, hi).reshape(-1)
    w = synth.wtable_rows(42, feats, lo, hi)
    ob, om = bias.copy(), (mats.copy() if mats.size else None)
    loss = refport.Model(kind, F, K, fc).backward(B, index, w, ob, emb, om, targets)
    ids, G = refport.make_embedding_grad(emb, feats, K)
    return loss, ids, G, om


@pytest.mark.parametrize("name,fc", [("fm", []), ("deepfm", [48, 24])])
def test_graph_not_replayed_after_reallocation(gpu_pkg, name, fc):
    """step(B) captures a graph; predict(4B) grows the handle's workspaces (cudaFree + cudaMalloc); the
    next step(B) must not replay the stale graph.  Results are checked against the oracle each time."""
    synth = gpu_pkg.synth
    F, K, rows, B = 39, 16, 39 * 300, 128
    model = gpu_pkg.make_model(name, F, K, fc)
    table = gpu_pkg.EmbeddingTable(rows, K)
    table.init_uniform(42, -0.3, 0.3)
    mats = synth.init_mats(7, model.getMatsSize())
    bias = np.array([0.1], np.float32)
    ps = gpu_pkg.ParRecModel(model, table)
    ps.setParams(bias, mats)
    ep = C.c_int64(0)

    def step(i):
        _, feats = synth.make_feats(77, i, B, F, rows)
        targets = synth.make_targets(77, feats, B, F)
        loss = ps.optimize(feats, targets) / B
        res = ps.stepResults()
        oloss, ids, G, om = _oracle_step(synth, name, fc, F, K, feats, targets, bias, mats)
        assert abs(loss - oloss) <= 2e-5 * abs(oloss), i
        assert np.array_equal(res["unique"], ids), i
        assert_close(res["emb_grad"], G, what=f"emb_grad step {i}", rtol=1e-4)   # unfiltered batch: kinks allowed for

    for i in range(3):                 # eager, captured, replayed
        step(i)
    gpu_pkg._lib.check(gpu_pkg.lib().b200rec_alloc_epoch(C.byref(ep)))
    e0 = ep.value
    _, big = synth.make_feats(78, 0, 4 * B, F, rows)
    preds = ps.predict(big, 4 * B)     # larger batch: first / second / X / acts / scratch grow
    assert np.all(np.isfinite(preds))
    gpu_pkg._lib.check(gpu_pkg.lib().b200rec_alloc_epoch(C.byref(ep)))
    assert ep.value > e0, "the larger predict batch was expected to reallocate workspaces"
    for i in range(3, 7):              # eager again, re-captured, replayed
        step(i)
    model.close(); table.close()


@pytest.mark.parametrize("optim", ["sgd", "adam"])
def test_optimizer_kernels_skip_out_of_range_ids(gpu_pkg, optim):
    """An id outside [0, rows) in `unique` must never become a device write outside the table: the row
    is skipped, the others are updated, and the table's status word reports B200REC_ERR_INDEX."""
    import torch
    lib, L = gpu_pkg.lib(), gpu_pkg._lib
    rows, K = 64, 16
    table = gpu_pkg.EmbeddingTable(rows, K)
    table.init_uniform(3, -0.1, 0.1)
    e0, w0 = table.read(0, rows)
    uniq = torch.tensor([2, 7, rows + 5, -3, 9], dtype=torch.int32, device="cuda")
    n = torch.tensor([5], dtype=torch.int32, device="cuda")
    G = torch.ones(5 * K, dtype=torch.float32, device="cuda")
    gw = torch.ones(5, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    if optim == "sgd":
        L.check(lib.b200rec_table_apply_sgd_dev(table.handle, 5, n.data_ptr(), uniq.data_ptr(), G.data_ptr(),
                                                gw.data_ptr(), 0.5, None))
    else:
        L.check(lib.b200rec_table_apply_optimizer_dev(table.handle, L.OPTIMIZERS["adam"], 0.5, 0.99, 0.9, 1, 5,
                                                      n.data_ptr(), uniq.data_ptr(), G.data_ptr(), gw.data_ptr(), None))
    with pytest.raises(ValueError):
        L.check(lib.b200rec_table_status(table.handle, 1, None))
    L.check(lib.b200rec_table_status(table.handle, 0, None))      # cleared by the reset above
    e1, w1 = table.read(0, rows)
    touched = np.zeros(rows, bool)
    touched[[2, 7, 9]] = True
    assert np.array_equal(e1[~touched], e0[~touched]) and np.array_equal(w1[~touched], w0[~touched])
    assert np.all(e1[touched] < e0[touched]) and np.all(w1[touched] < w0[touched])
    table.close()


def test_adam_step_counter_on_device_matches_host_step(gpu_pkg):
    """b200rec_*_apply_optimizer_stepdev_dev (update count read from device memory: what a replayed graph
    needs) == the host-step form, bit for bit, for steps 1..3."""
    import torch
    lib, L = gpu_pkg.lib(), gpu_pkg._lib
    rows, K, U = 128, 16, 20
    tabs = [gpu_pkg.EmbeddingTable(rows, K) for _ in range(2)]
    for t in tabs:
        t.init_uniform(5, -0.1, 0.1)
    uniq = torch.arange(3, 3 + U, dtype=torch.int32, device="cuda")
    n = torch.tensor([U], dtype=torch.int32, device="cuda")
    ctr = torch.zeros(1, dtype=torch.int32, device="cuda")
    rng = np.random.default_rng(1)
    for step in (1, 2, 3):
        G = torch.from_numpy(rng.standard_normal(U * K).astype(np.float32)).cuda()
        gw = torch.from_numpy(rng.standard_normal(U).astype(np.float32)).cuda()
        ctr.fill_(step)
        torch.cuda.synchronize()
        L.check(lib.b200rec_table_apply_optimizer_dev(tabs[0].handle, 3, 0.01, 0.99, 0.9, step, U, n.data_ptr(),
                                                      uniq.data_ptr(), G.data_ptr(), gw.data_ptr(), None))
        L.check(lib.b200rec_table_apply_optimizer_stepdev_dev(tabs[1].handle, 3, 0.01, 0.99, 0.9, ctr.data_ptr(), U,
                                                              n.data_ptr(), uniq.data_ptr(), G.data_ptr(),
                                                              gw.data_ptr(), None))
        for t in tabs:
            L.check(lib.b200rec_table_status(t.handle, 0, None))
    a, b = tabs[0].read(0, rows), tabs[1].read(0, rows)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    for t in tabs:
        t.close()
