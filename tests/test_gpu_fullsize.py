"""Parity at the BENCHMARK configuration and on ordinary (unfiltered) batches.

* B = 8192, the batch every number of bench.py is quoted on (BASELINE configs[1] / [2]: 64 row tiles,
  multi-wave split-K weight gradients, 1024 CIN row tiles), through the resident step, against the
  oracle run in sample chunks: samples are independent, so the oracle's backward over a chunk of Bc
  samples scaled by Bc / B (the loss is a batch mean, DeepFM.scala:105-117) sums to the full-batch result.
  No sample is filtered away from the ReLU kinks here.
* ReLU kinks: a pre-activation closer to zero than the arithmetic error may legitimately come out on
  either side in two correct fp32 paths, and one flipped unit changes ITS sample's gradient by O(1).
  Nothing is filtered here; instead the comparison is per sample (_kink_aware): every sample must meet
  the 1e-5 bar unless one of its pre-activations is kink-adjacent, the exceptions are counted and
  reported, and the summed weight gradient is allowed exactly their contribution on top of the bar.
"""
import json
import os

import numpy as np
import pytest

from common import _report, assert_close
from oracle import refport

pytestmark = pytest.mark.gpu
F, K = 39, 16


def _near_kink(o, B):
    """Per sample: the distance of its closest ReLU pre-activation to zero, relative to the layer's scale."""
    near = np.full(B, np.inf)
    for z in o.relu_preactivations(B):
        near = np.minimum(near, np.abs(z).min(axis=1) / np.abs(z).max())
    return near


def _kink_aware(name, B, per_sample, near, extra_of, mats_got, mats_want, mats64=None, rtol=1e-5):
    """The parity bar on a batch that was NOT filtered away from the ReLU kinks.

    per_sample: [(what, got[B, ...], want[B, ...])] -- tensors with one slice per sample.  A sample may
    exceed the bar only if one of its ReLU pre-activations is kink-adjacent (|z| < 1e-4 of the layer's
    scale: two correct fp32 paths may then disagree about the unit's sign, which changes that sample's
    gradient by O(1)); any violation elsewhere is an arithmetic bug and fails.  The violating samples
    are counted and reported, and must be fewer than the samples within 1e-5 of a kink.
    mats: the summed weight gradient is held to rtol * scale plus, per violating sample, twice the largest
    element of that sample's own contribution (extra_of(sample id) -> |contribution of that one sample|; a
    flipped unit moves whole rows of the sample's outer products, including rows the oracle has at
    exactly zero, so the allowance is per sample, not per element), nothing more."""
    bad = np.zeros(B, bool)
    for what, got, want in per_sample:
        got = np.asarray(got, np.float64).reshape(B, -1)
        want = np.asarray(want, np.float64).reshape(B, -1)
        tol = rtol * np.abs(want).max()
        err = np.abs(got - want).max(axis=1)
        bad |= err > tol
        ok = err <= tol
        sc = np.abs(want).max()
        _report(dict(what=f"{name} {what}: samples over the 1e-5 bar", n=int((~ok).sum()), of=int(B),
                     per_sample_err_over_scale_p50=float(np.percentile(err, 50) / sc),
                     per_sample_err_over_scale_p99=float(np.percentile(err, 99) / sc),
                     max_err_over_scale_rest=float(err[ok].max() / sc) if ok.any() else 0.0,
                     max_err_over_scale_all=float(err.max() / sc)))
    ids = np.nonzero(bad)[0]
    _report(dict(what=f"{name}: samples whose closest pre-activation is within 1e-6 / 1e-5 / 1e-4 of a kink",
                 n=[int((near < m).sum()) for m in (1e-6, 1e-5, 1e-4)], violating=int(ids.size)))
    assert np.all(near[ids] < 1e-4), (name, "a sample away from every ReLU kink misses the 1e-5 bar", ids[near[ids] >= 1e-4][:5])
    assert ids.size <= max(2, int((near < 1e-5).sum())), (name, ids.size, int((near < 1e-5).sum()))
    if mats_got is not None:
        got, want = np.asarray(mats_got, np.float64), np.asarray(mats_want, np.float64)
        scale = np.abs(want).max()
        tol = rtol * scale + np.zeros_like(want)
        if mats64 is not None:
            tol += min(2.0 * np.abs(want - mats64).max(), rtol * scale)
        assert ids.size <= 64, (name, "too many violating samples", ids.size)
        for i in ids:
            tol += 2.0 * float(extra_of(np.array([i])).max())
        err = np.abs(got - want)
        _report(dict(what=f"{name} mats_grad (kink-aware)", max_err_over_scale=float(err.max() / scale),
                     violating_samples=int(ids.size)))
        assert np.all(err <= tol + 1e-30), (name, "mats_grad", float((err - tol).max()), scale)
    return ids


def _chunked_oracle(kind, fc, cin, B, feats, targets, E_of, W_of, bias, mats, chunk, dtype=np.float32, only=None):
    """-> loss, dE[N,K], dw[N], dbias, dmats (fp64 accumulation over chunks), near[B].
    only: restrict to these samples (their contribution to the full-batch result, same 1/B scale)."""
    o = refport.Model(kind, F, K, fc, cin, dtype=dtype)
    sel = np.arange(B) if only is None else np.asarray(only)
    n = sel.size
    dE = np.zeros((n * F, K), dtype)
    dw = np.zeros(n * F, dtype)
    near = np.full(n, np.inf)
    gm = np.zeros(mats.size, np.float64)
    gb, loss = 0.0, 0.0
    fm = feats.reshape(B, F)
    for c0 in range(0, n, chunk):
        s = sel[c0:c0 + chunk]
        Bc = s.size
        f = np.ascontiguousarray(fm[s].reshape(-1))
        idx = np.repeat(np.arange(Bc, dtype=np.int32), F)
        emb = E_of(f).reshape(-1).astype(dtype)
        w = W_of(f).astype(dtype)
        bb, mm = bias.astype(dtype), mats.astype(dtype)
        l = o.backward(Bc, idx, w, bb, emb, mm, targets[s])
        sc = Bc / B
        loss += l * sc
        dE[c0 * F:(c0 + Bc) * F] = emb.reshape(-1, K) * dtype(sc)
        dw[c0 * F:(c0 + Bc) * F] = w * dtype(sc)
        gm += mm.astype(np.float64) * sc
        gb += float(bb[0]) * sc
        near[c0:c0 + Bc] = _near_kink(o, Bc)
    return loss, dE, dw, gb, gm, near


@pytest.mark.parametrize("name,fc,cin,chunk,with64", [
    ("deepfm", [400, 400, 400], [], 1024, True),
    ("xdeepfm", [400, 400, 400], [200, 200, 200], 256, False),
])
def test_benchmark_batch_parity(gpu_pkg, name, fc, cin, chunk, with64):
    synth = gpu_pkg.synth
    B, rows = 8192, 39 * (1 << 18)
    model = gpu_pkg.make_model(name, F, K, fc, cin)
    table = gpu_pkg.EmbeddingTable(rows, K)
    table.init_uniform(42)
    mats = synth.init_mats(42, model.getMatsSize())
    bias = np.array([0.1], np.float32)
    ps = gpu_pkg.ParRecModel(model, table)
    ps.setParams(bias, mats)
    _, feats = synth.make_feats(1234, 7, B, F, rows)        # an ordinary batch: nothing filtered
    targets = synth.make_targets(1234, feats, B, F)
    for _ in range(3):                                       # eager, captured, replayed: the bench's own path
        loss = ps.optimize(feats, targets) / B
    res = ps.stepResults()
    # per-nnz gradients of the step (before the per-id sums): the per-sample picture
    lib, L = gpu_pkg.lib(), gpu_pkg._lib
    import ctypes as C
    import torch
    pe, pw = C.c_void_p(), C.c_void_p()
    L.check(lib.b200rec_step_nnz_grad_ptrs(model.handle, C.byref(pe), C.byref(pw)))

    def dev_array(ptr, n):
        class _A:
            __cuda_array_interface__ = dict(shape=(n,), typestr="<f4", data=(ptr, False), version=2)
        return torch.as_tensor(_A(), device="cuda").cpu().numpy().copy()

    L.check(lib.b200rec_model_sync(model.handle))
    g_dE = dev_array(pe.value, B * F * K).reshape(B, F * K)
    g_dw = dev_array(pw.value, B * F).reshape(B, F)
    E_of = lambda f: synth.table_rows(42, f, K)
    W_of = lambda f: synth.wtable_rows(42, f)
    args = (name, fc, cin, B, feats, targets, E_of, W_of, bias, mats, chunk)
    oloss, dE, dw, gb, gm, near = _chunked_oracle(*args)
    gm64 = None
    if with64:
        _, _, _, _, gm64, near = _chunked_oracle(*args, dtype=np.float64)
    assert abs(loss - oloss) <= 1e-5 * abs(oloss), (loss, oloss)
    extra = lambda ids: np.abs(_chunked_oracle(*args, only=ids)[4])
    viol = _kink_aware(f"{name} B=8192", B, [("dembedding", g_dE, dE.reshape(B, F * K)), ("dweights", g_dw, dw.reshape(B, F))],
                       near, extra, res["mats_grad"], gm, gm64)
    # the dedup scatter-add of the step: ids bit-exact; sums compared on the ids no violating sample touches
    ids, G = refport.make_embedding_grad(dE.reshape(-1), feats, K)
    _, gw = refport.make_weights_grad(dw, feats)
    assert np.array_equal(res["unique"], ids)
    clean = ~np.isin(ids, np.unique(feats.reshape(B, F)[viol])) if viol.size else np.ones(ids.size, bool)
    assert_close(res["emb_grad"][clean], G[clean], what=f"{name} B=8192 emb_grad (ids of clean samples)")
    assert_close(res["w_grad"][clean], gw[clean], what=f"{name} B=8192 w_grad (ids of clean samples)")
    assert abs(res["bias_grad"] - gb) <= 1e-5 * abs(gb) + 1e-6 * np.abs(dw).sum() / F
    model.close(); table.close()


@pytest.mark.parametrize("name,fc,cin,depth", [
    ("deepfm", [400, 400, 400], [], 0),
    ("xdeepfm", [128, 64], [48, 32], 0),
    ("dcn", [128, 64], [], 4),
    ("pnn", [128, 64], [], 0),
])
@pytest.mark.parametrize("seed", [11, 12])
def test_unfiltered_batch(gpu_pkg, name, fc, cin, depth, seed):
    """Internal<M>Model.forward / .backward on a batch taken as it comes (no away_from_kinks)."""
    B = 512
    rng = np.random.default_rng(seed)
    n = B * F
    index = np.repeat(np.arange(B, dtype=np.int32), F)
    w = rng.uniform(-0.05, 0.05, n).astype(np.float32)
    bias = np.array([0.1], np.float32)
    emb = rng.uniform(-0.3, 0.3, n * K).astype(np.float32)
    mats = gpu_pkg.synth.init_mats(seed, refport.mats_size(name, F, K, fc, cin, depth))
    targets = (rng.uniform(0, 1, B) < 0.4).astype(np.float32)
    o32 = refport.Model(name, F, K, fc, cin, depth)
    o64 = refport.Model(name, F, K, fc, cin, depth, np.float64)
    m = gpu_pkg.make_model(name, F, K, fc, cin, depth)
    p64 = o64.forward(B, index, w, bias, emb, mats)
    near = _near_kink(o64, B)
    assert_close(m.forward(B, index, w, bias, emb, mats), o32.forward(B, index, w, bias, emb, mats),
                 what=f"{name} unfiltered preds", ref64=p64)      # the forward is continuous at a kink
    gw, gb, ge, gm = w.copy(), bias.copy(), emb.copy(), mats.copy()
    loss = m.backward(B, index, gw, gb, ge, gm, targets)
    rw, rb, re, rm = w.copy(), bias.copy(), emb.copy(), mats.copy()
    rloss = o32.backward(B, index, rw, rb, re, rm, targets)
    dw, db, de, dm = (a.astype(np.float64) for a in (w, bias, emb, mats))
    o64.backward(B, index, dw, db, de, dm, targets)
    assert abs(loss - rloss) <= 1e-5 * abs(rloss)

    def extra(ids):      # |contribution of these samples to the full-batch mats gradient|
        sub = np.concatenate([np.arange(i * F, (i + 1) * F) for i in ids])
        sw, sb, sm = w[sub].copy(), bias.copy(), mats.copy()
        se = np.ascontiguousarray(emb.reshape(B, F * K)[ids].reshape(-1))
        o32.backward(len(ids), np.repeat(np.arange(len(ids), dtype=np.int32), F), sw, sb, se, sm, targets[ids])
        return np.abs(sm.astype(np.float64)) * (len(ids) / B)

    _kink_aware(f"{name} unfiltered seed {seed}", B,
                [("dembedding", ge.reshape(B, -1), re.reshape(B, -1)), ("dweights", gw.reshape(B, -1), rw.reshape(B, -1))],
                near, extra, gm, rm, dm)
    m.close()
