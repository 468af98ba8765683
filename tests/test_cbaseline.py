"""The CPU baseline bench.py times (oracle/cbaseline.{cpp,py}) against the
oracle: the fastutil-style hash maps reproduce makeEmbeddings / makeWeights / make*Grad /
distinctIntIndices bit for bit, and the torch-CPU (MKL) module graph matches refport within 1e-5."""
import numpy as np
import pytest

from oracle import cbaseline as cb
from oracle import refport

F, K = 39, 16


@pytest.fixture(scope="module")
def synth(pkg):
    cb.build()
    return pkg.synth


@pytest.mark.parametrize("threads", [1, 3])
def test_hash_maps_bit_exact(synth, threads):
    B, rows = 96, 39 * 64          # small vocabulary: many duplicate ids, id 0 (fastutil's null key) included
    index, feats = synth.make_feats(1234, 0, B, F, rows)
    feats[:5] = 0
    ids = np.unique(feats).astype(np.int32)
    E, wv = synth.table_rows(42, ids, K), synth.wtable_rows(42, ids)
    p = cb.Pulled(ids, E, wv, threads)
    emb, ww = cb.make_embeddings(p, feats, threads), cb.make_weights(p, feats)
    pos = np.searchsorted(ids, feats)
    assert np.array_equal(emb, refport.make_embeddings(E, pos))
    assert np.array_equal(ww, refport.make_weights(wv, pos))
    assert np.array_equal(cb.distinct(feats), refport.distinct_int_indices(feats))
    rng = np.random.default_rng(0)
    ge = rng.standard_normal(feats.size * K).astype(np.float32)
    gw = rng.standard_normal(feats.size).astype(np.float32)
    i1, G1 = refport.make_embedding_grad(ge, feats, K)
    _, w1 = refport.make_weights_grad(gw, feats)
    i2, G2, w2, secs = cb.scatter_add(feats, ge, gw, K, threads=threads)
    assert np.array_equal(i1, i2) and np.array_equal(G1, G2) and np.array_equal(w1, w2)   # nnz order kept
    assert secs > 0


def test_hash_map_rehash_and_empty(synth):
    # an undersized expected capacity forces rehashes; an empty batch is a no-op
    feats = np.arange(1, 5000, dtype=np.int32).repeat(2)
    g = np.ones(feats.size, np.float32)
    ids, _, gw, _ = cb.scatter_add(feats, None, g, K, ids=np.unique(feats).astype(np.int32))
    assert np.array_equal(gw, np.full(ids.size, 2.0, np.float32))
    assert cb.distinct(np.zeros(0, np.int32)).size == 0


@pytest.mark.parametrize("kind,fc,cin", [("lr", [], []), ("fm", [], []), ("deepfm", [32, 16], []),
                                         ("xdeepfm", [32, 16], [8, 8]), ("xdeepfm", [24], [10])])
def test_torch_model_matches_refport(synth, kind, fc, cin):
    B, rows = 48, 39 * 512
    index, feats = synth.make_feats(1234, 1, B, F, rows)
    targets = synth.make_targets(1234, feats, B, F)
    emb = synth.table_rows(42, feats, K).reshape(-1) if kind != "lr" else None
    w = synth.wtable_rows(42, feats)
    mats = synth.init_mats(42, refport.mats_size(kind, F, K, fc, cin, 0))
    cp = lambda a: None if a is None or not a.size else a.copy()
    a = [cp(w), np.array([0.1], np.float32), cp(emb), cp(mats)]
    b = [cp(w), np.array([0.1], np.float32), cp(emb), cp(mats)]
    l1 = refport.Model(kind, F, K, fc, cin).backward(B, index, a[0], a[1], a[2], a[3], targets)
    l2 = cb.TorchModel(kind, F, K, fc, cin).backward(B, index, b[0], b[1], b[2], b[3], targets)
    assert abs(l1 - l2) <= 1e-6 * abs(l1)
    for x, y, what in zip(a, b, ("weights", "bias", "embedding", "mats")):
        if x is not None:
            assert np.abs(x - y).max() <= 1e-5 * np.abs(x).max() + 1e-12, what


def test_run_steps_reports_phases_and_threads(synth):
    r = cb.run_steps("deepfm", F, K, [16], [], 0, 64, 39 * 256, synth, 1234, 42, threads=2, budget_s=0.5, max_steps=2)
    assert r["steps"] == 2 and r["threads"] == 2 and r["value"] > 0
    assert set(r["phases_ms"]) == {"gather", "mats_copy", "dense", "scatter_add"}
    import torch
    assert torch.get_num_threads() >= 1
