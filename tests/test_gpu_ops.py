"""GPU parity of the reference's own modules (nn/*.scala), the lookup and the scatter-add.
Integer / copy work is bit-exact; in-order sums are bit-exact too (same order as the reference)."""
import numpy as np
import pytest

from common import assert_close
from oracle import refport

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,n,n_out,sorted_", [(1, 1, 1, True), (16, 200, 1, True), (16, 200, 3, False),
                                               (5, 0, 1, True), (300, 5000, 2, True)])
def test_scatter_module(gpu_pkg, B, n, n_out, sorted_):
    rng = np.random.default_rng(n + B)
    index = rng.integers(0, B, n).astype(np.int32)
    if sorted_:
        index.sort()
    x = rng.standard_normal((n, n_out)).astype(np.float32)
    mod = gpu_pkg.Scatter(B, n_out)
    out = mod.updateOutput((x, index))
    ref = refport.scatter_update_output(x, index, B, n_out)
    assert np.array_equal(out, ref)  # sequential i order on both sides -> bit-exact
    go = rng.standard_normal((B, n_out)).astype(np.float32)
    gi = mod.updateGradInput((x, index), go)[0]
    assert np.array_equal(gi, refport.scatter_update_grad_input(go, index, B))


def test_scatter_module_bad_index(gpu_pkg):
    mod = gpu_pkg.Scatter(4, 1)
    with pytest.raises(ValueError):
        mod.updateOutput((np.ones(3, np.float32), np.array([0, 4, 1], np.int32)))
    with pytest.raises(ValueError):
        mod.updateGradInput((np.ones(3, np.float32), np.array([0, -1, 1], np.int32)), np.ones((4, 1), np.float32))


@pytest.mark.parametrize("B,F,K", [(3, 4, 5), (17, 39, 16), (1, 2, 1)])
def test_gather_dotproduct_modules(gpu_pkg, B, F, K):
    rng = np.random.default_rng(F)
    x = rng.standard_normal((B, F, K)).astype(np.float32)
    rows, cols = refport.pnn_pairs(F)
    P = len(rows)
    g = gpu_pkg.Gather(B, P, K)
    ro, co = g.updateOutput((x, rows, cols))
    rr, rc = refport.gather_update_output(x, rows, cols)
    assert np.array_equal(ro, rr) and np.array_equal(co, rc)
    d = gpu_pkg.DotProduct2()
    out = d.updateOutput((ro, co))
    assert np.array_equal(out, refport.dotproduct2_update_output(rr, rc))  # cmul + sequential sum
    go = rng.standard_normal((B, P)).astype(np.float32)
    ga, gb = d.updateGradInput((ro, co), go)
    ra, rb = refport.dotproduct2_update_grad_input(rr, rc, go)
    assert np.array_equal(ga, ra) and np.array_equal(gb, rb)
    gi = g.updateGradInput((x, rows, cols), (ga, gb))[0]
    assert np.array_equal(gi, refport.gather_update_grad_input(x.shape, rows, cols, ra, rb))


@pytest.mark.parametrize("B,F,K", [(1, 1, 4), (9, 39, 16), (33, 7, 6), (8, 3, 64), (5, 40, 128)])
def test_second_order_encoder(gpu_pkg, B, F, K):
    rng = np.random.default_rng(K)
    e = rng.uniform(-1, 1, B * F * K).astype(np.float32)
    enc = gpu_pkg.SecondOrderEncoder(B, F, K)
    out = enc.forward(e)
    ref = refport.second_order_fwd(e, B, F, K)
    ref64 = refport.second_order_fwd(e.astype(np.float64), B, F, K)
    assert_close(out, ref, what="second", ref64=ref64)
    go = rng.standard_normal(B).astype(np.float32)
    gi = enc.backward(e, go)
    rg = refport.second_order_bwd(e, go, B, F, K)
    rg64 = refport.second_order_bwd(e.astype(np.float64), go.astype(np.float64), B, F, K)
    assert_close(gi, rg, what="second grad", ref64=rg64)
    with pytest.raises(ValueError):
        enc.forward(e[:-1])


@pytest.mark.parametrize("mode", [1, 0], ids=["tc3xtf32", "simt"])
@pytest.mark.parametrize("B,I,O", [(1, 1, 1), (37, 53, 29), (256, 624, 400), (130, 400, 1), (64, 7, 300),
                                   (1000, 741, 400), (300, 100, 257)])
def test_linear_module(gpu_pkg, B, I, O, mode):
    gpu_pkg._lib.set_default_gemm_mode(mode)
    rng = np.random.default_rng(I * O)
    x = rng.standard_normal((B, I)).astype(np.float32)
    w = (rng.standard_normal((O, I)) / np.sqrt(I)).astype(np.float32)
    b = rng.standard_normal(O).astype(np.float32)
    lin = gpu_pkg.Linear(I, O, True, w, b)
    y = lin.updateOutput(x)
    y64 = x.astype(np.float64) @ w.astype(np.float64).T + b
    assert_close(y, refport.linear_fwd(x, w, b), what="linear y", ref64=y64)
    gy = rng.standard_normal((B, O)).astype(np.float32)
    gx = lin.updateGradInput(x, gy)
    rgx, rgw, rgb = refport.linear_bwd(x, w, gy)
    assert_close(gx, rgx, what="linear gx", ref64=gy.astype(np.float64) @ w.astype(np.float64))
    lin.accGradParameters(x, gy)
    lin.accGradParameters(x, gy)   # accumulates (BigDL accGradParameters)
    gw64 = 2 * (gy.astype(np.float64).T @ x.astype(np.float64))
    assert_close(lin.gradWeight, 2 * rgw, what="linear gW", ref64=gw64)
    assert_close(lin.gradBias, 2 * rgb, what="linear gb", ref64=2 * gy.astype(np.float64).sum(0))
    gpu_pkg._lib.set_default_gemm_mode(1)


def test_tf32_single_pass_is_not_parity_grade(gpu_pkg):
    """Mode 2 (one TF32 pass) exists for comparison only: its error is ~1e-3, which is why the
    parity path uses the 3xTF32 split."""
    rng = np.random.default_rng(0)
    B, I, O = 256, 512, 256
    x = rng.standard_normal((B, I)).astype(np.float32)
    w = (rng.standard_normal((O, I)) / np.sqrt(I)).astype(np.float32)
    y64 = x.astype(np.float64) @ w.astype(np.float64).T
    errs = {}
    for mode in (1, 2):
        gpu_pkg._lib.set_default_gemm_mode(mode)
        y = gpu_pkg.Linear(I, O, False, w).updateOutput(x)
        errs[mode] = np.abs(y - y64).max() / np.abs(y64).max()
    gpu_pkg._lib.set_default_gemm_mode(1)
    assert errs[1] < 3e-6 and 1e-4 < errs[2] < 1e-2, errs


@pytest.mark.parametrize("rows,K,n", [(10, 16, 1), (1000, 16, 5000), (77, 8, 300), (50, 6, 100), (33, 64, 64),
                                      (40, 4, 0)])
def test_lookup_bit_exact(gpu_pkg, rows, K, n):
    rng = np.random.default_rng(rows)
    t = gpu_pkg.EmbeddingTable(rows, K)
    E = rng.standard_normal((rows, K)).astype(np.float32)
    w = rng.standard_normal(rows).astype(np.float32)
    t.write(0, E, w)
    feats = rng.integers(0, rows, n).astype(np.int32)
    e, ww = t.lookup(feats)
    assert np.array_equal(e, refport.make_embeddings(E, feats))
    assert np.array_equal(ww, refport.make_weights(w, feats))
    if n:
        bad = feats.copy()
        bad[0] = rows
        with pytest.raises(ValueError):
            t.lookup(bad)
    t.close()


def test_table_init_matches_host_hash(gpu_pkg):
    synth = gpu_pkg.synth
    rows, K = 4099, 16
    t = gpu_pkg.EmbeddingTable(rows, K)
    t.init_uniform(42)
    E, w = t.read(0, rows)
    ids = np.arange(rows)
    assert np.array_equal(E, synth.table_rows(42, ids, K))
    assert np.array_equal(w, synth.wtable_rows(42, ids))
    # a strided shard regenerates exactly its slice of the global table
    t2 = gpu_pkg.EmbeddingTable(1000, K)
    t2.init_uniform(42, row_offset=3, row_stride=4)
    E2, w2 = t2.read(0, 1000)
    gids = 3 + 4 * np.arange(1000)
    assert np.array_equal(E2, synth.table_rows(42, gids, K))
    assert np.array_equal(w2, synth.wtable_rows(42, gids))
    t.close(); t2.close()


@pytest.mark.parametrize("n,rows,K", [(1, 5, 16), (5000, 50, 16), (20000, 100000, 16), (3000, 7, 8),
                                      (999, 30, 6), (4000, 3, 64), (0, 5, 16)])
def test_scatter_add_bit_exact(gpu_pkg, n, rows, K):
    """makeEmbeddingGrad sums duplicate ids in nnz order (Int2FloatOpenHashMap.addTo, i ascending);
    the sorted-index segmented sum keeps that order -> bit-identical, including hot ids with
    thousands of rows (the long-segment kernel)."""
    rng = np.random.default_rng(n + K)
    # power-law ids: a few very hot rows
    feats = np.minimum((rng.pareto(0.7, n)).astype(np.int64), rows - 1).astype(np.int32)
    dE = rng.standard_normal((n, K)).astype(np.float32)
    dw = rng.standard_normal(n).astype(np.float32)
    ids, G, gw = gpu_pkg.scatter_add(feats, dE, dw, dim=K)
    if n == 0:
        assert len(ids) == 0
        return
    rids, rG = refport.make_embedding_grad(dE.reshape(-1), feats, K)
    _, rgw = refport.make_weights_grad(dw, feats)
    assert np.array_equal(ids, rids)
    assert np.array_equal(G, rG)
    assert np.array_equal(gw, rgw)
    assert np.array_equal(gpu_pkg.distinct(feats), refport.distinct_int_indices(feats))


@pytest.mark.gpu
@pytest.mark.parametrize("pattern", ["random31", "equal", "ascending", "descending", "two_values", "powerlaw"])
@pytest.mark.parametrize("n", [2047, 2048, 2049, 4097, 700001])
def test_scatter_add_sort_stress(gpu_pkg, pattern, n):
    """The sort half is the library's own radix sort (segsum.cu: rs_hist / rs_rowscan / rs_scatter, 2048-item tiles,
    8-bit digits, ranks by warp match): tile-boundary sizes, many tiles, all 31 key bits, degenerate digit
    distributions -- ids, sums (in non-zero order, bit-exact) and the distinct set against the oracle."""
    rng = np.random.default_rng(n)
    K = 4
    if pattern == "random31":
        feats = rng.integers(0, 2**31 - 1, n)
    elif pattern == "equal":
        feats = np.full(n, 123456789)
    elif pattern == "ascending":
        feats = np.arange(n) * 3001 % (2**31 - 1)
        feats.sort()
    elif pattern == "descending":
        feats = np.sort(rng.integers(0, 2**31 - 1, n))[::-1]
    elif pattern == "two_values":
        feats = np.where(rng.random(n) < 0.5, 255, 256)          # a carry between the first two digits
    else:
        feats = np.minimum(rng.pareto(0.7, n).astype(np.int64), 10**6)
    feats = np.ascontiguousarray(feats).astype(np.int32)
    dE = rng.standard_normal((n, K)).astype(np.float32)
    dw = rng.standard_normal(n).astype(np.float32)
    ids, G, gw = gpu_pkg.scatter_add(feats, dE, dw, dim=K)
    rids, rG = refport.make_embedding_grad(dE.reshape(-1), feats, K)
    _, rgw = refport.make_weights_grad(dw, feats)
    assert np.array_equal(ids, rids)
    assert np.array_equal(G, rG)
    assert np.array_equal(gw, rgw)
