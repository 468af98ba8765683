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


def test_native_parser_matches_line_parser(pkg):
    """b200rec_parse_samples (host C++ in libb200rec.so) == the line-array parser on the same text,
    libsvm and libffm, with blank lines, tabs, CRLF and scientific notation; errors name the line."""
    d = pkg.data
    rng = np.random.default_rng(3)
    for fmt in ("libsvm", "libffm"):
        lines = []
        for r in range(300):
            n = int(rng.integers(0, 12))
            toks = []
            for _ in range(n):
                k, v = int(rng.integers(1, 2**31)), float(np.float32(rng.standard_normal()))
                toks.append((f"{int(rng.integers(0, 39))}:" if fmt == "libffm" else "") + f"{k}:{v!r}")
            lines.append(f"{int(rng.integers(0, 2))} " + " ".join(toks))
            if r % 50 == 7:
                lines.append("")
        want = d.parse(lines, fmt)
        got = d.parse_text("\n".join(lines) + "\n", fmt)
        for a, b in zip(got, want):
            assert a.dtype == b.dtype and np.array_equal(a, b)
        got = d.parse_text("\r\n".join(lines).replace(" ", " \t ").encode(), fmt)     # CRLF, runs of blanks, no final newline
        for a, b in zip(got, want):
            assert np.array_equal(a, b)
    idx, feats, vals, tg, fields = d.parse_text("1 3:7:2.5e-1 0:1:1\n-1 38:5:1", "libffm", with_fields=True)
    assert fields.tolist() == [3, 0, 38] and feats.tolist() == [6, 0, 4] and idx.tolist() == [0, 0, 1]
    assert vals.tolist() == [0.25, 1.0, 1.0] and tg.tolist() == [1.0, -1.0]
    e = d.parse_text("", "libsvm")
    assert all(len(a) == 0 for a in e)
    for bad, what in [("1 3:1\n0 x:1", "line 2"), ("1 3", "line 1"), ("1 3:1:1", "line 1"), ("a 1:1", "label"),
                      ("1 0:1", "1-based"), ("1 4294967297:1", "31 bits"), ("1 2:1z", "line 1")]:
        with pytest.raises(ValueError, match=what):
            d.parse_text(bad, "libsvm")
    with pytest.raises(ValueError, match="field"):
        d.parse_text("1 3:1", "libffm")
    with pytest.raises(ValueError):
        d.parse_text("1 1:1", "csv")


def test_native_parser_floats_round_like_strtof(pkg):
    """The decimal fast path (exact double quotient, one rounding to float, ties sent to strtof) gives
    the float the C library gives, on plain decimals, exponents, specials and near-tie literals."""
    import ctypes
    libc = ctypes.CDLL(None)
    libc.strtof.restype = ctypes.c_float
    libc.strtof.argtypes = [ctypes.c_char_p, ctypes.c_void_p]
    rng = np.random.default_rng(0)
    x = rng.standard_normal(6000) * 10.0 ** rng.integers(-6, 6, 6000)
    vals = [repr(float(np.float32(v))) for v in x[:2000]] + [f"{v:.9f}" for v in x[2000:4000]] + \
           [f"{v:.3f}" for v in x[4000:]]
    vals += ["0.1", "1e-3", "3.4028235e38", "1e-45", "16777217", "0.30000001192092896", "8388608.5", "8388609.5",
             "1.00000005960464477539", "-0", "+.5", "5.", "inf", "-inf", "1E2", "123456789012345678"]
    got = pkg.data.parse_text("\n".join(f"1 1:{v}" for v in vals))[2]
    want = np.array([libc.strtof(v.encode(), None) for v in vals], np.float32)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


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


def test_bench_helpers_on_cpu():
    """bench.py's pure-host helpers: the gather pattern ceiling is read from the committed micro-benchmark output,
    and the algorithmic work table names the phases the per-kernel pass tags."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_for_test", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    c = bench.gather_ceiling_us(8192 * 39)
    assert c is not None and 5.0 < c < 100.0          # microseconds for one batch of random 64-B rows
    assert bench.gather_ceiling_us(12345) is None      # only sizes the micro-benchmark measured
    w = bench.algorithmic_work("deepfm", [400, 400, 400], [], 0, 8192, 150000)
    for phase in ("gather_fm_fwd", "dense_fwd", "dense_bwd", "emb_grad", "scatter_add"):
        assert phase in w and w[phase][1] > 0
    assert w["gather_fm_fwd"][0] == "hbm" and w["dense_bwd"][0] == "tensor"
