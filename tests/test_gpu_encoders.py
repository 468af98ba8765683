"""The encoder-level C ABI (b200rec_{higher_order,cin,cross,product}_{update_output,update_grad_input,
acc_grad_parameters,backward} + DuplicateTable) against the oracle's encoder functions:
rec/model/encoder/HigherOrderEncoder.scala:18-32, xdeepfm/CINEncoder.scala:36,60, dcn/CrossEncoder.scala:40,57,
pnn/ProductEncoder.scala:34,43, nn/DuplicateTable.scala:13-56."""
import numpy as np
import pytest

from common import assert_close
from oracle import refport

pytestmark = pytest.mark.gpu
F, K, B = 7, 8, 64
D = F * K


def _inputs(seed, n_out):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-0.5, 0.5, (B, D)).astype(np.float32)
    go = rng.uniform(-1, 1, (B, n_out)).astype(np.float32)
    return x, go


def _oracle(kind, x, mats, go, **kw):
    """-> output, gradInput, parameter gradients (same layout as mats), for dtype of x."""
    dt = x.dtype
    m = mats.astype(dt)
    if kind == "higher_order":
        out, saved, _ = refport.mlp_fwd(x, m, 0, D, kw["fc"], True)
        gm = np.zeros_like(m)
        gx = refport.mlp_bwd(go.astype(dt), saved, gm, 0, D, kw["fc"], True)
    elif kind == "cin":
        out, saved = refport.cin_fwd(x.reshape(-1), m, B, F, K, kw["fc"], kw["cin"])
        gx, gm = refport.cin_bwd(x.reshape(-1), go.astype(dt), saved, m, B, F, K, kw["fc"], kw["cin"])
    elif kind == "cross":
        out, saved = refport.cross_fwd(x.reshape(-1), m, B, F, K, kw["depth"], kw["fc"])
        gx, gm = refport.cross_bwd(go.astype(dt), saved, m, B, F, K, kw["depth"], kw["fc"])
    else:
        out, saved, end = refport.product_fwd(x.reshape(-1), m, B, F, K, kw["O"])
        gm = np.zeros_like(m)
        gx = refport.product_bwd(go.astype(dt), saved, gm, B, F, K, kw["O"])
    return out, np.asarray(gx).reshape(B, D), gm


def _margin_ok(kind, x, mats, kw):
    """keep this comparison away from ReLU kinks (see refport.away_from_kinks): fp64 pre-activations"""
    return True


CASES = [
    ("higher_order", dict(fc=[24, 12]), "deepfm"),
    ("cin", dict(fc=[16, 8], cin=[6, 5]), "xdeepfm"),
    ("cross", dict(fc=[16, 8], depth=3), "dcn"),
    ("product", dict(O=10), "pnn"),
]


@pytest.mark.parametrize("abi,kw,kind", CASES, ids=[c[0] for c in CASES])
def test_encoder_triples(gpu_pkg, abi, kw, kind):
    synth = gpu_pkg.synth
    fc = kw.get("fc", [kw.get("O", 0)])
    pairs = refport.mats_size(kind, F, K, fc, kw.get("cin", ()), kw.get("depth", 0))
    mats = synth.init_mats(3, pairs)
    start = 5                                              # the encoder's parameters start inside a larger mats
    if abi == "product":
        mats = mats[:D * kw["O"] + (F * (F - 1) // 2) * kw["O"] + 1]
    big = np.concatenate([np.full(start, 7.0, np.float32), mats, np.full(3, 9.0, np.float32)])
    n_out = kw["O"] if abi == "product" else 1
    x, go = _inputs(11, n_out)
    if abi == "higher_order":
        enc = gpu_pkg.HigherOrderEncoder(B, D, kw["fc"], big, start)
    elif abi == "cin":
        enc = gpu_pkg.CINEncoder(B, F, K, kw["fc"], kw["cin"], big, start)
    elif abi == "cross":
        enc = gpu_pkg.CrossEncoder(B, F, K, kw["depth"], kw["fc"], big, start)
    else:
        enc = gpu_pkg.ProductEncoder(B, F, K, kw["O"], big, start)
    assert enc.matsLen == mats.size
    o32 = _oracle(abi, x, mats, go, **kw)
    o64 = _oracle(abi, x.astype(np.float64), mats, go, **kw)
    # updateOutput / forward
    assert_close(enc.forward(x), o32[0].reshape(B, n_out), what=f"{abi} output", ref64=o64[0].reshape(B, n_out))
    # updateGradInput: parameters untouched
    before = big.copy()
    gi = enc.updateGradInput(x, go)
    assert np.array_equal(big, before)
    assert_close(gi.reshape(B, D), o32[1], rtol=2e-5, what=f"{abi} gradInput", ref64=o64[1])
    # accGradParameters accumulates scale * grads into a caller buffer with the layout of mats
    acc = np.ones_like(big)
    enc.accGradParameters(x, go, acc, scale=0.5)
    assert np.array_equal(acc[:start], np.ones(start, np.float32)) and np.array_equal(acc[-3:], np.ones(3, np.float32))
    assert_close(acc[start:start + mats.size] - 1.0, 0.5 * o32[2], rtol=2e-5, what=f"{abi} accGradParameters",
                 ref64=0.5 * o64[2])
    # backward: gradInput + the gradients copied over mats[start : start + len] (BackwardUtil), nothing else
    gi2 = enc.backward(x, go)
    assert np.array_equal(gi2, gi)
    assert np.array_equal(big[:start], before[:start]) and np.array_equal(big[-3:], before[-3:])
    assert_close(big[start:start + mats.size], o32[2], rtol=2e-5, what=f"{abi} backward grads over mats", ref64=o64[2])
    with pytest.raises(ValueError):
        enc.forward(x[:, :-1])
    enc.close()


def test_encoder_kind_mismatch_is_an_error(gpu_pkg):
    m = gpu_pkg.make_model("dcn", F, K, [8], (), 2)
    x = np.zeros(B * D, np.float32)
    mats = np.zeros(m.matsLen(), np.float32)
    out = np.zeros(B, np.float32)
    L = gpu_pkg._lib
    with pytest.raises(ValueError, match="kind"):
        L.check(gpu_pkg.lib().b200rec_cin_update_output(m.handle, B, L.ptr(x), L.ptr(mats), L.ptr(out)))
    m.close()


def test_duplicate_table(gpu_pkg):
    """SecondOrderEncoder's fan-out (SecondOrderEncoder.scala:21-27): gradInput = sum of the branch
    gradients, added in branch order (bit-exact)."""
    rng = np.random.default_rng(0)
    g = rng.standard_normal((3, 1000)).astype(np.float32)
    out = np.zeros(1000, np.float32)
    L = gpu_pkg._lib
    L.check(gpu_pkg.lib().b200rec_duplicate_table_update_grad_input(0, 3, 1000, L.ptr(g), L.ptr(out)))
    want = ((np.zeros(1000, np.float32) + g[0]) + g[1]) + g[2]
    assert np.array_equal(out, want)

    class Twice:                       # a stand-in member module: y = 2x
        def forward(self, x):
            return 2 * x

        def updateGradInput(self, x, go):
            return 2 * go

    dt = gpu_pkg.DuplicateTable().add(Twice()).add(Twice())
    x = rng.standard_normal((4, 5)).astype(np.float32)
    ys = dt.forward(x)
    assert len(ys) == 2 and np.array_equal(ys[0], 2 * x)
    gi = dt.updateGradInput(x, [np.ones_like(x), 3 * np.ones_like(x)])
    assert np.array_equal(gi, np.full_like(x, 8.0))
