"""Shared helpers for the parity tests: seeded inputs in the reference's flat-buffer ABI."""
import numpy as np

from oracle import refport

CONFIGS = {
    # kind: (constructor kwargs for the oracle / package)
    "lr": dict(),
    "fm": dict(),
    "deepfm": dict(fc_dims=[48, 24]),
    "xdeepfm": dict(fc_dims=[40, 24], cin_dims=[12, 10, 8]),
    "xdeepfm1": dict(fc_dims=[40, 24], cin_dims=[12]),      # L=1: the reference-exact CIN case (SURVEY B-2)
    "dcn": dict(fc_dims=[40, 24], cross_depth=3),
    "pnn": dict(fc_dims=[36, 20]),
}


def kind_of(name):
    return "xdeepfm" if name.startswith("xdeepfm") else name


def make_inputs(name, B, F, K, seed=0, scale=0.5, canonical=True):
    """Flat buffers (index, weights, bias, embedding, mats, targets) for one batch."""
    cfg = CONFIGS[name]
    kind = kind_of(name)
    rng = np.random.default_rng(seed)
    n = B * F
    index = np.repeat(np.arange(B, dtype=np.int32), F)
    weights = rng.uniform(-scale, scale, n).astype(np.float32)
    bias = np.array([0.1], np.float32)
    embedding = rng.uniform(-scale, scale, n * K).astype(np.float32) if kind != "lr" else None
    pairs = refport.mats_size(kind, F, K, cfg.get("fc_dims", ()), cfg.get("cin_dims", ()), cfg.get("cross_depth", 0))
    mats = None
    if pairs:
        blocks = []
        for i in range(0, len(pairs), 2):
            a, b = pairs[i], pairs[i + 1]
            s = 1.0 / np.sqrt(max(a, 1))
            blocks.append(rng.uniform(-s, s, a * b).astype(np.float32) if not (a == 1 and b == 1)
                          else rng.uniform(-0.1, 0.1, 1).astype(np.float32))
        mats = np.concatenate(blocks)
    targets = (rng.uniform(0, 1, B) < 0.4).astype(np.float32)
    return index, weights, bias, embedding, mats, targets


def oracle_model(name, F, K, dtype=np.float32):
    cfg = CONFIGS[name]
    return refport.Model(kind_of(name), F, K, cfg.get("fc_dims", ()), cfg.get("cin_dims", ()),
                         cfg.get("cross_depth", 0), dtype=dtype)


def pkg_model(pkg, name, F, K):
    cfg = CONFIGS[name]
    return pkg.make_model(kind_of(name), F, K, cfg.get("fc_dims", ()), cfg.get("cin_dims", ()),
                          cfg.get("cross_depth", 0))


def assert_close(got, want, rtol=1e-5, what="", ref64=None, atol=0.0):
    """|got - want| <= rtol * max|want| (+ the fp32 oracle's own distance to its fp64 twin when
    given: both sides are fp32 roundings of the same exact value, so the fp32 oracle is only
    known to that precision)."""
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    if got.size == 0:
        return
    scale = np.abs(want).max()
    tol = rtol * scale + atol
    if ref64 is not None:
        tol += 2.0 * np.abs(want - np.asarray(ref64, np.float64)).max()
    err = np.abs(got - want).max()
    assert err <= tol + 1e-30, f"{what}: max err {err:.3e} > tol {tol:.3e} (scale {scale:.3e})"


away_from_kinks = refport.away_from_kinks
