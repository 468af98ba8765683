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


REPORT = None   # parity report: one JSON line per comparison (gpurun_out/ travels back from the GPU box)


def _report(rec):
    global REPORT
    import json
    import os
    if REPORT is None:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        d = os.path.join(root, "gpurun_out")
        REPORT = os.path.join(d, "parity_report.jsonl") if os.path.isdir(d) else ""
    if REPORT:
        try:
            with open(REPORT, "a") as f:
                f.write(json.dumps(rec) + "\n")
        except OSError:
            pass


def assert_close(got, want, rtol=1e-5, what="", ref64=None, atol=0.0):
    """The parity bar: |got - want| <= rtol * max|want| for every element (north_star: 1e-5 relative,
    fp32).  `want` is the fp32 oracle -- itself one rounding path of the exact value -- so when its
    fp64 twin `ref64` is given the oracle's own distance to it is added, but NEVER more than another
    rtol * max|want|: an inaccurate oracle cannot widen the bar past 2e-5 of the tensor's scale.

    Besides the assertion, every call appends the element-wise picture to the parity report
    (gpurun_out/parity_report.jsonl): relative-error percentiles over the entries that carry signal
    (|want| >= 1e-3 * max|want|; smaller entries are sums that cancelled, where ANY fp32 path has large
    relative error) and, with ref64, the GPU's rms error against the exact value next to the fp32
    oracle's own."""
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    if got.size == 0:
        return
    scale = np.abs(want).max()
    tol = rtol * scale + atol
    noise = 0.0
    if ref64 is not None:
        noise = 2.0 * np.abs(want - np.asarray(ref64, np.float64)).max()
        tol += min(noise, rtol * scale)
    diff = np.abs(got - want)
    err = diff.max()
    rec = dict(what=what, n=int(got.size), scale=float(scale), max_err_over_scale=float(err / scale) if scale else 0.0,
               tol_over_scale=float(tol / scale) if scale else 0.0, oracle_noise_over_scale=float(noise / scale) if scale else 0.0)
    sig = np.abs(want) >= 1e-3 * scale
    if scale > 0 and sig.any():
        rel = diff[sig] / np.abs(want[sig])
        rec.update(signal_fraction=float(sig.mean()), rel_p50=float(np.percentile(rel, 50)),
                   rel_p99=float(np.percentile(rel, 99)), rel_p999=float(np.percentile(rel, 99.9)), rel_max=float(rel.max()))
    if ref64 is not None:
        r64 = np.asarray(ref64, np.float64)
        rec.update(rms_err_gpu_vs_exact=float(np.sqrt(np.mean((got - r64) ** 2))),
                   rms_err_oracle32_vs_exact=float(np.sqrt(np.mean((want - r64) ** 2))))
    _report(rec)
    assert err <= tol + 1e-30, f"{what}: max err {err:.3e} > tol {tol:.3e} (scale {scale:.3e})"


away_from_kinks = refport.away_from_kinks
