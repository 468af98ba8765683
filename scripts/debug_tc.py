import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
from oracle import refport
rng = np.random.default_rng(0)

def rel(a, b):
    return np.abs(a - b).max() / (np.abs(b).max() + 1e-30)

# 1) plain linear at several shapes, each op separately, vs fp64
for (B, I, O) in [(256, 512, 256), (700, 624, 400), (700, 400, 400), (128, 32, 16), (128, 40, 16), (128, 64, 208), (128, 64, 224), (300, 400, 624)]:
    x = rng.standard_normal((B, I)).astype(np.float32)
    w = (rng.standard_normal((O, I)) / np.sqrt(I)).astype(np.float32)
    gy = rng.standard_normal((B, O)).astype(np.float32)
    y64 = x.astype(np.float64) @ w.astype(np.float64).T
    gx64 = gy.astype(np.float64) @ w.astype(np.float64)
    gw64 = gy.astype(np.float64).T @ x.astype(np.float64)
    out = []
    for mode in (0, 1, 2):
        pkg._lib.set_default_gemm_mode(mode)
        lin = pkg.Linear(I, O, False, w)
        y = lin.updateOutput(x)
        gx = lin.updateGradInput(x, gy)
        lin.accGradParameters(x, gy)
        out.append((rel(y, y64), rel(gx, gx64), rel(lin.gradWeight, gw64)))
    print((B, I, O), " ".join("mode%d fwd %.1e gx %.1e gw %.1e |" % ((m,) + o) for m, o in enumerate(out)))
    # where are the errors?
    pkg._lib.set_default_gemm_mode(1)
    lin = pkg.Linear(I, O, False, w)
    y = lin.updateOutput(x)
    e = np.abs(y - y64)
    bad = np.argwhere(e > 1e-4 * np.abs(y64).max())
    if len(bad):
        print("   bad fwd entries:", len(bad), "rows", np.unique(bad[:, 0])[:10], "cols", np.unique(bad[:, 1])[:20])
    gx = lin.updateGradInput(x, gy)
    e = np.abs(gx - gx64)
    bad = np.argwhere(e > 1e-4 * np.abs(gx64).max())
    if len(bad):
        print("   bad gx entries:", len(bad), "rows", np.unique(bad[:, 0])[:10], "cols", np.unique(bad[:, 1])[:20])
pkg._lib.set_default_gemm_mode(1)

# 2) model level: mode 1 vs mode 0 per tensor
for kind, B, kw in [("deepfm", 700, dict(fc=[400, 400, 400])), ("xdeepfm", 96, dict(fc=[400, 400], cin=[200, 200, 200])),
                    ("xdeepfm", 40, dict(fc=[64], cin=[100, 50])), ("xdeepfm", 257, dict(fc=[40, 24], cin=[12]))]:
    F, K = 39, 16
    n = B * F
    index = np.repeat(np.arange(B, dtype=np.int32), F)
    w = rng.uniform(-0.05, 0.05, n).astype(np.float32)
    bias = np.array([0.1], np.float32)
    emb = rng.uniform(-0.3, 0.3, n * K).astype(np.float32)
    fc, cin = kw.get("fc", ()), kw.get("cin", ())
    pairs = refport.mats_size(kind, F, K, fc, cin, 0)
    mats = pkg.synth.init_mats(3, pairs)
    targets = (rng.uniform(0, 1, B) < 0.4).astype(np.float32)
    res = {}
    for mode in (0, 1):
        m = pkg.make_model(kind, F, K, fc, cin, 0)
        m.setGemmMode(mode)
        p = m.forward(B, index, w, bias, emb, mats)
        gw, gb, ge, gm = w.copy(), bias.copy(), emb.copy(), mats.copy()
        loss = m.backward(B, index, gw, gb, ge, gm, targets)
        res[mode] = (p, ge, gm)
        m.close()
    print(kind, B, kw, "preds %.1e dE %.1e dmats %.1e" % tuple(rel(res[1][i], res[0][i]) for i in range(3)))
    off = 0
    for i in range(0, len(pairs), 2):
        sz = pairs[i] * pairs[i + 1]
        a, b = res[1][2][off:off + sz], res[0][2][off:off + sz]
        print("    block", pairs[i], pairs[i + 1], "rel %.1e" % rel(a, b), "scale %.2e" % np.abs(b).max())
        off += sz
