import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
rng = np.random.default_rng(0)
for (B, I, O) in [(256, 64, 128), (256, 512, 128), (256, 2048, 128), (256, 8192, 128), (256, 32768, 128)]:
    for dist in ("normal", "positive"):
        x = rng.standard_normal((B, I)).astype(np.float32)
        w = (rng.standard_normal((O, I)) / np.sqrt(I)).astype(np.float32)
        if dist == "positive":
            x, w = np.abs(x), np.abs(w)
        y64 = x.astype(np.float64) @ w.astype(np.float64).T
        row = []
        for mode in (0, 1):
            pkg._lib.set_default_gemm_mode(mode)
            y = pkg.Linear(I, O, False, w).updateOutput(x)
            d = (y - y64)
            row.append("mode%d max %.2e rms %.2e mean(signed*sign(y)) %.2e" % (
                mode, np.abs(d).max() / np.abs(y64).max(), np.sqrt((d ** 2).mean()) / np.abs(y64).max(),
                (d * np.sign(y64)).mean() / np.abs(y64).max()))
        ynp = x @ w.T
        print((B, I, O), dist, " | ".join(row), "| numpy32 max %.2e" % (np.abs(ynp - y64).max() / np.abs(y64).max()))
