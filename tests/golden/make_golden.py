"""Generate the committed golden fixtures (tests/golden/*.npz).

The reference ships no golden vectors and cannot run here (Scala/JVM), so these
fixtures are produced by the INDEPENDENT fp64 torch-autograd implementation (tests/independent.py,
written from the papers' formulas, not from oracle/refport.py).  They pin the oracle
(tests/test_oracle.py) and, through it, the CUDA path.  Inputs are float32 values stored exactly;
outputs are float64.

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import independent  # noqa: E402
from common import CONFIGS, kind_of, make_inputs  # noqa: E402

CASES = [(name, B, F, K) for name in CONFIGS for (B, F, K) in [(6, 5, 4), (12, 39, 16)]]


def main():
    for name, B, F, K in CASES:
        cfg = CONFIGS[name]
        index, w, bias, emb, mats, targets = make_inputs(name, B, F, K, seed=B + F)
        r = independent.run(kind_of(name), B, F, K, index, w, bias, emb, mats, targets,
                            cfg.get("fc_dims", ()), cfg.get("cin_dims", ()), cfg.get("cross_depth", 0))
        out = dict(index=index, weights=w, bias=bias, targets=targets, pred=r["pred"], loss=np.float64(r["loss"]),
                   gw=r["gw"], gb=r["gb"])
        if emb is not None:
            out.update(embedding=emb, ge=r["ge"])
        if mats is not None:
            out.update(mats=mats, gm=r["gm"])
        np.savez_compressed(os.path.join(HERE, f"{name}_B{B}_F{F}_K{K}.npz"), **out)
        print("wrote", name, B, F, K)


if __name__ == "__main__":
    main()
