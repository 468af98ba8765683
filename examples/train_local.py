#!/usr/bin/env python
"""The reference's driver loop (rec/example/DeepFMLocalExample.scala:36-56) on one B200:
epochs of optimize over mini-batches, then predict + AUC, printing `epoch=… loss=… auc=… time=…ms`.

    python examples/train_local.py --model deepfm --epochs 3 [--input file.libsvm] [--optim adam]
Without --input a synthetic Criteo-shaped set is generated (and round-tripped through the libsvm
text parser, so the same code path serves real files).
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--input", default=None)
    ap.add_argument("--model", default="deepfm")
    ap.add_argument("--batchSize", type=int, default=2048)
    ap.add_argument("--stepSize", type=float, default=0.0025)     # the examples' default
    ap.add_argument("--optim", default="adam")                     # DeepFMLocalExample.scala:32
    ap.add_argument("--inputDim", type=int, default=39 * 4096)
    ap.add_argument("--nFields", type=int, default=39)
    ap.add_argument("--embeddingDim", type=int, default=16)
    ap.add_argument("--fcDims", default="64,32")
    ap.add_argument("--cinDims", default="32,32")
    ap.add_argument("--crossDepth", type=int, default=3)
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--samples", type=int, default=16384)
    a = ap.parse_args()
    b = g.load_package()
    F, K = a.nFields, a.embeddingDim
    fc = [int(x) for x in a.fcDims.split(",") if x]
    cin = [int(x) for x in a.cinDims.split(",") if x]
    if a.input:
        text = open(a.input, "rb").read()
    else:
        _, feats = b.synth.make_feats(1234, 0, a.samples, F, a.inputDim)
        text = "\n".join(b.data.to_libsvm(feats, b.synth.make_targets(1234, feats, a.samples, F), F)).encode()
    # the whole file through the native parser (b200rec_parse_samples), then mini-batches of whole samples
    index, cols_all, _, targets_all = b.data.parse_text(text, "libsvm")
    if len(targets_all) * F != len(cols_all) or not np.array_equal(index, np.repeat(np.arange(len(targets_all)), F)):
        raise ValueError("every sample needs exactly nFields features")
    model = b.make_model(a.model, F, K, fc, cin if a.model == "xdeepfm" else (), a.crossDepth)
    table = b.EmbeddingTable(a.inputDim, K if a.model != "lr" else 0)
    table.init_uniform(42)
    ps = b.ParRecModel(model, table)
    ps.setParams(np.array([0.0], np.float32), b.synth.init_mats(42, model.getMatsSize()))
    batches = [(cols_all[i * F:(i + a.batchSize) * F], targets_all[i:i + a.batchSize])
               for i in range(0, len(targets_all), a.batchSize)]
    for epoch in range(1, a.epochs + 1):
        t0 = time.time()
        loss_sum, n = 0.0, 0
        for cols, targets in batches:
            loss_sum += ps.optimize(cols, targets)       # loss * batchSize (ParRecModel.scala:477)
            ps.applyOptimizer(a.optim, a.stepSize)
            n += len(targets)
        scores = [(t, ps.predict(c, len(t))) for c, t in batches]
        auc = b.metrics.auc(np.concatenate([t for t, _ in scores]), np.concatenate([p for _, p in scores]))
        print(f"epoch={epoch} loss={loss_sum / n:.6f} auc={auc:.6f} time={int((time.time() - t0) * 1000)}ms")


if __name__ == "__main__":
    main()
