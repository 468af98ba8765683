"""Evaluation metric of the reference's driver loop (host side, like the reference).

rec/example/DeepFMLocalExample.scala:45-52 zips (target, prediction) pairs of all batches and calls
Angel's `new AUC().calculate(...)` on the driver; Angel's class is third-party (absent): the
rank-sum (Mann-Whitney) form without tie correction is restated here, parity unpinned.
"""
from __future__ import annotations

import numpy as np


def auc(targets, preds) -> float:
    """Rank-sum AUC.  Labels are thresholded `> 0` like the models do (DeepFM.scala:106)."""
    t = np.asarray(targets) > 0
    p = np.asarray(preds, dtype=np.float64)
    order = np.argsort(p, kind="stable")
    ranks = np.empty(order.size, dtype=np.float64)
    ranks[order] = np.arange(1, order.size + 1)
    npos = int(t.sum())
    nneg = t.size - npos
    if npos == 0 or nneg == 0:
        return float("nan")
    return float((ranks[t].sum() - npos * (npos + 1) / 2.0) / (npos * nneg))
