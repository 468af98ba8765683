"""Batch builder: libsvm / libffm text -> (index, feats, values, targets), the arrays the path takes.

Mirrors rec/data/SampleParser.scala (under /root/reference/src/main/scala/io/yaochi/recommendation):
  :15-21  parse(lines, "libsvm" | "libffm")
  :23-51  parseLIBSVM:  "<label> <key>:<value> ..."            keys are 1-based in files -> key - 1 (:37)
  :53-85  parseLIBFFM:  "<label> <field>:<key>:<value> ..."    (:69)
Rows are emitted sample-major: index[i] = sample of non-zero i (the COO row, :30-46).  Feature VALUES
are parsed but the models ignore them (RecModel.scala:135,151 binds "values" to the row indices and
never reads it: SURVEY B-5); they are returned for completeness.
Host-side text handling, like the reference's (it runs in the Spark task, not on the PS or in BigDL).
"""
from __future__ import annotations

import numpy as np


def parse(lines, fmt="libsvm"):
    fmt = fmt.lower()
    if fmt == "libsvm":
        return parse_libsvm(lines)
    if fmt == "libffm":
        return parse_libffm(lines)
    raise ValueError(f"unknown data format {fmt!r}")   # SampleParser.parse raises AngelException


def _finish(rows, cols, vals, targets):
    return (np.asarray(rows, np.int32), np.asarray(cols, np.int32), np.asarray(vals, np.float32),
            np.asarray(targets, np.float32))


def parse_libsvm(lines):
    rows, cols, vals, targets = [], [], [], []
    for r, line in enumerate(l for l in lines if l.strip()):
        parts = line.split()
        targets.append(float(parts[0]))
        for kv in parts[1:]:
            k, v = kv.split(":")
            rows.append(r)
            cols.append(int(k) - 1)
            vals.append(float(v))
    return _finish(rows, cols, vals, targets)


def parse_libffm(lines):
    rows, cols, vals, targets = [], [], [], []
    for r, line in enumerate(l for l in lines if l.strip()):
        parts = line.split()
        targets.append(float(parts[0]))
        for fkv in parts[1:]:
            _, k, v = fkv.split(":")
            rows.append(r)
            cols.append(int(k) - 1)
            vals.append(float(v))
    return _finish(rows, cols, vals, targets)


def to_libsvm(feats, targets, n_fields):
    """Inverse of parse_libsvm for synthetic batches (one id per field): text lines, 1-based keys."""
    f = np.asarray(feats).reshape(-1, n_fields)
    return [f"{int(t)} " + " ".join(f"{int(k) + 1}:1" for k in row) for t, row in zip(targets, f)]
