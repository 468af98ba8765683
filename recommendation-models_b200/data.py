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

import ctypes as C

import numpy as np

from . import _lib as L

FORMATS = {"libsvm": 0, "libffm": 1}


def parse_text(text, fmt="libsvm", with_fields=False):
    """Native parser (csrc/parser.cu, `b200rec_parse_samples`) over a whole text blob (bytes or str):
    -> (index, feats, values, targets[, fields]).  Multi-threaded over line ranges.
    Malformed lines raise ValueError naming the line."""
    fmt = fmt.lower()
    if fmt not in FORMATS:
        raise ValueError(f"unknown data format {fmt!r}")
    blob = text.encode() if isinstance(text, str) else bytes(text)
    ns, nz = C.c_int64(0), C.c_int64(0)
    # upper bounds from two memchr-speed scans: a sample per line, a non-zero per "key:value" token
    cap_s = blob.count(b"\n") + 1
    cap_z = blob.count(b":") // (2 if fmt == "libffm" else 1) + 1
    targets = np.empty(cap_s, np.float32)
    index, feats = np.empty(cap_z, np.int32), np.empty(cap_z, np.int32)
    values = np.empty(cap_z, np.float32)
    fields = np.empty(cap_z, np.int32) if with_fields else None
    L.check(L.lib().b200rec_parse_samples(FORMATS[fmt], blob, len(blob), cap_s, cap_z, L.ptr(targets),
                                          L.ptr(index), L.ptr(feats), L.ptr(fields), L.ptr(values),
                                          C.byref(ns), C.byref(nz)))
    targets, index, feats, values = targets[:ns.value], index[:nz.value], feats[:nz.value], values[:nz.value]
    fields = fields[:nz.value] if with_fields else None
    out = (index, feats, values, targets)
    return out + (fields,) if with_fields else out


def parse(lines, fmt="libsvm"):
    """Line-array form of SampleParser.parse (pure Python; `parse_text` is the fast path)."""
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
