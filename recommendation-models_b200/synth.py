"""Synthetic Criteo-shaped inputs: 39 fields, one id per field per
sample, global ids = field offset + power-law local id.  Everything is a pure function of
(seed, counters) through splitmix64, so the host (numpy) and the device
(csrc/table.cu: b200rec_table_init_uniform) produce bit-identical values and any row of a
100M-row table can be regenerated without materialising the table on the host.

The local id is drawn octave-uniformly (pick a bit length e uniformly in 0..floor(log2 V),
then a uniform id in [2^e - 1, 2^(e+1) - 2]) -- a log-uniform (Zipf s~1) law like the
exp(u ln V) form in SURVEY 8d, but integer-only, so it does not depend on libm rounding.
"""
from __future__ import annotations

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
GOLDEN = np.uint64(0x9E3779B97F4A7C15)


def splitmix64(x):
    """splitmix64 finaliser on uint64 arrays (wrap-around arithmetic)."""
    with np.errstate(over="ignore"):
        z = (np.asarray(x, dtype=np.uint64) + GOLDEN) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def u01_24(h):
    """Top 24 bits of a hash as an exactly-representable float32 in [0,1)."""
    return ((h >> np.uint64(40)).astype(np.float32)) * np.float32(1.0 / 16777216.0)


def hash_uniform(seed, counters, lo, hi):
    """float32 lo + (hi-lo)*u, one fp32 multiply then one fp32 add (no fma), u = u01_24(hash).

    Mirrors ``hash_uniform`` in csrc/table.cu bit for bit.
    """
    with np.errstate(over="ignore"):
        h = splitmix64(np.uint64(seed) * GOLDEN + np.asarray(counters, dtype=np.uint64))
    u = u01_24(h)
    return (u * np.float32(np.float32(hi) - np.float32(lo))).astype(np.float32) + np.float32(lo)


def field_layout(input_dim, n_fields):
    """Per-field vocab floor(inputDim/F) (remainder to the last field) and global offsets."""
    v = np.full(n_fields, input_dim // n_fields, dtype=np.int64)
    v[-1] += input_dim - int(v.sum())
    off = np.concatenate([[0], np.cumsum(v)[:-1]]).astype(np.int64)
    return off, v


def make_feats(seed, step, batch_size, n_fields, input_dim):
    """-> (index int32[N], feats int32[N]); index[i] = i // F (sample-major, the order
    SampleParser.parseLIBSVM emits, rec/data/SampleParser.scala:30-46)."""
    off, voc = field_layout(input_dim, n_fields)
    b = np.arange(batch_size, dtype=np.uint64)[:, None]
    f = np.arange(n_fields, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        ctr = (np.uint64(step) * np.uint64(batch_size) + b) * np.uint64(n_fields) + f
        h1 = splitmix64(np.uint64(seed) * GOLDEN + ctr)
        h2 = splitmix64(h1)
    nbits = np.floor(np.log2(voc)).astype(np.uint64)[None, :]          # L = floor(log2 V)
    e = h1 % (nbits + np.uint64(1))
    one = np.uint64(1)
    local = ((one << e) - one) + (h2 & ((one << e) - one))
    local = np.minimum(local, (voc[None, :] - 1).astype(np.uint64)).astype(np.int64)
    feats = (off[None, :] + local).astype(np.int32).reshape(-1)
    index = np.repeat(np.arange(batch_size, dtype=np.int32), n_fields)
    return index, feats


def make_targets(seed, feats, batch_size, n_fields):
    """Bernoulli(sigmoid(teacher logit)) labels from a hidden LR teacher (0/1 floats)."""
    tw = hash_uniform(seed + 7, feats.astype(np.uint64), -1.0, 1.0).reshape(batch_size, n_fields)
    z = tw.astype(np.float64).sum(1) * (2.0 / np.sqrt(n_fields))
    p = 1.0 / (1.0 + np.exp(-z))
    u = hash_uniform(seed + 11, np.arange(batch_size, dtype=np.uint64) + np.uint64(1 << 40), 0.0, 1.0)
    return (u < p).astype(np.float32)


def table_rows(seed, ids, k, lo=-0.05, hi=0.05):
    """Rows ``ids`` of the hash-initialised embedding table: value(row, col) = hash_uniform(seed,
    row*k + col)."""
    ids = np.asarray(ids, dtype=np.uint64)
    ctr = ids[:, None] * np.uint64(k) + np.arange(k, dtype=np.uint64)[None, :]
    return hash_uniform(seed, ctr, lo, hi)


def wtable_rows(seed, ids, lo=-0.05, hi=0.05):
    """First-order weight of each id (a K=1 table with its own stream)."""
    return hash_uniform(seed + 1, np.asarray(ids, dtype=np.uint64), lo, hi)


def init_mats(seed, pairs):
    """Dense params: every (in,out) block of getMatsSize ~ U(+-1/sqrt(in)); (out,1) bias blocks
    and scalar blocks 0.  ``pairs`` is the flat getMatsSize list."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(0, len(pairs), 2):
        a, b = int(pairs[i]), int(pairs[i + 1])
        if b == 1 and i >= 2 and int(pairs[i - 1]) == a:      # the (out,1) bias block after a (in,out) block
            out.append(np.zeros(a, dtype=np.float32))
        elif a == 1 and b == 1:
            out.append(np.zeros(1, dtype=np.float32))
        else:
            s = 1.0 / np.sqrt(a)
            out.append(rng.uniform(-s, s, size=a * b).astype(np.float32))
    return np.concatenate(out) if out else np.zeros(0, np.float32)
