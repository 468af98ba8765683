"""CPU restatement of the reference hot path (numpy).  TEST INFRASTRUCTURE ONLY.

This file is the *oracle*: a plain-numpy restatement of what
yaochitc/recommendation-models computes on its data-parallel hot path
(embedding lookup, first/second-order FM terms, MLP tower, CIN, DCN cross,
PNN inner products, output head + BCE, per-nnz gradient write-back, hash-map
gradient scatter-add).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The
product (``recommendation-models_b200``) never does.

PARITY UNPINNED: the reference ships no tests, fixtures or golden vectors and
cannot be built here (Scala on a JVM with BigDL 0.9.1 / Angel 2.3.1 / fastutil
8.2.2 as un-vendored Maven dependencies, ``pom.xml:19-53``; no JVM, no jars,
no network).  The arithmetic of the third-party BigDL layers is restated from
their published semantics (Torch7-style ``nn``): Linear ``y = x W^T + b``,
``Sum``/``Mean`` sequential along the dimension, ``Power(2)``, ``CSubTable``,
``CAddTable``, ``CAdd``, ``MulConstant``, ``ReLU``, ``Sigmoid``, ``MM``,
``BCECriterion`` (eps = 1e-12, sizeAverage).  What pins this file instead:
an independently written torch-fp64 autograd implementation of the *papers'*
formulas (``tests/independent.py``), finite differences, hand-computed small
cases, and the committed fixtures in ``tests/golden`` produced by that
independent implementation.

All paths below are relative to
``/root/reference/src/main/scala``:
  nn/  = com/intel/analytics/bigdl/nn/
  rec/ = io/yaochi/recommendation/

Every function takes ``dtype`` (np.float32 = the reference's precision,
np.float64 = the "twin" used to bound rounding noise in tolerance tests).
"""
from __future__ import annotations

import numpy as np

EPS_BCE = 1e-12  # BigDL BCECriterion eps (third-party; restated)


# ----------------------------------------------------------------------------
# Custom BigDL modules of the reference (nn/*.scala)
# ----------------------------------------------------------------------------
def scatter_update_output(inp, index, batch_size, n_output=1):
    """nn/Scatter.scala:17-36.  out[index[i], :] += inp[i, :], sequential in i.

    ``require(index < batchSize)`` (Scatter.scala:29-30) -> ValueError, the
    Python stand-in for IllegalArgumentException.
    """
    inp = np.asarray(inp)
    index = np.asarray(index)
    n = index.shape[0]
    x = inp.reshape(n, n_output)
    if n and index.max() >= batch_size:
        bad = int(index[index >= batch_size][0])
        raise ValueError(f"index should smaller than {batch_size}, but got {bad}")
    out = np.zeros((batch_size, n_output), dtype=inp.dtype)
    # np.add.at is unbuffered and walks i = 0..n-1 in order: the reference's order.
    np.add.at(out, index, x)
    return out


def scatter_update_grad_input(grad_output, index, batch_size):
    """nn/Scatter.scala:38-59.  grad_in[i, :] = grad_output[index[i], :]."""
    index = np.asarray(index)
    if index.size and index.max() >= batch_size:
        bad = int(index[index >= batch_size][0])
        raise ValueError(f"index should smaller than {batch_size}, but got {bad}")
    return np.ascontiguousarray(grad_output[index])


def pnn_pairs(n_fields):
    """rec/model/pnn/ProductEncoder.scala:110-120 (calcIndices): i<j, lexicographic."""
    rows, cols = [], []
    for i in range(n_fields):
        for j in range(i + 1, n_fields):
            rows.append(i)
            cols.append(j)
    return np.asarray(rows, np.int32), np.asarray(cols, np.int32)


def gather_update_output(x, rows, cols):
    """nn/Gather.scala:19-48.  x:[B,F,K] -> (x[:, rows, :], x[:, cols, :]) each [B,P,K]."""
    return np.ascontiguousarray(x[:, rows, :]), np.ascontiguousarray(x[:, cols, :])


def gather_update_grad_input(x_shape, rows, cols, g_row, g_col):
    """nn/Gather.scala:50-78.  Accumulates pair p = 0..P-1 in order; row then col.

    The reference forgets to zero ``gradTensor`` (Gather.scala:63-64, SURVEY B-9);
    it is correct only because modules are rebuilt per call.  Zeroed here.
    """
    g = np.zeros(x_shape, dtype=g_row.dtype)
    for p in range(len(rows)):
        g[:, rows[p], :] += g_row[:, p, :]
        g[:, cols[p], :] += g_col[:, p, :]
    return g


def dotproduct2_update_output(a, b):
    """nn/DotProduct2.scala:16-26.  cmul then sum over dim 3 (sequential in k)."""
    buf = a * b
    out = np.zeros(buf.shape[:2], dtype=buf.dtype)
    for k in range(buf.shape[2]):
        out += buf[:, :, k]
    return out


def dotproduct2_update_grad_input(a, b, grad_output):
    """nn/DotProduct2.scala:28-53.  (b*go, a*go) with go broadcast along k."""
    go = grad_output[:, :, None]
    return b * go, a * go


# ----------------------------------------------------------------------------
# BigDL third-party layers (restated)
# ----------------------------------------------------------------------------
def _seq_sum(x, axis):
    """BigDL Sum(dimension): sequential accumulate along ``axis`` in the tensor dtype."""
    x = np.moveaxis(x, axis, 0)
    acc = np.zeros(x.shape[1:], dtype=x.dtype)
    for i in range(x.shape[0]):
        acc = acc + x[i]
    return acc


def linear_fwd(x, w, b=None):
    """BigDL Linear via rec/util/LayerUtil.scala:7-24: y = x W^T (+ b), W:[out,in]."""
    y = x @ w.T
    if b is not None:
        y = y + b
    return y


def linear_bwd(x, w, gy, with_bias=True):
    """Linear.updateGradInput / accGradParameters: gx = gy W, gW = gy^T x, gb = sum_b gy."""
    gx = gy @ w
    gw = gy.T @ x
    gb = gy.sum(axis=0, dtype=gy.dtype) if with_bias else None
    return gx, gw, gb


def relu(x):
    return np.maximum(x, 0)


def sigmoid(x):
    one = x.dtype.type(1)
    return one / (one + np.exp(-x))


def bce_forward(p, t):
    """BigDL BCECriterion(sizeAverage=true).updateOutput, eps=1e-12 (third-party, restated):
    ``buffer = log(x + eps)``; ``sum += buffer . t``; ``buffer = log((1+eps) - x)``;
    ``sum += buffer . (1 - t)``; ``output = -sum / n``.  ``(1 + eps)`` is rounded to the tensor
    dtype first (== 1.0f in fp32), exactly like ``ev.fromType(1.0 + eps)``."""
    dt = p.dtype.type
    a = np.log(p + dt(EPS_BCE))
    b = np.log(dt(1.0 + EPS_BCE) - p)
    s = float((t * a).sum(dtype=p.dtype)) + float(((dt(1) - t) * b).sum(dtype=p.dtype))
    return dt(-s / p.size)


def bce_backward(p, t):
    """BCECriterion.updateGradInput: -(t - x) / (((1+eps) - x)(x + eps)) * (1/n)."""
    dt = p.dtype.type
    return (p - t) / ((dt(1.0 + EPS_BCE) - p) * (p + dt(EPS_BCE))) * dt(1.0 / p.size)


def head_forward(branches, bias):
    """CAddTable + Sigmoid, e.g. rec/model/deepfm/DeepFM.scala:128-134.

    ``branches``: list of [B,1]; ``bias``: [1] broadcast (SURVEY B-8).
    CAddTable adds left to right.
    """
    logit = branches[0].copy()
    for br in branches[1:]:
        logit = logit + br
    logit = logit + bias.reshape(1, 1)
    return sigmoid(logit)


def head_backward(p, targets):
    """BCE backward then Sigmoid backward (g * (1-p) * p); rec/model/deepfm/DeepFM.scala:105-112.

    targets are thresholded ``label > 0`` (DeepFM.scala:106).  Returns
    (loss, dlogit[B,1], dbias[1]).
    """
    dt = p.dtype.type
    t = (np.asarray(targets).reshape(-1, 1) > 0).astype(p.dtype)
    loss = bce_forward(p, t)
    g = bce_backward(p, t)
    dlogit = g * ((dt(1) - p) * p)
    dbias = np.asarray([dlogit.sum(dtype=p.dtype)], dtype=p.dtype)
    return loss, dlogit, dbias


# ----------------------------------------------------------------------------
# Encoders (rec/model/encoder/*.scala)
# ----------------------------------------------------------------------------
def first_order_fwd(weights, index, batch_size):
    """rec/model/encoder/FirstOrderEncoder.scala:7-17 -> Scatter(batchSize, 1)."""
    return scatter_update_output(weights, index, batch_size, 1)


def first_order_bwd(dlogit, index, batch_size):
    return scatter_update_grad_input(dlogit, index, batch_size).reshape(-1)


def second_order_fwd(embedding, batch_size, n_fields, k):
    """rec/model/encoder/SecondOrderEncoder.scala:19-34.

    Reshape [B,F,K]; DuplicateTable{Sum(2)->Power(2), Power(2)->Sum(2)}; CSubTable;
    Mean(dim 2, squeeze=false); MulConstant(0.5).  NB mean over K (SURVEY B-3).
    """
    dt = embedding.dtype.type
    v = embedding.reshape(batch_size, n_fields, k)
    s = _seq_sum(v, 1)
    sq = s * s
    q = _seq_sum(v * v, 1)
    d = sq - q
    m = _seq_sum(d, 1).reshape(batch_size, 1) / dt(k)
    return m * dt(0.5)


def second_order_bwd(embedding, dlogit, batch_size, n_fields, k):
    """Module-by-module backward of the graph above."""
    dt = embedding.dtype.type
    v = embedding.reshape(batch_size, n_fields, k)
    g = dlogit.reshape(batch_size, 1) * dt(0.5)              # MulConstant
    gd = np.repeat(g / dt(k), k, axis=1)                      # Mean backward [B,K]
    s = _seq_sum(v, 1)
    # branch 1: Sum -> Power(2): d/dv = 2*s*gd broadcast over f
    g1 = np.repeat((dt(2) * s * gd)[:, None, :], n_fields, axis=1)
    # branch 2 (subtracted): Power(2) -> Sum: d/dv = 2*v*(-gd)
    g2 = dt(2) * v * (-gd)[:, None, :]
    return (g1 + g2).reshape(-1)                              # DuplicateTable sums branch grads


def mlp_layout(in_dim, fc_dims, with_head):
    """Offsets of (W,b) blocks for a Linear stack inside ``mats`` (HigherOrderEncoder.scala:46-58)."""
    blocks, off, d = [], 0, in_dim
    dims = list(fc_dims) + ([1] if with_head else [])
    for o in dims:
        blocks.append((off, d, o))
        off += d * o + o
        d = o
    return blocks, off


def mlp_fwd(x, mats, start, in_dim, fc_dims, with_head):
    """rec/model/encoder/HigherOrderEncoder.scala:34-58 (Linear+ReLU)*n [+ Linear->1].

    Returns (out, saved) where saved holds each layer's input and pre-activation.
    """
    blocks, total = mlp_layout(in_dim, fc_dims, with_head)
    saved, h = [], x
    for li, (off, d, o) in enumerate(blocks):
        w = mats[start + off:start + off + d * o].reshape(o, d)
        b = mats[start + off + d * o:start + off + d * o + o]
        y = linear_fwd(h, w, b)
        is_head = with_head and li == len(blocks) - 1
        saved.append((h, w, y, is_head))
        h = y if is_head else relu(y)
    return h, saved, start + total


def mlp_bwd(gout, saved, gmats, start, in_dim, fc_dims, with_head):
    """Backward + rec/util/BackwardUtil.scala:6-31 (grads written at the params' offsets)."""
    blocks, _ = mlp_layout(in_dim, fc_dims, with_head)
    g = gout
    for (off, d, o), (h, w, y, is_head) in zip(reversed(blocks), reversed(saved)):
        if not is_head:
            g = g * (y > 0)
        gx, gw, gb = linear_bwd(h, w, g)
        gmats[start + off:start + off + d * o] = gw.reshape(-1)
        gmats[start + off + d * o:start + off + d * o + o] = gb
        g = gx
    return g


# ----------------------------------------------------------------------------
# mats sizes (getMatsSize of each model)
# ----------------------------------------------------------------------------
def mats_size(kind, n_fields, k, fc_dims=(), cin_dims=(), cross_depth=0):
    """(in,out) pairs exactly as getMatsSize emits them; total = sum(in*out).

    DeepFM.scala:15-20, XDeepFM.scala:15-28, DCN.scala:15-32, PNN.scala:15-25.
    LR / FM have no mats.
    """
    d = n_fields * k
    fc_dims, cin_dims = list(fc_dims), list(cin_dims)
    pairs = []

    def fc(dims):
        for i in range(1, len(dims)):
            pairs.extend([dims[i - 1], dims[i], dims[i], 1])

    if kind == "deepfm":
        fc([d] + fc_dims + [1])
    elif kind == "xdeepfm":
        fc([d] + fc_dims)
        cd = [n_fields] + cin_dims
        for i in range(1, len(cd)):
            pairs.extend([n_fields * cd[i - 1], cd[i], cd[i], 1])
        pairs.extend([sum(cin_dims) + fc_dims[-1], 1])
    elif kind == "dcn":
        for _ in range(cross_depth):
            pairs.extend([d, 1])
        for _ in range(cross_depth):
            pairs.extend([1, 1])
        fc([d] + fc_dims)
        pairs.extend([d + fc_dims[-1], 1])
    elif kind == "pnn":
        p = n_fields * (n_fields - 1) // 2
        pairs.extend([d, fc_dims[0], p, fc_dims[0], 1, 1])
        fc(fc_dims + [1])
    elif kind in ("lr", "fm"):
        pass
    else:
        raise ValueError(kind)
    return pairs


def mats_len(kind, n_fields, k, fc_dims=(), cin_dims=(), cross_depth=0):
    """rec/model/ParRecModel.scala:107-113."""
    p = mats_size(kind, n_fields, k, fc_dims, cin_dims, cross_depth)
    return int(sum(p[i] * p[i + 1] for i in range(0, len(p), 2)))


# ----------------------------------------------------------------------------
# CIN (rec/model/xdeepfm/CINEncoder.scala)
# ----------------------------------------------------------------------------
def cin_fwd(embedding, mats, batch_size, n_fields, k, fc_dims, cin_dims, materialise_z=True):
    """CINEncoder.forward :36-58.  L-layer semantics of SURVEY B-2 (== reference at L=1).

    mats layout (CINEncoder.scala:123-148,173-176): [DNN (W,b)...][CIN (W_l,b_l)...][W_out].
    """
    d = n_fields * k
    r = batch_size * k
    v = embedding.reshape(batch_size, n_fields, k)
    # shapeModule :105-110  -> x0[r=(b,k), f]
    x0 = np.ascontiguousarray(v.transpose(0, 2, 1)).reshape(r, n_fields)
    dnn_out, dnn_saved, off = mlp_fwd(embedding.reshape(batch_size, d), mats, 0, d, fc_dims, False)
    xs, ys, ws = [x0], [], []
    h = n_fields
    for c in cin_dims:
        w = mats[off:off + n_fields * h * c].reshape(c, n_fields * h)
        b = mats[off + n_fields * h * c:off + n_fields * h * c + c]
        off += n_fields * h * c + c
        xl = xs[-1]
        # MM(transB): Z[r, i, j] = x0[r,i] * x^{l-1}[r,j]   (:152)
        z = (x0[:, :, None] * xl[:, None, :]).reshape(r, n_fields * h)
        y = linear_fwd(z, w, b)
        ys.append(y)
        ws.append(w)
        xs.append(relu(y))
        h = c
    # sumModule :159-165: pool over k, concat layers
    pooled = [_seq_sum(x.reshape(batch_size, k, -1), 1) for x in xs[1:]]
    joined = np.concatenate(pooled + [dnn_out], axis=1)
    w_out = mats[off:off + joined.shape[1]].reshape(1, -1)
    out = linear_fwd(joined, w_out)
    saved = dict(x0=x0, xs=xs, ys=ys, ws=ws, dnn_saved=dnn_saved, joined=joined, w_out=w_out, off_out=off)
    return out, saved


def cin_bwd(embedding, gout, saved, mats, batch_size, n_fields, k, fc_dims, cin_dims):
    """CINEncoder.backward :60-103 as a single reverse pass (equal to the :76-86 path sum)."""
    d = n_fields * k
    r = batch_size * k
    gm = np.zeros_like(mats)
    x0, xs, ys, ws = saved["x0"], saved["xs"], saved["ys"], saved["ws"]
    joined, w_out, off_out = saved["joined"], saved["w_out"], saved["off_out"]
    gj, gw_out, _ = linear_bwd(joined, w_out, gout, with_bias=False)
    gm[off_out:off_out + joined.shape[1]] = gw_out.reshape(-1)
    csum = int(sum(cin_dims))
    g_dnn_out = gj[:, csum:]
    gx_dnn = mlp_bwd(g_dnn_out, saved["dnn_saved"], gm, 0, d, fc_dims, False)
    # offsets of CIN blocks
    _, off = mlp_layout(d, fc_dims, False)
    offs, h = [], n_fields
    for c in cin_dims:
        offs.append((off, h, c))
        off += n_fields * h * c + c
        h = c
    gx0 = np.zeros_like(x0)
    g_next = None
    col = csum
    for l in range(len(cin_dims) - 1, -1, -1):
        o, h, c = offs[l]
        col -= c
        gp = gj[:, col:col + c]                                  # pooled grad [B,c]
        gx = np.repeat(gp[:, None, :], k, axis=1).reshape(r, c)  # Sum backward: broadcast over k
        if g_next is not None:
            gx = gx + g_next
        gy = gx * (ys[l] > 0)
        xl = xs[l]
        z = (x0[:, :, None] * xl[:, None, :]).reshape(r, n_fields * h)
        gz, gw, gb = linear_bwd(z, ws[l], gy)
        gm[o:o + n_fields * h * c] = gw.reshape(-1)
        gm[o + n_fields * h * c:o + n_fields * h * c + c] = gb
        gz = gz.reshape(r, n_fields, h)
        gx0 = gx0 + (gz * xl[:, None, :]).sum(axis=2, dtype=gz.dtype)
        g_next = (gz * x0[:, :, None]).sum(axis=1, dtype=gz.dtype)
    gx0 = gx0 + g_next                                           # :85 (layer-1 input is x0)
    ge = gx0.reshape(batch_size, k, n_fields).transpose(0, 2, 1).reshape(batch_size, d)
    ge = ge + gx_dnn                                             # :102
    return ge.reshape(-1), gm


# ----------------------------------------------------------------------------
# DCN cross (rec/model/dcn/CrossEncoder.scala)
# ----------------------------------------------------------------------------
def cross_fwd(embedding, mats, batch_size, n_fields, k, cross_depth, fc_dims):
    """CrossEncoder.forward :40-55.  mats: [w_0..w_{L-1}][c_0..c_{L-1}][DNN][W_out] (:134-185)."""
    d = n_fields * k
    x0 = embedding.reshape(batch_size, d)
    ws = [mats[l * d:(l + 1) * d] for l in range(cross_depth)]
    cs = mats[cross_depth * d:cross_depth * d + cross_depth]
    xs, ss = [x0], []
    for l in range(cross_depth):
        s = xs[-1] @ ws[l].reshape(d, 1)                         # Linear(D->1), no bias
        ss.append(s)
        xs.append(x0 * s + xs[-1] + cs[l])                       # MM, CAddTable, CAdd(scalar)
    dnn_start = cross_depth * d + cross_depth
    dnn_out, dnn_saved, off = mlp_fwd(x0, mats, dnn_start, d, fc_dims, False)
    joined = np.concatenate([xs[-1], dnn_out], axis=1)
    w_out = mats[off:off + joined.shape[1]].reshape(1, -1)
    out = linear_fwd(joined, w_out)
    return out, dict(x0=x0, xs=xs, ss=ss, ws=ws, dnn_saved=dnn_saved, joined=joined, w_out=w_out,
                     off_out=off, dnn_start=dnn_start)


def cross_bwd(gout, saved, mats, batch_size, n_fields, k, cross_depth, fc_dims):
    """CrossEncoder.backward :57-105."""
    d = n_fields * k
    gm = np.zeros_like(mats)
    x0, xs, ss, ws = saved["x0"], saved["xs"], saved["ss"], saved["ws"]
    joined, w_out, off_out = saved["joined"], saved["w_out"], saved["off_out"]
    gj, gw_out, _ = linear_bwd(joined, w_out, gout, with_bias=False)
    gm[off_out:off_out + joined.shape[1]] = gw_out.reshape(-1)
    gx_dnn = mlp_bwd(gj[:, d:], saved["dnn_saved"], gm, saved["dnn_start"], d, fc_dims, False)
    g = gj[:, :d]
    gx0 = np.zeros_like(x0)
    for l in range(cross_depth - 1, -1, -1):
        gm[cross_depth * d + l] = g.sum(dtype=g.dtype)           # CAdd(1): gradBias = sum of all
        gx0 = gx0 + g * ss[l]                                    # MM backward wrt x0
        gs = (g * x0).sum(axis=1, keepdims=True, dtype=g.dtype)  # MM backward wrt s
        gm[l * d:(l + 1) * d] = (gs.T @ xs[l]).reshape(-1)       # Linear gradWeight
        g = g + gs @ ws[l].reshape(1, d)                         # CAddTable + Linear gradInput
    gx0 = gx0 + gx_dnn + g                                       # :102
    return gx0.reshape(-1), gm


# ----------------------------------------------------------------------------
# PNN product layer (rec/model/pnn/ProductEncoder.scala)
# ----------------------------------------------------------------------------
def product_fwd(embedding, mats, batch_size, n_fields, k, out_dim):
    """ProductEncoder.forward :34-41.  mats: [W_z(O x D)][W_p(O x P)][c] (:78-108)."""
    d = n_fields * k
    rows, cols = pnn_pairs(n_fields)
    p = len(rows)
    x = embedding.reshape(batch_size, d)
    wz = mats[:d * out_dim].reshape(out_dim, d)
    wp = mats[d * out_dim:d * out_dim + p * out_dim].reshape(out_dim, p)
    c = mats[d * out_dim + p * out_dim]
    lz = linear_fwd(x, wz)
    v = embedding.reshape(batch_size, n_fields, k)
    a, b = gather_update_output(v, rows, cols)
    ip = dotproduct2_update_output(a, b)
    lp = linear_fwd(ip, wp)
    pre = (lz + lp) + c
    end = d * out_dim + p * out_dim + 1
    return relu(pre), dict(x=x, v=v, a=a, b=b, ip=ip, pre=pre, wz=wz, wp=wp, rows=rows, cols=cols), end


def product_bwd(gout, saved, gm, batch_size, n_fields, k, out_dim):
    """ProductEncoder.backward :43-70."""
    d = n_fields * k
    p = len(saved["rows"])
    g = gout * (saved["pre"] > 0)
    gm[d * out_dim + p * out_dim] = g.sum(dtype=g.dtype)
    gx, gwz, _ = linear_bwd(saved["x"], saved["wz"], g, with_bias=False)
    gm[:d * out_dim] = gwz.reshape(-1)
    gip, gwp, _ = linear_bwd(saved["ip"], saved["wp"], g, with_bias=False)
    gm[d * out_dim:d * out_dim + p * out_dim] = gwp.reshape(-1)
    ga, gb = dotproduct2_update_grad_input(saved["a"], saved["b"], gip)
    gv = gather_update_grad_input(saved["v"].shape, saved["rows"], saved["cols"], ga, gb)
    return gx + gv.reshape(batch_size, d)


# ----------------------------------------------------------------------------
# Models: Internal<M>Model.forward / .backward with the flat in/out-aliased ABI
# ----------------------------------------------------------------------------
class Model:
    """One class for the six model kinds; mirrors Internal*Model (flat arrays in, grads in place).

    lr      rec/model/lr/LR.scala:42-90
    fm      (no class in the reference; SURVEY B-1: DeepFM minus HigherOrderEncoder)
    deepfm  rec/model/deepfm/DeepFM.scala:51-125
    xdeepfm rec/model/xdeepfm/XDeepFM.scala:58-126
    dcn     rec/model/dcn/DCN.scala:62-130
    pnn     rec/model/pnn/PNN.scala:56-132
    """

    def __init__(self, kind, n_fields=0, embedding_dim=0, fc_dims=(), cin_dims=(), cross_depth=0,
                 dtype=np.float32):
        self.kind, self.f, self.k = kind, n_fields, embedding_dim
        self.fc, self.cin, self.depth = list(fc_dims), list(cin_dims), cross_depth
        self.dtype = dtype

    def mats_size(self):
        return mats_size(self.kind, self.f, self.k, self.fc, self.cin, self.depth)

    def mats_len(self):
        return mats_len(self.kind, self.f, self.k, self.fc, self.cin, self.depth)

    def _cast(self, a):
        return None if a is None else np.asarray(a, dtype=self.dtype)

    def _run(self, batch_size, index, weights, bias, embedding, mats, targets):
        kd, f, k, dt = self.kind, self.f, self.k, self.dtype
        w = self._cast(weights)
        b = self._cast(bias)
        e = self._cast(embedding)
        m = self._cast(mats)
        index = np.asarray(index, dtype=np.int64)
        if kd != "lr" and e.size != batch_size * f * k:
            # Reshape(Array(batchSize, nFields, embeddingDim)) fails in BigDL
            raise ValueError(f"embedding has {e.size} elements, expected {batch_size}*{f}*{k}")
        first = first_order_fwd(w, index, batch_size)
        branches = [first]
        ctx = {}
        if kd in ("fm", "deepfm"):
            branches.append(second_order_fwd(e, batch_size, f, k))
        if kd == "deepfm":
            hi, ctx["mlp"], _ = mlp_fwd(e.reshape(batch_size, f * k), m, 0, f * k, self.fc, True)
            branches.append(hi)
        elif kd == "xdeepfm":
            out, ctx["cin"] = cin_fwd(e, m, batch_size, f, k, self.fc, self.cin)
            branches.append(out)
        elif kd == "dcn":
            out, ctx["cross"] = cross_fwd(e, m, batch_size, f, k, self.depth, self.fc)
            branches.append(out)
        elif kd == "pnn":
            h, ctx["prod"], end = product_fwd(e, m, batch_size, f, k, self.fc[0])
            out, ctx["mlp"], _ = mlp_fwd(h, m, end, self.fc[0], self.fc[1:], True)
            ctx["end"] = end
            branches.append(out)
        p = head_forward(branches, b)
        self.last_ctx = ctx
        if targets is None:
            return p.reshape(-1).astype(dt)
        loss, dlogit, dbias = head_backward(p, self._cast(targets))
        gw = first_order_bwd(dlogit, index, batch_size)
        ge, gm = None, None
        if kd in ("fm", "deepfm"):
            ge = second_order_bwd(e, dlogit, batch_size, f, k)
        if kd == "deepfm":
            gm = np.zeros_like(m)
            gx = mlp_bwd(dlogit, ctx["mlp"], gm, 0, f * k, self.fc, True)
            ge = ge + gx.reshape(-1)                       # rec/util/GradUtil.scala:23-34
        elif kd == "xdeepfm":
            ge, gm = cin_bwd(e, dlogit, ctx["cin"], m, batch_size, f, k, self.fc, self.cin)
        elif kd == "dcn":
            ge, gm = cross_bwd(dlogit, ctx["cross"], m, batch_size, f, k, self.depth, self.fc)
        elif kd == "pnn":
            gm = np.zeros_like(m)
            gh = mlp_bwd(dlogit, ctx["mlp"], gm, ctx["end"], self.fc[0], self.fc[1:], True)
            ge = product_bwd(gh, ctx["prod"], gm, batch_size, f, k, self.fc[0]).reshape(-1)
        return loss, gw, dbias, ge, gm

    def relu_preactivations(self, batch_size):
        """Pre-activations of every ReLU of the last forward, one [B, units] array per layer.
        Tests use them to keep inputs away from the ReLU kinks, where ANY two fp32 implementations
        (MKL vs OpenBLAS vs this GPU path) may legitimately disagree about the sign."""
        ctx, out = self.last_ctx, []
        def mlp(saved):
            for (_, _, y, is_head) in saved:
                if not is_head:
                    out.append(y.reshape(batch_size, -1))
        if "mlp" in ctx:
            mlp(ctx["mlp"])
        if "cin" in ctx:
            mlp(ctx["cin"]["dnn_saved"])
            out.extend(y.reshape(batch_size, -1) for y in ctx["cin"]["ys"])
        if "cross" in ctx:
            mlp(ctx["cross"]["dnn_saved"])
        if "prod" in ctx:
            out.append(ctx["prod"]["pre"].reshape(batch_size, -1))
        return out

    def forward(self, batch_size, index, weights, bias, embedding=None, mats=None):
        """-> float[B] sigmoid(logit)   (e.g. DeepFM.scala:54-81)."""
        return self._run(batch_size, index, weights, bias, embedding, mats, None)

    def backward(self, batch_size, index, weights, bias, embedding=None, mats=None, targets=None):
        """-> loss; overwrites weights/bias/embedding/mats with their gradients
        (rec/util/GradUtil.scala:7-42, rec/util/BackwardUtil.scala:6-42)."""
        loss, gw, gb, ge, gm = self._run(batch_size, index, weights, bias, embedding, mats, targets)
        weights[...] = gw
        bias[...] = gb
        if ge is not None:
            embedding[...] = ge
        if gm is not None:
            mats[...] = gm
        return float(loss)


def away_from_kinks(model64, B, F, K, index, w, bias, emb, mats, margin=2e-4, keep=None):
    """Select `keep` of the B candidate samples whose ReLU pre-activations (fp64 oracle) all stay at
    least `margin` (relative to the layer's largest pre-activation) away from zero.

    ReLU is discontinuous in its derivative: a pre-activation within rounding distance of zero can
    come out on either side in two correct fp32 implementations (the reference's MKL sgemm and the
    oracle's OpenBLAS already differ there), and one flipped unit changes a weight-gradient row by
    O(1/sqrt(B)) -- far above any arithmetic tolerance.  Parity of the ARITHMETIC is therefore
    tested on inputs that avoid the kinks; samples are independent, so dropping some is harmless.
    Returns (index, w, emb, sample_ids) for the kept samples."""
    keep = keep or B
    c64 = lambda a: None if a is None else np.asarray(a, np.float64)
    model64.forward(B, index, c64(w), c64(bias), c64(emb), c64(mats))
    pre = model64.relu_preactivations(B)
    if not pre:
        ids = np.arange(keep)
    else:
        m = np.full(B, np.inf)
        for z in pre:
            m = np.minimum(m, np.abs(z).min(axis=1) / np.abs(z).max())
        order = np.argsort(-m)
        ids = np.sort(order[:keep])
        assert m[ids].min() >= margin, f"only {(m >= margin).sum()} of {B} candidates clear the margin; need {keep}"
    w2 = w.reshape(B, F)[ids].reshape(-1)
    e2 = None if emb is None else emb.reshape(B, F * K)[ids].reshape(-1)
    idx2 = np.repeat(np.arange(keep, dtype=np.int32), F)
    return idx2, np.ascontiguousarray(w2), (None if e2 is None else np.ascontiguousarray(e2)), ids


# ----------------------------------------------------------------------------
# Gather / scatter-add around the model (rec/model/ParRecModel.scala:270-345)
# ----------------------------------------------------------------------------
def make_embeddings(table, feats):
    """ParRecModel.makeEmbeddings :300-306 -> float[N*K], row-major [N,K] (pure copy)."""
    return np.ascontiguousarray(table[np.asarray(feats, np.int64)]).reshape(-1)


def make_weights(wtable, feats):
    """ParRecModel.makeWeights :279-284."""
    return np.ascontiguousarray(wtable[np.asarray(feats, np.int64)])


def distinct_int_indices(feats):
    """ParRecModel.distinctIntIndices :337-345.  Reference order is hash order (unspecified);
    compared as a sorted set."""
    return np.unique(np.asarray(feats).astype(np.int32))


def make_embedding_grad(buf, feats, k):
    """ParRecModel.makeEmbeddingGrad :316-328.  Per distinct id, rows summed in nnz order
    i = 0..N-1 in fp32 (Int2FloatOpenHashMap.addTo).  Returns (sorted ids, G[U,K])."""
    feats = np.asarray(feats, np.int64)
    ids, inv = np.unique(feats, return_inverse=True)
    g = np.zeros((ids.size, k), dtype=buf.dtype)
    np.add.at(g, inv, buf.reshape(-1, k))     # unbuffered, in order of i
    return ids.astype(np.int32), g


def make_weights_grad(buf, feats):
    """ParRecModel.makeWeightsGrad :293-298."""
    feats = np.asarray(feats, np.int64)
    ids, inv = np.unique(feats, return_inverse=True)
    g = np.zeros(ids.size, dtype=buf.dtype)
    np.add.at(g, inv, buf)
    return ids.astype(np.int32), g


# ----------------------------------------------------------------------------
# Optimizers (rec/optim/Async*.scala -> Angel Async*Func PSFs, third-party: textbook forms, unpinned)
# ----------------------------------------------------------------------------
def optimizer_update(kind, w, g, state, lr, p1=None, p2=None, step=1, eps=1e-7):
    """One update of the values `w` (any shape) with gradient `g`; `state` is a dict holding the slots.
    sgd; momentum (p1 = 0.9); adagrad (p1 = factor 0.9, EMA of g^2); adam (p1 = gamma 0.99, p2 = beta 0.9).
    fp32 arithmetic like the GPU kernel (csrc/optim.cu)."""
    f = np.float32
    w, g = w.astype(f), g.astype(f)
    lr = f(lr)
    if kind == "sgd":
        return w - lr * g
    if kind == "momentum":
        mu = f(0.9 if p1 is None else p1)
        v = mu * state.get("s1", np.zeros_like(w)) + g
        state["s1"] = v
        return w - lr * v
    if kind == "adagrad":
        fac = f(0.9 if p1 is None else p1)
        s = fac * state.get("s1", np.zeros_like(w)) + (f(1) - fac) * g * g
        state["s1"] = s
        return w - lr * g / (np.sqrt(s) + f(eps))
    if kind == "adam":
        gamma, beta = f(0.99 if p1 is None else p1), f(0.9 if p2 is None else p2)
        m = beta * state.get("s1", np.zeros_like(w)) + (f(1) - beta) * g
        v = gamma * state.get("s2", np.zeros_like(w)) + (f(1) - gamma) * g * g
        state["s1"], state["s2"] = m, v
        c1 = f(1.0 / (1.0 - float(beta) ** step))
        c2 = f(1.0 / (1.0 - float(gamma) ** step))
        return w - lr * (m * c1) / (np.sqrt(v * c2) + f(eps))
    raise ValueError(kind)


# ----------------------------------------------------------------------------
# AUC (Angel AUC().calculate, rec/example/DeepFMLocalExample.scala:45-52) -- rank-sum
# ----------------------------------------------------------------------------
def auc(targets, preds):
    """Rank-sum AUC; ties broken by sort order like a plain sort-based implementation
    (Angel's is third-party; restated, no tie correction)."""
    t = np.asarray(targets) > 0
    order = np.argsort(np.asarray(preds), kind="stable")
    ranks = np.empty(len(order), dtype=np.float64)
    ranks[order] = np.arange(1, len(order) + 1)
    npos = int(t.sum())
    nneg = len(t) - npos
    if npos == 0 or nneg == 0:
        return float("nan")
    return float((ranks[t].sum() - npos * (npos + 1) / 2.0) / (npos * nneg))
