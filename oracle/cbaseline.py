"""CPU BASELINE of the reference path.  TEST + BENCH INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.build()`` (which merely compiles it) and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module; the product never does.

What it times, following rec/model/ParRecModel.scala:439-478 (optimizeBiasWeightEmbeddingMats) minus
the PS RPC, op for op and *keeping the reference's costs*:

  gather        makeEmbeddings / makeWeights (:300-306, :279-284): N*K + N ``get`` calls on K + 1
                hash-backed sparse vectors -- ``oracle/cbaseline.cpp`` (open addressing, fastutil's
                algorithm restated)
  mats_copy     every Linear re-created from ``mats`` per call (rec/util/LayerUtil.scala:13-14) and the
                gradients copied back (rec/util/BackwardUtil.scala:18,26)
  dense         Internal<M>Model.backward (forward inside, e.g. deepfm/DeepFM.scala:93-124): one pass
                and one temporary per BigDL module, CIN with the materialised outer product Z
                (xdeepfm/CINEncoder.scala:150-157), sgemm through MKL (torch-CPU ``addmm`` /
                ``mm`` / ``bmm`` -- the BLAS family BigDL links)
  scatter_add   makeEmbeddingGrad / makeWeightsGrad (:316-328, :293-298): N*K + N ``addTo`` calls

Two variants: ``threads=1`` (Spark ``local[1]`` + BigDL's default single MKL thread:
the faithful one) and ``threads=all`` (generous).  Thread counts are set explicitly through
``torch.set_num_threads`` and the C++ calls' ``threads`` argument, so an inherited
``OMP_NUM_THREADS=1`` (torchrun sets it) cannot change them.

The dense restatement below mirrors ``oracle/refport.py`` function for function (tests compare the
two); dcn / pnn reuse refport's numpy code for the dense part.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cbaseline.cpp")
LIB = os.path.join(HERE, "libcbaseline.so")
_lib = None


def build(force=False):
    """g++ -O3 -march=native -fopenmp -> oracle/libcbaseline.so (git-ignored, travels with gpurun)."""
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    # -march=native would not survive the trip to the GPU box's CPU; x86-64-v3 (AVX2) is common ground
    cmd = ["g++", "-O3", "-march=x86-64-v3", "-fopenmp", "-shared", "-fPIC", "-std=c++17", SRC, "-o", LIB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"g++ failed for {SRC}:\n{r.stderr}")
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
        L.cb_pull.restype = vp
        L.cb_pull.argtypes = [i32, i64, vp, vp, vp, i32]
        L.cb_pull_free.argtypes = [vp]
        L.cb_make_embeddings.argtypes = [vp, i64, vp, vp, i32]
        L.cb_make_weights.argtypes = [vp, i64, vp, vp]
        L.cb_distinct.restype = i64
        L.cb_distinct.argtypes = [i64, vp, vp]
        L.cb_scatter_add.argtypes = [i32, i64, vp, vp, vp, i64, i64, vp, vp, vp, i32]
        L.cb_last_addto_seconds.restype = C.c_double
        L.cb_max_threads.restype = i32
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def host_threads():
    """Usable host cores (the affinity mask, not the box's total)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ---- gather / scatter-add through the hash maps -------------------------------------------------------
class Pulled:
    """The K + 1 hash-backed sparse vectors pullEmbeddings / pullWeights return (untimed: PS RPC)."""

    def __init__(self, ids, rows, w, threads=1):
        ids = np.ascontiguousarray(ids, np.int32)
        self.K = 0 if rows is None else int(rows.shape[1])
        self.U = int(ids.size)
        rows = None if rows is None else np.ascontiguousarray(rows, np.float32)
        w = None if w is None else np.ascontiguousarray(w, np.float32)
        self.h = lib().cb_pull(self.K, self.U, _p(ids), _p(rows), _p(w), threads)

    def close(self):
        if self.h:
            lib().cb_pull_free(self.h)
            self.h = None

    def __del__(self):
        self.close()


def make_embeddings(pulled, feats, threads=1):
    feats = np.ascontiguousarray(feats, np.int32)
    buf = np.empty(feats.size * pulled.K, np.float32)
    lib().cb_make_embeddings(pulled.h, feats.size, _p(feats), _p(buf), threads)
    return buf


def make_weights(pulled, feats):
    feats = np.ascontiguousarray(feats, np.int32)
    buf = np.empty(feats.size, np.float32)
    lib().cb_make_weights(pulled.h, feats.size, _p(feats), _p(buf))
    return buf


def distinct(feats):
    feats = np.ascontiguousarray(feats, np.int32)
    out = np.empty(feats.size, np.int32)
    u = lib().cb_distinct(feats.size, _p(feats), _p(out))
    return out[:u].copy()


def scatter_add(feats, emb_grad, w_grad, k, ids=None, threads=1, want_out=True):
    """-> (ids sorted, G[U,K], gw[U], seconds spent in the addTo loops)."""
    feats = np.ascontiguousarray(feats, np.int32)
    if ids is None:
        ids = np.unique(feats).astype(np.int32)
    U = ids.size
    G = np.empty((U, k), np.float32) if (want_out and emb_grad is not None) else None
    gw = np.empty(U, np.float32) if (want_out and w_grad is not None) else None
    eg = None if emb_grad is None else np.ascontiguousarray(emb_grad, np.float32)
    wg = None if w_grad is None else np.ascontiguousarray(w_grad, np.float32)
    lib().cb_scatter_add(k if eg is not None else 0, feats.size, _p(feats), _p(eg), _p(wg), U, U, _p(ids),
                         _p(G), _p(gw), threads)
    return ids, G, gw, float(lib().cb_last_addto_seconds())


# ---- the dense part through torch-CPU (MKL) --------------------------------------------------------------
def _t():
    import torch
    return torch


class TorchModel:
    """Internal{LR,FM,DeepFM,XDeepFM}Model.backward with torch-CPU tensors: the same passes and
    temporaries as oracle/refport.py (which cites the Scala line by line), sgemm by MKL."""

    KINDS = ("lr", "fm", "deepfm", "xdeepfm")

    def __init__(self, kind, n_fields, k, fc_dims=(), cin_dims=()):
        assert kind in self.KINDS
        self.kind, self.f, self.k = kind, n_fields, k
        self.fc, self.cin = list(fc_dims), list(cin_dims)
        self.t_copy = 0.0

    # rec/util/LayerUtil.scala:7-24: a fresh Linear per call, weights copied out of mats
    def _linear_params(self, mats, off, d, o, bias=True):
        torch = _t()
        t0 = time.perf_counter()
        w = mats[off:off + d * o].reshape(o, d).clone()
        b = mats[off + d * o:off + d * o + o].clone() if bias else None
        self.t_copy += time.perf_counter() - t0
        return w, b

    def _mlp_fwd(self, x, mats, start, in_dim, dims, with_head):
        torch = _t()
        saved, h, off, d = [], x, start, in_dim
        all_dims = list(dims) + ([1] if with_head else [])
        for li, o in enumerate(all_dims):
            w, b = self._linear_params(mats, off, d, o)
            y = torch.addmm(b, h, w.t())
            is_head = with_head and li == len(all_dims) - 1
            saved.append((h, w, y, is_head, off, d, o))
            h = y if is_head else torch.relu(y)
            off += d * o + o
            d = o
        return h, saved, off

    def _mlp_bwd(self, gout, saved, gmats):
        torch = _t()
        g = gout
        for (h, w, y, is_head, off, d, o) in reversed(saved):
            if not is_head:
                g = g * (y > 0)
            gx = torch.mm(g, w)
            gw = torch.mm(g.t(), h)
            gb = g.sum(dim=0)
            t0 = time.perf_counter()
            gmats[off:off + d * o] = gw.reshape(-1)      # BackwardUtil.linearBackward :6-31
            gmats[off + d * o:off + d * o + o] = gb
            self.t_copy += time.perf_counter() - t0
            g = gx
        return g

    def backward(self, B, index, weights, bias, embedding, mats, targets):
        """numpy arrays in; overwrites them with gradients like the reference; returns the loss."""
        torch = _t()
        kd, F, K = self.kind, self.f, self.k
        self.t_copy = 0.0
        tw = torch.from_numpy(weights)
        idx = torch.from_numpy(np.asarray(index, np.int64))
        # FirstOrderEncoder -> Scatter(batchSize, 1)  (nn/Scatter.scala:17-36)
        first = torch.zeros(B, 1).index_add_(0, idx, tw.reshape(-1, 1))
        branches = [first]
        tm = torch.from_numpy(mats) if mats is not None else None
        gm = torch.zeros_like(tm) if tm is not None else None
        if kd != "lr":
            te = torch.from_numpy(embedding)
            v = te.reshape(B, F, K)
        if kd in ("fm", "deepfm"):
            # SecondOrderEncoder.scala:19-34: Sum, Power, Power, Sum, CSubTable, Mean, MulConstant
            s = v.sum(dim=1)
            sq = s * s
            vv = v * v
            q = vv.sum(dim=1)
            dd = sq - q
            second = dd.mean(dim=1, keepdim=True) * 0.5
            branches.append(second)
        if kd == "deepfm":
            hi, mlp_saved, _ = self._mlp_fwd(te.reshape(B, F * K), tm, 0, F * K, self.fc, True)
            branches.append(hi)
        elif kd == "xdeepfm":
            R = B * K
            x0 = v.transpose(1, 2).contiguous().reshape(R, F)           # shapeModule :105-110
            dnn_out, dnn_saved, off = self._mlp_fwd(te.reshape(B, F * K), tm, 0, F * K, self.fc, False)
            xs, ys, ws, offs = [x0], [], [], []
            h = F
            for c in self.cin:
                w, b = self._linear_params(tm, off, F * h, c)
                offs.append((off, h, c))
                off += F * h * c + c
                z = torch.bmm(x0.unsqueeze(2), xs[-1].unsqueeze(1)).reshape(R, F * h)   # MM(transB) :152
                y = torch.addmm(b, z, w.t())
                ys.append(y)
                ws.append(w)
                xs.append(torch.relu(y))
                h = c
            pooled = [x.reshape(B, K, -1).sum(dim=1) for x in xs[1:]]    # sumModule :159-165
            joined = torch.cat(pooled + [dnn_out], dim=1)
            w_out, _ = self._linear_params(tm, off, joined.shape[1], 1, bias=False)
            branches.append(torch.mm(joined, w_out.t()))
            off_out = off
        # CAddTable + Sigmoid + BCECriterion (DeepFM.scala:105-117,127-134)
        logit = branches[0].clone()
        for br in branches[1:]:
            logit = logit + br
        logit = logit + torch.from_numpy(bias).reshape(1, 1)
        p = torch.sigmoid(logit)
        t = (torch.from_numpy(targets).reshape(-1, 1) > 0).float()
        eps = 1e-12
        loss = -((t * torch.log(p + eps)).sum() + ((1 - t) * torch.log((1.0 + eps) - p)).sum()) / B
        g = (p - t) / (((1.0 + eps) - p) * (p + eps)) * (1.0 / B)
        dlogit = g * ((1 - p) * p)
        dbias = dlogit.sum().reshape(1)
        gw = dlogit[idx].reshape(-1)                                     # Scatter backward :38-59
        ge = None
        if kd in ("fm", "deepfm"):
            gg = dlogit * 0.5
            gd = (gg / K).repeat(1, K)
            g1 = (2 * s * gd).unsqueeze(1).repeat(1, F, 1)
            g2 = 2 * v * (-gd).unsqueeze(1)
            ge = (g1 + g2).reshape(-1)
        if kd == "deepfm":
            gx = self._mlp_bwd(dlogit, mlp_saved, gm)
            ge = ge + gx.reshape(-1)                                      # GradUtil.scala:23-34
        elif kd == "xdeepfm":
            gj = torch.mm(dlogit, w_out)
            gm[off_out:off_out + joined.shape[1]] = torch.mm(dlogit.t(), joined).reshape(-1)
            csum = sum(self.cin)
            gx_dnn = self._mlp_bwd(gj[:, csum:], dnn_saved, gm)
            gx0 = torch.zeros_like(x0)
            g_next, col = None, csum
            for l in range(len(self.cin) - 1, -1, -1):
                o, h, c = offs[l]
                col -= c
                gx = gj[:, col:col + c].unsqueeze(1).repeat(1, K, 1).reshape(R, c)
                if g_next is not None:
                    gx = gx + g_next
                gy = gx * (ys[l] > 0)
                xl = xs[l]
                z = torch.bmm(x0.unsqueeze(2), xl.unsqueeze(1)).reshape(R, F * h)
                gz = torch.mm(gy, ws[l]).reshape(R, F, h)
                gm[o:o + F * h * c] = torch.mm(gy.t(), z).reshape(-1)
                gm[o + F * h * c:o + F * h * c + c] = gy.sum(dim=0)
                gx0 = gx0 + torch.bmm(gz, xl.unsqueeze(2)).squeeze(2)
                g_next = torch.bmm(x0.unsqueeze(1), gz).squeeze(1)
            gx0 = gx0 + g_next                                            # CINEncoder.scala:85
            ge = gx0.reshape(B, K, F).transpose(1, 2).reshape(B, F * K) + gx_dnn
            ge = ge.reshape(-1)
        weights[...] = gw.numpy()
        bias[...] = dbias.numpy()
        if ge is not None:
            embedding[...] = ge.numpy()
        if gm is not None:
            t0 = time.perf_counter()
            mats[...] = gm.numpy()                                        # makeMatsGrad
            self.t_copy += time.perf_counter() - t0
        return float(loss)


def make_dense(kind, n_fields, k, fc, cin, depth):
    """(object with .backward(B, index, w, bias, emb, mats, targets), description of its BLAS)."""
    if kind in TorchModel.KINDS:
        return TorchModel(kind, n_fields, k, fc, cin), "torch-CPU/MKL sgemm"
    from . import refport
    return refport.Model(kind, n_fields, k, fc, cin, depth), "numpy/OpenBLAS sgemm"


def run_steps(kind, n_fields, k, fc, cin, depth, batch, rows, synth, seed_data, seed_params, threads,
              budget_s, max_steps=None, warmup=1, first_step=0):
    """Time optimize-without-RPC steps of the reference path on `threads` host threads.
    -> dict(value samples/s, ms_per_step, steps, phases_ms{gather, mats_copy, dense, scatter_add}, blas)."""
    torch = _t()
    prev_threads = torch.get_num_threads()
    torch.set_num_threads(int(threads))
    try:
        dense, blas = make_dense(kind, n_fields, k, fc, cin, depth)
        from . import refport
        mats = synth.init_mats(seed_params, refport.mats_size(kind, n_fields, k, fc, cin, depth))
        ph = dict(gather=0.0, mats_copy=0.0, dense=0.0, scatter_add=0.0)
        total, done, step = 0.0, 0, 0
        t_start = time.perf_counter()
        while True:
            index, feats = synth.make_feats(seed_data, first_step + step, batch, n_fields, rows)
            targets = synth.make_targets(seed_data, feats, batch, n_fields)
            ids = np.unique(feats).astype(np.int32)
            # the PS pull (rows of the distinct ids -> K + 1 sparse vectors) is RPC: untimed
            E = synth.table_rows(seed_params, ids, k) if kind != "lr" else None
            wv = synth.wtable_rows(seed_params, ids)
            pulled = Pulled(ids, E, wv, threads)
            t0 = time.perf_counter()
            emb = make_embeddings(pulled, feats, threads) if kind != "lr" else None
            ww = make_weights(pulled, feats)
            t1 = time.perf_counter()
            bias = np.array([0.1], np.float32)
            m = mats.copy() if mats.size else None                     # makeMats
            t2 = time.perf_counter()
            dense.backward(batch, index, ww, bias, emb, m, targets)
            t3 = time.perf_counter()
            _, _, _, addto_s = scatter_add(feats, emb, ww, k, ids=ids, threads=threads, want_out=False)
            pulled.close()
            step += 1
            if step > warmup:
                copy_s = (t2 - t1) + getattr(dense, "t_copy", 0.0)
                ph["gather"] += t1 - t0
                ph["mats_copy"] += copy_s
                ph["dense"] += (t3 - t2) - getattr(dense, "t_copy", 0.0)
                ph["scatter_add"] += addto_s
                total += (t1 - t0) + (t3 - t1) + addto_s
                done += 1
            if max_steps is not None and done >= max_steps:
                break
            if time.perf_counter() - t_start > budget_s and done >= 2:
                break
        return dict(value=batch * done / total, ms_per_step=1e3 * total / done, steps=done, threads=int(threads),
                    phases_ms={p: round(1e3 * v / done, 3) for p, v in ph.items()}, blas=blas)
    finally:
        torch.set_num_threads(prev_threads)
