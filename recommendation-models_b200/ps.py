"""Device-resident stand-in for the PS half of ParRecModel: the table that replaces the Angel
PSMatrix / PSVector objects, and the optimize / predict flows over it.

Mirrors rec/model/ParRecModel.scala (under /root/reference/src/main/scala/io/yaochi/recommendation):
  :74-105  initMats                      -> EmbeddingTable(rows, dim)
  :279-306 makeWeights / makeEmbeddings  -> EmbeddingTable.lookup
  :293-328 make*Grad                     -> scatter_add / the fused step
  :337-345 distinctIntIndices            -> distinct
  :348-363, :439-478 optimize            -> ParRecModel.optimize
  :519-533, :555-567 predict             -> ParRecModel.predict
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


class EmbeddingTable:
    """`embedding` matrix [rows, dim] + first-order `weights` vector [rows], resident in HBM."""

    def __init__(self, rows, dim, device=0):
        self.rows, self.dim, self.device = int(rows), int(dim), device
        h = C.c_void_p()
        L.check(L.lib().b200rec_table_create(self.rows, self.dim, device, C.byref(h)))
        self.handle = h

    def close(self):
        if getattr(self, "handle", None):
            L.lib().b200rec_table_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init_uniform(self, seed, lo=-0.05, hi=0.05, row_offset=0, row_stride=1):
        L.check(L.lib().b200rec_table_init_uniform(self.handle, seed, lo, hi, row_offset, row_stride))

    def write(self, row0, embedding=None, weights=None):
        embedding, weights = L.f32(embedding), L.f32(weights)
        n = embedding.shape[0] if embedding is not None else weights.shape[0]
        L.check(L.lib().b200rec_table_write(self.handle, row0, n, L.ptr(embedding), L.ptr(weights)))

    def read(self, row0, nrows):
        e = np.zeros((nrows, self.dim), np.float32)
        w = np.zeros(nrows, np.float32)
        L.check(L.lib().b200rec_table_read(self.handle, row0, nrows, L.ptr(e), L.ptr(w)))
        return e, w

    def ptrs(self):
        e, w = C.c_void_p(), C.c_void_p()
        L.check(L.lib().b200rec_table_ptrs(self.handle, C.byref(e), C.byref(w)))
        return e.value, w.value

    def lookup(self, feats):
        """makeEmbeddings + makeWeights: (float32[N*K], float32[N]) -- bit-exact copies."""
        feats = L.i32(feats)
        n = feats.shape[0]
        e = np.zeros(n * self.dim, np.float32)
        w = np.zeros(n, np.float32)
        L.check(L.lib().b200rec_table_lookup(self.handle, n, L.ptr(feats), L.ptr(e), L.ptr(w)))
        return e, w


def distinct(feats, device=0):
    """distinctIntIndices -> ascending int32[U]."""
    feats = L.i32(feats)
    out = np.zeros(max(1, feats.shape[0]), np.int32)
    n = C.c_int64(0)
    L.check(L.lib().b200rec_distinct(device, feats.shape[0], L.ptr(feats), L.ptr(out), C.byref(n)))
    return out[:n.value].copy()


def scatter_add(feats, embedding_grad=None, weights_grad=None, dim=0, device=0):
    """makeEmbeddingGrad + makeWeightsGrad -> (ids[U], G[U,dim] | None, gw[U] | None)."""
    feats = L.i32(feats)
    n = feats.shape[0]
    eg, wg = L.f32(embedding_grad), L.f32(weights_grad)
    ids = np.zeros(max(1, n), np.int32)
    G = np.zeros((max(1, n), dim), np.float32) if eg is not None else None
    gw = np.zeros(max(1, n), np.float32) if wg is not None else None
    u = C.c_int64(0)
    L.check(L.lib().b200rec_scatter_add(device, dim, n, L.ptr(feats), L.ptr(eg), L.ptr(wg), L.ptr(ids),
                                        L.ptr(G), L.ptr(gw), C.byref(u)))
    U = u.value
    return ids[:U].copy(), (G[:U].copy() if G is not None else None), (gw[:U].copy() if gw is not None else None)


class ParRecModel:
    """optimize / predict over a resident table (single GPU).  `model` is an Internal<M>Model."""

    def __init__(self, model, table):
        self.model, self.table = model, table

    def setParams(self, bias, mats=None):
        bias, mats = L.f32(bias), L.f32(mats)
        L.check(L.lib().b200rec_model_set_params(self.model.handle, L.ptr(bias), L.ptr(mats)))

    def getParams(self):
        bias = np.zeros(1, np.float32)
        mats = np.zeros(max(1, self.model.matsLen()), np.float32)
        L.check(L.lib().b200rec_model_get_params(self.model.handle, L.ptr(bias), L.ptr(mats)))
        return bias, mats[:self.model.matsLen()]

    def optimize(self, feats, targets):
        """ParRecModel.optimize :348-363 -> loss * batchSize (:477).  feats int32[B*F], one id per
        field per sample, sample-major."""
        feats, targets = L.i32(feats), L.f32(targets)
        B = targets.shape[0]
        self._last_nnz = feats.shape[0]
        loss = C.c_float(0)
        L.check(L.lib().b200rec_step(self.model.handle, self.table.handle, B, L.ptr(feats),
                                     L.ptr(targets), C.byref(loss)))
        return float(loss.value) * B

    def stage(self, feats, targets):
        """Input prefetch: start copying a batch to the GPU (pinned int32 / float32 arrays overlap with
        the running step; they must stay alive until the matching optimizeStaged returns)."""
        feats, targets = L.i32(feats), L.f32(targets)
        L.check(L.lib().b200rec_stage_batch(self.model.handle, targets.shape[0], L.ptr(feats), L.ptr(targets)))
        self._staged = getattr(self, "_staged", [])
        self._staged.append((feats, targets))

    def optimizeStaged(self):
        """optimize() on the batch staged first -> loss * batchSize."""
        loss = C.c_float(0)
        L.check(L.lib().b200rec_step_staged(self.model.handle, self.table.handle, C.byref(loss)))
        feats, targets = self._staged.pop(0)
        self._last_nnz = feats.shape[0]
        return float(loss.value) * targets.shape[0]

    def optimizeStagedAsync(self):
        """Enqueue optimize() on the batch staged first and return at once (at most two in flight);
        waitLoss() returns the losses in order.  `stage(i+1); optimizeStagedAsync(); waitLoss()` of the
        previous step keeps the GPU busy while every loss is still read by the host, one step late."""
        L.check(L.lib().b200rec_step_staged_async(self.model.handle, self.table.handle))
        feats, targets = self._staged.pop(0)
        self._last_nnz = feats.shape[0]
        self._inflight = getattr(self, "_inflight", [])
        self._inflight.append(targets.shape[0])

    def waitLoss(self):
        """-> loss * batchSize of the oldest step enqueued by optimizeStagedAsync."""
        loss = C.c_float(0)
        L.check(L.lib().b200rec_step_wait(self.model.handle, self.table.handle, C.byref(loss)))
        return float(loss.value) * self._inflight.pop(0)

    def predict(self, feats, batchSize):
        feats = L.i32(feats)
        preds = np.zeros(batchSize, np.float32)
        L.check(L.lib().b200rec_predict(self.model.handle, self.table.handle, batchSize, L.ptr(feats),
                                        L.ptr(preds)))
        return preds

    # ---- checkpoint (the reference has no format of its own: ParRecModel.scala:66-123
    # only initialises) ------------------------------------------------------------------------------
    def save(self, path, chunk_rows=1 << 20):
        """bias, mats and the table (embedding + first-order weights) to one .npz, read from the GPU in
        row chunks.  Optimizer slots are not saved."""
        bias, mats = self.getParams()
        t = self.table
        emb = np.empty((t.rows, t.dim), np.float32)
        w = np.empty(t.rows, np.float32)
        for r0 in range(0, t.rows, chunk_rows):
            n = min(chunk_rows, t.rows - r0)
            emb[r0:r0 + n], w[r0:r0 + n] = t.read(r0, n)
        np.savez(path, bias=bias, mats=mats, embedding=emb, weights=w,
                 meta=np.array([t.rows, t.dim, self.model.matsLen()], np.int64))

    def load(self, path, chunk_rows=1 << 20):
        z = np.load(path)
        rows, dim, mats_len = (int(v) for v in z["meta"])
        t = self.table
        if (rows, dim, mats_len) != (t.rows, t.dim, self.model.matsLen()):
            raise ValueError(f"checkpoint is for rows={rows} dim={dim} mats={mats_len}, "
                             f"this model has rows={t.rows} dim={t.dim} mats={self.model.matsLen()}")
        self.setParams(z["bias"], z["mats"] if mats_len else None)
        emb, w = z["embedding"], z["weights"]
        for r0 in range(0, rows, chunk_rows):
            n = min(chunk_rows, rows - r0)
            t.write(r0, emb[r0:r0 + n] if dim else None, w[r0:r0 + n])

    # ---- optimizer step: rec/optim/OptimUtils.scala:5-12 + Async*.scala defaults -------------------
    _DEFAULTS = {"sgd": (0.0, 0.0), "momentum": (0.9, 0.0), "adagrad": (0.9, 0.0), "adam": (0.99, 0.9)}

    def applyOptimizer(self, optim, stepSize, p1=None, p2=None):
        """What the PS does with the pushed gradients (ParRecModel.push* :201-267 -> optim.asycUpdate):
        update the touched table rows and the dense params in place on the GPU.  `optim` is the
        reference's name ("sgd" | "momentum" | "adagrad" | "adam"; anything else raises, like the
        MatchError of OptimUtils.scala:6-11).  Arithmetic: textbook forms, parity unpinned."""
        name = optim.lower()
        if name not in L.OPTIMIZERS:
            raise ValueError(f"unknown optimizer {optim!r}")
        d1, d2 = self._DEFAULTS[name]
        p1 = d1 if p1 is None else p1
        p2 = d2 if p2 is None else p2
        self._opt_step = getattr(self, "_opt_step", 0) + 1
        lib, m, t = L.lib(), self.model, self.table
        ptrs = [C.c_void_p() for _ in range(7)]
        L.check(lib.b200rec_step_result_ptrs(m.handle, *[C.byref(p) for p in ptrs]))
        _, n_unique, unique, emb_grad, w_grad, _, _ = [p.value for p in ptrs]
        sp = C.c_void_p()
        L.check(lib.b200rec_model_stream(m.handle, C.byref(sp)))
        cap = self._last_nnz
        L.check(lib.b200rec_table_apply_optimizer_dev(t.handle, L.OPTIMIZERS[name], stepSize, p1, p2,
                                                      self._opt_step, cap, n_unique, unique, emb_grad,
                                                      w_grad, sp))
        L.check(lib.b200rec_model_apply_optimizer_dev(m.handle, L.OPTIMIZERS[name], stepSize, p1, p2,
                                                      self._opt_step, sp))
        L.check(lib.b200rec_model_sync(m.handle))

    def stepResults(self):
        """Gradients of the last optimize: dict(loss, unique, emb_grad, w_grad, bias_grad, mats_grad)."""
        m = self.model
        nnz_cap = None
        loss, nu, bg = C.c_float(0), C.c_int64(0), C.c_float(0)
        L.check(L.lib().b200rec_step_results(m.handle, C.byref(loss), C.byref(nu), None, None, None,
                                             C.byref(bg), None))
        U = nu.value
        uniq = np.zeros(max(1, U), np.int32)
        K = m.embeddingDim
        G = np.zeros((max(1, U), max(1, K)), np.float32)
        gw = np.zeros(max(1, U), np.float32)
        gm = np.zeros(max(1, m.matsLen()), np.float32)
        L.check(L.lib().b200rec_step_results(m.handle, None, None, L.ptr(uniq),
                                             L.ptr(G) if m.kind != "lr" else None, L.ptr(gw), None,
                                             L.ptr(gm) if m.matsLen() else None))
        return dict(loss=float(loss.value), unique=uniq[:U], emb_grad=G[:U, :K] if m.kind != "lr" else None,
                    w_grad=gw[:U], bias_grad=float(bg.value), mats_grad=gm[:m.matsLen()])
