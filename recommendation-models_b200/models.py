"""Internal<M>Model of the reference with the same constructor and forward/backward signatures,
running on the B200 through the C ABI.

Mirrors (under /root/reference/src/main/scala/io/yaochi/recommendation/model):
  lr/LR.scala:42-90            InternalLRModel
  deepfm/DeepFM.scala:51-125   InternalDeepFMModel
  xdeepfm/XDeepFM.scala:58-126 InternalXDeepFMModel
  dcn/DCN.scala:62-130         InternalDCNModel
  pnn/PNN.scala:56-132         InternalPNNModel
  (FM: no class in the reference, SURVEY B-1)  InternalFMModel = DeepFM minus HigherOrderEncoder
`backward` overwrites weights / bias / embedding / mats with their gradients and returns the loss,
like rec/util/GradUtil.scala:7-42 and rec/util/BackwardUtil.scala:6-42.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


class _InternalModel:
    kind = None

    def __init__(self, nFields=0, embeddingDim=0, fcDims=(), cinDims=(), crossDepth=0, device=0):
        self.nFields, self.embeddingDim = int(nFields), int(embeddingDim)
        self.fcDims, self.cinDims, self.crossDepth = list(fcDims), list(cinDims), int(crossDepth)
        self.device = device
        fc = (C.c_int * max(1, len(self.fcDims)))(*self.fcDims)
        cin = (C.c_int * max(1, len(self.cinDims)))(*self.cinDims)
        h = C.c_void_p()
        L.check(L.lib().b200rec_model_create(L.KINDS[self.kind], self.nFields, self.embeddingDim, fc,
                                             len(self.fcDims), cin, len(self.cinDims), self.crossDepth,
                                             device, C.byref(h)))
        self.handle = h

    def close(self):
        if getattr(self, "handle", None):
            L.lib().b200rec_model_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def setGemmMode(self, mode):
        """0 fp32 FFMA, 1 3xTF32 tcgen05, 2 1xTF32 tcgen05 (not parity grade)."""
        L.check(L.lib().b200rec_model_set_gemm_mode(self.handle, int(mode)))

    def getMatsSize(self):
        n = C.c_int(0)
        L.check(L.lib().b200rec_model_mats_size(self.handle, None, 0, C.byref(n)))
        buf = (C.c_int * max(1, n.value))()
        L.check(L.lib().b200rec_model_mats_size(self.handle, buf, n.value, C.byref(n)))
        return [int(buf[i]) for i in range(n.value)]

    def matsLen(self):
        n = C.c_int64(0)
        L.check(L.lib().b200rec_model_mats_len(self.handle, C.byref(n)))
        return n.value

    @staticmethod
    def _inplace(a, name):
        if a is None:
            return None
        if not (isinstance(a, np.ndarray) and a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
                and a.flags["WRITEABLE"]):
            raise ValueError(f"{name} must be a writable C-contiguous float32 array "
                             "(backward overwrites it with its gradient)")
        return a

    def forward(self, batchSize, index, weights, bias, embedding=None, mats=None):
        index, weights, bias = L.i32(index), L.f32(weights), L.f32(bias)
        embedding, mats = L.f32(embedding), L.f32(mats)
        preds = np.zeros(batchSize, np.float32)
        L.check(L.lib().b200rec_forward(self.handle, batchSize, index.shape[0], L.ptr(index),
                                        L.ptr(weights), L.ptr(bias), L.ptr(embedding), L.ptr(mats),
                                        L.ptr(preds)))
        return preds

    def backward(self, batchSize, index, weights, bias, embedding=None, mats=None, targets=None):
        index = L.i32(index)
        weights, bias = self._inplace(weights, "weights"), self._inplace(bias, "bias")
        embedding, mats = self._inplace(embedding, "embedding"), self._inplace(mats, "mats")
        targets = L.f32(targets)
        loss = C.c_float(0)
        L.check(L.lib().b200rec_backward(self.handle, batchSize, index.shape[0], L.ptr(index),
                                         L.ptr(weights), L.ptr(bias), L.ptr(embedding), L.ptr(mats),
                                         L.ptr(targets), C.byref(loss)))
        return float(loss.value)


class InternalLRModel(_InternalModel):
    kind = "lr"

    def __init__(self, device=0, nFields=0):
        super().__init__(nFields=nFields, device=device)


class InternalFMModel(_InternalModel):
    kind = "fm"

    def __init__(self, nFields, embeddingDim, device=0):
        super().__init__(nFields, embeddingDim, device=device)


class InternalDeepFMModel(_InternalModel):
    kind = "deepfm"

    def __init__(self, nFields, embeddingDim, fcDims, device=0):
        super().__init__(nFields, embeddingDim, fcDims, device=device)


class InternalXDeepFMModel(_InternalModel):
    kind = "xdeepfm"

    def __init__(self, nFields, embeddingDim, fcDims, cinDims, device=0):
        super().__init__(nFields, embeddingDim, fcDims, cinDims, device=device)


class InternalDCNModel(_InternalModel):
    kind = "dcn"

    def __init__(self, nFields, embeddingDim, crossDepth, fcDims, device=0):
        super().__init__(nFields, embeddingDim, fcDims, crossDepth=crossDepth, device=device)


class InternalPNNModel(_InternalModel):
    kind = "pnn"

    def __init__(self, nFields, embeddingDim, fcDims, device=0):
        super().__init__(nFields, embeddingDim, fcDims, device=device)


def make_model(kind, nFields=0, embeddingDim=0, fcDims=(), cinDims=(), crossDepth=0, device=0):
    kind = kind.lower()
    if kind == "lr":
        return InternalLRModel(device, nFields)
    if kind == "fm":
        return InternalFMModel(nFields, embeddingDim, device)
    if kind == "deepfm":
        return InternalDeepFMModel(nFields, embeddingDim, fcDims, device)
    if kind == "xdeepfm":
        return InternalXDeepFMModel(nFields, embeddingDim, fcDims, cinDims, device)
    if kind == "dcn":
        return InternalDCNModel(nFields, embeddingDim, crossDepth, fcDims, device)
    if kind == "pnn":
        return InternalPNNModel(nFields, embeddingDim, fcDims, device)
    raise ValueError(f"unknown model kind {kind!r}")
