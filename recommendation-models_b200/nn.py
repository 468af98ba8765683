"""The reference's own BigDL modules and encoders, with their method names
(updateOutput / updateGradInput / accGradParameters), computed by libb200rec on the B200.

Mirrors (paths under /root/reference/src/main/scala):
  com/intel/analytics/bigdl/nn/Scatter.scala, Gather.scala, DotProduct2.scala
  io/yaochi/recommendation/model/encoder/{FirstOrder,SecondOrder}Encoder.scala
  BigDL Linear as built by io/yaochi/recommendation/util/LayerUtil.scala:7-24
Tensors are numpy float32 arrays (the Scala shim passes Array[Float] storage the same way).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


class _Module:
    def __init__(self, device=0):
        self.device = device
        self.output = None
        self.gradInput = None

    def forward(self, input):
        return self.updateOutput(input)

    def backward(self, input, gradOutput):
        g = self.updateGradInput(input, gradOutput)
        self.accGradParameters(input, gradOutput)
        return g

    def accGradParameters(self, input, gradOutput):
        pass

    def clearState(self):
        self.output = None
        self.gradInput = None
        return self


class Scatter(_Module):
    """nn/Scatter.scala: Scatter(batchSize, nOutput).  input = (tensor[n(,nOutput)], index[n])."""

    def __init__(self, batchSize, nOutput, device=0):
        super().__init__(device)
        self.batchSize, self.nOutput = batchSize, nOutput

    def updateOutput(self, input):
        x, index = L.f32(input[0]), L.i32(input[1])
        n = index.shape[0]
        out = np.zeros((self.batchSize, self.nOutput), np.float32)
        L.check(L.lib().b200rec_scatter_update_output(self.device, self.batchSize, self.nOutput, n,
                                                      L.ptr(x), L.ptr(index), L.ptr(out)))
        self.output = out
        return out

    def updateGradInput(self, input, gradOutput):
        index = L.i32(input[1])
        n = index.shape[0]
        go = L.f32(gradOutput).reshape(self.batchSize, self.nOutput)
        gi = np.zeros((n, self.nOutput), np.float32)
        L.check(L.lib().b200rec_scatter_update_grad_input(self.device, self.batchSize, self.nOutput, n,
                                                          L.ptr(index), L.ptr(go), L.ptr(gi)))
        self.gradInput = (gi, None)
        return self.gradInput


class Gather(_Module):
    """nn/Gather.scala: Gather(batchSize, numPairs, embeddingSize).  input = (x[B,F,K], rows, cols)."""

    def __init__(self, batchSize, numPairs, embeddingSize, device=0):
        super().__init__(device)
        self.batchSize, self.numPairs, self.embeddingSize = batchSize, numPairs, embeddingSize

    def updateOutput(self, input):
        x, rows, cols = L.f32(input[0]), L.i32(input[1]), L.i32(input[2])
        B, F, K = x.shape
        ro = np.zeros((B, self.numPairs, K), np.float32)
        co = np.zeros_like(ro)
        L.check(L.lib().b200rec_gather_update_output(self.device, B, F, self.numPairs, K, L.ptr(x),
                                                     L.ptr(rows), L.ptr(cols), L.ptr(ro), L.ptr(co)))
        self.output = (ro, co)
        return self.output

    def updateGradInput(self, input, gradOutput):
        x, rows, cols = input[0], L.i32(input[1]), L.i32(input[2])
        B, F, K = x.shape
        gr, gc = L.f32(gradOutput[0]), L.f32(gradOutput[1])
        gi = np.zeros((B, F, K), np.float32)
        L.check(L.lib().b200rec_gather_update_grad_input(self.device, B, F, self.numPairs, K,
                                                         L.ptr(rows), L.ptr(cols), L.ptr(gr),
                                                         L.ptr(gc), L.ptr(gi)))
        self.gradInput = (gi, None, None)
        return self.gradInput


class DotProduct2(_Module):
    """nn/DotProduct2.scala: input = (a[B,P,K], b[B,P,K]) -> [B,P]."""

    def updateOutput(self, input):
        a, b = L.f32(input[0]), L.f32(input[1])
        Bn, P, K = a.shape
        out = np.zeros((Bn, P), np.float32)
        L.check(L.lib().b200rec_dotproduct2_update_output(self.device, Bn * P, K, L.ptr(a), L.ptr(b),
                                                          L.ptr(out)))
        self.output = out
        return out

    def updateGradInput(self, input, gradOutput):
        a, b, go = L.f32(input[0]), L.f32(input[1]), L.f32(gradOutput)
        Bn, P, K = a.shape
        ga, gb = np.zeros_like(a), np.zeros_like(b)
        L.check(L.lib().b200rec_dotproduct2_update_grad_input(self.device, Bn * P, K, L.ptr(a),
                                                              L.ptr(b), L.ptr(go), L.ptr(ga), L.ptr(gb)))
        self.gradInput = (ga, gb)
        return self.gradInput


class FirstOrderEncoder:
    """rec/model/encoder/FirstOrderEncoder.scala:7-17."""

    def __init__(self, batchSize, device=0):
        self.batchSize = batchSize
        self.module = Scatter(batchSize, 1, device)

    def forward(self, weights, index):
        return self.module.forward((weights, index))

    def backward(self, weights, index, gradOutput):
        return self.module.backward((weights, index), gradOutput)[0].reshape(-1)


class SecondOrderEncoder:
    """rec/model/encoder/SecondOrderEncoder.scala:8-35."""

    def __init__(self, batchSize, nFields, embeddingDim, device=0):
        self.batchSize, self.nFields, self.embeddingDim, self.device = batchSize, nFields, embeddingDim, device

    def forward(self, input):
        x = L.f32(input).reshape(-1)
        if x.size != self.batchSize * self.nFields * self.embeddingDim:
            raise ValueError("Reshape: element count mismatch")
        out = np.zeros((self.batchSize, 1), np.float32)
        L.check(L.lib().b200rec_second_order_update_output(self.device, self.batchSize, self.nFields,
                                                           self.embeddingDim, L.ptr(x), L.ptr(out)))
        return out

    def backward(self, input, gradOutput):
        x = L.f32(input).reshape(-1)
        go = L.f32(gradOutput).reshape(-1)
        gi = np.zeros_like(x)
        L.check(L.lib().b200rec_second_order_update_grad_input(self.device, self.batchSize, self.nFields,
                                                               self.embeddingDim, L.ptr(x), L.ptr(go),
                                                               L.ptr(gi)))
        return gi


class Linear(_Module):
    """BigDL Linear(inputSize, outputSize, withBias, initWeight, initBias) as LayerUtil builds it:
    weight [out,in] row-major; accGradParameters accumulates into gradWeight / gradBias."""

    def __init__(self, inputSize, outputSize, withBias=True, initWeight=None, initBias=None, device=0):
        super().__init__(device)
        self.inputSize, self.outputSize = inputSize, outputSize
        self.weight = L.f32(initWeight, copy=True).reshape(outputSize, inputSize) if initWeight is not None \
            else np.zeros((outputSize, inputSize), np.float32)
        self.bias = (L.f32(initBias, copy=True).reshape(outputSize) if initBias is not None
                     else np.zeros(outputSize, np.float32)) if withBias else None
        self.gradWeight = np.zeros_like(self.weight)
        self.gradBias = np.zeros_like(self.bias) if withBias else None
        self.scaleW = 1.0

    def updateOutput(self, input):
        x = L.f32(input).reshape(-1, self.inputSize)
        y = np.zeros((x.shape[0], self.outputSize), np.float32)
        L.check(L.lib().b200rec_linear_update_output(self.device, x.shape[0], self.inputSize,
                                                     self.outputSize, L.ptr(x), L.ptr(self.weight),
                                                     L.ptr(self.bias), 0, L.ptr(y)))
        self.output = y
        return y

    def updateGradInput(self, input, gradOutput):
        gy = L.f32(gradOutput).reshape(-1, self.outputSize)
        gx = np.zeros((gy.shape[0], self.inputSize), np.float32)
        L.check(L.lib().b200rec_linear_update_grad_input(self.device, gy.shape[0], self.inputSize,
                                                         self.outputSize, L.ptr(gy), L.ptr(self.weight),
                                                         L.ptr(gx)))
        self.gradInput = gx
        return gx

    def accGradParameters(self, input, gradOutput):
        x = L.f32(input).reshape(-1, self.inputSize)
        gy = L.f32(gradOutput).reshape(-1, self.outputSize)
        L.check(L.lib().b200rec_linear_acc_grad_parameters(self.device, x.shape[0], self.inputSize,
                                                           self.outputSize, L.ptr(x), L.ptr(gy),
                                                           self.scaleW, L.ptr(self.gradWeight),
                                                           L.ptr(self.gradBias)))

    def zeroGradParameters(self):
        self.gradWeight[...] = 0
        if self.gradBias is not None:
            self.gradBias[...] = 0


# ---- the encoders (dense branches), rec/model/{encoder,xdeepfm,dcn,pnn}/*Encoder.scala -------------------
class _Encoder:
    """forward(input) -> output, backward(input, gradOutput) -> gradInput with the parameter gradients
    copied over `self.mats[start:...]` at the parameters' offsets -- the reference encoders' contract
    (e.g. HigherOrderEncoder.scala:18-32).  The BigDL triple is available too."""
    _abi = None
    _kind = None

    def __init__(self, batchSize, mats, start, handle_args, device=0):
        from .models import make_model
        self.batchSize, self.mats, self.start = int(batchSize), mats, int(start)
        self._model = make_model(self._kind, *handle_args, device=device)
        n = C.c_int64(0)
        L.check(L.lib().b200rec_encoder_mats_len(self._model.handle, C.byref(n)))
        self.matsLen = n.value
        if self.start < 0 or self.start + self.matsLen > len(mats):
            raise ValueError(f"mats has {len(mats)} elements, the encoder needs [{self.start}, {self.start + self.matsLen})")
        self.output = self.gradInput = None

    def _params(self):
        return np.ascontiguousarray(self.mats[self.start:self.start + self.matsLen], np.float32)

    def _fn(self, suffix):
        return getattr(L.lib(), f"b200rec_{self._abi}_{suffix}")

    def _in(self, input):
        x = np.ascontiguousarray(np.asarray(input, np.float32).reshape(-1))
        if x.size != self.batchSize * self.inputDim:
            raise ValueError(f"input has {x.size} elements, expected {self.batchSize}*{self.inputDim}")
        return x

    def updateOutput(self, input):
        x, p = self._in(input), self._params()
        out = np.zeros((self.batchSize, self.outputDim), np.float32)
        L.check(self._fn("update_output")(self._model.handle, self.batchSize, L.ptr(x), L.ptr(p), L.ptr(out)))
        self.output = out
        return out

    forward = updateOutput

    def updateGradInput(self, input, gradOutput):
        x, p, go = self._in(input), self._params(), L.f32(np.asarray(gradOutput).reshape(-1))
        gi = np.zeros(x.size, np.float32)
        L.check(self._fn("update_grad_input")(self._model.handle, self.batchSize, L.ptr(x), L.ptr(p), L.ptr(go), L.ptr(gi)))
        self.gradInput = gi
        return gi

    def accGradParameters(self, input, gradOutput, gradMats, scale=1.0):
        """gradMats[start:...] += scale * parameter gradients (BigDL accGradParameters accumulates)."""
        x, p, go = self._in(input), self._params(), L.f32(np.asarray(gradOutput).reshape(-1))
        g = np.ascontiguousarray(gradMats[self.start:self.start + self.matsLen], np.float32)
        L.check(self._fn("acc_grad_parameters")(self._model.handle, self.batchSize, L.ptr(x), L.ptr(p), L.ptr(go),
                                                float(scale), L.ptr(g)))
        gradMats[self.start:self.start + self.matsLen] = g

    def backward(self, input, gradOutput):
        x, p, go = self._in(input), self._params(), L.f32(np.asarray(gradOutput).reshape(-1))
        gi = np.zeros(x.size, np.float32)
        L.check(self._fn("backward")(self._model.handle, self.batchSize, L.ptr(x), L.ptr(p), L.ptr(go), L.ptr(gi)))
        self.mats[self.start:self.start + self.matsLen] = p      # BackwardUtil.linearBackward: grads over mats
        self.gradInput = gi
        return gi

    def close(self):
        self._model.close()


class HigherOrderEncoder(_Encoder):
    """rec/model/encoder/HigherOrderEncoder.scala: HigherOrderEncoder(batchSize, inputDim, fcDims, mats, start)."""
    _abi, _kind = "higher_order", "deepfm"

    def __init__(self, batchSize, inputDim, fcDims, mats, start=0, reshape=True, device=0):
        self.inputDim, self.outputDim = int(inputDim), 1
        super().__init__(batchSize, mats, start, (int(inputDim), 1, list(fcDims)), device)


class CINEncoder(_Encoder):
    """rec/model/xdeepfm/CINEncoder.scala: CINEncoder(batchSize, nFields, embeddingDim, fcDims, cinDims, mats, start)."""
    _abi, _kind = "cin", "xdeepfm"

    def __init__(self, batchSize, nFields, embeddingDim, fcDims, cinDims, mats, start=0, device=0):
        self.inputDim, self.outputDim = nFields * embeddingDim, 1
        super().__init__(batchSize, mats, start, (nFields, embeddingDim, list(fcDims), list(cinDims)), device)


class CrossEncoder(_Encoder):
    """rec/model/dcn/CrossEncoder.scala: CrossEncoder(batchSize, nFields, embeddingDim, crossDepth, fcDims, mats, start)."""
    _abi, _kind = "cross", "dcn"

    def __init__(self, batchSize, nFields, embeddingDim, crossDepth, fcDims, mats, start=0, device=0):
        self.inputDim, self.outputDim = nFields * embeddingDim, 1
        super().__init__(batchSize, mats, start, (nFields, embeddingDim, list(fcDims), (), int(crossDepth)), device)


class ProductEncoder(_Encoder):
    """rec/model/pnn/ProductEncoder.scala: ProductEncoder(batchSize, nFields, embeddingDim, outputDim, mats, start)."""
    _abi, _kind = "product", "pnn"

    def __init__(self, batchSize, nFields, embeddingDim, outputDim, mats, start=0, device=0):
        self.inputDim, self.outputDim = nFields * embeddingDim, int(outputDim)
        super().__init__(batchSize, mats, start, (nFields, embeddingDim, [int(outputDim)]), device)


class DuplicateTable(_Module):
    """nn/DuplicateTable.scala: fan the input out to the member modules, sum their input gradients."""

    def __init__(self, device=0):
        super().__init__(device)
        self.modules = []

    def add(self, module):
        self.modules.append(module)
        return self

    def updateOutput(self, input):
        self.output = [m.forward(input) for m in self.modules]
        return self.output

    def updateGradInput(self, input, gradOutput):
        gs = [np.asarray(m.updateGradInput(input, g), np.float32).reshape(-1) for m, g in zip(self.modules, gradOutput)]
        stack = np.ascontiguousarray(np.stack(gs)) if gs else np.zeros((0, 0), np.float32)
        out = np.zeros(stack.shape[1] if gs else 0, np.float32)
        L.check(L.lib().b200rec_duplicate_table_update_grad_input(self.device, len(gs), out.size, L.ptr(stack), L.ptr(out)))
        self.gradInput = out.reshape(np.asarray(input).shape)
        return self.gradInput
