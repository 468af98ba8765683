"""The reference's own BigDL modules and encoders, with their method names
(updateOutput / updateGradInput / accGradParameters), computed by libb200rec on the B200.

Mirrors (paths under /root/reference/src/main/scala):
  com/intel/analytics/bigdl/nn/Scatter.scala, Gather.scala, DotProduct2.scala
  io/yaochi/recommendation/model/encoder/{FirstOrder,SecondOrder}Encoder.scala
  BigDL Linear as built by io/yaochi/recommendation/util/LayerUtil.scala:7-24
Tensors are numpy float32 arrays (the Scala shim passes Array[Float] storage the same way).
"""
from __future__ import annotations

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
