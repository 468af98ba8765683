// Drop-in bodies for the reference's Internal<M>Model classes (NOT compiled in this repo: no JVM in
// the build image).  Same constructor arguments and method signatures as
// rec/model/deepfm/DeepFM.scala:51-125 etc.; the BigDL graph is replaced by one C-ABI call.
package io.yaochi.recommendation.b200

import com.sun.jna.Pointer
import com.sun.jna.ptr.{FloatByReference, PointerByReference}
import B200Rec.{check, lib}

object Kind { val LR = 0; val FM = 1; val DeepFM = 2; val XDeepFM = 3; val DCN = 4; val PNN = 5 }

abstract class B200InternalModel(kind: Int, nFields: Int, embeddingDim: Int, fcDims: Array[Int],
                                 cinDims: Array[Int], crossDepth: Int, device: Int = 0) {
  private val handle: Pointer = {
    val out = new PointerByReference()
    check(lib.b200rec_model_create(kind, nFields, embeddingDim, fcDims, fcDims.length, cinDims,
      cinDims.length, crossDepth, device, out))
    out.getValue
  }

  def forward(batchSize: Int, index: Array[Int], weights: Array[Float], bias: Array[Float],
              embedding: Array[Float], mats: Array[Float]): Array[Float] = {
    val preds = new Array[Float](batchSize)
    check(lib.b200rec_forward(handle, batchSize, index.length, index, weights, bias, embedding, mats, preds))
    preds
  }

  /** Returns the mean BCE loss; weights / bias / embedding / mats come back holding gradients. */
  def backward(batchSize: Int, index: Array[Int], weights: Array[Float], bias: Array[Float],
               embedding: Array[Float], mats: Array[Float], targets: Array[Float]): Float = {
    val loss = new FloatByReference()
    check(lib.b200rec_backward(handle, batchSize, index.length, index, weights, bias, embedding, mats,
      targets, loss))
    loss.getValue
  }

  def close(): Unit = check(lib.b200rec_model_destroy(handle))
}

class InternalDeepFMModel(nFields: Int, embeddingDim: Int, fcDims: Array[Int])
  extends B200InternalModel(Kind.DeepFM, nFields, embeddingDim, fcDims, Array.empty, 0)

class InternalXDeepFMModel(nFields: Int, embeddingDim: Int, fcDims: Array[Int], cinDims: Array[Int])
  extends B200InternalModel(Kind.XDeepFM, nFields, embeddingDim, fcDims, cinDims, 0)

class InternalDCNModel(nFields: Int, embeddingDim: Int, crossDepth: Int, fcDims: Array[Int])
  extends B200InternalModel(Kind.DCN, nFields, embeddingDim, fcDims, Array.empty, crossDepth)

class InternalPNNModel(nFields: Int, embeddingDim: Int, fcDims: Array[Int])
  extends B200InternalModel(Kind.PNN, nFields, embeddingDim, fcDims, Array.empty, 0)
