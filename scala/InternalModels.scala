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

/** rec/model/lr/LR.scala:42-90 -- InternalLRModel has no constructor arguments and its forward / backward
  * take no embedding / mats (RecModelType.BIAS_WEIGHT): the same C calls with NULL for both. */
class InternalLRModel extends B200InternalModel(Kind.LR, 0, 0, Array.empty, Array.empty, 0) {
  def forward(batchSize: Int, index: Array[Int], weights: Array[Float], bias: Array[Float]): Array[Float] =
    forward(batchSize, index, weights, bias, null, null)
  def backward(batchSize: Int, index: Array[Int], weights: Array[Float], bias: Array[Float],
               targets: Array[Float]): Float =
    backward(batchSize, index, weights, bias, null, null, targets)
}

/** FM (BASELINE configs[0]): the reference has the BIAS_WEIGHT_EMBEDDING plumbing
  * (rec/model/ParRecModel.scala:401-437, rec/model/RecModel.scala:27-35,74-84) but no model class
  * (SURVEY B-1); this is DeepFM minus the HigherOrderEncoder: first order + SecondOrderEncoder + bias. */
class InternalFMModel(nFields: Int, embeddingDim: Int)
  extends B200InternalModel(Kind.FM, nFields, embeddingDim, Array.empty, Array.empty, 0) {
  def forward(batchSize: Int, index: Array[Int], weights: Array[Float], bias: Array[Float],
              embedding: Array[Float]): Array[Float] =
    forward(batchSize, index, weights, bias, embedding, null)
  def backward(batchSize: Int, index: Array[Int], weights: Array[Float], bias: Array[Float],
               embedding: Array[Float], targets: Array[Float]): Float =
    backward(batchSize, index, weights, bias, embedding, null, targets)
}

class InternalDeepFMModel(nFields: Int, embeddingDim: Int, fcDims: Array[Int])
  extends B200InternalModel(Kind.DeepFM, nFields, embeddingDim, fcDims, Array.empty, 0)

class InternalXDeepFMModel(nFields: Int, embeddingDim: Int, fcDims: Array[Int], cinDims: Array[Int])
  extends B200InternalModel(Kind.XDeepFM, nFields, embeddingDim, fcDims, cinDims, 0)

class InternalDCNModel(nFields: Int, embeddingDim: Int, crossDepth: Int, fcDims: Array[Int])
  extends B200InternalModel(Kind.DCN, nFields, embeddingDim, fcDims, Array.empty, crossDepth)

class InternalPNNModel(nFields: Int, embeddingDim: Int, fcDims: Array[Int])
  extends B200InternalModel(Kind.PNN, nFields, embeddingDim, fcDims, Array.empty, 0)

/** The encoders with the reference's own constructor arguments and forward / backward signatures
  * (rec/model/encoder/HigherOrderEncoder.scala:18-32, xdeepfm/CINEncoder.scala:36,60,
  * dcn/CrossEncoder.scala:40,57, pnn/ProductEncoder.scala:34,43): Tensor storage in, Tensor out, the
  * parameter gradients copied over mats(start ...) by backward like BackwardUtil.linearBackward. */
abstract class B200Encoder(kind: Int, batchSize: Int, nFields: Int, embeddingDim: Int, fcDims: Array[Int],
                           cinDims: Array[Int], crossDepth: Int, mats: Array[Float], start: Int,
                           outputDim: Int) {
  import com.intel.analytics.bigdl.tensor.Tensor
  protected val handle: Pointer = {
    val out = new PointerByReference()
    check(lib.b200rec_model_create(kind, nFields, embeddingDim, fcDims, fcDims.length, cinDims, cinDims.length,
      crossDepth, 0, out))
    out.getValue
  }
  private val matsLen: Int = { val n = new com.sun.jna.ptr.LongByReference(); check(lib.b200rec_encoder_mats_len(handle, n)); n.getValue.toInt }
  protected def fwd(in: Array[Float], p: Array[Float], out: Array[Float]): Int
  protected def bwd(in: Array[Float], p: Array[Float], go: Array[Float], gi: Array[Float]): Int

  def forward(input: Tensor[Float]): Tensor[Float] = {
    val out = new Array[Float](batchSize * outputDim)
    check(fwd(input.contiguous().storage().array(), java.util.Arrays.copyOfRange(mats, start, start + matsLen), out))
    Tensor(out, Array(batchSize, outputDim))
  }
  def backward(input: Tensor[Float], gradOutput: Tensor[Float]): Tensor[Float] = {
    val p = java.util.Arrays.copyOfRange(mats, start, start + matsLen)
    val gi = new Array[Float](batchSize * nFields * embeddingDim)
    check(bwd(input.contiguous().storage().array(), p, gradOutput.contiguous().storage().array(), gi))
    System.arraycopy(p, 0, mats, start, matsLen)          // BackwardUtil.linearBackward: grads over mats
    Tensor(gi, Array(batchSize, nFields * embeddingDim))
  }
}
class HigherOrderEncoder(batchSize: Int, inputDim: Int, fcDims: Array[Int], mats: Array[Float], start: Int = 0)
  extends B200Encoder(Kind.DeepFM, batchSize, inputDim, 1, fcDims, Array.empty, 0, mats, start, 1) {
  protected def fwd(in: Array[Float], p: Array[Float], out: Array[Float]): Int = lib.b200rec_higher_order_update_output(handle, batchSize, in, p, out)
  protected def bwd(in: Array[Float], p: Array[Float], go: Array[Float], gi: Array[Float]): Int = lib.b200rec_higher_order_backward(handle, batchSize, in, p, go, gi)
}
class CINEncoder(batchSize: Int, nFields: Int, embeddingDim: Int, fcDims: Array[Int], cinDims: Array[Int],
                 mats: Array[Float], start: Int = 0)
  extends B200Encoder(Kind.XDeepFM, batchSize, nFields, embeddingDim, fcDims, cinDims, 0, mats, start, 1) {
  protected def fwd(in: Array[Float], p: Array[Float], out: Array[Float]): Int = lib.b200rec_cin_update_output(handle, batchSize, in, p, out)
  protected def bwd(in: Array[Float], p: Array[Float], go: Array[Float], gi: Array[Float]): Int = lib.b200rec_cin_backward(handle, batchSize, in, p, go, gi)
}
class CrossEncoder(batchSize: Int, nFields: Int, embeddingDim: Int, crossDepth: Int, fcDims: Array[Int],
                   mats: Array[Float], start: Int = 0)
  extends B200Encoder(Kind.DCN, batchSize, nFields, embeddingDim, fcDims, Array.empty, crossDepth, mats, start, 1) {
  protected def fwd(in: Array[Float], p: Array[Float], out: Array[Float]): Int = lib.b200rec_cross_update_output(handle, batchSize, in, p, out)
  protected def bwd(in: Array[Float], p: Array[Float], go: Array[Float], gi: Array[Float]): Int = lib.b200rec_cross_backward(handle, batchSize, in, p, go, gi)
}
class ProductEncoder(batchSize: Int, nFields: Int, embeddingDim: Int, outputDim: Int, mats: Array[Float],
                     start: Int = 0)
  extends B200Encoder(Kind.PNN, batchSize, nFields, embeddingDim, Array(outputDim), Array.empty, 0, mats, start, outputDim) {
  protected def fwd(in: Array[Float], p: Array[Float], out: Array[Float]): Int = lib.b200rec_product_update_output(handle, batchSize, in, p, out)
  protected def bwd(in: Array[Float], p: Array[Float], go: Array[Float], gi: Array[Float]): Int = lib.b200rec_product_backward(handle, batchSize, in, p, go, gi)
}

/** The resident replacement of ParRecModel's pull / optimize / push cycle
  * (rec/model/ParRecModel.scala:439-478): table and dense params live on the GPU, a step is one call.
  * `stage` + `stepAsync` + `waitLoss` is the asynchronous form (cf. asyncPullEmbeddings /
  * asyncPushEmbedding, :193-196, :261-264): the next batch is copied under the running step and every
  * loss still reaches the JVM, one step late.  Arrays passed to `stage` must be direct / pinned
  * buffers that stay alive until the matching `waitLoss` (JNA `Memory`), not JVM heap arrays. */
class B200ResidentTrainer(model: Pointer, table: Pointer) {
  def optimize(batchSize: Int, feats: Array[Int], targets: Array[Float]): Float = {
    val loss = new FloatByReference()
    check(lib.b200rec_step(model, table, batchSize, feats, targets, loss))
    loss.getValue * batchSize                       // ParRecModel.scala:477 returns loss * batchSize
  }
  def stage(batchSize: Int, feats: Pointer, targets: Pointer): Unit =
    check(lib.b200rec_stage_batch(model, batchSize, feats, targets))
  def stepAsync(): Unit = check(lib.b200rec_step_staged_async(model, table))
  def waitLoss(): Float = {
    val loss = new FloatByReference()
    check(lib.b200rec_step_wait(model, table, loss))
    loss.getValue
  }
}

/** rec/data/SampleParser.scala:23-85 through the native, multi-threaded parser. */
object B200SampleParser {
  def parse(text: Array[Byte], libffm: Boolean): (Array[Int], Array[Int], Array[Float], Array[Float]) = {
    val ns = new com.sun.jna.ptr.LongByReference(); val nz = new com.sun.jna.ptr.LongByReference()
    check(lib.b200rec_parse_samples(if (libffm) 1 else 0, text, text.length, 0, 0, null, null, null, null, null, ns, nz))
    val targets = new Array[Float](ns.getValue.toInt)
    val index = new Array[Int](nz.getValue.toInt); val feats = new Array[Int](nz.getValue.toInt)
    val values = new Array[Float](nz.getValue.toInt)
    check(lib.b200rec_parse_samples(if (libffm) 1 else 0, text, text.length, targets.length, index.length,
      targets, index, feats, null, values, ns, nz))
    (index, feats, values, targets)
  }
}
