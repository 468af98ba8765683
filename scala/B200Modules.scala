// BigDL AbstractModule wrappers over the C ABI for the reference's custom modules (NOT compiled in
// this repo: no JVM / BigDL jars in the build image).  They live in BigDL's package like the
// originals (nn/Scatter.scala, nn/Gather.scala, nn/DotProduct2.scala).
package com.intel.analytics.bigdl.nn

import com.intel.analytics.bigdl.nn.abstractnn.AbstractModule
import com.intel.analytics.bigdl.tensor.Tensor
import com.intel.analytics.bigdl.utils.{T, Table}
import io.yaochi.recommendation.b200.B200Rec.{check, lib}

/** nn/Scatter.scala:17-59 -- output[index(i), :] += input(i, :);  gradInput(i, :) = gradOutput(index(i), :) */
class B200Scatter(batchSize: Int, nOutput: Int, device: Int = 0)
  extends AbstractModule[Table, Tensor[Float], Float] {

  private def ints(t: Tensor[Int]): Array[Int] = t.contiguous().storage().array()

  override def updateOutput(input: Table): Tensor[Float] = {
    val x = input[Tensor[Float]](1).contiguous()
    val index = input[Tensor[Int]](2)
    output.resize(batchSize, nOutput)
    check(lib.b200rec_scatter_update_output(device, batchSize, nOutput, index.nElement(),
      x.storage().array(), ints(index), output.storage().array()))
    output
  }

  override def updateGradInput(input: Table, gradOutput: Tensor[Float]): Table = {
    val index = input[Tensor[Int]](2)
    val g = Tensor[Float](index.nElement(), nOutput)
    check(lib.b200rec_scatter_update_grad_input(device, batchSize, nOutput, index.nElement(),
      ints(index), gradOutput.contiguous().storage().array(), g.storage().array()))
    gradInput = T(g, Tensor[Int]())
    gradInput
  }
  // parameter-free: accGradParameters is the inherited no-op
}
