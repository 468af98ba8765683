// Dump golden fixtures FROM THE REAL REFERENCE (tests/golden/README.md).  NOT compiled in this repo (no
// JVM in the build image); drop this file into the reference's src/main/scala and build with its pom.
// The Internal<M>Model classes are package-private (e.g. `private[deepfm] class InternalDeepFMModel`,
// rec/model/deepfm/DeepFM.scala:51), hence the one-line hooks declared inside their packages.
package io.yaochi.recommendation.model.lr { object B200Hook { def make() = new InternalLRModel } }
package io.yaochi.recommendation.model.deepfm {
  object B200Hook { def make(f: Int, k: Int, fc: Array[Int]) = new InternalDeepFMModel(f, k, fc) }
}
package io.yaochi.recommendation.model.xdeepfm {
  object B200Hook { def make(f: Int, k: Int, fc: Array[Int], cin: Array[Int]) = new InternalXDeepFMModel(f, k, fc, cin) }
}
package io.yaochi.recommendation.model.dcn {
  object B200Hook { def make(f: Int, k: Int, depth: Int, fc: Array[Int]) = new InternalDCNModel(f, k, depth, fc) }
}
package io.yaochi.recommendation.model.pnn {
  object B200Hook { def make(f: Int, k: Int, fc: Array[Int]) = new InternalPNNModel(f, k, fc) }
}

package io.yaochi.recommendation.b200 {

  import java.io.File
  import java.nio.{ByteBuffer, ByteOrder}
  import java.nio.file.{Files, Paths}

  import io.yaochi.recommendation.model.{dcn, deepfm, lr, pnn, xdeepfm}

  object DumpFixtures {
    private def bytes(p: String): ByteBuffer = ByteBuffer.wrap(Files.readAllBytes(Paths.get(p))).order(ByteOrder.LITTLE_ENDIAN)
    private def floats(p: String): Array[Float] =
      if (!new File(p).exists()) null else { val b = bytes(p).asFloatBuffer(); val a = new Array[Float](b.remaining()); b.get(a); a }
    private def ints(p: String): Array[Int] = { val b = bytes(p).asIntBuffer(); val a = new Array[Int](b.remaining()); b.get(a); a }
    private def write(p: String, a: Array[Float]): Unit = if (a != null) {
      val b = ByteBuffer.allocate(4 * a.length).order(ByteOrder.LITTLE_ENDIAN); b.asFloatBuffer().put(a); Files.write(Paths.get(p), b.array())
    }
    private def dims(s: String): Array[Int] = if (s.isEmpty) Array.empty else s.split(",").map(_.toInt)

    def main(args: Array[String]): Unit = {
      for (dir <- new File(args(0)).listFiles().filter(_.isDirectory).sortBy(_.getName)) {
        val d = dir.getPath + "/"
        val meta = scala.io.Source.fromFile(d + "meta.txt").getLines().map(_.split("=", 2)).map(a => a(0) -> (if (a.length > 1) a(1) else "")).toMap
        val (b, f, k) = (meta("batchSize").toInt, meta("nFields").toInt, meta("embeddingDim").toInt)
        val (fc, cin, depth) = (dims(meta("fcDims")), dims(meta("cinDims")), meta("crossDepth").toInt)
        val index = ints(d + "in_index.bin")
        def in(n: String) = floats(d + s"in_$n.bin")
        val targets = in("targets")
        try {
          // forward on fresh copies, then backward on fresh copies (backward overwrites its inputs)
          val (pred, loss, w, bias, e, m) = meta("kind") match {
            case "lr" =>
              val mdl = lr.B200Hook.make()
              val p = mdl.forward(b, index, in("weights"), in("bias"))
              val (w, bi) = (in("weights"), in("bias"))
              (p, mdl.backward(b, index, w, bi, targets), w, bi, null, null)
            case "deepfm" =>
              val mdl = deepfm.B200Hook.make(f, k, fc)
              val p = mdl.forward(b, index, in("weights"), in("bias"), in("embedding"), in("mats"))
              val (w, bi, e, m) = (in("weights"), in("bias"), in("embedding"), in("mats"))
              (p, mdl.backward(b, index, w, bi, e, m, targets), w, bi, e, m)
            case "xdeepfm" =>
              val mdl = xdeepfm.B200Hook.make(f, k, fc, cin)
              val p = mdl.forward(b, index, in("weights"), in("bias"), in("embedding"), in("mats"))
              val (w, bi, e, m) = (in("weights"), in("bias"), in("embedding"), in("mats"))
              (p, mdl.backward(b, index, w, bi, e, m, targets), w, bi, e, m)
            case "dcn" =>
              val mdl = dcn.B200Hook.make(f, k, depth, fc)
              val p = mdl.forward(b, index, in("weights"), in("bias"), in("embedding"), in("mats"))
              val (w, bi, e, m) = (in("weights"), in("bias"), in("embedding"), in("mats"))
              (p, mdl.backward(b, index, w, bi, e, m, targets), w, bi, e, m)
            case "pnn" =>
              val mdl = pnn.B200Hook.make(f, k, fc)
              val p = mdl.forward(b, index, in("weights"), in("bias"), in("embedding"), in("mats"))
              val (w, bi, e, m) = (in("weights"), in("bias"), in("embedding"), in("mats"))
              (p, mdl.backward(b, index, w, bi, e, m, targets), w, bi, e, m)
            case other => throw new IllegalArgumentException(s"no reference class for kind $other (FM: SURVEY B-1)")
          }
          write(d + "out_pred.bin", pred); write(d + "out_loss.bin", Array(loss))
          write(d + "out_weights.bin", w); write(d + "out_bias.bin", bias)
          write(d + "out_embedding.bin", e); write(d + "out_mats.bin", m)
          println(s"dumped ${dir.getName}: loss $loss")
        } catch {
          case t: Throwable => println(s"FAILED ${dir.getName}: $t")   // expected for multi-layer CIN (SURVEY B-2) and fm
        }
      }
    }
  }
}
