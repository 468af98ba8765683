// Output head shared by all models: CAddTable -> Sigmoid -> BCECriterion and their backward.
//
// Reference: rec/model/deepfm/DeepFM.scala:105-117,127-134 (identical blocks in LR.scala:72-83,
// XDeepFM.scala:107-118, DCN.scala:111-122, PNN.scala:112-124).  BigDL's BCECriterion (third
// party, restated in oracle/refport.py): loss = -mean[t log(p+eps) + (1-t) log((1+eps)-p)],
// grad = (p-t) / (((1+eps)-p)(p+eps)) / B with eps = 1e-12 rounded to fp32; Sigmoid backward
// multiplies by (1-p) p.  Labels are thresholded `label > 0` (DeepFM.scala:106).
#include "kernels.h"

namespace b200rec {

__global__ void __launch_bounds__(256) head_kernel(Head h, float* part) {
  B200_PDL_ENTRY();
  __shared__ float sh_l[8], sh_b[8];
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  float li = 0.f, gi = 0.f;
  if (b < h.B) {
    float z = h.br[0][b];
#pragma unroll
    for (int i = 1; i < 4; ++i)   // static indices: a dynamic one would spill the parameter struct to local memory
      if (i < h.n_br) z += h.br[i][b];
    z += __ldg(h.bias);
    const float p = 1.0f / (1.0f + expf(-z));
    if (h.preds) h.preds[b] = p;
    if (h.targets) {
      const float eps = 1e-12f;
      const float one_eps = (float)(1.0 + 1e-12);  // == 1.0f, as ev.fromType(1.0 + eps)
      const float t = h.targets[b] > 0.f ? 1.f : 0.f;
      li = t * logf(p + eps) + (1.f - t) * logf(one_eps - p);
      const float g = (p - t) / ((one_eps - p) * (p + eps)) * (1.0f / (float)h.B);
      gi = g * ((1.f - p) * p);
      h.dlogit[b] = gi;
    }
  }
  if (h.targets) {
    li = warp_sum(li);
    gi = warp_sum(gi);
    if ((threadIdx.x & 31) == 0) { sh_l[threadIdx.x >> 5] = li; sh_b[threadIdx.x >> 5] = gi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float a = 0.f, c = 0.f;
      for (int i = 0; i < (blockDim.x >> 5); ++i) { a += sh_l[i]; c += sh_b[i]; }
      part[2 * blockIdx.x] = a;
      part[2 * blockIdx.x + 1] = c;
    }
  }
}

__global__ void head_finish_kernel(int nparts, int B, const float* part, float* loss, float* dbias,
                                   float* dbias2) {
  B200_PDL_ENTRY();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double a = 0.0;  // BCECriterion accumulates the two dot products in a double
    float c = 0.f;
    for (int i = 0; i < nparts; ++i) { a += (double)part[2 * i]; c += part[2 * i + 1]; }
    if (loss) loss[0] = (float)(-a / (double)B);
    if (dbias) dbias[0] = c;
    if (dbias2) dbias2[0] = c;
  }
}

int head_run(const Head& h, DevBuf& scratch, cudaStream_t st) {
  if (h.B <= 0) return B200REC_OK;
  const int blocks = cdiv(h.B, 256);
  B200_TRY(scratch.reserve((size_t)blocks * 2 * sizeof(float)));
  B200_LAUNCH(head_kernel, blocks, 256, 0, st, h, scratch.as<float>());
  if (h.targets)
    B200_LAUNCH(head_finish_kernel, 1, 32, 0, st, blocks, h.B, scratch.as<float>(), h.loss, h.dbias, h.dbias2);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

}  // namespace b200rec
