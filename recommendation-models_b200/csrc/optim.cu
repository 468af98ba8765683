// Optimizer step on the touched table rows and on the dense parameters.
//
// Reference: rec/optim/{AsyncSGD,AsyncMomentum,AsyncAdagrad,AsyncAdam}.scala package
// (gradients, hyper-parameters, slot offset) into Angel's Async{SGD,Momentum,Adagrad,Adam}Func PSFs,
// which run ON THE PS against slot rows interleaved with the weights (rec/model/ParRecModel.scala:
// 75,79,96,115: numSlots = 1 / 2 / 2 / 3).  The PSF arithmetic is third-party (Angel 2.3.1, absent):
// PARITY UNPINNED -- the textbook forms below are used, with the reference's default hyper-parameters
// (AsyncMomentum momentum 0.9; AsyncAdagrad factor 0.9; AsyncAdam gamma 0.99, beta 0.9):
//   sgd      : w -= lr g
//   momentum : v = mu v + g ;                     w -= lr v
//   adagrad  : s = f s + (1-f) g^2 ;               w -= lr g / (sqrt(s) + eps)
//   adam     : m = beta m + (1-beta) g ; v = gamma v + (1-gamma) g^2 ;
//              w -= lr (m / (1-beta^t)) / (sqrt(v / (1-gamma^t)) + eps)
// Only rows whose id occurred in the batch are touched (the PS applies the pushed sparse gradient the
// same way), so the access pattern is the scatter-add's: HBM-bound over U rows.
#include "kernels.h"

namespace b200rec {

constexpr float OPT_EPS = 1e-7f;

__device__ __forceinline__ float opt_update(int kind, float w, float g, float* s1, float* s2, float lr,
                                            float p1, float p2, float c1, float c2) {
  switch (kind) {
    case B200REC_OPT_SGD:
      return w - lr * g;
    case B200REC_OPT_MOMENTUM: {
      const float v = p1 * (*s1) + g;
      *s1 = v;
      return w - lr * v;
    }
    case B200REC_OPT_ADAGRAD: {
      const float s = p1 * (*s1) + (1.f - p1) * g * g;
      *s1 = s;
      return w - lr * g / (sqrtf(s) + OPT_EPS);
    }
    default: {  // adam: p1 = gamma (second moment), p2 = beta (first moment)
      const float m = p2 * (*s1) + (1.f - p2) * g;
      const float v = p1 * (*s2) + (1.f - p1) * g * g;
      *s1 = m;
      *s2 = v;
      return w - lr * (m * c1) / (sqrtf(v * c2) + OPT_EPS);
    }
  }
}

// Adam's bias corrections from a DEVICE update counter (a replayed CUDA graph carries no host step
// number): thread 0 of the block computes them once in double, like bias_corrections() on the host.
__device__ __forceinline__ void device_corrections(int kind, float p1, float p2, const int* step_dev,
                                                   float& c1, float& c2) {
  __shared__ float sc[2];
  if (step_dev == nullptr) return;   // uniform: the corrections came from the host
  if (threadIdx.x == 0) {
    const int step = *step_dev;
    sc[0] = 1.f; sc[1] = 1.f;
    if (kind == B200REC_OPT_ADAM && step > 0) {
      sc[0] = (float)(1.0 / (1.0 - pow((double)p2, (double)step)));
      sc[1] = (float)(1.0 / (1.0 - pow((double)p1, (double)step)));
    }
  }
  __syncthreads();
  c1 = sc[0]; c2 = sc[1];
}

__global__ void __launch_bounds__(256) opt_rows_kernel(int kind, int K, long long rows, const int* n_unique,
                                                       const int* unique, const float* G,
                                                       const float* gw, float lr, float p1, float p2,
                                                       float c1, float c2, const int* step_dev,
                                                       float* table, float* wtable,
                                                       float* s1e, float* s2e, float* s1w, float* s2w,
                                                       int* err) {
  B200_PDL_ENTRY();
  device_corrections(kind, p1, p2, step_dev, c1, c2);
  const int U = *n_unique;
  const int KK = K + 1;  // column K = the first-order weight
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < (long long)U * KK;
       t += (long long)gridDim.x * blockDim.x) {
    const long long seg = t / KK;
    const int k = (int)(t - seg * KK);
    const long long id = unique[seg];
    if (id < 0 || id >= rows) {   // the reference throws on such an id; never write outside the table
      if (err && k == 0) atomicOr(err, DEV_BAD_ID);
      continue;
    }
    if (k < K) {
      if (!G) continue;
      const long long o = id * K + k;
      table[o] = opt_update(kind, table[o], G[seg * K + k], s1e ? s1e + o : nullptr,
                            s2e ? s2e + o : nullptr, lr, p1, p2, c1, c2);
    } else {
      if (!gw || !wtable) continue;
      wtable[id] = opt_update(kind, wtable[id], gw[seg], s1w ? s1w + id : nullptr,
                              s2w ? s2w + id : nullptr, lr, p1, p2, c1, c2);
    }
  }
}

__global__ void __launch_bounds__(256) opt_dense_kernel(int kind, long long n, const float* g, float lr,
                                                        float p1, float p2, float c1, float c2,
                                                        const int* step_dev, float* w,
                                                        float* s1, float* s2) {
  B200_PDL_ENTRY();
  device_corrections(kind, p1, p2, step_dev, c1, c2);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    w[i] = opt_update(kind, w[i], g[i], s1 ? s1 + i : nullptr, s2 ? s2 + i : nullptr, lr, p1, p2, c1, c2);
}

static void bias_corrections(int kind, float p1, float p2, long long step, float* c1, float* c2) {
  *c1 = 1.f;
  *c2 = 1.f;
  if (kind == B200REC_OPT_ADAM && step > 0) {
    *c1 = (float)(1.0 / (1.0 - pow((double)p2, (double)step)));
    *c2 = (float)(1.0 / (1.0 - pow((double)p1, (double)step)));
  }
}

int opt_rows(int kind, int K, long long rows, long long cap, const int* n_unique, const int* unique,
             const float* G, const float* gw, float lr, float p1, float p2, long long step,
             const int* step_dev, float* table, float* wtable, float* s1e, float* s2e, float* s1w,
             float* s2w, int* err, cudaStream_t st) {
  if (cap <= 0) return B200REC_OK;
  float c1, c2;
  bias_corrections(kind, p1, p2, step, &c1, &c2);
  int grid = cdiv(cap * (K + 1), 256);
  if (grid > 148 * 16) grid = 148 * 16;
  ProfTag tag("optimizer");
  B200_LAUNCH(opt_rows_kernel, grid, 256, 0, st, kind, K, rows, n_unique, unique, G, gw, lr, p1, p2, c1, c2,
              step_dev, table, wtable, s1e, s2e, s1w, s2w, err);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

int opt_dense(int kind, long long n, const float* g, float lr, float p1, float p2, long long step,
              const int* step_dev, float* w, float* s1, float* s2, cudaStream_t st) {
  if (n <= 0) return B200REC_OK;
  float c1, c2;
  bias_corrections(kind, p1, p2, step, &c1, &c2);
  int grid = cdiv(n, 256);
  if (grid > 148 * 16) grid = 148 * 16;
  ProfTag tag("optimizer");
  B200_LAUNCH(opt_dense_kernel, grid, 256, 0, st, kind, n, g, lr, p1, p2, c1, c2, step_dev, w, s1, s2);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

}  // namespace b200rec
