// DCN cross layers and PNN inner products: per-sample fused kernels (HBM-bound; the row stays in
// registers / shared memory across all layers or pairs).
//
// Reference:
//   rec/model/dcn/CrossEncoder.scala:40-55,111-132  x_{l+1} = x0 * (x_l . w_l) + x_l + c_l
//     (Linear(D->1) no bias :138, MM :118, CAddTable + scalar CAdd :128-129,148)
//   rec/model/dcn/CrossEncoder.scala:57-105         its backward
//   rec/model/pnn/ProductEncoder.scala:84-89,110-120  ip[b,p] = <v_i, v_j>, i<j lexicographic
//   nn/Gather.scala:50-78 + nn/DotProduct2.scala:28-53  backward of the pair gather / dot
// The reference materialises x_l per layer and two [B,P,K] pair copies (2 x 388 MB at B=8192);
// here nothing but the [B,D] input and output ever touches HBM.
#include "kernels.h"

namespace b200rec {

// ------------------------------------------------------------------------------------------------
// DCN cross forward: one warp per sample, NPL elements per lane (D <= 32*NPL).
// ------------------------------------------------------------------------------------------------
template <int NPL>
__global__ void __launch_bounds__(256) cross_fwd_kernel(int B, int D, int L, const float* X,
                                                        const float* w, const float* c, float* xL,
                                                        float* s_out) {
  B200_PDL_ENTRY();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int b = blockIdx.x * wpb + (threadIdx.x >> 5); b < B; b += gridDim.x * wpb) {
    float x0[NPL], x[NPL];
#pragma unroll
    for (int q = 0; q < NPL; ++q) {
      const int d = q * 32 + lane;
      x0[q] = d < D ? X[(long long)b * D + d] : 0.f;
      x[q] = x0[q];
    }
    for (int l = 0; l < L; ++l) {
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < NPL; ++q) {
        const int d = q * 32 + lane;
        if (d < D) s = fmaf(x[q], __ldg(w + (long long)l * D + d), s);
      }
      s = warp_sum(s);
      if (lane == 0) s_out[(long long)l * B + b] = s;
      const float cl = __ldg(c + l);
#pragma unroll
      for (int q = 0; q < NPL; ++q) x[q] = fmaf(x0[q], s, x[q]) + cl;
    }
#pragma unroll
    for (int q = 0; q < NPL; ++q) {
      const int d = q * 32 + lane;
      if (d < D) xL[(long long)b * D + d] = x[q];
    }
  }
}

// backward: per sample, walks the layers in reverse.  Per-sample scalars for the parameter
// gradients go to u[l,b] = gs_l[b] * alpha_l[b], gs[l,b], gcs[l,b] (sum_d g) and are reduced over
// the batch by cross_dw_* below, using x_l = x0 * alpha_l + beta_l with
// alpha_l = 1 + sum_{m<l} s_m, beta_l = sum_{m<l} c_m (closed form of the recurrence).
template <int NPL>
__global__ void __launch_bounds__(256) cross_bwd_kernel(int B, int D, int L, const float* X,
                                                        const float* w, const float* s_in,
                                                        const float* g_xL, float* dX, float* u,
                                                        float* gs_out, float* gcs) {
  B200_PDL_ENTRY();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int b = blockIdx.x * wpb + (threadIdx.x >> 5); b < B; b += gridDim.x * wpb) {
    float x0[NPL], g[NPL], gx0[NPL];
#pragma unroll
    for (int q = 0; q < NPL; ++q) {
      const int d = q * 32 + lane;
      x0[q] = d < D ? X[(long long)b * D + d] : 0.f;
      g[q] = d < D ? g_xL[(long long)b * D + d] : 0.f;
      gx0[q] = 0.f;
    }
    float alpha = 1.f;
    for (int l = 0; l < L; ++l) alpha += s_in[(long long)l * B + b];
    for (int l = L - 1; l >= 0; --l) {
      const float sl = s_in[(long long)l * B + b];
      alpha -= sl;  // alpha_l = 1 + sum_{m<l} s_m
      float gsum = 0.f, gs = 0.f;
#pragma unroll
      for (int q = 0; q < NPL; ++q) {
        gsum += g[q];
        gs = fmaf(g[q], x0[q], gs);
        gx0[q] = fmaf(g[q], sl, gx0[q]);
      }
      gsum = warp_sum(gsum);
      gs = warp_sum(gs);
      if (lane == 0) {
        gcs[(long long)l * B + b] = gsum;
        gs_out[(long long)l * B + b] = gs;
        u[(long long)l * B + b] = gs * alpha;
      }
#pragma unroll
      for (int q = 0; q < NPL; ++q) {
        const int d = q * 32 + lane;
        if (d < D) g[q] = fmaf(gs, __ldg(w + (long long)l * D + d), g[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < NPL; ++q) {
      const int d = q * 32 + lane;
      if (d < D) dX[(long long)b * D + d] = gx0[q] + g[q];
    }
  }
}

constexpr int CROSS_MAXL = 16;
constexpr int CROSS_CHUNKS = 64;

// stage 1: part[c][l][d] = sum_{b in chunk c} u[l,b] X[b,d];  ps[c][l] = sum gs, pc[c][l] = sum gcs
__global__ void cross_dw_stage1(int B, int D, int L, const float* X, const float* u,
                                const float* gs, const float* gcs, int rows_per_chunk, float* part,
                                float* ps, float* pc) {
  B200_PDL_ENTRY();
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y;
  const int r0 = c * rows_per_chunk, r1 = min(B, r0 + rows_per_chunk);
  if (d < D) {
    float acc[CROSS_MAXL];
#pragma unroll
    for (int l = 0; l < CROSS_MAXL; ++l) acc[l] = 0.f;
    for (int b = r0; b < r1; ++b) {
      const float x = X[(long long)b * D + d];
#pragma unroll
      for (int l = 0; l < CROSS_MAXL; ++l)
        if (l < L) acc[l] = fmaf(__ldg(u + (long long)l * B + b), x, acc[l]);
    }
#pragma unroll
    for (int l = 0; l < CROSS_MAXL; ++l)
      if (l < L) part[((long long)c * L + l) * D + d] = acc[l];
  }
  if (blockIdx.x == 0 && threadIdx.x < 2 * L) {
    const int l = threadIdx.x % L;
    const float* src = threadIdx.x < L ? gs : gcs;
    float s = 0.f;
    for (int b = r0; b < r1; ++b) s += src[(long long)l * B + b];
    (threadIdx.x < L ? ps : pc)[c * L + l] = s;
  }
}
// stage 2: gw[l,d] = sum_c part + beta_l * sum_c ps ;  gc[l] = sum_c pc
__global__ void cross_dw_stage2(int D, int L, const float* cvec, const float* part, const float* ps,
                                const float* pc, float* gw, float* gc) {
  B200_PDL_ENTRY();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < L * D) {
    const int l = t / D, d = t - l * D;
    float s = 0.f, sg = 0.f, beta = 0.f;
    for (int c = 0; c < CROSS_CHUNKS; ++c) {
      s += part[((long long)c * L + l) * D + d];
      sg += ps[c * L + l];
    }
    for (int m = 0; m < l; ++m) beta += __ldg(cvec + m);
    gw[t] = fmaf(beta, sg, s);
  }
  if (t < L) {
    float s = 0.f;
    for (int c = 0; c < CROSS_CHUNKS; ++c) s += pc[c * L + t];
    gc[t] = s;
  }
}

static int cross_npl(int D) {
  const int need = (D + 31) / 32;
  if (need <= 8) return 8;
  if (need <= 16) return 16;
  if (need <= 32) return 32;
  if (need <= 64) return 64;
  return 0;
}

int cross_fwd(int B, int D, int L, const float* X, const float* w, const float* c, float* xL,
              float* s, cudaStream_t st) {
  B200_REQUIRE(L <= CROSS_MAXL, B200REC_ERR_ARG, "cross depth %d > %d", L, CROSS_MAXL);
  int grid = cdiv((long long)B * 32, 256);
  if (grid > 148 * 8) grid = 148 * 8;
  if (grid < 1) grid = 1;
  switch (cross_npl(D)) {
    case 8: B200_LAUNCH(cross_fwd_kernel<8>, grid, 256, 0, st, B, D, L, X, w, c, xL, s); break;
    case 16: B200_LAUNCH(cross_fwd_kernel<16>, grid, 256, 0, st, B, D, L, X, w, c, xL, s); break;
    case 32: B200_LAUNCH(cross_fwd_kernel<32>, grid, 256, 0, st, B, D, L, X, w, c, xL, s); break;
    case 64: B200_LAUNCH(cross_fwd_kernel<64>, grid, 256, 0, st, B, D, L, X, w, c, xL, s); break;
    default: set_error("cross: nFields*embeddingDim = %d > 2048 unsupported", D); return B200REC_ERR_ARG;
  }
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

int cross_bwd(int B, int D, int L, const float* X, const float* w, const float* c, const float* s,
              const float* g_xL, float* dX, float* gw, float* gc, DevBuf& scratch,
              cudaStream_t st) {
  const size_t n_lb = (size_t)L * B;
  const size_t floats = 3 * n_lb + (size_t)CROSS_CHUNKS * L * D + 2 * (size_t)CROSS_CHUNKS * L;
  B200_TRY(scratch.reserve(floats * sizeof(float)));
  float* u = scratch.as<float>();
  float* gs = u + n_lb;
  float* gcs = gs + n_lb;
  float* part = gcs + n_lb;
  float* ps = part + (size_t)CROSS_CHUNKS * L * D;
  float* pc = ps + (size_t)CROSS_CHUNKS * L;
  int grid = cdiv((long long)B * 32, 256);
  if (grid > 148 * 8) grid = 148 * 8;
  if (grid < 1) grid = 1;
  switch (cross_npl(D)) {
    case 8: B200_LAUNCH(cross_bwd_kernel<8>, grid, 256, 0, st, B, D, L, X, w, s, g_xL, dX, u, gs, gcs); break;
    case 16: B200_LAUNCH(cross_bwd_kernel<16>, grid, 256, 0, st, B, D, L, X, w, s, g_xL, dX, u, gs, gcs); break;
    case 32: B200_LAUNCH(cross_bwd_kernel<32>, grid, 256, 0, st, B, D, L, X, w, s, g_xL, dX, u, gs, gcs); break;
    case 64: B200_LAUNCH(cross_bwd_kernel<64>, grid, 256, 0, st, B, D, L, X, w, s, g_xL, dX, u, gs, gcs); break;
    default: set_error("cross: nFields*embeddingDim = %d > 2048 unsupported", D); return B200REC_ERR_ARG;
  }
  dim3 g1(cdiv(D, 128), CROSS_CHUNKS);
  B200_LAUNCH(cross_dw_stage1, g1, 128, 0, st, B, D, L, X, u, gs, gcs, cdiv(B, CROSS_CHUNKS), part,
              ps, pc);
  B200_LAUNCH(cross_dw_stage2, cdiv((long long)L * D, 128), 128, 0, st, D, L, c, part, ps, pc, gw,
              gc);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

// ------------------------------------------------------------------------------------------------
// PNN inner products.  One CTA per sample; the [F,K] tile of the sample sits in shared memory.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int pair_index(int i, int j, int F) {  // i < j
  return i * F - (i * (i + 1)) / 2 + (j - i - 1);
}

__global__ void __launch_bounds__(128) pnn_ip_fwd_kernel(int B, int F, int K, const float* X,
                                                         float* ip) {
  B200_PDL_ENTRY();
  extern __shared__ float sv[];  // [F][K+1]
  const int P = F * (F - 1) / 2;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    for (int t = threadIdx.x; t < F * K; t += blockDim.x)
      sv[(t / K) * (K + 1) + (t % K)] = X[(long long)b * F * K + t];
    __syncthreads();
    // thread -> pair p; recover (i,j) by walking rows
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
      int i = 0, rem = p;
      while (rem >= F - 1 - i) { rem -= F - 1 - i; ++i; }
      const int j = i + 1 + rem;
      float acc = 0.f;
      for (int k = 0; k < K; ++k)  // cmul then sum (DotProduct2.scala:24-25): no fma
        acc = __fadd_rn(acc, __fmul_rn(sv[i * (K + 1) + k], sv[j * (K + 1) + k]));
      ip[(long long)b * P + p] = acc;
    }
    __syncthreads();
  }
}

// dX[b,i,k] (+)= sum_{j<i} gip[pair(j,i)] v[j,k] + sum_{j>i} gip[pair(i,j)] v[j,k]  (j ascending:
// the order nn/Gather.scala:70-76 accumulates, p ascending)
__global__ void __launch_bounds__(128) pnn_ip_bwd_kernel(int B, int F, int K, const float* X,
                                                         const float* gip, float* dX,
                                                         bool accumulate) {
  B200_PDL_ENTRY();
  extern __shared__ float sm[];  // v [F][K+1] then g [P]
  const int P = F * (F - 1) / 2;
  float* sv = sm;
  float* sg = sm + F * (K + 1);
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    for (int t = threadIdx.x; t < F * K; t += blockDim.x)
      sv[(t / K) * (K + 1) + (t % K)] = X[(long long)b * F * K + t];
    for (int p = threadIdx.x; p < P; p += blockDim.x) sg[p] = gip[(long long)b * P + p];
    __syncthreads();
    for (int t = threadIdx.x; t < F * K; t += blockDim.x) {
      const int i = t / K, k = t - i * K;
      float acc = 0.f;
      for (int j = 0; j < i; ++j) acc = __fadd_rn(acc, __fmul_rn(sg[pair_index(j, i, F)], sv[j * (K + 1) + k]));
      for (int j = i + 1; j < F; ++j) acc = __fadd_rn(acc, __fmul_rn(sg[pair_index(i, j, F)], sv[j * (K + 1) + k]));
      float* o = dX + (long long)b * F * K + t;
      *o = accumulate ? *o + acc : acc;
    }
    __syncthreads();
  }
}

int pnn_ip_fwd(int B, int F, int K, const float* X, float* ip, cudaStream_t st) {
  if (B <= 0) return B200REC_OK;
  const size_t smem = (size_t)F * (K + 1) * sizeof(float);
  int grid = B < 148 * 16 ? B : 148 * 16;
  B200_LAUNCH(pnn_ip_fwd_kernel, grid, 128, smem, st, B, F, K, X, ip);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}
int pnn_ip_bwd(int B, int F, int K, const float* X, const float* gip, float* dX, bool accumulate,
               cudaStream_t st) {
  if (B <= 0) return B200REC_OK;
  const size_t smem = ((size_t)F * (K + 1) + (size_t)F * (F - 1) / 2) * sizeof(float);
  int grid = B < 148 * 16 ? B : 148 * 16;
  B200_LAUNCH(pnn_ip_bwd_kernel, grid, 128, smem, st, B, F, K, X, gip, dX, accumulate);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

}  // namespace b200rec
