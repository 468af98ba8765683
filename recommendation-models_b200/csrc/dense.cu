// Dense (Linear) layers on the exact-fp32 SIMT GEMM, plus the small reductions around them.
//
// Reference: BigDL Linear as instantiated by rec/util/LayerUtil.scala:7-24 (y = x W^T + b with
// W:[out,in] row-major inside `mats`), its gradients copied back by rec/util/BackwardUtil.scala:6-31,
// stacked by rec/model/encoder/HigherOrderEncoder.scala:34-58.
#include "gemm_simt.cuh"
#include "kernels.h"

namespace b200rec {

__global__ void splitk_reduce_kernel(const float* ws, int splits, long long MN, float scale,
                                     bool accumulate, float* out) {
  B200_PDL_ENTRY();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < MN;
       i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += ws[z * MN + i];
    s *= scale;
    out[i] = accumulate ? out[i] + s : s;
  }
}

static int grid1d(long long n, int threads = 256, int cap = 148 * 16) {
  int g = cdiv(n, threads);
  if (g > cap) g = cap;
  return g < 1 ? 1 : g;
}

int splitk_reduce(const float* ws, int splits, long long MN, float scale, bool accumulate,
                  float* out, cudaStream_t st) {
  B200_LAUNCH(splitk_reduce_kernel, grid1d(MN), 256, 0, st, ws, splits, MN, scale, accumulate, out);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

// the same for a weight-gradient GEMM that carried the bias gradient as column K of its [N x ldw] partials
__global__ void splitk_reduce_wb_kernel(const float* ws, int splits, int N, int K, int ldw, float scale,
                                        bool accumulate, float* gw, float* gb) {
  B200_PDL_ENTRY();
  const long long total = (long long)N * (K + 1), MN = (long long)N * ldw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / (K + 1)), k = (int)(i - (long long)n * (K + 1));
    const float* p = ws + (long long)n * ldw + k;
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += p[z * MN];
    s *= scale;
    float* out = k < K ? gw + (long long)n * K + k : gb + n;
    *out = accumulate ? *out + s : s;
  }
}
int splitk_reduce_wb(const float* ws, int splits, int N, int K, int ldw, float scale, bool accumulate,
                     float* gw, float* gb, cudaStream_t st) {
  B200_LAUNCH(splitk_reduce_wb_kernel, grid1d((long long)N * (K + 1)), 256, 0, st, ws, splits, N, K, ldw, scale,
              accumulate, gw, gb);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

int linear_fwd(int M, int N, int K, const float* x, const float* w, const float* b, bool relu,
               float* y, cudaStream_t st, int mode) {
  if (N > 1 && mode) return tc_linear_fwd(M, N, K, x, w, b, relu, y, mode == 1 ? 3 : 1, st);
  if (N == 1) {  // Linear(K -> 1): a row dot product
    B200_TRY(gemv_rows(M, K, x, K, w, b, false, y, st));
    if (relu) return add_bias_relu(M, y, nullptr, y, st);
    return B200REC_OK;
  }
  return gemm_simt(M, N, K, 1, RowMajorOp{x, K}, RowMajorOp{w, K}, EpBiasAct{y, N, b, relu}, st);
}

int linear_bwd_input(int M, int N, int K, const float* gy, const float* w, const float* mask,
                     float* gx, bool accumulate, cudaStream_t st, int mode) {
  if (mode) return tc_linear_bwd_input(M, N, K, gy, w, mask, gx, accumulate, mode == 1 ? 3 : 1, st);
  // gx[M,K] = gy[M,N] W[N,K]: contraction over N
  return gemm_simt(M, K, N, 1, RowMajorOp{gy, N}, ColMajorOp{w, K},
                   EpMaskAcc{gx, K, mask, K, accumulate}, st);
}

// column sums of gy[M,N] in two fixed-order stages
__global__ void colsum_stage1(int M, int N, const float* g, int rows_per_chunk, float* part) {
  B200_PDL_ENTRY();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y;
  if (n >= N) return;
  const int r0 = c * rows_per_chunk, r1 = min(M, r0 + rows_per_chunk);
  float s = 0.f;
  int r = r0;
  for (; r + 8 <= r1; r += 8) {  // 8 independent loads in flight, added in row order
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = g[(long long)(r + u) * N + n];
#pragma unroll
    for (int u = 0; u < 8; ++u) s += v[u];
  }
  for (; r < r1; ++r) s += g[(long long)r * N + n];
  part[(long long)c * N + n] = s;
}
// one warp per column: lanes stride over the chunks, then a fixed xor tree (deterministic)
__global__ void colsum_stage2(int N, int chunks, const float* part, float scale, bool accumulate,
                              float* out) {
  B200_PDL_ENTRY();
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  float s = 0.f;
  for (int c = lane; c < chunks; c += 32) s += part[(long long)c * N + n];
  s = warp_sum(s);
  if (lane == 0) {
    s *= scale;
    out[n] = accumulate ? out[n] + s : s;
  }
}

int colsum(int M, int N, const float* g, float scale, bool accumulate, float* out, float* part,
           cudaStream_t st) {
  // row chunks of >= 32 rows, at most COLSUM_CHUNKS of them: enough CTAs to pull HBM bandwidth
  int chunks = M / 32;
  if (chunks > COLSUM_CHUNKS) chunks = COLSUM_CHUNKS;
  if (chunks < 1) chunks = 1;
  const int rows_per_chunk = cdiv(M, chunks);
  chunks = cdiv(M, rows_per_chunk);
  dim3 g1(cdiv(N, 128), chunks);
  B200_LAUNCH(colsum_stage1, g1, 128, 0, st, M, N, g, rows_per_chunk, part);
  B200_LAUNCH(colsum_stage2, cdiv((long long)N * 32, 256), 256, 0, st, N, chunks, part, scale, accumulate, out);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

int linear_bwd_params(int M, int N, int K, const float* x, const float* gy, float scale,
                      bool accumulate, float* gw, float* gb, DevBuf& scratch, cudaStream_t st,
                      int mode) {
  if (mode)
    return tc_linear_bwd_params(M, N, K, x, gy, scale, accumulate, gw, gb, scratch, mode == 1 ? 3 : 1, st);
  // gw[N,K] = gy^T x : output N x K, contraction over the batch M
  const int splits = pick_splits(N, K, M);
  const long long MN = (long long)N * K;
  B200_TRY(scratch.reserve(((size_t)splits * MN + (size_t)COLSUM_CHUNKS * N) * sizeof(float)));
  float* ws = scratch.as<float>();
  B200_TRY(gemm_simt(N, K, M, splits, ColMajorOp{gy, N}, ColMajorOp{x, K}, EpPartial{ws, MN, K},
                     st));
  B200_TRY(splitk_reduce(ws, real_splits(M, splits), MN, scale, accumulate, gw, st));
  if (gb) B200_TRY(colsum(M, N, gy, scale, accumulate, gb, ws + (size_t)splits * MN, st));
  return B200REC_OK;
}

// out[m] (+)= x[m,:] . w + b0      one warp per row
__global__ void gemv_rows_kernel(int M, int K, const float* x, int ldx, const float* w,
                                 const float* b0, bool accumulate, float* out) {
  B200_PDL_ENTRY();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int m = blockIdx.x * wpb + (threadIdx.x >> 5); m < M; m += gridDim.x * wpb) {
    float s = 0.f;
    for (int k = lane; k < K; k += 32) s = fmaf(x[(long long)m * ldx + k], __ldg(w + k), s);
    s = warp_sum(s);
    if (lane == 0) {
      if (b0) s += __ldg(b0);
      out[m] = accumulate ? out[m] + s : s;
    }
  }
}
int gemv_rows(int M, int K, const float* x, int ldx, const float* w, const float* b0,
              bool accumulate, float* out, cudaStream_t st) {
  if (M <= 0) return B200REC_OK;
  B200_LAUNCH(gemv_rows_kernel, grid1d((long long)M * 32), 256, 0, st, M, K, x, ldx, w, b0,
              accumulate, out);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

// g[m,k] = d[m] * w[k] (* (mask[m,k] > 0))
__global__ void outer_rows_kernel(int M, int K, const float* d, const float* w, const float* mask,
                                  int ldm, float* g, int ldg) {
  B200_PDL_ENTRY();
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < (long long)M * K;
       t += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(t / K), k = (int)(t - (long long)m * K);
    float v = __ldg(d + m) * __ldg(w + k);
    if (mask && !(mask[(long long)m * ldm + k] > 0.f)) v = 0.f;
    g[(long long)m * ldg + k] = v;
  }
}
int outer_rows(int M, int K, const float* d, const float* w, const float* mask, int ldm, float* g,
               int ldg, cudaStream_t st) {
  if ((long long)M * K <= 0) return B200REC_OK;
  B200_LAUNCH(outer_rows_kernel, grid1d((long long)M * K), 256, 0, st, M, K, d, w, mask, ldm, g,
              ldg);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

// gw[k] = sum_m d[m] x[m,k]   (two fixed-order stages)
__global__ void wcolsum_stage1(int M, int K, const float* d, const float* x, int ldx,
                               int rows_per_chunk, float* part) {
  B200_PDL_ENTRY();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y;
  if (k >= K) return;
  const int r0 = c * rows_per_chunk, r1 = min(M, r0 + rows_per_chunk);
  float s = 0.f;
  for (int r = r0; r < r1; ++r) s = fmaf(__ldg(d + r), x[(long long)r * ldx + k], s);
  part[(long long)c * K + k] = s;
}
int wcolsum(int M, int K, const float* d, const float* x, int ldx, float* gw, DevBuf& scratch,
            cudaStream_t st) {
  int chunks = M / 32;
  if (chunks > COLSUM_CHUNKS) chunks = COLSUM_CHUNKS;
  if (chunks < 1) chunks = 1;
  const int rows_per_chunk = cdiv(M, chunks);
  chunks = cdiv(M, rows_per_chunk);
  B200_TRY(scratch.reserve((size_t)chunks * K * sizeof(float)));
  float* part = scratch.as<float>();
  dim3 g1(cdiv(K, 128), chunks);
  B200_LAUNCH(wcolsum_stage1, g1, 128, 0, st, M, K, d, x, ldx, rows_per_chunk, part);
  B200_LAUNCH(colsum_stage2, cdiv((long long)K * 32, 256), 256, 0, st, K, chunks, part, 1.0f, false, gw);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

// Linear(K -> 1) backward in ONE pass over its input a[M,K] (HigherOrderEncoder.scala:46-58, the
// "-> 1" layer): g[m,k] = d[m] w[k] (a[m,k] > 0 if masked) for the layer below, and the chunk partials
// of gw[k] = sum_m d[m] a[m,k] and gb = sum_m d[m]; colsum_stage2 finishes [gw | gb] (adjacent in mats).
__global__ void __launch_bounds__(128) head_bwd_kernel(int M, int K, const float* d, const float* a,
                                                       const float* w, bool mask, float* g,
                                                       int rows_per_chunk, float* part) {
  B200_PDL_ENTRY();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y;
  const int r0 = c * rows_per_chunk, r1 = min(M, r0 + rows_per_chunk);
  if (k < K) {
    const float wk = __ldg(w + k);
    float s = 0.f;
    for (int r = r0; r < r1; ++r) {
      const float dm = __ldg(d + r);
      const float av = a[(long long)r * K + k];
      if (g) g[(long long)r * K + k] = (mask && !(av > 0.f)) ? 0.f : dm * wk;
      s = fmaf(dm, av, s);
    }
    part[(long long)c * (K + 1) + k] = s;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    float sd = 0.f;
    for (int r = r0; r < r1; ++r) sd += __ldg(d + r);
    part[(long long)c * (K + 1) + K] = sd;
  }
}

// the same with 4 columns per thread (128-bit loads / stores, 4 rows in flight); K % 4 == 0, 16-B aligned
constexpr int HB_UNR = 8;
__global__ void __launch_bounds__(128) head_bwd_vec_kernel(int M, int K, const float* d, const float* a,
                                                           const float* w, bool mask, float* g,
                                                           int rows_per_chunk, float* part) {
  B200_PDL_ENTRY();
  const int k = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  const int c = blockIdx.y;
  const int r0 = c * rows_per_chunk, r1 = min(M, r0 + rows_per_chunk);
  if (k < K) {
    const float4 wk = ldg_f4(w + k);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int rb = r0; rb < r1; rb += HB_UNR) {
      float4 av[HB_UNR];
      float dm[HB_UNR];
#pragma unroll
      for (int u = 0; u < HB_UNR; ++u) {
        const int r = rb + u;
        dm[u] = r < r1 ? __ldg(d + r) : 0.f;
        av[u] = r < r1 ? ld_stream_f4(a + (long long)r * K + k) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < HB_UNR; ++u) {
        const int r = rb + u;
        if (r < r1) {
          if (g) {
            float4 o;
            o.x = (mask && !(av[u].x > 0.f)) ? 0.f : dm[u] * wk.x;
            o.y = (mask && !(av[u].y > 0.f)) ? 0.f : dm[u] * wk.y;
            o.z = (mask && !(av[u].z > 0.f)) ? 0.f : dm[u] * wk.z;
            o.w = (mask && !(av[u].w > 0.f)) ? 0.f : dm[u] * wk.w;
            st_f4(g + (long long)r * K + k, o);
          }
          s.x = fmaf(dm[u], av[u].x, s.x); s.y = fmaf(dm[u], av[u].y, s.y);
          s.z = fmaf(dm[u], av[u].z, s.z); s.w = fmaf(dm[u], av[u].w, s.w);
        }
      }
    }
    float* p = part + (long long)c * (K + 1) + k;   // row stride K + 1: not 16-B aligned
    p[0] = s.x; p[1] = s.y; p[2] = s.z; p[3] = s.w;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    float sd = 0.f;
    for (int r = r0; r < r1; ++r) sd += __ldg(d + r);
    part[(long long)c * (K + 1) + K] = sd;
  }
}

int head_layer_bwd(int M, int K, const float* d, const float* a, const float* w, bool mask, float* g,
                   float* gw_gb, DevBuf& scratch, cudaStream_t st) {
  if (M <= 0) return B200REC_OK;
  const bool vec = K % 4 == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(w) |
                                   reinterpret_cast<uintptr_t>(g)) & 15) == 0;
  int chunks = M / (vec ? 16 : 32);   // 4 columns per thread: more row chunks keep the thread count up
  if (chunks > COLSUM_CHUNKS) chunks = COLSUM_CHUNKS;
  if (chunks < 1) chunks = 1;
  const int rows_per_chunk = cdiv(M, chunks);
  chunks = cdiv(M, rows_per_chunk);
  B200_TRY(scratch.reserve((size_t)chunks * (K + 1) * sizeof(float)));
  float* part = scratch.as<float>();
  if (vec) {
    dim3 g1(cdiv(K / 4, 128), chunks);
    B200_LAUNCH(head_bwd_vec_kernel, g1, 128, 0, st, M, K, d, a, w, mask, g, rows_per_chunk, part);
  } else {
    dim3 g1(cdiv(K, 128), chunks);
    B200_LAUNCH(head_bwd_kernel, g1, 128, 0, st, M, K, d, a, w, mask, g, rows_per_chunk, part);
  }
  B200_LAUNCH(colsum_stage2, cdiv((long long)(K + 1) * 32, 256), 256, 0, st, K + 1, chunks, part, 1.0f, false,
              gw_gb);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

// deterministic sum of n floats
__global__ void reduce_stage1(long long n, const float* x, float* part) {
  B200_PDL_ENTRY();
  __shared__ float sh[8];
  float s = 0.f;
  const long long per = (n + gridDim.x - 1) / gridDim.x;
  const long long b0 = blockIdx.x * per, b1 = min(n, b0 + per);
  for (long long i = b0 + threadIdx.x; i < b1; i += blockDim.x) s += x[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) t += sh[i];
    part[blockIdx.x] = t;
  }
}
__global__ void reduce_stage2(int nparts, const float* part, float scale, float* out) {
  B200_PDL_ENTRY();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < nparts; ++i) t += part[i];
    out[0] = t * scale;
  }
}
int reduce_sum(long long n, const float* x, float scale, float* out, DevBuf& scratch,
               cudaStream_t st) {
  const int parts = 64;
  B200_TRY(scratch.reserve(parts * sizeof(float)));
  B200_LAUNCH(reduce_stage1, parts, 256, 0, st, n, x, scratch.as<float>());
  B200_LAUNCH(reduce_stage2, 1, 32, 0, st, parts, scratch.as<float>(), scale, out);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

// h = relu(a + c0)  (c0 optional scalar)
__global__ void add_bias_relu_kernel(long long n, const float* a, const float* c0, float* h) {
  B200_PDL_ENTRY();
  const float c = c0 ? __ldg(c0) : 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    h[i] = fmaxf(a[i] + c, 0.f);
}
int add_bias_relu(long long n, const float* a, const float* c0, float* h, cudaStream_t st) {
  if (n <= 0) return B200REC_OK;
  B200_LAUNCH(add_bias_relu_kernel, grid1d(n), 256, 0, st, n, a, c0, h);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}
__global__ void relu_mask_kernel(long long n, const float* g, const float* h, float* out) {
  B200_PDL_ENTRY();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = h[i] > 0.f ? g[i] : 0.f;
}
int relu_mask(long long n, const float* g, const float* h, float* out, cudaStream_t st) {
  if (n <= 0) return B200REC_OK;
  B200_LAUNCH(relu_mask_kernel, grid1d(n), 256, 0, st, n, g, h, out);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

__global__ void axpy_kernel(long long n, const float* x, float* y) {
  B200_PDL_ENTRY();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    y[i] += x[i];
}
int axpy(long long n, const float* x, float* y, cudaStream_t st) {
  if (n <= 0) return B200REC_OK;
  B200_LAUNCH(axpy_kernel, grid1d(n), 256, 0, st, n, x, y);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

int pnn_lp_fwd(int B, int P, int O, const float* ip, const float* wp, const float* prev,
               const float* c0, float* h, cudaStream_t st, int mode) {
  if (mode) return tc_pnn_lp_fwd(B, P, O, ip, wp, prev, c0, h, mode == 1 ? 3 : 1, st);
  // ProductEncoder.scala:97-108: CAddTable(lz, lp) -> CAdd(scalar) -> ReLU
  return gemm_simt(B, O, P, 1, RowMajorOp{ip, P}, RowMajorOp{wp, P}, EpAddBiasRelu2{h, O, prev, c0},
                   st);
}

}  // namespace b200rec
