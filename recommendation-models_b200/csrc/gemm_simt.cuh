// fp32 SIMT GEMM with functor operands:  C(m,n) = sum_k A(m,k) * B(n,k).
//
// This is the exact-fp32 (FFMA) contraction used wherever a tcgen05 path does not exist yet.  The
// operands are *functors*, so the CIN outer product Z[r,(i,j)] = x0[r,i]*x[r,j]
// (rec/model/xdeepfm/CINEncoder.scala:150-157, MM(transB) + Linear) is generated while the tile is
// loaded and never written to memory (it would be R x F*H floats = 4.1 GB at H=200).
//
// Tile 128x128x8, 256 threads, 8x8 outputs per thread as 2x2 blocks of 4x4 (conflict-free 128-bit
// shared loads), register double buffering, optional split-K over blockIdx.z with a fixed-order
// second pass (no float atomics).
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace b200rec {

constexpr int GBM = 128, GBN = 128, GBK = 8, GPAD = 4;

// ---- operand functors: value(row, k); kMajor = consecutive k are contiguous in memory ----------
struct RowMajorOp {  // v(r,k) = p[r*ld + k]
  const float* p; long long ld;
  static constexpr bool kMajor = true;
  __device__ __forceinline__ float operator()(int r, int k) const { return __ldg(p + r * ld + k); }
};
struct ColMajorOp {  // v(r,k) = p[k*ld + r]
  const float* p; long long ld;
  static constexpr bool kMajor = false;
  __device__ __forceinline__ float operator()(int r, int k) const { return __ldg(p + k * ld + r); }
};
// Z[r, k=(i,j)] = x0[r,i] * x[r,j],  i = k / H, j = k % H      (k-major: j runs fastest)
struct CinZOp {
  const float* x0; const float* x; int F, H;
  static constexpr bool kMajor = true;
  __device__ __forceinline__ float operator()(int r, int k) const {
    const int i = k / H, j = k - i * H;
    return __ldg(x0 + (long long)r * F + i) * __ldg(x + (long long)r * H + j);
  }
};
// Z^T as a B operand for dW: v(n=(i,j), k=r) = x0[r,i] * x[r,j]    (n-major)
struct CinZtOp {
  const float* x0; const float* x; int F, H;
  static constexpr bool kMajor = false;
  __device__ __forceinline__ float operator()(int n, int r) const {
    const int i = n / H, j = n - i * H;
    return __ldg(x0 + (long long)r * F + i) * __ldg(x + (long long)r * H + j);
  }
};

// ---- epilogues: ep(m, n, acc, z) ---------------------------------------------------------------
struct EpBiasAct {  // y = act(acc + bias[n])
  float* y; long long ld; const float* bias; bool relu;
  __device__ __forceinline__ void operator()(int m, int n, float acc, int) const {
    float v = acc + (bias ? __ldg(bias + n) : 0.f);
    if (relu) v = fmaxf(v, 0.f);
    y[m * ld + n] = v;
  }
};
struct EpMaskAcc {  // g = acc * (mask > 0), optionally accumulated
  float* g; long long ld; const float* mask; long long ldm; bool accumulate;
  __device__ __forceinline__ void operator()(int m, int n, float acc, int) const {
    if (mask && !(__ldg(mask + m * ldm + n) > 0.f)) acc = 0.f;
    if (accumulate) acc += g[m * ld + n];
    g[m * ld + n] = acc;
  }
};
struct EpPartial {  // split-K partial: ws[z][m][n]
  float* ws; long long MN; long long ld;
  __device__ __forceinline__ void operator()(int m, int n, float acc, int z) const {
    ws[z * MN + m * ld + n] = acc;
  }
};
struct EpAddBiasRelu2 {  // PNN: h = relu(prev + acc + c0)
  float* h; long long ld; const float* prev; const float* c0;
  __device__ __forceinline__ void operator()(int m, int n, float acc, int) const {
    h[m * ld + n] = fmaxf((prev[m * ld + n] + acc) + __ldg(c0), 0.f);
  }
};

template <class AOp, class BOp, class Ep>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(int M, int N, int K, int k_chunk, AOp aop, BOp bop, Ep ep) {
  B200_PDL_ENTRY();
  __shared__ __align__(16) float As[2][GBK][GBM + GPAD];
  __shared__ __align__(16) float Bs[2][GBK][GBN + GPAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
  const int z = blockIdx.z;
  const int k_begin = z * k_chunk;
  const int k_end = min(K, k_begin + k_chunk);

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float ra[4], rb[4];
  auto load_tile = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int r, kk;
      if (AOp::kMajor) { kk = tid & 7; r = (tid >> 3) + 32 * i; }
      else             { r = tid & 127; kk = (tid >> 7) + 2 * i; }
      const int gm = m0 + r, gk = k0 + kk;
      ra[i] = (gm < M && gk < k_end) ? aop(gm, gk) : 0.f;
      if (BOp::kMajor) { kk = tid & 7; r = (tid >> 3) + 32 * i; }
      else             { r = tid & 127; kk = (tid >> 7) + 2 * i; }
      const int gn = n0 + r, gk2 = k0 + kk;
      rb[i] = (gn < N && gk2 < k_end) ? bop(gn, gk2) : 0.f;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int r, kk;
      if (AOp::kMajor) { kk = tid & 7; r = (tid >> 3) + 32 * i; }
      else             { r = tid & 127; kk = (tid >> 7) + 2 * i; }
      As[buf][kk][r] = ra[i];
      if (BOp::kMajor) { kk = tid & 7; r = (tid >> 3) + 32 * i; }
      else             { r = tid & 127; kk = (tid >> 7) + 2 * i; }
      Bs[buf][kk][r] = rb[i];
    }
  };

  int buf = 0;
  if (k_begin < k_end) {
    load_tile(k_begin);
    store_tile(0);
  }
  __syncthreads();
  for (int k0 = k_begin; k0 < k_end; k0 += GBK) {
    const bool more = k0 + GBK < k_end;
    if (more) load_tile(k0 + GBK);
#pragma unroll
    for (int kk = 0; kk < GBK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      store_tile(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n < N) ep(m, n, acc[i][j], z);
    }
  }
}

// out (+)= scale * sum_z ws[z]   (z ascending: fixed order)            -- dense.cu
int splitk_reduce(const float* ws, int splits, long long MN, float scale, bool accumulate,
                  float* out, cudaStream_t st);
int splitk_reduce_wb(const float* ws, int splits, int N, int K, int ldw, float scale, bool accumulate,
                     float* gw, float* gb, cudaStream_t st);
// colsum / COLSUM_CHUNKS: kernels.h
// the split count gemm_simt really uses for (K, splits)
static inline int real_splits(int K, int splits) {
  if (splits < 1) splits = 1;
  int k_chunk = (((K + splits - 1) / splits + 7) / 8) * 8;
  if (k_chunk < 8) k_chunk = 8;
  return K > 0 ? (K + k_chunk - 1) / k_chunk : 1;
}

template <class AOp, class BOp, class Ep>
static int gemm_simt(int M, int N, int K, int splits, AOp a, BOp b, Ep ep, cudaStream_t st) {
  if (M <= 0 || N <= 0) return B200REC_OK;
  if (splits < 1) splits = 1;
  int k_chunk = ((cdiv(K, splits) + GBK - 1) / GBK) * GBK;
  if (k_chunk < GBK) k_chunk = GBK;
  splits = K > 0 ? cdiv(K, k_chunk) : 1;
  dim3 grid(cdiv(N, GBN), cdiv(M, GBM), splits);
  B200_LAUNCH((gemm_simt_kernel<AOp, BOp, Ep>), grid, 256, 0, st, M, N, K, k_chunk, a, b, ep);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

// number of K splits that fills the machine for an (M,N) output
static inline int pick_splits(int M, int N, int K) {
  const long long tiles = (long long)cdiv(M, GBM) * cdiv(N, GBN);
  if (tiles >= 148 * 2) return 1;
  long long s = (148LL * 2 + tiles - 1) / tiles;
  const long long max_s = K / (GBK * 8) > 0 ? K / (GBK * 8) : 1;
  if (s > max_s) s = max_s;
  if (s > 64) s = 64;
  return (int)(s < 1 ? 1 : s);
}

}  // namespace b200rec
