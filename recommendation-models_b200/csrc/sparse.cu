// HBM-bound sparse kernels: embedding-row gather, fused first-order + FM second-order forward,
// embedding-gradient producer, and the reference's Scatter / Gather / DotProduct2 modules.
//
// Reference (relative to /root/reference/src/main/scala):
//   rec/model/ParRecModel.scala:279-284,300-306   makeWeights / makeEmbeddings   (gather)
//   nn/Scatter.scala:17-59                        first-order segment sum and its backward
//   rec/model/encoder/SecondOrderEncoder.scala:19-34   0.5*mean_k[(sum_f v)^2 - sum_f v^2]
//   rec/util/GradUtil.scala:7-42                  per-nnz gradient write-back
//
// Layout: an embedding row is K contiguous floats (64 B at K=16).  A row is read by
// LPR = K/4 adjacent lanes with one 128-bit load each, so a warp moves 32/LPR rows per load
// instruction and the [B,F,K] activation is written with fully coalesced 512-B warp stores.
#include "kernels.h"

namespace b200rec {

// ------------------------------------------------------------------------------------------------
// Fused gather + first-order + second-order forward.  One warp per sample.
//   GATHER : rows come from table[feats[.]] (and are optionally written to X) instead of emb_in
//   LPR    : lanes per row (K = 4*LPR)
// ------------------------------------------------------------------------------------------------
template <int LPR, bool GATHER>
__global__ void __launch_bounds__(256) fm_fwd_kernel(SparseFwd a) {
  B200_PDL_ENTRY();
  constexpr int RPW = 32 / LPR;  // rows per warp-wide load
  constexpr int UNR = 5;         // independent row loads in flight per lane (39 fields = 5 x 8)
  const int K = 4 * LPR;
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR;
  const int slot = lane / LPR;
  const int warps_per_block = blockDim.x >> 5;
  const int F = a.F;
  for (int b = blockIdx.x * warps_per_block + (threadIdx.x >> 5); b < a.B;
       b += gridDim.x * warps_per_block) {
    float4 S = make_float4(0.f, 0.f, 0.f, 0.f), Q = S;
    float wsum = 0.f;
    const long long base = (long long)b * F;
    for (int f0 = 0; f0 < F; f0 += RPW * UNR) {
      long long id[UNR];
      bool ok[UNR];
      float4 v[UNR];
      float w[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int f = f0 + u * RPW + slot;
        ok[u] = f < F;
        id[u] = 0;
        if (GATHER && ok[u]) {
          long long x = a.feats[base + f];
          if (x < 0 || x >= a.rows) {
            if (a.err) atomicOr(a.err, DEV_BAD_ID);
            x = x < 0 ? -1 : 0;   // negative (e.g. a slot dropped by a full exchange bucket): a ZERO row, never another id's
          }
          id[u] = x;
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int f = f0 + u * RPW + slot;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        w[u] = 0.f;
        if (ok[u]) {
          if (GATHER) {
            if (id[u] >= 0) {
              v[u] = ldg_f4(a.table + id[u] * K + sub * 4);
              if (sub == 0 && a.wtable) w[u] = __ldg(a.wtable + id[u]);
            }
          } else {
            v[u] = ld_stream_f4(a.emb_in + (base + f) * K + sub * 4);
            if (sub == 0 && a.w_in) w[u] = __ldg(a.w_in + base + f);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int f = f0 + u * RPW + slot;
        if (ok[u]) {
          if (GATHER && a.X) st_f4(a.X + (base + f) * K + sub * 4, v[u]);
          if (GATHER && a.w_out && sub == 0) a.w_out[base + f] = w[u];
        }
        S.x += v[u].x; S.y += v[u].y; S.z += v[u].z; S.w += v[u].w;
        // Power(2) then Sum (SecondOrderEncoder.scala:24-26): rounded squares, then added -- no fma
        Q.x = __fadd_rn(Q.x, __fmul_rn(v[u].x, v[u].x)); Q.y = __fadd_rn(Q.y, __fmul_rn(v[u].y, v[u].y));
        Q.z = __fadd_rn(Q.z, __fmul_rn(v[u].z, v[u].z)); Q.w = __fadd_rn(Q.w, __fmul_rn(v[u].w, v[u].w));
        wsum += w[u];
      }
    }
    // combine the RPW row slots (fixed xor tree => deterministic)
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) {
      S.x += __shfl_xor_sync(0xffffffffu, S.x, o); S.y += __shfl_xor_sync(0xffffffffu, S.y, o);
      S.z += __shfl_xor_sync(0xffffffffu, S.z, o); S.w += __shfl_xor_sync(0xffffffffu, S.w, o);
      Q.x += __shfl_xor_sync(0xffffffffu, Q.x, o); Q.y += __shfl_xor_sync(0xffffffffu, Q.y, o);
      Q.z += __shfl_xor_sync(0xffffffffu, Q.z, o); Q.w += __shfl_xor_sync(0xffffffffu, Q.w, o);
      wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
    }
    if (a.S && slot == 0) st_f4(a.S + (long long)b * K + sub * 4, S);
    if (a.second) {
      // Sum -> Power(2), CSubTable: (rounded S^2) - Q, so that F = 1 gives exactly 0 like the reference
      float d = __fsub_rn(__fmul_rn(S.x, S.x), Q.x) + __fsub_rn(__fmul_rn(S.y, S.y), Q.y) +
                __fsub_rn(__fmul_rn(S.z, S.z), Q.z) + __fsub_rn(__fmul_rn(S.w, S.w), Q.w);
#pragma unroll
      for (int o = 1; o < LPR; o <<= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      if (lane == 0) a.second[b] = 0.5f * (d / (float)K);
    }
    if (a.first && lane == 0) a.first[b] = wsum;
  }
}

// Generic K (not 4*2^n): one warp per sample, lane strides over k.  Same outputs.
template <bool GATHER>
__global__ void __launch_bounds__(256) fm_fwd_generic_kernel(SparseFwd a) {
  B200_PDL_ENTRY();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int F = a.F, K = a.K;
  for (int b = blockIdx.x * wpb + (threadIdx.x >> 5); b < a.B; b += gridDim.x * wpb) {
    const long long base = (long long)b * F;
    float d = 0.f;
    for (int k = lane; k < K; k += 32) {
      float s = 0.f, q = 0.f;
      for (int f = 0; f < F; ++f) {
        float v;
        if (GATHER) {
          long long id = a.feats[base + f];
          if (id < 0 || id >= a.rows) { if (a.err) atomicOr(a.err, DEV_BAD_ID); id = 0; }
          v = __ldg(a.table + id * K + k);
          if (a.X) a.X[(base + f) * K + k] = v;
        } else {
          v = a.emb_in[(base + f) * K + k];
        }
        s += v;
        q = __fadd_rn(q, __fmul_rn(v, v));
      }
      if (a.S) a.S[(long long)b * K + k] = s;
      d += __fsub_rn(__fmul_rn(s, s), q);
    }
    d = warp_sum(d);
    float wsum = 0.f;
    for (int f = lane; f < F; f += 32) {
      float w = 0.f;
      if (GATHER) {
        long long id = a.feats[base + f];
        if (id < 0 || id >= a.rows) id = 0;
        if (a.wtable) w = __ldg(a.wtable + id);
        if (a.w_out) a.w_out[base + f] = w;
      } else if (a.w_in) {
        w = a.w_in[base + f];
      }
      wsum += w;
    }
    wsum = warp_sum(wsum);
    if (lane == 0) {
      if (a.second) a.second[b] = 0.5f * (d / (float)K);
      if (a.first) a.first[b] = wsum;
    }
  }
}

static int lpr_for(int K) {
  if (K % 4) return 0;
  int l = K / 4;
  return (l >= 1 && l <= 32 && (l & (l - 1)) == 0) ? l : 0;
}

template <bool GATHER>
static int launch_fm_fwd(const SparseFwd& a, cudaStream_t st) {
  const int threads = 256, wpb = threads / 32;
  int grid = cdiv(a.B, wpb);
  if (grid > 148 * 32) grid = 148 * 32;
  if (grid < 1) grid = 1;
  switch (lpr_for(a.K)) {
    case 1: B200_LAUNCH((fm_fwd_kernel<1, GATHER>), grid, threads, 0, st, a); break;
    case 2: B200_LAUNCH((fm_fwd_kernel<2, GATHER>), grid, threads, 0, st, a); break;
    case 4: B200_LAUNCH((fm_fwd_kernel<4, GATHER>), grid, threads, 0, st, a); break;
    case 8: B200_LAUNCH((fm_fwd_kernel<8, GATHER>), grid, threads, 0, st, a); break;
    case 16: B200_LAUNCH((fm_fwd_kernel<16, GATHER>), grid, threads, 0, st, a); break;
    case 32: B200_LAUNCH((fm_fwd_kernel<32, GATHER>), grid, threads, 0, st, a); break;
    default: B200_LAUNCH((fm_fwd_generic_kernel<GATHER>), grid, threads, 0, st, a); break;
  }
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

int sparse_fwd(const SparseFwd& a, cudaStream_t st) {
  if (a.B <= 0) return B200REC_OK;
  return a.feats ? launch_fm_fwd<true>(a, st) : launch_fm_fwd<false>(a, st);
}

// ------------------------------------------------------------------------------------------------
// Embedding-gradient producer:  dE[i,:] = (dlogit_b / K)(S_b - v_i) [second order]  + dX[i,:]
//                               dw[i]   = dlogit[index[i]]              (nn/Scatter.scala:38-59)
// Elementwise over [B*F, K]; may run in place over X.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) emb_grad_kernel(SparseBwd a, long long n_vec, int kv) {
  B200_PDL_ENTRY();
  // one thread per float4 of the [B*F, K] grid (kv = K/4 vectors per row)
  const int F = a.F, K = a.K;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n_vec;
       t += (long long)gridDim.x * blockDim.x) {
    const long long row = t / kv;
    const int sub = (int)(t - row * kv);
    const int b = (int)(row / F);
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.dX) g = ld_stream_f4(a.dX + row * K + sub * 4);
    if (a.S) {
      const float c = __ldg(a.dlogit + b) / (float)K;
      const float4 s = ldg_f4(a.S + (long long)b * K + sub * 4);
      const float4 v = ld_stream_f4(a.X + row * K + sub * 4);
      g.x = fmaf(c, s.x - v.x, g.x); g.y = fmaf(c, s.y - v.y, g.y);
      g.z = fmaf(c, s.z - v.z, g.z); g.w = fmaf(c, s.w - v.w, g.w);
    }
    const long long orow = a.out_slot ? (long long)a.out_slot[row] : row;
    if (orow < 0) continue;   // no slot (full exchange bucket, already flagged): the gradient is dropped
    st_f4(a.dE + orow * K + sub * 4, g);
    if (sub == 0 && a.dw) {
      const int bi = a.index ? a.index[row] : b;
      a.dw[orow] = __ldg(a.dlogit + bi);
    }
  }
}

__global__ void __launch_bounds__(256) emb_grad_generic_kernel(SparseBwd a, long long n) {
  B200_PDL_ENTRY();
  const int F = a.F, K = a.K;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n;
       t += (long long)gridDim.x * blockDim.x) {
    const long long row = t / K;
    const int k = (int)(t - row * K);
    const int b = (int)(row / F);
    float g = a.dX ? a.dX[t] : 0.f;
    if (a.S) g = fmaf(a.dlogit[b] / (float)K, a.S[(long long)b * K + k] - a.X[t], g);
    const long long orow = a.out_slot ? (long long)a.out_slot[row] : row;
    if (orow < 0) continue;
    a.dE[orow * K + k] = g;
    if (k == 0 && a.dw) a.dw[orow] = a.dlogit[a.index ? a.index[row] : b];
  }
}

__global__ void dw_only_kernel(long long n, int F, const int* index, const float* dlogit,
                               const int* out_slot, float* dw) {
  B200_PDL_ENTRY();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const long long o = out_slot ? (long long)out_slot[i] : i;
    if (o >= 0) dw[o] = dlogit[index ? index[i] : (int)(i / F)];
  }
}

int sparse_bwd(const SparseBwd& a, cudaStream_t st) {
  const long long rows = (long long)a.B * a.F;
  if (rows <= 0) return B200REC_OK;
  if (!a.dE) {  // LR: only the first-order gradient
    B200_LAUNCH(dw_only_kernel, cdiv(rows, 256), 256, 0, st, rows, a.F, a.index, a.dlogit, a.out_slot, a.dw);
    B200_CHECK_LAUNCH();
    return B200REC_OK;
  }
  if (a.K % 4 == 0) {
    const int kv = a.K / 4;
    const long long n_vec = rows * kv;
    int grid = cdiv(n_vec, 256);
    if (grid > 148 * 16) grid = 148 * 16;
    B200_LAUNCH(emb_grad_kernel, grid, 256, 0, st, a, n_vec, kv);
  } else {
    const long long n = rows * a.K;
    int grid = cdiv(n, 256);
    if (grid > 148 * 16) grid = 148 * 16;
    B200_LAUNCH(emb_grad_generic_kernel, grid, 256, 0, st, a, n);
  }
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

// ------------------------------------------------------------------------------------------------
// Plain lookup (makeEmbeddings / makeWeights): bit-exact copies.
// ------------------------------------------------------------------------------------------------
template <bool PAD>
__global__ void __launch_bounds__(256) lookup_kernel(long long rows, int K, long long n,
                                                     const int* feats, const float* table,
                                                     const float* wtable, float* emb_out,
                                                     float* w_out, int* err) {
  B200_PDL_ENTRY();
  const int kv = K / 4;  // caller guarantees K % 4 == 0 on this path
  const long long n_vec = n * kv;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n_vec;
       t += (long long)gridDim.x * blockDim.x) {
    const long long i = t / kv;
    const int sub = (int)(t - i * kv);
    long long id = feats[i];
    if (PAD && id < 0) {  // exchange padding: zero row
      if (emb_out) st_f4(emb_out + i * K + sub * 4, make_float4(0.f, 0.f, 0.f, 0.f));
      if (sub == 0 && w_out) w_out[i] = 0.f;
      continue;
    }
    if (id < 0 || id >= rows) {
      if (err) atomicOr(err, DEV_BAD_ID);
      id = 0;
    }
    if (emb_out) st_f4(emb_out + i * K + sub * 4, ldg_f4(table + id * K + sub * 4));
    if (sub == 0 && w_out) w_out[i] = __ldg(wtable + id);
  }
}

template <bool PAD>
__global__ void lookup_generic_kernel(long long rows, int K, long long n, const int* feats,
                                      const float* table, const float* wtable, float* emb_out,
                                      float* w_out, int* err) {
  B200_PDL_ENTRY();
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n * K;
       t += (long long)gridDim.x * blockDim.x) {
    const long long i = t / K;
    const int k = (int)(t - i * K);
    long long id = feats[i];
    if (PAD && id < 0) {
      if (emb_out) emb_out[t] = 0.f;
      if (k == 0 && w_out) w_out[i] = 0.f;
      continue;
    }
    if (id < 0 || id >= rows) {
      if (err) atomicOr(err, DEV_BAD_ID);
      id = 0;
    }
    if (emb_out) emb_out[t] = table[id * K + k];
    if (k == 0 && w_out) w_out[i] = wtable[id];
  }
}

template <bool PAD>
static int lookup_rows_impl(long long rows, int K, long long n, const int* feats, const float* table,
                            const float* wtable, float* emb_out, float* w_out, int* err,
                            cudaStream_t st) {
  if (n <= 0) return B200REC_OK;
  if (K % 4 == 0) {
    int grid = cdiv(n * (K / 4), 256);
    if (grid > 148 * 16) grid = 148 * 16;
    B200_LAUNCH(lookup_kernel<PAD>, grid, 256, 0, st, rows, K, n, feats, table, wtable, emb_out,
                w_out, err);
  } else {
    int grid = cdiv(n * K, 256);
    if (grid > 148 * 16) grid = 148 * 16;
    B200_LAUNCH(lookup_generic_kernel<PAD>, grid, 256, 0, st, rows, K, n, feats, table, wtable,
                emb_out, w_out, err);
  }
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}
int lookup_rows(long long rows, int K, long long n, const int* feats, const float* table,
                const float* wtable, float* emb_out, float* w_out, int* err, cudaStream_t st) {
  return lookup_rows_impl<false>(rows, K, n, feats, table, wtable, emb_out, w_out, err, st);
}
int lookup_rows_padded(long long rows, int K, long long n, const int* feats, const float* table,
                       const float* wtable, float* emb_out, float* w_out, int* err,
                       cudaStream_t st) {
  return lookup_rows_impl<true>(rows, K, n, feats, table, wtable, emb_out, w_out, err, st);
}

// ------------------------------------------------------------------------------------------------
// nn/Scatter.scala generic: out[index[i], c] += in[i, c], i ascending (bit-exact order).
// One thread per (output row b, column c) walks the inputs whose index == b in order.  With the
// non-decreasing index the parser emits, the run of b is found by binary search; an unsorted
// index falls back to a full in-order scan of the n inputs per output row (correct, slow, rare).
// ------------------------------------------------------------------------------------------------
__global__ void index_check_kernel(long long n, int B, const int* index, int* flags, int* err) {
  B200_PDL_ENTRY();
  // flags[0] |= 1 when index is not non-decreasing
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = index[i];
    if (x < 0 || x >= B) atomicOr(err, DEV_BAD_INDEX);
    if (i > 0 && index[i - 1] > x) atomicOr(flags, 1);
  }
}

__global__ void scatter_fwd_kernel(int B, int n_out, long long n, const float* in,
                                   const int* index, float* out, const int* flags) {
  B200_PDL_ENTRY();
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)B * n_out) return;
  const int b = (int)(t / n_out), c = (int)(t % n_out);
  float acc = 0.f;
  if (*flags == 0) {
    long long lo = 0, hi = n;  // first i with index[i] >= b
    while (lo < hi) {
      long long mid = (lo + hi) >> 1;
      if (index[mid] < b) lo = mid + 1; else hi = mid;
    }
    for (long long i = lo; i < n && index[i] == b; ++i) acc += in[i * n_out + c];
  } else {
    for (long long i = 0; i < n; ++i)
      if (index[i] == b) acc += in[i * n_out + c];
  }
  out[t] = acc;
}

__global__ void scatter_bwd_kernel(int B, int n_out, long long n, const int* index,
                                   const float* gout, float* gin, int* err) {
  B200_PDL_ENTRY();
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n * n_out;
       t += (long long)gridDim.x * blockDim.x) {
    const long long i = t / n_out;
    const int c = (int)(t - i * n_out);
    int b = index[i];
    if (b < 0 || b >= B) { atomicOr(err, DEV_BAD_INDEX); b = 0; }
    gin[t] = gout[(long long)b * n_out + c];
  }
}

int scatter_fwd(int B, int n_out, long long n, const float* in, const int* index, float* out,
                int* err, cudaStream_t st) {
  // err[0] = DevErr word, err[1] = sortedness flag (both zeroed by the caller)
  if (n > 0) {
    int grid = cdiv(n, 256);
    if (grid > 148 * 8) grid = 148 * 8;
    B200_LAUNCH(index_check_kernel, grid, 256, 0, st, n, B, index, err + 1, err);
  }
  if ((long long)B * n_out > 0)
    B200_LAUNCH(scatter_fwd_kernel, cdiv((long long)B * n_out, 128), 128, 0, st, B, n_out, n, in,
                index, out, err + 1);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

int scatter_bwd(int B, int n_out, long long n, const int* index, const float* gout, float* gin,
                int* err, cudaStream_t st) {
  if (n * n_out <= 0) return B200REC_OK;
  int grid = cdiv(n * n_out, 256);
  if (grid > 148 * 16) grid = 148 * 16;
  B200_LAUNCH(scatter_bwd_kernel, grid, 256, 0, st, B, n_out, n, index, gout, gin, err);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

// ------------------------------------------------------------------------------------------------
// nn/Gather.scala (PNN pair gather) and nn/DotProduct2.scala as stand-alone modules.  The PNN
// model itself never materialises the pair copies (pnn.cu); these exist for the module ABI.
// ------------------------------------------------------------------------------------------------
__global__ void pair_gather_fwd_kernel(int B, int F, int P, int K, const float* in,
                                       const int* rows, const int* cols, float* ro, float* co) {
  B200_PDL_ENTRY();
  const long long n = (long long)B * P * K;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n;
       t += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(t % K);
    const long long bp = t / K;
    const int p = (int)(bp % P);
    const long long b = bp / P;
    ro[t] = in[(b * F + rows[p]) * K + k];
    co[t] = in[(b * F + cols[p]) * K + k];
  }
}

// one thread per (b, field, k): accumulate pairs in ascending p, row-role before col-role at equal p
__global__ void pair_gather_bwd_kernel(int B, int F, int P, int K, const int* rows,
                                       const int* cols, const float* gr, const float* gc,
                                       float* gin) {
  B200_PDL_ENTRY();
  const long long n = (long long)B * F * K;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n;
       t += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(t % K);
    const long long bf = t / K;
    const int f = (int)(bf % F);
    const long long b = bf / F;
    float acc = 0.f;
    for (int p = 0; p < P; ++p) {
      if (rows[p] == f) acc += gr[(b * P + p) * K + k];
      if (cols[p] == f) acc += gc[(b * P + p) * K + k];
    }
    gin[t] = acc;
  }
}

__global__ void dot2_fwd_kernel(long long n, int K, const float* a, const float* b, float* out) {
  B200_PDL_ENTRY();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < K; ++k)  // cmul then sum (DotProduct2.scala:24-25): no fma
      acc = __fadd_rn(acc, __fmul_rn(a[i * K + k], b[i * K + k]));
    out[i] = acc;
  }
}

__global__ void dot2_bwd_kernel(long long n, int K, const float* a, const float* b,
                                const float* go, float* ga, float* gb) {
  B200_PDL_ENTRY();
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n * K;
       t += (long long)gridDim.x * blockDim.x) {
    const float g = go[t / K];
    ga[t] = __fmul_rn(b[t], g);
    gb[t] = __fmul_rn(a[t], g);
  }
}

static int grid_for(long long n) {
  int g = cdiv(n, 256);
  if (g > 148 * 16) g = 148 * 16;
  return g < 1 ? 1 : g;
}

int pair_gather_fwd(int B, int F, int P, int K, const float* in, const int* rows, const int* cols,
                    float* ro, float* co, cudaStream_t st) {
  B200_LAUNCH(pair_gather_fwd_kernel, grid_for((long long)B * P * K), 256, 0, st, B, F, P, K, in,
              rows, cols, ro, co);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}
int pair_gather_bwd(int B, int F, int P, int K, const int* rows, const int* cols, const float* gr,
                    const float* gc, float* gin, cudaStream_t st) {
  B200_LAUNCH(pair_gather_bwd_kernel, grid_for((long long)B * F * K), 256, 0, st, B, F, P, K, rows,
              cols, gr, gc, gin);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}
int dot2_fwd(long long n, int K, const float* a, const float* b, float* out, cudaStream_t st) {
  B200_LAUNCH(dot2_fwd_kernel, grid_for(n), 256, 0, st, n, K, a, b, out);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}
int dot2_bwd(long long n, int K, const float* a, const float* b, const float* go, float* ga,
             float* gb, cudaStream_t st) {
  B200_LAUNCH(dot2_bwd_kernel, grid_for(n * K), 256, 0, st, n, K, a, b, go, ga, gb);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

}  // namespace b200rec
