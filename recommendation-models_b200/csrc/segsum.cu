// Sorted-index segmented scatter-add (deterministic, reference order).
//
// Reference: rec/model/ParRecModel.scala:316-328 (makeEmbeddingGrad) and :293-298
// (makeWeightsGrad): for every non-zero i = 0..N-1 IN ORDER, grads(j).addTo(feats(i), buf(i*K+j)).
// So the gradient of a distinct id is the fp32 sum of its rows in increasing i.
//
// Here: a stable LSD radix sort of (id, i) pairs groups equal ids while keeping i ascending;
// segment heads give the distinct ids (ascending); each segment is then summed strictly in
// order.  Loads are issued in parallel, only the fp32 adds are sequential, so the result is
// bit-identical to the reference's hash-map accumulation regardless of grid size -- there are
// no float atomics anywhere.
//   * short segments (<= LONG_T rows): LPR = K/4 lanes per segment, 128-bit loads
//   * long segments  (hot ids of the power law): one warp per segment; 32/LPR rows are loaded
//     per instruction, then folded into the accumulator in order through shuffles.
// The sort half depends only on the ids, so the driver runs it on a side stream while the
// dense math runs (model.cu).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "kernels.h"

namespace b200rec {

static constexpr int LONG_T = 32;   // (8 was measured slower: 3 000 one-warp segments instead of 700)

int SegSumWorkspace::reserve(long long n) {
  if (n <= cap_n) return B200REC_OK;
  const size_t ni = (size_t)n + 8;
  B200_TRY(keys_b.reserve(ni * 4));
  B200_TRY(vals_a.reserve(ni * 4));
  B200_TRY(vals_b.reserve(ni * 4));
  B200_TRY(seg_start.reserve(ni * 4));
  B200_TRY(long_list.reserve(ni / LONG_T * 4 + 64));
  B200_TRY(counters.reserve(64));
  size_t sort_bytes = 0, scan_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const unsigned*)nullptr, (unsigned*)nullptr,
                                  (const unsigned*)nullptr, (unsigned*)nullptr, (int)n, 0, 32);
  cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, (const int*)nullptr, (int*)nullptr, (int)n);
  B200_TRY(cub_tmp.reserve((sort_bytes > scan_bytes ? sort_bytes : scan_bytes) * 2 + 1024));
  cap_n = n;
  return B200REC_OK;
}

void SegSumWorkspace::release() {
  keys_a.release(); keys_b.release(); vals_a.release(); vals_b.release(); cub_tmp.release();
  seg_start.release(); long_list.release(); counters.release();
  cap_n = 0;
}

__global__ void iota_kernel(long long n, unsigned* v, int* counters) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) v[i] = (unsigned)i;
  if (i == 0) { counters[0] = 0; counters[1] = 0; }
}

__global__ void head_flag_kernel(long long n, const unsigned* keys, int* flags) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// seg_idx[i] = 1-based segment number of sorted position i (inclusive scan of the head flags)
__global__ void seg_heads_kernel(long long n, const unsigned* keys, const int* seg_idx,
                                 int* seg_start, int* unique, int* n_unique, bool drop_pad) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = seg_idx[i];
  if (i == 0 || keys[i] != keys[i - 1]) {
    seg_start[s - 1] = (int)i;
    unique[s - 1] = (int)keys[i];
  }
  if (i == n - 1) {
    // exchange padding (key 0xffffffff) sorts last: its segment is not a feature id
    *n_unique = (drop_pad && keys[i] == 0xffffffffu) ? s - 1 : s;
    seg_start[s] = (int)n;
  }
}

// hot ids (segments longer than LONG_T) -> long_list, still in the sort half (off the critical path)
__global__ void seg_long_kernel(const int* n_unique, const int* seg_start, int* long_list, int* long_count) {
  const int U = *n_unique;
  for (long long seg = blockIdx.x * (long long)blockDim.x + threadIdx.x; seg < U;
       seg += (long long)gridDim.x * blockDim.x)
    if (seg_start[seg + 1] - seg_start[seg] > LONG_T) long_list[atomicAdd(long_count, 1)] = (int)seg;
}

int segsum_sort(SegSumWorkspace& ws, const SegSum& a, cudaStream_t st) {
  ProfTag tag("scatter_sort");
  B200_TRY(ws.reserve(a.n));
  int* counters = ws.counters.as<int>();
  if (a.n <= 0) {
    B200_CUDA(cudaMemsetAsync(a.n_unique, 0, sizeof(int), st));
    return B200REC_OK;
  }
  const int n = (int)a.n;
  unsigned* vals_in = ws.vals_a.as<unsigned>();
  unsigned* perm = ws.vals_b.as<unsigned>();
  unsigned* keys_sorted = ws.keys_b.as<unsigned>();
  B200_LAUNCH(iota_kernel, cdiv(n, 256), 256, 0, st, (long long)n, vals_in, counters);
  size_t tmp = ws.cub_tmp.cap;
  int bits = a.key_bits < 1 ? 1 : (a.key_bits > 32 ? 32 : a.key_bits);
  if (tl_prof) tl_prof->begin("cub::DeviceRadixSort::SortPairs", st);
  B200_CUDA(cub::DeviceRadixSort::SortPairs(ws.cub_tmp.p, tmp, (const unsigned*)a.feats,
                                            keys_sorted, vals_in, perm, n, 0, bits, st));
  if (tl_prof) tl_prof->end(st);
  g_launches.fetch_add(2 + (bits + 7) / 8, std::memory_order_relaxed);  // cub's own kernels
  // head flags -> inclusive scan -> segment table.  vals_a (the iota) is free again: reuse it.
  int* seg_idx = ws.vals_a.as<int>();
  B200_LAUNCH(head_flag_kernel, cdiv(n, 256), 256, 0, st, (long long)n, keys_sorted, seg_idx);
  tmp = ws.cub_tmp.cap;
  if (tl_prof) tl_prof->begin("cub::DeviceScan::InclusiveSum", st);
  B200_CUDA(cub::DeviceScan::InclusiveSum(ws.cub_tmp.p, tmp, seg_idx, seg_idx, n, st));
  if (tl_prof) tl_prof->end(st);
  g_launches.fetch_add(2, std::memory_order_relaxed);
  B200_LAUNCH(seg_heads_kernel, cdiv(n, 256), 256, 0, st, (long long)n, keys_sorted, seg_idx,
              ws.seg_start.as<int>(), a.unique, a.n_unique, a.drop_pad);
  {
    int grid = cdiv(n, 256);
    if (grid > 148 * 8) grid = 148 * 8;
    B200_LAUNCH(seg_long_kernel, grid, 256, 0, st, a.n_unique, ws.seg_start.as<int>(), ws.long_list.as<int>(),
                counters);
  }
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

// inv[i] = index of the distinct id of non-zero i (valid after segsum_sort on the same workspace)
__global__ void seg_inverse_kernel(long long n, const unsigned* perm, const int* seg_idx, int* inv) {
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p < n) inv[perm[p]] = seg_idx[p] - 1;
}
int segsum_inverse(SegSumWorkspace& ws, long long n, int* inv, cudaStream_t st) {
  if (n <= 0) return B200REC_OK;
  B200_LAUNCH(seg_inverse_kernel, cdiv(n, 256), 256, 0, st, n, ws.vals_b.as<unsigned>(), ws.vals_a.as<int>(), inv);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

// ---- in-order segment sums --------------------------------------------------------------------
template <int LPR>
__device__ __forceinline__ void segsum_short_role(int block, int n_blocks, const int* n_unique,
                                                  const int* seg_start, const unsigned* perm,
                                                  const float* dE, const float* dw, float* G, float* gw) {
  constexpr int K = 4 * LPR;
  constexpr int UNR = 4;
  const int U = *n_unique;
  const long long tid = block * (long long)blockDim.x + threadIdx.x;
  const int sub = (int)(tid % LPR);
  const long long n_groups = ((long long)n_blocks * blockDim.x) / LPR;
  // SB segments per iteration: their (start, length), first position and first row are loaded as
  // three batches of independent loads (the chain start -> perm -> row is paid once per batch, and
  // most segments have one row); the remaining rows of a segment follow in order.
  constexpr int SB = 4;
  for (long long seg0 = tid / LPR; seg0 < U; seg0 += n_groups * SB) {
    int start[SB], len[SB];
#pragma unroll
    for (int i = 0; i < SB; ++i) {
      const long long seg = seg0 + i * n_groups;
      start[i] = 0; len[i] = 0;
      if (seg < U) {
        start[i] = seg_start[seg];
        len[i] = seg_start[seg + 1] - start[i];
        if (len[i] > LONG_T) len[i] = 0;   // a hot id: the long role sums it
      }
    }
    unsigned p0[SB];
#pragma unroll
    for (int i = 0; i < SB; ++i) p0[i] = len[i] > 0 ? perm[start[i]] : 0u;
    float4 v0[SB];
    float w0[SB];
#pragma unroll
    for (int i = 0; i < SB; ++i) {
      v0[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      w0[i] = 0.f;
      if (len[i] > 0) {
        if (dE) v0[i] = ldg_f4(dE + (long long)p0[i] * K + sub * 4);
        if (dw && sub == 0) w0[i] = __ldg(dw + p0[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < SB; ++i) {
      if (len[i] <= 0) continue;
      const long long seg = seg0 + i * n_groups;
      float4 acc = make_float4(0.f + v0[i].x, 0.f + v0[i].y, 0.f + v0[i].z, 0.f + v0[i].w);
      float aw = 0.f + w0[i];
      for (int j0 = 1; j0 < len[i]; j0 += UNR) {
        float4 v[UNR];
        float w[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          w[u] = 0.f;
          if (j0 + u < len[i]) {
            const long long p = perm[start[i] + j0 + u];
            if (dE) v[u] = ldg_f4(dE + p * K + sub * 4);
            if (dw && sub == 0) w[u] = __ldg(dw + p);
          }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          if (j0 + u < len[i]) {
            acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
            aw += w[u];
          }
        }
      }
      if (G) st_f4(G + seg * K + sub * 4, acc);
      if (gw && sub == 0) gw[seg] = aw;
    }
  }
}

template <int LPR>
__device__ __forceinline__ void segsum_long_role(int block, int n_blocks, const int* seg_start,
                                                 const unsigned* perm, const float* dE,
                                                 const float* dw, float* G, float* gw,
                                                 const int* long_list, const int* long_count) {
  // One warp per hot id.  Rows are LOADED 32/LPR x UNR at a time (all loads of a batch in flight,
  // the positions of the next batch prefetched meanwhile) and ADDED strictly in non-zero order
  // through shuffles, so the result equals the reference's sequential hash-map accumulation.
  constexpr int K = 4 * LPR;
  constexpr int RPW = 32 / LPR;
  constexpr int UNR = 8;
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR, slot = lane / LPR;
  const int n_long = *long_count;
  const int warp = (int)((block * (long long)blockDim.x + threadIdx.x) >> 5);
  const int n_warps = (int)(((long long)n_blocks * blockDim.x) >> 5);
  for (int wi = warp; wi < n_long; wi += n_warps) {
    const int seg = long_list[wi];
    const int start = seg_start[seg];
    const int len = seg_start[seg + 1] - start;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float aw = 0.f;
    unsigned pnext[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int j = u * RPW + slot;
      pnext[u] = j < len ? perm[start + j] : 0u;
    }
    for (int base = 0; base < len; base += RPW * UNR) {
      float4 v[UNR];
      float w[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int j = base + u * RPW + slot;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        w[u] = 0.f;
        if (j < len) {
          const long long p = pnext[u];
          if (dE) v[u] = ldg_f4(dE + p * K + sub * 4);
          if (dw && sub == 0) w[u] = __ldg(dw + p);
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {  // positions of the next batch: their latency hides under the fold
        const int j = base + RPW * UNR + u * RPW + slot;
        pnext[u] = j < len ? perm[start + j] : 0u;
      }
      // fold the RPW*UNR rows into the accumulator strictly in order (all lanes keep a copy)
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        if (base + u * RPW >= len) break;   // warp-uniform: nothing left in this batch
#pragma unroll
        for (int s = 0; s < RPW; ++s) {
          const int src = s * LPR + sub;
          const float x = __shfl_sync(0xffffffffu, v[u].x, src);
          const float y = __shfl_sync(0xffffffffu, v[u].y, src);
          const float z = __shfl_sync(0xffffffffu, v[u].z, src);
          const float q = __shfl_sync(0xffffffffu, v[u].w, src);
          const float ww = __shfl_sync(0xffffffffu, w[u], s * LPR);
          if (base + u * RPW + s < len) {
            acc.x += x; acc.y += y; acc.z += z; acc.w += q;
            aw += ww;
          }
        }
      }
    }
    if (G && slot == 0) st_f4(G + (long long)seg * K + sub * 4, acc);
    if (gw && lane == 0) gw[seg] = aw;
  }
}

// ONE launch: the first l_blocks blocks take the hot ids (one warp each: they run longest, so they
// start first), the others the short segments.  The hot-id list comes from the sort half.
template <int LPR>
__global__ void __launch_bounds__(256) segsum_kernel(int l_blocks, const int* n_unique, const int* seg_start,
                                                     const unsigned* perm, const float* dE, const float* dw,
                                                     float* G, float* gw, const int* long_list,
                                                     const int* long_count) {
  if ((int)blockIdx.x < l_blocks)
    segsum_long_role<LPR>(blockIdx.x, l_blocks, seg_start, perm, dE, dw, G, gw, long_list, long_count);
  else
    segsum_short_role<LPR>(blockIdx.x - l_blocks, gridDim.x - l_blocks, n_unique, seg_start, perm, dE, dw, G, gw);
}

// any K: one thread per (segment, k), strictly sequential
__global__ void segsum_generic_kernel(const int* n_unique, const int* seg_start,
                                      const unsigned* perm, int K, const float* dE,
                                      const float* dw, float* G, float* gw) {
  const int U = *n_unique;
  const int KK = K + 1;  // column K = the first-order weight gradient
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < (long long)U * KK;
       t += (long long)gridDim.x * blockDim.x) {
    const long long seg = t / KK;
    const int k = (int)(t - seg * KK);
    const int start = seg_start[seg], end = seg_start[seg + 1];
    float acc = 0.f;
    if (k < K) {
      if (!dE) continue;
      for (int j = start; j < end; ++j) acc += dE[(long long)perm[j] * K + k];
      G[seg * K + k] = acc;
    } else {
      if (!dw) continue;
      for (int j = start; j < end; ++j) acc += dw[perm[j]];
      gw[seg] = acc;
    }
  }
}

template <int LPR>
static int launch_segsum(SegSumWorkspace& ws, const SegSum& a, cudaStream_t st) {
  const int* seg_start = ws.seg_start.as<int>();
  const unsigned* perm = ws.vals_b.as<unsigned>();
  int* counters = ws.counters.as<int>();
  int* long_list = ws.long_list.as<int>();
  long long groups = a.n;  // upper bound on the number of segments
  int grid = cdiv(groups * LPR, 256);
  if (grid > 148 * 8) grid = 148 * 8;
  long long max_long = a.n / LONG_T + 1;
  int lgrid = cdiv(max_long * 32, 256);
  if (lgrid > 148) lgrid = 148;
  B200_LAUNCH((segsum_kernel<LPR>), lgrid + grid, 256, 0, st, lgrid, a.n_unique, seg_start, perm, a.dE, a.dw,
              a.G, a.gw, long_list, counters);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

int segsum_reduce(SegSumWorkspace& ws, const SegSum& a, cudaStream_t st) {
  ProfTag tag("scatter_add");
  if (a.n <= 0) return B200REC_OK;
  int lpr = 0;
  if (a.K % 4 == 0) {
    int l = a.K / 4;
    if (l >= 1 && l <= 32 && (l & (l - 1)) == 0) lpr = l;
  }
  switch (lpr) {
    case 1: return launch_segsum<1>(ws, a, st);
    case 2: return launch_segsum<2>(ws, a, st);
    case 4: return launch_segsum<4>(ws, a, st);
    case 8: return launch_segsum<8>(ws, a, st);
    case 16: return launch_segsum<16>(ws, a, st);
    case 32: return launch_segsum<32>(ws, a, st);
    default: {
      int grid = cdiv(a.n * (a.K + 1), 256);
      if (grid > 148 * 8) grid = 148 * 8;
      B200_LAUNCH(segsum_generic_kernel, grid, 256, 0, st, a.n_unique, ws.seg_start.as<int>(),
                  ws.vals_b.as<unsigned>(), a.K, a.dE, a.dw, a.G, a.gw);
      B200_CHECK_LAUNCH();
      return B200REC_OK;
    }
  }
}

// rec/optim/AsyncSGD.scala:10-31 applies w -= lr * g on the PS (textbook SGD; Angel's PSF source
// is third-party, parity unpinned).  Touched rows only.
__global__ void apply_sgd_kernel(int K, long long rows, const int* n_unique, const int* unique, const float* G,
                                 const float* gw, float lr, float* table, float* wtable, int* err) {
  const int U = *n_unique;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < (long long)U * K;
       t += (long long)gridDim.x * blockDim.x) {
    const long long seg = t / K;
    const int k = (int)(t - seg * K);
    const long long id = unique[seg];
    if (id < 0 || id >= rows) {   // never write outside the table (the reference throws on such ids)
      if (err && k == 0) atomicOr(err, DEV_BAD_ID);
      continue;
    }
    if (G) table[id * K + k] -= lr * G[t];
    if (k == 0 && gw && wtable) wtable[id] -= lr * gw[seg];
  }
}

int apply_sgd(int K, long long rows, long long cap, const int* n_unique, const int* unique, const float* G,
              const float* gw, float lr, float* table, float* wtable, int* err, cudaStream_t st) {
  if (cap <= 0) return B200REC_OK;
  int grid = cdiv(cap * K, 256);
  if (grid > 148 * 8) grid = 148 * 8;
  B200_LAUNCH(apply_sgd_kernel, grid, 256, 0, st, K, rows, n_unique, unique, G, gw, lr, table, wtable, err);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

}  // namespace b200rec
