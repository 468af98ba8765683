// Sorted-index segmented scatter-add (deterministic, reference order).
//
// Reference: rec/model/ParRecModel.scala:316-328 (makeEmbeddingGrad) and :293-298
// (makeWeightsGrad): for every non-zero i = 0..N-1 IN ORDER, grads(j).addTo(feats(i), buf(i*K+j)).
// So the gradient of a distinct id is the fp32 sum of its rows in increasing i.
//
// Here: a stable LSD radix sort of (id, i) pairs groups equal ids while keeping i ascending (own kernels, below:
// a batch is 10^5 .. 10^6 keys of 24 - 32 bits, where a library onesweep sort is bound by its tile-to-tile look-back
// chain -- 0.10 ms for 320 k pairs -- while three short launches per 8-bit digit have no serial chain at all);
// segment heads give the distinct ids (ascending); each segment is then summed strictly in
// order.  Loads are issued in parallel, only the fp32 adds are sequential, so the result is
// bit-identical to the reference's hash-map accumulation regardless of grid size -- there are
// no float atomics anywhere.
//   * short segments (<= LONG_T = 4 rows, the bulk of a power-law batch): LPR = K/4 lanes per segment;
//     its (up to) 4 positions and then its 4 rows are loaded as two batches of independent 128-bit
//     loads -- straight-line code, no per-segment loop to diverge on
//   * longer segments (the hot ids): one warp per segment; 32/LPR rows are loaded per instruction, 64
//     rows per batch with the next batch's positions prefetched, then folded into the accumulator in
//     order through shuffles.  Their list is built in the sort half.
// The sort half depends only on the ids, so the driver runs it on a side stream while the
// dense math runs (model.cu).
//
// FUSED gradient producer (resident step): instead of reading a per-nnz gradient dE[N,K] that an
// elementwise kernel wrote a moment earlier, the reduce can compute each row on the fly,
//     dE[p,:] = (dlogit_b / K) (S_b - X[p,:]) + dX[p,:],   dw[p] = dlogit_b,   b = p / F
// (SecondOrderEncoder backward + GradUtil.embeddingGrad / weightsGrad, rec/util/GradUtil.scala:7-42), with
// the same arithmetic as emb_grad_kernel (sparse.cu), so the sums are bit-identical to the two-kernel
// path while the 2 x N x K x 4 bytes of the dE round trip never touch HBM.
#include <cstdlib>

#include "kernels.h"
#include "p2p_dev.cuh"

namespace b200rec {

static constexpr int LONG_T = 4;    // segments longer than this take the warp-per-segment role
static constexpr int BIG_T = 64;    // ... and beyond this the block-per-segment role (shared-memory staging)
static constexpr int BIG_CH = 448;  // rows of a big segment staged per round (448 x 68 B = 30 KB of shared memory)
static constexpr int BIG_MAX_K = 16; // widest row the staging buffer is sized for

int SegSumWorkspace::reserve(long long n) {
  if (n <= cap_n) return B200REC_OK;
  const size_t ni = (size_t)n + 8;
  B200_TRY(keys_b.reserve(ni * 4));
  B200_TRY(vals_a.reserve(ni * 4));
  B200_TRY(vals_b.reserve(ni * 4));
  B200_TRY(seg_start.reserve(ni * 4));
  // [0, ni/LONG_T): segments of LONG_T+1 .. BIG_T rows; [ni/LONG_T, ...): segments of more than BIG_T rows
  B200_TRY(long_list.reserve((ni / LONG_T + ni / BIG_T) * 4 + 128));
  B200_TRY(counters.reserve(64));
  B200_TRY(keys_a.reserve(ni * 4));
  // radix-sort histograms [256 digits][tiles] (+ the per-tile segment-head counts of the segment scan)
  const size_t tiles = (ni + RS_TILE - 1) / RS_TILE;
  B200_TRY(cub_tmp.reserve((2 * RS_BINS * tiles + 2 * RS_BINS + 2 * tiles + 64) * 4));   // up to 512 bins; SEG_TILE = RS_TILE / 2
  cap_n = n;
  return B200REC_OK;
}

void SegSumWorkspace::release() {
  keys_a.release(); keys_b.release(); vals_a.release(); vals_b.release(); cub_tmp.release();
  seg_start.release(); long_list.release(); counters.release();
  cap_n = 0;
}

// ---- stable LSD radix sort of (key, position) pairs, 8 bits per pass, three launches per pass ------------
//   rs_hist   : tile t (2048 consecutive items) counts its digits                    -> hist[digit][tile]
//   rs_rowscan: one block per digit: exclusive prefix over the tiles + the digit's total (rs_scatter scans the 256 totals)
//   rs_scatter: the tile walks its items in 8 rounds of 256 (round, thread) = index order; inside a round the rank
//               of an item among equal digits = match.any inside the warp + per-warp digit counts scanned over the
//               8 warps, so equal keys keep their input order (the pass is stable; nothing is atomic)
// The first pass takes the position as the value (no iota array).
// (every kernel of the sort half is a 256-thread block with few registers: it has to fit NEXT TO a resident GEMM
// CTA -- 320 threads x 168 registers -- or it would wait for an SM between two GEMM launches and stall them)
__device__ __forceinline__ int block_excl_scan_256(int v, int* warp_sums, int* total) {   // blockDim.x == 256
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[w] = inc;
  __syncthreads();
  int wbase = 0, tot = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int x = warp_sums[q];
    wbase += q < w ? x : 0;
    tot += x;
  }
  if (total) *total = tot;
  __syncthreads();
  return wbase + inc - v;
}

// DB = bits per digit: 8 (default), or 9 (B200REC_SORT_DB9=1) where that saves a whole pass -- see segsum_sort
template <int DB>
__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(int n, const unsigned* __restrict__ keys, int shift,
                                                              int* __restrict__ hist, int tiles, int group, int* counters) {
  B200_PDL_ENTRY();
  constexpr int BINS = 1 << DB, PER = BINS / RS_THREADS;   // PER consecutive digits per thread
  __shared__ int cnt[BINS];
  if (counters && blockIdx.x == 0 && threadIdx.x < 2) counters[threadIdx.x] = 0;   // the hot-id list counters
  // a block takes `group` consecutive tiles: one tile per block when the sort is on the critical path (FM / LR),
  // ~20 fat blocks when it runs beside the dense math, where the GEMMs leave 20 of the 148 SMs idle and anything
  // spread over all SMs takes issue slots from their producer warps
  for (int tile = blockIdx.x * group; tile < min(tiles, (blockIdx.x + 1) * group); ++tile) {
#pragma unroll
  for (int j = 0; j < PER; ++j) cnt[threadIdx.x * PER + j] = 0;
  __syncthreads();
  const int base = tile * RS_TILE;
  unsigned d[RS_TILE / RS_THREADS];
#pragma unroll
  for (int j = 0; j < RS_TILE / RS_THREADS; ++j) {      // all loads of the tile in flight
    const int i = base + j * RS_THREADS + threadIdx.x;
    d[j] = i < n ? ((keys[i] >> shift) & (BINS - 1)) : 0xffffu;
  }
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < RS_TILE / RS_THREADS; ++j) {
    // one shared-memory atomic per distinct digit of the warp (the high digits of feature ids take few values)
    const unsigned peers = __match_any_sync(0xffffffffu, d[j]);
    if (d[j] != 0xffffu && (peers & ((1u << lane) - 1u)) == 0) atomicAdd(&cnt[d[j]], __popc(peers));
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < PER; ++j) hist[(threadIdx.x * PER + j) * tiles + tile] = cnt[threadIdx.x * PER + j];
  __syncthreads();
  }
}

__global__ void __launch_bounds__(256) rs_scan_kernel(int* __restrict__ a, int total) {
  B200_PDL_ENTRY();   // one block: exclusive scan in place
  __shared__ int warp_sums[8];
  const int per = (total + 255) / 256;
  const int b0 = threadIdx.x * per, b1 = min(total, b0 + per);
  int s = 0;
  for (int i = b0; i < b1; ++i) s += a[i];
  int run = block_excl_scan_256(s, warp_sums, nullptr);
  for (int i = b0; i < b1; ++i) {
    const int v = a[i];
    a[i] = run;
    run += v;
  }
}

// one block per digit: hist[d][0 .. tiles) becomes its own exclusive prefix, totals[d] the digit's count
__global__ void __launch_bounds__(RS_THREADS) rs_rowscan_kernel(int* __restrict__ hist, int tiles, int* __restrict__ totals) {
  B200_PDL_ENTRY();
  __shared__ int wsum[RS_THREADS / 32];
  __shared__ int carry_s;
  int* row = hist + (size_t)blockIdx.x * tiles;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int c0 = 0; c0 < tiles; c0 += RS_THREADS) {
    const int i = c0 + threadIdx.x;
    const int v = i < tiles ? row[i] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    int wbase = 0;
#pragma unroll
    for (int q = 0; q < RS_THREADS / 32; ++q) wbase += q < w ? wsum[q] : 0;
    const int carry = carry_s;
    if (i < tiles) row[i] = carry + wbase + inc - v;
    __syncthreads();
    if (threadIdx.x == RS_THREADS - 1) carry_s = carry + wbase + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) totals[blockIdx.x] = carry_s;
}

template <int DB>
__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(int n, const unsigned* __restrict__ keys_in,
                                                                 const unsigned* __restrict__ vals_in, int shift,
                                                                 const int* __restrict__ hist, int tiles, int group,
                                                                 const int* __restrict__ totals,
                                                                 unsigned* __restrict__ keys_out,
                                                                 unsigned* __restrict__ vals_out) {
  B200_PDL_ENTRY();
  constexpr int BINS = 1 << DB, PER = BINS / RS_THREADS;   // thread t owns the PER consecutive digits t * PER ...
  __shared__ int base[BINS];                           // next output slot of every digit for this tile
  __shared__ int wcnt[RS_THREADS / 32][BINS];          // per-warp digit counts of the current round
  __shared__ int wsum[RS_THREADS / 32];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  for (int tile = blockIdx.x * group; tile < min(tiles, (blockIdx.x + 1) * group); ++tile) {
  const int tile0 = tile * RS_TILE;
  constexpr int NJ = RS_TILE / RS_THREADS;
  unsigned keyr[NJ], valr[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {                       // the tile's loads fly together, ahead of the ranking rounds
    const int i = tile0 + j * RS_THREADS + t;
    keyr[j] = i < n ? keys_in[i] : 0u;
    valr[j] = (i < n && vals_in) ? vals_in[i] : (unsigned)i;
  }
  {  // first slot of a digit for this tile = (all smaller digits) + (the same digit in the tiles before this one)
    int totj[PER], tot = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) { totj[j] = totals[t * PER + j]; tot += totj[j]; }
    int inc = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    int wbase = 0;
#pragma unroll
    for (int q = 0; q < RS_THREADS / 32; ++q) wbase += q < w ? wsum[q] : 0;
    int run = wbase + inc - tot;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      base[t * PER + j] = run + hist[(t * PER + j) * tiles + tile];
      run += totj[j];
    }
  }
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int i = tile0 + j * RS_THREADS + t;
    if (tile0 + j * RS_THREADS >= n) break;            // block-uniform
    const bool valid = i < n;
    const unsigned key = keyr[j], val = valr[j];
    const unsigned d = valid ? ((key >> shift) & (BINS - 1)) : 0xffffu;
#pragma unroll
    for (int q = 0; q < RS_THREADS / 32; ++q)
#pragma unroll
      for (int jj = 0; jj < PER; ++jj) wcnt[q][t * PER + jj] = 0;
    __syncthreads();
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    if (valid && rank == 0) wcnt[w][d] = __popc(peers);
    __syncthreads();
    int acc[PER];                                      // thread t scans its digits over the warps
#pragma unroll
    for (int jj = 0; jj < PER; ++jj) {
      acc[jj] = 0;
#pragma unroll
      for (int q = 0; q < RS_THREADS / 32; ++q) {
        const int c = wcnt[q][t * PER + jj];
        wcnt[q][t * PER + jj] = acc[jj];
        acc[jj] += c;
      }
    }
    __syncthreads();
    if (valid) {
      const int pos = base[d] + wcnt[w][d] + rank;
      keys_out[pos] = key;
      vals_out[pos] = val;
    }
    __syncthreads();
#pragma unroll
    for (int jj = 0; jj < PER; ++jj) base[t * PER + jj] += acc[jj];
  }
  __syncthreads();
  }
}

// ---- segments of the sorted keys: three launches, no library scan ----------------------------------------
//   seg_count : heads (key != previous key) per tile of 1024 positions
//   rs_scan   : exclusive scan of the tile counts
//   seg_write : seg_idx[i] = 1-based segment number of sorted position i; seg_start / unique at the heads
constexpr int SEG_TILE = 1024, SEG_THREADS = 256, SEG_PER = SEG_TILE / SEG_THREADS;   // a thread owns 4 consecutive positions
__global__ void __launch_bounds__(SEG_THREADS) seg_count_kernel(int n, const unsigned* __restrict__ keys, int* __restrict__ tile_heads) {
  B200_PDL_ENTRY();
  __shared__ int warp_sums[8];
  const int i0 = blockIdx.x * SEG_TILE + threadIdx.x * SEG_PER;
  int h = 0;
  unsigned prev = i0 > 0 && i0 - 1 < n ? keys[i0 - 1] : 0u;
#pragma unroll
  for (int u = 0; u < SEG_PER; ++u) {
    const int i = i0 + u;
    const unsigned k = i < n ? keys[i] : 0u;
    h += (i < n && (i == 0 || k != prev)) ? 1 : 0;
    prev = k;
  }
  int tot;
  block_excl_scan_256(h, warp_sums, &tot);
  if (threadIdx.x == 0) tile_heads[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(SEG_THREADS) seg_write_kernel(int n, const unsigned* __restrict__ keys,
                                                                const int* __restrict__ tile_off, int* __restrict__ seg_idx,
                                                                int* __restrict__ seg_start, int* __restrict__ unique,
                                                                int* __restrict__ n_unique, bool drop_pad) {
  B200_PDL_ENTRY();
  __shared__ int warp_sums[8];
  const int i0 = blockIdx.x * SEG_TILE + threadIdx.x * SEG_PER;
  unsigned k[SEG_PER];
  int hd[SEG_PER], h = 0;
  unsigned prev = i0 > 0 && i0 - 1 < n ? keys[i0 - 1] : 0u;
#pragma unroll
  for (int u = 0; u < SEG_PER; ++u) {
    const int i = i0 + u;
    k[u] = i < n ? keys[i] : 0u;
    hd[u] = (i < n && (i == 0 || k[u] != prev)) ? 1 : 0;
    h += hd[u];
    prev = k[u];
  }
  int s = tile_off[blockIdx.x] + block_excl_scan_256(h, warp_sums, nullptr);   // heads before this thread's positions
#pragma unroll
  for (int u = 0; u < SEG_PER; ++u) {
    const int i = i0 + u;
    if (i >= n) break;
    s += hd[u];                      // inclusive: 1-based segment number of position i
    seg_idx[i] = s;
    if (hd[u]) {
      seg_start[s - 1] = i;
      unique[s - 1] = (int)k[u];
    }
    if (i == n - 1) {
      // exchange padding (key 0xffffffff) sorts last: its segment is not a feature id
      *n_unique = (drop_pad && k[u] == 0xffffffffu) ? s - 1 : s;
      seg_start[s] = n;
    }
  }
}

// hot ids -> two lists by length class, still in the sort half (off the critical path).  The order
// inside a list is whatever the atomics give: it only decides which warp / block sums which segment.
__global__ void seg_long_kernel(const int* n_unique, const int* seg_start, int* long_list, int* big_list,
                                int* counts, int big_t) {
  B200_PDL_ENTRY();
  const int U = *n_unique;
  for (long long seg = blockIdx.x * (long long)blockDim.x + threadIdx.x; seg < U;
       seg += (long long)gridDim.x * blockDim.x) {
    const int len = seg_start[seg + 1] - seg_start[seg];
    if (len > big_t) big_list[atomicAdd(counts + 1, 1)] = (int)seg;
    else if (len > LONG_T) long_list[atomicAdd(counts, 1)] = (int)seg;
  }
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
  unsigned* perm = ws.vals_b.as<unsigned>();          // the sorted pairs always end in (keys_b, vals_b)
  unsigned* keys_sorted = ws.keys_b.as<unsigned>();
  int* hist = ws.cub_tmp.as<int>();
  const int bits = a.key_bits < 1 ? 1 : (a.key_bits > 32 ? 32 : a.key_bits);
  // 9-bit digits would save a whole pass at 17-18 and 25-27 key bits (the 100 M-row ids of the sharded step: 3 passes
  // instead of 4) and do cut the sort's kernel time by 6-10 % -- but the 2-GPU step measured 4 % SLOWER with them
  // (0.477 vs 0.458 ms, same box: the 512-bin scatter blocks hold twice the shared memory and run longer chains
  // beside the GEMMs).  Off unless B200REC_SORT_DB9=1.
  static const bool db9 = [] { const char* e = std::getenv("B200REC_SORT_DB9"); return e && e[0] == '1'; }();
  const int db = db9 && (bits + 8) / 9 < (bits + 7) / 8 ? 9 : 8;
  const int passes = (bits + db - 1) / db, bins = 1 << db;
  int* totals = hist + (size_t)bins * cdiv(n, RS_TILE);
  const int tiles = cdiv(n, RS_TILE);
  static const int bg_group = [] { const char* e = std::getenv("B200REC_SORT_GROUP"); return e && *e ? atoi(e) : 1; }();
  const int group = a.background && bg_group > 1 ? bg_group : 1;   // (measured: 1 is best; ~20 fat blocks make the sort the critical path)
  const int blocks = cdiv(tiles, group);
  const unsigned* kin = (const unsigned*)a.feats;
  const unsigned* vin = nullptr;                      // first pass: the value is the position itself
  for (int p = 0; p < passes; ++p) {
    const bool to_b = ((passes - 1 - p) & 1) == 0;    // ping-pong so that the last pass writes the b buffers
    unsigned* kout = to_b ? ws.keys_b.as<unsigned>() : ws.keys_a.as<unsigned>();
    unsigned* vout = to_b ? ws.vals_b.as<unsigned>() : ws.vals_a.as<unsigned>();
    if (db == 9) {
      B200_LAUNCH(rs_hist_kernel<9>, blocks, RS_THREADS, 0, st, n, kin, db * p, hist, tiles, group, p == 0 ? counters : nullptr);
      B200_LAUNCH(rs_rowscan_kernel, bins, RS_THREADS, 0, st, hist, tiles, totals);
      B200_LAUNCH(rs_scatter_kernel<9>, blocks, RS_THREADS, 0, st, n, kin, vin, db * p, hist, tiles, group, totals, kout, vout);
    } else {
      B200_LAUNCH(rs_hist_kernel<8>, blocks, RS_THREADS, 0, st, n, kin, db * p, hist, tiles, group, p == 0 ? counters : nullptr);
      B200_LAUNCH(rs_rowscan_kernel, bins, RS_THREADS, 0, st, hist, tiles, totals);
      B200_LAUNCH(rs_scatter_kernel<8>, blocks, RS_THREADS, 0, st, n, kin, vin, db * p, hist, tiles, group, totals, kout, vout);
    }
    kin = kout;
    vin = vout;
  }
  // segment table.  vals_a is free again (the last pass read it at most): it becomes seg_idx.
  int* seg_idx = ws.vals_a.as<int>();
  const int stiles = cdiv(n, SEG_TILE);
  int* tile_heads = totals + bins;
  B200_LAUNCH(seg_count_kernel, stiles, SEG_THREADS, 0, st, n, keys_sorted, tile_heads);
  B200_LAUNCH(rs_scan_kernel, 1, 256, 0, st, tile_heads, stiles);
  B200_LAUNCH(seg_write_kernel, stiles, SEG_THREADS, 0, st, n, keys_sorted, tile_heads, seg_idx, ws.seg_start.as<int>(),
              a.unique, a.n_unique, a.drop_pad);
  {
    int grid = cdiv(n, 256);
    if (grid > 148 * 8) grid = 148 * 8;
    B200_LAUNCH(seg_long_kernel, grid, 256, 0, st, a.n_unique, ws.seg_start.as<int>(), ws.long_list.as<int>(),
                ws.long_list.as<int>() + (ws.cap_n + 8) / LONG_T, counters,
                a.K <= BIG_MAX_K ? BIG_T : 0x7fffffff);   // wider rows than the staging buffer: warp role only
  }
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

// inv[i] = index of the distinct id of non-zero i (valid after segsum_sort on the same workspace)
__global__ void seg_inverse_kernel(long long n, const unsigned* perm, const int* seg_idx, int* inv) {
  B200_PDL_ENTRY();
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p < n) inv[perm[p]] = seg_idx[p] - 1;
}
int segsum_inverse(SegSumWorkspace& ws, long long n, int* inv, cudaStream_t st) {
  if (n <= 0) return B200REC_OK;
  B200_LAUNCH(seg_inverse_kernel, cdiv(n, 256), 256, 0, st, n, ws.vals_b.as<unsigned>(), ws.vals_a.as<int>(), inv);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

// ---- where a row of the per-nnz gradient comes from ---------------------------------------------
struct RowSrc {
  const float* dE = nullptr;      // plain: [n,K]
  const float* dw = nullptr;      // plain: [n]
  // fused (X or dX given): computed per row, see the header
  const float* X = nullptr;
  const float* dX = nullptr;
  const float* S = nullptr;
  const float* dlogit = nullptr;
  int F = 1;
  float* dE_keep = nullptr;       // fused: also materialise the per-nnz gradients (tests / callers that ask)
  float* dw_keep = nullptr;
  bool has_e = false, has_w = false, fused = false;
};

template <int K>
__device__ __forceinline__ void load_row(const RowSrc& s, long long p, int sub, float4& v, float& w) {
  if (!s.fused) {
    if (s.has_e) v = ldg_f4(s.dE + p * K + sub * 4);
    if (s.has_w && sub == 0) w = __ldg(s.dw + p);
    return;
  }
  const int b = (int)(p / s.F);
  const float g0 = __ldg(s.dlogit + b);
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
  if (s.dX) g = ld_stream_f4(s.dX + p * K + sub * 4);
  if (s.S) {   // exactly emb_grad_kernel's arithmetic
    const float c = g0 / (float)K;
    const float4 sv = ldg_f4(s.S + (long long)b * K + sub * 4);
    const float4 x = ld_stream_f4(s.X + p * K + sub * 4);
    g.x = fmaf(c, sv.x - x.x, g.x); g.y = fmaf(c, sv.y - x.y, g.y);
    g.z = fmaf(c, sv.z - x.z, g.z); g.w = fmaf(c, sv.w - x.w, g.w);
  }
  v = g;
  if (sub == 0) w = g0;
  if (s.dE_keep) st_f4(s.dE_keep + p * K + sub * 4, g);
  if (s.dw_keep && sub == 0) s.dw_keep[p] = g0;
}

// ---- where a summed row goes ---------------------------------------------------------------------
struct RowSink {
  float* G = nullptr;
  float* gw = nullptr;
  bool push = false;            // store into the owners' peer buffers instead (SegPush)
  SegPush p;
};
template <int K>
__device__ __forceinline__ void sink_row(const RowSink& o, long long seg, int sub, const float4& acc) {
  if (!o.push) {
    if (o.G) st_f4(o.G + seg * K + sub * 4, acc);
    return;
  }
  const int s = o.p.dst[seg];
  if (s < 0) return;            // no slot (full exchange bucket, already flagged): the gradient is dropped
  const int ow = s / o.p.cap;
  const long long slot = (long long)o.p.c.rank * o.p.cap + (s - ow * o.p.cap);
  st_f4(peer_sel(o.p.grad_in.p, ow) + slot * K + sub * 4, acc);
}
__device__ __forceinline__ void sink_w(const RowSink& o, long long seg, float aw) {
  if (!o.push) {
    if (o.gw) o.gw[seg] = aw;
    return;
  }
  const int s = o.p.dst[seg];
  if (s < 0) return;
  const int ow = s / o.p.cap;
  peer_sel(o.p.gw_in.p, ow)[(long long)o.p.c.rank * o.p.cap + (s - ow * o.p.cap)] = aw;
}

// ---- in-order segment sums --------------------------------------------------------------------
template <int LPR>
__device__ __forceinline__ void segsum_short_role(int block, int n_blocks, long long n, const int* seg_idx,
                                                  const unsigned* perm, const RowSrc& src, const RowSink& out) {
  // Walks SORTED POSITIONS, not segments: a lane group looks at position p, and if p is the head of a
  // segment of at most LONG_T rows it sums that segment.  seg_idx (the 1-based segment number of every
  // sorted position, left by the sort half) and perm are read at p-1 .. p+LONG_T: contiguous, coalesced
  // loads, and only TWO dependent load levels (positions -> rows) instead of three through seg_start.
  constexpr int K = 4 * LPR;
  const long long tid = block * (long long)blockDim.x + threadIdx.x;
  const int sub = (int)(tid % LPR);
  const long long n_groups = ((long long)n_blocks * blockDim.x) / LPR;
  for (long long p = tid / LPR; p < n; p += n_groups) {
    const int s0 = seg_idx[p];
    const int sprev = p > 0 ? seg_idx[p - 1] : 0;
    int sn[LONG_T];
    unsigned q[LONG_T];
#pragma unroll
    for (int j = 0; j < LONG_T; ++j) {
      sn[j] = p + 1 + j < n ? seg_idx[p + 1 + j] : 0;   // segment of position p + 1 + j (0: past the end)
      q[j] = p + j < n ? perm[p + j] : 0u;
    }
    if (s0 == sprev) continue;                          // not a segment head
    int len = 1;
#pragma unroll
    for (int j = 0; j < LONG_T; ++j) len += (len == j + 1 && sn[j] == s0) ? 1 : 0;
    if (len > LONG_T) continue;                         // a longer segment: the warp / block roles sum it
    float4 v[LONG_T];
    float w[LONG_T];
#pragma unroll
    for (int j = 0; j < LONG_T; ++j) {
      v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      w[j] = 0.f;
      if (j < len) load_row<K>(src, (long long)q[j], sub, v[j], w[j]);
    }
    float4 acc = make_float4(0.f + v[0].x, 0.f + v[0].y, 0.f + v[0].z, 0.f + v[0].w);
    float aw = 0.f + w[0];
#pragma unroll
    for (int j = 1; j < LONG_T; ++j) {
      if (j < len) {
        acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w;
        aw += w[j];
      }
    }
    const long long seg = s0 - 1;
    if (src.has_e) sink_row<K>(out, seg, sub, acc);
    if (src.has_w && sub == 0) sink_w(out, seg, aw);
  }
}

template <int LPR>
__device__ __forceinline__ void segsum_long_role(int block, int n_blocks, const int* seg_start,
                                                 const unsigned* perm, const RowSrc& src, const RowSink& out,
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
        if (j < len) load_row<K>(src, (long long)pnext[u], sub, v[u], w[u]);
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
    if (src.has_e && slot == 0) sink_row<K>(out, seg, sub, acc);
    if (src.has_w && lane == 0) sink_w(out, seg, aw);
  }
}

// One BLOCK per very hot id (more than BIG_T rows; the hottest id of a Criteo-shaped batch of 8192 owns
// ~430).  The only sequential part of an in-order sum is its chain of fp32 adds; everything else is made
// parallel: all 256 threads fetch up to BIG_CH rows of the segment into shared memory in one round of
// independent loads, then K + 1 lanes of warp 0 (one per component, one for the first-order weight) walk
// the staged rows in non-zero order -- one LDS + one FADD per row instead of five shuffles, and no
// load latency inside the chain.
template <int LPR>
__device__ __forceinline__ void segsum_big_role(int block, int n_blocks, const int* seg_start,
                                                const unsigned* perm, const RowSrc& src, const RowSink& out,
                                                const int* big_list, const int* big_count, float* rows_s,
                                                float* w_s) {
  constexpr int K = 4 * LPR;
  const int n_big = *big_count;
  const int tid = threadIdx.x;
  for (int bi = block; bi < n_big; bi += n_blocks) {
    const int seg = big_list[bi];
    const int start = seg_start[seg];
    const int len = seg_start[seg + 1] - start;
    float acc = 0.f;     // lane k < K of warp 0: component k; lane K: the weight gradient
    for (int base = 0; base < len; base += BIG_CH) {
      const int n = min(BIG_CH, len - base);
      for (int q = tid; q < n * LPR; q += blockDim.x) {
        const int j = q / LPR, sub = q - j * LPR;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        float w = 0.f;
        load_row<K>(src, (long long)perm[start + base + j], sub, v, w);
        *reinterpret_cast<float4*>(rows_s + j * K + sub * 4) = v;
        if (sub == 0) w_s[j] = w;
      }
      __syncthreads();
      if (tid <= K) {
        const float* col = tid < K ? rows_s + tid : w_s;
        const int stride = tid < K ? K : 1;
#pragma unroll 8
        for (int j = 0; j < n; ++j) acc += col[j * stride];
      }
      __syncthreads();
    }
    if (src.has_e) {   // lanes 0 .. K-1 hold one component each: regroup to the 128-bit stores of the sink
      const float a1 = __shfl_down_sync(0xffffffffu, acc, 1), a2 = __shfl_down_sync(0xffffffffu, acc, 2),
                  a3 = __shfl_down_sync(0xffffffffu, acc, 3);
      if (tid < K && (tid & 3) == 0) sink_row<K>(out, seg, tid >> 2, make_float4(acc, a1, a2, a3));
    }
    if (tid == K && src.has_w) sink_w(out, seg, acc);
  }
}

// ONE launch: the first l_blocks blocks take the hot ids (one warp each: they run longest, so they
// start first), the others the short segments.  The hot-id list comes from the sort half.
template <int LPR>
__global__ void __launch_bounds__(256, 4) segsum_kernel(int b_blocks, int l_blocks, long long n, const int* seg_idx,
                                                     const int* seg_start, const unsigned* perm,
                                                     const RowSrc src, const RowSink out,
                                                     const int* long_list, const int* big_list,
                                                     const int* counts) {
  B200_PDL_ENTRY();
  constexpr int SK = 4 * LPR <= BIG_MAX_K ? 4 * LPR : 1;   // wider rows never reach the big role (seg_long_kernel)
  __shared__ __align__(16) float rows_s[BIG_CH * SK];
  __shared__ float w_s[BIG_CH];
  const int b = (int)blockIdx.x;
  if (b < b_blocks) {
    if (4 * LPR <= BIG_MAX_K)
      segsum_big_role<LPR>(b, b_blocks, seg_start, perm, src, out, big_list, counts + 1, rows_s, w_s);
  } else if (b < b_blocks + l_blocks) {
    segsum_long_role<LPR>(b - b_blocks, l_blocks, seg_start, perm, src, out, long_list, counts);
  } else {
    segsum_short_role<LPR>(b - b_blocks - l_blocks, gridDim.x - b_blocks - l_blocks, n, seg_idx, perm, src, out);
  }
  if (out.push) p2p_signal(out.p.c, 2);   // every thread of every block arrives here: phase-2 flag to all owners
}

// any K: one thread per (segment, k), strictly sequential
__global__ void segsum_generic_kernel(const int* n_unique, const int* seg_start,
                                      const unsigned* perm, int K, const float* dE,
                                      const float* dw, float* G, float* gw) {
  B200_PDL_ENTRY();
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
  int grid = cdiv(a.n * LPR, 256);   // one lane group per sorted position (grid-stride beyond 8 blocks per SM)
  if (grid > 148 * 8) grid = 148 * 8;
  long long max_long = a.n / (LONG_T + 1) + 1;
  int lgrid = cdiv(max_long * 32, 256);
  if (lgrid > 148 * 4) lgrid = 148 * 4;     // ~5 000 segments of a Criteo-shaped batch of 8192 exceed 4 rows: one warp each
  RowSrc src;
  src.dE = a.dE; src.dw = a.dw;
  src.fused = a.fused;
  if (a.fused) {
    src.X = a.fX; src.dX = a.fdX; src.S = a.fS; src.dlogit = a.fdlogit; src.F = a.fF > 0 ? a.fF : 1;
    src.dE_keep = a.keep_dE; src.dw_keep = a.keep_dw;
  }
  src.has_e = a.fused ? (a.G != nullptr) : (a.dE != nullptr);
  src.has_w = a.fused ? (a.gw != nullptr) : (a.dw != nullptr);
  const int* big_list = long_list + (ws.cap_n + 8) / LONG_T;
  const int bgrid = 148;     // one block per very hot id; a Criteo-shaped batch of 8192 has ~270 of them
  RowSink out;
  out.G = a.fused || a.dE ? a.G : nullptr;
  out.gw = a.fused || a.dw ? a.gw : nullptr;
  if (a.push) { out.push = true; out.p = *a.push; }
  B200_LAUNCH((segsum_kernel<LPR>), bgrid + lgrid + grid, 256, 0, st, bgrid, lgrid, a.n, ws.vals_a.as<int>(), seg_start, perm, src,
              out, long_list, big_list, counters);
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
  B200_PDL_ENTRY();
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
