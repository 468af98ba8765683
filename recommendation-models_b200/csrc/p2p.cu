// Row-sharded table exchange over NVLink peer memory: the id dispatch, the row return and the
// gradient push are STORES INTO THE PEERS' BUFFERS issued by the producing kernels themselves
// (fused gather + dispatch), with release/acquire flags in peer memory instead of NCCL all-to-alls.
//
// Reference: the worker <-> PS pull / push of rec/model/ParRecModel.scala:174-177,193-196 (pull
// embedding rows of the batch's ids) and :247-250,261-264 (push their gradients).  Buffers are
// symmetric allocations (same layout on every rank; the host passes the peer-mapped pointers):
//   ids_in [4][G*cap]   local rows requested by each source (block s written by source s; -1 padding).
//                       A ring of 4: ids are dispatched one step ahead, so while step t runs peers may
//                       already write buffer t+1, a slow owner may still read t-1, and t+3 is reset
//   rows_in[G*cap*K], w_in[G*cap]        rows returned by each owner (block o written by owner o)
//   grad_in[G*cap*K], gw_in[G*cap]       per-nnz gradients from each source (block s)
//   dense_in[mats]      this rank's dense gradients, read by every peer (one-shot allreduce, phase 3)
//   dense_out[mats]     two-shot allreduce only: the reduced slices, stored by their owners (phase 4)
//   flags  [5][G]       flags[phase][src] = step number, written by src after its data (release.sys)
// The step number lives in a device counter (advanced by p2p_begin_step) so that a captured CUDA
// graph of the whole step can be replayed.
// A writer kernel ends with: __syncthreads(), then thread 0's __threadfence_system() and one atomicInc
// per block, and the LAST block stores the step number into every peer's flag.  A one-warp wait kernel
// spins (acquire.sys) until all G flags of a phase reach the step; after B200REC_P2P_TIMEOUT_MS it gives up
// and sets DEV_PEER_TIMEOUT in the handle's status word (no trap: a late peer must not kill the job).
#include <cstdlib>

#include "kernels.h"
#include "p2p_dev.cuh"

namespace b200rec {

// threads 0..world-1 spin until every source's flag of `phase` reaches the step; then the block syncs.
// A peer that is merely LATE (data-loader hiccup, first graph capture, checkpoint save on its host) must
// not kill the job: the wait is bounded by wall time (globaltimer, B200REC_P2P_TIMEOUT_MS, default 30 s)
// and on expiry it sets DEV_PEER_TIMEOUT in the handle's sticky status word and returns -- the host
// reports it (b200rec_model_sync / the sharded step's check) instead of every rank dying on a trap.
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void p2p_wait_block(const int* flags, int phase, int world, int step,
                                               unsigned long long timeout_ns, int* status) {
  const int s = threadIdx.x;
  if (s < world) {
    if (ld_acquire_sys(flags + phase * world + s) < step) {
      const unsigned long long t0 = global_ns();
      unsigned spins = 0;
      while (ld_acquire_sys(flags + phase * world + s) < step) {
        if ((++spins & 63u) == 0 && global_ns() - t0 > timeout_ns) {
          if (status) atomicOr(status, DEV_PEER_TIMEOUT);
          break;
        }
      }
    }
  }
  __syncthreads();
}

static unsigned long long p2p_timeout_ns() {
  static unsigned long long v = [] {
    const char* e = getenv("B200REC_P2P_TIMEOUT_MS");
    const double ms = e ? atof(e) : 30000.0;
    return (unsigned long long)((ms > 1.0 ? ms : 1.0) * 1e6);
  }();
  return v;
}

__global__ void p2p_wait_kernel(const int* flags, int phase, int world, int step, const int* step_ptr,
                                unsigned long long timeout_ns, int* status) {
  B200_PDL_ENTRY();
  p2p_wait_block(flags, phase, world, step > 0 ? step : *step_ptr - step, timeout_ns, status);
}

int p2p_wait(const int* flags, int phase, int world, int step, const int* step_ptr, int* status, cudaStream_t st) {
  ProfTag tag("p2p_wait");
  B200_LAUNCH(p2p_wait_kernel, 1, 32, 0, st, flags, phase, world, step, step_ptr, p2p_timeout_ns(), status);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

// start of a step: advance the device step counter and reset the ids buffer of the NEXT step to the
// -1 padding (nobody writes that buffer before this rank's next phase-0 signal)
__global__ void p2p_begin_step_kernel(int* step_ctr, int* ids_next, long long n) {
  B200_PDL_ENTRY();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    ids_next[i] = -1;
  if (blockIdx.x == 0 && threadIdx.x == 0) *step_ctr += 1;
}
int p2p_begin_step(int* step_ctr, int* ids_next, long long n, cudaStream_t st) {
  ProfTag tag("p2p_dispatch_ids");
  int grid = cdiv(n > 0 ? n : 1, 1024);
  if (grid > 148 * 4) grid = 148 * 4;
  B200_LAUNCH(p2p_begin_step_kernel, grid, 256, 0, st, step_ctr, ids_next, n);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

// writer kernels run as few fat blocks (2 per SM): every block ends with a system-scope fence
constexpr int EX_THREADS = 512;
constexpr int EX_UNR = 4;

// ---- phase 3: dense gradients, one-shot allreduce over peer loads -----------------------------------
// Reference: the dense gradients go to the PS like the embedding ones (ParRecModel.scala:247-264);
// data-parallel replicas need their SUM.  Every rank publishes its vector in symmetric memory and
// then reads all G vectors, adding them in rank order: all replicas get bit-identical sums.
__global__ void __launch_bounds__(EX_THREADS) p2p_publish_kernel(long long n, const float* src, float* mine, P2P c) {
  B200_PDL_ENTRY();
  const long long n4 = n >> 2;
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(mine);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x)
    d4[i] = s4[i];
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) mine[n4 * 4 + threadIdx.x] = src[n4 * 4 + threadIdx.x];
  p2p_signal(c, 3);
}

__global__ void __launch_bounds__(256) p2p_reduce_kernel(long long n, float* dst, P2P c, PeerF bufs) {
  B200_PDL_ENTRY();
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 v[P2P_MAX];
#pragma unroll
    for (int p = 0; p < P2P_MAX; ++p)
      if (p < c.world) v[p] = __ldcg(reinterpret_cast<const float4*>(bufs.p[p]) + i);
    float4 a = v[0];
#pragma unroll
    for (int p = 1; p < P2P_MAX; ++p)
      if (p < c.world) {
        a.x = __fadd_rn(a.x, v[p].x); a.y = __fadd_rn(a.y, v[p].y);
        a.z = __fadd_rn(a.z, v[p].z); a.w = __fadd_rn(a.w, v[p].w);
      }
    reinterpret_cast<float4*>(dst)[i] = a;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = n4 * 4 + threadIdx.x;
    float a = __ldcg(bufs.p[0] + i);
#pragma unroll
    for (int p = 1; p < P2P_MAX; ++p)
      if (p < c.world) a = __fadd_rn(a, __ldcg(bufs.p[p] + i));
    dst[i] = a;
  }
}

// Two-shot form for larger groups: rank r sums slice r of all published vectors (rank order) and
// stores the result into every peer's `out` buffer (flag 4); then everyone copies `out` back.  NVLink
// bytes per rank: 2 (G-1)/G n instead of (G-1) n.  n4 = float4 count of the zero-padded vectors.
__global__ void __launch_bounds__(256) p2p_reduce_scatter_kernel(long long n4, P2P c, PeerF bufs, PeerF outs) {
  B200_PDL_ENTRY();
  const long long lo = n4 * c.rank / c.world, hi = n4 * (c.rank + 1) / c.world;
  for (long long i = lo + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hi;
       i += (long long)gridDim.x * blockDim.x) {
    float4 v[P2P_MAX];
#pragma unroll
    for (int p = 0; p < P2P_MAX; ++p)
      if (p < c.world) v[p] = __ldcg(reinterpret_cast<const float4*>(bufs.p[p]) + i);
    float4 a = v[0];
#pragma unroll
    for (int p = 1; p < P2P_MAX; ++p)
      if (p < c.world) {
        a.x = __fadd_rn(a.x, v[p].x); a.y = __fadd_rn(a.y, v[p].y);
        a.z = __fadd_rn(a.z, v[p].z); a.w = __fadd_rn(a.w, v[p].w);
      }
#pragma unroll
    for (int p = 0; p < P2P_MAX; ++p)
      if (p < c.world) reinterpret_cast<float4*>(outs.p[p])[i] = a;
  }
  p2p_signal(c, 4);
}

__global__ void __launch_bounds__(256) p2p_copy_back_kernel(long long n, const float* src, float* dst) {
  B200_PDL_ENTRY();
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x)
    reinterpret_cast<float4*>(dst)[i] = __ldcg(reinterpret_cast<const float4*>(src) + i);
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) dst[n4 * 4 + threadIdx.x] = __ldcg(src + n4 * 4 + threadIdx.x);
}

// bufs: the published vectors (padded to a multiple of 4 floats, pad zero-initialised by the host);
// outs (may be empty: one-shot): the result buffers of the two-shot form
int p2p_allreduce(long long n, float* inout, const int* flags_local, const P2P& c, const PeerF& bufs,
                  const PeerF* outs, int* status, cudaStream_t st) {
  if (n <= 0) return B200REC_OK;
  ProfTag tag("p2p_allreduce");
  int grid = cdiv(n / 4 > 0 ? n / 4 : 1, 256);
  if (grid > 148 * 4) grid = 148 * 4;
  int pgrid = cdiv(n / 4 > 0 ? n / 4 : 1, EX_THREADS);
  if (pgrid > 148 * 2) pgrid = 148 * 2;
  B200_LAUNCH(p2p_publish_kernel, pgrid, EX_THREADS, 0, st, n, (const float*)inout, bufs.p[c.rank], c);
  // a one-warp kernel does the spinning: the reduce blocks must not hold SMs while a peer is late
  B200_LAUNCH(p2p_wait_kernel, 1, 32, 0, st, flags_local, 3, c.world, c.step, c.step_ptr, p2p_timeout_ns(), status);
  if (!outs) {
    B200_LAUNCH(p2p_reduce_kernel, grid, 256, 0, st, n, inout, c, bufs);
  } else {
    const long long n4 = (n + 3) / 4;
    int g2 = cdiv(cdiv(n4, c.world), 256);
    if (g2 > 148 * 2) g2 = 148 * 2;
    if (g2 < 1) g2 = 1;
    B200_LAUNCH(p2p_reduce_scatter_kernel, g2, 256, 0, st, n4, c, bufs, *outs);
    B200_LAUNCH(p2p_wait_kernel, 1, 32, 0, st, flags_local, 4, c.world, c.step, c.step_ptr, p2p_timeout_ns(), status);
    B200_LAUNCH(p2p_copy_back_kernel, grid, 256, 0, st, n, (const float*)outs->p[c.rank], inout);
  }
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

// ---- phase 0: ids to their owners ------------------------------------------------------------------
// Ranking (stable counting sort by owner, csrc/shard.cu) and placement in ONE kernel: an item's sorted
// position gives its slot; its local row is stored straight into the owner's ids_in.
constexpr int PCS_ITEMS = 2048;   // must equal CS_ITEMS of shard.cu
constexpr int PCS_MAXW = 32;

__global__ void __launch_bounds__(256) p2p_rank_place_kernel(long long n, const int* n_dev, long long period,
                                                             int cap, const int* feats,
                                                             const int* block_offsets, int* dst,
                                                             int* overflow, P2P c, PeerI ids_in) {
  B200_PDL_ENTRY();
  __shared__ int run[PCS_MAXW];
  __shared__ int wcnt[8][PCS_MAXW];
  if (n_dev) n = min(n, (long long)*n_dev);
  const int world = c.world;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < PCS_MAXW) run[threadIdx.x] = 0;
  __syncthreads();
  const long long base = (long long)blockIdx.x * PCS_ITEMS;
  for (int k0 = 0; k0 < PCS_ITEMS; k0 += 256) {
    const long long i = base + k0 + threadIdx.x;
    long long id = 0;
    int o = -1;
    if (i < n) {
      id = feats[i];
      o = shard_owner(id, world, period);
    }
    int my_rank = 0;
    for (int q = 0; q < world; ++q) {
      const unsigned m = __ballot_sync(0xffffffffu, o == q);
      if (o == q) my_rank = __popc(m & ((1u << lane) - 1));
      if (lane == 0) wcnt[warp][q] = __popc(m);
    }
    __syncthreads();
    if (o >= 0) {
      int before = run[o];
      for (int w = 0; w < warp; ++w) before += wcnt[w][o];
      const int slot = block_offsets[blockIdx.x * world + o] + before + my_rank;   // position within owner o
      if (slot >= cap) {
        // the bucket is full: flag it (the host raises) and give the id NO slot -- its rows read as zero
        // and its gradient is dropped, never another id's slot
        atomicOr(overflow, 1);
        dst[i] = -1;
      } else {
        peer_sel(ids_in.p, o)[(long long)c.rank * cap + slot] = (int)shard_local_row(id, world, period);   // store into the owner's memory
        dst[i] = o * cap + slot;
      }
    }
    __syncthreads();
    if (threadIdx.x < world) {
      int t = 0;
      for (int w = 0; w < 8; ++w) t += wcnt[w][threadIdx.x];
      run[threadIdx.x] += t;
    }
    __syncthreads();
  }
  p2p_signal(c, 0);
}

int shard_hist_scan(ShardPlanWorkspace& ws, long long n, const int* n_dev, int world, long long period,
                    const int* feats, cudaStream_t st);

int p2p_plan(ShardPlanWorkspace& ws, long long n, const int* n_dev, long long period, int cap,
             const int* feats, int* dst, int* overflow, const P2P& c, const PeerI& ids_in, cudaStream_t st) {
  ProfTag tag("p2p_dispatch_ids");
  B200_TRY(shard_hist_scan(ws, n, n_dev, c.world, period, feats, st));
  const int n_blocks = cdiv(n > 0 ? n : 1, PCS_ITEMS);
  B200_LAUNCH(p2p_rank_place_kernel, n_blocks, 256, 0, st, n, n_dev, period, cap, feats, ws.keys.as<int>(),
              dst, overflow, c, ids_in);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

// per-nnz slot = slot of the non-zero's distinct id  (dispatch over distinct ids: dedup before exchange)
__global__ void p2p_compose_kernel(long long n, const unsigned* perm, const int* seg_idx,
                                   const int* dst_unique, int* dst) {
  B200_PDL_ENTRY();
  // sorted position p of the batch's ids: non-zero perm[p] belongs to distinct id seg_idx[p] - 1
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p < n) dst[perm[p]] = dst_unique[seg_idx[p] - 1];
}
int p2p_compose(SegSumWorkspace& ws, long long n, const int* dst_unique, int* dst, cudaStream_t st) {
  if (n <= 0) return B200REC_OK;
  ProfTag tag("p2p_dispatch_ids");
  B200_LAUNCH(p2p_compose_kernel, cdiv(n, 256), 256, 0, st, n, ws.vals_b.as<unsigned>(), ws.vals_a.as<int>(),
              dst_unique, dst);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

// ---- phase 1: owner-side gather, rows stored straight into the requesters' buffers -----------------
// Few fat blocks (2 per SM, 512 threads): every block ends with a system-scope fence, which is the
// expensive part of a writer kernel; each thread keeps EX_UNR independent id -> row chains in flight.
static int ex_grid(long long items) {
  int grid = cdiv(cdiv(items, EX_UNR), EX_THREADS);
  if (grid > 148 * 2) grid = 148 * 2;
  return grid < 1 ? 1 : grid;
}

template <int LPR>
__global__ void __launch_bounds__(EX_THREADS) p2p_gather_kernel(long long rows, int cap, const int* ids_in,
                                                                const float* table, const float* wtable, P2P c,
                                                                PeerF rows_in, PeerF w_in, int* err) {
  B200_PDL_ENTRY();
  constexpr int K = 4 * LPR;
  const long long n_vec = (long long)c.world * cap * LPR;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; t0 < n_vec; t0 += EX_UNR * stride) {
    long long id[EX_UNR];
#pragma unroll
    for (int u = 0; u < EX_UNR; ++u) {
      const long long t = t0 + u * stride;
      id[u] = t < n_vec ? (long long)ids_in[t / LPR] : -1;   // slot in my ids_in: source * cap + slot
      if (id[u] >= rows) {
        atomicOr(err, DEV_BAD_ID);
        id[u] = 0;
      }
    }
    float4 v[EX_UNR];
    float w[EX_UNR];
#pragma unroll
    for (int u = 0; u < EX_UNR; ++u) {
      const int sub = (int)((t0 + u * stride) % LPR);
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      w[u] = 0.f;
      if (id[u] >= 0) {                    // negative = padding: the requester never reads this slot
        if (table) v[u] = ldg_f4(table + id[u] * K + sub * 4);
        if (sub == 0) w[u] = __ldg(wtable + id[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < EX_UNR; ++u) {
      if (id[u] < 0) continue;
      const long long t = t0 + u * stride;
      const long long j = t / LPR;
      const int sub = (int)(t - j * LPR);
      const int src = (int)(j / cap);
      const long long slot = j - (long long)src * cap;
      const long long o = (long long)c.rank * cap + slot;   // my block in the requester's buffers
      if (table) st_f4(peer_sel(rows_in.p, src) + o * K + sub * 4, v[u]);
      if (sub == 0) peer_sel(w_in.p, src)[o] = w[u];
    }
  }
  p2p_signal(c, 1);
}

int p2p_gather(long long rows, int K, int cap, const int* ids_in, const float* table, const float* wtable,
               const P2P& c, const PeerF& rows_in, const PeerF& w_in, int* err, cudaStream_t st) {
  ProfTag tag("p2p_gather_rows");
  const long long n = (long long)c.world * cap;
  const int grid = ex_grid(n * (K / 4 > 0 ? K / 4 : 1));
  switch (K) {
    case 4: B200_LAUNCH(p2p_gather_kernel<1>, grid, EX_THREADS, 0, st, rows, cap, ids_in, table, wtable, c, rows_in, w_in, err); break;
    case 8: B200_LAUNCH(p2p_gather_kernel<2>, grid, EX_THREADS, 0, st, rows, cap, ids_in, table, wtable, c, rows_in, w_in, err); break;
    case 16: B200_LAUNCH(p2p_gather_kernel<4>, grid, EX_THREADS, 0, st, rows, cap, ids_in, table, wtable, c, rows_in, w_in, err); break;
    case 32: B200_LAUNCH(p2p_gather_kernel<8>, grid, EX_THREADS, 0, st, rows, cap, ids_in, table, wtable, c, rows_in, w_in, err); break;
    case 64: B200_LAUNCH(p2p_gather_kernel<16>, grid, EX_THREADS, 0, st, rows, cap, ids_in, table, wtable, c, rows_in, w_in, err); break;
    default: set_error("p2p exchange supports embeddingDim in {4,8,16,32,64}, got %d", K); return B200REC_ERR_ARG;
  }
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

// ---- phase 2: per-nnz gradients stored into the owners' buffers (after the dense backward) ----------
template <int LPR>
__global__ void __launch_bounds__(EX_THREADS) p2p_push_grads_kernel(long long n, const int* n_dev, int cap,
                                                                    const int* dst, const float* dE,
                                                                    const float* dw, P2P c, PeerF grad_in,
                                                                    PeerF gw_in) {
  B200_PDL_ENTRY();
  constexpr int K = 4 * LPR;
  if (n_dev) n = min(n, (long long)*n_dev);
  const long long n_vec = n * LPR;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; t0 < n_vec; t0 += EX_UNR * stride) {
    int s[EX_UNR];
    float4 v[EX_UNR];
    float w[EX_UNR];
#pragma unroll
    for (int u = 0; u < EX_UNR; ++u) {
      const long long t = t0 + u * stride;
      const long long i = t / LPR;
      const int sub = (int)(t - i * LPR);
      s[u] = -1;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      w[u] = 0.f;
      if (t < n_vec) {
        s[u] = dst[i];
        if (dE) v[u] = ld_stream_f4(dE + i * K + sub * 4);
        if (sub == 0) w[u] = dw[i];
      }
    }
#pragma unroll
    for (int u = 0; u < EX_UNR; ++u) {
      if (s[u] < 0) continue;
      const int sub = (int)((t0 + u * stride) % LPR);
      const int o = s[u] / cap;
      const long long slot = (long long)c.rank * cap + (s[u] - o * cap);
      if (dE) st_f4(peer_sel(grad_in.p, o) + slot * K + sub * 4, v[u]);
      if (sub == 0) peer_sel(gw_in.p, o)[slot] = w[u];
    }
  }
  p2p_signal(c, 2);
}

int p2p_push_grads(long long n, const int* n_dev, int K, int cap, const int* dst, const float* dE,
                   const float* dw, const P2P& c, const PeerF& grad_in, const PeerF& gw_in, cudaStream_t st) {
  ProfTag tag("p2p_push_grads");
  const int grid = ex_grid(n * (K / 4 > 0 ? K / 4 : 1));
  switch (K) {
    case 4: B200_LAUNCH(p2p_push_grads_kernel<1>, grid, EX_THREADS, 0, st, n, n_dev, cap, dst, dE, dw, c, grad_in, gw_in); break;
    case 8: B200_LAUNCH(p2p_push_grads_kernel<2>, grid, EX_THREADS, 0, st, n, n_dev, cap, dst, dE, dw, c, grad_in, gw_in); break;
    case 16: B200_LAUNCH(p2p_push_grads_kernel<4>, grid, EX_THREADS, 0, st, n, n_dev, cap, dst, dE, dw, c, grad_in, gw_in); break;
    case 32: B200_LAUNCH(p2p_push_grads_kernel<8>, grid, EX_THREADS, 0, st, n, n_dev, cap, dst, dE, dw, c, grad_in, gw_in); break;
    case 64: B200_LAUNCH(p2p_push_grads_kernel<16>, grid, EX_THREADS, 0, st, n, n_dev, cap, dst, dE, dw, c, grad_in, gw_in); break;
    default: set_error("p2p exchange supports embeddingDim in {4,8,16,32,64}, got %d", K); return B200REC_ERR_ARG;
  }
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

}  // namespace b200rec
