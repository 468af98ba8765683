// Row-sharded table exchange over NVLink peer memory: the id dispatch, the row return and the
// gradient push are STORES INTO THE PEERS' BUFFERS issued by the producing kernels themselves
// (fused gather + dispatch), with release/acquire flags in peer memory instead of NCCL all-to-alls.
//
// Reference: the worker <-> PS pull / push of rec/model/ParRecModel.scala:174-177,193-196 (pull
// embedding rows of the batch's ids) and :247-250,261-264 (push their gradients).  Buffers are
// symmetric allocations (same layout on every rank; the host passes the peer-mapped pointers):
//   ids_in [2][G*cap]   local rows requested by each source (block s written by source s; -1 padding;
//                       double-buffered by step parity so the owner can reset the next one)
//   rows_in[G*cap*K], w_in[G*cap]        rows returned by each owner (block o written by owner o)
//   grad_in[G*cap*K], gw_in[G*cap]       per-nnz gradients from each source (block s)
//   flags  [3][G]       flags[phase][src] = step number, written by src after its data (release.sys)
// A writer kernel ends with: __threadfence_system() by every thread, __syncthreads(), one atomicInc
// per block, and the LAST block stores the step number into every peer's flag.  A one-warp wait kernel
// spins (acquire.sys) until all G flags of a phase reach the step, trapping after ~2 s.
#include "kernels.h"

namespace b200rec {

__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// end-of-kernel signal: every thread of every block must call this (convergently)
__device__ __forceinline__ void p2p_signal(const P2P& c, int phase) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned last = gridDim.x * gridDim.y - 1;
    if (atomicInc(c.block_counter, last) == last) {
      __threadfence_system();
      for (int p = 0; p < c.world; ++p) st_release_sys(c.flags[p] + phase * c.world + c.rank, c.step);
    }
  }
}

__global__ void p2p_wait_kernel(const int* flags, int phase, int world, int step) {
  const int s = threadIdx.x;
  if (s < world) {
    const long long t0 = clock64();
    while (ld_acquire_sys(flags + phase * world + s) < step)
      if (clock64() - t0 > 4000000000LL) __trap();   // a lost peer must fail the launch, not hang
  }
  __syncthreads();
}

int p2p_wait(const int* flags, int phase, int world, int step, cudaStream_t st) {
  ProfTag tag("p2p_wait");
  B200_LAUNCH(p2p_wait_kernel, 1, 32, 0, st, flags, phase, world, step);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

// ---- phase 0: ids to their owners ------------------------------------------------------------------
__global__ void p2p_place_kernel(long long n, const int* n_dev, int cap, const int* feats,
                                 const unsigned* owner_sorted, const unsigned* perm, const int* offsets,
                                 int* dst, int* overflow, P2P c, PeerI ids_in) {
  if (n_dev) n = min(n, (long long)*n_dev);
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p < n) {
    const int o = (int)owner_sorted[p];
    const int slot = (int)p - offsets[o];
    const long long i = perm[p];
    if (slot >= cap) {
      atomicOr(overflow, 1);
      dst[i] = o * cap;
    } else {
      ids_in.p[o][(long long)c.rank * cap + slot] = feats[i] / c.world;   // store into the owner's memory
      dst[i] = o * cap + slot;
    }
  }
  p2p_signal(c, 0);
}

// shard.cu: owner keys -> stable sort -> per-owner offsets
int p2p_plan(ShardPlanWorkspace& ws, long long n, const int* n_dev, long long period, int cap,
             const int* feats, int* dst, int* overflow, const P2P& c, const PeerI& ids_in, cudaStream_t st) {
  ProfTag tag("p2p_dispatch_ids");
  B200_TRY(shard_sort(ws, n, n_dev, c.world, period, feats, st));
  int grid = cdiv(n > 0 ? n : 1, 256);
  B200_LAUNCH(p2p_place_kernel, grid, 256, 0, st, n, n_dev, cap, feats, ws.keys_sorted.as<unsigned>(),
              ws.perm.as<unsigned>(), ws.offsets.as<int>(), dst, overflow, c, ids_in);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

// per-nnz slot = slot of the non-zero's distinct id  (dispatch over distinct ids: dedup before exchange)
__global__ void p2p_compose_kernel(long long n, const int* inv, const int* dst_unique, int* dst) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = dst_unique[inv[i]];
}
int p2p_compose(long long n, const int* inv, const int* dst_unique, int* dst, cudaStream_t st) {
  if (n <= 0) return B200REC_OK;
  ProfTag tag("p2p_dispatch_ids");
  B200_LAUNCH(p2p_compose_kernel, cdiv(n, 256), 256, 0, st, n, inv, dst_unique, dst);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

// ---- phase 1: owner-side gather, rows stored straight into the requesters' buffers -----------------
template <int LPR>
__global__ void __launch_bounds__(256) p2p_gather_kernel(long long rows, int cap, const int* ids_in,
                                                         const float* table, const float* wtable, P2P c,
                                                         PeerF rows_in, PeerF w_in, int* err) {
  constexpr int K = 4 * LPR;
  const long long n_vec = (long long)c.world * cap * LPR;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n_vec;
       t += (long long)gridDim.x * blockDim.x) {
    const long long j = t / LPR;          // slot in my ids_in: source * cap + slot
    const int sub = (int)(t - j * LPR);
    long long id = ids_in[j];
    if (id < 0) continue;                 // padding: the requester never reads this slot
    if (id >= rows) {
      atomicOr(err, DEV_BAD_ID);
      id = 0;
    }
    const int src = (int)(j / cap);
    const long long slot = j - (long long)src * cap;
    const long long o = (long long)c.rank * cap + slot;   // my block in the requester's buffers
    if (table) st_f4(rows_in.p[src] + o * K + sub * 4, ldg_f4(table + id * K + sub * 4));
    if (sub == 0) w_in.p[src][o] = __ldg(wtable + id);
  }
  p2p_signal(c, 1);
}

int p2p_gather(long long rows, int K, int cap, const int* ids_in, const float* table, const float* wtable,
               const P2P& c, const PeerF& rows_in, const PeerF& w_in, int* err, cudaStream_t st) {
  ProfTag tag("p2p_gather_rows");
  const long long n = (long long)c.world * cap;
  int grid = cdiv(n * (K / 4 > 0 ? K / 4 : 1), 256);
  if (grid > 148 * 16) grid = 148 * 16;
  if (grid < 1) grid = 1;
  switch (K) {
    case 4: B200_LAUNCH(p2p_gather_kernel<1>, grid, 256, 0, st, rows, cap, ids_in, table, wtable, c, rows_in, w_in, err); break;
    case 8: B200_LAUNCH(p2p_gather_kernel<2>, grid, 256, 0, st, rows, cap, ids_in, table, wtable, c, rows_in, w_in, err); break;
    case 16: B200_LAUNCH(p2p_gather_kernel<4>, grid, 256, 0, st, rows, cap, ids_in, table, wtable, c, rows_in, w_in, err); break;
    case 32: B200_LAUNCH(p2p_gather_kernel<8>, grid, 256, 0, st, rows, cap, ids_in, table, wtable, c, rows_in, w_in, err); break;
    case 64: B200_LAUNCH(p2p_gather_kernel<16>, grid, 256, 0, st, rows, cap, ids_in, table, wtable, c, rows_in, w_in, err); break;
    default: set_error("p2p exchange supports embeddingDim in {4,8,16,32,64}, got %d", K); return B200REC_ERR_ARG;
  }
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

// ---- phase 2: per-nnz gradients stored into the owners' buffers (after the dense backward) ----------
template <int LPR>
__global__ void __launch_bounds__(256) p2p_push_grads_kernel(long long n, const int* n_dev, int cap,
                                                             const int* dst, const float* dE,
                                                             const float* dw, P2P c, PeerF grad_in,
                                                             PeerF gw_in) {
  constexpr int K = 4 * LPR;
  if (n_dev) n = min(n, (long long)*n_dev);
  const long long n_vec = n * LPR;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n_vec;
       t += (long long)gridDim.x * blockDim.x) {
    const long long i = t / LPR;
    const int sub = (int)(t - i * LPR);
    const int s = dst[i];
    const int o = s / cap;
    const long long slot = (long long)c.rank * cap + (s - o * cap);
    if (dE) st_f4(grad_in.p[o] + slot * K + sub * 4, ld_stream_f4(dE + i * K + sub * 4));
    if (sub == 0) gw_in.p[o][slot] = dw[i];
  }
  p2p_signal(c, 2);
}

int p2p_push_grads(long long n, const int* n_dev, int K, int cap, const int* dst, const float* dE,
                   const float* dw, const P2P& c, const PeerF& grad_in, const PeerF& gw_in, cudaStream_t st) {
  ProfTag tag("p2p_push_grads");
  int grid = cdiv(n * (K / 4 > 0 ? K / 4 : 1), 256);
  if (grid > 148 * 16) grid = 148 * 16;
  if (grid < 1) grid = 1;
  switch (K) {
    case 4: B200_LAUNCH(p2p_push_grads_kernel<1>, grid, 256, 0, st, n, n_dev, cap, dst, dE, dw, c, grad_in, gw_in); break;
    case 8: B200_LAUNCH(p2p_push_grads_kernel<2>, grid, 256, 0, st, n, n_dev, cap, dst, dE, dw, c, grad_in, gw_in); break;
    case 16: B200_LAUNCH(p2p_push_grads_kernel<4>, grid, 256, 0, st, n, n_dev, cap, dst, dE, dw, c, grad_in, gw_in); break;
    case 32: B200_LAUNCH(p2p_push_grads_kernel<8>, grid, 256, 0, st, n, n_dev, cap, dst, dE, dw, c, grad_in, gw_in); break;
    case 64: B200_LAUNCH(p2p_push_grads_kernel<16>, grid, 256, 0, st, n, n_dev, cap, dst, dE, dw, c, grad_in, gw_in); break;
    default: set_error("p2p exchange supports embeddingDim in {4,8,16,32,64}, got %d", K); return B200REC_ERR_ARG;
  }
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

}  // namespace b200rec
