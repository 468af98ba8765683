// Row-sharded table: planning the id exchange of one batch.
//
// Reference: the Angel PS shards the `embedding` / `weights` matrices over PS nodes by contiguous
// feature-id ranges (ColumnRangePartitioner, rec/model/ParRecModel.scala:77,81,98,116) and each
// worker pulls the rows of its batch by RPC (:174-177,193-196) and pushes the gradients back
// (:247-250,261-264).  Here every GPU owns the rows
//     owner(id) = (id + id / period) % world,   local row = id / world
// (`period` a multiple of `world`: the `world` consecutive ids of a block rotate over the ranks, and
// the rotation advances every `period` ids so the hot first ids of the 39 fields -- which are all
// congruent mod 8 when the per-field vocab is a power of two -- do not pile up on rank 0), and the
// pull / push are NCCL all-to-all exchanges of fixed-capacity slot buffers:
//     send_ids[owner * cap + slot] = local row (or -1 padding)
// Slot order within an owner is the non-zero order i ascending (stable counting sort), so the
// owner's in-order segment sum sees a deterministic order: by source rank, then by i.
#include "kernels.h"

namespace b200rec {

int ShardPlanWorkspace::reserve(long long n) {
  const size_t ni = (size_t)(n > 0 ? n : 1) + 8;
  B200_TRY(keys.reserve(ni * 4));
  B200_TRY(keys_sorted.reserve(ni * 4));
  B200_TRY(vals.reserve(ni * 4));
  B200_TRY(perm.reserve(ni * 4));
  B200_TRY(offsets.reserve(64 * 4));
  return B200REC_OK;
}
void ShardPlanWorkspace::release() {
  keys.release(); keys_sorted.release(); vals.release(); perm.release(); offsets.release();
  cub_tmp.release();
}

// ---- stable counting sort of the non-zeros by owner (world <= 32): three tiny kernels ---------------
// A radix sort of 1..3 key bits is all fixed overhead; here: (1) per-block owner histograms,
// (2) one block scans them into per-(block, owner) offsets, (3) every block ranks its items per owner
// with ballots (stable: block order, then thread order) and scatters owner / position to sorted order.
constexpr int CS_ITEMS = 2048;   // items per block
constexpr int CS_MAXW = 32;

__device__ __forceinline__ int owner_of(long long id, int world, long long period) {
  return shard_owner(id, world, period);
}

__global__ void __launch_bounds__(256) shard_hist_kernel(long long n, const int* n_dev, int world,
                                                         long long period, const int* feats,
                                                         int* block_counts, int* send_ids,
                                                         long long n_send) {
  B200_PDL_ENTRY();
  __shared__ int cnt[CS_MAXW];
  if (n_dev) n = min(n, (long long)*n_dev);
  if (threadIdx.x < CS_MAXW) cnt[threadIdx.x] = 0;
  __syncthreads();
  const long long base = (long long)blockIdx.x * CS_ITEMS;
  for (int k = threadIdx.x; k < CS_ITEMS; k += 256) {
    const long long i = base + k;
    if (i < n) atomicAdd(&cnt[owner_of(feats[i], world, period)], 1);   // integer counts: order-free
  }
  __syncthreads();
  if (threadIdx.x < world) block_counts[blockIdx.x * world + threadIdx.x] = cnt[threadIdx.x];
  // the slot buffer starts as all padding
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n_send; i += (long long)gridDim.x * 256)
    send_ids[i] = -1;
}

// offsets[o] = first sorted position of owner o (o = 0..world); block_counts -> exclusive offsets.
// The counts are staged in shared memory in chunks (coalesced) so the serial scan runs at smem latency.
__global__ void __launch_bounds__(256) shard_scan_kernel(int n_blocks, int world, int* block_counts,
                                                         int* offsets) {
  B200_PDL_ENTRY();
  constexpr int CHUNK = 4096;            // ints of block_counts per pass
  __shared__ int buf[CHUNK];
  __shared__ int tot[CS_MAXW + 1];
  const int o = threadIdx.x;
  int run = 0;
  const int blocks_per_pass = CHUNK / world;
  for (int b0 = 0; b0 < n_blocks; b0 += blocks_per_pass) {
    const int nb = min(blocks_per_pass, n_blocks - b0);
    for (int k = threadIdx.x; k < nb * world; k += 256) buf[k] = block_counts[b0 * world + k];
    __syncthreads();
    if (o < world) {
      for (int b = 0; b < nb; ++b) {
        const int c = buf[b * world + o];
        buf[b * world + o] = run;
        run += c;
      }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nb * world; k += 256) block_counts[b0 * world + k] = buf[k];
    __syncthreads();
  }
  if (o < world) tot[o] = run;
  __syncthreads();
  if (o == 0) {
    int r = 0;
    for (int k = 0; k < world; ++k) {
      offsets[k] = r;
      r += tot[k];
    }
    offsets[world] = r;
  }
}

__global__ void __launch_bounds__(256) shard_rank_kernel(long long n, const int* n_dev, int world,
                                                         long long period, const int* feats,
                                                         const int* block_offsets, const int* offsets,
                                                         unsigned* owner_sorted, unsigned* perm) {
  B200_PDL_ENTRY();
  if (n_dev) n = min(n, (long long)*n_dev);
  __shared__ int run[CS_MAXW];            // items of each owner already ranked in this block
  __shared__ int wcnt[8][CS_MAXW];        // per-warp counts of the current pass
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < CS_MAXW) run[threadIdx.x] = 0;
  __syncthreads();
  const long long base = (long long)blockIdx.x * CS_ITEMS;
  for (int k0 = 0; k0 < CS_ITEMS; k0 += 256) {
    const long long i = base + k0 + threadIdx.x;
    const int o = i < n ? owner_of(feats[i], world, period) : -1;
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
      const int pos = offsets[o] + block_offsets[blockIdx.x * world + o] + before + my_rank;
      owner_sorted[pos] = (unsigned)o;
      perm[pos] = (unsigned)i;
    }
    __syncthreads();
    if (threadIdx.x < world) {
      int c = 0;
      for (int w = 0; w < 8; ++w) c += wcnt[w][threadIdx.x];
      run[threadIdx.x] += c;
    }
    __syncthreads();
  }
}

__global__ void shard_place_kernel(long long n, int world, long long period, int cap, const int* feats,
                                   const unsigned* owner_sorted, const unsigned* perm,
                                   const int* offsets, int* send_ids, int* dst, int* overflow) {
  B200_PDL_ENTRY();
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int o = (int)owner_sorted[p];
  const int slot = (int)p - offsets[o];
  const long long i = perm[p];
  if (slot >= cap) {  // bucket overflow: flagged (the host raises); the non-zero gets NO slot (zero row, gradient dropped)
    atomicOr(overflow, 1);
    dst[i] = -1;
    return;
  }
  send_ids[(long long)o * cap + slot] = (int)shard_local_row(feats[i], world, period);
  dst[i] = o * cap + slot;
}

// owner keys -> stable sort -> per-owner offsets (ws.keys_sorted / ws.perm / ws.offsets).  When
// send_ids is given it is reset to all padding on the way.
static int shard_sort_impl(ShardPlanWorkspace& ws, long long n, const int* n_dev, int world,
                           long long period, const int* feats, int* send_ids, long long n_send,
                           cudaStream_t st) {
  B200_REQUIRE(world >= 1 && world <= CS_MAXW, B200REC_ERR_ARG, "world size %d out of range", world);
  B200_REQUIRE(shard_period_ok(world, period), B200REC_ERR_ARG,
               "shard period %lld must be a positive multiple of the world size %d (or negative: -rows per rank of "
               "a contiguous-range partition)", period, world);
  B200_TRY(ws.reserve(n));
  const int n_blocks = cdiv(n > 0 ? n : 1, CS_ITEMS);
  B200_TRY(ws.keys.reserve((size_t)n_blocks * world * sizeof(int) + 64));   // block counts / offsets
  int* block_counts = ws.keys.as<int>();
  B200_LAUNCH(shard_hist_kernel, n_blocks, 256, 0, st, n, n_dev, world, period, feats, block_counts, send_ids, n_send);
  B200_LAUNCH(shard_scan_kernel, 1, 256, 0, st, n_blocks, world, block_counts, ws.offsets.as<int>());
  B200_LAUNCH(shard_rank_kernel, n_blocks, 256, 0, st, n, n_dev, world, period, feats, block_counts,
              ws.offsets.as<int>(), ws.keys_sorted.as<unsigned>(), ws.perm.as<unsigned>());
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

// histogram + scan only (the caller ranks and places in one kernel): block offsets in ws.keys,
// owner offsets in ws.offsets
int shard_hist_scan(ShardPlanWorkspace& ws, long long n, const int* n_dev, int world, long long period,
                    const int* feats, cudaStream_t st) {
  B200_REQUIRE(world >= 1 && world <= CS_MAXW, B200REC_ERR_ARG, "world size %d out of range", world);
  B200_REQUIRE(shard_period_ok(world, period), B200REC_ERR_ARG,
               "shard period %lld must be a positive multiple of the world size %d (or negative: -rows per rank of "
               "a contiguous-range partition)", period, world);
  B200_TRY(ws.reserve(n));
  const int n_blocks = cdiv(n > 0 ? n : 1, CS_ITEMS);
  B200_TRY(ws.keys.reserve((size_t)n_blocks * world * sizeof(int) + 64));
  int* block_counts = ws.keys.as<int>();
  B200_LAUNCH(shard_hist_kernel, n_blocks, 256, 0, st, n, n_dev, world, period, feats, block_counts,
              (int*)nullptr, 0LL);
  B200_LAUNCH(shard_scan_kernel, 1, 256, 0, st, n_blocks, world, block_counts, ws.offsets.as<int>());
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

int shard_sort(ShardPlanWorkspace& ws, long long n, const int* n_dev, int world, long long period,
               const int* feats, cudaStream_t st) {
  return shard_sort_impl(ws, n, n_dev, world, period, feats, nullptr, 0, st);
}

int shard_plan(ShardPlanWorkspace& ws, long long n, int world, long long period, int cap,
               const int* feats, int* send_ids, int* dst, int* overflow, cudaStream_t st) {
  ProfTag tag("shard_plan");
  B200_TRY(shard_sort_impl(ws, n, nullptr, world, period, feats, send_ids, (long long)world * cap, st));
  if (n > 0)
    B200_LAUNCH(shard_place_kernel, cdiv(n, 256), 256, 0, st, n, world, period, cap, feats,
                ws.keys_sorted.as<unsigned>(), ws.perm.as<unsigned>(), ws.offsets.as<int>(), send_ids,
                dst, overflow);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

}  // namespace b200rec
