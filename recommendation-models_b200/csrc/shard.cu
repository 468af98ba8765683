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
#include <cub/device/device_radix_sort.cuh>

#include "kernels.h"

namespace b200rec {

int ShardPlanWorkspace::reserve(long long n) {
  const size_t ni = (size_t)(n > 0 ? n : 1) + 8;
  B200_TRY(keys.reserve(ni * 4));
  B200_TRY(keys_sorted.reserve(ni * 4));
  B200_TRY(vals.reserve(ni * 4));
  B200_TRY(perm.reserve(ni * 4));
  B200_TRY(offsets.reserve(64 * 4));
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned*)nullptr, (unsigned*)nullptr,
                                  (const unsigned*)nullptr, (unsigned*)nullptr, (int)ni, 0, 8);
  B200_TRY(cub_tmp.reserve(bytes + 1024));
  return B200REC_OK;
}
void ShardPlanWorkspace::release() {
  keys.release(); keys_sorted.release(); vals.release(); perm.release(); offsets.release();
  cub_tmp.release();
}

__global__ void shard_owner_kernel(long long n, int world, long long period, const int* feats,
                                   unsigned* owner, unsigned* iota, int* send_ids, long long n_send) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t < n) {
    const long long id = feats[t];
    owner[t] = (unsigned)((id + id / period) % world);
    iota[t] = (unsigned)t;
  }
  // the slot buffer starts as all padding
  for (long long i = t; i < n_send; i += (long long)gridDim.x * blockDim.x) send_ids[i] = -1;
}

// offsets[o] = first sorted position whose owner >= o  (o = 0..world)
__global__ void shard_offsets_kernel(long long n, int world, const unsigned* owner_sorted,
                                     int* offsets) {
  const int o = threadIdx.x;
  if (o > world) return;
  long long lo = 0, hi = n;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (owner_sorted[mid] < (unsigned)o) lo = mid + 1; else hi = mid;
  }
  offsets[o] = (int)lo;
}

__global__ void shard_place_kernel(long long n, int world, int cap, const int* feats,
                                   const unsigned* owner_sorted, const unsigned* perm,
                                   const int* offsets, int* send_ids, int* dst, int* overflow) {
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int o = (int)owner_sorted[p];
  const int slot = (int)p - offsets[o];
  const long long i = perm[p];
  if (slot >= cap) {  // bucket overflow: reported, the row is dropped (dst points at slot 0)
    atomicOr(overflow, 1);
    dst[i] = o * cap;
    return;
  }
  send_ids[(long long)o * cap + slot] = feats[i] / world;
  dst[i] = o * cap + slot;
}

int shard_plan(ShardPlanWorkspace& ws, long long n, int world, long long period, int cap,
               const int* feats, int* send_ids, int* dst, int* overflow, cudaStream_t st) {
  ProfTag tag("shard_plan");
  B200_REQUIRE(world >= 1 && world <= 32, B200REC_ERR_ARG, "world size %d out of range", world);
  B200_REQUIRE(period >= world && period % world == 0, B200REC_ERR_ARG,
               "shard period %lld must be a positive multiple of the world size %d", period, world);
  B200_TRY(ws.reserve(n));
  const long long n_send = (long long)world * cap;
  const long long cover = n > n_send ? n : n_send;
  int grid = cdiv(cover, 256);
  if (grid > 148 * 16) grid = 148 * 16;
  if (grid < cdiv(n, 256)) grid = cdiv(n, 256);
  if (grid < 1) grid = 1;
  B200_LAUNCH(shard_owner_kernel, grid, 256, 0, st, n, world, period, feats, ws.keys.as<unsigned>(),
              ws.vals.as<unsigned>(), send_ids, n_send);
  if (n > 0) {
    size_t tmp = ws.cub_tmp.cap;
    int bits = 1;
    while ((1 << bits) < world) ++bits;
    if (tl_prof) tl_prof->begin("cub::DeviceRadixSort::SortPairs", st);
    B200_CUDA(cub::DeviceRadixSort::SortPairs(ws.cub_tmp.p, tmp, ws.keys.as<unsigned>(),
                                              ws.keys_sorted.as<unsigned>(), ws.vals.as<unsigned>(),
                                              ws.perm.as<unsigned>(), (int)n, 0, bits, st));
    if (tl_prof) tl_prof->end(st);
    g_launches.fetch_add(3, std::memory_order_relaxed);
  }
  B200_LAUNCH(shard_offsets_kernel, 1, 64, 0, st, n, world, ws.keys_sorted.as<unsigned>(),
              ws.offsets.as<int>());
  if (n > 0)
    B200_LAUNCH(shard_place_kernel, cdiv(n, 256), 256, 0, st, n, world, cap, feats,
                ws.keys_sorted.as<unsigned>(), ws.perm.as<unsigned>(), ws.offsets.as<int>(), send_ids,
                dst, overflow);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

}  // namespace b200rec
