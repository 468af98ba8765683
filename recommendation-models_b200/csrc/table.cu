// Device-resident parameter table (the stand-in for the Angel PS matrices) -- initialisation.
//
// Reference: rec/model/ParRecModel.scala:66-69,95-105 creates the `embedding` PS matrix
// ((K*slots) x inputDim, dimension-major) and fills it with Angel's XavierUniform PSF (third party,
// parity unpinned).  Here the table is row-major [rows, K] (one 64-B row per id at K=16) and is
// filled from a counter hash, so host (recommendation-models_b200/synth.py: hash_uniform) and device
// produce bit-identical values and a shard can regenerate exactly its slice of a 100M-row table.
#include "kernels.h"

namespace b200rec {

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  uint64_t z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// lo + (hi-lo) * u, u = top 24 bits / 2^24; one fp32 multiply then one fp32 add (no fma) so numpy
// reproduces it exactly.
__device__ __forceinline__ float hash_uniform(uint64_t seed, uint64_t ctr, float lo, float span) {
  const uint64_t h = splitmix64(seed * 0x9E3779B97F4A7C15ull + ctr);
  const float u = (float)(uint32_t)(h >> 40) * (1.0f / 16777216.0f);
  return __fadd_rn(__fmul_rn(u, span), lo);
}

__global__ void __launch_bounds__(256) table_init_kernel(float* table, float* wtable,
                                                         long long rows, int K, uint64_t seed,
                                                         float lo, float span, long long row_offset,
                                                         long long row_stride) {
  B200_PDL_ENTRY();
  const long long n = rows * K;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n;
       t += (long long)gridDim.x * blockDim.x) {
    const long long r = t / K;
    const int k = (int)(t - r * K);
    const uint64_t g = (uint64_t)(row_offset + r * row_stride);
    table[t] = hash_uniform(seed, g * (uint64_t)K + (uint64_t)k, lo, span);
    if (k == 0 && wtable) wtable[r] = hash_uniform(seed + 1, g, lo, span);
  }
}

// local row q of rank g holds global id q*world + ((g - (q*world)/period) mod world)   (shard.cu)
__global__ void __launch_bounds__(256) table_init_sharded_kernel(float* table, float* wtable,
                                                                 long long rows, int K, uint64_t seed,
                                                                 float lo, float span, int rank,
                                                                 int world, long long period) {
  B200_PDL_ENTRY();
  const long long n = rows * K;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n;
       t += (long long)gridDim.x * blockDim.x) {
    const long long q = t / K;
    const int k = (int)(t - q * K);
    uint64_t g;
    if (period < 0) {   // contiguous ranges of -period rows per rank
      g = (uint64_t)((long long)rank * (-period) + q);
    } else {
      const long long base = q * world;
      long long tt = ((long long)rank - (base / period)) % world;
      if (tt < 0) tt += world;
      g = (uint64_t)(base + tt);
    }
    table[t] = hash_uniform(seed, g * (uint64_t)K + (uint64_t)k, lo, span);
    if (k == 0 && wtable) wtable[q] = hash_uniform(seed + 1, g, lo, span);
  }
}

int table_init_uniform_sharded(float* table, float* wtable, long long rows, int K, uint64_t seed,
                               float lo, float hi, int rank, int world, long long period,
                               cudaStream_t st) {
  if (rows <= 0) return B200REC_OK;
  int grid = cdiv(rows * K, 256);
  if (grid > 148 * 32) grid = 148 * 32;
  B200_LAUNCH(table_init_sharded_kernel, grid, 256, 0, st, table, wtable, rows, K, seed, lo, hi - lo,
              rank, world, period);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

int table_init_uniform(float* table, float* wtable, long long rows, int K, uint64_t seed, float lo,
                       float hi, long long row_offset, long long row_stride, cudaStream_t st) {
  if (rows <= 0) return B200REC_OK;
  int grid = cdiv(rows * K, 256);
  if (grid > 148 * 32) grid = 148 * 32;
  B200_LAUNCH(table_init_kernel, grid, 256, 0, st, table, wtable, rows, K, seed, lo, hi - lo,
              row_offset, row_stride);
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

}  // namespace b200rec
