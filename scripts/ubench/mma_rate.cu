// Micro-benchmark: cycles per tcgen05.mma (cta_group::1, M=128) for kind::tf32 (K=8) and kind::f16/bf16 (K=16)
// at several N, operands in shared memory (SWIZZLE_128B K-major tiles, arbitrary contents).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I recommendation-models_b200/csrc -I include \
//        scripts/ubench/mma_rate.cu -o scripts/ubench/mma_rate && ./scripts/ubench/mma_rate
#include <cstdio>
#include "tc_gemm.cuh"
using namespace b200rec::tc;

__global__ void __launch_bounds__(128, 1) rate_kernel(int bn, int reps, int mode, long long* out) {
  extern __shared__ char smem_raw[];
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  // fill operands with something finite
  for (int i = threadIdx.x; i < (16384 + 32768) * 2 / 4; i += blockDim.x) reinterpret_cast<float*>(base)[i] = 0.001f * (i & 255);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tmem_slot), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    char* a0 = base; char* a1 = base + 16384; char* b0 = base + 32768; char* b1 = b0 + 32768;
    const uint64_t da = make_desc(smem_u32(a0)), dac = make_desc(smem_u32(a1));
    const uint64_t db = make_desc(smem_u32(b0)), dbc = make_desc(smem_u32(b1));
    const uint32_t id_t = make_idesc(bn), id_b = make_idesc_bf16(bn);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t adv = (uint64_t)(ks * 32 >> 4);
        if (mode == 0 || mode == 2) mma_tf32(tmem, da + adv, db + adv, id_t, 1u);
        if (mode == 3) { mma_tf32(tmem, dac + adv, db + adv, id_t, 1u); mma_tf32(tmem, da + adv, dbc + adv, id_t, 1u); mma_tf32(tmem, da + adv, db + adv, id_t, 1u); }
        if (mode == 1 || mode == 2) mma_bf16(tmem, dac + adv, dbc + adv, id_b, 1u);
      }
    }
    const long long t1 = clock64();
    mma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024 + 2048);
  const char* names[] = {"tf32 x4", "bf16 x4", "tf32 x4 + bf16 x4", "tf32 x12 (3xTF32)"};
  const int per[] = {4, 4, 8, 12};
  for (int grid : {1, 148})
    for (int bn : {64, 128, 208, 256})
      for (int mode = 0; mode < 4; ++mode) {
        const int reps = 200;
        rate_kernel<<<grid, 128, 100 * 1024 + 2048>>>(bn, reps, mode, d);
        long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaDeviceSynchronize();
        printf("grid %3d N %3d %-20s issue %7.1f cyc/MMA  complete %7.1f cyc/MMA  (%6.1f cyc per K-block of 32) %s\n", grid, bn,
               names[mode], (double)h[0] / (reps * per[mode]), (double)h[1] / (reps * per[mode]), (double)h[1] / reps,
               e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  return 0;
}
