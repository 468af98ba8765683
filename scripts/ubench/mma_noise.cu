// Micro-benchmark: does CUDA-core work on the same SM slow tcgen05.mma down?  One thread issues the MMAs of a
// TF32 + bf16 K-block (4 + 4, M=128, N=208, operands in shared memory) 400 times; NW "noise" warps meanwhile run
//   mode 0: nothing (they wait)          mode 1: integer ALU chains (the producers' split arithmetic)
//   mode 2: ALU + 128-bit shared stores  mode 3: ALU + L1-hitting global loads      mode 4: shared stores only
//   mode 5: four independent integer chains per thread (issue-bound)
// placed on all four schedulers or only on schedulers 1-3 (the MMA warp is warp 0 = scheduler 0).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I recommendation-models_b200/csrc -I include \
//        scripts/ubench/mma_noise.cu -o scripts/ubench/mma_noise && ./scripts/ubench/mma_noise
#include <cstdio>
#include "tc_gemm.cuh"
using namespace b200rec::tc;

__global__ void __launch_bounds__(1024, 1) noise_kernel(int bn, int reps, int mode, int skip_sched0, const float* g, long long* out,
                                                        unsigned* sink, int commit_each = 0) {
  extern __shared__ char smem_raw[];
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint64_t ring[6];   // commit_each: one tcgen05.commit per K-block on a ring of barriers nobody waits for
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop;
  for (int i = threadIdx.x; i < (16384 + 32768) * 2 / 4; i += blockDim.x) reinterpret_cast<float*>(base)[i] = 0.001f * (i & 255);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    for (int i = 0; i < 6; ++i) mbar_init(smem_u32(&ring[i]), 1);
    fence_mbar_init();
    stop = 0;
  }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tmem_slot), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    char* a0 = base; char* a1 = base + 16384; char* b0 = base + 32768; char* b1 = b0 + 32768;
    const uint64_t da = make_desc(smem_u32(a0)), dac = make_desc(smem_u32(a1));
    const uint64_t db = make_desc(smem_u32(b0)), dbc = make_desc(smem_u32(b1));
    const uint32_t id_t = make_idesc(bn), id_b = make_idesc_bf16(bn);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t adv = (uint64_t)(ks * 32 >> 4);
        mma_tf32(tmem, da + adv, db + adv, id_t, 1u);
      }
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t adv = (uint64_t)(ks * 32 >> 4);
        mma_bf16(tmem, dac + adv, dbc + adv, id_b, 1u);
      }
      if (commit_each) mma_commit(smem_u32(&ring[r % 6]));
    }
    mma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) out[0] = t2 - t0;
    stop = 1;
  } else if (warp >= 1 && mode > 0 && !(skip_sched0 && (warp & 3) == 0)) {
    // noise: scratch region beyond the operands (never read by the MMAs)
    char* scratch = base + 2 * (16384 + 32768);
    unsigned x = threadIdx.x * 2654435761u, y = x ^ 0x9e3779b9u, z = x + 12345u, w = ~x;
    const float* gp = g + (threadIdx.x & 1023) * 4;
    long long iters = 0;
    while (!stop) {
#pragma unroll 8
      for (int i = 0; i < 32; ++i) {
        if (mode != 4 && mode != 5) {
          x = (x + 0x1000u) & 0xffffe000u; y = __byte_perm(y + 0x8000u, x, 0x7632); z = (z + x) ^ y; w = (w + 0x8000u) & z;
          x += w; y += z;
        }
        if (mode == 5) {   // four INDEPENDENT chains: issue-bound noise (the modes above are latency-bound chains)
          x = (x + 0x1000u) & 0xffffe000u; y = (y + 0x8000u) ^ 0x7632u; z = (z + 0x1001u) & 0xfffff000u; w = (w + 0x8001u) ^ 0x1234u;
          x = (x + 0x1000u) & 0xffffe000u; y = (y + 0x8000u) ^ 0x7632u; z = (z + 0x1001u) & 0xfffff000u; w = (w + 0x8001u) ^ 0x1234u;
        }
        if (mode == 2 || mode == 4) *reinterpret_cast<uint4*>(scratch + ((threadIdx.x * 16 + i * 16384) & 32767)) = make_uint4(x, y, z, w);
        if (mode == 3) { const float4 v = __ldg(reinterpret_cast<const float4*>(gp + ((i & 7) << 12))); x ^= __float_as_uint(v.x); }
      }
      ++iters;
    }
    sink[threadIdx.x] = x ^ y ^ z ^ w ^ (unsigned)iters;
    if (threadIdx.x == 32 && blockIdx.x == 0) out[1] = iters;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  unsigned* sink; cudaMalloc(&sink, 4096 * 4);
  float* g; cudaMalloc(&g, 1 << 20); cudaMemset(g, 0, 1 << 20);
  const int smem = 2 * (16384 + 32768) + 32768 + 2048;
  cudaFuncSetAttribute(noise_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[] = {"idle", "ALU chains", "ALU + STS.128", "ALU + LDG(L1 hit)", "STS.128 only", "ALU 4 indep. chains"};
  const int reps = 400, bn = 208;
  for (int grid : {148})
    for (int nw : {8, 16})
      for (int skip = 0; skip < 2; ++skip)
        for (int mode = 0; mode < 6; ++mode) {
          if (mode == 0 && skip) continue;
          cudaMemset(d, 0, 16);
          noise_kernel<<<grid, 32 * (1 + nw + (skip ? nw / 3 + 1 : 0)), smem>>>(bn, reps, mode, skip, g, d, sink);
          long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          cudaError_t e = cudaDeviceSynchronize();
          printf("grid %3d  noise warps %2d %-22s %-18s %7.1f cycles per K-block (8 MMAs)   noise iters %lld %s\n", grid, nw,
                 skip ? "(not on scheduler 0)" : "(all schedulers)", names[mode], (double)h[0] / reps, h[1],
                 e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
  // does a tcgen05.commit per K-block (what the pipeline needs to free its ring slots) cost tensor time?
  for (int ce = 0; ce < 2; ++ce) {
    cudaMemset(d, 0, 16);
    noise_kernel<<<148, 32 * 9, smem>>>(bn, reps, 0, 0, g, d, sink, ce);
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaDeviceSynchronize();
    printf("commit per K-block: %d   %7.1f cycles per K-block (8 MMAs) %s\n", ce, (double)h[0] / reps, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
