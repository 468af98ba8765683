// Device helpers of the NVLink peer-memory exchange shared by p2p.cu and the fused reduce + push of
// segsum.cu: release / acquire flag accesses, peer-pointer selection, the end-of-kernel signal.
#pragma once
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

// peer pointer i of a by-value kernel parameter array.  A dynamic index would make the compiler copy
// the whole parameter struct to LOCAL memory in every thread (ncu: 4.5 M local store sectors, 77 MB of
// DRAM writes in the gather kernel); a chain of selects on the constant bank does not.
template <class T>
__device__ __forceinline__ T* peer_sel(T* const (&p)[P2P_MAX], int i) {
  T* r = p[0];
#pragma unroll
  for (int k = 1; k < P2P_MAX; ++k)
    if (i == k) r = p[k];
  return r;
}

__device__ __forceinline__ int p2p_step(const P2P& c) { return c.step > 0 ? c.step : *c.step_ptr - c.step; }

// end-of-kernel signal: every thread of every block must call this (convergently).  The block barrier
// orders every thread's stores before thread 0 (CTA scope); thread 0's system-scope fence is cumulative,
// so those stores are visible system-wide before its counter increment and, in the last block, before
// the release stores of the flags.  (One fence per block: a membar.sys per thread made the writer
// kernels several times slower.)
__device__ __forceinline__ void p2p_signal(const P2P& c, int phase) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned last = gridDim.x * gridDim.y - 1;
    if (atomicInc(c.block_counter, last) == last) {
      __threadfence_system();
      const int step = p2p_step(c);
#pragma unroll
      for (int p = 0; p < P2P_MAX; ++p)
        if (p < c.world) st_release_sys(c.flags[p] + phase * c.world + c.rank, step);
    }
  }
}

}  // namespace b200rec
