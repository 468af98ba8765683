// Shared helpers for libb200rec (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <atomic>

#include "../../include/b200rec.h"

namespace b200rec {

// ---- error plumbing: nothing throws across the ABI -----------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

#define B200_CUDA(expr)                                                                  \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      ::b200rec::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                 \
                           cudaGetErrorString(_e));                                      \
      return B200REC_ERR_CUDA;                                                           \
    }                                                                                    \
  } while (0)

#define B200_TRY(expr)                 \
  do {                                 \
    int _s = (expr);                   \
    if (_s != B200REC_OK) return _s;   \
  } while (0)

#define B200_REQUIRE(cond, code, ...)      \
  do {                                     \
    if (!(cond)) {                         \
      ::b200rec::set_error(__VA_ARGS__);   \
      return (code);                       \
    }                                      \
  } while (0)

// ---- optional per-kernel timing (b200rec_profile_begin / _end): CUDA events on the launching
// stream around every launch of the calling thread.  Off by default; bench.py turns it on for a
// separate pass to get per-kernel durations for the roofline lines.
struct Prof {
  struct Rec { const char* tag; const char* name; cudaEvent_t a, b; };
  static constexpr int kMax = 1 << 15;
  Rec* recs = nullptr;
  int n = 0;
  void begin(const char* name, cudaStream_t st);
  void end(cudaStream_t st);
};
extern thread_local Prof* tl_prof;
extern thread_local const char* tl_tag;
struct DevBuf;
// weight-pack scratch of the handle / call that is currently running on this thread (tc_gemm.cu)
extern thread_local DevBuf* tl_pack;
// weights of the MLP tower packed ahead of time in two launches (tc_gemm.cu: tc_prepack_linear);
// the GEMM launchers look their operand up here before packing it themselves
struct PrePack {
  struct Entry { const float* w; int fmt, N, K; size_t off; };   // fmt 0: forward, 1: gradInput
  Entry e[16];
  int n = 0;
  DevBuf* blob = nullptr;
};
extern thread_local PrePack* tl_prepack;
struct PackScope {
  DevBuf* prev;
  explicit PackScope(DevBuf* b) : prev(tl_pack) { tl_pack = b; }
  ~PackScope() { tl_pack = prev; }
};
struct ProfTag {  // names the phase the following launches belong to
  const char* prev;
  explicit ProfTag(const char* t) : prev(tl_tag) { tl_tag = t; }
  ~ProfTag() { tl_tag = prev; }
};

// Programmatic dependent launch (PDL).  A step is ~45 dependent launches; between two of them the GPU pays the
// launch latency of the second and, for the GEMMs, its prologue (barrier init, TMEM allocation) with the SMs idle.
// Every kernel of the library therefore starts with B200_PDL_ENTRY(): `griddepcontrol.launch_dependents` lets the
// NEXT kernel of the stream be scheduled while this one runs, `griddepcontrol.wait` then holds this kernel until the
// PREVIOUS one has completed and its writes are visible -- nothing of global memory is touched before it, so the
// data dependencies are exactly those of plain stream order.  The launches carry the attribute
// cudaLaunchAttributeProgrammaticStreamSerialization (B200REC_PDL=0 turns it off); both instructions are no-ops in
// a kernel launched without it.  (The GEMMs wait AFTER their prologue: tc_gemm.cuh.)
// Measured on one B200 with the attribute on EVERY launch: DeepFM -1.3 %, but FM +19 % and xDeepFM +3 % slower (blocks of
// the next kernel, parked at their wait, take warp slots and registers from the running one; multi-wave kernels trigger
// late).  So only launches inside a PdlHint scope carry it: the single-wave GEMMs of the MLP tower.
extern int g_pdl;   // -1: not read yet
int pdl_enabled();  // B200REC_PDL: 0 off, 1 every launch, 2 (default) launches marked by PdlHint, 3 = 2 + multi-wave GEMMs
extern thread_local int tl_pdl_hint;
struct PdlHint {    // the launches inside the scope carry the attribute in mode 2 / 3
  int prev;
  explicit PdlHint(int on) : prev(tl_pdl_hint) { tl_pdl_hint = on; }
  ~PdlHint() { tl_pdl_hint = prev; }
};
#ifdef __CUDACC__
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#define B200_PDL_ENTRY()          \
  do {                            \
    ::b200rec::griddep_launch();  \
    ::b200rec::griddep_wait();    \
  } while (0)

template <class... KArgs, class... Args>
static inline void launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  const int mode = pdl_enabled();
  // (also marking every grid of <= 64 blocks measured within noise, every grid of <= 1100 blocks +3 % slower)
  cfg.numAttrs = (mode == 1 || (mode >= 2 && tl_pdl_hint)) ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// Every kernel launch of this library goes through LAUNCH so b200rec_launch_count is honest.
#define B200_LAUNCH(kernel, grid, block, smem, stream, ...)                 \
  do {                                                                      \
    ::b200rec::Prof* _prof = ::b200rec::tl_prof;                            \
    if (_prof) _prof->begin(#kernel, (stream));                             \
    ::b200rec::launch_kernel(kernel, (grid), (block), (smem), (stream), __VA_ARGS__); \
    if (_prof) _prof->end((stream));                                        \
    ::b200rec::g_launches.fetch_add(1, std::memory_order_relaxed);          \
  } while (0)

#define B200_LAUNCH_NAMED(name, kernel, grid, block, smem, stream, ...)     \
  do {                                                                      \
    ::b200rec::Prof* _prof = ::b200rec::tl_prof;                            \
    if (_prof) _prof->begin((name), (stream));                              \
    ::b200rec::launch_kernel(kernel, (grid), (block), (smem), (stream), __VA_ARGS__); \
    if (_prof) _prof->end((stream));                                        \
    ::b200rec::g_launches.fetch_add(1, std::memory_order_relaxed);          \
  } while (0)

#define B200_CHECK_LAUNCH() B200_CUDA(cudaGetLastError())

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- device buffer that grows on demand ---------------------------------------------------------
// g_alloc_epoch counts the times a buffer that already held memory was FREED (grown or released).
// A captured CUDA graph bakes device pointers in: every graph remembers the epoch of its capture and
// is re-captured (never replayed) once the epoch has moved, so a replay cannot touch freed memory.
extern std::atomic<long long> g_alloc_epoch;
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return B200REC_OK;
    if (p) {
      cudaFree(p);
      g_alloc_epoch.fetch_add(1, std::memory_order_relaxed);
    }
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
      return B200REC_ERR_NOMEM;
    }
    cap = want;
    return B200REC_OK;
  }
  void release() {
    if (p) {
      cudaFree(p);
      g_alloc_epoch.fetch_add(1, std::memory_order_relaxed);
    }
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

// ---- device helpers ----------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ float4 ldg_f4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
// streaming (read-once) 128-bit load: do not pollute L1
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_f4(float* p, const float4& v) {
  *reinterpret_cast<float4*>(p) = v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif

}  // namespace b200rec
