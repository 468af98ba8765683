// tcgen05 (5th-gen tensor core) GEMM for sm_100a with SOFTWARE operand producers and fp32-class
// accuracy through an error-compensated 3xTF32 split.
//
//   C(m,n) = sum_k A(m,k) * B(n,k)        A: [M,K], B: [N,K]; fp32 in, fp32 accumulate (TMEM)
//
// Why software producers: the dominant contraction of the path is the CIN layer
// (rec/model/xdeepfm/CINEncoder.scala:150-157: MM(transB) builds Z[r,(i,j)] = x0[r,i]*x[r,j], then
// Linear(F*H -> C)).  Z is R x F*H floats (4.1 GB at H=200) -- it must never exist in HBM.  Here the
// A tile of the GEMM is GENERATED in shared memory by the CTA's threads (one multiply per element),
// directly in the canonical K-major SWIZZLE_128B layout tcgen05.mma reads through its smem
// descriptor, and the 1x1 compression runs on the tensor cores with the accumulator in TMEM.
//
// Why 3xTF32: the parity bar is 1e-5 relative in fp32 (BigDL's Linear/MM are MKL sgemm).  A single
// TF32 pass carries 10 mantissa bits (about 1e-3).  Each operand is split as x = hi + lo with
// hi = x & 0xffffe000 (exactly what the tensor core keeps of an fp32 word) and lo = x - hi (exact),
// and D += Ahi*Bhi + Alo*Bhi + Ahi*Blo: the dropped lo*lo term is 2^-22 relative.  The producers
// write both parts, so the split costs two ALU ops per element and no extra memory pass.
//
// Structure (one CTA = one 128 x BN output tile, BN <= 256, 256 threads, 1 CTA / SM):
//   * 2-stage ring of {A_hi, A_lo, B_hi, B_lo} tiles, 128 B (32 fp32) of K per row per stage
//   * all 8 warps produce stage s (global -> registers one stage ahead -> split -> swizzled STS),
//     fence.proxy.async, __syncthreads; one elected thread issues up to 4 K-steps x 3 passes of
//     tcgen05.mma.cta_group::1.kind::tf32 and tcgen05.commit's the stage's "empty" mbarrier;
//     production of stage s+1 overlaps the MMAs of stage s (the tensor core runs asynchronously)
//   * epilogue: tcgen05.ld 32x32b.x16 (thread = accumulator row), functor, vectorised stores.
#pragma once
#include "common.cuh"

namespace b200rec {
namespace tc {

constexpr int BM = 128;        // UMMA M, cta_group::1
constexpr int BK = 32;         // fp32 per K-block: one 128-byte swizzle row
constexpr int UK = 8;          // K of one tcgen05.mma.kind::tf32
constexpr int STAGES = 2;
constexpr int THREADS = 256;
constexpr int MAX_BN = 256;
constexpr int TMEM_COLS = 256;
constexpr int A_TILE_BYTES = BM * 128;        // 16 KB
constexpr int B_TILE_BYTES = MAX_BN * 128;    // 32 KB
constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;  // 96 KB
constexpr int EXTRA_BYTES = 24 * 1024;        // producer scratch (CIN keeps its x0 tile here)
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EXTRA_BYTES + 1024 /*align*/ + 64 /*barriers*/;

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {  // same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols)
               : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], TF32 inputs, fp32 accumulate
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(addr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (unused for swizzled K-major: 1)
//   [32,46) stride byte offset >> 4 = 1024 B between 8-row groups | [46,48) version = 1
//   [61,64) layout type = 2 (SWIZZLE_128B).  Tiles are 1024-B aligned, so base_offset = 0.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// cute::UMMA::InstrDescriptor: c_format F32 (1) @4, a/b_format TF32 (2) @7/@10, K-major both,
// N>>3 @17, M>>4 @24.
__device__ __forceinline__ uint32_t make_idesc(int bn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// byte offset of 16-byte chunk c (0..7) of row r inside a [rows x 128 B] SWIZZLE_128B K-major tile
__device__ __forceinline__ uint32_t swz(int r, int c) {
  return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4));
}
// x = hi + lo, hi = the 19 bits the tensor core keeps; both stored as fp32 words
__device__ __forceinline__ void split_store(char* hi, char* lo, uint32_t off, float4 v) {
  float4 h, l;
  h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
  h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
  h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
  h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
  l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
  *reinterpret_cast<float4*>(hi + off) = h;
  *reinterpret_cast<float4*>(lo + off) = l;
}

// ---- K schedules: which 32-wide window of the contraction each stage covers ---------------------
struct KPlain {  // kb-th block of [k_begin, k_end)
  int k_begin, k_end;
  __device__ __forceinline__ int nkb() const { return (k_end - k_begin + BK - 1) / BK; }
  __device__ __forceinline__ int k0(int kb) const { return k_begin + kb * BK; }
  __device__ __forceinline__ int kvalid(int kb) const { return min(BK, k_end - k0(kb)); }
};
// CIN: the contraction index is (i, j), i over F fields, j over H columns; stages walk j-blocks of 32
// in the outer loop and i in the inner loop so the x segment stays in registers across all i.
struct KCin {
  int F, H;
  __device__ __forceinline__ int njb() const { return (H + BK - 1) / BK; }
  __device__ __forceinline__ int nkb() const { return F * njb(); }
  __device__ __forceinline__ int fi(int kb) const { return kb % F; }
  __device__ __forceinline__ int jb(int kb) const { return kb / F; }
  __device__ __forceinline__ int kvalid(int kb) const { return min(BK, H - jb(kb) * BK); }
};

// ---- producers --------------------------------------------------------------------------------------
// A producer fills a [ROWS x 32] tile (ROWS = 128 for A, BN for B) of one stage.  Interface:
//   init(extra_smem, row0, tid)  once;  prefetch(kb) global -> registers;  store(kb, hi, lo).
// Every store() writes all 8 chunks of every row of the tile (zeros beyond kvalid) so that partially
// valid K-steps never multiply stale shared memory.

// value(r, k) = p[r * ld + koff(kb) + kk]   (row-major, contraction contiguous)
template <int MAXT, class Sched>
struct RowMajorProd {
  const float* p; long long ld; int rows_total; int tile_rows; Sched s; bool vec;
  int row0, tid;
  float4 reg[MAXT];
  __device__ __forceinline__ long long koff(int kb) const;
  __device__ __forceinline__ void init(char*, int r0, int t) { row0 = r0; tid = t; }
  __device__ __forceinline__ void prefetch(int kb) {
    const int kv = s.kvalid(kb);
    const long long ko = koff(kb);
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
      const int q = tid + t * THREADS;
      const int r = q >> 3, c = q & 7;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < tile_rows && row0 + r < rows_total && 4 * c < kv) {
        const float* src = p + (long long)(row0 + r) * ld + ko + 4 * c;
        if (vec && 4 * c + 3 < kv) {
          v = __ldg(reinterpret_cast<const float4*>(src));
        } else {
          v.x = __ldg(src);
          if (4 * c + 1 < kv) v.y = __ldg(src + 1);
          if (4 * c + 2 < kv) v.z = __ldg(src + 2);
          if (4 * c + 3 < kv) v.w = __ldg(src + 3);
        }
      }
      reg[t] = v;
    }
  }
  __device__ __forceinline__ void store(int, char* hi, char* lo) {
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
      const int q = tid + t * THREADS;
      const int r = q >> 3, c = q & 7;
      if (r < tile_rows) split_store(hi, lo, swz(r, c), reg[t]);
    }
  }
};
template <int MAXT>
struct RowMajorPlain : RowMajorProd<MAXT, KPlain> {};
template <>
template <>
__device__ __forceinline__ long long RowMajorProd<4, KPlain>::koff(int kb) const { return s.k0(kb); }
template <>
template <>
__device__ __forceinline__ long long RowMajorProd<8, KPlain>::koff(int kb) const { return s.k0(kb); }
template <>
template <>
__device__ __forceinline__ long long RowMajorProd<8, KCin>::koff(int kb) const {
  return (long long)s.fi(kb) * s.H + s.jb(kb) * BK;   // W[c, i*H + j]
}
template <>
template <>
__device__ __forceinline__ long long RowMajorProd<4, KCin>::koff(int kb) const {
  return (long long)s.fi(kb) * s.H + s.jb(kb) * BK;
}

}  // namespace tc
}  // namespace b200rec
