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
constexpr int TMEM_COLS = 512;       // [0,256) chunk accumulator P, [256,512) running sum S
constexpr int TMEM_S = 256;
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
  // the loaded registers are valid only after wait::ld; "+r" ties their uses to the wait
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]),
                 "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]),
                 "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t addr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(addr));
}
__device__ __forceinline__ void tmem_wait_ld2(uint32_t* a, uint32_t* b) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]),
                 "+r"(a[7]), "+r"(a[8]), "+r"(a[9]), "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]),
                 "+r"(a[14]), "+r"(a[15]), "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]),
                 "+r"(b[5]), "+r"(b[6]), "+r"(b[7]), "+r"(b[8]), "+r"(b[9]), "+r"(b[10]), "+r"(b[11]),
                 "+r"(b[12]), "+r"(b[13]), "+r"(b[14]), "+r"(b[15])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_wait_ld1(uint32_t* a) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]),
                 "+r"(a[7]), "+r"(a[8]), "+r"(a[9]), "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]),
                 "+r"(a[14]), "+r"(a[15])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t addr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
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
struct KPlain {  // kb-th block of [k_begin, k_end); split-K over blockIdx.z
  int k_begin, k_end, k_chunk;
  __device__ __forceinline__ KPlain for_split(int z) const {
    KPlain r = *this;
    r.k_begin = k_begin + z * k_chunk;
    r.k_end = min(k_end, r.k_begin + k_chunk);
    return r;
  }
  __device__ __forceinline__ int nkb() const { return (k_end - k_begin + BK - 1) / BK; }
  __device__ __forceinline__ int k0(int kb) const { return k_begin + kb * BK; }
  __device__ __forceinline__ int kvalid(int kb) const { return min(BK, k_end - k0(kb)); }
  // offset of the block along a contiguous contraction dimension / along rows of a [k][n] matrix
  __device__ __forceinline__ long long row_koff(int kb) const { return k0(kb); }
  __device__ __forceinline__ long long col_off(int kb, long long ld) const { return k0(kb) * ld; }
};
// CIN: the contraction index is (i, j): i over F fields, j over H columns of the second factor.
// Stages walk j-blocks of 32 in the OUTER loop and i in the INNER loop, so the producer keeps its
// x[r, j-block] segment in registers across all F fields (x is read once, not F times).
struct KCin {
  int F, H;   // H = width of the second factor (x^{l-1} in the forward, dY in the dx GEMM)
  int Hp;     // dx GEMM only: row stride of field i inside W's columns (= H_{l-1})
  __device__ __forceinline__ KCin for_split(int) const { return *this; }
  __device__ __forceinline__ int njb() const { return (H + BK - 1) / BK; }
  __device__ __forceinline__ int nkb() const { return F * njb(); }
  __device__ __forceinline__ int fi(int kb) const { return kb % F; }
  __device__ __forceinline__ int jb(int kb) const { return kb / F; }
  __device__ __forceinline__ int kvalid(int kb) const { return min(BK, H - jb(kb) * BK); }
  // forward: W[c, i*H + j]  (row-major over the contraction)
  __device__ __forceinline__ long long row_koff(int kb) const { return (long long)fi(kb) * H + jb(kb) * BK; }
  // dx: B(n = j', k = (i, c)) = W[c, i*Hp + j']  ->  [k][n] matrix with row stride ld = F*Hp
  __device__ __forceinline__ long long col_off(int kb, long long ld) const {
    return (long long)jb(kb) * BK * ld + (long long)fi(kb) * Hp;
  }
};

// ---- producers --------------------------------------------------------------------------------------
// A producer fills the [rows x 32] tile of one stage for one operand.  Interface:
//   init(extra_smem, row0, tid) once; prefetch(kb): global -> registers; store(kb, hi, lo): registers
//   -> split -> swizzled shared memory.  store() writes all 8 chunks of every tile row (zeros beyond
//   kvalid) so a partially valid K-step never multiplies stale shared memory.

// value(r, kk) = p[r * ld + sched.row_koff(kb) + kk]     (contraction contiguous in memory)
template <int MAXT, class Sched>
struct RowProd {
  const float* p; long long ld; int rows_total; int tile_rows; bool vec;
  Sched s;
  int row0, tid;
  float4 reg[MAXT];
  __device__ __forceinline__ void init(char*, int r0, int t) { row0 = r0; tid = t; }
  __device__ __forceinline__ void prefetch(int kb) {
    const int kv = s.kvalid(kb);
    const long long ko = s.row_koff(kb);
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

// value(r, kk) = p[sched.col_off(kb, ld) + kk * ld + r]    (tile rows contiguous in memory: the
// producer transposes while it stores).  ROWS = 128 or 256 (power of two >= tile rows).
template <int ROWS, class Sched>
struct ColProd {
  static constexpr int MAXT = ROWS * 8 / THREADS;
  const float* p; long long ld; int rows_total; int tile_rows;
  Sched s;
  int row0, tid;
  float4 reg[MAXT];
  __device__ __forceinline__ void init(char*, int r0, int t) { row0 = r0; tid = t; }
  __device__ __forceinline__ void prefetch(int kb) {
    const int kv = s.kvalid(kb);
    const int r = tid & (ROWS - 1);
    const bool rok = r < tile_rows && row0 + r < rows_total;
    const float* src = p + s.col_off(kb, ld) + row0 + r;
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
      const int c = (tid / ROWS) + t * (THREADS / ROWS);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (rok) {
        if (4 * c + 0 < kv) v.x = __ldg(src + (long long)(4 * c + 0) * ld);
        if (4 * c + 1 < kv) v.y = __ldg(src + (long long)(4 * c + 1) * ld);
        if (4 * c + 2 < kv) v.z = __ldg(src + (long long)(4 * c + 2) * ld);
        if (4 * c + 3 < kv) v.w = __ldg(src + (long long)(4 * c + 3) * ld);
      }
      reg[t] = v;
    }
  }
  __device__ __forceinline__ void store(int, char* hi, char* lo) {
    const int r = tid & (ROWS - 1);
    if (r >= tile_rows) return;
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
      const int c = (tid / ROWS) + t * (THREADS / ROWS);
      split_store(hi, lo, swz(r, c), reg[t]);
    }
  }
};

// CIN outer product as the A operand:  A(r, (i, j)) = x0[r, i] * x[r, j]   (CINEncoder.scala:152)
// x0 tile of the CTA's 128 rows lives in the extra shared memory; the x segment of the current
// j-block lives in registers across the F inner stages.
struct CinZProd {
  const float* x0; const float* x; int R; bool vec;   // vec: H % 4 == 0 (16-B aligned x rows)
  KCin s;
  int row0, tid;
  float* x0s;
  float4 xr[4];
  float x0v[4];
  static constexpr int MAX_F = EXTRA_BYTES / (BM * 4);
  __device__ __forceinline__ void init(char* extra, int r0, int t) {
    row0 = r0; tid = t;
    x0s = reinterpret_cast<float*>(extra);
    for (int idx = tid; idx < BM * s.F; idx += THREADS) {
      const int r = idx / s.F, i = idx - r * s.F;
      x0s[idx] = (row0 + r < R) ? __ldg(x0 + (long long)(row0 + r) * s.F + i) : 0.f;
    }
  }
  __device__ __forceinline__ void prefetch(int kb) {
    const int i = s.fi(kb);
    if (i == 0) {
      const int kv = s.kvalid(kb);
      const int j0 = s.jb(kb) * BK;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int q = tid + t * THREADS;
        const int r = q >> 3, c = q & 7;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < R && 4 * c < kv) {
          const float* src = x + (long long)(row0 + r) * s.H + j0 + 4 * c;
          if (vec && 4 * c + 3 < kv) {
            v = __ldg(reinterpret_cast<const float4*>(src));
          } else {
            v.x = __ldg(src);
            if (4 * c + 1 < kv) v.y = __ldg(src + 1);
            if (4 * c + 2 < kv) v.z = __ldg(src + 2);
            if (4 * c + 3 < kv) v.w = __ldg(src + 3);
          }
        }
        xr[t] = v;
      }
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) x0v[t] = x0s[((tid >> 3) + 32 * t) * s.F + i];
  }
  __device__ __forceinline__ void store(int, char* hi, char* lo) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int q = tid + t * THREADS;
      const int r = q >> 3, c = q & 7;
      const float a = x0v[t];
      split_store(hi, lo, swz(r, c), make_float4(a * xr[t].x, a * xr[t].y, a * xr[t].z, a * xr[t].w));
    }
  }
};

// Z^T as the A operand of the weight-gradient GEMM:  A(m = (i, j), k = r) = x0[r, i] * x[r, j]
struct CinZtProd {
  const float* x0; const float* x; int F, H;
  KPlain s;
  int tid, im, jm;
  bool valid;
  float4 reg[4];
  __device__ __forceinline__ void init(char*, int r0, int t) {
    tid = t;
    const int m = r0 + (t & 127);
    valid = m < F * H;
    im = valid ? m / H : 0;
    jm = valid ? m - im * H : 0;
  }
  __device__ __forceinline__ void prefetch(int kb) {
    const int k0 = s.k0(kb);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int c = (tid >> 7) + 2 * t;
      float e[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long r = k0 + 4 * c + u;
        e[u] = (valid && r < s.k_end) ? __ldg(x0 + r * F + im) * __ldg(x + r * H + jm) : 0.f;
      }
      reg[t] = make_float4(e[0], e[1], e[2], e[3]);
    }
  }
  __device__ __forceinline__ void store(int, char* hi, char* lo) {
#pragma unroll
    for (int t = 0; t < 4; ++t) split_store(hi, lo, swz(tid & 127, (tid >> 7) + 2 * t), reg[t]);
  }
};

// ---- epilogues: ep(m, n0, v[16], nv, z) for one accumulator row chunk ------------------------------
struct EpBiasAct {  // y = act(acc + bias[n])
  float* y; long long ld; const float* bias; bool relu;
  static constexpr bool kRowReduce = false;
  __device__ __forceinline__ void operator()(int m, int n0, float* v, int nv, int) const {
    float* dst = y + (long long)m * ld + n0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (i < nv) {
        float t = v[i] + (bias ? __ldg(bias + n0 + i) : 0.f);
        v[i] = relu ? fmaxf(t, 0.f) : t;
      }
    }
    if (nv == 16 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(dst + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i < nv) dst[i] = v[i];
    }
  }
};
struct EpMaskAcc {  // g = acc * (mask > 0) (+ g)
  float* g; long long ld; const float* mask; long long ldm; bool accumulate;
  static constexpr bool kRowReduce = false;
  __device__ __forceinline__ void operator()(int m, int n0, float* v, int nv, int) const {
    float* dst = g + (long long)m * ld + n0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (i < nv) {
        float t = v[i];
        if (mask && !(__ldg(mask + (long long)m * ldm + n0 + i) > 0.f)) t = 0.f;
        if (accumulate) t += dst[i];
        dst[i] = t;
      }
    }
  }
};
struct EpPartial {  // split-K partial: ws[z][m][n]
  float* ws; long long MN; long long ld;
  static constexpr bool kRowReduce = false;
  __device__ __forceinline__ void operator()(int m, int n0, float* v, int nv, int z) const {
    float* dst = ws + (long long)z * MN + (long long)m * ld + n0;
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < nv) dst[i] = v[i];
  }
};
struct EpPartialT {  // split-K partial stored transposed: ws[z][n][m]  (lanes = consecutive m: coalesced)
  float* ws; long long MN; long long ldt;
  static constexpr bool kRowReduce = false;
  __device__ __forceinline__ void operator()(int m, int n0, float* v, int nv, int z) const {
    float* dst = ws + (long long)z * MN + (long long)n0 * ldt + m;
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < nv) dst[(long long)i * ldt] = v[i];
  }
};
struct EpAddBiasRelu2 {  // PNN: h = relu(prev + acc + c0)   (ProductEncoder.scala:97-108)
  float* h; long long ld; const float* prev; const float* c0;
  static constexpr bool kRowReduce = false;
  __device__ __forceinline__ void operator()(int m, int n0, float* v, int nv, int) const {
    const float c = __ldg(c0);
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < nv) {
        const long long o = (long long)m * ld + n0 + i;
        h[o] = fmaxf((prev[o] + v[i]) + c, 0.f);
      }
  }
};
// CIN input gradient w.r.t. x0:  out[m, tile] += sum_n acc[m, n] * x[m, n]   (one N tile per field)
struct EpRowDot {
  const float* x; long long ldx; float* out; long long ldo;
  static constexpr bool kRowReduce = true;
  __device__ __forceinline__ float dot(int m, int nl0, const float* v, int nv) const {
    float s = 0.f;
    const float* xr = x + (long long)m * ldx + nl0;
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < nv) s = fmaf(v[i], __ldg(xr + i), s);
    return s;
  }
  __device__ __forceinline__ void finish(int m, int tile, float s) const { out[(long long)m * ldo + tile] += s; }
};


// ---- the kernel ---------------------------------------------------------------------------------------
// grid = (n_tiles, m_tiles, splits).  Tile n covers columns [n * n_stride, n * n_stride + n_valid).
template <class AP, class BP, class Sched, class Ep, int PASSES>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_kernel(int M, int N, int bn, int n_stride, int n_valid, int kc, Sched sched, AP ap, BP bp,
               Ep ep) {
  // kc > 0: the tensor core accumulates at most kc K-blocks (kc * 32 of K) into the TMEM chunk
  // accumulator P; every chunk is then added into the running sum S (TMEM columns 256..) by the
  // CUDA cores with round-to-nearest.  tcgen05.mma's own accumulation truncates toward zero, so the
  // error of one long TMEM accumulation grows linearly with K (5.7e-5 at K = 8192, measured); chunked
  // it stays at the fp32-blocked-sum level.  kc = 0: plain single accumulation.
  extern __shared__ char smem_raw[];
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  char* extra = base + STAGES * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(extra + EXTRA_BYTES);   // stage[0], stage[1] drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * BM;
  const int n0 = blockIdx.x * n_stride;
  const Sched s = sched.for_split(blockIdx.z);
  ap.s = s;
  bp.s = s;

  if (tid == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  ap.init(extra, m0, tid);
  bp.init(extra, n0, tid);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t idesc = make_idesc(bn);
  const int nkb = s.nkb();

  if (nkb > 0) {
    ap.prefetch(0);
    bp.prefetch(0);
  }
  for (int kb = 0; kb < nkb; ++kb) {
    const int st = kb & 1;
    char* a_hi = base + st * STAGE_BYTES;
    char* a_lo = a_hi + A_TILE_BYTES;
    char* b_hi = a_lo + A_TILE_BYTES;
    char* b_lo = b_hi + B_TILE_BYTES;
    // the MMAs that read this stage two iterations ago must have drained
    if (kb >= STAGES) mbar_wait(smem_u32(&bars[st]), ((kb >> 1) - 1) & 1);
    ap.store(kb, a_hi, a_lo);
    bp.store(kb, b_hi, b_lo);
    if (kb + 1 < nkb) {  // next stage's global loads fly while the tensor core works
      ap.prefetch(kb + 1);
      bp.prefetch(kb + 1);
    }
    const bool chunk_start = kc > 0 ? (kb % kc == 0) : (kb == 0);
    if (kc > 0 && chunk_start && kb > 0) {
      // the previous chunk ended with stage kb-1: wait for its MMAs, then S (+)= P
      mbar_wait(smem_u32(&bars[st ^ 1]), ((kb - 1) >> 1) & 1);
      tc_fence_after();
      const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
      for (int ch = warp >> 2; ch * 16 < bn; ch += 2) {
        uint32_t p[16], q[16];
        tmem_ld16_nowait(tmem + lane_addr + ch * 16, p);
        if (kb > kc) {
          tmem_ld16_nowait(tmem + lane_addr + TMEM_S + ch * 16, q);
          tmem_wait_ld2(p, q);
#pragma unroll
          for (int i = 0; i < 16; ++i) p[i] = __float_as_uint(__uint_as_float(p[i]) + __uint_as_float(q[i]));
        } else {
          tmem_wait_ld1(p);
        }
        tmem_st16(tmem + lane_addr + TMEM_S + ch * 16, p);
      }
      tmem_wait_st();
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const int ksteps = (s.kvalid(kb) + UK - 1) / UK;
      const uint64_t dah = make_desc(smem_u32(a_hi)), dal = make_desc(smem_u32(a_lo));
      const uint64_t dbh = make_desc(smem_u32(b_hi)), dbl = make_desc(smem_u32(b_lo));
      for (int ks = 0; ks < ksteps; ++ks) {
        const uint64_t adv = (uint64_t)(ks * UK * 4 >> 4);  // 32 B per K-step inside the swizzle row
        mma_tf32(tmem, dah + adv, dbh + adv, idesc, (!chunk_start || ks != 0) ? 1u : 0u);
        if (PASSES == 3) {
          mma_tf32(tmem, dal + adv, dbh + adv, idesc, 1u);
          mma_tf32(tmem, dah + adv, dbl + adv, idesc, 1u);
        }
      }
      mma_commit(smem_u32(&bars[st]));
    }
  }

  // ---- epilogue: thread = accumulator row (TMEM lane), 16 columns per tcgen05.ld ------------------
  if (nkb > 0) {
    mbar_wait(smem_u32(&bars[(nkb - 1) & 1]), ((nkb - 1) >> 1) & 1);
    tc_fence_after();
    const bool add_s = kc > 0 && nkb > kc;   // result = S + P (the last chunk is still in P)
    const int lane_base = (warp & 3) * 32;
    const int row = m0 + lane_base + lane;
    const int ncols = min(n_valid, N - n0);
    float v[16];
    if constexpr (Ep::kRowReduce) {
      if (warp < 4) {
        float acc = 0.f;
        for (int ch = 0; ch * 16 < ncols; ++ch) {
          tmem_ld16(tmem + ((uint32_t)lane_base << 16) + ch * 16, v);
          if (add_s) {
            float sv[16];
            tmem_ld16(tmem + ((uint32_t)lane_base << 16) + TMEM_S + ch * 16, sv);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += sv[i];
          }
          if (row < M) acc += ep.dot(row, ch * 16, v, min(16, ncols - ch * 16));
        }
        if (row < M) ep.finish(row, blockIdx.x, acc);
      }
    } else {
      for (int ch = warp >> 2; ch * 16 < ncols; ch += 2) {
        tmem_ld16(tmem + ((uint32_t)lane_base << 16) + ch * 16, v);
        if (add_s) {
          float sv[16];
          tmem_ld16(tmem + ((uint32_t)lane_base << 16) + TMEM_S + ch * 16, sv);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += sv[i];
        }
        if (row < M) ep(row, n0 + ch * 16, v, min(16, ncols - ch * 16), blockIdx.z);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace tc
}  // namespace b200rec
