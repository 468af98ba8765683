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
// Why an error-compensated split: the parity bar is 1e-5 relative in fp32 (BigDL's Linear/MM are MKL
// sgemm).  A single TF32 pass carries 10 mantissa bits (about 1e-3).  Each operand is split as
// x = hi + lo with hi = tf32_rn(x) (11 significant bits: exactly what the tensor core keeps of an fp32
// word) and lo = x - hi (exact, |lo| <= 2^-12 |x|), and
//     D += Ahi*Bhi  (kind::tf32)  +  Ahi*Blo + Alo*Bhi  (kind::f16 on bf16 copies).
// The two correction terms are 2^-12 of the result, so their operands only need bf16's 8 bits (their
// rounding costs 2^-9 * 2^-12 = 2^-21 relative, the level of the dropped lo*lo term): they run as ONE
// bf16 contraction of K = 64 per 32-wide K-block -- the A correction row is [hi | lo] (32 + 32 bf16 =
// one 128-byte swizzle row), the B correction row is [lo | hi] -- at twice the TF32 rate and half its
// shared-memory operand traffic.  Per K-block: 4 TF32 MMAs + 4 BF16 MMAs instead of 12 TF32 MMAs
// (round 1's 3xTF32): the tensor ceiling rises from 1/3 to 1/2 of the TF32 peak at the same accuracy
// (measured 1.5e-6 -> see DESIGN.md).  The producers write both tiles; same shared-memory footprint.
//
// Structure (one CTA = one 128 x BN output tile, BN <= 256, 1 CTA / SM, warp-specialised; see
// gemm_ws_kernel below): a 2-stage ring of {A_hi, A_lo} tiles filled by 8 producer warps
// (global -> registers one or two stages ahead -> split -> swizzled STS, 128 B = 32 fp32 of K per
// row per stage), a B ring that is either filled by the same warps or -- for weight operands --
// fetched as pre-packed stage images with cp.async.bulk, and one MMA thread issuing up to
// 4 K-steps x 3 passes of tcgen05.mma.cta_group::1.kind::tf32 per stage; mbarriers only, no
// __syncthreads in the main loop; epilogue tcgen05.ld 32x32b.x16 (thread = accumulator row).
#pragma once
#include <type_traits>

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

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
  // a pipeline bug must fail the launch, not hang the GPU: trap after ~2 s of waiting
  const long long t0 = clock64();
  while (!mbar_try(bar, parity))
    if (clock64() - t0 > 4000000000LL) __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try(bar, parity)) return;   // fast path: no clock reads
  if (mbar_try(bar, parity)) return;
  mbar_wait_slow(bar, parity);
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// 1-D bulk async copy global -> shared (TMA engine, no tensor map), completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// 1-D bulk async copy shared -> global (TMA engine); completion tracked by the thread's bulk group
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_wait_read() {   // the shared-memory source may be reused / released afterwards
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
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
// D[tmem] (+)= A * B, BF16 inputs (K = 16 per instruction), fp32 accumulate
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// All MMAs of one 32-wide K-block into the accumulator columns at `acc`: `ksteps` TF32 K-steps on the hi
// tiles and, with PASSES == 3, as many bf16 K-steps (16 positions = 8 K values x {hi*lo, lo*hi}) on the
// correction tiles.  boff: descriptor offset of the first B row.
__device__ __forceinline__ void mma_block_tf32(uint32_t acc, uint64_t da_hi, uint64_t db_hi, uint64_t boff,
                                               uint32_t idesc, int ksteps, bool fresh) {
  for (int ks = 0; ks < ksteps; ++ks) {
    const uint64_t adv = (uint64_t)(ks * UK * 4 >> 4);
    mma_tf32(acc, da_hi + adv, db_hi + adv + boff, idesc, (!fresh || ks != 0) ? 1u : 0u);
  }
}
__device__ __forceinline__ void mma_block_bf16(uint32_t acc, uint64_t da_corr, uint64_t db_corr, uint64_t boff,
                                               uint32_t idesc_bf, int ksteps) {
  for (int ks = 0; ks < ksteps; ++ks) {
    const uint64_t adv = (uint64_t)(ks * 32 >> 4);
    mma_bf16(acc, da_corr + adv, db_corr + adv + boff, idesc_bf, 1u);
  }
}
template <int PASSES>
__device__ __forceinline__ void mma_block(uint32_t acc, uint64_t da_hi, uint64_t da_corr, uint64_t db_hi,
                                          uint64_t db_corr, uint64_t boff, uint32_t idesc, uint32_t idesc_bf,
                                          int ksteps, int kvalid, bool fresh) {
  mma_block_tf32(acc, da_hi, db_hi, boff, idesc, ksteps, fresh);
  if (PASSES == 3) mma_block_bf16(acc, da_corr, db_corr, boff, idesc_bf, ksteps);
  (void)kvalid;
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
// the same for kind::f16 with BF16 (1) operands, fp32 accumulate
__device__ __forceinline__ uint32_t make_idesc_bf16(int bn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// byte offset of 16-byte chunk c (0..7) of row r inside a [rows x 128 B] SWIZZLE_128B K-major tile
__device__ __forceinline__ uint32_t swz(int r, int c) {
  return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4));
}
__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ uint32_t bf16x2_rn(float lo_elem, float hi_elem) {   // {hi_elem, lo_elem} packed
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
}
// The 4 consecutive K values v of tile row r, 16-byte chunk c (K = 4c .. 4c+3 of the 32-wide block):
//   hi tile   [rows x 32 tf32]  chunk c  <- tf32_rn(v)
//   corr tile [rows x 64 bf16]  chunk c  <- A operand { bf16(hi) x4 | bf16(lo) x4 },  B operand { lo x4 | hi x4 }
// The correction contraction only needs A and B to pair the same K value at the same position, so the
// hi / lo copies of a K quad sit side by side in ONE 16-byte chunk: one conflict-free 128-bit store per
// tile (a [hi half | lo half] row layout costs two 4-way-conflicted 64-bit stores), and bf16 K-step s
// (16 positions = chunks 2s, 2s+1) covers exactly the K values of TF32 K-step s.
// Both tiles are K-major SWIZZLE_128B.
__device__ __forceinline__ void split_store(char* hi, char* corr, int r, int c, float4 v, bool is_b) {
  const uint32_t off = swz(r, c);
#ifdef B200_SPLIT_CVT
  float4 h;
  h.x = tf32_rn(v.x); h.y = tf32_rn(v.y); h.z = tf32_rn(v.z); h.w = tf32_rn(v.w);
  *reinterpret_cast<float4*>(hi + off) = h;
  const uint32_t h0 = bf16x2_rn(h.x, h.y), h1 = bf16x2_rn(h.z, h.w);
  const uint32_t l0 = bf16x2_rn(v.x - h.x, v.y - h.y), l1 = bf16x2_rn(v.z - h.z, v.w - h.w);
#else
  // The same split with integer ALU ops only (the conversion pipe is narrow and the producers are on the
  // critical path of the two-deep A ring): tf32 = round the 13 low mantissa bits away (add half an ulp,
  // mask), bf16 = add half an ulp to the fp32 word and keep its upper 16 bits (PRMT packs two).
  const uint32_t bx = (__float_as_uint(v.x) + 0x1000u) & 0xffffe000u, by = (__float_as_uint(v.y) + 0x1000u) & 0xffffe000u;
  const uint32_t bz = (__float_as_uint(v.z) + 0x1000u) & 0xffffe000u, bw = (__float_as_uint(v.w) + 0x1000u) & 0xffffe000u;
  *reinterpret_cast<uint4*>(hi + off) = make_uint4(bx, by, bz, bw);
  const float lx = v.x - __uint_as_float(bx), ly = v.y - __uint_as_float(by);
  const float lz = v.z - __uint_as_float(bz), lw = v.w - __uint_as_float(bw);
  const uint32_t h0 = __byte_perm(bx + 0x8000u, by + 0x8000u, 0x7632), h1 = __byte_perm(bz + 0x8000u, bw + 0x8000u, 0x7632);
  const uint32_t l0 = __byte_perm(__float_as_uint(lx) + 0x8000u, __float_as_uint(ly) + 0x8000u, 0x7632);
  const uint32_t l1 = __byte_perm(__float_as_uint(lz) + 0x8000u, __float_as_uint(lw) + 0x8000u, 0x7632);
#endif
  *reinterpret_cast<uint4*>(corr + off) = is_b ? make_uint4(l0, l1, h0, h1) : make_uint4(h0, h1, l0, l1);
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
  int F, H;   // F = length of the first factor, H = width of the second factor
  int Hp;     // dx GEMM only: row stride of field i inside W's columns (= H_{l-1})
  long long KS;  // row-major B: offset of first-factor index i along the contraction (H fwd, F*H dx0)
  unsigned magicF;  // ceil(2^32 / F): kb / F == __umulhi(kb, magicF) while kb * F < 2^32 (host: make_kcin)
  __device__ __forceinline__ KCin for_split(int) const { return *this; }
  __device__ __forceinline__ int njb() const { return (H + BK - 1) / BK; }
  __device__ __forceinline__ int nkb() const { return F * njb(); }
  __device__ __forceinline__ int jb(int kb) const { return (int)__umulhi((unsigned)kb, magicF); }
  __device__ __forceinline__ int fi(int kb) const { return kb - jb(kb) * F; }
  __device__ __forceinline__ int kvalid(int kb) const { return min(BK, H - jb(kb) * BK); }
  // forward: W[c, i*H + j]  (row-major over the contraction)
  __device__ __forceinline__ long long row_koff(int kb) const { return (long long)fi(kb) * KS + jb(kb) * BK; }
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
  bool is_b = false;   // B operand: the correction tile is [lo | hi] instead of [hi | lo] (split_store)
  float4 reg[MAXT];
  float4 reg2[MAXT];   // second prefetch slot (packed-B kernel: loads fly two stages ahead)
  template <int SLOT> __device__ __forceinline__ void prefetch2(int kb) {
    if constexpr (SLOT == 1) load(kb, reg2); else load(kb, reg);
  }
  template <int SLOT> __device__ __forceinline__ void store2(int, char* hi, char* lo) {
    if constexpr (SLOT == 1) put(reg2, hi, lo); else put(reg, hi, lo);
  }
  static constexpr int DIST = 2;
  __device__ __forceinline__ void init(char*, int r0, int t) { row0 = r0; tid = t; }
  __device__ __forceinline__ void prefetch(int kb) { load(kb, reg); }
  __device__ __forceinline__ void store(int, char* hi, char* lo) { put(reg, hi, lo); }
  __device__ __forceinline__ void load(int kb, float4 (&dst)[MAXT]) {
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
      dst[t] = v;
    }
  }
  __device__ __forceinline__ void put(const float4 (&src)[MAXT], char* hi, char* lo) {
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
      const int q = tid + t * THREADS;
      const int r = q >> 3, c = q & 7;
      if (r < tile_rows) split_store(hi, lo, r, c, src[t], is_b);
    }
  }
};

// value(r, kk) = p[sched.col_off(kb, ld) + kk * ld + r]    (tile rows contiguous in memory: the
// producer transposes while it stores).  ROWS = 128 or 256 (power of two >= tile rows).
// vec path (ld % 4 == 0, 16-B aligned, rows_total % 4 == 0): a thread loads a 4 (kk) x 4 (rows) block
// with four 128-bit loads and transposes it in registers into the four rows' 16-B chunks -- 4x fewer
// load instructions than the element-wise path.
template <int ROWS, class Sched>
struct ColProd {
  static constexpr int MAXT = ROWS * 8 / THREADS;   // float4 registers per thread per stage
  static constexpr int NIT = ROWS * 2 / THREADS;    // vec path: 4x4 blocks per thread
  const float* p; long long ld; int rows_total; int tile_rows; bool vec;
  Sched s;
  int row0, tid;
  bool is_b = false;   // B operand: the correction tile is [lo | hi] instead of [hi | lo] (split_store)
  // Row `ones_row` of the operand reads as 1.0 at every valid k (and is otherwise past rows_total): the weight-
  // gradient GEMM gy^T [x | 1] then yields the bias gradient (column sums of gy) as one more output column, and
  // the two column-sum launches per layer disappear (BackwardUtil.linearBackward's gradBias).
  int ones_row = -1;
  float4 reg[MAXT];
  static constexpr int DIST = 1;
  template <int SLOT> __device__ __forceinline__ void prefetch2(int kb) { prefetch(kb); }
  template <int SLOT> __device__ __forceinline__ void store2(int kb, char* hi, char* lo) { store(kb, hi, lo); }
  __device__ __forceinline__ void init(char*, int r0, int t) { row0 = r0; tid = t; }
  __device__ __forceinline__ float4 ones4(int c, int kv) const {
    return make_float4(4 * c + 0 < kv ? 1.f : 0.f, 4 * c + 1 < kv ? 1.f : 0.f, 4 * c + 2 < kv ? 1.f : 0.f,
                       4 * c + 3 < kv ? 1.f : 0.f);
  }
  __device__ __forceinline__ void prefetch(int kb) {
    const int kv = s.kvalid(kb);
    if (vec) {
#pragma unroll
      for (int t = 0; t < NIT; ++t) {
        const int q = tid + t * THREADS;
        const int c = q & 7, rg = q >> 3;
        const int r = 4 * rg;
        const bool rok = r < tile_rows && row0 + r < rows_total;
        const float* src = p + s.col_off(kb, ld) + row0 + r;
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          v[u] = (rok && 4 * c + u < kv) ? __ldg(reinterpret_cast<const float4*>(src + (long long)(4 * c + u) * ld))
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
        reg[4 * t + 0] = make_float4(v[0].x, v[1].x, v[2].x, v[3].x);
        reg[4 * t + 1] = make_float4(v[0].y, v[1].y, v[2].y, v[3].y);
        reg[4 * t + 2] = make_float4(v[0].z, v[1].z, v[2].z, v[3].z);
        reg[4 * t + 3] = make_float4(v[0].w, v[1].w, v[2].w, v[3].w);
        if (ones_row >= 0 && r < tile_rows) {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (row0 + r + e == ones_row) reg[4 * t + e] = ones4(c, kv);
        }
      }
      return;
    }
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
      if (ones_row >= 0 && r < tile_rows && row0 + r == ones_row) v = ones4(c, kv);
      reg[t] = v;
    }
  }
  __device__ __forceinline__ void store(int, char* hi, char* lo) {
    if (vec) {
#pragma unroll
      for (int t = 0; t < NIT; ++t) {
        const int q = tid + t * THREADS;
        const int c = q & 7, r = 4 * (q >> 3);
        if (r < tile_rows) {
#pragma unroll
          for (int e = 0; e < 4; ++e) split_store(hi, lo, r + e, c, reg[4 * t + e], is_b);
        }
      }
      return;
    }
    const int r = tid & (ROWS - 1);
    if (r >= tile_rows) return;
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
      const int c = (tid / ROWS) + t * (THREADS / ROWS);
      split_store(hi, lo, r, c, reg[t], is_b);
    }
  }
};

// CIN outer product as the A operand:  A(r, (i, j)) = x0[r, i] * x[r, j]   (CINEncoder.scala:152)
// x0 tile of the CTA's 128 rows lives in the extra shared memory; the x segment of the current
// j-block lives in registers across the F inner stages.
struct CinZProd {
  const float* x0; const float* x; int R; bool vec;   // vec: H % 4 == 0 (16-B aligned x rows)
  bool x0_in_smem;                                     // false: no scratch smem (packed-B kernel)
  KCin s;
  int row0, tid;
  float* x0s;
  const float* x0row[4];   // the thread's four x0 rows (nullptr beyond R): hoisted out of the per-stage loads
  float4 xr[4];
  float x0v[4];
  static constexpr int DIST = 1;
  template <int SLOT> __device__ __forceinline__ void prefetch2(int kb) { prefetch(kb); }
  template <int SLOT> __device__ __forceinline__ void store2(int kb, char* hi, char* lo) { store(kb, hi, lo); }
  __device__ __forceinline__ void init(char* extra, int r0, int t) {
    row0 = r0; tid = t;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = (t >> 3) + 32 * q;
      x0row[q] = row0 + r < R ? x0 + (long long)(row0 + r) * s.F : nullptr;
    }
    x0s = reinterpret_cast<float*>(extra);
    if (x0_in_smem)
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
    for (int t = 0; t < 4; ++t) {
      const int r = (tid >> 3) + 32 * t;
      x0v[t] = x0_in_smem ? x0s[r * s.F + i] : (x0row[t] ? __ldg(x0row[t] + i) : 0.f);
    }
  }
  __device__ __forceinline__ void store(int, char* hi, char* lo) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int q = tid + t * THREADS;
      const int r = q >> 3, c = q & 7;
      const float a = x0v[t];
      split_store(hi, lo, r, c, make_float4(a * xr[t].x, a * xr[t].y, a * xr[t].z, a * xr[t].w), false);
    }
  }
};

// Z^T as the A operand of the weight-gradient GEMM:  A(m = (i, j), k = r) = x0[r, i] * x[r, j]
// vec path (H % 4 == 0): a thread owns 4 consecutive m (= 4 consecutive j of one field) x 4 rows r:
// four 128-bit loads of x, four scalar loads of x0, a register transpose, four 16-B chunk stores.
struct CinZtProd {
  const float* x0; const float* x; int F, H; bool vec;
  KPlain s;
  int tid, im, jm;
  bool valid;
  float4 reg[4];
  static constexpr int DIST = 1;
  template <int SLOT> __device__ __forceinline__ void prefetch2(int kb) { prefetch(kb); }
  template <int SLOT> __device__ __forceinline__ void store2(int kb, char* hi, char* lo) { store(kb, hi, lo); }
  __device__ __forceinline__ void init(char*, int r0, int t) {
    tid = t;
    const int m = vec ? r0 + 4 * (t >> 3) : r0 + (t & 127);
    valid = m < F * H;
    im = valid ? m / H : 0;
    jm = valid ? m - im * H : 0;
  }
  __device__ __forceinline__ void prefetch(int kb) {
    const int k0 = s.k0(kb);
    if (vec) {
      const int c = tid & 7;
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long r = k0 + 4 * c + u;
        if (valid && r < s.k_end) {
          const float a = __ldg(x0 + r * F + im);
          const float4 xv = __ldg(reinterpret_cast<const float4*>(x + r * H + jm));
          v[u] = make_float4(a * xv.x, a * xv.y, a * xv.z, a * xv.w);
        } else {
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      reg[0] = make_float4(v[0].x, v[1].x, v[2].x, v[3].x);
      reg[1] = make_float4(v[0].y, v[1].y, v[2].y, v[3].y);
      reg[2] = make_float4(v[0].z, v[1].z, v[2].z, v[3].z);
      reg[3] = make_float4(v[0].w, v[1].w, v[2].w, v[3].w);
      return;
    }
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
    if (vec) {
      const int c = tid & 7, r = 4 * (tid >> 3);
#pragma unroll
      for (int e = 0; e < 4; ++e) split_store(hi, lo, r + e, c, reg[e], false);
      return;
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) split_store(hi, lo, tid & 127, (tid >> 7) + 2 * t, reg[t], false);
  }
};

// The same operand from TRANSPOSED copies of the factors (x0T [F x R], xT [H x R], made once per layer by a
// tiled transpose): A(m = (i, j), k = r) = x0T[i, r] * xT[j, r].  The contraction index r is now contiguous in
// memory, so a thread fetches its 4 K values of a tile row with two 128-bit loads and multiplies them
// component-wise -- no scalar x0 loads, no 4x4 register transposes, and the first layer (H = F = 39, rows of x
// not 16-byte aligned) takes the same vector path as the others.
struct CinZtProdT {
  const float* x0T; const float* xT; int F, H; long long ldr;   // ldr = R: row stride of the transposed copies
  KPlain s;
  int tid;
  const float* p0[4];   // x0T row of the (i) of this thread's four tile rows, nullptr beyond F*H
  const float* p1[4];   // xT row of their (j)
  // Two register sets of RAW factors: prefetch only issues loads (nothing in it depends on a loaded value, so it
  // never stalls), the products are formed when the stage is stored two stages later.  With the multiply in the
  // prefetch the L2 latency under load (1500-2000 cycles) was exposed once per stage (profiles/r02s_trace.txt).
  float4 ra[2][4], rb[2][4];
  static constexpr int DIST = 2;
  __device__ __forceinline__ void init(char*, int r0, int t) {
    tid = t;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int m = r0 + (t >> 3) + 32 * q;
      const bool valid = m < F * H;
      const int im = valid ? m / H : 0;
      const int jm = valid ? m - im * H : 0;
      p0[q] = valid ? x0T + (long long)im * ldr : nullptr;
      p1[q] = valid ? xT + (long long)jm * ldr : nullptr;
    }
  }
  template <int SLOT> __device__ __forceinline__ void prefetch2(int kb) {
    const int k = s.k0(kb) + 4 * (tid & 7);          // R % 4 == 0 (host check): a quad is inside [0, k_end) or outside
    const bool in = k < s.k_end;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
      if (in && p0[q]) {
        a = __ldg(reinterpret_cast<const float4*>(p0[q] + k));
        b = __ldg(reinterpret_cast<const float4*>(p1[q] + k));
      }
      ra[SLOT][q] = a;
      rb[SLOT][q] = b;
    }
  }
  template <int SLOT> __device__ __forceinline__ void store2(int, char* hi, char* lo) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 a = ra[SLOT][q], b = rb[SLOT][q];
      split_store(hi, lo, (tid >> 3) + 32 * q, tid & 7, make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w), false);
    }
  }
};

// ---- epilogues: ep(m, n0, v[16], nv, z) for one accumulator row chunk ------------------------------
// An epilogue with load_aux / apply has its global inputs fetched one chunk ahead by the kernel.
template <class T, class = void> struct has_aux : std::false_type {};
template <class T>
struct has_aux<T, std::void_t<decltype(&T::load_aux)>> : std::true_type {};
// Staged epilogues (round 2, late).  With thread = accumulator row, a warp's 128-bit store touches 32 different
// 128-byte lines (one 16-byte piece of 32 rows): 32 LSU wavefronts per instruction, and the epilogue of a 128 x 208
// tile took ~9000 cycles -- 20-29 % of a short-K CTA's life (profiles/r02j, r02z timelines).  A staged epilogue
// first parks the tile in the (now idle) operand rings, row-major with a 4-float pad (conflict-free 128-bit
// stores: consecutive rows are 4 banks apart modulo 32), then writes it out ROW-wise:
//   kStaged == 1: a row of the tile is one cp.async.bulk shared -> global copy, issued by one thread per row;
//   kStaged == 2: warps walk rows, lanes walk 16-byte column groups: coalesced mask / accumulate / store.
template <class T, class = void> struct staged_kind : std::integral_constant<int, 0> {};
template <class T>
struct staged_kind<T, std::void_t<decltype(T::kStaged)>> : std::integral_constant<int, T::kStaged> {};
template <class T, class = void> struct has_bias_stage : std::false_type {};
template <class T>
struct has_bias_stage<T, std::void_t<decltype(&T::apply_staged)>> : std::true_type {};
struct EpBiasAct {  // y = act(acc + bias[n])
  float* y; long long ld; const float* bias; bool relu;
  static constexpr bool kRowReduce = false;
  static constexpr int kStaged = 1;
  __device__ __forceinline__ bool staged_ok(int m0, int n0, int ncols) const {
    return ld % 4 == 0 && ncols % 4 == 0 && ((reinterpret_cast<uintptr_t>(y + (long long)m0 * ld + n0) & 15) == 0);
  }
  __device__ __forceinline__ void pre(float* v, const float* bias_s) const {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      v[i] += bias_s[i];
      if (relu) v[i] = fmaxf(v[i], 0.f);
    }
  }
  __device__ __forceinline__ float* row_ptr(int m, int n0, int) const { return y + (long long)m * ld + n0; }
  // staged form: the kernel copies the tile's bias segment to shared memory once (its global load is issued
  // before the wait for the last MMA), so the per-chunk epilogue has no global load on its critical path
  // (round 1: four 128-bit bias loads per 16-column chunk, ~7 dependent round trips per thread)
  __device__ __forceinline__ float bias_at(int n, int n_end) const { return (bias && n < n_end) ? __ldg(bias + n) : 0.f; }
  __device__ __forceinline__ void apply_staged(int m, int n0, float* v, int nv, const float* bias_s) const {
    float* dst = y + (long long)m * ld + n0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      v[i] += bias_s[i];                      // zero beyond the tile / when there is no bias
      if (relu) v[i] = fmaxf(v[i], 0.f);
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
  __device__ __forceinline__ void operator()(int m, int n0, float* v, int nv, int) const {
    float* dst = y + (long long)m * ld + n0;
    if (bias && nv == 16 && ((reinterpret_cast<uintptr_t>(bias + n0) & 15) == 0)) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + n0 + 4 * q));
        v[4 * q] += b4.x; v[4 * q + 1] += b4.y; v[4 * q + 2] += b4.z; v[4 * q + 3] += b4.w;
      }
    } else if (bias) {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i < nv) v[i] += __ldg(bias + n0 + i);
    }
    if (relu) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
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
  static constexpr int kStaged = 2;
  __device__ __forceinline__ bool staged_ok(int m0, int n0, int ncols) const {
    return ld % 4 == 0 && ncols % 4 == 0 && ((reinterpret_cast<uintptr_t>(g + (long long)m0 * ld + n0) & 15) == 0) &&
           (!mask || (ldm % 4 == 0 && ((reinterpret_cast<uintptr_t>(mask + (long long)m0 * ldm + n0) & 15) == 0)));
  }
  __device__ __forceinline__ void pre(float*, const float*) const {}
  __device__ __forceinline__ float4 load_mask4(int m, int n) const {
    return mask ? __ldg(reinterpret_cast<const float4*>(mask + (long long)m * ldm + n)) : make_float4(1.f, 1.f, 1.f, 1.f);
  }
  __device__ __forceinline__ void store4(int m, int n, float4 o, const float4& k) const {
    float* dst = g + (long long)m * ld + n;
    if (!(k.x > 0.f)) o.x = 0.f;
    if (!(k.y > 0.f)) o.y = 0.f;
    if (!(k.z > 0.f)) o.z = 0.f;
    if (!(k.w > 0.f)) o.w = 0.f;
    if (accumulate) {
      const float4 p = *reinterpret_cast<const float4*>(dst);
      o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
    }
    *reinterpret_cast<float4*>(dst) = o;
  }
  // the mask chunk of a row is loaded one chunk ahead of its use (aux), so that its latency hides
  // under the TMEM load / stores of the previous chunk instead of serialising the epilogue
  __device__ __forceinline__ bool vec_ok(int m, int n0, int nv) const {
    const float* dst = g + (long long)m * ld + n0;
    const float* mk = mask ? mask + (long long)m * ldm + n0 : nullptr;
    return nv == 16 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) &&
           (!mk || (reinterpret_cast<uintptr_t>(mk) & 15) == 0);
  }
  __device__ __forceinline__ void load_aux(int m, int n0, int nv, float4* aux) const {
    if (!mask || !vec_ok(m, n0, nv)) return;
    const float* mk = mask + (long long)m * ldm + n0;
#pragma unroll
    for (int q = 0; q < 4; ++q) aux[q] = __ldg(reinterpret_cast<const float4*>(mk + 4 * q));
  }
  __device__ __forceinline__ void apply(int m, int n0, float* v, int nv, int, const float4* aux) const {
    float* dst = g + (long long)m * ld + n0;
    const float* mk = mask ? mask + (long long)m * ldm + n0 : nullptr;
    if (vec_ok(m, n0, nv)) {  // a thread owns 64 contiguous bytes of its row: 128-bit accesses
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 o = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        if (mk) {
          const float4 k = aux[q];
          if (!(k.x > 0.f)) o.x = 0.f;
          if (!(k.y > 0.f)) o.y = 0.f;
          if (!(k.z > 0.f)) o.z = 0.f;
          if (!(k.w > 0.f)) o.w = 0.f;
        }
        if (accumulate) {
          const float4 p = *reinterpret_cast<const float4*>(dst + 4 * q);
          o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
        }
        *reinterpret_cast<float4*>(dst + 4 * q) = o;
      }
      return;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (i < nv) {
        float t = v[i];
        if (mk && !(__ldg(mk + i) > 0.f)) t = 0.f;
        if (accumulate) t += dst[i];
        dst[i] = t;
      }
    }
  }
  __device__ __forceinline__ void operator()(int m, int n0, float* v, int nv, int z) const {
    float4 aux[4];
    load_aux(m, n0, nv, aux);
    apply(m, n0, v, nv, z, aux);
  }
};
struct EpPartial {  // split-K partial: ws[z][m][n]
  float* ws; long long MN; long long ld;
  static constexpr bool kRowReduce = false;
  static constexpr int kStaged = 1;
  __device__ __forceinline__ bool staged_ok(int m0, int n0, int ncols) const {
    return ld % 4 == 0 && MN % 4 == 0 && ncols % 4 == 0 && ((reinterpret_cast<uintptr_t>(ws + (long long)m0 * ld + n0) & 15) == 0);
  }
  __device__ __forceinline__ void pre(float*, const float*) const {}
  __device__ __forceinline__ float* row_ptr(int m, int n0, int z) const { return ws + (long long)z * MN + (long long)m * ld + n0; }
  __device__ __forceinline__ void operator()(int m, int n0, float* v, int nv, int z) const {
    float* dst = ws + (long long)z * MN + (long long)m * ld + n0;
    if (nv == 16 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<float4*>(dst + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      return;
    }
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
// CIN input gradient w.r.t. x0:  out[m, tile] += sum_n acc[m, n] * x[m, n]   (one N tile per field).
// The CTA's x tile [128 x ncols] is staged in shared memory with coalesced loads (row stride ncols+1:
// thread = row reads are then bank-conflict free); dZ = gy W never leaves TMEM.
struct EpRowDot {
  const float* x; long long ldx; float* out; long long ldo;
  static constexpr bool kRowReduce = true;
  __device__ __forceinline__ void load_tile(float* xs, int m0, int M, int ncols, int tid) const {
    // one warp per row pair, lanes over the columns: coalesced, no per-element division, and all
    // (up to 14) loads of an iteration in flight before the first store (latency, not bandwidth, rules)
    const int stride = ncols + 1;
    const int lane = tid & 31;
    for (int r = 2 * (tid >> 5); r < BM; r += 2 * (THREADS / 32)) {
      float v[2][7];
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const bool ok = m0 + r + rr < M;
        const float* src = x + (long long)(m0 + r + rr) * ldx;
#pragma unroll
        for (int u = 0; u < 7; ++u) {
          const int j = lane + 32 * u;
          v[rr][u] = (ok && j < ncols) ? __ldg(src + j) : 0.f;
        }
      }
#pragma unroll
      for (int rr = 0; rr < 2; ++rr)
#pragma unroll
        for (int u = 0; u < 7; ++u) {
          const int j = lane + 32 * u;
          if (j < ncols) xs[(r + rr) * stride + j] = v[rr][u];
        }
    }
  }
  __device__ __forceinline__ float dot(const float* xs, int ncols, int rl, int nl0, const float* v, int nv) const {
    float s = 0.f;
    const float* xr = xs + rl * (ncols + 1) + nl0;
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < nv) s = fmaf(v[i], xr[i], s);
    return s;
  }
  __device__ __forceinline__ void finish(int m, int tile, float s) const { out[(long long)m * ldo + tile] += s; }
};

// ====================================================================================================
// Packed-B variant.  Wherever the B operand is a WEIGHT matrix (Linear forward / gradInput, CIN
// forward, CIN dx0 / dx), it is split into hi / lo and laid out ONCE per step by pack_b_kernel as the
// exact shared-memory images of the GEMM's stages: blob[(n_tile * nkb + kb)] = {hi image, lo image},
// each bn x 128 B, already SWIZZLE_128B.  The GEMM then fetches a whole B stage with two
// cp.async.bulk copies issued by one thread into a 3-deep ring (2-deep when bn > 208), completion by
// mbarrier transaction count -- no registers, no ALU, and the copy of stage kb+1 is in flight for two
// MMA stages.  The A operand stays software-produced (generated CIN outer product, or activations
// split on the fly with loads two stages ahead).
// ====================================================================================================
constexpr int TCB_A_BYTES = STAGES * 2 * A_TILE_BYTES;   // 64 KB
__host__ __device__ inline int tcb_nb(int bn) { return bn <= 208 ? 3 : 2; }
__host__ __device__ inline int tcb_smem_bytes(int bn) {
  return 1024 + TCB_A_BYTES + tcb_nb(bn) * 2 * bn * 128 + 128;
}

template <class BP, class Sched>
__global__ void __launch_bounds__(THREADS) pack_b_kernel(int bn, int n_stride, Sched sched, BP bp, char* blob) {
  B200_PDL_ENTRY();
  const int nkb = sched.nkb();
  const int tile = blockIdx.x;
  bp.s = sched;
  bp.is_b = true;
  bp.init(nullptr, tile * n_stride, threadIdx.x);
  for (int kb = blockIdx.y; kb < nkb; kb += gridDim.y) {
    char* hi = blob + ((size_t)tile * nkb + kb) * (size_t)(2 * bn * 128);
    bp.prefetch(kb);
    bp.store(kb, hi, hi + bn * 128);
  }
}

// several Linear weights in one launch (blockIdx.z = job): W[N,K] row-major ->
//   TRANSPOSED = false: B(n, k) = W[n*K + k]          (forward:   y = x W^T)
//   TRANSPOSED = true : B(k_out, n) = W[n*K + k_out]  (gradInput: gx = gy W, contraction over N)
struct PackJob { const float* w; int N, K, bn, nkb, n_tiles; long long off; };
struct PackJobs { PackJob j[8]; };
template <bool TRANSPOSED>
__global__ void __launch_bounds__(THREADS) pack_linear_multi_kernel(PackJobs jobs, char* blob) {
  B200_PDL_ENTRY();
  const PackJob jb = jobs.j[blockIdx.z];
  const int tile = blockIdx.x;
  if (tile >= jb.n_tiles) return;
  const bool al = (reinterpret_cast<uintptr_t>(jb.w) & 15) == 0;
  for (int kb = blockIdx.y; kb < jb.nkb; kb += gridDim.y) {
    char* hi = blob + jb.off + ((size_t)tile * jb.nkb + kb) * (size_t)(2 * jb.bn * 128);
    if (!TRANSPOSED) {
      RowProd<8, KPlain> bp{jb.w, jb.K, jb.N, jb.bn, (jb.K % 4 == 0) && al};
      bp.s = KPlain{0, jb.K, jb.K};
      bp.is_b = true;
      bp.init(nullptr, tile * jb.bn, threadIdx.x);
      bp.prefetch(kb);
      bp.store(kb, hi, hi + jb.bn * 128);
    } else {
      ColProd<256, KPlain> bp{jb.w, jb.K, jb.K, jb.bn, (jb.K % 4 == 0) && al};
      bp.s = KPlain{0, jb.N, jb.N};
      bp.is_b = true;
      bp.init(nullptr, tile * jb.bn, threadIdx.x);
      bp.prefetch(kb);
      bp.store(kb, hi, hi + jb.bn * 128);
    }
  }
}

// ----------------------------------------------------------------------------------------------------
// Warp-specialised kernel (the one the launchers use).  10 warps:
//   warp 0      : one thread issues tcgen05.mma (4 TF32 + 4 bf16 K-steps per stage) and commits the
//                 stage's "empty" mbarriers; it never touches operand data
//   warp 1      : one thread issues the cp.async.bulk copies of packed B stages into their ring, paced by
//                 the ring's own "bempty" mbarriers.  (Round 1 had producer thread 0 issue them inside its
//                 stage routine: the in-kernel timeline showed ~550 cycles per stage between that
//                 thread's wait and its first store -- the bulk-copy issue -- and the stage's "full"
//                 barrier needs every producer, so every stage's MMAs started that much later.)
//   warps 2..9  : 256 producer threads fill the A stage (and the B stage when B is not packed),
//                 arrive on the stage's "full" mbarrier, and run up to STAGES ahead of the tensor
//                 core -- there is no __syncthreads in the main loop.  At every accumulation-chunk boundary
//                 the producers add the TMEM chunk accumulator P into the running sum S (RN) and
//                 release the MMA warp through the "drained" mbarrier.  They are the epilogue warps.
// barriers: full[0..1] (256 arrivals), empty[0..1] (tcgen05.commit), bfull[0..2] (tx bytes), bempty[0..2]
// (tcgen05.commit), drained.
// ----------------------------------------------------------------------------------------------------
constexpr int WS_THREADS = 64 + THREADS;   // MMA warp + B-loader warp + 8 producer / epilogue warps
constexpr int WS_PROD0 = 64;               // first producer thread
constexpr int KC_SHORT = 8;
#ifdef B200_TC_TRACE
// debug timeline of CTA (0,0,0): [role][kb][slot] clock64 stamps (built only into libb200rec_trace.so)
__device__ long long g_tc_trace[3 * 512 * 4];
__device__ long long g_tc_trace_dw[3 * 512 * 4];   // copy taken after a deep CIN layer's dW launch
#define TC_TRACE(role, kb, slot)                                                          \
  do {                                                                                    \
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (kb) < 512)              \
      g_tc_trace[((role) * 512 + (kb)) * 4 + (slot)] = clock64();                         \
  } while (0)
#else
#define TC_TRACE(role, kb, slot) do {} while (0)
#endif
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__host__ __device__ inline int ws_nb(bool packed, int bn) { return packed ? tcb_nb(bn) : 2; }
__host__ __device__ inline int ws_smem_bytes(bool packed, int bn) {
  return 1024 + TCB_A_BYTES + ws_nb(packed, bn) * 2 * bn * 128 + 128;
}

template <class AP, class BP, class Sched, class Ep, int PASSES, bool PACKED>
__global__ void __launch_bounds__(WS_THREADS, 1)   // registers are granted per 4 warps: 9 warps -> cap 168
gemm_ws_kernel(int M, int N, int bn, int n_stride, int n_valid, int kc, Sched sched, AP ap, BP bp,
               const char* __restrict__ bblob, int blob_nkb, int blob_kb_per_split, Ep ep) {
  // packed B with split-K: the blob holds the stages of the WHOLE contraction ([n_tile][blob_nkb]);
  // split z starts at stage z * blob_kb_per_split
  griddep_launch();   // the next kernel of the stream may be scheduled (it waits for this one to complete)
  extern __shared__ char smem_raw[];
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int NB = ws_nb(PACKED, bn);
  const int b_stage = 2 * bn * 128;
  char* bbase = base + TCB_A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(bbase + NB * b_stage);
  const uint32_t bar_full = smem_u32(&bars[0]);    // +8*st
  const uint32_t bar_bfull = smem_u32(&bars[2]);   // +8*sb
  const uint32_t bar_drained = smem_u32(&bars[5]);    // second column half of a chunk drained into S
  const uint32_t bar_half = smem_u32(&bars[6]);       // first column half of a chunk's last stage computed
  const uint32_t bar_drained0 = smem_u32(&bars[7]);   // first column half drained into S
  // ONE tcgen05.commit per stage, on done[kb % 6]: it frees A slot kb % 2 (the producers of stage kb + 2
  // wait for it) and B slot kb % 3 (the loader of stage kb + 3 does).  6 = lcm of the two ring depths, so a
  // barrier completes once per 6 stages and nobody can fall a whole phase behind: the MMAs of stage kb + 6
  // need the A stage and the B stage that are produced only after these waits.
  const uint32_t bar_done = smem_u32(&bars[8]);       // +8*(kb % 6)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  // Accumulation-chunk boundaries are pipelined by column halves [0, h0) / [h0, bn): the chunk's last
  // stage and the next chunk's first stage issue their MMAs half by half, so that the drain of one
  // half (P -> S, CUDA cores) runs under the tensor work of the other instead of idling the pipe.
#ifdef B200_NO_HALVES
  const int h0 = bn, h1 = 0;
#else
  const int h0 = ((bn / 16 + 1) / 2) * 16, h1 = bn - h0;
#endif
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM;
  const int n0 = blockIdx.x * n_stride;
  const Sched s = sched.for_split(blockIdx.z);
  const int nkb = s.nkb();
  // A contraction of at most KC_SHORT K-blocks (K <= 256) stays in ONE accumulation: the truncation
  // drift is linear in K (5.7e-5 at K = 8192 -> 1.8e-6 at 256, the level of the 3xTF32 split itself),
  // so the P -> S drain would buy nothing (CIN dx0 runs 7-stage CTAs).
  // (the host passes the threshold in the high bits of kc: kc = chunk | short_threshold << 8)
  {
    const int kc_short = kc >> 8;
    kc &= 255;
    if (nkb <= kc_short) kc = 0;
  }

  if (threadIdx.x == 0) {
    mbar_init(bar_full, THREADS); mbar_init(bar_full + 8, THREADS);
    mbar_init(bar_bfull, 1); mbar_init(bar_bfull + 8, 1); mbar_init(bar_bfull + 16, 1);
    mbar_init(bar_drained, THREADS);
    mbar_init(bar_half, 1);
    mbar_init(bar_drained0, THREADS);
    for (int i = 0; i < 6; ++i) mbar_init(bar_done + 8 * i, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // Programmatic dependent launch: everything above touched only shared / tensor memory, so it ran while the
  // previous kernel of the stream was still finishing; its results are needed (and this kernel's writes allowed)
  // from here on.
  griddep_wait();

  if (warp == 0) {
    // ===================================== MMA issuer =====================================
    // elect.sync (not `lane == 0`): ptxas then knows a single thread runs the block and keeps the
    // descriptors / TMEM address in uniform registers instead of re-broadcasting them per MMA
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    if (leader) {
      const uint32_t idesc = make_idesc(bn), idesc0 = make_idesc(h0), idesc1 = make_idesc(h1 > 0 ? h1 : 16);
      const uint32_t idesc_bf = make_idesc_bf16(bn), idesc0_bf = make_idesc_bf16(h0),
                     idesc1_bf = make_idesc_bf16(h1 > 0 ? h1 : 16);
      int sb = 0, bphase = 0, cpos = 0, cidx = 0, d6 = 0;   // B ring slot / its phase, position in / index of the chunk, kb % 6
      for (int kb = 0; kb < nkb; ++kb) {
        const int st = kb & 1;
        const bool chunk_start = kc > 0 ? (cpos == 0) : (kb == 0);
        const bool opens = kc > 0 && chunk_start && kb > 0;            // first stage after a boundary
        const bool closes = kc > 0 && cpos == kc - 1 && kb + 1 < nkb;   // last stage before a boundary
        TC_TRACE(0, kb, 0);
        mbar_wait(bar_full + 8 * st, (kb >> 1) & 1);
        TC_TRACE(0, kb, 1);
        if (PACKED) mbar_wait(bar_bfull + 8 * sb, bphase);
        TC_TRACE(0, kb, 2);
        tc_fence_after();
        char* a_hi = base + st * 2 * A_TILE_BYTES;
        char* b_hi = bbase + sb * b_stage;
        const uint64_t dah = make_desc(smem_u32(a_hi)), dal = make_desc(smem_u32(a_hi + A_TILE_BYTES));
        const uint64_t dbh = make_desc(smem_u32(b_hi)), dbl = make_desc(smem_u32(b_hi + bn * 128));
        const int kvalid = s.kvalid(kb);
        const int ksteps = (kvalid + UK - 1) / UK;
        const bool fresh = chunk_start;   // the first MMA of a chunk overwrites the accumulator
        if (opens || closes) {
          const uint64_t boff = (uint64_t)(h0 * 128 >> 4);   // B rows [h0, bn) of the stage image
#ifdef B200_BOUNDARY_OLD
          if (opens) {
            mbar_wait(bar_drained0, (cidx - 1) & 1);
            tc_fence_after();
          }
          mma_block<PASSES>(tmem, dah, dal, dbh, dbl, 0, idesc0, idesc0_bf, ksteps, kvalid, fresh);
          if (closes) mma_commit(bar_half);
          if (opens) {
            mbar_wait(bar_drained, (cidx - 1) & 1);
            tc_fence_after();
          }
          if (h1 > 0) mma_block<PASSES>(tmem + h0, dah, dal, dbh, dbl, boff, idesc1, idesc1_bf, ksteps, kvalid, fresh);
#else
          // Only the TF32 MMAs are split by column halves: a bf16 MMA costs the same 132 cycles at any N
          // (profiles/r02f_mma_rate_ubench.txt), so the corrections run once at full width -- before the halves in
          // a stage that closes a chunk, after them in a stage that opens one.
          const bool corr_first = !opens && !fresh;   // (a fresh stage's first TF32 MMA overwrites the accumulator)
          if (PASSES == 3 && corr_first) mma_block_bf16(tmem, dal, dbl, 0, idesc_bf, ksteps);
          if (opens) {
            mbar_wait(bar_drained0, (cidx - 1) & 1);
            tc_fence_after();
          }
          mma_block_tf32(tmem, dah, dbh, 0, idesc0, ksteps, fresh);
          if (closes && corr_first) mma_commit(bar_half);
          if (opens) {
            mbar_wait(bar_drained, (cidx - 1) & 1);
            tc_fence_after();
          }
          if (h1 > 0) mma_block_tf32(tmem + h0, dah, dbh, boff, idesc1, ksteps, fresh);
          if (!corr_first) {
            if (PASSES == 3) mma_block_bf16(tmem, dal, dbl, 0, idesc_bf, ksteps);
            if (closes) mma_commit(bar_half);
          }
#endif
        } else {
          mma_block<PASSES>(tmem, dah, dal, dbh, dbl, 0, idesc, idesc_bf, ksteps, kvalid, fresh);
        }
        mma_commit(bar_done + 8 * d6);
        TC_TRACE(0, kb, 3);
        if (++d6 == 6) d6 = 0;
        if (++sb == NB) { sb = 0; bphase ^= 1; }
        if (kc > 0 && ++cpos == kc) { cpos = 0; ++cidx; }
      }
    }
  } else if (warp == 1) {
    // ===================================== B loader =====================================
    if (PACKED) {
      uint32_t leader;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
      if (leader) {
        const char* myblob = bblob + ((size_t)blockIdx.x * blob_nkb + (size_t)blockIdx.z * blob_kb_per_split) * b_stage;
        int sb = 0, round = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          if (kb >= NB) mbar_wait(bar_done + 8 * ((kb - NB) % 6), ((kb - NB) / 6) & 1);   // the MMAs of stage kb - NB are done
#ifdef B200_EXPERIMENT_HALF_B   // timing experiment only (wrong results): is the stage period set by the L2 -> SM bytes?
          mbar_expect_tx(bar_bfull + 8 * sb, (uint32_t)b_stage / 2);
          bulk_g2s(smem_u32(bbase + sb * b_stage), myblob + (size_t)kb * b_stage, (uint32_t)b_stage / 2,
                   bar_bfull + 8 * sb);
#else
          mbar_expect_tx(bar_bfull + 8 * sb, (uint32_t)b_stage);
          bulk_g2s(smem_u32(bbase + sb * b_stage), myblob + (size_t)kb * b_stage, (uint32_t)b_stage,
                   bar_bfull + 8 * sb);
#endif
          if (++sb == NB) { sb = 0; ++round; }
        }
      }
    }
  } else {
    // ===================================== producers =====================================
    const int tid = threadIdx.x - WS_PROD0;
    const int pw = warp - 2;               // 0..7
    ap.s = s;
    ap.init(nullptr, m0, tid);
    if (!PACKED) {
      bp.s = s;
      bp.is_b = true;
      bp.init(nullptr, n0, tid);
    }
    if (nkb > 0) {
      ap.template prefetch2<0>(0);
      if (!PACKED) bp.prefetch(0);
    }
    if (AP::DIST == 2 && nkb > 1) ap.template prefetch2<1>(1);
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;

    int pcpos = 0;   // position of stage kb inside its accumulation chunk
    auto stage = [&](auto slot_tag, int kb) {
      constexpr int P = decltype(slot_tag)::value;
      char* a_hi = base + P * 2 * A_TILE_BYTES;
      if (tid == 0) TC_TRACE(1, kb, 0);
      if (kb >= STAGES) mbar_wait(bar_done + 8 * ((kb - 2) % 6), ((kb - 2) / 6) & 1);   // MMA(kb-2) done: stage P is free
      if (tid == 0) TC_TRACE(1, kb, 1);
#ifndef B200_EXPERIMENT_NO_ASTORE   // timing experiment only (wrong results): the pipeline without producer work
      ap.template store2<P>(kb, a_hi, a_hi + A_TILE_BYTES);
#endif
      if (tid == 0) TC_TRACE(1, kb, 2);
      if (!PACKED) {
        char* b_hi = bbase + P * b_stage;
        bp.store(kb, b_hi, b_hi + bn * 128);
      }
      // hand the stage to the MMA warp FIRST, then start the loads of a later stage: the prefetch (address
      // math + 4..8 global loads per thread) used to sit between the stores and the arrive and delayed every
      // stage's MMAs by a few hundred cycles
      fence_proxy_async();
      mbar_arrive(bar_full + 8 * P);
      if (tid == 0) TC_TRACE(1, kb, 3);
      if (!PACKED && kb + 1 < nkb) bp.prefetch(kb + 1);
      if (kb + AP::DIST < nkb) ap.template prefetch2<(AP::DIST == 2 ? P : 1 - P)>(kb + AP::DIST);
      if (kc > 0 && kb > 0 && pcpos == 0) {
        // Stage kb (just produced, so the MMA warp can start it the moment it is released) opens a new
        // accumulation chunk; the previous one ended with stage kb-1, whose MMAs ran as two column
        // halves.  As each half completes: S (+)= P with round-to-nearest for its columns, then the
        // MMA warp is released for that half.  All TMEM loads of a group of 16-column pieces are issued
        // before the first wait (fewer TMEM round trips).
        const bool have_s = kb > kc;
        const int ch0 = pw >> 2;
        const int bpar = (kb / kc - 1) & 1;        // boundary number: each barrier completes once per boundary
#ifndef B200_DRAIN_DG
#define B200_DRAIN_DG 1
#endif
        constexpr int DG = B200_DRAIN_DG;   // pieces per TMEM round trip (the drain now hides under the other half: registers matter more)
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {   // one copy of the drain code for both halves
          if (half == 0) mbar_wait(bar_half, bpar);
          else mbar_wait(bar_done + 8 * ((kb - 1) % 6), ((kb - 1) / 6) & 1);
          tc_fence_after();
          const int c_lo = half ? h0 / 16 : 0, c_hi = half ? bn / 16 : h0 / 16;
          const int first = c_lo + ((c_lo & 1) != ch0 ? 1 : 0);   // this thread's parity of 16-column pieces
#pragma unroll 1
          for (int c = first; c < c_hi; c += 2 * DG) {
            uint32_t p[DG][16], q[DG][16];
#pragma unroll
            for (int i = 0; i < DG; ++i) {
              const int ch = c + 2 * i;
              if (ch < c_hi) {
                tmem_ld16_nowait(tmem + lane_addr + ch * 16, p[i]);
                if (have_s) tmem_ld16_nowait(tmem + lane_addr + TMEM_S + ch * 16, q[i]);
              }
            }
#pragma unroll
            for (int i = 0; i < DG; ++i) {
              const int ch = c + 2 * i;
              if (ch < c_hi) {
                if (have_s) {
                  tmem_wait_ld2(p[i], q[i]);
#pragma unroll
                  for (int e = 0; e < 16; ++e)
                    p[i][e] = __float_as_uint(__uint_as_float(p[i][e]) + __uint_as_float(q[i][e]));
                } else {
                  tmem_wait_ld1(p[i]);
                }
                tmem_st16(tmem + lane_addr + TMEM_S + ch * 16, p[i]);
              }
            }
          }
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive(half ? bar_drained : bar_drained0);
          if (tid == 0) TC_TRACE(2, kb, half ? 0 : 1);
        }
      }
      if (kc > 0 && ++pcpos == kc) pcpos = 0;
    };
    for (int kb0 = 0; kb0 < nkb; kb0 += 2) {
      stage(std::integral_constant<int, 0>{}, kb0);
      if (kb0 + 1 < nkb) stage(std::integral_constant<int, 1>{}, kb0 + 1);
    }

    // ---- epilogue: thread = accumulator row (TMEM lane), 16 columns per tcgen05.ld ----------------
    if (nkb > 0) {
      if (tid == 0) TC_TRACE(2, 511, 0);
      const int row = m0 + (warp & 3) * 32 + lane;
      const int ncols = min(n_valid, N - n0);
      bool staged_on = false;
      if constexpr (staged_kind<Ep>::value != 0) {
#ifndef B200_NO_STAGED_EP
        staged_on = ep.staged_ok(m0, n0, ncols) && ncols <= 256;   // (256 columns: two 16-byte groups per lane and row)
#endif
      }
      [[maybe_unused]] float4 aux[4], aux_next[4];
      if constexpr (has_aux<Ep>::value) {   // first chunk's global inputs: in flight while the last MMAs drain
        if (!staged_on && row < M && (pw >> 2) * 16 < ncols)
          ep.load_aux(row, n0 + (pw >> 2) * 16, min(16, ncols - (pw >> 2) * 16), aux);
      }
      [[maybe_unused]] float bias_reg = 0.f;
      if constexpr (has_bias_stage<Ep>::value) bias_reg = ep.bias_at(n0 + tid, n0 + ncols);   // in flight under the wait
      mbar_wait(bar_done + 8 * ((nkb - 1) % 6), ((nkb - 1) / 6) & 1);
      if (tid == 0) TC_TRACE(2, 511, 1);
      tc_fence_after();
      [[maybe_unused]] float* bias_s = reinterpret_cast<float*>(base);   // A stage 0 is free: every MMA has completed
      if constexpr (has_bias_stage<Ep>::value) {
        bias_s[tid] = bias_reg;                                           // THREADS = 256 >= MAX_BN columns
        asm volatile("bar.sync 1, %0;" ::"r"(THREADS) : "memory");
      }
      const bool add_s = kc > 0 && nkb > kc;   // result = S + P (the last chunk is still in P)
      float v[16];
      if constexpr (Ep::kRowReduce) {
        // x tile -> shared memory (the B ring is free now), then thread = row dots; the two column
        // halves of a row are reduced by different warps and combined through shared memory
        float* xs = reinterpret_cast<float*>(bbase);
        ep.load_tile(xs, m0, M, ncols, tid);
        asm volatile("bar.sync 1, %0;" ::"r"(THREADS) : "memory");
        const int rl = (warp & 3) * 32 + lane;
        float acc = 0.f;
        for (int ch = pw >> 2; ch * 16 < ncols; ch += 2) {
          tmem_ld16(tmem + lane_addr + ch * 16, v);
          if (add_s) {
            float sv[16];
            tmem_ld16(tmem + lane_addr + TMEM_S + ch * 16, sv);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += sv[i];
          }
          acc += ep.dot(xs, ncols, rl, ch * 16, v, min(16, ncols - ch * 16));
        }
        float* red = reinterpret_cast<float*>(base);   // A stage 0 is free now
        if (pw >= 4) red[(warp & 3) * 32 + lane] = acc;
        asm volatile("bar.sync 1, %0;" ::"r"(THREADS) : "memory");
        if (pw < 4 && row < M) ep.finish(row, blockIdx.x, acc + red[(warp & 3) * 32 + lane]);
      } else if (staged_kind<Ep>::value != 0 && staged_on) {
        if constexpr (staged_kind<Ep>::value != 0) {
          // ---- phase 1: TMEM -> registers (-> bias / ReLU) -> staging tile, thread = accumulator row ----
          float* stg = reinterpret_cast<float*>(base + 1024);   // (the first KB holds the staged bias)
          const int lds = bn + 4;
          const int rl = (warp & 3) * 32 + lane;
          float* srow = stg + rl * lds;
          for (int ch = pw >> 2; ch * 16 < ncols; ch += 4) {
            const int ch2 = ch + 2;
            const bool two = ch2 * 16 < ncols;
            uint32_t pr[16], sr[16], pr2[16], sr2[16];
            float v2[16];
            tmem_ld16_nowait(tmem + lane_addr + ch * 16, pr);
            if (add_s) tmem_ld16_nowait(tmem + lane_addr + TMEM_S + ch * 16, sr);
            if (two) {
              tmem_ld16_nowait(tmem + lane_addr + ch2 * 16, pr2);
              if (add_s) tmem_ld16_nowait(tmem + lane_addr + TMEM_S + ch2 * 16, sr2);
            }
            if (add_s) {
              tmem_wait_ld2(pr, sr);
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(pr[i]) + __uint_as_float(sr[i]);
            } else {
              tmem_wait_ld1(pr);
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(pr[i]);
            }
            if (two) {
              if (add_s) {
                tmem_wait_ld2(pr2, sr2);
#pragma unroll
                for (int i = 0; i < 16; ++i) v2[i] = __uint_as_float(pr2[i]) + __uint_as_float(sr2[i]);
              } else {
                tmem_wait_ld1(pr2);
#pragma unroll
                for (int i = 0; i < 16; ++i) v2[i] = __uint_as_float(pr2[i]);
              }
            }
            ep.pre(v, bias_s + ch * 16);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              *reinterpret_cast<float4*>(srow + ch * 16 + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            if (two) {
              ep.pre(v2, bias_s + ch2 * 16);
#pragma unroll
              for (int q = 0; q < 4; ++q)
                *reinterpret_cast<float4*>(srow + ch2 * 16 + 4 * q) = make_float4(v2[4 * q], v2[4 * q + 1], v2[4 * q + 2], v2[4 * q + 3]);
            }
          }
          if constexpr (staged_kind<Ep>::value == 1) fence_proxy_async();   // the bulk copies read through the async proxy
          if (tid == 0) TC_TRACE(2, 510, 0);
          asm volatile("bar.sync 1, %0;" ::"r"(THREADS) : "memory");
          if (tid == 0) TC_TRACE(2, 510, 1);
          // ---- phase 2: row-wise write-out ----
          if constexpr (staged_kind<Ep>::value == 1) {
#ifndef B200_EP_STG_ROWS   // one cp.async.bulk per row; coalesced 128-bit stores by all warps measured the same or 1 % slower
            if (tid < BM && m0 + tid < M) {
              bulk_s2g(ep.row_ptr(m0 + tid, n0, blockIdx.z), smem_u32(stg + tid * lds), (uint32_t)ncols * 4u);
              bulk_commit_wait_read();
            }
#else
            const int nq = ncols >> 2;                    // 16-byte column groups of a row
            for (int r = pw; r < BM && m0 + r < M; r += THREADS / 32) {
              float* dst = ep.row_ptr(m0 + r, n0, blockIdx.z);
              const float* src = stg + r * lds;
              for (int c4 = lane; c4 < nq; c4 += 32)
                *reinterpret_cast<float4*>(dst + 4 * c4) = *reinterpret_cast<const float4*>(src + 4 * c4);
            }
#endif
          } else {
            const int nq = ncols >> 2;                    // 16-byte column groups of a row
            constexpr int RR = 8;                         // rows per warp and round: their mask loads fly together
            for (int r0 = pw * RR; r0 < BM; r0 += 8 * RR) {
              float4 mk[RR][2];
#pragma unroll
              for (int rr = 0; rr < RR; ++rr) {
                const int m = m0 + r0 + rr;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const int c4 = lane + 32 * h;
                  mk[rr][h] = (m < M && c4 < nq) ? ep.load_mask4(m, n0 + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
              }
#pragma unroll
              for (int rr = 0; rr < RR; ++rr) {
                const int m = m0 + r0 + rr;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const int c4 = lane + 32 * h;
                  if (m < M && c4 < nq)
                    ep.store4(m, n0 + 4 * c4, *reinterpret_cast<const float4*>(stg + (r0 + rr) * lds + 4 * c4), mk[rr][h]);
                }
              }
            }
          }
        }
      } else {
        if constexpr (has_aux<Ep>::value) {
          for (int ch = pw >> 2; ch * 16 < ncols; ch += 2) {
            uint32_t pr[16], sr[16];
            tmem_ld16_nowait(tmem + lane_addr + ch * 16, pr);
            if (add_s) tmem_ld16_nowait(tmem + lane_addr + TMEM_S + ch * 16, sr);
            if (row < M && (ch + 2) * 16 < ncols)
              ep.load_aux(row, n0 + (ch + 2) * 16, min(16, ncols - (ch + 2) * 16), aux_next);
            if (add_s) {
              tmem_wait_ld2(pr, sr);
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(pr[i]) + __uint_as_float(sr[i]);
            } else {
              tmem_wait_ld1(pr);
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(pr[i]);
            }
            if (row < M) ep.apply(row, n0 + ch * 16, v, min(16, ncols - ch * 16), blockIdx.z, aux);
#pragma unroll
            for (int q = 0; q < 4; ++q) aux[q] = aux_next[q];
          }
        } else {
        // two 16-column chunks per round: their TMEM loads (P and, when there is a running sum, S) are all in
        // flight before the first wait -- half as many TMEM round trips on the epilogue's critical path
        for (int ch = pw >> 2; ch * 16 < ncols; ch += 4) {
          const int ch2 = ch + 2;
          const bool two = ch2 * 16 < ncols;
          uint32_t pr[16], sr[16], pr2[16], sr2[16];
          float v2[16];
          tmem_ld16_nowait(tmem + lane_addr + ch * 16, pr);
          if (add_s) tmem_ld16_nowait(tmem + lane_addr + TMEM_S + ch * 16, sr);
          if (two) {
            tmem_ld16_nowait(tmem + lane_addr + ch2 * 16, pr2);
            if (add_s) tmem_ld16_nowait(tmem + lane_addr + TMEM_S + ch2 * 16, sr2);
          }
          if (add_s) {
            tmem_wait_ld2(pr, sr);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(pr[i]) + __uint_as_float(sr[i]);
          } else {
            tmem_wait_ld1(pr);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(pr[i]);
          }
          if (two) {
            if (add_s) {
              tmem_wait_ld2(pr2, sr2);      // (already complete: wait::ld covers every earlier load; this ties the registers)
#pragma unroll
              for (int i = 0; i < 16; ++i) v2[i] = __uint_as_float(pr2[i]) + __uint_as_float(sr2[i]);
            } else {
              tmem_wait_ld1(pr2);
#pragma unroll
              for (int i = 0; i < 16; ++i) v2[i] = __uint_as_float(pr2[i]);
            }
          }
          if constexpr (has_bias_stage<Ep>::value) {
            if (row < M) ep.apply_staged(row, n0 + ch * 16, v, min(16, ncols - ch * 16), bias_s + ch * 16);
            if (two && row < M) ep.apply_staged(row, n0 + ch2 * 16, v2, min(16, ncols - ch2 * 16), bias_s + ch2 * 16);
          } else {
            if (row < M) ep(row, n0 + ch * 16, v, min(16, ncols - ch * 16), blockIdx.z);
            if (two && row < M) ep(row, n0 + ch2 * 16, v2, min(16, ncols - ch2 * 16), blockIdx.z);
          }
        }
        }
      }
    }
  }
  if (threadIdx.x == WS_PROD0) TC_TRACE(2, 511, 2);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
  if (threadIdx.x == 0) TC_TRACE(2, 511, 3);
}

}  // namespace tc
}  // namespace b200rec
