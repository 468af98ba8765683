// CIN layer backward, both input gradients from ONE short-K GEMM (default whenever F <= 40 and
// C <= 256; B200REC_CIN_FUSED=0 selects the older two-GEMM path, see tc_cin_layer_bwd).
//
// Reference: rec/model/xdeepfm/CINEncoder.scala:105-177 (backward of MM + Linear of one CIN layer).
//   dZ[r, (i,j)] = sum_c gy[r,c] W[c, i*H + j]                    [R x F*H], contraction C <= 256
//   gx_in[r, j] (+)= sum_i dZ[r,(i,j)] x0[r,i]       gx0[r, i] += sum_j dZ[r,(i,j)] x_in[r,j]
// The older path computes gx0 from dZ with one seven-stage CTA per (row tile, field) and gx_in
// with a second, long-K GEMM on the generated operand (x0 (x) gy).  Here one persistent CTA per row
// tile sweeps all of dZ: N tiles of JT = 4 values of j, columns ordered n = jl * FP + i (FP = 40 >= F,
// pad columns are zero rows of the weight image), two TMEM accumulators so that the epilogue of tile
// t runs under the MMAs of tile t + 1.  In the epilogue a thread owns an accumulator row and half of
// the tile's columns (2 values of j): it finishes gx_in[r, j] for its two j inside the tile and keeps
// FP running sums of gx0[r, :] in registers for the whole sweep.  dZ never leaves TMEM; every dZ
// element feeds two FMAs instead of a second GEMM (half the tensor work of the older path; measured: 16.2 ms -> 5.5 ms per xDeepFM step).
// K <= 256 means one accumulation per tile (no P -> S drain, tc_gemm.cuh: KC_SHORT).
#pragma once
#include "tc_gemm.cuh"

namespace b200rec {
namespace tc {

constexpr int DZ_JT = 4;                      // values of j per N tile
constexpr int DZ_FP = 40;                     // columns per j (fields padded to a multiple of 8)
constexpr int DZ_BN = DZ_JT * DZ_FP;          // 160 accumulator columns
constexpr int DZ_HALF = DZ_BN / 2;            // columns per epilogue thread (2 values of j)
constexpr int DZ_CHUNKS = DZ_HALF / 16;       // 5 tcgen05.ld of 16 columns per thread and tile
constexpr int DZ_NB = 3;                      // B ring depth
constexpr int DZ_B_STAGE = 2 * DZ_BN * 128;   // hi | lo image of one K-block
constexpr int DZ_XS = DZ_FP + 1;              // row stride of the x0 tile in shared memory (conflict free)
constexpr int DZ_G = 3;                       // N tiles (accumulators) that share an A stage: 3 x 160 TMEM columns
__host__ __device__ inline int dz_smem_bytes() {
  return 1024 + TCB_A_BYTES + DZ_NB * DZ_B_STAGE + BM * DZ_XS * 4 + 256;   // 256: 16 mbarriers + the TMEM slot
}

// weight image: blob[tile * nkb + kb] = {hi, lo}, row n = jl * FP + i  <->  W[c, i*H + tile*JT + jl],
// column = c - 32 kb; rows with i >= F or j >= H and columns with c >= C are zero
__global__ void __launch_bounds__(THREADS) dz_pack_kernel(int F, int H, int C, const float* __restrict__ W,
                                                          char* blob) {
  B200_PDL_ENTRY();
  const int tile = blockIdx.x, kb = blockIdx.y, nkb = gridDim.y;
  char* hi = blob + ((size_t)tile * nkb + kb) * (size_t)DZ_B_STAGE;
  char* lo = hi + DZ_BN * 128;
  const long long FH = (long long)F * H;
  for (int q = threadIdx.x; q < DZ_BN * 8; q += THREADS) {
    const int r = q >> 3, c = q & 7;
    const int jl = r / DZ_FP, i = r - jl * DZ_FP;
    const int j = tile * DZ_JT + jl;
    const int k0 = kb * BK + 4 * c;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (i < F && j < H) {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (k0 + e < C) v[e] = __ldg(W + (long long)(k0 + e) * FH + (long long)i * H + j);
    }
    split_store(hi, lo, r, c, make_float4(v[0], v[1], v[2], v[3]), true);
  }
}

template <int PASSES>
__global__ void __launch_bounds__(WS_THREADS, 1)
cin_dz_kernel(int R, int F, int H, int C, const float* __restrict__ x0, const float* __restrict__ x_in,
              RowProd<4, KPlain> ap, const char* __restrict__ blob, float* gx_in, bool gx_acc, float* gx0) {
  // N tiles are swept in GROUPS of DZ_G = 3 that share every A stage: the gy tile of a K-block is split and
  // stored once per group and feeds the MMAs of three accumulators (TMEM columns 0 / 160 / 320), so the
  // producers do a third of round 1's work per MMA and the full / done hand-shakes happen once per 24 MMAs.
  // The epilogue of group g - 1 (15 pieces of 16 columns) runs on the producer warps right after they have
  // put the first two A stages of group g in flight; the MMA warp re-uses accumulator t as soon as its tile
  // has been read (accempty[t]).
  griddep_launch();   // (programmatic dependent launch: see gemm_ws_kernel)
  extern __shared__ char smem_raw[];
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  char* bbase = base + TCB_A_BYTES;
  float* xs0 = reinterpret_cast<float*>(bbase + DZ_NB * DZ_B_STAGE);      // [BM][DZ_XS]
  uint64_t* bars = reinterpret_cast<uint64_t*>(xs0 + BM * DZ_XS);
  const uint32_t bar_full = smem_u32(&bars[0]);        // +8*st : A stage filled (THREADS arrivals)
  const uint32_t bar_bfull = smem_u32(&bars[2]);       // +8*sb : B stage landed (tx bytes)
  const uint32_t bar_accfull = smem_u32(&bars[5]);     // +8*(grp & 1): a group's MMAs done (commit)
  const uint32_t bar_accempty = smem_u32(&bars[7]);    // +8*t  : the epilogue has read accumulator t (THREADS)
  const uint32_t bar_bdone = smem_u32(&bars[10]);      // +8*(bs % 6): the MMAs of B stage bs are done (commit); frees
                                                       // B slot bs % 3 and, for a group's last tile, the A slot
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM;
  const KPlain s{0, C, C};
  const int nkb = s.nkb();
  const int n_tiles = (H + DZ_JT - 1) / DZ_JT;
  const int n_groups = (n_tiles + DZ_G - 1) / DZ_G;
  const int AS = n_groups * nkb;                 // A stages of the whole sweep: as = grp * nkb + kb
  auto tiles_in = [&](int grp) { return min(DZ_G, n_tiles - grp * DZ_G); };
  // B stages issued before A stage `as` (every group before the last is full)
  auto b_before = [&](int as) { const int grp = as / nkb, kb = as - grp * nkb; return grp * DZ_G * nkb + kb * tiles_in(grp); };

  if (threadIdx.x == 0) {
    mbar_init(bar_full, THREADS); mbar_init(bar_full + 8, THREADS);
    mbar_init(bar_bfull, 1); mbar_init(bar_bfull + 8, 1); mbar_init(bar_bfull + 16, 1);
    mbar_init(bar_accfull, 1); mbar_init(bar_accfull + 8, 1);
    for (int i = 0; i < DZ_G; ++i) mbar_init(bar_accempty + 8 * i, THREADS);
    for (int i = 0; i < 6; ++i) mbar_init(bar_bdone + 8 * i, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  griddep_wait();   // first global access below
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ===================================== MMA issuer =====================================
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    if (leader) {
      const uint32_t idesc = make_idesc(DZ_BN), idesc_bf = make_idesc_bf16(DZ_BN);
      int bs = 0, as = 0;
      for (int grp = 0; grp < n_groups; ++grp) {
        const int nt = tiles_in(grp);
        for (int kb = 0; kb < nkb; ++kb, ++as) {
          const int st = as & 1;
          mbar_wait(bar_full + 8 * st, (as >> 1) & 1);
          tc_fence_after();
          char* a_hi = base + st * 2 * A_TILE_BYTES;
          const uint64_t dah = make_desc(smem_u32(a_hi)), dal = make_desc(smem_u32(a_hi + A_TILE_BYTES));
          const int kvalid = s.kvalid(kb);
          for (int t = 0; t < nt; ++t, ++bs) {
            if (kb == 0 && grp > 0) {   // the epilogue of group grp - 1 must have read this accumulator
              mbar_wait(bar_accempty + 8 * t, (grp - 1) & 1);
              tc_fence_after();
            }
            const int sb = bs % DZ_NB;
            mbar_wait(bar_bfull + 8 * sb, (bs / DZ_NB) & 1);
            tc_fence_after();
            char* b_hi = bbase + sb * DZ_B_STAGE;
            const uint64_t dbh = make_desc(smem_u32(b_hi)), dbl = make_desc(smem_u32(b_hi + DZ_BN * 128));
            mma_block<PASSES>(tmem + t * DZ_BN, dah, dal, dbh, dbl, 0, idesc, idesc_bf, (kvalid + UK - 1) / UK, kvalid,
                              kb == 0);
            mma_commit(bar_bdone + 8 * (bs % 6));
          }
        }
        mma_commit(bar_accfull + 8 * (grp & 1));
      }
    }
  } else if (warp == 1) {
    // ===================================== B loader (see gemm_ws_kernel) =====================================
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    if (leader) {
      int bs = 0;
      for (int grp = 0; grp < n_groups; ++grp) {
        const int nt = tiles_in(grp);
        for (int kb = 0; kb < nkb; ++kb)
          for (int t = 0; t < nt; ++t, ++bs) {
            const int sb = bs % DZ_NB;
            if (bs >= DZ_NB) mbar_wait(bar_bdone + 8 * ((bs - DZ_NB) % 6), ((bs - DZ_NB) / 6) & 1);
            mbar_expect_tx(bar_bfull + 8 * sb, (uint32_t)DZ_B_STAGE);
            bulk_g2s(smem_u32(bbase + sb * DZ_B_STAGE),
                     blob + ((size_t)(grp * DZ_G + t) * nkb + kb) * DZ_B_STAGE, (uint32_t)DZ_B_STAGE,
                     bar_bfull + 8 * sb);
          }
      }
    }
  } else {
    // ============================ producers / epilogue warps ============================
    const int tid = threadIdx.x - WS_PROD0;
    const int pw = warp - 2;                          // 0..7
    const int set = pw >> 2;                          // which half of a tile's columns (2 values of j)
    const int rl = (warp & 3) * 32 + lane;            // accumulator row = TMEM lane
    const int row = m0 + rl;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    ap.s = s;
    ap.init(nullptr, m0, tid);
    // x0 tile -> shared memory once (zero beyond F and beyond R)
    for (int q = tid; q < BM * DZ_FP; q += THREADS) {
      const int r = q / DZ_FP, i = q - r * DZ_FP;
      xs0[r * DZ_XS + i] = (m0 + r < R && i < F) ? __ldg(x0 + (long long)(m0 + r) * F + i) : 0.f;
    }
    asm volatile("bar.sync 1, %0;" ::"r"(THREADS) : "memory");
    if (AS > 0) ap.template prefetch2<0>(0);
    if (AS > 1) ap.template prefetch2<1>(1 % nkb);

    float dx0acc[DZ_FP];
#pragma unroll
    for (int i = 0; i < DZ_FP; ++i) dx0acc[i] = 0.f;
    float dxacc[2] = {0.f, 0.f};
    float xj[2], xj_next[2];
    auto load_xj = [&](int tile, float (&dst)[2]) {   // x_in[row, j] of this thread's two j in `tile`
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int j = tile * DZ_JT + 2 * set + u;
        dst[u] = (row < R && j < H && tile < n_tiles) ? __ldg(x_in + (long long)row * H + j) : 0.f;
      }
    };
    load_xj(0, xj_next);

    // one 16-column piece of a tile's epilogue; CI (0..4) is static so that every index into the
    // register arrays is (a run-time one would move dx0acc to local memory)
    auto epi_body = [&](auto ci_tag, const float* v) {
      constexpr int CI = decltype(ci_tag)::value;
      const float* xr = xs0 + rl * DZ_XS;
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int col = 16 * CI + e;            // column inside this thread's half: jl * FP + i
        const int jl = col / DZ_FP, i = col - jl * DZ_FP;
        dxacc[jl] = fmaf(v[e], xr[i], dxacc[jl]);
        dx0acc[i] = fmaf(v[e], xj[jl], dx0acc[i]);
      }
    };
    auto store_dx = [&](int tile, int u) {
      const int j = tile * DZ_JT + 2 * set + u;
      if (row < R && j < H) {
        float* dst = gx_in + (long long)row * H + j;
        *dst = gx_acc ? *dst + dxacc[u] : dxacc[u];
      }
    };
    auto epi_chunk = [&](int tile, int c) {
      const int grp = tile / DZ_G, t = tile - grp * DZ_G;
      if (c == 0) {
        mbar_wait(bar_accfull + 8 * (grp & 1), (grp >> 1) & 1);   // (complete already for the group's later tiles)
        tc_fence_after();
        xj[0] = xj_next[0]; xj[1] = xj_next[1];
        dxacc[0] = 0.f; dxacc[1] = 0.f;
      }
      float v[16];
      tmem_ld16(tmem + lane_addr + t * DZ_BN + set * DZ_HALF + c * 16, v);
      switch (c) {
        case 0: epi_body(std::integral_constant<int, 0>{}, v); break;
        case 1: epi_body(std::integral_constant<int, 1>{}, v); break;
        case 2: epi_body(std::integral_constant<int, 2>{}, v); store_dx(tile, 0); break;   // columns 32..39 end j 0
        case 3: epi_body(std::integral_constant<int, 3>{}, v); break;
        default:
          epi_body(std::integral_constant<int, 4>{}, v);
          store_dx(tile, 1);
          load_xj(tile + 1, xj_next);       // next tile's x values: in flight for a whole tile
          tc_fence_before();
          mbar_arrive(bar_accempty + 8 * t);
          break;
      }
    };
    auto epilogue_group = [&](int grp) {
      const int t0 = grp * DZ_G, t1 = min(n_tiles, t0 + DZ_G);
      for (int tile = t0; tile < t1; ++tile)
        for (int c = 0; c < DZ_CHUNKS; ++c) epi_chunk(tile, c);
    };

    int grp = 0, kb = 0;
    auto stage = [&](auto slot_tag, int as) {
      constexpr int P = decltype(slot_tag)::value;
      char* a_hi = base + P * 2 * A_TILE_BYTES;
      if (as >= STAGES) {   // the MMAs of A stage as - 2 are done when its LAST B stage is
        const int bl = b_before(as - 2) + tiles_in((as - 2) / nkb) - 1;
        mbar_wait(bar_bdone + 8 * (bl % 6), (bl / 6) & 1);
      }
      ap.template store2<P>(kb, a_hi, a_hi + A_TILE_BYTES);
      fence_proxy_async();
      mbar_arrive(bar_full + 8 * P);
      if (as + 2 < AS) {     // after the hand-off: the stage's MMAs must not wait for these loads to be issued
        int kb2 = kb + 2;
        if (kb2 >= nkb) kb2 -= nkb;
        if (kb2 >= nkb) kb2 -= nkb;   // nkb == 1
        ap.template prefetch2<P>(kb2);
      }
      // the previous group's epilogue, once the first two A stages of this group are in flight (one when the
      // contraction has a single K-block)
      if (grp > 0 && kb == (nkb > 1 ? 1 : 0)) epilogue_group(grp - 1);
      if (++kb == nkb) { kb = 0; ++grp; }
    };
    for (int a0 = 0; a0 < AS; a0 += 2) {
      stage(std::integral_constant<int, 0>{}, a0);
      if (a0 + 1 < AS) stage(std::integral_constant<int, 1>{}, a0 + 1);
    }
    if (n_groups > 0) epilogue_group(n_groups - 1);

    // gx0[row, :] += the two column halves' sums (the x0 tile is no longer read: reuse it)
    asm volatile("bar.sync 1, %0;" ::"r"(THREADS) : "memory");
    if (set == 1) {
#pragma unroll
      for (int i = 0; i < DZ_FP; ++i) xs0[rl * DZ_XS + i] = dx0acc[i];
    }
    asm volatile("bar.sync 1, %0;" ::"r"(THREADS) : "memory");
    if (set == 0 && row < R) {
#pragma unroll
      for (int i = 0; i < DZ_FP; ++i)
        if (i < F) gx0[(long long)row * F + i] += dx0acc[i] + xs0[rl * DZ_XS + i];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace tc
}  // namespace b200rec
