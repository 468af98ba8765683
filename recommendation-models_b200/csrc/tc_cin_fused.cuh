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
constexpr int DZ_TMEM_BUF = 256;              // column offset of the second accumulator
__host__ __device__ inline int dz_smem_bytes() {
  return 1024 + TCB_A_BYTES + DZ_NB * DZ_B_STAGE + BM * DZ_XS * 4 + 256;
}

// weight image: blob[tile * nkb + kb] = {hi, lo}, row n = jl * FP + i  <->  W[c, i*H + tile*JT + jl],
// column = c - 32 kb; rows with i >= F or j >= H and columns with c >= C are zero
__global__ void __launch_bounds__(THREADS) dz_pack_kernel(int F, int H, int C, const float* __restrict__ W,
                                                          char* blob) {
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
  extern __shared__ char smem_raw[];
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  char* bbase = base + TCB_A_BYTES;
  float* xs0 = reinterpret_cast<float*>(bbase + DZ_NB * DZ_B_STAGE);      // [BM][DZ_XS]
  uint64_t* bars = reinterpret_cast<uint64_t*>(xs0 + BM * DZ_XS);
  const uint32_t bar_full = smem_u32(&bars[0]);        // +8*st : A stage filled (THREADS arrivals)
  const uint32_t bar_bfull = smem_u32(&bars[2]);       // +8*sb : B stage landed (tx bytes)
  const uint32_t bar_accfull = smem_u32(&bars[5]);     // +8*buf: a tile's MMAs done (commit)
  const uint32_t bar_accempty = smem_u32(&bars[7]);    // +8*buf: the tile's epilogue read it (THREADS)
  const uint32_t bar_done = smem_u32(&bars[9]);        // +8*(g % 6): the MMAs of stage g are done (one commit per
                                                       // stage frees A slot g % 2 and B slot g % 3, see gemm_ws_kernel)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM;
  const KPlain s{0, C, C};
  const int nkb = s.nkb();
  const int n_tiles = (H + DZ_JT - 1) / DZ_JT;
  const int G = n_tiles * nkb;   // stages of the whole sweep; stage g = (tile g / nkb, K-block g % nkb)

  if (threadIdx.x == 0) {
    mbar_init(bar_full, THREADS); mbar_init(bar_full + 8, THREADS);
    mbar_init(bar_bfull, 1); mbar_init(bar_bfull + 8, 1); mbar_init(bar_bfull + 16, 1);
    mbar_init(bar_accfull, 1); mbar_init(bar_accfull + 8, 1);
    mbar_init(bar_accempty, THREADS); mbar_init(bar_accempty + 8, THREADS);
    for (int i = 0; i < 6; ++i) mbar_init(bar_done + 8 * i, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ===================================== MMA issuer =====================================
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    if (leader) {
      const uint32_t idesc = make_idesc(DZ_BN), idesc_bf = make_idesc_bf16(DZ_BN);
      int g = 0, sb = 0, bphase = 0;
      for (int t = 0; t < n_tiles; ++t) {
        const int buf = t & 1;
        if (t >= 2) {   // the epilogue of tile t - 2 must have read this accumulator
          mbar_wait(bar_accempty + 8 * buf, ((t >> 1) - 1) & 1);
          tc_fence_after();
        }
        const uint32_t acc = tmem + buf * DZ_TMEM_BUF;
        for (int kb = 0; kb < nkb; ++kb, ++g) {
          const int st = g & 1;
          mbar_wait(bar_full + 8 * st, (g >> 1) & 1);
          mbar_wait(bar_bfull + 8 * sb, bphase);
          tc_fence_after();
          char* a_hi = base + st * 2 * A_TILE_BYTES;
          char* b_hi = bbase + sb * DZ_B_STAGE;
          const uint64_t dah = make_desc(smem_u32(a_hi)), dal = make_desc(smem_u32(a_hi + A_TILE_BYTES));
          const uint64_t dbh = make_desc(smem_u32(b_hi)), dbl = make_desc(smem_u32(b_hi + DZ_BN * 128));
          const int kvalid = s.kvalid(kb);
          mma_block<PASSES>(acc, dah, dal, dbh, dbl, 0, idesc, idesc_bf, (kvalid + UK - 1) / UK, kvalid, kb == 0);
          mma_commit(bar_done + 8 * (g % 6));
          if (++sb == DZ_NB) { sb = 0; bphase ^= 1; }
        }
        mma_commit(bar_accfull + 8 * buf);
      }
    }
  } else if (warp == 1) {
    // ===================================== B loader (see gemm_ws_kernel) =====================================
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    if (leader) {
      int sb = 0, round = 0;
      for (int g = 0; g < G; ++g) {
        if (g >= DZ_NB) mbar_wait(bar_done + 8 * ((g - DZ_NB) % 6), ((g - DZ_NB) / 6) & 1);
        mbar_expect_tx(bar_bfull + 8 * sb, (uint32_t)DZ_B_STAGE);
        bulk_g2s(smem_u32(bbase + sb * DZ_B_STAGE), blob + (size_t)g * DZ_B_STAGE, (uint32_t)DZ_B_STAGE,
                 bar_bfull + 8 * sb);
        if (++sb == DZ_NB) { sb = 0; ++round; }
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
    if (G > 0) ap.template prefetch2<0>(0);
    if (G > 1) ap.template prefetch2<1>(1 % nkb);

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
      const int buf = tile & 1;
      if (c == 0) {
        mbar_wait(bar_accfull + 8 * buf, (tile >> 1) & 1);
        tc_fence_after();
        xj[0] = xj_next[0]; xj[1] = xj_next[1];
        dxacc[0] = 0.f; dxacc[1] = 0.f;
      }
      float v[16];
      tmem_ld16(tmem + lane_addr + buf * DZ_TMEM_BUF + set * DZ_HALF + c * 16, v);
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
          mbar_arrive(bar_accempty + 8 * buf);
          break;
      }
    };

    int t = 0, kb = 0, epi_done = 0;   // stage g = (t, kb); pieces of tile t - 1 already drained
    auto stage = [&](auto slot_tag, int g) {
      constexpr int P = decltype(slot_tag)::value;
      char* a_hi = base + P * 2 * A_TILE_BYTES;
      if (g >= STAGES) mbar_wait(bar_done + 8 * ((g - 2) % 6), ((g - 2) / 6) & 1);   // MMA(g-2) done: stage P is free
      ap.template store2<P>(kb, a_hi, a_hi + A_TILE_BYTES);
      fence_proxy_async();
      mbar_arrive(bar_full + 8 * P);
      if (g + 2 < G) {     // after the hand-off: the stage's MMAs must not wait for these loads to be issued
        int kb2 = kb + 2;
        if (kb2 >= nkb) kb2 -= nkb;
        if (kb2 >= nkb) kb2 -= nkb;   // nkb == 1
        ap.template prefetch2<P>(kb2);
      }
      // the previous tile's epilogue, one piece per stage from its second stage on (stage (t, 1) could
      // only be produced after the last MMA of tile t - 1 completed), the rest at the tile's last stage
      if (t > 0) {
        const int want = kb == nkb - 1 ? DZ_CHUNKS : (kb < DZ_CHUNKS ? kb : DZ_CHUNKS);
        while (epi_done < want) epi_chunk(t - 1, epi_done++);
      }
      if (++kb == nkb) { kb = 0; ++t; epi_done = 0; }
    };
    for (int g0 = 0; g0 < G; g0 += 2) {
      stage(std::integral_constant<int, 0>{}, g0);
      if (g0 + 1 < G) stage(std::integral_constant<int, 1>{}, g0 + 1);
    }
    if (n_tiles > 0)
      for (int c = 0; c < DZ_CHUNKS; ++c) epi_chunk(n_tiles - 1, c);

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
