// Host-side launchers of the tcgen05 GEMM (tc_gemm.cuh) for the shapes of the path:
//   BigDL Linear forward / gradInput / accGradParameters (rec/util/LayerUtil.scala:7-24,
//   rec/model/encoder/HigherOrderEncoder.scala:34-58) and the CIN layer forward / backward
//   (rec/model/xdeepfm/CINEncoder.scala:60-103,150-157).
#include "tc_gemm.cuh"
#include "tc_cin_fused.cuh"
#include <cstdlib>
#include "gemm_simt.cuh"
#include "kernels.h"

namespace b200rec {
namespace tc {

// K-blocks per TMEM accumulation chunk in the parity-grade mode, and the longest contraction (in
// K-blocks) that stays in ONE accumulation.  tcgen05.mma accumulates with truncation, so a long
// accumulation drifts linearly with the number of accumulating instructions; every chunk is therefore
// added into a running sum with round-to-nearest by the CUDA cores (the "drain": 2 x 106 KB of TMEM reads
// per boundary at 64 B/clk -- about as long as the MMAs of 4 K-blocks, which is why the chunk is as long
// as the error budget allows).  Measured on B200 (profiles/r02h_kc_sweep.txt): chunks of 8 K-blocks keep a
// K = 512 contraction at < 3e-6 of the output's largest magnitude (16-block single accumulations reach
// 3.3e-6, a third of the whole 1e-5 budget in one layer) and cost ~5 % of CIN-forward time against 16.
// B200REC_KC / B200REC_KC_SHORT override them (measurement only).
static int env_int(const char* name, int dflt) {
  const char* e = std::getenv(name);
  return e && *e ? std::atoi(e) : dflt;
}
static int kc_precise() { static const int v = env_int("B200REC_KC", 8); return v < 1 ? 1 : (v > 255 ? 255 : v); }
static int kc_short() { static const int v = env_int("B200REC_KC_SHORT", 8); return v < 0 ? 0 : v; }
static int round16(int n) { return (n + 15) / 16 * 16; }
// widest tile <= 256 that splits N evenly
static int pick_bn(int N) {
  const int tiles = (N + MAX_BN - 1) / MAX_BN;
  int bn = round16((N + tiles - 1) / tiles);
  return bn < 16 ? 16 : bn;
}

template <class AP, class BP, class Sched, class Ep, bool PACKED>
static int launch_ws(const char* name, dim3 grid, int smem, int M, int N, int bn, int n_stride, int n_valid,
                     Sched sched, AP ap, BP bp, const char* blob, int blob_nkb, int blob_kb_per_split, Ep ep,
                     int passes, cudaStream_t st) {
  const int kc = passes == 3 ? (kc_precise() | (kc_short() << 8)) : 0;
  // programmatic dependent launch for single-wave grids (the next kernel's prologue under this one's tail)
  PdlHint pdl_hint((long long)grid.x * grid.y * grid.z <= 148 || pdl_enabled() == 3);
  const int smem_max = ws_smem_bytes(PACKED, PACKED ? 208 : 256);
  if (passes == 3) {
    auto k = gemm_ws_kernel<AP, BP, Sched, Ep, 3, PACKED>;
    static bool attr_done = false;
    if (!attr_done) {
      B200_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
      attr_done = true;
    }
    B200_LAUNCH_NAMED(name, k, grid, WS_THREADS, smem, st, M, N, bn, n_stride, n_valid, kc, sched, ap, bp, blob, blob_nkb,
                      blob_kb_per_split, ep);
  } else {
    auto k = gemm_ws_kernel<AP, BP, Sched, Ep, 1, PACKED>;
    static bool attr_done = false;
    if (!attr_done) {
      B200_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
      attr_done = true;
    }
    B200_LAUNCH_NAMED(name, k, grid, WS_THREADS, smem, st, M, N, bn, n_stride, n_valid, kc, sched, ap, bp, blob, blob_nkb,
                      blob_kb_per_split, ep);
  }
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

// both operands produced by the CTA's producer warps (activation x activation contractions)
template <class AP, class BP, class Sched, class Ep>
static int launch(const char* name, int M, int N, int bn, int n_stride, int n_valid, int n_tiles, int splits,
                  Sched sched, AP ap, BP bp, Ep ep, int passes, cudaStream_t st) {
  if (M <= 0 || N <= 0) return B200REC_OK;
  dim3 grid(n_tiles, cdiv(M, BM), splits);
  return launch_ws<AP, BP, Sched, Ep, false>(name, grid, ws_smem_bytes(false, bn), M, N, bn, n_stride, n_valid,
                                             sched, ap, bp, nullptr, 0, 0, ep, passes, st);
}

static KCin make_kcin(int F, int H, int Hp, long long KS) {
  return KCin{F, H, Hp, KS, (unsigned)((0x100000000ULL + (unsigned)F - 1) / (unsigned)F)};
}
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
// the 4x4 register-transpose path of ColProd needs 16-B aligned rows of 4
static bool colvec(const float* p, long long ld, int rows_total) {
  return ld % 4 == 0 && rows_total % 4 == 0 && aligned16(p);
}

// split-K factor from a wave model: CTAs run in waves of 148 (one per SM); a CTA costs its stages plus
// a fixed prologue / epilogue (~6 stages).  Minimise waves x per-CTA cost.
static int pick_splits_waves(int tiles, int nkb, int max_s) {
  int best = 1;
  long long best_t = -1;
  for (int sp = 1; sp <= max_s && sp <= nkb; ++sp) {
    const long long waves = ((long long)tiles * sp + 147) / 148;
    const long long t = waves * ((nkb + sp - 1) / sp + 6);
    if (best_t < 0 || t < best_t) { best_t = t; best = sp; }
  }
  return best;
}

// widest tile <= 208 (3-deep B ring fits) that splits N evenly
static int pick_bn_packed(int N) {
  const int tiles = (N + 207) / 208;
  int bn = round16((N + tiles - 1) / tiles);
  return bn < 16 ? 16 : bn;
}

// pack the weight operand into stage images, then run the bulk-copy GEMM on it
// `sched` is the schedule of the WHOLE contraction (what gets packed); with splits > 1 (KPlain only)
// split z covers k_chunk of it and reads its stages from the same blob.
template <class AP, class BP, class Sched, class Ep>
static int launch_packed(const char* name, int M, int N, int bn, int n_stride, int n_valid, int n_tiles,
                         int nkb, Sched sched, AP ap, BP bp, Ep ep, int passes, cudaStream_t st,
                         const char* prepacked = nullptr, int splits = 1, int k_chunk = 0) {
  if (M <= 0 || N <= 0) return B200REC_OK;
  char* blob = const_cast<char*>(prepacked);
  if (!blob) {
    B200_REQUIRE(tl_pack, B200REC_ERR_STATE, "no weight-pack buffer bound to this thread");
    const size_t blob_bytes = (size_t)n_tiles * nkb * 2 * bn * 128;
    B200_TRY(tl_pack->reserve(blob_bytes));
    blob = tl_pack->as<char>();
    int py = 148 * 8 / (n_tiles > 0 ? n_tiles : 1);   // enough CTAs to pull HBM bandwidth on big operands
    if (py < 1) py = 1;
    dim3 pg(n_tiles, nkb < py ? nkb : py);
    B200_LAUNCH_NAMED("tc_pack_weights", (pack_b_kernel<BP, Sched>), pg, THREADS, 0, st, bn, n_stride, sched, bp, blob);
  }
  dim3 grid(n_tiles, cdiv(M, BM), splits);
  return launch_ws<AP, BP, Sched, Ep, true>(name, grid, ws_smem_bytes(true, bn), M, N, bn, n_stride, n_valid,
                                            sched, ap, bp, (const char*)blob, nkb, splits > 1 ? k_chunk / BK : 0,
                                            ep, passes, st);
}

}  // namespace tc

using namespace tc;

// the stage images of a Linear weight packed ahead by tc_prepack_linear, or nullptr
static const char* find_prepacked(const float* w, int fmt, int N, int K) {
  const PrePack* pp = tl_prepack;
  if (!pp || !pp->blob) return nullptr;
  for (int i = 0; i < pp->n; ++i)
    if (pp->e[i].w == w && pp->e[i].fmt == fmt && pp->e[i].N == N && pp->e[i].K == K)
      return pp->blob->as<char>() + pp->e[i].off;
  return nullptr;
}

int tc_prepack_linear(PrePack& pp, const float* mats, int in_dim, const int* dims, int n_layers,
                      const long long* w_off, bool with_dx, cudaStream_t st, cudaStream_t st_dx) {
  pp.n = 0;
  if (n_layers <= 0 || n_layers > 8 || !pp.blob) return B200REC_OK;
  PackJobs fwd{}, dx{};
  size_t off = 0;
  int max_tiles_f = 0, max_nkb_f = 0, max_tiles_d = 0, max_nkb_d = 0, nf = 0, nd = 0;
  int K = in_dim;
  for (int l = 0; l < n_layers; ++l) {
    const int N = dims[l];
    const float* w = mats + w_off[l];
    if (N > 1) {   // Linear(K -> 1) runs as a row dot product, not a GEMM
      {
        const int bn = pick_bn_packed(N), nt = cdiv(N, bn), nkb = cdiv(K, BK);
        fwd.j[nf++] = PackJob{w, N, K, bn, nkb, nt, (long long)off};
        pp.e[pp.n++] = PrePack::Entry{w, 0, N, K, off};
        off += (size_t)nt * nkb * 2 * bn * 128;
        max_tiles_f = nt > max_tiles_f ? nt : max_tiles_f;
        max_nkb_f = nkb > max_nkb_f ? nkb : max_nkb_f;
      }
      if (with_dx) {
        const int bn = pick_bn_packed(K), nt = cdiv(K, bn), nkb = cdiv(N, BK);
        dx.j[nd++] = PackJob{w, N, K, bn, nkb, nt, (long long)off};
        pp.e[pp.n++] = PrePack::Entry{w, 1, N, K, off};
        off += (size_t)nt * nkb * 2 * bn * 128;
        max_tiles_d = nt > max_tiles_d ? nt : max_tiles_d;
        max_nkb_d = nkb > max_nkb_d ? nkb : max_nkb_d;
      }
    }
    K = N;
  }
  if (!pp.n) return B200REC_OK;
  B200_TRY(pp.blob->reserve(off));
  char* blob = pp.blob->as<char>();
  if (nf) {
    dim3 g(max_tiles_f, max_nkb_f < 32 ? max_nkb_f : 32, nf);
    B200_LAUNCH_NAMED("tc_pack_weights", pack_linear_multi_kernel<false>, g, THREADS, 0, st, fwd, blob);
  }
  if (nd) {
    dim3 g(max_tiles_d, max_nkb_d < 32 ? max_nkb_d : 32, nd);
    B200_LAUNCH_NAMED("tc_pack_weights", pack_linear_multi_kernel<true>, g, THREADS, 0, st_dx, dx, blob);
  }
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

// y[M,N] = act(x[M,K] W[N,K]^T + b)
int tc_linear_fwd(int M, int N, int K, const float* x, const float* w, const float* b, bool relu,
                  float* y, int passes, cudaStream_t st) {
  const int bn = pick_bn_packed(N);
  KPlain s{0, K, K};
  RowProd<4, KPlain> ap{x, K, M, BM, (K % 4 == 0) && aligned16(x)};
  RowProd<8, KPlain> bp{w, K, N, bn, (K % 4 == 0) && aligned16(w)};
  return launch_packed("tc_linear_fwd", M, N, bn, bn, bn, cdiv(N, bn), cdiv(K, BK), s, ap, bp,
                       tc::EpBiasAct{y, N, b, relu}, passes, st, find_prepacked(w, 0, N, K));
}

// PNN product layer: h = relu(prev + ip Wp^T + c0)
int tc_pnn_lp_fwd(int B, int P, int O, const float* ip, const float* wp, const float* prev,
                  const float* c0, float* h, int passes, cudaStream_t st) {
  const int bn = pick_bn_packed(O);
  KPlain s{0, P, P};
  RowProd<4, KPlain> ap{ip, P, B, BM, (P % 4 == 0) && aligned16(ip)};
  RowProd<8, KPlain> bp{wp, P, O, bn, (P % 4 == 0) && aligned16(wp)};
  return launch_packed("tc_pnn_lp_fwd", B, O, bn, bn, bn, cdiv(O, bn), cdiv(P, BK), s, ap, bp,
                       tc::EpAddBiasRelu2{h, O, prev, c0}, passes, st);
}

// gx[M,K] = (gy[M,N] W[N,K]) (* mask > 0) (+ gx)
int tc_linear_bwd_input(int M, int N, int K, const float* gy, const float* w, const float* mask,
                        float* gx, bool accumulate, int passes, cudaStream_t st) {
  const int bn = pick_bn_packed(K);
  KPlain s{0, N, N};
  RowProd<4, KPlain> ap{gy, N, M, BM, (N % 4 == 0) && aligned16(gy)};
  ColProd<256, KPlain> bp{w, K, K, bn, colvec(w, K, K)};   // B(k_out, n) = W[n*K + k_out]
  return launch_packed("tc_linear_dx", M, K, bn, bn, bn, cdiv(K, bn), cdiv(N, BK), s, ap, bp,
                       tc::EpMaskAcc{gx, K, mask, K, accumulate}, passes, st, find_prepacked(w, 1, N, K));
}

// gw[N,K] (+)= scale * gy[M,N]^T x[M,K]  (split-K over the batch, fixed-order reduce) ; gb likewise
int tc_linear_bwd_params(int M, int N, int K, const float* x, const float* gy, float scale,
                         bool accumulate, float* gw, float* gb, DevBuf& scratch, int passes,
                         cudaStream_t st) {
  // With a bias gradient asked for, the B operand is [x | 1]: output column K is gb = colsum(gy).  The partials
  // then have K + 4 columns (16-byte rows for the staged epilogue; columns K + 1 .. K + 3 are zero).
  static const bool fuse_gb = [] { const char* e = std::getenv("B200REC_FUSE_BIAS_GRAD"); return !(e && e[0] == '0'); }();
  const bool with_gb = gb != nullptr && fuse_gb;
  const int Kc = with_gb ? K + 4 : K;
  const int bn = pick_bn(Kc);
  const int tiles = cdiv(N, BM) * cdiv(Kc, bn);
  int splits = pick_splits_waves(tiles, cdiv(M, BK), M / 256 > 0 ? M / 256 : 1);
  int k_chunk = ((cdiv(M, splits) + BK - 1) / BK) * BK;
  splits = cdiv(M, k_chunk);
  const long long MN = (long long)N * Kc;
  B200_TRY(scratch.reserve(((size_t)splits * MN + (size_t)COLSUM_CHUNKS * N) * sizeof(float)));
  float* ws = scratch.as<float>();
  KPlain s{0, M, k_chunk};
  ColProd<128, KPlain> ap{gy, N, N, BM, colvec(gy, N, N)};  // A(n, m) = gy[m*N + n]
  ColProd<256, KPlain> bp{x, K, K, bn, colvec(x, K, K)};   // B(k, m) = x[m*K + k]: packed once (all splits)
  if (with_gb) bp.ones_row = K;
  B200_TRY(launch_packed("tc_linear_dW", N, Kc, bn, bn, bn, cdiv(Kc, bn), cdiv(M, BK), s, ap, bp,
                         tc::EpPartial{ws, MN, Kc}, passes, st, nullptr, splits, k_chunk));
  if (with_gb) return splitk_reduce_wb(ws, splits, N, K, Kc, scale, accumulate, gw, gb, st);
  B200_TRY(splitk_reduce(ws, splits, MN, scale, accumulate, gw, st));
  if (gb) B200_TRY(colsum(M, N, gy, scale, accumulate, gb, ws + (size_t)splits * MN, st));
  return B200REC_OK;
}

// ---- CIN --------------------------------------------------------------------------------------------
bool tc_cin_supported(int F, int H, int C) {
  return F >= 1 && H >= 1 && H <= 208 && C >= 1;   // dx0 runs one N tile per field: H must fit a 208-wide tile
}

// x_out[r,c] = relu(sum_{i,j} x0[r,i] x_in[r,j] W[c, i*H+j] + b[c])
int tc_cin_layer_fwd(int R, int F, int H, int C, const float* x0, const float* x_in, const float* W,
                     const float* b, float* x_out, int passes, cudaStream_t st) {
  const int bn = pick_bn_packed(C);
  KCin s = make_kcin(F, H, 0, H);
  CinZProd ap{x0, x_in, R, (H % 4 == 0) && aligned16(x_in), false};
  RowProd<8, KCin> bp{W, (long long)F * H, C, bn, (H % 4 == 0) && aligned16(W)};
  return launch_packed("tc_cin_fwd", R, C, bn, bn, bn, cdiv(C, bn), F * cdiv(H, BK), s, ap, bp,
                       tc::EpBiasAct{x_out, C, b, true}, passes, st);
}

// Both CIN input gradients from one persistent short-K GEMM per row tile (tc_cin_fused.cuh) whenever the
// layer fits it (F <= 40 fields, C <= 256); B200REC_CIN_FUSED=0 keeps the two-GEMM path below.
// CIN dW from transposed copies of the factors (CinZtProdT); B200REC_CIN_DW_T=0 keeps round 1's producer
static bool cin_dw_transposed_enabled() {
  static const bool on = [] { const char* e = std::getenv("B200REC_CIN_DW_T"); return !(e && e[0] == '0'); }();
  return on;
}
static bool cin_fused_enabled() {
  static const bool on = [] { const char* e = std::getenv("B200REC_CIN_FUSED"); return !(e && e[0] == '0'); }();
  return on;
}
static int tc_cin_bwd_inputs_fused(int R, int F, int H, int C, const float* x0, const float* x_in,
                                   const float* W, const float* gy, float* gx_in, bool gx_in_acc, float* gx0,
                                   int passes, cudaStream_t st) {
  const int nkb = cdiv(C, BK), n_tiles = cdiv(H, DZ_JT);
  B200_REQUIRE(tl_pack, B200REC_ERR_STATE, "no weight-pack buffer bound to this thread");
  B200_TRY(tl_pack->reserve((size_t)n_tiles * nkb * DZ_B_STAGE));
  char* blob = tl_pack->as<char>();
  B200_LAUNCH_NAMED("tc_pack_weights", dz_pack_kernel, dim3(n_tiles, nkb), THREADS, 0, st, F, H, C, W, blob);
  RowProd<4, KPlain> ap{gy, C, R, BM, (C % 4 == 0) && aligned16(gy)};
  const int smem = dz_smem_bytes();
  if (passes == 3) {
    auto k = cin_dz_kernel<3>;
    static bool attr_done = false;
    if (!attr_done) { B200_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr_done = true; }
    B200_LAUNCH_NAMED("tc_cin_dz", k, cdiv(R, BM), WS_THREADS, smem, st, R, F, H, C, x0, x_in, ap, (const char*)blob,
                      gx_in, gx_in_acc, gx0);
  } else {
    auto k = cin_dz_kernel<1>;
    static bool attr_done = false;
    if (!attr_done) { B200_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr_done = true; }
    B200_LAUNCH_NAMED("tc_cin_dz", k, cdiv(R, BM), WS_THREADS, smem, st, R, F, H, C, x0, x_in, ap, (const char*)blob,
                      gx_in, gx_in_acc, gx0);
  }
  B200_CHECK_LAUNCH();
  return B200REC_OK;
}

// layer backward (gy already ReLU-masked):
//   gW[c, (i,j)] = sum_r gy[r,c] x0[r,i] x_in[r,j]          (split-K over r, fixed-order reduce)
//   gx0[r,i]    += sum_j (gy W)[r,(i,j)] x_in[r,j]           (one N tile per field, row-dot epilogue)
//   gx_in[r,j]   = sum_{i,c} x0[r,i] gy[r,c] W[c, i*H+j]     (the forward kernel with (x0 (x) gy) as A)
int tc_cin_layer_bwd(int R, int F, int H, int C, const float* x0, const float* x_in, const float* W,
                     const float* gy, float* gW, float* gb, float* gx_in, float* gx0, bool gx_in_acc,
                     DevBuf& scratch, int passes, cudaStream_t st) {
  const int FH = F * H;
  {  // ---- gW^T tile: out(m = (i,j), n = c), contraction over r
    const int bn = pick_bn(C);
    const int tiles = cdiv(FH, BM) * cdiv(C, bn);
    int splits = pick_splits_waves(tiles, cdiv(R, BK), R / 512 > 0 ? (R / 512 < 32 ? R / 512 : 32) : 1);
    int k_chunk = ((cdiv(R, splits) + BK - 1) / BK) * BK;
    splits = cdiv(R, k_chunk);
    const long long MN = (long long)C * FH;
    // + transposed copies of the two factors (x0T [F x R], xT [H x R]) for the generated Z^T operand
    const bool transposed = R % 4 == 0 && cin_dw_transposed_enabled();
    const size_t n_part = ((size_t)splits * MN + (size_t)COLSUM_CHUNKS * C + 3) / 4 * 4;
    B200_TRY(scratch.reserve((n_part + (transposed ? (size_t)R * (F + H) : 0)) * sizeof(float)));
    float* ws = scratch.as<float>();
    KPlain s{0, R, k_chunk};
    ColProd<256, KPlain> bp{gy, C, C, bn, colvec(gy, C, C)};   // B(c, r) = gy[r*C + c]: packed once (all splits)
    if (transposed) {
      float* x0T = ws + n_part;
      float* xT = x0T + (size_t)R * F;
      B200_TRY(transpose2d(R, F, x0, x0T, st));
      if (x_in == x0 && H == F) xT = x0T;                 // first layer: the second factor is x0 itself
      else B200_TRY(transpose2d(R, H, x_in, xT, st));
      CinZtProdT ap{x0T, xT, F, H, (long long)R};
      B200_TRY(launch_packed("tc_cin_dW", FH, C, bn, bn, bn, cdiv(C, bn), cdiv(R, BK), s, ap, bp,
                             tc::EpPartialT{ws, MN, FH}, passes, st, nullptr, splits, k_chunk));
    } else {
      CinZtProd ap{x0, x_in, F, H, (H % 4 == 0) && aligned16(x_in)};
      B200_TRY(launch_packed("tc_cin_dW", FH, C, bn, bn, bn, cdiv(C, bn), cdiv(R, BK), s, ap, bp,
                             tc::EpPartialT{ws, MN, FH}, passes, st, nullptr, splits, k_chunk));
    }
#ifdef B200_TC_TRACE
    if (H != F) {   // keep the timeline of a deep layer's dW launch (later launches overwrite g_tc_trace)
      void *src = nullptr, *dst = nullptr;
      cudaGetSymbolAddress(&src, tc::g_tc_trace);
      cudaGetSymbolAddress(&dst, tc::g_tc_trace_dw);
      cudaMemcpyAsync(dst, src, sizeof(long long) * 3 * 512 * 4, cudaMemcpyDeviceToDevice, st);
    }
#endif
    B200_TRY(splitk_reduce(ws, splits, MN, 1.0f, false, gW, st));
    B200_TRY(colsum(R, C, gy, 1.0f, false, gb, ws + (size_t)splits * MN, st));
  }
  if (cin_fused_enabled() && F <= DZ_FP && C <= KC_SHORT * BK)
    return tc_cin_bwd_inputs_fused(R, F, H, C, x0, x_in, W, gy, gx_in, gx_in_acc, gx0, passes, st);
  {  // ---- gx0: per field i, T_i[r, j] = sum_c gy[r,c] W[c, i*H + j];  gx0[r,i] += <T_i[r,:], x_in[r,:]>
     //      (one N tile per field: full-width MMAs; T = dZ stays in TMEM, the x tile is staged in smem)
    const int bn = round16(H);
    KPlain s{0, C, C};
    RowProd<4, KPlain> ap{gy, C, R, BM, (C % 4 == 0) && aligned16(gy)};
    ColProd<256, KPlain> bp{W, FH, FH, bn, colvec(W, FH, FH) && H % 4 == 0};   // B(n = i*H + j, c) = W[c*FH + n]
    B200_TRY(launch_packed("tc_cin_dx0", R, FH, bn, H, H, F, cdiv(C, BK), s, ap, bp, tc::EpRowDot{x_in, H, gx0, F},
                           passes, st));
  }
  {  // ---- gx_in[r, j] = sum_{(i,c)} (x0[r,i] gy[r,c]) W[c, i*H + j]
    const int bn = pick_bn_packed(H);
    KCin s = make_kcin(F, C, H, C);
    CinZProd ap{x0, gy, R, (C % 4 == 0) && aligned16(gy), false};
    ColProd<256, KCin> bp{W, FH, H, bn, colvec(W, FH, H) && H % 4 == 0};   // B(j, (i,c)) = W[c*FH + i*H + j]
    B200_TRY(launch_packed("tc_cin_dx", R, H, bn, bn, bn, cdiv(H, bn), F * cdiv(C, BK), s, ap, bp,
                           tc::EpMaskAcc{gx_in, H, nullptr, 0, gx_in_acc}, passes, st));
  }
  return B200REC_OK;
}

}  // namespace b200rec

#ifdef B200_TC_TRACE
extern "C" int b200rec_debug_tc_trace(long long* out, int n) {
  return (int)cudaMemcpyFromSymbol(out, b200rec::tc::g_tc_trace, sizeof(long long) * (size_t)n);
}
extern "C" int b200rec_debug_tc_trace_dw(long long* out, int n) {
  return (int)cudaMemcpyFromSymbol(out, b200rec::tc::g_tc_trace_dw, sizeof(long long) * (size_t)n);
}
#endif
