// Model / table handle definitions (opaque through the C ABI).
#pragma once
#include <vector>

#include "kernels.h"

namespace b200rec {

struct MlpPlan {
  int in_dim = 0;
  std::vector<int> dims;
  bool head = false;
  long long off = 0, end = 0;
  std::vector<long long> w_off, b_off;
  void build(int in, const std::vector<int>& d, bool with_head, long long start);
};

struct RunArgs {
  int B = 0;
  long long nnz = 0;
  const int* index = nullptr;  // device; nullptr = canonical (index[i] = i / F)
  // resident source (gather fused into the forward)
  const int* feats = nullptr;
  const float* table_emb = nullptr;
  const float* table_w = nullptr;
  long long table_rows = 0;
  // flat source (rows already gathered)
  const float* w_nz = nullptr;
  const float* emb = nullptr;
  // dense params
  const float* bias = nullptr;
  const float* mats = nullptr;
  const float* targets = nullptr;  // nullptr => forward only
  float* preds = nullptr;
  // backward outputs
  float* dw_out = nullptr;     // [nnz]
  float* dE_out = nullptr;     // [nnz*K]  (may alias emb / X)
  float* dbias_out = nullptr;  // [1]
  float* gmats_out = nullptr;  // [mats_len]
  float* loss_out = nullptr;   // [1]
  const int* out_slot = nullptr;  // sharded path: per-nnz gradients are written to these slots
  // resident step: leave the per-nnz gradient producer (GradUtil.scala:7-42) to the fused segment
  // reduce (segsum.cu); run() then only records where its inputs are (Model::deferred)
  bool defer_sparse_bwd = false;
  // encoder-only call (the reference's HigherOrderEncoder / CINEncoder / CrossEncoder / ProductEncoder
  // .forward / .backward): `emb` is the encoder input [B, F*K], no sparse terms, no head
  bool encoder_only = false;
  float* enc_out = nullptr;         // [B]   ([B, fc[0]] for the PNN product encoder)
  const float* enc_grad = nullptr;  // gradOutput, same shape; nullptr => forward only
  float* enc_dx = nullptr;          // gradInput [B, F*K] out
};

struct DeferredGrad {       // inputs of the per-nnz gradient the fused scatter-add computes on the fly
  const float* X = nullptr;      // gathered rows [nnz,K]
  const float* S = nullptr;      // [B,K] or nullptr
  const float* dX = nullptr;     // [nnz,K] or nullptr
  const float* dlogit = nullptr; // [B]
  bool valid = false;
};

}  // namespace b200rec

struct b200rec_table_s {
  long long rows = 0;
  int dim = 0;
  int device = 0;
  cudaStream_t stream = nullptr;
  b200rec::DevBuf emb, w, stage_i, stage_e, stage_w, err;
  b200rec::DevBuf s1e, s2e, s1w, s2w;  // optimizer slots (allocated on first use)
};

struct b200rec_model_s {
  int kind = 0, F = 0, K = 0, D = 0, depth = 0, device = 0;
  std::vector<int> fc, cin, pairs;
  long long mats_len = 0;
  b200rec::MlpPlan mlp;
  std::vector<long long> cin_w, cin_b;
  long long out_off = 0;
  cudaStream_t stream = nullptr, side = nullptr, side2 = nullptr, side3 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_fork2 = nullptr, ev_join2 = nullptr;
  cudaEvent_t ev_fork3 = nullptr, ev_join3 = nullptr;
  // auxiliary stream of the dense branch: work that is off the dx chain (the gradInput weight images,
  // the bias-gradient column sums) runs beside the GEMMs; always joined before run() returns
  cudaStream_t aux = nullptr;
  cudaEvent_t ev_aux_fork = nullptr, ev_aux_pack = nullptr, ev_aux_join = nullptr, ev_aux_cs[2] = {nullptr, nullptr};
  bool aux_pack_pending = false, aux_open = false;
  b200rec::DevBuf scratch_aux, wpack_aux;
  // input prefetch of the host-facing step (b200rec_stage_batch / b200rec_step_staged): two staged
  // batches at most, copied on their own stream while the previous step computes
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_staged[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};
  b200rec::DevBuf stage_f[2], stage_t[2];
  int stage_B[2] = {0, 0}, stage_head = 0, stage_count = 0;
  bool stage_used[2] = {false, false};
  // steps enqueued by b200rec_step_staged_async whose loss has not been waited for yet (<= 2)
  cudaEvent_t ev_done[2] = {nullptr, nullptr};
  int async_B[2] = {0, 0}, async_head = 0, async_count = 0;
  float* h_scal = nullptr;  // pinned 16 words: loss, dbias, n_unique, -, err, sorted
  bool params_set = false;
  // CUDA graph of the resident step (one per (B, table, gemm_mode)); captured after one eager warm-up
  cudaGraphExec_t graph_exec = nullptr;
  int graph_B = 0, graph_mode = -1, graph_nodes = 0, graph_warm_B = 0;
  long long graph_epoch = -1, graph_warm_epoch = -1;   // g_alloc_epoch at capture / at the eager warm-up
  const void* graph_table = nullptr;
  bool graph_enabled = true;
  int gemm_mode = 0;  // 0 fp32 SIMT, 1 split-precision tcgen05 (TF32 + bf16 corrections), 2 1xTF32 tcgen05
  b200rec::DeferredGrad deferred;
  bool fused_scatter = false;    // resident step: per-nnz gradient computed inside the segment reduce (measured slower: off)
  bool keep_nnz_grads = false;   // resident step: also materialise the per-nnz gradients (b200rec_step_nnz_grad_ptrs)
  int last_B = 0;
  long long last_nnz = 0;
  using DevBuf = b200rec::DevBuf;
  DevBuf p_bias, p_mats, gmats, scal;
  DevBuf d_feats, d_targets, d_index, stage_a, stage_b;
  DevBuf X, wnz, S, first, second, branch, preds, dlogit, dXd, dw, gA, gB, scratch;
  DevBuf uniq, G, gwU, wpack, wpack_mlp;
  b200rec::PrePack prepack;
  DevBuf s1m, s2m;  // optimizer slots of [mats | bias]
  DevBuf enc_go, enc_o, enc_dx;   // encoder-level calls: gradOutput in, output / gradInput out
  DevBuf p2p_ctr;   // [0] block-completion counter of the peer-exchange kernels, [1] device step counter
  bool join_pending[3] = {false, false, false};   // a side-stream sort was forked and not joined yet
  bool capturing = false;                        // b200rec_capture_begin .. _end
  long long capture_l0 = 0;
  struct UserGraph { cudaGraphExec_t exec; int nodes; long long epoch; };
  std::vector<UserGraph> user_graphs;   // executable, kernel nodes, g_alloc_epoch at capture
  DevBuf x0, gx0, gy, gnA, gnB, pooled, gpooled;
  DevBuf xL, s_cross, g_xL;
  DevBuf ip, gip, pre, hbuf;
  std::vector<DevBuf> acts, xl;
  // sort workspaces: 0 the batch's ids, 1 the ids an owner received, 2 the NEXT batch's ids (prefetch)
  b200rec::SegSumWorkspace seg, seg2, seg3;
  b200rec::SegSumWorkspace& ws_of(int i) { return i == 0 ? seg : (i == 1 ? seg2 : seg3); }
  cudaStream_t side_of(int i) const { return i == 0 ? side : (i == 1 ? side2 : side3); }
  cudaEvent_t fork_of(int i) const { return i == 0 ? ev_fork : (i == 1 ? ev_fork2 : ev_fork3); }
  cudaEvent_t join_of(int i) const { return i == 0 ? ev_join : (i == 1 ? ev_join2 : ev_join3); }
  b200rec::ShardPlanWorkspace plan;

  int init(int kind, int F, int K, const int* fc, int n_fc, const int* cin, int n_cin, int depth,
           int device);
  void destroy();
  int reserve(int B, long long nnz);
  int mlp_forward(int B, const float* x_in, const float* mats, const float** last_out,
                  float* head_out, cudaStream_t st);
  int mlp_backward(int B, const float* x_in, const float* mats, float* gm, float* dx,
                   const float* in_mask, cudaStream_t st);
  int aux_join(cudaStream_t st);
  int mlp_head_backward(int B, const float* x_in, const float* mats, float* gm, const float* dlg,
                        float* dx_if_no_hidden, cudaStream_t st);
  int run(const b200rec::RunArgs& a, cudaStream_t st);
};

namespace b200rec {
using Model = ::b200rec_model_s;
using Table = ::b200rec_table_s;
}  // namespace b200rec
