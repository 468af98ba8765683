// Model handles: Internal<M>Model.forward / .backward of the reference on device buffers, and the
// resident step (lookup -> forward -> backward -> dedup scatter-add).
//
// Reference (relative to /root/reference/src/main/scala/io/yaochi/recommendation):
//   model/lr/LR.scala:42-90, model/deepfm/DeepFM.scala:51-125, model/xdeepfm/XDeepFM.scala:58-126,
//   model/dcn/DCN.scala:62-130, model/pnn/PNN.scala:56-132, model/encoder/*.scala,
//   util/{LayerUtil,BackwardUtil,GradUtil}.scala, model/ParRecModel.scala:439-478.
// The reference rebuilds every encoder and copies all of `mats` into fresh layers on every call
// and re-runs the forward inside backward (SURVEY B-11); here parameters are read in place and
// one pass computes forward and backward.
#include <vector>
#include <cstring>

#include "kernels.h"
#include "model.h"

namespace b200rec {

// ---- layout of a Linear stack inside mats (HigherOrderEncoder.scala:46-58) ----------------------
void MlpPlan::build(int in, const std::vector<int>& d, bool with_head, long long start) {
  in_dim = in; dims = d; head = with_head; off = start;
  w_off.clear(); b_off.clear();
  long long o = start;
  int prev = in;
  std::vector<int> all = d;
  if (with_head) all.push_back(1);
  for (int n : all) {
    w_off.push_back(o); o += (long long)prev * n;
    b_off.push_back(o); o += n;
    prev = n;
  }
  end = o;
}

static std::vector<int> mats_pairs_of(int kind, int F, int K, const std::vector<int>& fc,
                                      const std::vector<int>& cin, int depth) {
  // getMatsSize: DeepFM.scala:15-20, XDeepFM.scala:15-28, DCN.scala:15-32, PNN.scala:15-25
  std::vector<int> p;
  const int D = F * K;
  auto fcpairs = [&](std::vector<int> dims) {
    for (size_t i = 1; i < dims.size(); ++i) {
      p.push_back(dims[i - 1]); p.push_back(dims[i]); p.push_back(dims[i]); p.push_back(1);
    }
  };
  if (kind == B200REC_DEEPFM) {
    std::vector<int> d{D}; d.insert(d.end(), fc.begin(), fc.end()); d.push_back(1);
    fcpairs(d);
  } else if (kind == B200REC_XDEEPFM) {
    std::vector<int> d{D}; d.insert(d.end(), fc.begin(), fc.end());
    fcpairs(d);
    int h = F, sum = 0;
    for (int c : cin) { p.push_back(F * h); p.push_back(c); p.push_back(c); p.push_back(1); h = c; sum += c; }
    p.push_back(sum + fc.back()); p.push_back(1);
  } else if (kind == B200REC_DCN) {
    for (int i = 0; i < depth; ++i) { p.push_back(D); p.push_back(1); }
    for (int i = 0; i < depth; ++i) { p.push_back(1); p.push_back(1); }
    std::vector<int> d{D}; d.insert(d.end(), fc.begin(), fc.end());
    fcpairs(d);
    p.push_back(D + fc.back()); p.push_back(1);
  } else if (kind == B200REC_PNN) {
    const int P = F * (F - 1) / 2;
    p.push_back(D); p.push_back(fc[0]); p.push_back(P); p.push_back(fc[0]); p.push_back(1); p.push_back(1);
    std::vector<int> d(fc.begin(), fc.end()); d.push_back(1);
    fcpairs(d);
  }
  return p;
}

}  // namespace b200rec

using namespace b200rec;

int b200rec_model_s::init(int kind_, int F_, int K_, const int* fc_, int n_fc, const int* cin_, int n_cin,
                int depth_, int device_) {
  gemm_mode = g_default_gemm_mode;
  kind = kind_; F = F_; K = K_; D = F_ * K_; depth = depth_; device = device_;
  fc.assign(fc_, fc_ + n_fc);
  cin.assign(cin_, cin_ + n_cin);
  const bool needs_fc = kind == B200REC_DEEPFM || kind == B200REC_XDEEPFM || kind == B200REC_DCN ||
                        kind == B200REC_PNN;
  B200_REQUIRE(kind >= B200REC_LR && kind <= B200REC_PNN, B200REC_ERR_ARG, "unknown model kind %d", kind);
  if (kind != B200REC_LR) B200_REQUIRE(F > 0 && K > 0, B200REC_ERR_ARG, "nFields / embeddingDim must be positive");
  for (int d : fc) B200_REQUIRE(d > 0, B200REC_ERR_ARG, "fcDims must be positive");
  for (int d : cin) B200_REQUIRE(d > 0, B200REC_ERR_ARG, "cinDims must be positive");
  if (kind == B200REC_XDEEPFM) B200_REQUIRE(!fc.empty() && !cin.empty() && cin.size() <= 8, B200REC_ERR_ARG, "xDeepFM needs fcDims and 1..8 cinDims");
  if (kind == B200REC_DCN) B200_REQUIRE(!fc.empty() && depth >= 1 && depth <= 16, B200REC_ERR_ARG, "DCN needs fcDims and crossDepth in 1..16");
  if (kind == B200REC_PNN) B200_REQUIRE(!fc.empty() && F >= 2, B200REC_ERR_ARG, "PNN needs fcDims and >= 2 fields");
  (void)needs_fc;
  pairs = mats_pairs_of(kind, F, K, fc, cin, depth);
  mats_len = 0;
  for (size_t i = 0; i + 1 < pairs.size(); i += 2) mats_len += (long long)pairs[i] * pairs[i + 1];
  // parameter plans
  if (kind == B200REC_DEEPFM) {
    mlp.build(D, fc, true, 0);
  } else if (kind == B200REC_XDEEPFM) {
    mlp.build(D, fc, false, 0);
    long long o = mlp.end;
    int h = F;
    cin_w.clear(); cin_b.clear();
    for (int c : cin) { cin_w.push_back(o); o += (long long)F * h * c; cin_b.push_back(o); o += c; h = c; }
    out_off = o;
  } else if (kind == B200REC_DCN) {
    mlp.build(D, fc, false, (long long)depth * D + depth);
    out_off = mlp.end;
  } else if (kind == B200REC_PNN) {
    const int P = F * (F - 1) / 2, O = fc[0];
    std::vector<int> rest(fc.begin() + 1, fc.end());
    mlp.build(O, rest, true, (long long)D * O + (long long)P * O + 1);
  }
  B200_CUDA(cudaSetDevice(device));
  B200_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  B200_CUDA(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
  B200_CUDA(cudaStreamCreateWithFlags(&side2, cudaStreamNonBlocking));
  B200_CUDA(cudaStreamCreateWithFlags(&side3, cudaStreamNonBlocking));
  B200_CUDA(cudaStreamCreateWithFlags(&aux, cudaStreamNonBlocking));
  B200_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
  for (cudaEvent_t* e : {&ev_staged[0], &ev_staged[1], &ev_consumed[0], &ev_consumed[1], &ev_done[0], &ev_done[1]})
    B200_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  for (cudaEvent_t* e : {&ev_aux_fork, &ev_aux_pack, &ev_aux_join, &ev_aux_cs[0], &ev_aux_cs[1]})
    B200_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  B200_CUDA(cudaEventCreateWithFlags(&ev_fork3, cudaEventDisableTiming));
  B200_CUDA(cudaEventCreateWithFlags(&ev_join3, cudaEventDisableTiming));
  B200_CUDA(cudaEventCreateWithFlags(&ev_fork2, cudaEventDisableTiming));
  B200_CUDA(cudaEventCreateWithFlags(&ev_join2, cudaEventDisableTiming));
  B200_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
  B200_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
  B200_CUDA(cudaMallocHost(&h_scal, 64));
  B200_TRY(scal.reserve(64));
  B200_CUDA(cudaMemset(scal.p, 0, 64));
  B200_TRY(p_bias.reserve(4));
  B200_TRY(p_mats.reserve((size_t)(mats_len > 0 ? mats_len : 1) * 4));
  B200_TRY(gmats.reserve((size_t)(mats_len + 1) * 4));  // + the bias gradient at the end
  return B200REC_OK;
}

void b200rec_model_s::destroy() {
  cudaSetDevice(device);
  if (stream) cudaStreamSynchronize(stream);
  if (side) cudaStreamSynchronize(side);
  if (side2) cudaStreamSynchronize(side2);
  if (side3) cudaStreamSynchronize(side3);
  if (aux) cudaStreamSynchronize(aux);
  if (copy_stream) cudaStreamSynchronize(copy_stream);
  DevBuf* bufs[] = {&p_bias, &p_mats, &gmats, &scal, &d_feats, &d_targets, &d_index, &X, &wnz, &S,
                    &first, &second, &branch, &preds, &dlogit, &dXd, &dw, &gA, &gB, &scratch,
                    &uniq, &G, &gwU, &wpack, &wpack_mlp, &s1m, &s2m, &p2p_ctr, &x0, &gx0, &gy, &gnA, &gnB, &pooled, &gpooled, &xL, &s_cross,
                    &g_xL, &ip, &gip, &pre, &hbuf, &stage_a, &stage_b, &enc_go, &enc_o, &enc_dx};
  for (DevBuf* b : bufs) b->release();
  for (auto& b : acts) b.release();
  for (auto& b : xl) b.release();
  seg.release();
  seg2.release();
  seg3.release();
  plan.release();
  if (graph_exec) cudaGraphExecDestroy(graph_exec);
  for (auto& g : user_graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  graph_exec = nullptr;
  if (h_scal) cudaFreeHost(h_scal);
  if (ev_fork) cudaEventDestroy(ev_fork);
  if (ev_join) cudaEventDestroy(ev_join);
  if (stream) cudaStreamDestroy(stream);
  if (side) cudaStreamDestroy(side);
  if (side2) cudaStreamDestroy(side2);
  if (side3) cudaStreamDestroy(side3);
  if (aux) cudaStreamDestroy(aux);
  if (copy_stream) cudaStreamDestroy(copy_stream);
  for (cudaEvent_t e : {ev_staged[0], ev_staged[1], ev_consumed[0], ev_consumed[1], ev_done[0], ev_done[1]})
    if (e) cudaEventDestroy(e);
  for (int i = 0; i < 2; ++i) { stage_f[i].release(); stage_t[i].release(); }
  for (cudaEvent_t e : {ev_aux_fork, ev_aux_pack, ev_aux_join, ev_aux_cs[0], ev_aux_cs[1]})
    if (e) cudaEventDestroy(e);
  scratch_aux.release();
  wpack_aux.release();
  if (ev_fork3) cudaEventDestroy(ev_fork3);
  if (ev_join3) cudaEventDestroy(ev_join3);
  if (ev_fork2) cudaEventDestroy(ev_fork2);
  if (ev_join2) cudaEventDestroy(ev_join2);
}

int b200rec_model_s::reserve(int B, long long nnz) {
  const size_t f = sizeof(float);
  const size_t b = (size_t)(B > 0 ? B : 1), n = (size_t)(nnz > 0 ? nnz : 1);
  B200_TRY(first.reserve(b * f)); B200_TRY(second.reserve(b * f)); B200_TRY(branch.reserve(b * f));
  B200_TRY(preds.reserve(b * f)); B200_TRY(dlogit.reserve(b * f));
  B200_TRY(dw.reserve(n * f));
  if (kind == B200REC_LR) return B200REC_OK;
  B200_TRY(S.reserve(b * K * f));
  int maxw = D;
  for (int d : fc) maxw = d > maxw ? d : maxw;
  if (kind != B200REC_FM) {
    B200_TRY(dXd.reserve(b * D * f));
    B200_TRY(gA.reserve(b * maxw * f)); B200_TRY(gB.reserve(b * maxw * f));
    acts.resize(mlp.dims.size());
    for (size_t l = 0; l < mlp.dims.size(); ++l) B200_TRY(acts[l].reserve(b * mlp.dims[l] * f));
  }
  if (kind == B200REC_XDEEPFM) {
    const size_t R = b * K;
    int maxc = F, sum = 0;
    for (int c : cin) { maxc = c > maxc ? c : maxc; sum += c; }
    B200_TRY(x0.reserve(R * F * f)); B200_TRY(gx0.reserve(R * F * f));
    B200_TRY(gy.reserve(R * maxc * f)); B200_TRY(gnA.reserve(R * maxc * f)); B200_TRY(gnB.reserve(R * maxc * f));
    xl.resize(cin.size());
    for (size_t l = 0; l < cin.size(); ++l) B200_TRY(xl[l].reserve(R * cin[l] * f));
    B200_TRY(pooled.reserve(b * sum * f)); B200_TRY(gpooled.reserve(b * sum * f));
  } else if (kind == B200REC_DCN) {
    B200_TRY(xL.reserve(b * D * f)); B200_TRY(g_xL.reserve(b * D * f));
    B200_TRY(s_cross.reserve(b * depth * f));
  } else if (kind == B200REC_PNN) {
    const size_t P = (size_t)F * (F - 1) / 2;
    B200_TRY(ip.reserve(b * P * f)); B200_TRY(gip.reserve(b * P * f));
    B200_TRY(pre.reserve(b * fc[0] * f)); B200_TRY(hbuf.reserve(b * fc[0] * f));
  }
  return B200REC_OK;
}

// ---- MLP tower ---------------------------------------------------------------------------------
// forward: returns the last hidden activation (or x_in when there is no hidden layer); when the
// plan has a head, head_out[B] = last . w_o + b_o.
int b200rec_model_s::mlp_forward(int B, const float* x_in, const float* mats, const float** last_out,
                       float* head_out, cudaStream_t st) {
  const float* h = x_in;
  int in = mlp.in_dim;
  for (size_t l = 0; l < mlp.dims.size(); ++l) {
    float* y = acts[l].as<float>();
    B200_TRY(linear_fwd(B, mlp.dims[l], in, h, mats + mlp.w_off[l], mats + mlp.b_off[l], true, y, st, gemm_mode));
    h = y;
    in = mlp.dims[l];
  }
  if (mlp.head) {
    const size_t hl = mlp.dims.size();
    B200_TRY(gemv_rows(B, in, h, in, mats + mlp.w_off[hl], mats + mlp.b_off[hl], false, head_out, st));
  }
  if (last_out) *last_out = h;
  return B200REC_OK;
}

// backward.  On entry gA holds the gradient w.r.t. the pre-activation of the last hidden layer
// (already ReLU-masked) -- or, with no hidden layer, nothing.  Writes the parameter gradients at
// the parameters' offsets in gm and the input gradient to dx (masked by in_mask > 0 if given).
int b200rec_model_s::mlp_backward(int B, const float* x_in, const float* mats, float* gm, float* dx,
                        const float* in_mask, cudaStream_t st) {
  float* g = gA.as<float>();
  float* g2 = gB.as<float>();
  const int L = (int)mlp.dims.size();
  for (int l = L - 1; l >= 0; --l) {
    const int out = mlp.dims[l];
    const int in = l == 0 ? mlp.in_dim : mlp.dims[l - 1];
    const float* xp = l == 0 ? x_in : acts[l - 1].as<float>();
    // parameter gradients (weight GEMM + split-K reduce + bias column sums) on the auxiliary stream:
    // only the gradInput GEMMs are on the chain to the embedding gradients.  (Under the per-kernel
    // profiler everything stays on one stream so that every kernel is timed alone.)
    cudaStream_t ax = tl_prof ? st : aux;
    if (ax != st) {
      B200_CUDA(cudaEventRecord(ev_aux_fork, st));
      B200_CUDA(cudaStreamWaitEvent(aux, ev_aux_fork, 0));
      aux_open = true;
    }
    {
      PackScope aux_pack(&wpack_aux);   // the main stream's GEMMs pack into wpack meanwhile
      B200_TRY(linear_bwd_params(B, out, in, xp, g, 1.0f, false, gm + mlp.w_off[l], gm + mlp.b_off[l],
                                 scratch_aux, ax, gemm_mode));
    }
    B200_CUDA(cudaEventRecord(ev_aux_cs[l & 1], ax));
    if (l > 0) {
      // g2 held the gradient of layer l + 1: its parameter gradients must have been computed
      if (l + 1 < L) B200_CUDA(cudaStreamWaitEvent(st, ev_aux_cs[(l + 1) & 1], 0));
      B200_TRY(linear_bwd_input(B, out, in, g, mats + mlp.w_off[l], acts[l - 1].as<float>(), g2, false, st, gemm_mode));
      float* t = g; g = g2; g2 = t;
    } else if (dx) {
      B200_TRY(linear_bwd_input(B, out, in, g, mats + mlp.w_off[l], in_mask, dx, false, st, gemm_mode));
    }
  }
  return B200REC_OK;
}

// everything forked to the auxiliary stream is joined into st (a captured graph must end joined)
int b200rec_model_s::aux_join(cudaStream_t st) {
  if (!aux_open) return B200REC_OK;
  B200_CUDA(cudaEventRecord(ev_aux_join, aux));
  B200_CUDA(cudaStreamWaitEvent(st, ev_aux_join, 0));
  aux_open = false;
  aux_pack_pending = false;
  return B200REC_OK;
}

// head layer of an MLP (Linear(last -> 1) with bias) backward: fills gA (masked) + param grads
int b200rec_model_s::mlp_head_backward(int B, const float* x_in, const float* mats, float* gm,
                             const float* dlg, float* dx_if_no_hidden, cudaStream_t st) {
  const size_t hl = mlp.dims.size();
  const int last = hl ? mlp.dims[hl - 1] : mlp.in_dim;
  const float* a = hl ? acts[hl - 1].as<float>() : x_in;
  // W_o (1 x last) and b_o are adjacent in mats (LayerUtil.scala:12-24): one fused pass fills both
  // gradients and the (ReLU-masked) gradient for the layer below
  float* gbelow = hl ? gA.as<float>() : dx_if_no_hidden;
  B200_TRY(head_layer_bwd(B, last, dlg, a, mats + mlp.w_off[hl], hl > 0, gbelow, gm + mlp.w_off[hl], scratch, st));
  return B200REC_OK;
}

// ---- the whole forward (+ backward) on device buffers -------------------------------------------
int b200rec_model_s::run(const RunArgs& a, cudaStream_t st) {
  PackScope pack_scope(&wpack);
  struct PrePackScope { PrePack* prev = tl_prepack; ~PrePackScope() { tl_prepack = prev; } } prepack_scope;
  prepack.n = 0;
  prepack.blob = &wpack_mlp;
  tl_prepack = nullptr;
  const int B = a.B;
  const long long nnz = a.nnz;
  const bool enc = a.encoder_only;
  const bool train = enc ? a.enc_grad != nullptr : a.targets != nullptr;
  const bool has_emb = kind != B200REC_LR;
  if (has_emb) B200_REQUIRE(nnz == (long long)B * F, B200REC_ERR_SHAPE,
                            "nnz %lld != batchSize %d * nFields %d (Reshape to [B,F,K] would fail)", nnz, B, F);
  B200_TRY(reserve(B, nnz));
  int* err = scal.as<int>() + 4;  // err[0] DevErr, err[1] sortedness
  B200_CUDA(cudaMemsetAsync(err, 0, 2 * sizeof(int), st));
  const bool second_order = kind == B200REC_FM || kind == B200REC_DEEPFM;
  const bool canonical = a.index == nullptr;
  const float* mats = a.mats;

  // ---- sparse forward: (gather) + first-order + second-order -----------------------------------
  struct PhaseScope { const char* prev = tl_tag; ~PhaseScope() { tl_tag = prev; } } phase_scope;
  auto phase = [](const char* name) { tl_tag = name; };
  phase("gather_fm_fwd");
  const float* Xp = a.emb;
  if (enc) {
    // encoder-only: nothing sparse to do
  } else if (a.table_emb) {
    SparseFwd s;
    s.B = B; s.F = F; s.K = has_emb ? K : 0; s.rows = a.table_rows; s.feats = a.feats;
    s.table = a.table_emb; s.wtable = a.table_w;
    if (has_emb) { B200_TRY(X.reserve((size_t)nnz * K * sizeof(float))); s.X = X.as<float>(); Xp = s.X; }
    s.first = first.as<float>();
    if (second_order) s.second = second.as<float>();
    if (second_order && train) s.S = S.as<float>();
    s.err = err;
    if (!has_emb) {  // LR from a table: gather the weights, then sum canonically
      B200_TRY(wnz.reserve((size_t)nnz * sizeof(float)));
      B200_TRY(lookup_rows(a.table_rows, 4, nnz, a.feats, a.table_emb, a.table_w, nullptr, wnz.as<float>(), err, st));
      SparseFwd t; t.B = B; t.F = (int)(B ? nnz / B : 0); t.K = 0; t.w_in = wnz.as<float>(); t.first = first.as<float>();
      B200_TRY(sparse_fwd(t, st));
    } else {
      B200_TRY(sparse_fwd(s, st));
    }
  } else {
    if (canonical) {
      SparseFwd s;
      s.B = B; s.F = has_emb ? F : (int)(B ? nnz / B : 0); s.K = has_emb ? K : 0;
      s.emb_in = a.emb; s.w_in = a.w_nz; s.first = first.as<float>();
      if (second_order) s.second = second.as<float>();
      if (second_order && train) s.S = S.as<float>();
      if (!has_emb) B200_REQUIRE(B == 0 || nnz % B == 0, B200REC_ERR_SHAPE, "canonical index needs nnz %% batchSize == 0");
      B200_TRY(sparse_fwd(s, st));
    } else {
      B200_TRY(scatter_fwd(B, 1, nnz, a.w_nz, a.index, first.as<float>(), err, st));
      if (second_order) {
        SparseFwd s;
        s.B = B; s.F = F; s.K = K; s.emb_in = a.emb; s.second = second.as<float>();
        if (train) s.S = S.as<float>();
        B200_TRY(sparse_fwd(s, st));
      }
    }
  }

  // ---- dense branch forward ---------------------------------------------------------------------
  phase("dense_fwd");
  if (gemm_mode && !mlp.dims.empty()) {
    // every Linear weight of the tower, in the forward and (training) the gradInput stage layout, in
    // two launches instead of one per GEMM
    cudaStream_t st_dx = st;
    if (train && !tl_prof) {   // the gradInput images are first needed in the backward: pack them beside the forward
      B200_CUDA(cudaEventRecord(ev_aux_fork, st));
      B200_CUDA(cudaStreamWaitEvent(aux, ev_aux_fork, 0));
      st_dx = aux;
      aux_open = true;
    }
    B200_TRY(tc_prepack_linear(prepack, mats, mlp.in_dim, mlp.dims.data(), (int)mlp.dims.size(),
                               mlp.w_off.data(), train, st, st_dx));
    if (st_dx != st) {
      B200_CUDA(cudaEventRecord(ev_aux_pack, aux));
      aux_pack_pending = true;
    }
    tl_prepack = &prepack;
  }
  Head h;
  h.B = B;
  h.br[h.n_br++] = first.as<float>();
  if (second_order) h.br[h.n_br++] = second.as<float>();
  const float* last = nullptr;
  float* br = branch.as<float>();
  int csum = 0;
  for (int c : cin) csum += c;
  if (kind == B200REC_DEEPFM) {
    B200_TRY(mlp_forward(B, Xp, mats, &last, br, st));
    h.br[h.n_br++] = br;
  } else if (kind == B200REC_XDEEPFM) {
    const int R = B * K;
    B200_TRY(mlp_forward(B, Xp, mats, &last, nullptr, st));
    B200_TRY(cin_transpose_in(B, F, K, Xp, x0.as<float>(), st));
    const float* xin = x0.as<float>();
    int hdim = F, col = 0;
    for (size_t l = 0; l < cin.size(); ++l) {
      B200_TRY(cin_layer_fwd(R, F, hdim, cin[l], x0.as<float>(), xin, mats + cin_w[l], mats + cin_b[l], xl[l].as<float>(), st, gemm_mode));
      B200_TRY(cin_pool(B, K, cin[l], xl[l].as<float>(), pooled.as<float>(), csum, col, st));
      xin = xl[l].as<float>();
      hdim = cin[l];
      col += cin[l];
    }
    // outputModule: Linear([pooled, dnn] -> 1), no bias (CINEncoder.scala:167-176)
    B200_TRY(gemv_rows(B, csum, pooled.as<float>(), csum, mats + out_off, nullptr, false, br, st));
    B200_TRY(gemv_rows(B, fc.back(), last, fc.back(), mats + out_off + csum, nullptr, true, br, st));
    h.br[h.n_br++] = br;
  } else if (kind == B200REC_DCN) {
    const float* cw = mats;
    const float* cc = mats + (long long)depth * D;
    B200_TRY(cross_fwd(B, D, depth, Xp, cw, cc, xL.as<float>(), s_cross.as<float>(), st));
    B200_TRY(mlp_forward(B, Xp, mats, &last, nullptr, st));
    B200_TRY(gemv_rows(B, D, xL.as<float>(), D, mats + out_off, nullptr, false, br, st));
    B200_TRY(gemv_rows(B, fc.back(), last, fc.back(), mats + out_off + D, nullptr, true, br, st));
    h.br[h.n_br++] = br;
  } else if (kind == B200REC_PNN) {
    const int P = F * (F - 1) / 2, O = fc[0];
    const float* wz = mats;
    const float* wp = mats + (long long)D * O;
    const float* c0 = wp + (long long)P * O;
    B200_TRY(linear_fwd(B, O, D, Xp, wz, nullptr, false, pre.as<float>(), st, gemm_mode));
    B200_TRY(pnn_ip_fwd(B, F, K, Xp, ip.as<float>(), st));
    B200_TRY(pnn_lp_fwd(B, P, O, ip.as<float>(), wp, pre.as<float>(), c0, hbuf.as<float>(), st, gemm_mode));
    if (!enc) {   // (the ProductEncoder alone ends here; PNN.scala:77-79 stacks a HigherOrderEncoder on it)
      B200_TRY(mlp_forward(B, hbuf.as<float>(), mats, &last, br, st));
      h.br[h.n_br++] = br;
    }
  }
  if (enc) {
    const size_t n_out = kind == B200REC_PNN ? (size_t)B * fc[0] : (size_t)B;
    const float* src = kind == B200REC_PNN ? hbuf.as<float>() : br;
    if (a.enc_out) B200_CUDA(cudaMemcpyAsync(a.enc_out, src, n_out * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (!train) return aux_join(st);
  }

  // ---- head ---------------------------------------------------------------------------------------
  if (!enc) {
    phase("head");
    h.bias = a.bias;
    h.targets = a.targets;
    h.preds = a.preds ? a.preds : preds.as<float>();
    h.dlogit = dlogit.as<float>();
    h.loss = a.loss_out;
    h.dbias = a.dbias_out;
    if (a.gmats_out == gmats.as<float>()) h.dbias2 = gmats.as<float>() + mats_len;  // [mats grad | bias grad]
    B200_TRY(head_run(h, scratch, st));
    if (!train) return aux_join(st);
  }

  // ---- dense branch backward ----------------------------------------------------------------------
  phase("dense_bwd");
  if (aux_pack_pending) {
    B200_CUDA(cudaStreamWaitEvent(st, ev_aux_pack, 0));
    aux_pack_pending = false;
  }
  const float* dlg = enc ? a.enc_grad : dlogit.as<float>();   // an encoder's gradOutput is [B,1]
  float* gm = a.gmats_out;
  float* dxd = nullptr;  // dense-branch gradient w.r.t. the embedding input
  if (kind == B200REC_DEEPFM) {
    dxd = dXd.as<float>();
    B200_TRY(mlp_head_backward(B, Xp, mats, gm, dlg, dxd, st));
    B200_TRY(mlp_backward(B, Xp, mats, gm, dxd, nullptr, st));
  } else if (kind == B200REC_XDEEPFM) {
    const int R = B * K;
    dxd = dXd.as<float>();
    // output Linear (no bias): gW_out = dlogit^T [pooled, dnn];  grads of its two inputs
    B200_TRY(wcolsum(B, csum, dlg, pooled.as<float>(), csum, gm + out_off, scratch, st));
    B200_TRY(wcolsum(B, fc.back(), dlg, last, fc.back(), gm + out_off + csum, scratch, st));
    B200_TRY(outer_rows(B, csum, dlg, mats + out_off, nullptr, 0, gpooled.as<float>(), csum, st));
    B200_TRY(outer_rows(B, fc.back(), dlg, mats + out_off + csum, last, fc.back(), gA.as<float>(), fc.back(), st));
    B200_TRY(mlp_backward(B, Xp, mats, gm, dxd, nullptr, st));
    // CIN reverse pass (CINEncoder.scala:76-86 as one sweep)
    B200_CUDA(cudaMemsetAsync(gx0.p, 0, (size_t)R * F * sizeof(float), st));
    const float* g_next = nullptr;
    float* gn_cur = gnA.as<float>();
    float* gn_other = gnB.as<float>();
    int col = csum;
    for (int l = (int)cin.size() - 1; l >= 0; --l) {
      const int C = cin[l];
      const int hdim = l == 0 ? F : cin[l - 1];
      const float* xin = l == 0 ? x0.as<float>() : xl[l - 1].as<float>();
      col -= C;
      B200_TRY(cin_gy(B, K, C, gpooled.as<float>(), csum, col, g_next, xl[l].as<float>(), gy.as<float>(), st));
      B200_TRY(cin_layer_bwd(R, F, hdim, C, x0.as<float>(), xin, mats + cin_w[l], gy.as<float>(),
                             gm + cin_w[l], gm + cin_b[l], gn_cur, gx0.as<float>(), scratch, st, gemm_mode));
      g_next = gn_cur;
      float* t = gn_cur; gn_cur = gn_other; gn_other = t;
    }
    // layer 1's input is x0 itself (CINEncoder.scala:85): fold its gradient in, un-transpose, add
    B200_TRY(cin_transpose_out(B, F, K, gx0.as<float>(), dxd, true, st));
    B200_TRY(cin_transpose_out(B, F, K, g_next, dxd, true, st));
  } else if (kind == B200REC_DCN) {
    dxd = dXd.as<float>();
    B200_TRY(wcolsum(B, D, dlg, xL.as<float>(), D, gm + out_off, scratch, st));
    B200_TRY(wcolsum(B, fc.back(), dlg, last, fc.back(), gm + out_off + D, scratch, st));
    B200_TRY(outer_rows(B, D, dlg, mats + out_off, nullptr, 0, g_xL.as<float>(), D, st));
    B200_TRY(outer_rows(B, fc.back(), dlg, mats + out_off + D, last, fc.back(), gA.as<float>(), fc.back(), st));
    B200_TRY(mlp_backward(B, Xp, mats, gm, dxd, nullptr, st));
    // cross backward writes into g_xL in place, then dxd += that
    B200_TRY(cross_bwd(B, D, depth, Xp, mats, mats + (long long)depth * D, s_cross.as<float>(),
                       g_xL.as<float>(), g_xL.as<float>(), gm, gm + (long long)depth * D, scratch, st));
    B200_TRY(axpy(B * (long long)D, g_xL.as<float>(), dxd, st));
  } else if (kind == B200REC_PNN) {
    const int P = F * (F - 1) / 2, O = fc[0];
    const float* wz = mats;
    const float* wp = mats + (long long)D * O;
    dxd = dXd.as<float>();
    float* gh = pre.as<float>();  // reuse: gradient w.r.t. the product layer pre-activation
    if (enc) {   // ProductEncoder.backward alone: gradOutput is [B, O], masked by the layer's own ReLU
      B200_TRY(relu_mask((long long)B * O, a.enc_grad, hbuf.as<float>(), gh, st));
    } else {
      B200_TRY(mlp_head_backward(B, hbuf.as<float>(), mats, gm, dlg, gh, st));
      if (mlp.dims.empty()) {
        B200_TRY(relu_mask((long long)B * O, gh, hbuf.as<float>(), gh, st));
      } else {
        B200_TRY(mlp_backward(B, hbuf.as<float>(), mats, gm, gh, hbuf.as<float>(), st));
      }
    }
    // ProductEncoder.backward :43-70
    B200_TRY(reduce_sum((long long)B * O, gh, 1.0f, gm + (long long)D * O + (long long)P * O, scratch, st));
    B200_TRY(linear_bwd_params(B, O, D, Xp, gh, 1.0f, false, gm, nullptr, scratch, st, gemm_mode));
    B200_TRY(linear_bwd_params(B, O, P, ip.as<float>(), gh, 1.0f, false, gm + (long long)D * O, nullptr, scratch, st, gemm_mode));
    B200_TRY(linear_bwd_input(B, O, D, gh, wz, nullptr, dxd, false, st, gemm_mode));
    B200_TRY(linear_bwd_input(B, O, P, gh, wp, nullptr, gip.as<float>(), false, st, gemm_mode));
    B200_TRY(pnn_ip_bwd(B, F, K, Xp, gip.as<float>(), dxd, true, st));
  }

  B200_TRY(aux_join(st));
  if (enc) {
    if (a.enc_dx && dxd)
      B200_CUDA(cudaMemcpyAsync(a.enc_dx, dxd, (size_t)B * D * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return B200REC_OK;
  }

  // ---- per-nnz gradients (GradUtil.scala:7-42) ------------------------------------------------------
  deferred.valid = false;
  if (a.defer_sparse_bwd && has_emb && canonical && !a.out_slot) {
    // the resident step's scatter-add computes them on the fly (segsum.cu: fused gradient producer)
    deferred.X = Xp; deferred.S = second_order ? S.as<float>() : nullptr; deferred.dX = dxd;
    deferred.dlogit = dlg; deferred.valid = true;
    return B200REC_OK;
  }
  phase("emb_grad");
  SparseBwd sb;
  sb.B = B; sb.F = has_emb ? F : (int)(B ? nnz / B : 0); sb.K = has_emb ? K : 0;
  sb.X = Xp; sb.S = second_order ? S.as<float>() : nullptr; sb.dX = dxd; sb.dlogit = dlg;
  sb.index = a.index; sb.dE = has_emb ? a.dE_out : nullptr; sb.dw = a.dw_out; sb.out_slot = a.out_slot;
  if (!has_emb && !canonical) {
    B200_TRY(scatter_bwd(B, 1, nnz, a.index, dlg, a.dw_out, err, st));
  } else {
    B200_TRY(sparse_bwd(sb, st));
  }
  return B200REC_OK;
}
