// Internal (C++) interface between the kernel translation units and the C ABI / model driver.
#pragma once
#include "common.cuh"

namespace b200rec {

// ---------------------------------------------------------------- sparse.cu (HBM-bound) ------
// Device-side status word written by kernels that validate ids / indices.
enum DevErr : int { DEV_OK = 0, DEV_BAD_ID = 1, DEV_BAD_INDEX = 2, DEV_PEER_TIMEOUT = 4, DEV_OVERFLOW = 8 };

struct SparseFwd {
  int B = 0, F = 0, K = 0;
  long long rows = 0;            // table rows (gather mode) for the id range check
  const int* feats = nullptr;    // [B*F] global ids; nullptr => rows are already in emb_in
  const float* table = nullptr;  // [rows,K]
  const float* wtable = nullptr; // [rows]
  const float* emb_in = nullptr; // [B*F*K]  (flat mode)
  const float* w_in = nullptr;   // [B*F]    (flat mode)
  float* X = nullptr;            // [B,F*K] gathered rows out (gather mode, optional)
  float* w_out = nullptr;        // [B*F]   gathered first-order weights out (optional)
  float* first = nullptr;        // [B]  sum_f w   (canonical index only)          (optional)
  float* second = nullptr;       // [B]  0.5/K sum_k (S^2 - Q)                     (optional)
  float* S = nullptr;            // [B,K] sum_f v  (saved for the backward)        (optional)
  int* err = nullptr;            // DevErr word
};
int sparse_fwd(const SparseFwd& a, cudaStream_t st);

struct SparseBwd {
  int B = 0, F = 0, K = 0;
  const float* X = nullptr;       // [B,F*K] rows (v)
  const float* S = nullptr;       // [B,K]  or nullptr when the model has no second-order term
  const float* dX = nullptr;      // [B,F*K] dense-branch input grad, or nullptr
  const float* dlogit = nullptr;  // [B]
  const int* index = nullptr;     // [B*F] or nullptr (canonical)
  float* dE = nullptr;            // [B*F*K] out (may alias X)
  float* dw = nullptr;            // [B*F] out
  const int* out_slot = nullptr;  // sharded path: row i of the gradients goes to slot out_slot[i]
};
int sparse_bwd(const SparseBwd& a, cudaStream_t st);

// nn/Scatter.scala generic forms (arbitrary index, n_output columns)
int scatter_fwd(int B, int n_out, long long n, const float* in, const int* index, float* out,
                int* err, cudaStream_t st);
int scatter_bwd(int B, int n_out, long long n, const int* index, const float* gout, float* gin,
                int* err, cudaStream_t st);
int lookup_rows(long long rows, int K, long long n, const int* feats, const float* table,
                const float* wtable, float* emb_out, float* w_out, int* err, cudaStream_t st);
// same with negative ids allowed (exchange padding): their output rows are zero
int lookup_rows_padded(long long rows, int K, long long n, const int* feats, const float* table,
                       const float* wtable, float* emb_out, float* w_out, int* err, cudaStream_t st);
// nn/Gather.scala / nn/DotProduct2.scala
int pair_gather_fwd(int B, int F, int P, int K, const float* in, const int* rows, const int* cols,
                    float* row_out, float* col_out, cudaStream_t st);
int pair_gather_bwd(int B, int F, int P, int K, const int* rows, const int* cols,
                    const float* g_row, const float* g_col, float* gin, cudaStream_t st);
int dot2_fwd(long long n, int K, const float* a, const float* b, float* out, cudaStream_t st);
int dot2_bwd(long long n, int K, const float* a, const float* b, const float* go, float* ga,
             float* gb, cudaStream_t st);

// ---------------------------------------------------------------- segsum.cu (scatter-add) ----
constexpr int RS_THREADS = 256, RS_TILE = 2048, RS_BINS = 256;   // radix sort: threads per tile, items per tile, 8-bit digits
struct SegSumWorkspace {
  DevBuf keys_a, keys_b, vals_a, vals_b, cub_tmp, seg_start, long_list, counters;
  long long cap_n = 0;
  int reserve(long long n);
  void release();
};
struct SegSum {
  long long n = 0;
  int K = 0;
  int key_bits = 31;              // radix-sort only the bits that can be set
  const int* feats = nullptr;     // [n]
  const float* dE = nullptr;      // [n,K] or nullptr
  const float* dw = nullptr;      // [n]   or nullptr
  int* unique = nullptr;          // [n] out (ascending)
  float* G = nullptr;             // [n,K] out (first U rows valid)
  float* gw = nullptr;            // [n]   out
  int* n_unique = nullptr;        // device int out
  bool drop_pad = false;          // keys equal to -1 (exchange padding) form no segment
  bool background = false;        // the sort half runs beside dense math (side stream): few fat blocks instead of one per tile
  // fused gradient producer (segsum.cu): rows are computed from the step's saved tensors instead of
  // being read from dE / dw:  dE[p] = (dlogit_b / K)(S_b - X[p]) + dX[p], dw[p] = dlogit_b, b = p / F
  bool fused = false;
  const float* fX = nullptr;      // [n,K] gathered rows        (needed when fS is given)
  const float* fdX = nullptr;     // [n,K] dense-branch input gradient, or nullptr
  const float* fS = nullptr;      // [B,K] sum over fields, or nullptr (no second-order term)
  const float* fdlogit = nullptr; // [B]
  int fF = 0;                     // non-zeros per sample
  float* keep_dE = nullptr;       // optional: also write the per-nnz gradients
  float* keep_dw = nullptr;
  const struct SegPush* push = nullptr;   // fused gradient push: sums go to the owners' buffers, not to G / gw
};
// sort half (depends only on feats: can run on a side stream while the dense math runs)
int segsum_sort(SegSumWorkspace& ws, const SegSum& a, cudaStream_t st);
// reduce half (needs dE / dw)
int segsum_reduce(SegSumWorkspace& ws, const SegSum& a, cudaStream_t st);
int segsum_inverse(SegSumWorkspace& ws, long long n, int* inv, cudaStream_t st);
// ---------------------------------------------------------------- shard.cu (row-sharded table) --
// owner(id) = (id + id / period) % world, local row = id / world (period is a multiple of world, so
// the `world` consecutive ids of a block rotate over the ranks: a bijection id <-> (owner, row)).
// period < 0 selects the reference's own partitioning instead: CONTIGUOUS RANGES of -period rows per rank
// (ColumnRangePartitioner, rec/model/ParRecModel.scala:77,81,98,116): owner = id / rows_per_rank.
__host__ __device__ inline int shard_owner(long long id, int world, long long period) {
  if (period < 0) {
    const long long o = id / (-period);
    return (int)(o < world ? o : world - 1);
  }
  return (int)((id + id / period) % world);
}
__host__ __device__ inline long long shard_local_row(long long id, int world, long long period) {
  if (period < 0) return id - (long long)shard_owner(id, world, period) * (-period);
  return id / world;
}
__host__ __device__ inline bool shard_period_ok(int world, long long period) {
  return period < 0 || (period >= world && period % world == 0);
}
struct ShardPlanWorkspace {
  DevBuf keys, keys_sorted, vals, perm, offsets, cub_tmp;
  int reserve(long long n);
  void release();
};
// send_ids[world*cap]: local row ids grouped by owner in slot order (-1 = padding);
// dst[n]: slot (owner*cap + position) of non-zero i; overflow[0] |= 1 when a bucket exceeds cap.
int shard_sort(ShardPlanWorkspace& ws, long long n, const int* n_dev, int world, long long period,
               const int* feats, cudaStream_t st);
int shard_plan(ShardPlanWorkspace& ws, long long n, int world, long long period, int cap,
               const int* feats, int* send_ids, int* dst, int* overflow, cudaStream_t st);
// ---------------------------------------------------------------- p2p.cu (NVLink peer exchange) --
constexpr int P2P_MAX = 8;
struct PeerI { int* p[P2P_MAX]; };
struct PeerF { float* p[P2P_MAX]; };
struct P2P {
  int world, rank, step;
  const int* step_ptr;       // device step counter: step <= 0 means "*step_ptr - step" (CUDA-graph replays
                             // carry no host step; -1 = the step after the current one, for prefetches)
  unsigned* block_counter;   // local, zero-initialised; wraps back to 0 through atomicInc
  int* flags[P2P_MAX];       // every rank's flags[3][world]
};
// Fused local pre-reduce + gradient push (segsum.cu): the in-order per-id sums of a rank's batch are stored
// straight into the owners' grad_in / gw_in slots (dst[seg] = owner * cap + slot, < 0: no slot) instead of
// a local array that a second kernel would copy out; the kernel ends with the phase-2 signal.
struct SegPush {
  P2P c;
  PeerF grad_in, gw_in;
  const int* dst = nullptr;   // [U] slot of every distinct id
  int cap = 0;
};
// status (may be NULL): sticky device word; DEV_PEER_TIMEOUT is set when a peer's flag does not arrive in time
int p2p_wait(const int* flags, int phase, int world, int step, const int* step_ptr, int* status, cudaStream_t st);
int p2p_begin_step(int* step_ctr, int* ids_next, long long n, cudaStream_t st);
// dense gradients: inout -> my symmetric buffer, then the sum over all ranks (rank order) back into inout
int p2p_allreduce(long long n, float* inout, const int* flags_local, const P2P& c, const PeerF& bufs,
                  const PeerF* outs, int* status, cudaStream_t st);
// n_dev (optional): device count of valid ids (<= n); ids beyond it are ignored
int p2p_plan(ShardPlanWorkspace& ws, long long n, const int* n_dev, long long period, int cap,
             const int* feats, int* dst, int* overflow, const P2P& c, const PeerI& ids_in, cudaStream_t st);
int p2p_compose(SegSumWorkspace& ws, long long n, const int* dst_unique, int* dst, cudaStream_t st);
int p2p_gather(long long rows, int K, int cap, const int* ids_in, const float* table, const float* wtable,
               const P2P& c, const PeerF& rows_in, const PeerF& w_in, int* err, cudaStream_t st);
int p2p_push_grads(long long n, const int* n_dev, int K, int cap, const int* dst, const float* dE,
                   const float* dw, const P2P& c, const PeerF& grad_in, const PeerF& gw_in, cudaStream_t st);
int table_init_uniform_sharded(float* table, float* wtable, long long rows, int K, uint64_t seed,
                               float lo, float hi, int rank, int world, long long period,
                               cudaStream_t st);
int apply_sgd(int K, long long rows, long long cap, const int* n_unique, const int* unique, const float* G,
              const float* gw, float lr, float* table, float* wtable, int* err, cudaStream_t st);

// ---------------------------------------------------------------- optim.cu ----------------------
// step_dev (may be NULL): device update counter used instead of `step` (CUDA-graph replays); ids outside
// [0, rows) are skipped and flagged in err (DEV_BAD_ID)
int opt_rows(int kind, int K, long long rows, long long cap, const int* n_unique, const int* unique,
             const float* G, const float* gw, float lr, float p1, float p2, long long step,
             const int* step_dev, float* table, float* wtable, float* s1e, float* s2e, float* s1w,
             float* s2w, int* err, cudaStream_t st);
int opt_dense(int kind, long long n, const float* g, float lr, float p1, float p2, long long step,
              const int* step_dev, float* w, float* s1, float* s2, cudaStream_t st);

// ---------------------------------------------------------------- dense.cu (SIMT fp32) --------
// y[M,N] = act(x[M,K] W[N,K]^T + b[N])   (BigDL Linear + optional ReLU)
// mode: 0 = fp32 FFMA (SIMT), 1 = 3xTF32 tcgen05, 2 = 1xTF32 tcgen05
int linear_fwd(int M, int N, int K, const float* x, const float* w, const float* b, bool relu,
               float* y, cudaStream_t st, int mode = 0);
// gx[M,K] = (gy[M,N] W[N,K]) (* (mask[M,K] > 0) if mask)
int linear_bwd_input(int M, int N, int K, const float* gy, const float* w, const float* mask,
                     float* gx, bool accumulate, cudaStream_t st, int mode = 0);
// gw[N,K] (+)= scale * gy[M,N]^T x[M,K] ; gb[N] (+)= scale * colsum(gy)   (split-K, fixed order)
int linear_bwd_params(int M, int N, int K, const float* x, const float* gy, float scale,
                      bool accumulate, float* gw, float* gb, DevBuf& scratch, cudaStream_t st,
                      int mode = 0);
// out[M] = x[M,K] . w[K] (+ b0)      (Linear(K -> 1))
int gemv_rows(int M, int K, const float* x, int ldx, const float* w, const float* b0,
              bool accumulate, float* out, cudaStream_t st);
// g[M,K] = d[M] (x) w[K]  (* (mask>0))   -- backward of Linear(K->1) w.r.t. its input
int outer_rows(int M, int K, const float* d, const float* w, const float* mask, int ldm,
               float* g, int ldg, cudaStream_t st);
// gw[K] = sum_m d[m] * x[m,k] ;  deterministic two-stage
int wcolsum(int M, int K, const float* d, const float* x, int ldx, float* gw, DevBuf& scratch,
            cudaStream_t st);
// Linear(K -> 1) backward fused: g (optional) = d (x) w (* (a > 0) if mask); gw_gb[K+1] = [sum d a | sum d]
int head_layer_bwd(int M, int K, const float* d, const float* a, const float* w, bool mask, float* g,
                   float* gw_gb, DevBuf& scratch, cudaStream_t st);
// sum of n floats into out[0] (deterministic two-stage); scale applied
int reduce_sum(long long n, const float* x, float scale, float* out, DevBuf& scratch,
               cudaStream_t st);

// ---------------------------------------------------------------- head.cu ---------------------
// logit = sum branches + bias ; p = sigmoid ; (optional) BCE loss + dlogit + dbias
struct Head {
  int B = 0;
  const float* br[4] = {nullptr, nullptr, nullptr, nullptr};  // up to 4 branches [B]
  int n_br = 0;
  const float* bias = nullptr;     // [1]
  const float* targets = nullptr;  // [B] or nullptr (forward only)
  float* preds = nullptr;          // [B]
  float* dlogit = nullptr;         // [B]
  float* loss = nullptr;           // [1]
  float* dbias = nullptr;          // [1]
  float* dbias2 = nullptr;         // second copy, adjacent to the mats gradient (one allreduce)
};
int head_run(const Head& h, DevBuf& scratch, cudaStream_t st);

// ---------------------------------------------------------------- cin.cu ----------------------
struct CinDims {
  int B, F, K;
  int n_layers;
  int H[9];  // H[0]=F, H[l]=cin_dims[l-1]
};
// out[cols, rows] = in[rows, cols]^T
int transpose2d(long long rows, int cols, const float* in, float* out, cudaStream_t st);
// x0[(b,k),f] = X[b,f,k]
int cin_transpose_in(int B, int F, int K, const float* X, float* x0, cudaStream_t st);
// ge[b,f,k] (+)= gx0[(b,k),f]
int cin_transpose_out(int B, int F, int K, const float* gx0, float* gE, bool accumulate,
                      cudaStream_t st);
// x_out[r,c] = relu( sum_{i,j} x0[r,i] x_in[r,j] W[c,i*H+j] + b[c] )
int cin_layer_fwd(int R, int F, int H, int C, const float* x0, const float* x_in, const float* W,
                  const float* b, float* x_out, cudaStream_t st, int mode = 0);
// pooled[b, col0 + c] = sum_k x[(b,k), c]
int cin_pool(int B, int K, int C, const float* x, float* pooled, int ld, int col0, cudaStream_t st);
// gy[r,c] = (gp[b, col0+c] + (g_next? g_next[r,c] : 0)) * (x_out[r,c] > 0)
int cin_gy(int B, int K, int C, const float* gp, int ld, int col0, const float* g_next,
           const float* x_out, float* gy, cudaStream_t st);
// layer backward: gW[c, i*H+j] = sum_r gy Z ; gb ; gx_in[r,j] ; gx0[r,i] +=
int cin_layer_bwd(int R, int F, int H, int C, const float* x0, const float* x_in, const float* W,
                  const float* gy, float* gW, float* gb, float* gx_in, float* gx0,
                  DevBuf& scratch, cudaStream_t st, int mode = 0);

// ---------------------------------------------------------------- cross.cu (DCN) ---------------
// forward: xL[B,D], s[L,B]
int cross_fwd(int B, int D, int L, const float* X, const float* w, const float* c, float* xL,
              float* s, cudaStream_t st);
// backward: g_xL[B,D] in -> dX[B,D] out, gw[L,D], gc[L]
int cross_bwd(int B, int D, int L, const float* X, const float* w, const float* c, const float* s,
              const float* g_xL, float* dX, float* gw, float* gc, DevBuf& scratch,
              cudaStream_t st);

// ---------------------------------------------------------------- pnn.cu -----------------------
// ip[b,p] = <v_i, v_j>, pairs i<j lexicographic
int pnn_ip_fwd(int B, int F, int K, const float* X, float* ip, cudaStream_t st);
// h = relu(prev + x W^T + c0)   (second product GEMM of the PNN layer)
int pnn_lp_fwd(int B, int P, int O, const float* ip, const float* wp, const float* prev,
               const float* c0, float* h, cudaStream_t st, int mode = 0);
// dX[b,i,:] (+)= sum_{j != i} gip[b,pair(i,j)] v_j   (j ascending = reference order)
int pnn_ip_bwd(int B, int F, int K, const float* X, const float* gip, float* dX, bool accumulate,
               cudaStream_t st);
// h = relu(a + b + c0)  /  g = gout * (h > 0)
int add_bias_relu(long long n, const float* a, const float* c0, float* h, cudaStream_t st);
int relu_mask(long long n, const float* g, const float* h, float* out, cudaStream_t st);
// y += x
int axpy(long long n, const float* x, float* y, cudaStream_t st);

// ---------------------------------------------------------------- tc_gemm.cu (tcgen05) --------
// passes: 3 = error-compensated 3xTF32 (fp32-class accuracy), 1 = plain TF32
int tc_linear_fwd(int M, int N, int K, const float* x, const float* w, const float* b, bool relu,
                  float* y, int passes, cudaStream_t st);
int tc_pnn_lp_fwd(int B, int P, int O, const float* ip, const float* wp, const float* prev,
                  const float* c0, float* h, int passes, cudaStream_t st);
int tc_linear_bwd_input(int M, int N, int K, const float* gy, const float* w, const float* mask,
                        float* gx, bool accumulate, int passes, cudaStream_t st);
int tc_linear_bwd_params(int M, int N, int K, const float* x, const float* gy, float scale,
                         bool accumulate, float* gw, float* gb, DevBuf& scratch, int passes,
                         cudaStream_t st);
int tc_prepack_linear(PrePack& pp, const float* mats, int in_dim, const int* dims, int n_layers,
                      const long long* w_off, bool with_dx, cudaStream_t st, cudaStream_t st_dx);
// out[n] (+)= scale * sum_m g[m,n]; part = COLSUM_CHUNKS*N floats of scratch (fixed-order two stage) -- dense.cu
constexpr int COLSUM_CHUNKS = 512;
int colsum(int M, int N, const float* g, float scale, bool accumulate, float* out, float* part,
           cudaStream_t st);
bool tc_cin_supported(int F, int H, int C);
int tc_cin_layer_fwd(int R, int F, int H, int C, const float* x0, const float* x_in, const float* W,
                     const float* b, float* x_out, int passes, cudaStream_t st);
int tc_cin_layer_bwd(int R, int F, int H, int C, const float* x0, const float* x_in, const float* W,
                     const float* gy, float* gW, float* gb, float* gx_in, float* gx0, bool gx_in_acc,
                     DevBuf& scratch, int passes, cudaStream_t st);
extern int g_default_gemm_mode;

// ---------------------------------------------------------------- table.cu ---------------------
int table_init_uniform(float* table, float* wtable, long long rows, int K, uint64_t seed, float lo,
                       float hi, long long row_offset, long long row_stride, cudaStream_t st);

}  // namespace b200rec
