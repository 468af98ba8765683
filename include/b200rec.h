/*
 * b200rec.h -- C ABI of libb200rec.so: the B200 (sm_100a) hot path of
 * yaochitc/recommendation-models behind the reference's own operator surface.
 *
 * Reference paths are relative to /root/reference/src/main/scala:
 *   nn/  = com/intel/analytics/bigdl/nn/        rec/ = io/yaochi/recommendation/
 *
 * Conventions
 *  - Every function returns an int status: 0 = ok, negative = error (enum below).
 *    b200rec_last_error() returns the message of the calling thread's last failure.
 *    Nothing throws or aborts across the ABI.  A Scala shim turns a non-zero status into
 *    IllegalArgumentException, which is what the reference's `require`s raise
 *    (nn/Scatter.scala:29-30).
 *  - The caller owns every array it passes; the callee never keeps a pointer after return.
 *  - Functions without a suffix take HOST pointers (what JNA hands over for Array[Float] /
 *    Array[Int]) and do the host<->device copies themselves on the handle's stream.
 *    Functions with the suffix _dev take DEVICE pointers and an explicit cudaStream_t (as
 *    void*); they never synchronise the host unless a scalar result is requested.
 *  - A handle is not thread safe; distinct handles are independent.  Every call sets the
 *    handle's CUDA device first, so JVM threads may call from anywhere.
 *  - Gradient write-back follows the reference: backward OVERWRITES weights / bias /
 *    embedding / mats with their gradients (rec/util/GradUtil.scala:7-42,
 *    rec/util/BackwardUtil.scala:6-42).
 *  - There is no CPU fallback.  Without a usable sm_100 device every compute entry point
 *    fails with B200REC_ERR_CUDA.
 */
#ifndef B200REC_H_
#define B200REC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200REC_ABI_VERSION 1

typedef struct b200rec_model_s* b200rec_model_t;
typedef struct b200rec_table_s* b200rec_table_t;

enum b200rec_status {
  B200REC_OK = 0,
  B200REC_ERR_ARG = -1,      /* null pointer, negative size, unknown kind ...            */
  B200REC_ERR_SHAPE = -2,    /* nnz != B*F where the reference's Reshape would fail       */
  B200REC_ERR_INDEX = -3,    /* index >= batchSize (nn/Scatter.scala:29) or id >= rows   */
  B200REC_ERR_CUDA = -4,     /* CUDA runtime error / no sm_100 device                    */
  B200REC_ERR_STATE = -5,    /* call order (e.g. backward_step before params were set)   */
  B200REC_ERR_NOMEM = -6
};

/* rec/optim/OptimUtils.scala:5-12 ("sgd" | "momentum" | "adagrad" | "adam") */
enum b200rec_optimizer {
  B200REC_OPT_SGD = 0,       /* AsyncSGD.scala:10-31      1 slot                             */
  B200REC_OPT_MOMENTUM = 1,  /* AsyncMomentum.scala:10-33 2 slots, p1 = momentum (0.9)       */
  B200REC_OPT_ADAGRAD = 2,   /* AsyncAdagrad.scala:10-33  2 slots, p1 = factor (0.9)         */
  B200REC_OPT_ADAM = 3       /* AsyncAdam.scala:10-36     3 slots, p1 = gamma (0.99), p2 = beta (0.9) */
};

/* rec/model/RecModelType.scala:3-9 decides which buffers exist; `kind` implies it. */
enum b200rec_kind {
  B200REC_LR = 0,       /* rec/model/lr/LR.scala:42-90            BIAS_WEIGHT                 */
  B200REC_FM = 1,       /* no class in the reference (SURVEY B-1) BIAS_WEIGHT_EMBEDDING       */
  B200REC_DEEPFM = 2,   /* rec/model/deepfm/DeepFM.scala:51-125   BIAS_WEIGHT_EMBEDDING_MATS  */
  B200REC_XDEEPFM = 3,  /* rec/model/xdeepfm/XDeepFM.scala:58-126                             */
  B200REC_DCN = 4,      /* rec/model/dcn/DCN.scala:62-130                                     */
  B200REC_PNN = 5       /* rec/model/pnn/PNN.scala:56-132                                     */
};

/* ---- library ------------------------------------------------------------------------- */
int b200rec_abi_version(void);
const char* b200rec_last_error(void);
/* Number of usable sm_100 devices (0 and B200REC_ERR_CUDA when there is none). */
int b200rec_device_count(int* count);
/* Kernels launched by this library in the calling process since load (bench.py's
 * gpu_launches claim is read from here). */
int b200rec_launch_count(int64_t* count);

/* Per-kernel device timing of the calling thread's launches (CUDA events on the launching stream).
 * The reference has wall-clock phase counters (rec/model/ParRecModel.scala:34-40,584-626); this is
 * their device-side equivalent.  _end synchronises the device and writes lines
 * "phase|kernel|launches|total_ms\n" into buf (truncated to cap; *needed = full size). */
int b200rec_profile_begin(void);
int b200rec_profile_end(char* buf, int64_t cap, int64_t* needed);
/* Median elapsed time (microseconds) of an event pair around an EMPTY kernel on `device`: what the
 * per-launch event bracketing of a profile pass adds to each kernel's time. */
int b200rec_profile_overhead_us(int device, float* us);

/* ---- model: Internal<M>Model (constructor args are the reference's) -------------------- */
/* DeepFM.scala:51-53, XDeepFM.scala:58-61, DCN.scala:62-65, PNN.scala:56-58.
 * fc_dims / cin_dims may be NULL when the count is 0.  `device` = CUDA ordinal. */
int b200rec_model_create(int kind, int n_fields, int embedding_dim,
                         const int* fc_dims, int n_fc,
                         const int* cin_dims, int n_cin,
                         int cross_depth, int device, b200rec_model_t* out);
int b200rec_model_destroy(b200rec_model_t m);
/* getMatsSize (DeepFM.scala:15-20, XDeepFM.scala:15-28, DCN.scala:15-32, PNN.scala:15-25):
 * writes up to `cap` ints of (in,out) pairs, returns the count in *n. */
int b200rec_model_mats_size(b200rec_model_t m, int* pairs, int cap, int* n);
/* sum(in*out) over the pairs (rec/model/ParRecModel.scala:107-113). */
int b200rec_model_mats_len(b200rec_model_t m, int64_t* len);
/* The handle's stream (cudaStream_t) so callers can order their own work against it. */
int b200rec_model_stream(b200rec_model_t m, void** stream);
/* Arithmetic of the dense contractions (BigDL Linear / MM -> MKL sgemm in the reference,
 * rec/util/LayerUtil.scala:22-24, rec/model/xdeepfm/CINEncoder.scala:152):
 *   0 = fp32 FFMA (SIMT), 1 = 3xTF32 error-compensated tcgen05 (fp32-class accuracy, default where a
 *   tensor-core kernel exists), 2 = single-pass TF32 tcgen05 (about 1e-3 relative; NOT parity grade). */
int b200rec_model_set_gemm_mode(b200rec_model_t m, int mode);
/* Process-wide default (new model handles and the stand-alone Linear module calls use it). */
int b200rec_set_default_gemm_mode(int mode);
/* Wait for the handle's streams and report a deferred index / id error of a *_dev call. */
int b200rec_model_sync(b200rec_model_t m);

/* Internal<M>Model.forward (e.g. DeepFM.scala:54-81): host arrays in, preds[B] out.
 * index[nnz] = COO row of each non-zero; weights[nnz]; bias[1]; embedding[nnz*K] (NULL for
 * LR); mats[mats_len] (NULL for LR/FM). */
int b200rec_forward(b200rec_model_t m, int batch_size, int64_t nnz, const int* index,
                    const float* weights, const float* bias, const float* embedding,
                    const float* mats, float* preds);
/* Internal<M>Model.backward (e.g. DeepFM.scala:83-124): returns the mean BCE loss in *loss
 * and overwrites weights / bias / embedding / mats with their gradients. */
int b200rec_backward(b200rec_model_t m, int batch_size, int64_t nnz, const int* index,
                     float* weights, float* bias, float* embedding, float* mats,
                     const float* targets, float* loss);
/* Same two calls on device memory.  `index` may be NULL = canonical COO rows (index[i] = i/F,
 * what SampleParser emits).  loss_dev is a device float. */
int b200rec_forward_dev(b200rec_model_t m, int batch_size, int64_t nnz, const int* index,
                        const float* weights, const float* bias, const float* embedding,
                        const float* mats, float* preds, void* stream);
int b200rec_backward_dev(b200rec_model_t m, int batch_size, int64_t nnz, const int* index,
                         float* weights, float* bias, float* embedding, float* mats,
                         const float* targets, float* loss_dev, void* stream);

/* ---- table: the device-resident stand-in for the PS matrices ---------------------------- */
/* ParRecModel.initMats(inputDim) / initMats(inputDim, embeddingDim) :74-105.  One table holds
 * the `embedding` matrix (rows x dim, row-major; the reference's PS layout is dim-major,
 * which is not observable through the ABI) and the first-order `weights` vector (rows).
 * row_begin/row_end: the contiguous or strided shard this table owns (see shard_mod). */
int b200rec_table_create(int64_t rows, int dim, int device, b200rec_table_t* out);
int b200rec_table_destroy(b200rec_table_t t);
/* Deterministic U(lo,hi) init from a counter hash: value(row,col) = f(seed, global_row*dim+col);
 * global_row = row_offset + local_row * row_stride (so a shard reproduces its slice of the
 * global table).  ParRecModel.initEmbedding :66-69 (Xavier on the PS; parity unpinned). */
int b200rec_table_init_uniform(b200rec_table_t t, uint64_t seed, float lo, float hi,
                               int64_t row_offset, int64_t row_stride);
/* Host <-> table rows [row0, row0+nrows). Either pointer may be NULL. */
int b200rec_table_write(b200rec_table_t t, int64_t row0, int64_t nrows,
                        const float* embedding, const float* weights);
int b200rec_table_read(b200rec_table_t t, int64_t row0, int64_t nrows,
                       float* embedding, float* weights);
/* Device pointers of the table storage (embedding rows, weights). */
int b200rec_table_ptrs(b200rec_table_t t, float** embedding, float** weights);

/* makeEmbeddings + makeWeights (ParRecModel.scala:300-306, :279-284):
 * embedding_out[i*K+j] = E[feats[i], j]; weights_out[i] = w[feats[i]].  Bit-exact copies. */
int b200rec_table_lookup(b200rec_table_t t, int64_t nnz, const int* feats,
                         float* embedding_out, float* weights_out);
int b200rec_table_lookup_dev(b200rec_table_t t, int64_t nnz, const int* feats,
                             float* embedding_out, float* weights_out, void* stream);

/* distinctIntIndices (ParRecModel.scala:337-345).  Output is sorted ascending (the reference's
 * order is hash order, i.e. unspecified).  unique_out needs room for nnz ints. */
int b200rec_distinct(int device, int64_t nnz, const int* feats, int* unique_out, int64_t* n_unique);

/* makeEmbeddingGrad + makeWeightsGrad (ParRecModel.scala:316-328, :293-298): per distinct id,
 * the rows of embedding_grad (and entries of weights_grad) are summed IN NNZ ORDER i=0..N-1 in
 * fp32 -- the order of Int2FloatOpenHashMap.addTo -- so the result is bit-identical to the
 * reference.  unique_out[U] ascending, emb_out[U*K], w_out[U]; buffers sized for nnz.
 * embedding_grad or weights_grad may be NULL (then the matching output is untouched). */
int b200rec_scatter_add(int device, int dim, int64_t nnz, const int* feats,
                        const float* embedding_grad, const float* weights_grad,
                        int* unique_out, float* emb_out, float* w_out, int64_t* n_unique);

/* ---- resident training / prediction step -------------------------------------------------
 * ParRecModel.optimizeBiasWeightEmbeddingMats :439-478 and predictBiasWeightEmbeddingMats
 * :555-567 without the PS RPC: lookup (gather) -> forward -> backward -> dedup scatter-add,
 * parameters and gradients resident in HBM. */
/* Dense params the PS would hold: bias[1] and mats[mats_len] (host).  makeBias/makeMats. */
int b200rec_model_set_params(b200rec_model_t m, const float* bias, const float* mats);
int b200rec_model_get_params(b200rec_model_t m, float* bias, float* mats);
/* One step from HOST ids: feats[B*F] and targets[B] are copied in, *loss is copied out.
 * Gradients stay on the device until fetched. */
int b200rec_step(b200rec_model_t m, b200rec_table_t t, int batch_size,
                 const int* feats, const float* targets, float* loss);
/* Same with DEVICE ids / targets; no host synchronisation; loss stays in the handle. */
int b200rec_step_dev(b200rec_model_t m, b200rec_table_t t, int batch_size,
                     const int* feats, const float* targets, void* stream);
/* The training step replays as a CUDA graph after one eager warm-up per (batch size, table); 0 turns
 * that off (plain stream launches). */
int b200rec_model_set_graph(b200rec_model_t m, int enabled);
/* Batch builder (host code): libsvm ("<label> <key>:<value> ...", format 0) or libffm
 * ("<label> <field>:<key>:<value> ...", format 1) text, one sample per line, to the COO arrays the
 * path takes -- rec/data/SampleParser.scala:23-51 / :53-85.  Non-zero i of sample r becomes
 * index[i] = r, feats[i] = key - 1 (keys are 1-based in files, :37 / :69), in file order
 * (sample-major).  fields (libffm) and values may be NULL.  With every output NULL the call only
 * counts (*n_samples, *nnz) so the caller can size the arrays.  Malformed text -> B200REC_ERR_ARG
 * naming the line (the reference throws NumberFormatException); a key outside [1, 2^31] ->
 * B200REC_ERR_INDEX.  `text` need not be NUL-terminated; blank lines are skipped. */
int b200rec_parse_samples(int format, const char* text, int64_t n_bytes, int64_t cap_samples,
                          int64_t cap_nnz, float* targets, int* index, int* feats, int* fields,
                          float* values, int64_t* n_samples, int64_t* nnz);

/* b200rec_step with input prefetch: _stage_batch copies a batch's ids and labels (pinned host memory,
 * or the copy does not overlap; the arrays must stay valid until the matching _step_staged returns)
 * to the device on a copy stream and returns at once; _step_staged runs the step on the batch staged
 * first and returns its loss.  Staging batch i+1 before stepping batch i hides the host->device copy
 * under the step.  At most two batches are staged (B200REC_ERR_STATE otherwise, or when none is). */
int b200rec_stage_batch(b200rec_model_t m, int batch_size, const int* feats, const float* targets);
int b200rec_step_staged(b200rec_model_t m, b200rec_table_t t, float* loss);
/* The same without waiting: _step_staged_async enqueues the step on the batch staged first plus the
 * read-back of its loss and returns; _step_wait blocks until the OLDEST enqueued step is done and
 * returns its loss (and the device status: bad ids surface here).  Up to two steps may be in flight,
 * so the loop "stage(i+1); step_staged_async(i); step_wait(i-1)" keeps the GPU busy while the host
 * reads every step's loss one step late. */
int b200rec_step_staged_async(b200rec_model_t m, b200rec_table_t t);
int b200rec_step_wait(b200rec_model_t m, b200rec_table_t t, float* loss);

/* Predict: preds[B] = sigmoid(logit) (ParRecModel.predict :519-533). */
int b200rec_predict(b200rec_model_t m, b200rec_table_t t, int batch_size,
                    const int* feats, float* preds);
int b200rec_predict_dev(b200rec_model_t m, b200rec_table_t t, int batch_size,
                        const int* feats, float* preds, void* stream);
/* Device pointers of the resident dense params (bias[1], mats[mats_len]) -- e.g. for an NCCL
 * broadcast / an optimizer kernel of the caller. */
int b200rec_model_param_ptrs(b200rec_model_t m, float** bias, float** mats);
/* Results of the last step (host copies).  Any pointer may be NULL.
 *  loss; n_unique; unique[U]; emb_grad[U*K]; w_grad[U]; bias_grad[1]; mats_grad[mats_len]. */
int b200rec_step_results(b200rec_model_t m, float* loss, int64_t* n_unique, int* unique,
                         float* emb_grad, float* w_grad, float* bias_grad, float* mats_grad);
/* Device pointers of the same (valid until the next step on this handle).  mats_grad points at
 * mats_len + 1 floats: the bias gradient is repeated at mats_grad[mats_len] so one allreduce covers
 * all dense gradients. */
int b200rec_step_result_ptrs(b200rec_model_t m, float** loss, int** n_unique, int** unique,
                             float** emb_grad, float** w_grad, float** bias_grad,
                             float** mats_grad);
/* Per-nnz gradients of the last step before dedup: dE[nnz*K], dw[nnz] (device pointers) -- what
 * GradUtil.embeddingGrad / weightsGrad leave in the `embedding` / `weights` buffers
 * (rec/util/GradUtil.scala:7-42). */
int b200rec_step_nnz_grad_ptrs(b200rec_model_t m, float** emb_grad, float** w_grad);
/* Variant of the resident step's scatter-add that computes the per-nnz gradient inside the segment
 * reduce instead of reading the array an elementwise kernel wrote (csrc/segsum.cu; bit-identical sums).
 * It saves the dE round trip through HBM but gathers two arrays at sorted-id positions instead of one;
 * measured slower on B200, so it is OFF by default.  keep_nnz_grads != 0 makes the fused
 * form store the per-nnz gradients too (b200rec_step_nnz_grad_ptrs stays valid). */
int b200rec_model_set_fused_scatter(b200rec_model_t m, int enabled, int keep_nnz_grads);
/* Split step for a row-sharded table: the caller gathered rows itself (all-to-all)
 * and hands over device buffers emb[B*F*K], w[B*F]; grads are written in place (per nnz). */
int b200rec_step_gathered_dev(b200rec_model_t m, int batch_size, float* emb, float* w,
                              const float* targets, void* stream);
/* makeEmbeddingGrad / makeWeightsGrad on DEVICE buffers (the owner-side reduce of a row-sharded
 * table): sort + in-order segment sums of (feats[nnz], emb_grad[nnz*dim], w_grad[nnz]) into
 * unique_out / emb_out / w_out (sized for nnz) and *n_unique_dev.  Uses the handle's workspace. */
int b200rec_segsum_dev(b200rec_model_t m, int dim, int64_t nnz, int key_bits, int drop_pad,
                       const int* feats, const float* emb_grad, const float* w_grad, int* unique_out,
                       float* emb_out, float* w_out, int* n_unique_dev, void* stream);

/* The two halves of b200rec_segsum_dev, for overlap: the sort needs only the ids and runs on the
 * handle's side stream forked from `stream`; the reduce joins it.  Pass the same arguments to both. */
/* `ws` selects one of the handle's three sort workspaces / side streams: 0 and 2 for a batch's own
 * ids (two, so the NEXT batch's ids can be sorted while the current step runs), 1 for the ids an
 * owner received.  The sorts overlap the dense math of the step. */
int b200rec_segsum_sort_dev(b200rec_model_t m, int ws, int dim, int64_t nnz, int key_bits, int drop_pad,
                            const int* feats, int* unique_out, int* n_unique_dev, void* stream);
int b200rec_segsum_reduce_dev(b200rec_model_t m, int ws, int dim, int64_t nnz, int key_bits, int drop_pad,
                              const int* feats, const float* emb_grad, const float* w_grad,
                              int* unique_out, float* emb_out, float* w_out, int* n_unique_dev,
                              void* stream);
/* make `stream` wait for the sort of workspace ws */
int b200rec_segsum_join_dev(b200rec_model_t m, int ws, void* stream);
/* inv_out[i] = position of non-zero i's id in the sorted distinct ids of workspace ws's last sort */
int b200rec_segsum_inverse_dev(b200rec_model_t m, int ws, int64_t nnz, int* inv_out, void* stream);

/* ---- row-sharded table: the PS pull / push as NCCL all-to-all of fixed-capacity slot buffers -------
 * The Angel PS range-shards the matrices over PS nodes (ColumnRangePartitioner,
 * rec/model/ParRecModel.scala:77,81,98,116); workers pull rows (:174-177,193-196) and push gradients
 * (:247-250,261-264).  Here GPU `owner(id) = (id + id / period) % world` holds row `id / world`
 * (period > 0, a multiple of world: balanced under per-field power-law ids), or -- period < 0, the
 * reference's own ColumnRangePartitioner layout -- GPU `id / (-period)` holds row `id % (-period)`:
 * contiguous ranges of -period rows per rank.
 * The collectives themselves are issued by the host (torch.distributed / NCCL) between these calls. */
/* Fill a shard: local row q of `rank` holds the hash-initialised values of its global id. */
int b200rec_table_init_uniform_sharded(b200rec_table_t t, uint64_t seed, float lo, float hi, int rank,
                                       int world, int64_t period);
/* Plan the exchange of one batch: send_ids[world*cap] = local rows grouped by owner in non-zero order
 * (-1 padding); dst[nnz] = slot of non-zero i; *overflow |= 1 if a bucket exceeds cap. */
int b200rec_shard_plan_dev(b200rec_model_t m, int64_t nnz, int world, int64_t period, int cap,
                           const int* feats, int* send_ids, int* dst, int* overflow, void* stream);
/* Owner-side gather of received slot ids (negative = padding -> zero rows). */
int b200rec_table_lookup_padded_dev(b200rec_table_t t, int64_t n, const int* local_rows,
                                    float* embedding_out, float* weights_out, void* stream);
/* forward + backward where the rows of non-zero i are rows_emb[slots[i]] / rows_w[slots[i]] (the
 * buffers the all-to-all returned) and its gradients are written to grad_*_slots[slots[i]] (the
 * buffers the next all-to-all sends to the owners).  Dense gradients / loss stay in the handle. */
int b200rec_step_rows_dev(b200rec_model_t m, int batch_size, const int* slots, const float* rows_emb,
                          const float* rows_w, int64_t n_rows, const float* targets,
                          float* grad_emb_slots, float* grad_w_slots, int grads_by_slot, void* stream);

/* ---- the same exchange over NVLink peer memory instead of NCCL ------------------------------------
 * The id dispatch, the row return and the gradient push are stores into the PEERS' buffers issued by
 * the producing kernels (fused gather + dispatch); release/acquire flags in peer memory order them
 * (csrc/p2p.cu).  peer_* are arrays of `world` device pointers, one per rank, to symmetric buffers
 * (the host maps them, e.g. torch symmetric memory or cudaIpc): ids_in[world*cap] int (a ring kept
 * by the caller, reset to -1 ahead of use: 2 deep, or 4 deep when ids are dispatched one step ahead), rows_in[world*cap*K], w_in[world*cap], grad_in, gw_in
 * (same shapes), flags[5*world] int (zero-initialised).  `step` counts from 1 and must increase; pass
 * step <= 0 for (the model's device step counter - step), see b200rec_p2p_begin_step_dev: 0 = the
 * current step, -1 = the next one (an id dispatch issued one step ahead). */
/* n_dev (may be NULL): device count of valid ids, e.g. the distinct ids of the batch (dedup before the
 * exchange: every id is requested once, its gradient is pre-reduced locally and pushed once). */
int b200rec_p2p_dispatch_ids_dev(b200rec_model_t m, int64_t nnz, const int* n_dev, int world, int rank,
                                 int64_t period, int cap, int step, const int* feats,
                                 void* const* peer_ids_in, void* const* peer_flags, int* dst,
                                 int* overflow, void* stream);
/* Start of a sharded step: advances the model's DEVICE step counter (the p2p calls use it whenever
 * their `step` argument is <= 0, which is what a replayed CUDA graph needs) and fills ids_next[0..n)
 * with the -1 padding. */
int b200rec_p2p_begin_step_dev(b200rec_model_t m, int* ids_next, int64_t n, void* stream);

/* Allreduce (sum, rank order: bit-identical on every rank) of the dense gradients over NVLink peer
 * memory: inout is copied to this rank's symmetric buffer peer_bufs[rank] (n rounded up to 4 floats,
 * pad zero) and flagged (phase 3).  peer_out == NULL: one-shot, every rank reads all vectors with
 * peer loads.  Otherwise two-shot: rank r sums slice r and stores it into every peer_out[q] (phase 4),
 * then the result is copied back: 2(G-1)/G instead of (G-1) vectors over NVLink per rank.  Replaces
 * the worker -> PS push of the dense gradients (rec/model/ParRecModel.scala:247-264) for
 * data-parallel replicas. */
int b200rec_p2p_allreduce_dev(b200rec_model_t m, int64_t n, int world, int rank, int step, float* inout,
                              void* const* peer_bufs, void* const* peer_out, void* const* peer_flags,
                              const int* flags_local, void* stream);

/* Side streams: the sorts of workspace `ws` (0, 1, 2) run on a side stream of the model
 * (b200rec_segsum_sort_dev forks it from `stream`; b200rec_segsum_join_dev / _reduce_dev join it).
 * A caller may put more work there: _side_stream returns it, _side_fork_dev makes it wait for what
 * `stream` holds so far, _side_rejoin_dev moves the join point behind what the side stream holds now. */
int b200rec_model_side_stream(b200rec_model_t m, int ws, void** stream_out);
int b200rec_side_fork_dev(b200rec_model_t m, int ws, void* stream);
int b200rec_side_rejoin_dev(b200rec_model_t m, int ws);

/* Capture the device calls issued on the model's stream (and its side streams) between _begin and
 * _end into a CUDA graph; b200rec_graph_launch replays it.  Calls made while capturing are recorded,
 * not executed; every buffer must already have its final size (run the step eagerly once before). */
int b200rec_capture_begin(b200rec_model_t m, void* stream);
int b200rec_capture_end(b200rec_model_t m, int* graph_id, void* stream);
/* B200REC_ERR_STATE when a workspace of the library was reallocated after the capture (the graph would
 * replay freed addresses): run the step eagerly once and capture again. */
int b200rec_graph_launch(b200rec_model_t m, int graph_id, void* stream);
/* Count of device-buffer reallocations in this process; a captured graph (the library's own of the
 * resident step, or a caller's) is valid only while it is unchanged. */
int b200rec_alloc_epoch(int64_t* epoch);

/* slot of every non-zero from the slot of its distinct id (the ids' sort of workspace `ws` links them) */
int b200rec_p2p_compose_dst_dev(b200rec_model_t m, int ws, int64_t nnz, const int* dst_unique, int* dst,
                                void* stream);
/* spin (device side) until every rank's flag of `phase` (0 ids, 1 rows, 2 grads, 3 dense published, 4 dense reduced) has reached `step` */
int b200rec_p2p_wait_dev(b200rec_model_t m, const int* flags_local, int phase, int world, int step,
                         void* stream);
int b200rec_p2p_gather_dev(b200rec_model_t m, b200rec_table_t t, int world, int rank, int cap, int step,
                           const int* ids_in, void* const* peer_rows_in, void* const* peer_w_in,
                           void* const* peer_flags, void* stream);
int b200rec_p2p_push_grads_dev(b200rec_model_t m, int64_t nnz, const int* n_dev, int world, int rank,
                               int cap, int step, const int* dst, const float* emb_grad,
                               const float* w_grad,
                               void* const* peer_grad_in, void* const* peer_gw_in,
                               void* const* peer_flags, void* stream);
/* The two calls a rank makes between its backward and the push -- b200rec_segsum_reduce_dev (local in-order
 * pre-reduce per distinct id) and b200rec_p2p_push_grads_dev -- as ONE kernel: the per-id sums are stored
 * straight into the owners' grad_in / gw_in slots (dst_unique[u] = slot of distinct id u, from
 * b200rec_p2p_dispatch_ids_dev) and the kernel raises the phase-2 flags.  Same sums, same slots; one kernel
 * and one pass over the gradients less on the step's critical path. */
int b200rec_p2p_reduce_push_dev(b200rec_model_t m, int ws, int64_t nnz, int key_bits, const int* feats,
                                const float* emb_grad, const float* w_grad, int* unique, int* n_unique_dev,
                                int world, int rank, int cap, int step, const int* dst_unique,
                                void* const* peer_grad_in, void* const* peer_gw_in, void* const* peer_flags,
                                void* stream);
/* Plain SGD on the touched rows: E[id] -= lr * G[id], w[id] -= lr * gw[id] (rec/optim/
 * AsyncSGD.scala:10-31 applies the pushed gradient on the PS; textbook form, parity unpinned). */
int b200rec_table_apply_sgd_dev(b200rec_table_t t, int64_t n_unique_cap, const int* n_unique,
                                const int* unique, const float* emb_grad, const float* w_grad,
                                float lr, void* stream);

/* ---- optimizer step: what the PS does with a pushed gradient ------------------------
 * rec/optim/Async*.scala hand (gradient, hyper-parameters, slot offset) to Angel PSFs; the PSF
 * arithmetic is third-party: textbook forms, PARITY UNPINNED (csrc/optim.cu states them).  Optimizer
 * state ("slots", ParRecModel.scala:75,79,96,115) is allocated inside the handle on first use.
 * `step` = 1-based update count (Adam bias correction; AsyncAdam.numUpdates).
 * Touched rows only: unique[U], emb_grad[U*K], w_grad[U] as produced by the step (device pointers). */
int b200rec_table_apply_optimizer_dev(b200rec_table_t t, int optimizer, float lr, float p1, float p2,
                                      int64_t step, int64_t n_unique_cap, const int* n_unique,
                                      const int* unique, const float* emb_grad, const float* w_grad,
                                      void* stream);
/* Dense params of the handle (bias, mats) with the gradients of its last step. */
int b200rec_model_apply_optimizer_dev(b200rec_model_t m, int optimizer, float lr, float p1, float p2,
                                      int64_t step, void* stream);
/* The same two calls with the 1-based update count read from DEVICE memory (*step_dev) when the kernel
 * runs: what a replayed CUDA graph needs (AsyncAdam.numUpdates lives on the host in the reference,
 * rec/optim/AsyncAdam.scala:12,16-18).  b200rec_model_step_counter returns the model's own counter,
 * advanced by b200rec_p2p_begin_step_dev. */
int b200rec_table_apply_optimizer_stepdev_dev(b200rec_table_t t, int optimizer, float lr, float p1, float p2,
                                              const int* step_dev, int64_t n_unique_cap,
                                              const int* n_unique, const int* unique,
                                              const float* emb_grad, const float* w_grad, void* stream);
int b200rec_model_apply_optimizer_stepdev_dev(b200rec_model_t m, int optimizer, float lr, float p1, float p2,
                                              const int* step_dev, void* stream);
int b200rec_model_step_counter(b200rec_model_t m, int** step_dev);
/* Deferred status of the table's device kernels since the last call: an id outside [0, rows) seen by a
 * lookup / owner-side gather / optimizer kernel (never applied: such rows are skipped) ->
 * B200REC_ERR_INDEX.  reset != 0 clears the word.  Synchronises `stream` (NULL: the table's stream). */
int b200rec_table_status(b200rec_table_t t, int reset, void* stream);

/* ---- the reference's own BigDL modules (updateOutput / updateGradInput / accGradParameters) */
/* nn/Scatter.scala:17-36  output[index[i], :] += input[i, :]  (i ascending; bit-exact order). */
int b200rec_scatter_update_output(int device, int batch_size, int n_output, int64_t n,
                                  const float* input, const int* index, float* output);
/* nn/Scatter.scala:38-59  grad_input[i, :] = grad_output[index[i], :]. */
int b200rec_scatter_update_grad_input(int device, int batch_size, int n_output, int64_t n,
                                      const int* index, const float* grad_output, float* grad_input);
/* nn/Gather.scala:19-48   row_out[b,p,:] = input[b,rows[p],:], col_out[b,p,:] = input[b,cols[p],:]. */
int b200rec_gather_update_output(int device, int batch_size, int n_fields, int n_pairs, int dim,
                                 const float* input, const int* rows, const int* cols,
                                 float* row_out, float* col_out);
/* nn/Gather.scala:50-78   grad_input[b,rows[p],:] += g_row[b,p,:]; [b,cols[p],:] += g_col[b,p,:]
 * (p ascending, row before col; grad_input zeroed first, see SURVEY B-9). */
int b200rec_gather_update_grad_input(int device, int batch_size, int n_fields, int n_pairs, int dim,
                                     const int* rows, const int* cols, const float* g_row,
                                     const float* g_col, float* grad_input);
/* nn/DotProduct2.scala:16-26  out[b,p] = sum_k a[b,p,k]*b[b,p,k] (k ascending). */
int b200rec_dotproduct2_update_output(int device, int64_t n_rows, int dim, const float* a,
                                      const float* b, float* out);
/* nn/DotProduct2.scala:28-53  ga = b*go, gb = a*go. */
int b200rec_dotproduct2_update_grad_input(int device, int64_t n_rows, int dim, const float* a,
                                          const float* b, const float* grad_output,
                                          float* ga, float* gb);
/* rec/model/encoder/SecondOrderEncoder.scala:19-34 forward / backward.
 * out[b] = 0.5 * mean_k[(sum_f v)^2 - sum_f v^2];  grad_in = (g_b/K)(S - v). */
int b200rec_second_order_update_output(int device, int batch_size, int n_fields, int dim,
                                       const float* embedding, float* out);
int b200rec_second_order_update_grad_input(int device, int batch_size, int n_fields, int dim,
                                           const float* embedding, const float* grad_output,
                                           float* grad_input);
/* BigDL Linear as used by rec/util/LayerUtil.scala:7-24: y = x W^T + b, W:[out,in].
 * bias may be NULL.  relu != 0 fuses the following ReLU (HigherOrderEncoder.scala:40-43). */
int b200rec_linear_update_output(int device, int batch_size, int in_dim, int out_dim,
                                 const float* x, const float* w, const float* bias, int relu,
                                 float* y);
/* grad_x = gy W.   */
int b200rec_linear_update_grad_input(int device, int batch_size, int in_dim, int out_dim,
                                     const float* gy, const float* w, float* gx);
/* grad_w += scale * gy^T x ; grad_b += scale * sum_b gy  (accGradParameters accumulates). */
int b200rec_linear_acc_grad_parameters(int device, int batch_size, int in_dim, int out_dim,
                                       const float* x, const float* gy, float scale,
                                       float* grad_w, float* grad_b);

/* ---- the encoders: the dense branch of each model as the reference exposes it ------------------------
 * forward(input: Tensor[B, nFields*embeddingDim]) -> Tensor[B,1] and backward(input, gradOutput) ->
 * gradInput, with the parameter gradients copied over `mats` at the parameters' own offsets
 * (BackwardUtil.linearBackward / biasBackward, rec/util/BackwardUtil.scala:6-42):
 *   higher_order  rec/model/encoder/HigherOrderEncoder.scala:18-32   handle kind B200REC_DEEPFM
 *                 (nFields * embeddingDim = inputDim; for PNN's own tower over fcDims.head create the
 *                 handle with nFields = fcDims.head, embeddingDim = 1, fcDims = fcDims.tail)
 *   cin           rec/model/xdeepfm/CINEncoder.scala:36,60           handle kind B200REC_XDEEPFM
 *   cross         rec/model/dcn/CrossEncoder.scala:40,57             handle kind B200REC_DCN
 *   product       rec/model/pnn/ProductEncoder.scala:34,43           handle kind B200REC_PNN; output and
 *                 gradOutput are [B, fcDims.head], mats is the encoder's own prefix [W_z | W_p | c]
 * `mats` points at the encoder's first parameter (the reference's mats + start); its length is
 * b200rec_encoder_mats_len.  The BigDL method triple is kept for each:
 *   _update_output        = forward;
 *   _update_grad_input    = gradInput only (parameters untouched);
 *   _acc_grad_parameters  = grad_mats += scale * parameter gradients, same layout as mats (accumulates,
 *                           like BigDL's gradWeight);
 *   _backward             = the reference encoder's backward: gradInput out, `mats` OVERWRITTEN with the
 *                           gradients.
 * Host arrays; each call runs the branch's forward (+ backward) on the handle's stream and returns. */
int b200rec_encoder_mats_len(b200rec_model_t m, int64_t* len);
int b200rec_higher_order_update_output(b200rec_model_t m, int batch_size, const float* input, const float* mats,
                                       float* output);
int b200rec_higher_order_update_grad_input(b200rec_model_t m, int batch_size, const float* input, const float* mats,
                                           const float* grad_output, float* grad_input);
int b200rec_higher_order_acc_grad_parameters(b200rec_model_t m, int batch_size, const float* input, const float* mats,
                                             const float* grad_output, float scale, float* grad_mats);
int b200rec_higher_order_backward(b200rec_model_t m, int batch_size, const float* input, float* mats,
                                  const float* grad_output, float* grad_input);
int b200rec_cin_update_output(b200rec_model_t m, int batch_size, const float* input, const float* mats,
                              float* output);
int b200rec_cin_update_grad_input(b200rec_model_t m, int batch_size, const float* input, const float* mats,
                                  const float* grad_output, float* grad_input);
int b200rec_cin_acc_grad_parameters(b200rec_model_t m, int batch_size, const float* input, const float* mats,
                                    const float* grad_output, float scale, float* grad_mats);
int b200rec_cin_backward(b200rec_model_t m, int batch_size, const float* input, float* mats,
                         const float* grad_output, float* grad_input);
int b200rec_cross_update_output(b200rec_model_t m, int batch_size, const float* input, const float* mats,
                                float* output);
int b200rec_cross_update_grad_input(b200rec_model_t m, int batch_size, const float* input, const float* mats,
                                    const float* grad_output, float* grad_input);
int b200rec_cross_acc_grad_parameters(b200rec_model_t m, int batch_size, const float* input, const float* mats,
                                      const float* grad_output, float scale, float* grad_mats);
int b200rec_cross_backward(b200rec_model_t m, int batch_size, const float* input, float* mats,
                           const float* grad_output, float* grad_input);
int b200rec_product_update_output(b200rec_model_t m, int batch_size, const float* input, const float* mats,
                                  float* output);
int b200rec_product_update_grad_input(b200rec_model_t m, int batch_size, const float* input, const float* mats,
                                      const float* grad_output, float* grad_input);
int b200rec_product_acc_grad_parameters(b200rec_model_t m, int batch_size, const float* input, const float* mats,
                                        const float* grad_output, float scale, float* grad_mats);
int b200rec_product_backward(b200rec_model_t m, int batch_size, const float* input, float* mats,
                             const float* grad_output, float* grad_input);
/* nn/DuplicateTable.scala:13-56, the fan-out container of SecondOrderEncoder: forward hands the same
 * tensor to every branch (no data movement: the kernels read it in place; inside the models the fan-out
 * is fused), updateGradInput sums the branches' input gradients in branch order into a zeroed tensor
 * (:22-33).  grad_outputs: n_branches x len, branch-major. */
int b200rec_duplicate_table_update_grad_input(int device, int n_branches, int64_t len,
                                              const float* grad_outputs, float* grad_input);

#ifdef __cplusplus
}
#endif
#endif /* B200REC_H_ */
