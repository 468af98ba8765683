// The C ABI of libb200rec.so (include/b200rec.h).  Everything here is plumbing: argument checks,
// host <-> device staging, stream ordering, and status translation.  The arithmetic lives in the
// kernel translation units.
//
// Error convention: the reference raises IllegalArgumentException from Scala
// `require`s (nn/Scatter.scala:29-30,52-53; nn/DuplicateTable.scala:61-62); here every entry point
// returns a negative status and b200rec_last_error() carries the message.  Nothing throws.
#include <cstdlib>
#include <algorithm>
#include <cstdarg>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "kernels.h"
#include "model.h"

namespace b200rec {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
std::atomic<long long> g_alloc_epoch{0};
int g_pdl = -1;
int pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = std::getenv("B200REC_PDL");
    g_pdl = (e && e[0] >= '0' && e[0] <= '3') ? e[0] - '0' : 2;
  }
  return g_pdl;
}
thread_local int tl_pdl_hint = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

int g_default_gemm_mode = 1;  // 3xTF32 tcgen05 wherever a tensor-core kernel exists
thread_local Prof* tl_prof = nullptr;
thread_local const char* tl_tag = nullptr;
thread_local DevBuf* tl_pack = nullptr;
thread_local PrePack* tl_prepack = nullptr;

// Under the per-kernel profiler a step starts with this spin: the host gets ~0.5 ms ahead, the step's
// launches queue up behind it, and the event pairs then time kernels that run back to back instead
// of host launch gaps.
__global__ void prof_delay_kernel(unsigned long long ns) {
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do {
    __nanosleep(1000);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  } while (t - t0 < ns);
}

void Prof::begin(const char* name, cudaStream_t st) {
  if (n >= kMax) return;
  if (!recs) recs = new Rec[kMax]();
  Rec& r = recs[n];
  if (!r.a) { cudaEventCreate(&r.a); cudaEventCreate(&r.b); }
  r.tag = tl_tag; r.name = name;
  cudaEventRecord(r.a, st);
}
void Prof::end(cudaStream_t st) {
  if (n >= kMax) return;
  cudaEventRecord(recs[n].b, st);
  ++n;
}

// A usable device is an sm_100 (B200) one: the library carries sm_100a SASS only, no CPU path.
static int use_device(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    cudaGetLastError();
    set_error("no CUDA device (%s): libb200rec has no CPU fallback",
              e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
    return B200REC_ERR_CUDA;
  }
  B200_REQUIRE(device >= 0 && device < n, B200REC_ERR_ARG, "device %d out of range [0,%d)", device, n);
  static int cc_major[64];
  static bool probed[64];
  if (device < 64 && !probed[device]) {
    int major = 0;
    B200_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    cc_major[device] = major;
    probed[device] = true;
  }
  if (device < 64)
    B200_REQUIRE(cc_major[device] == 10, B200REC_ERR_CUDA,
                 "device %d has compute capability %d.x; libb200rec is built for sm_100a only",
                 device, cc_major[device]);
  B200_CUDA(cudaSetDevice(device));
  return B200REC_OK;
}

struct ScopedBuf : DevBuf {
  ~ScopedBuf() { release(); }
};

static int upload(DevBuf& b, const void* src, size_t bytes, cudaStream_t st) {
  B200_TRY(b.reserve(bytes ? bytes : 4));
  if (bytes) B200_CUDA(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, st));
  return B200REC_OK;
}
static int download(void* dst, const void* src, size_t bytes, cudaStream_t st) {
  if (bytes) B200_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
  return B200REC_OK;
}

// translate the device status word (already on the host) into a status
static int dev_status(int word, int batch_size, long long rows) {
  if (word & DEV_BAD_INDEX) {
    set_error("index should smaller than %d (nn/Scatter.scala:29-30)", batch_size);
    return B200REC_ERR_INDEX;
  }
  if (word & DEV_BAD_ID) {
    set_error("feature id outside the table's %lld rows", rows);
    return B200REC_ERR_INDEX;
  }
  if (word & DEV_PEER_TIMEOUT) {
    set_error("a peer GPU's exchange flag did not arrive within B200REC_P2P_TIMEOUT_MS: the step's results are invalid");
    return B200REC_ERR_STATE;
  }
  return B200REC_OK;
}

static int key_bits_for(long long rows) {
  int b = 1;
  while (b < 32 && (1LL << b) < rows) ++b;
  return b;
}

}  // namespace b200rec

using namespace b200rec;

#define B200_GUARD_BEGIN try {
#define B200_GUARD_END                                        \
  }                                                           \
  catch (const std::bad_alloc&) {                             \
    set_error("host allocation failed");                      \
    return B200REC_ERR_NOMEM;                                 \
  }                                                           \
  catch (...) {                                               \
    set_error("unexpected C++ exception");                    \
    return B200REC_ERR_ARG;                                   \
  }

extern "C" {

// ---- library -----------------------------------------------------------------------------------
int b200rec_abi_version(void) { return B200REC_ABI_VERSION; }
const char* b200rec_last_error(void) { return g_err; }

int b200rec_device_count(int* count) {
  B200_REQUIRE(count, B200REC_ERR_ARG, "count is NULL");
  *count = 0;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    cudaGetLastError();
    set_error("no CUDA device: libb200rec has no CPU fallback");
    return B200REC_ERR_CUDA;
  }
  int ok = 0;
  for (int d = 0; d < n; ++d) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess &&
        major == 10)
      ++ok;
  }
  *count = ok;
  if (!ok) {
    set_error("no sm_100 device among %d CUDA devices", n);
    return B200REC_ERR_CUDA;
  }
  return B200REC_OK;
}

int b200rec_launch_count(int64_t* count) {
  B200_REQUIRE(count, B200REC_ERR_ARG, "count is NULL");
  *count = (int64_t)g_launches.load();
  return B200REC_OK;
}

int b200rec_profile_begin(void) {
  static thread_local Prof prof;
  prof.n = 0;
  tl_prof = &prof;
  return B200REC_OK;
}

__global__ void prof_empty_kernel() {}

// What an event pair adds to the kernel it brackets: the median elapsed time of 64 bracketed empty
// kernels queued behind a spin (so that no host launch gap is in it).  A profile_begin/_end pass
// reports raw event times; a caller comparing them with durations (ncu gpu__time_duration, or a
// graph replay) subtracts this per launch.
int b200rec_profile_overhead_us(int device, float* us) {
  B200_GUARD_BEGIN
  B200_REQUIRE(us, B200REC_ERR_ARG, "NULL argument");
  B200_TRY(use_device(device));
  constexpr int N = 64;
  cudaStream_t st;
  B200_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  cudaEvent_t a[N], b[N];
  for (int i = 0; i < N; ++i) { cudaEventCreate(&a[i]); cudaEventCreate(&b[i]); }
  prof_delay_kernel<<<1, 1, 0, st>>>(500000ull);
  for (int i = 0; i < N; ++i) {
    cudaEventRecord(a[i], st);
    prof_empty_kernel<<<1, 32, 0, st>>>();
    cudaEventRecord(b[i], st);
  }
  cudaError_t e = cudaStreamSynchronize(st);
  std::vector<float> ms(N, 0.f);
  for (int i = 0; i < N; ++i) {
    if (e == cudaSuccess) cudaEventElapsedTime(&ms[i], a[i], b[i]);
    cudaEventDestroy(a[i]); cudaEventDestroy(b[i]);
  }
  cudaStreamDestroy(st);
  B200_CUDA(e);
  std::sort(ms.begin(), ms.end());
  *us = ms[N / 2] * 1000.f;
  return B200REC_OK;
  B200_GUARD_END
}

int b200rec_profile_end(char* buf, int64_t cap, int64_t* needed) {
  B200_GUARD_BEGIN
  Prof* p = tl_prof;
  tl_prof = nullptr;
  B200_REQUIRE(p, B200REC_ERR_STATE, "b200rec_profile_begin was not called on this thread");
  B200_CUDA(cudaDeviceSynchronize());
  struct Agg { std::string key; int count; double ms; };
  std::vector<Agg> agg;
  for (int i = 0; i < p->n; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p->recs[i].a, p->recs[i].b) != cudaSuccess) { cudaGetLastError(); continue; }
    std::string key = std::string(p->recs[i].tag ? p->recs[i].tag : "-") + "|" + p->recs[i].name;
    size_t j = 0;
    for (; j < agg.size(); ++j) if (agg[j].key == key) break;
    if (j == agg.size()) agg.push_back({key, 0, 0.0});
    agg[j].count++; agg[j].ms += ms;
  }
  std::string out;
  char line[64];
  for (auto& a : agg) {
    snprintf(line, sizeof line, "|%d|%.6f\n", a.count, a.ms);
    out += a.key + line;
  }
  if (needed) *needed = (int64_t)out.size() + 1;
  if (buf && cap > 0) {
    size_t ncopy = out.size() < (size_t)cap - 1 ? out.size() : (size_t)cap - 1;
    memcpy(buf, out.data(), ncopy);
    buf[ncopy] = 0;
  }
  return B200REC_OK;
  B200_GUARD_END
}

// ---- model ---------------------------------------------------------------------------------------
int b200rec_model_create(int kind, int n_fields, int embedding_dim, const int* fc_dims, int n_fc,
                         const int* cin_dims, int n_cin, int cross_depth, int device,
                         b200rec_model_t* out) {
  B200_GUARD_BEGIN
  B200_REQUIRE(out, B200REC_ERR_ARG, "out is NULL");
  *out = nullptr;
  B200_REQUIRE(n_fc >= 0 && n_cin >= 0, B200REC_ERR_ARG, "negative dims count");
  B200_REQUIRE(n_fc == 0 || fc_dims, B200REC_ERR_ARG, "fc_dims is NULL");
  B200_REQUIRE(n_cin == 0 || cin_dims, B200REC_ERR_ARG, "cin_dims is NULL");
  B200_TRY(use_device(device));
  Model* m = new Model();
  int s = m->init(kind, n_fields, embedding_dim, fc_dims, n_fc, cin_dims, n_cin, cross_depth, device);
  if (s != B200REC_OK) {
    m->destroy();
    delete m;
    return s;
  }
  *out = m;
  return B200REC_OK;
  B200_GUARD_END
}

int b200rec_model_destroy(b200rec_model_t m) {
  if (!m) return B200REC_OK;
  m->destroy();
  delete m;
  return B200REC_OK;
}

int b200rec_model_mats_size(b200rec_model_t m, int* pairs, int cap, int* n) {
  B200_REQUIRE(m && n, B200REC_ERR_ARG, "NULL argument");
  *n = (int)m->pairs.size();
  if (pairs)
    for (int i = 0; i < cap && i < *n; ++i) pairs[i] = m->pairs[i];
  return B200REC_OK;
}

int b200rec_model_mats_len(b200rec_model_t m, int64_t* len) {
  B200_REQUIRE(m && len, B200REC_ERR_ARG, "NULL argument");
  *len = m->mats_len;
  return B200REC_OK;
}

int b200rec_model_stream(b200rec_model_t m, void** stream) {
  B200_REQUIRE(m && stream, B200REC_ERR_ARG, "NULL argument");
  *stream = (void*)m->stream;
  return B200REC_OK;
}

int b200rec_set_default_gemm_mode(int mode) {
  B200_REQUIRE(mode >= 0 && mode <= 2, B200REC_ERR_ARG, "gemm mode must be 0 (fp32 SIMT), 1 (3xTF32 tcgen05) or 2 (1xTF32)");
  g_default_gemm_mode = mode;
  return B200REC_OK;
}

int b200rec_model_set_gemm_mode(b200rec_model_t m, int mode) {
  B200_REQUIRE(m, B200REC_ERR_ARG, "NULL model");
  B200_REQUIRE(mode >= 0 && mode <= 2, B200REC_ERR_ARG, "gemm mode must be 0 (fp32 SIMT), 1 (3xTF32 tcgen05) or 2 (1xTF32)");
  m->gemm_mode = mode;
  return B200REC_OK;
}

static int check_flat_args(Model* m, int B, long long nnz, const int* index, const float* weights,
                           const float* bias, const float* embedding, const float* mats) {
  B200_REQUIRE(m, B200REC_ERR_ARG, "NULL model");
  B200_REQUIRE(B >= 0 && nnz >= 0, B200REC_ERR_ARG, "negative batchSize / nnz");
  B200_REQUIRE(nnz < (1LL << 31), B200REC_ERR_ARG, "nnz must fit an Int (ParRecModel.scala:282)");
  B200_REQUIRE(index || nnz == 0, B200REC_ERR_ARG, "index is NULL");
  B200_REQUIRE(weights || nnz == 0, B200REC_ERR_ARG, "weights is NULL");
  B200_REQUIRE(bias, B200REC_ERR_ARG, "bias is NULL");
  if (m->kind != B200REC_LR) B200_REQUIRE(embedding || nnz == 0, B200REC_ERR_ARG, "embedding is NULL");
  if (m->mats_len > 0) B200_REQUIRE(mats, B200REC_ERR_ARG, "mats is NULL");
  if (m->kind != B200REC_LR)
    B200_REQUIRE(nnz == (long long)B * m->F, B200REC_ERR_SHAPE,
                 "nnz %lld != batchSize %d * nFields %d (Reshape to [B,F,K] would fail)", nnz, B, m->F);
  return B200REC_OK;
}

// host index == canonical COO rows (i / F)?  Then the fused per-sample kernels apply.
static bool is_canonical(const int* index, long long nnz, int B) {
  if (B <= 0 || nnz % B) return false;
  const long long f = nnz / B;
  if (f == 0) return true;
  for (long long i = 0; i < nnz; ++i)
    if (index[i] != (int)(i / f)) return false;
  return true;
}

static int run_flat_host(Model* m, int B, long long nnz, const int* index, float* weights,
                         float* bias, float* embedding, float* mats, const float* targets,
                         float* preds, float* loss) {
  B200_TRY(use_device(m->device));
  cudaStream_t st = m->stream;
  const bool train = targets != nullptr;
  const bool has_emb = m->kind != B200REC_LR;
  const size_t f = sizeof(float);
  // Scatter's require(index < batchSize) is checked on the host copy we already hold
  for (long long i = 0; i < nnz; ++i)
    B200_REQUIRE(index[i] >= 0 && index[i] < B, B200REC_ERR_INDEX,
                 "index should smaller than %d, but got %d (nn/Scatter.scala:29-30)", B, index[i]);
  const bool canonical = has_emb ? is_canonical(index, nnz, B) && (nnz == (long long)B * m->F)
                                 : is_canonical(index, nnz, B);
  B200_TRY(upload(m->wnz, weights, (size_t)nnz * f, st));
  B200_TRY(upload(m->stage_b, bias, f, st));
  if (!canonical) B200_TRY(upload(m->d_index, index, (size_t)nnz * sizeof(int), st));
  if (has_emb) B200_TRY(upload(m->X, embedding, (size_t)nnz * m->K * f, st));
  if (m->mats_len > 0) B200_TRY(upload(m->stage_a, mats, (size_t)m->mats_len * f, st));
  if (train) B200_TRY(upload(m->d_targets, targets, (size_t)B * f, st));
  B200_TRY(m->dw.reserve((size_t)(nnz ? nnz : 1) * f));
  B200_TRY(m->preds.reserve((size_t)(B ? B : 1) * f));
  RunArgs a;
  a.B = B; a.nnz = nnz;
  a.index = canonical ? nullptr : m->d_index.as<int>();
  a.w_nz = m->wnz.as<float>();
  a.emb = has_emb ? m->X.as<float>() : nullptr;
  a.bias = m->stage_b.as<float>();
  a.mats = m->mats_len > 0 ? m->stage_a.as<float>() : nullptr;
  a.targets = train ? m->d_targets.as<float>() : nullptr;
  a.preds = m->preds.as<float>();
  float* scal = m->scal.as<float>();
  if (train) {
    a.dw_out = m->dw.as<float>();
    a.dE_out = has_emb ? m->X.as<float>() : nullptr;  // in place, like the reference's Array.copy
    a.dbias_out = scal + 1;
    a.gmats_out = m->gmats.as<float>();
    a.loss_out = scal + 0;
  }
  B200_TRY(m->run(a, st));
  if (train) {
    B200_TRY(download(weights, m->dw.p, (size_t)nnz * f, st));
    if (has_emb) B200_TRY(download(embedding, m->X.p, (size_t)nnz * m->K * f, st));
    if (m->mats_len > 0) B200_TRY(download(mats, m->gmats.p, (size_t)m->mats_len * f, st));
    B200_TRY(download(m->h_scal, scal, 8 * f, st));
  } else {
    B200_TRY(download(preds, m->preds.p, (size_t)B * f, st));
    B200_TRY(download(m->h_scal, scal, 8 * f, st));
  }
  B200_CUDA(cudaStreamSynchronize(st));
  B200_TRY(dev_status(((int*)m->h_scal)[4], B, 0));
  if (train) {
    bias[0] = m->h_scal[1];
    if (loss) *loss = m->h_scal[0];
  }
  return B200REC_OK;
}

int b200rec_forward(b200rec_model_t m, int batch_size, int64_t nnz, const int* index,
                    const float* weights, const float* bias, const float* embedding,
                    const float* mats, float* preds) {
  B200_GUARD_BEGIN
  B200_TRY(check_flat_args(m, batch_size, nnz, index, weights, bias, embedding, mats));
  B200_REQUIRE(preds || batch_size == 0, B200REC_ERR_ARG, "preds is NULL");
  return run_flat_host(m, batch_size, nnz, index, (float*)weights, (float*)bias, (float*)embedding,
                       (float*)mats, nullptr, preds, nullptr);
  B200_GUARD_END
}

int b200rec_backward(b200rec_model_t m, int batch_size, int64_t nnz, const int* index,
                     float* weights, float* bias, float* embedding, float* mats,
                     const float* targets, float* loss) {
  B200_GUARD_BEGIN
  B200_TRY(check_flat_args(m, batch_size, nnz, index, weights, bias, embedding, mats));
  B200_REQUIRE(targets || batch_size == 0, B200REC_ERR_ARG, "targets is NULL");
  B200_REQUIRE(batch_size > 0, B200REC_ERR_ARG, "backward needs batchSize > 0");
  return run_flat_host(m, batch_size, nnz, index, weights, bias, embedding, mats, targets, nullptr,
                       loss);
  B200_GUARD_END
}

int b200rec_forward_dev(b200rec_model_t m, int batch_size, int64_t nnz, const int* index,
                        const float* weights, const float* bias, const float* embedding,
                        const float* mats, float* preds, void* stream) {
  B200_GUARD_BEGIN
  B200_REQUIRE(m && weights && bias && preds, B200REC_ERR_ARG, "NULL argument");
  B200_TRY(use_device(m->device));
  RunArgs a;
  a.B = batch_size; a.nnz = nnz; a.index = index; a.w_nz = weights; a.emb = embedding;
  a.bias = bias; a.mats = mats; a.preds = preds;
  return m->run(a, stream ? (cudaStream_t)stream : m->stream);
  B200_GUARD_END
}

int b200rec_backward_dev(b200rec_model_t m, int batch_size, int64_t nnz, const int* index,
                         float* weights, float* bias, float* embedding, float* mats,
                         const float* targets, float* loss_dev, void* stream) {
  B200_GUARD_BEGIN
  B200_REQUIRE(m && weights && bias && targets, B200REC_ERR_ARG, "NULL argument");
  B200_REQUIRE(batch_size > 0, B200REC_ERR_ARG, "backward needs batchSize > 0");
  B200_TRY(use_device(m->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : m->stream;
  const size_t f = sizeof(float);
  // the dense branches read `mats` and `bias` while their gradients are produced: write the
  // gradients to handle-owned buffers, then copy over the inputs (the reference does the same
  // Array.copy at the end, rec/util/BackwardUtil.scala:6-42).
  B200_TRY(m->dw.reserve((size_t)(nnz ? nnz : 1) * f));
  float* scal = m->scal.as<float>();
  RunArgs a;
  a.B = batch_size; a.nnz = nnz; a.index = index; a.w_nz = weights; a.emb = embedding;
  a.bias = bias; a.mats = mats; a.targets = targets;
  a.dw_out = m->dw.as<float>();
  a.dE_out = embedding;
  a.dbias_out = scal + 1;
  a.gmats_out = m->gmats.as<float>();
  a.loss_out = loss_dev ? loss_dev : scal + 0;
  B200_TRY(m->run(a, st));
  B200_CUDA(cudaMemcpyAsync(weights, m->dw.p, (size_t)nnz * f, cudaMemcpyDeviceToDevice, st));
  B200_CUDA(cudaMemcpyAsync(bias, scal + 1, f, cudaMemcpyDeviceToDevice, st));
  if (m->mats_len > 0 && mats)
    B200_CUDA(cudaMemcpyAsync(mats, m->gmats.p, (size_t)m->mats_len * f, cudaMemcpyDeviceToDevice, st));
  return B200REC_OK;
  B200_GUARD_END
}

int b200rec_model_sync(b200rec_model_t m) {
  B200_REQUIRE(m, B200REC_ERR_ARG, "NULL model");
  B200_TRY(use_device(m->device));
  B200_CUDA(cudaStreamSynchronize(m->stream));
  B200_CUDA(cudaStreamSynchronize(m->side));
  int word[3] = {0, 0, 0};   // [4] err of the last run, [5] sortedness, [6] sticky exchange status
  B200_CUDA(cudaMemcpy(word, m->scal.as<int>() + 4, 3 * sizeof(int), cudaMemcpyDeviceToHost));
  // a deferred error is reported once
  if (word[0] | word[2]) B200_CUDA(cudaMemset(m->scal.as<int>() + 4, 0, 3 * sizeof(int)));
  return dev_status(word[0] | word[2], m->last_B, 0);
}

// ---- table ---------------------------------------------------------------------------------------
int b200rec_table_create(int64_t rows, int dim, int device, b200rec_table_t* out) {
  B200_GUARD_BEGIN
  B200_REQUIRE(out, B200REC_ERR_ARG, "out is NULL");
  *out = nullptr;
  B200_REQUIRE(rows > 0 && rows < (1LL << 31), B200REC_ERR_ARG,
               "rows must be in [1, 2^31) (ids are Int, ParRecModel.scala:304)");
  B200_REQUIRE(dim >= 0 && dim <= 4096, B200REC_ERR_ARG, "embedding dim out of range");
  B200_TRY(use_device(device));
  Table* t = new Table();
  t->rows = rows; t->dim = dim; t->device = device;
  int s = B200REC_OK;
  do {
    if (cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking) != cudaSuccess) { s = B200REC_ERR_CUDA; set_error("cudaStreamCreate failed"); break; }
    if ((s = t->emb.reserve((size_t)rows * (dim ? dim : 1) * sizeof(float))) != B200REC_OK) break;
    if ((s = t->w.reserve((size_t)rows * sizeof(float))) != B200REC_OK) break;
    if ((s = t->err.reserve(16)) != B200REC_OK) break;
    if (cudaMemsetAsync(t->emb.p, 0, (size_t)rows * (dim ? dim : 1) * sizeof(float), t->stream) != cudaSuccess ||
        cudaMemsetAsync(t->w.p, 0, (size_t)rows * sizeof(float), t->stream) != cudaSuccess ||
        cudaMemsetAsync(t->err.p, 0, 16, t->stream) != cudaSuccess ||
        cudaStreamSynchronize(t->stream) != cudaSuccess) { s = B200REC_ERR_CUDA; set_error("table memset failed"); break; }
  } while (0);
  if (s != B200REC_OK) {
    b200rec_table_destroy(t);
    return s;
  }
  *out = t;
  return B200REC_OK;
  B200_GUARD_END
}

int b200rec_table_destroy(b200rec_table_t t) {
  if (!t) return B200REC_OK;
  cudaSetDevice(t->device);
  if (t->stream) cudaStreamSynchronize(t->stream);
  t->s1e.release(); t->s2e.release(); t->s1w.release(); t->s2w.release();
  t->emb.release(); t->w.release(); t->stage_i.release(); t->stage_e.release();
  t->stage_w.release(); t->err.release();
  if (t->stream) cudaStreamDestroy(t->stream);
  delete t;
  return B200REC_OK;
}

int b200rec_table_init_uniform(b200rec_table_t t, uint64_t seed, float lo, float hi,
                               int64_t row_offset, int64_t row_stride) {
  B200_REQUIRE(t, B200REC_ERR_ARG, "NULL table");
  B200_TRY(use_device(t->device));
  B200_TRY(table_init_uniform(t->emb.as<float>(), t->w.as<float>(), t->rows, t->dim ? t->dim : 1,
                              seed, lo, hi, row_offset, row_stride, t->stream));
  B200_CUDA(cudaStreamSynchronize(t->stream));
  return B200REC_OK;
}

int b200rec_table_write(b200rec_table_t t, int64_t row0, int64_t nrows, const float* embedding,
                        const float* weights) {
  B200_REQUIRE(t, B200REC_ERR_ARG, "NULL table");
  B200_REQUIRE(row0 >= 0 && nrows >= 0 && row0 + nrows <= t->rows, B200REC_ERR_INDEX,
               "rows [%lld, %lld) outside the table's %lld rows", (long long)row0,
               (long long)(row0 + nrows), t->rows);
  B200_TRY(use_device(t->device));
  if (embedding && t->dim)
    B200_CUDA(cudaMemcpyAsync(t->emb.as<float>() + row0 * t->dim, embedding,
                              (size_t)nrows * t->dim * sizeof(float), cudaMemcpyHostToDevice, t->stream));
  if (weights)
    B200_CUDA(cudaMemcpyAsync(t->w.as<float>() + row0, weights, (size_t)nrows * sizeof(float),
                              cudaMemcpyHostToDevice, t->stream));
  B200_CUDA(cudaStreamSynchronize(t->stream));
  return B200REC_OK;
}

int b200rec_table_read(b200rec_table_t t, int64_t row0, int64_t nrows, float* embedding,
                       float* weights) {
  B200_REQUIRE(t, B200REC_ERR_ARG, "NULL table");
  B200_REQUIRE(row0 >= 0 && nrows >= 0 && row0 + nrows <= t->rows, B200REC_ERR_INDEX,
               "rows [%lld, %lld) outside the table's %lld rows", (long long)row0,
               (long long)(row0 + nrows), t->rows);
  B200_TRY(use_device(t->device));
  if (embedding && t->dim)
    B200_CUDA(cudaMemcpyAsync(embedding, t->emb.as<float>() + row0 * t->dim,
                              (size_t)nrows * t->dim * sizeof(float), cudaMemcpyDeviceToHost, t->stream));
  if (weights)
    B200_CUDA(cudaMemcpyAsync(weights, t->w.as<float>() + row0, (size_t)nrows * sizeof(float),
                              cudaMemcpyDeviceToHost, t->stream));
  B200_CUDA(cudaStreamSynchronize(t->stream));
  return B200REC_OK;
}

int b200rec_table_ptrs(b200rec_table_t t, float** embedding, float** weights) {
  B200_REQUIRE(t, B200REC_ERR_ARG, "NULL table");
  if (embedding) *embedding = t->emb.as<float>();
  if (weights) *weights = t->w.as<float>();
  return B200REC_OK;
}

int b200rec_table_lookup_dev(b200rec_table_t t, int64_t nnz, const int* feats, float* embedding_out,
                             float* weights_out, void* stream) {
  B200_REQUIRE(t, B200REC_ERR_ARG, "NULL table");
  B200_REQUIRE(nnz >= 0 && (feats || nnz == 0), B200REC_ERR_ARG, "bad feats");
  B200_TRY(use_device(t->device));
  return lookup_rows(t->rows, t->dim ? t->dim : 4, nnz, feats, t->emb.as<float>(), t->w.as<float>(),
                     t->dim ? embedding_out : nullptr, weights_out, t->err.as<int>(),
                     stream ? (cudaStream_t)stream : t->stream);
}

int b200rec_table_lookup(b200rec_table_t t, int64_t nnz, const int* feats, float* embedding_out,
                         float* weights_out) {
  B200_GUARD_BEGIN
  B200_REQUIRE(t, B200REC_ERR_ARG, "NULL table");
  B200_REQUIRE(nnz >= 0 && (feats || nnz == 0), B200REC_ERR_ARG, "bad feats");
  B200_TRY(use_device(t->device));
  if (nnz == 0) return B200REC_OK;
  cudaStream_t st = t->stream;
  const int K = t->dim;
  B200_TRY(upload(t->stage_i, feats, (size_t)nnz * sizeof(int), st));
  if (embedding_out && K) B200_TRY(t->stage_e.reserve((size_t)nnz * K * sizeof(float)));
  if (weights_out) B200_TRY(t->stage_w.reserve((size_t)nnz * sizeof(float)));
  B200_CUDA(cudaMemsetAsync(t->err.p, 0, 4, st));
  B200_TRY(lookup_rows(t->rows, K ? K : 4, nnz, t->stage_i.as<int>(), t->emb.as<float>(),
                       t->w.as<float>(), (embedding_out && K) ? t->stage_e.as<float>() : nullptr,
                       weights_out ? t->stage_w.as<float>() : nullptr, t->err.as<int>(), st));
  if (embedding_out && K) B200_TRY(download(embedding_out, t->stage_e.p, (size_t)nnz * K * sizeof(float), st));
  if (weights_out) B200_TRY(download(weights_out, t->stage_w.p, (size_t)nnz * sizeof(float), st));
  int word = 0;
  B200_TRY(download(&word, t->err.p, sizeof(int), st));
  B200_CUDA(cudaStreamSynchronize(st));
  return dev_status(word, 0, t->rows);
  B200_GUARD_END
}

int b200rec_distinct(int device, int64_t nnz, const int* feats, int* unique_out, int64_t* n_unique) {
  return b200rec_scatter_add(device, 0, nnz, feats, nullptr, nullptr, unique_out, nullptr, nullptr,
                             n_unique);
}

int b200rec_scatter_add(int device, int dim, int64_t nnz, const int* feats,
                        const float* embedding_grad, const float* weights_grad, int* unique_out,
                        float* emb_out, float* w_out, int64_t* n_unique) {
  B200_GUARD_BEGIN
  B200_REQUIRE(nnz >= 0 && nnz < (1LL << 31), B200REC_ERR_ARG, "bad nnz");
  B200_REQUIRE(n_unique, B200REC_ERR_ARG, "n_unique is NULL");
  B200_REQUIRE(feats || nnz == 0, B200REC_ERR_ARG, "feats is NULL");
  B200_REQUIRE(dim >= 0, B200REC_ERR_ARG, "bad dim");
  B200_TRY(use_device(device));
  *n_unique = 0;
  if (nnz == 0) return B200REC_OK;
  for (int64_t i = 0; i < nnz; ++i)
    B200_REQUIRE(feats[i] >= 0, B200REC_ERR_INDEX, "negative feature id %d at %lld", feats[i], (long long)i);
  cudaStream_t st = cudaStreamPerThread;
  ScopedBuf d_feats, d_dE, d_dw, d_uniq, d_G, d_gw, d_n;
  SegSumWorkspace ws;
  struct WsGuard { SegSumWorkspace& w; ~WsGuard() { w.release(); } } guard{ws};
  const bool has_e = embedding_grad && dim > 0;
  B200_TRY(upload(d_feats, feats, (size_t)nnz * sizeof(int), st));
  if (has_e) B200_TRY(upload(d_dE, embedding_grad, (size_t)nnz * dim * sizeof(float), st));
  if (weights_grad) B200_TRY(upload(d_dw, weights_grad, (size_t)nnz * sizeof(float), st));
  B200_TRY(d_uniq.reserve((size_t)nnz * sizeof(int)));
  if (has_e) B200_TRY(d_G.reserve((size_t)nnz * dim * sizeof(float)));
  if (weights_grad) B200_TRY(d_gw.reserve((size_t)nnz * sizeof(float)));
  B200_TRY(d_n.reserve(sizeof(int)));
  SegSum a;
  a.n = nnz; a.K = dim > 0 ? dim : 4; a.key_bits = 31;
  a.feats = d_feats.as<int>();
  a.dE = has_e ? d_dE.as<float>() : nullptr;
  a.dw = weights_grad ? d_dw.as<float>() : nullptr;
  a.unique = d_uniq.as<int>();
  a.G = has_e ? d_G.as<float>() : nullptr;
  a.gw = weights_grad ? d_gw.as<float>() : nullptr;
  a.n_unique = d_n.as<int>();
  B200_TRY(segsum_sort(ws, a, st));
  if (a.dE || a.dw) B200_TRY(segsum_reduce(ws, a, st));
  int U = 0;
  B200_CUDA(cudaMemcpyAsync(&U, d_n.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  B200_CUDA(cudaStreamSynchronize(st));
  *n_unique = U;
  if (unique_out) B200_CUDA(cudaMemcpyAsync(unique_out, d_uniq.p, (size_t)U * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (emb_out && has_e) B200_CUDA(cudaMemcpyAsync(emb_out, d_G.p, (size_t)U * dim * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (w_out && weights_grad) B200_CUDA(cudaMemcpyAsync(w_out, d_gw.p, (size_t)U * sizeof(float), cudaMemcpyDeviceToHost, st));
  B200_CUDA(cudaStreamSynchronize(st));
  return B200REC_OK;
  B200_GUARD_END
}

// ---- resident step -------------------------------------------------------------------------------
int b200rec_model_set_params(b200rec_model_t m, const float* bias, const float* mats) {
  B200_REQUIRE(m && bias, B200REC_ERR_ARG, "NULL argument");
  B200_REQUIRE(mats || m->mats_len == 0, B200REC_ERR_ARG, "mats is NULL");
  B200_TRY(use_device(m->device));
  B200_CUDA(cudaMemcpyAsync(m->p_bias.p, bias, sizeof(float), cudaMemcpyHostToDevice, m->stream));
  if (m->mats_len)
    B200_CUDA(cudaMemcpyAsync(m->p_mats.p, mats, (size_t)m->mats_len * sizeof(float),
                              cudaMemcpyHostToDevice, m->stream));
  B200_CUDA(cudaStreamSynchronize(m->stream));
  m->params_set = true;
  return B200REC_OK;
}

int b200rec_model_get_params(b200rec_model_t m, float* bias, float* mats) {
  B200_REQUIRE(m, B200REC_ERR_ARG, "NULL model");
  B200_REQUIRE(m->params_set, B200REC_ERR_STATE, "params were never set");
  B200_TRY(use_device(m->device));
  if (bias) B200_CUDA(cudaMemcpyAsync(bias, m->p_bias.p, sizeof(float), cudaMemcpyDeviceToHost, m->stream));
  if (mats && m->mats_len)
    B200_CUDA(cudaMemcpyAsync(mats, m->p_mats.p, (size_t)m->mats_len * sizeof(float),
                              cudaMemcpyDeviceToHost, m->stream));
  B200_CUDA(cudaStreamSynchronize(m->stream));
  return B200REC_OK;
}

int b200rec_model_param_ptrs(b200rec_model_t m, float** bias, float** mats) {
  B200_REQUIRE(m, B200REC_ERR_ARG, "NULL model");
  if (bias) *bias = m->p_bias.as<float>();
  if (mats) *mats = m->p_mats.as<float>();
  return B200REC_OK;
}

static int check_step_args(Model* m, Table* t, int B) {
  B200_REQUIRE(m && t, B200REC_ERR_ARG, "NULL handle");
  B200_REQUIRE(m->params_set, B200REC_ERR_STATE, "b200rec_model_set_params must be called before a step");
  B200_REQUIRE(B > 0, B200REC_ERR_ARG, "batchSize must be positive");
  B200_REQUIRE(m->F > 0, B200REC_ERR_ARG, "the resident step needs nFields (one id per field per sample)");
  B200_REQUIRE(m->device == t->device, B200REC_ERR_ARG, "model on device %d, table on device %d", m->device, t->device);
  if (m->kind != B200REC_LR)
    B200_REQUIRE(t->dim == m->K, B200REC_ERR_SHAPE, "table dim %d != embeddingDim %d", t->dim, m->K);
  B200_REQUIRE((long long)B * m->F < (1LL << 31), B200REC_ERR_ARG, "batchSize * nFields must fit an Int");
  return B200REC_OK;
}

static int step_on_device(Model* m, Table* t, int B, const int* feats, const float* targets,
                          float* preds, cudaStream_t st) {
  const long long nnz = (long long)B * m->F;
  const bool train = targets != nullptr;
  const bool has_emb = m->kind != B200REC_LR;
  const size_t f = sizeof(float);
  float* scal = m->scal.as<float>();
  m->last_B = B; m->last_nnz = nnz;
  SegSum sg;
  if (train) {
    B200_TRY(m->uniq.reserve((size_t)nnz * sizeof(int)));
    if (has_emb) B200_TRY(m->G.reserve((size_t)nnz * m->K * f));
    B200_TRY(m->gwU.reserve((size_t)nnz * f));
    B200_TRY(m->dw.reserve((size_t)nnz * f));
    B200_TRY(m->seg.reserve(nnz));
    sg.n = nnz; sg.K = has_emb ? m->K : 4; sg.key_bits = key_bits_for(t->rows);
    sg.background = m->kind != B200REC_LR && m->kind != B200REC_FM;   // FM / LR: the sort IS the critical path
    sg.feats = feats;
    sg.unique = m->uniq.as<int>();
    sg.n_unique = m->scal.as<int>() + 2;
    // the sort half needs only the ids: run it beside the forward / dense math
    B200_CUDA(cudaEventRecord(m->ev_fork, st));
    B200_CUDA(cudaStreamWaitEvent(m->side, m->ev_fork, 0));
    B200_TRY(segsum_sort(m->seg, sg, m->side));
    B200_CUDA(cudaEventRecord(m->ev_join, m->side));
  }
  RunArgs a;
  a.B = B; a.nnz = nnz;
  a.feats = feats; a.table_emb = t->emb.as<float>(); a.table_w = t->w.as<float>();
  a.table_rows = t->rows;
  a.bias = m->p_bias.as<float>();
  a.mats = m->mats_len ? m->p_mats.as<float>() : nullptr;
  a.targets = targets;
  a.preds = preds;
  if (train) {
    a.dw_out = m->dw.as<float>();
    if (has_emb) {
      B200_TRY(m->X.reserve((size_t)nnz * m->K * f));
      a.dE_out = m->X.as<float>();  // per-nnz gradients overwrite the gathered rows
    }
    a.dbias_out = scal + 1;
    a.gmats_out = m->gmats.as<float>();
    a.loss_out = scal + 0;
    a.defer_sparse_bwd = m->fused_scatter;
  }
  B200_TRY(m->run(a, st));
  if (train) {
    B200_CUDA(cudaStreamWaitEvent(st, m->ev_join, 0));
    sg.G = has_emb ? m->G.as<float>() : nullptr;
    sg.gw = m->gwU.as<float>();
    if (m->deferred.valid) {
      // makeEmbeddingGrad / makeWeightsGrad with the per-nnz gradient computed inside the reduce
      sg.fused = true;
      sg.fX = m->deferred.X; sg.fS = m->deferred.S; sg.fdX = m->deferred.dX; sg.fdlogit = m->deferred.dlogit;
      sg.fF = m->F;
      if (m->keep_nnz_grads) { sg.keep_dE = m->X.as<float>(); sg.keep_dw = m->dw.as<float>(); }
    } else {
      sg.dE = has_emb ? m->X.as<float>() : nullptr;
      sg.dw = m->dw.as<float>();
    }
    B200_TRY(segsum_reduce(m->seg, sg, st));
  }
  return B200REC_OK;
}

// The resident training step as a CUDA graph: the ~45 dependent launches of one step replay with
// sub-microsecond gaps instead of a stream launch each.  Inputs are staged at fixed addresses
// (d_feats / d_targets), the first step of a configuration runs eagerly (it sizes every workspace and
// sets kernel attributes), the second is captured (both streams: the sort fork/join becomes graph
// edges), later ones replay.  Any capture failure falls back to eager launches for good.
static int step_train_graphed(Model* m, Table* t, int B, const int* feats, const float* targets,
                              cudaStream_t st) {
  const long long nnz = (long long)B * m->F;
  B200_TRY(m->d_feats.reserve((size_t)nnz * sizeof(int)));
  B200_TRY(m->d_targets.reserve((size_t)B * sizeof(float)));
  if (feats != m->d_feats.as<int>())
    B200_CUDA(cudaMemcpyAsync(m->d_feats.p, feats, (size_t)nnz * sizeof(int), cudaMemcpyDeviceToDevice, st));
  if (targets != m->d_targets.as<float>())
    B200_CUDA(cudaMemcpyAsync(m->d_targets.p, targets, (size_t)B * sizeof(float), cudaMemcpyDeviceToDevice, st));
  const int* f = m->d_feats.as<int>();
  const float* tg = m->d_targets.as<float>();
  if (tl_prof) prof_delay_kernel<<<1, 1, 0, st>>>(500000ull);
  if (!m->graph_enabled || tl_prof) return step_on_device(m, t, B, f, tg, nullptr, st);
  // The graph bakes in the addresses of the handle's workspaces.  Any call that grew one of them since
  // the capture (b200rec_predict with a larger batch, b200rec_forward / _backward, step_rows ...) freed
  // the old block: the allocation epoch moved and the graph must not be replayed.
  const long long epoch = g_alloc_epoch.load();
  const bool match = m->graph_exec && m->graph_B == B && m->graph_table == (const void*)t &&
                     m->graph_mode == m->gemm_mode && m->graph_epoch == epoch;
  if (!match) {
    if (m->graph_warm_B != B || m->graph_warm_epoch != epoch) {
      // eager warm-up: allocations and attribute calls happen here; the capture follows once a whole
      // eager step has run without moving the epoch again
      if (m->graph_exec) { cudaGraphExecDestroy(m->graph_exec); m->graph_exec = nullptr; }
      const int s0 = step_on_device(m, t, B, f, tg, nullptr, st);
      m->graph_warm_B = B;
      m->graph_warm_epoch = g_alloc_epoch.load();
      return s0;
    }
    if (m->graph_exec) { cudaGraphExecDestroy(m->graph_exec); m->graph_exec = nullptr; }
    cudaGraph_t graph = nullptr;
    const long long l0 = g_launches.load();
    bool ok = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    int s = B200REC_OK;
    if (ok) {
      s = step_on_device(m, t, B, f, tg, nullptr, st);
      ok = cudaStreamEndCapture(st, &graph) == cudaSuccess && s == B200REC_OK && graph;
    }
    if (ok) ok = cudaGraphInstantiate(&m->graph_exec, graph, 0) == cudaSuccess;
    if (graph) cudaGraphDestroy(graph);
    if (!ok) {
      cudaGetLastError();
      m->graph_enabled = false;
      m->graph_exec = nullptr;
      return step_on_device(m, t, B, f, tg, nullptr, st);
    }
    m->graph_nodes = (int)(g_launches.load() - l0);   // launches recorded, not executed, by the capture
    g_launches.fetch_sub(m->graph_nodes);
    m->graph_B = B; m->graph_table = (const void*)t; m->graph_mode = m->gemm_mode;
    m->graph_epoch = g_alloc_epoch.load();
    if (m->graph_epoch != epoch) {   // a workspace grew inside the capture: the graph is stale already
      cudaGraphExecDestroy(m->graph_exec);
      m->graph_exec = nullptr;
      m->graph_warm_epoch = -1;
      return step_on_device(m, t, B, f, tg, nullptr, st);
    }
    m->last_B = B; m->last_nnz = nnz;
  }
  B200_CUDA(cudaGraphLaunch(m->graph_exec, st));
  g_launches.fetch_add(m->graph_nodes, std::memory_order_relaxed);
  return B200REC_OK;
}

int b200rec_model_set_graph(b200rec_model_t m, int enabled) {
  B200_REQUIRE(m, B200REC_ERR_ARG, "NULL model");
  m->graph_enabled = enabled != 0;
  return B200REC_OK;
}

int b200rec_step_dev(b200rec_model_t m, b200rec_table_t t, int batch_size, const int* feats,
                     const float* targets, void* stream) {
  B200_GUARD_BEGIN
  B200_TRY(check_step_args(m, t, batch_size));
  B200_REQUIRE(feats && targets, B200REC_ERR_ARG, "NULL argument");
  B200_TRY(use_device(m->device));
  return step_train_graphed(m, t, batch_size, feats, targets, stream ? (cudaStream_t)stream : m->stream);
  B200_GUARD_END
}

int b200rec_step(b200rec_model_t m, b200rec_table_t t, int batch_size, const int* feats,
                 const float* targets, float* loss) {
  B200_GUARD_BEGIN
  B200_TRY(check_step_args(m, t, batch_size));
  B200_REQUIRE(feats && targets, B200REC_ERR_ARG, "NULL argument");
  B200_TRY(use_device(m->device));
  cudaStream_t st = m->stream;
  const long long nnz = (long long)batch_size * m->F;
  B200_TRY(upload(m->d_feats, feats, (size_t)nnz * sizeof(int), st));
  B200_TRY(upload(m->d_targets, targets, (size_t)batch_size * sizeof(float), st));
  B200_TRY(step_train_graphed(m, t, batch_size, m->d_feats.as<int>(), m->d_targets.as<float>(), st));
  B200_TRY(download(m->h_scal, m->scal.p, 8 * sizeof(float), st));
  B200_CUDA(cudaStreamSynchronize(st));
  B200_TRY(dev_status(((int*)m->h_scal)[4], batch_size, t->rows));
  if (loss) *loss = m->h_scal[0];
  return B200REC_OK;
  B200_GUARD_END
}

// ---- host-facing step with input prefetch ---------------------------------------------------------------
int b200rec_stage_batch(b200rec_model_t m, int batch_size, const int* feats, const float* targets) {
  B200_GUARD_BEGIN
  B200_REQUIRE(m && feats && targets && batch_size > 0, B200REC_ERR_ARG, "bad argument");
  B200_REQUIRE(m->stage_count < 2, B200REC_ERR_STATE, "two batches are staged already: call b200rec_step_staged");
  B200_TRY(use_device(m->device));
  const int slot = (m->stage_head + m->stage_count) & 1;
  const long long nnz = (long long)batch_size * m->F;
  // the slot's previous batch must have been consumed by the step that used it
  if (m->stage_used[slot]) B200_CUDA(cudaStreamWaitEvent(m->copy_stream, m->ev_consumed[slot], 0));
  B200_TRY(upload(m->stage_f[slot], feats, (size_t)nnz * sizeof(int), m->copy_stream));
  B200_TRY(upload(m->stage_t[slot], targets, (size_t)batch_size * sizeof(float), m->copy_stream));
  B200_CUDA(cudaEventRecord(m->ev_staged[slot], m->copy_stream));
  m->stage_B[slot] = batch_size;
  ++m->stage_count;
  return B200REC_OK;
  B200_GUARD_END
}

int b200rec_step_staged_async(b200rec_model_t m, b200rec_table_t t) {
  B200_GUARD_BEGIN
  B200_REQUIRE(m && t, B200REC_ERR_ARG, "NULL argument");
  B200_REQUIRE(m->stage_count > 0, B200REC_ERR_STATE, "no staged batch: call b200rec_stage_batch first");
  B200_REQUIRE(m->async_count < 2, B200REC_ERR_STATE, "two steps are in flight already: call b200rec_step_wait");
  const int slot = m->stage_head;
  const int batch_size = m->stage_B[slot];
  B200_TRY(check_step_args(m, t, batch_size));
  B200_TRY(use_device(m->device));
  cudaStream_t st = m->stream;
  B200_CUDA(cudaStreamWaitEvent(st, m->ev_staged[slot], 0));
  m->stage_head ^= 1;
  --m->stage_count;
  B200_TRY(step_train_graphed(m, t, batch_size, m->stage_f[slot].as<int>(), m->stage_t[slot].as<float>(), st));
  B200_CUDA(cudaEventRecord(m->ev_consumed[slot], st));
  m->stage_used[slot] = true;
  const int a = (m->async_head + m->async_count) & 1;   // result slot: 8 floats of the pinned scalar block each
  B200_TRY(download(m->h_scal + 8 * a, m->scal.p, 8 * sizeof(float), st));
  B200_CUDA(cudaEventRecord(m->ev_done[a], st));
  m->async_B[a] = batch_size;
  ++m->async_count;
  return B200REC_OK;
  B200_GUARD_END
}

int b200rec_step_wait(b200rec_model_t m, b200rec_table_t t, float* loss) {
  B200_GUARD_BEGIN
  B200_REQUIRE(m && t, B200REC_ERR_ARG, "NULL argument");
  B200_REQUIRE(m->async_count > 0, B200REC_ERR_STATE, "no step in flight");
  B200_TRY(use_device(m->device));
  const int a = m->async_head;
  B200_CUDA(cudaEventSynchronize(m->ev_done[a]));
  m->async_head ^= 1;
  --m->async_count;
  B200_TRY(dev_status(((int*)(m->h_scal + 8 * a))[4], m->async_B[a], t->rows));
  if (loss) *loss = m->h_scal[8 * a];
  return B200REC_OK;
  B200_GUARD_END
}

int b200rec_step_staged(b200rec_model_t m, b200rec_table_t t, float* loss) {
  B200_REQUIRE(m, B200REC_ERR_ARG, "NULL argument");
  B200_REQUIRE(m->async_count == 0, B200REC_ERR_STATE, "asynchronous steps are in flight: call b200rec_step_wait");
  B200_TRY(b200rec_step_staged_async(m, t));
  return b200rec_step_wait(m, t, loss);
}

int b200rec_predict(b200rec_model_t m, b200rec_table_t t, int batch_size, const int* feats,
                    float* preds) {
  B200_GUARD_BEGIN
  B200_TRY(check_step_args(m, t, batch_size));
  B200_REQUIRE(feats && preds, B200REC_ERR_ARG, "NULL argument");
  B200_TRY(use_device(m->device));
  cudaStream_t st = m->stream;
  const long long nnz = (long long)batch_size * m->F;
  B200_TRY(upload(m->d_feats, feats, (size_t)nnz * sizeof(int), st));
  B200_TRY(m->preds.reserve((size_t)batch_size * sizeof(float)));
  B200_TRY(step_on_device(m, t, batch_size, m->d_feats.as<int>(), nullptr, m->preds.as<float>(), st));
  B200_TRY(download(preds, m->preds.p, (size_t)batch_size * sizeof(float), st));
  B200_TRY(download(m->h_scal, m->scal.p, 8 * sizeof(float), st));
  B200_CUDA(cudaStreamSynchronize(st));
  return dev_status(((int*)m->h_scal)[4], batch_size, t->rows);
  B200_GUARD_END
}

int b200rec_predict_dev(b200rec_model_t m, b200rec_table_t t, int batch_size, const int* feats,
                        float* preds, void* stream) {
  B200_GUARD_BEGIN
  B200_TRY(check_step_args(m, t, batch_size));
  B200_REQUIRE(feats && preds, B200REC_ERR_ARG, "NULL argument");
  B200_TRY(use_device(m->device));
  return step_on_device(m, t, batch_size, feats, nullptr, preds,
                        stream ? (cudaStream_t)stream : m->stream);
  B200_GUARD_END
}

int b200rec_step_results(b200rec_model_t m, float* loss, int64_t* n_unique, int* unique,
                         float* emb_grad, float* w_grad, float* bias_grad, float* mats_grad) {
  B200_GUARD_BEGIN
  B200_REQUIRE(m, B200REC_ERR_ARG, "NULL model");
  B200_REQUIRE(m->last_B > 0, B200REC_ERR_STATE, "no step has run on this handle");
  B200_TRY(use_device(m->device));
  cudaStream_t st = m->stream;
  B200_TRY(download(m->h_scal, m->scal.p, 8 * sizeof(float), st));
  B200_CUDA(cudaStreamSynchronize(st));
  const int U = ((int*)m->h_scal)[2];
  if (loss) *loss = m->h_scal[0];
  if (bias_grad) *bias_grad = m->h_scal[1];
  if (n_unique) *n_unique = U;
  if (unique) B200_TRY(download(unique, m->uniq.p, (size_t)U * sizeof(int), st));
  if (emb_grad && m->kind != B200REC_LR) B200_TRY(download(emb_grad, m->G.p, (size_t)U * m->K * sizeof(float), st));
  if (w_grad) B200_TRY(download(w_grad, m->gwU.p, (size_t)U * sizeof(float), st));
  if (mats_grad && m->mats_len) B200_TRY(download(mats_grad, m->gmats.p, (size_t)m->mats_len * sizeof(float), st));
  B200_CUDA(cudaStreamSynchronize(st));
  return B200REC_OK;
  B200_GUARD_END
}

int b200rec_step_result_ptrs(b200rec_model_t m, float** loss, int** n_unique, int** unique,
                             float** emb_grad, float** w_grad, float** bias_grad,
                             float** mats_grad) {
  B200_REQUIRE(m, B200REC_ERR_ARG, "NULL model");
  if (loss) *loss = m->scal.as<float>() + 0;
  if (bias_grad) *bias_grad = m->scal.as<float>() + 1;
  if (n_unique) *n_unique = m->scal.as<int>() + 2;
  if (unique) *unique = m->uniq.as<int>();
  if (emb_grad) *emb_grad = m->G.as<float>();
  if (w_grad) *w_grad = m->gwU.as<float>();
  if (mats_grad) *mats_grad = m->gmats.as<float>();
  return B200REC_OK;
}

int b200rec_model_set_fused_scatter(b200rec_model_t m, int enabled, int keep_nnz_grads) {
  B200_REQUIRE(m, B200REC_ERR_ARG, "NULL model");
  m->fused_scatter = enabled != 0;
  m->keep_nnz_grads = keep_nnz_grads != 0;
  // the captured graph of the resident step bakes the choice in
  if (m->graph_exec) { cudaGraphExecDestroy(m->graph_exec); m->graph_exec = nullptr; }
  m->graph_warm_epoch = -1;
  return B200REC_OK;
}

int b200rec_step_nnz_grad_ptrs(b200rec_model_t m, float** emb_grad, float** w_grad) {
  B200_REQUIRE(m, B200REC_ERR_ARG, "NULL model");
  if (emb_grad) *emb_grad = m->X.as<float>();
  if (w_grad) *w_grad = m->dw.as<float>();
  return B200REC_OK;
}

int b200rec_step_gathered_dev(b200rec_model_t m, int batch_size, float* emb, float* w,
                              const float* targets, void* stream) {
  B200_GUARD_BEGIN
  B200_REQUIRE(m && w && targets, B200REC_ERR_ARG, "NULL argument");
  B200_REQUIRE(m->params_set, B200REC_ERR_STATE, "b200rec_model_set_params must be called before a step");
  B200_REQUIRE(batch_size > 0 && m->F > 0, B200REC_ERR_ARG, "batchSize and nFields must be positive");
  B200_REQUIRE(emb || m->kind == B200REC_LR, B200REC_ERR_ARG, "emb is NULL");
  B200_TRY(use_device(m->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : m->stream;
  const long long nnz = (long long)batch_size * m->F;
  float* scal = m->scal.as<float>();
  m->last_B = batch_size; m->last_nnz = nnz;
  B200_TRY(m->dw.reserve((size_t)nnz * sizeof(float)));
  RunArgs a;
  a.B = batch_size; a.nnz = nnz;
  a.w_nz = w; a.emb = emb;
  a.bias = m->p_bias.as<float>();
  a.mats = m->mats_len ? m->p_mats.as<float>() : nullptr;
  a.targets = targets;
  a.dw_out = m->dw.as<float>();
  a.dE_out = emb;
  a.dbias_out = scal + 1;
  a.gmats_out = m->gmats.as<float>();
  a.loss_out = scal + 0;
  B200_TRY(m->run(a, st));
  B200_CUDA(cudaMemcpyAsync(w, m->dw.p, (size_t)nnz * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return B200REC_OK;
  B200_GUARD_END
}

int b200rec_segsum_dev(b200rec_model_t m, int dim, int64_t nnz, int key_bits, int drop_pad,
                       const int* feats, const float* emb_grad, const float* w_grad, int* unique_out,
                       float* emb_out, float* w_out, int* n_unique_dev, void* stream) {
  B200_GUARD_BEGIN
  B200_REQUIRE(m && n_unique_dev && unique_out, B200REC_ERR_ARG, "NULL argument");
  B200_REQUIRE(nnz >= 0 && nnz < (1LL << 31), B200REC_ERR_ARG, "bad nnz");
  B200_TRY(use_device(m->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : m->stream;
  SegSum a;
  a.n = nnz; a.K = dim > 0 ? dim : 4; a.key_bits = drop_pad ? 32 : (key_bits > 0 ? key_bits : 31);
  a.drop_pad = drop_pad != 0;
  a.feats = feats; a.dE = dim > 0 ? emb_grad : nullptr; a.dw = w_grad;
  a.unique = unique_out; a.G = dim > 0 ? emb_out : nullptr; a.gw = w_out;
  a.n_unique = n_unique_dev;
  B200_TRY(segsum_sort(m->seg, a, st));
  if (a.dE || a.dw) B200_TRY(segsum_reduce(m->seg, a, st));
  return B200REC_OK;
  B200_GUARD_END
}

// The two halves of b200rec_segsum_dev for overlap: the sort depends only on the ids, so the owner
// runs it on the handle's side stream (forked from `stream`) while the dense math runs; the reduce
// joins it.  Same arguments in both calls.
static int fill_segsum(Model* m, int dim, int64_t nnz, int key_bits, int drop_pad, const int* feats,
                       const float* emb_grad, const float* w_grad, int* unique_out, float* emb_out,
                       float* w_out, int* n_unique_dev, SegSum& a) {
  B200_REQUIRE(m && n_unique_dev && unique_out, B200REC_ERR_ARG, "NULL argument");
  B200_REQUIRE(nnz >= 0 && nnz < (1LL << 31), B200REC_ERR_ARG, "bad nnz");
  // padding (-1 = all ones) still sorts last when one extra key bit is kept above the ids' bits
  a.n = nnz; a.K = dim > 0 ? dim : 4;
  a.key_bits = key_bits > 0 ? (drop_pad ? (key_bits < 31 ? key_bits + 1 : 32) : key_bits) : (drop_pad ? 32 : 31);
  a.drop_pad = drop_pad != 0;
  a.feats = feats; a.dE = dim > 0 ? emb_grad : nullptr; a.dw = w_grad;
  a.unique = unique_out; a.G = dim > 0 ? emb_out : nullptr; a.gw = w_out;
  a.n_unique = n_unique_dev;
  return B200REC_OK;
}

int b200rec_segsum_sort_dev(b200rec_model_t m, int ws, int dim, int64_t nnz, int key_bits, int drop_pad,
                            const int* feats, int* unique_out, int* n_unique_dev, void* stream) {
  B200_GUARD_BEGIN
  SegSum a;
  B200_TRY(fill_segsum(m, dim, nnz, key_bits, drop_pad, feats, nullptr, nullptr, unique_out, nullptr,
                       nullptr, n_unique_dev, a));
  B200_REQUIRE(ws >= 0 && ws <= 2, B200REC_ERR_ARG, "workspace must be 0, 1 or 2");
  B200_TRY(use_device(m->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : m->stream;
  cudaStream_t side = m->side_of(ws);
  B200_CUDA(cudaEventRecord(m->fork_of(ws), st));
  B200_CUDA(cudaStreamWaitEvent(side, m->fork_of(ws), 0));
  B200_TRY(segsum_sort(m->ws_of(ws), a, side));
  B200_CUDA(cudaEventRecord(m->join_of(ws), side));
  m->join_pending[ws] = true;
  return B200REC_OK;
  B200_GUARD_END
}

int b200rec_segsum_join_dev(b200rec_model_t m, int ws, void* stream) {
  B200_REQUIRE(m && ws >= 0 && ws <= 2, B200REC_ERR_ARG, "bad argument");
  B200_TRY(use_device(m->device));
  // joined once: a later capture must not wait on an event recorded by an earlier capture
  if (m->join_pending[ws]) B200_CUDA(cudaStreamWaitEvent(stream ? (cudaStream_t)stream : m->stream, m->join_of(ws), 0));
  m->join_pending[ws] = false;
  return B200REC_OK;
}

int b200rec_segsum_inverse_dev(b200rec_model_t m, int ws, int64_t nnz, int* inv_out, void* stream) {
  B200_REQUIRE(m && inv_out && ws >= 0 && ws <= 2, B200REC_ERR_ARG, "bad argument");
  B200_TRY(use_device(m->device));
  return segsum_inverse(m->ws_of(ws), nnz, inv_out, stream ? (cudaStream_t)stream : m->stream);
}

int b200rec_segsum_reduce_dev(b200rec_model_t m, int ws, int dim, int64_t nnz, int key_bits, int drop_pad,
                              const int* feats, const float* emb_grad, const float* w_grad,
                              int* unique_out, float* emb_out, float* w_out, int* n_unique_dev,
                              void* stream) {
  B200_GUARD_BEGIN
  SegSum a;
  B200_TRY(fill_segsum(m, dim, nnz, key_bits, drop_pad, feats, emb_grad, w_grad, unique_out, emb_out,
                       w_out, n_unique_dev, a));
  B200_REQUIRE(ws >= 0 && ws <= 2, B200REC_ERR_ARG, "workspace must be 0, 1 or 2");
  B200_TRY(use_device(m->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : m->stream;
  if (m->join_pending[ws]) B200_CUDA(cudaStreamWaitEvent(st, m->join_of(ws), 0));
  m->join_pending[ws] = false;
  if (a.dE || a.dw) B200_TRY(segsum_reduce(m->ws_of(ws), a, st));
  return B200REC_OK;
  B200_GUARD_END
}

// ---- row-sharded table ------------------------------------------------------------------
int b200rec_table_init_uniform_sharded(b200rec_table_t t, uint64_t seed, float lo, float hi, int rank,
                                       int world, int64_t period) {
  B200_REQUIRE(t, B200REC_ERR_ARG, "NULL table");
  B200_REQUIRE(world >= 1 && rank >= 0 && rank < world, B200REC_ERR_ARG, "bad rank %d / world %d", rank, world);
  B200_REQUIRE(shard_period_ok(world, period), B200REC_ERR_ARG,
               "shard period must be a positive multiple of the world size, or negative (-rows per rank: contiguous ranges)");
  B200_TRY(use_device(t->device));
  B200_TRY(table_init_uniform_sharded(t->emb.as<float>(), t->w.as<float>(), t->rows, t->dim ? t->dim : 1,
                                      seed, lo, hi, rank, world, period, t->stream));
  B200_CUDA(cudaStreamSynchronize(t->stream));
  return B200REC_OK;
}

int b200rec_table_lookup_padded_dev(b200rec_table_t t, int64_t n, const int* local_rows,
                                    float* embedding_out, float* weights_out, void* stream) {
  B200_REQUIRE(t, B200REC_ERR_ARG, "NULL table");
  B200_REQUIRE(n >= 0 && (local_rows || n == 0), B200REC_ERR_ARG, "bad ids");
  B200_TRY(use_device(t->device));
  return lookup_rows_padded(t->rows, t->dim ? t->dim : 4, n, local_rows, t->emb.as<float>(),
                            t->w.as<float>(), t->dim ? embedding_out : nullptr, weights_out,
                            t->err.as<int>(), stream ? (cudaStream_t)stream : t->stream);
}

int b200rec_shard_plan_dev(b200rec_model_t m, int64_t nnz, int world, int64_t period, int cap,
                           const int* feats, int* send_ids, int* dst, int* overflow, void* stream) {
  B200_GUARD_BEGIN
  B200_REQUIRE(m && send_ids && dst && overflow, B200REC_ERR_ARG, "NULL argument");
  B200_REQUIRE(nnz >= 0 && nnz < (1LL << 31) && cap > 0, B200REC_ERR_ARG, "bad sizes");
  B200_REQUIRE(feats || nnz == 0, B200REC_ERR_ARG, "feats is NULL");
  B200_TRY(use_device(m->device));
  return shard_plan(m->plan, nnz, world, period, cap, feats, send_ids, dst, overflow,
                    stream ? (cudaStream_t)stream : m->stream);
  B200_GUARD_END
}

int b200rec_step_rows_dev(b200rec_model_t m, int batch_size, const int* slots, const float* rows_emb,
                          const float* rows_w, int64_t n_rows, const float* targets,
                          float* grad_emb_slots, float* grad_w_slots, int grads_by_slot, void* stream) {
  B200_GUARD_BEGIN
  B200_REQUIRE(m && slots && rows_w && targets && grad_w_slots, B200REC_ERR_ARG, "NULL argument");
  B200_REQUIRE(m->params_set, B200REC_ERR_STATE, "b200rec_model_set_params must be called before a step");
  B200_REQUIRE(batch_size > 0 && m->F > 0, B200REC_ERR_ARG, "batchSize and nFields must be positive");
  B200_REQUIRE((rows_emb && grad_emb_slots) || m->kind == B200REC_LR, B200REC_ERR_ARG, "NULL row buffers");
  B200_TRY(use_device(m->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : m->stream;
  const long long nnz = (long long)batch_size * m->F;
  float* scal = m->scal.as<float>();
  m->last_B = batch_size; m->last_nnz = nnz;
  RunArgs a;
  a.B = batch_size; a.nnz = nnz;
  a.feats = slots; a.table_emb = rows_emb ? rows_emb : rows_w; a.table_w = rows_w; a.table_rows = n_rows;
  a.bias = m->p_bias.as<float>();
  a.mats = m->mats_len ? m->p_mats.as<float>() : nullptr;
  a.targets = targets;
  a.dw_out = grad_w_slots;
  a.dE_out = m->kind == B200REC_LR ? nullptr : grad_emb_slots;
  a.out_slot = grads_by_slot ? slots : nullptr;
  a.dbias_out = scal + 1;
  a.gmats_out = m->gmats.as<float>();
  a.loss_out = scal + 0;
  return m->run(a, st);
  B200_GUARD_END
}

// ---- NVLink peer-memory exchange (csrc/p2p.cu) -------------------------------------------------------
// ctr: which block-completion counter (kernels that may run concurrently need their own: 0 gather /
// push on the main stream, 2 id dispatch (prefetched on a side stream), 3 dense allreduce)
static int fill_p2p(Model* m, int world, int rank, int step, void* const* peer_flags, P2P& c, int ctr = 0) {
  B200_REQUIRE(m && peer_flags, B200REC_ERR_ARG, "NULL argument");
  B200_REQUIRE(world >= 1 && world <= P2P_MAX && rank >= 0 && rank < world, B200REC_ERR_ARG,
               "bad rank %d / world %d (peer exchange supports up to %d GPUs)", rank, world, P2P_MAX);
  if (!m->p2p_ctr.p) {
    B200_TRY(m->p2p_ctr.reserve(16));
    B200_CUDA(cudaMemset(m->p2p_ctr.p, 0, 16));
    B200_CUDA(cudaDeviceSynchronize());
  }
  c.world = world; c.rank = rank; c.step = step;
  c.block_counter = m->p2p_ctr.as<unsigned>() + ctr;
  c.step_ptr = m->p2p_ctr.as<int>() + 1;
  for (int p = 0; p < world; ++p) c.flags[p] = (int*)peer_flags[p];
  return B200REC_OK;
}

int b200rec_p2p_wait_dev(b200rec_model_t m, const int* flags_local, int phase, int world, int step,
                         void* stream) {
  B200_GUARD_BEGIN
  B200_REQUIRE(flags_local && phase >= 0 && phase < 5, B200REC_ERR_ARG, "bad argument");
  P2P c;
  void* none[P2P_MAX] = {};
  B200_TRY(fill_p2p(m, world, 0, step, none, c));
  B200_TRY(use_device(m->device));
  return p2p_wait(flags_local, phase, world, step, c.step_ptr, m->scal.as<int>() + 6,
                  stream ? (cudaStream_t)stream : m->stream);
  B200_GUARD_END
}

int b200rec_p2p_begin_step_dev(b200rec_model_t m, int* ids_next, int64_t n, void* stream) {
  B200_GUARD_BEGIN
  P2P c;
  void* none[P2P_MAX] = {};
  B200_TRY(fill_p2p(m, 1, 0, 0, none, c));
  B200_REQUIRE(ids_next || n == 0, B200REC_ERR_ARG, "NULL argument");
  B200_TRY(use_device(m->device));
  return p2p_begin_step(m->p2p_ctr.as<int>() + 1, ids_next, n, stream ? (cudaStream_t)stream : m->stream);
  B200_GUARD_END
}

int b200rec_p2p_allreduce_dev(b200rec_model_t m, int64_t n, int world, int rank, int step, float* inout,
                              void* const* peer_bufs, void* const* peer_out, void* const* peer_flags,
                              const int* flags_local, void* stream) {
  B200_GUARD_BEGIN
  P2P c;
  B200_TRY(fill_p2p(m, world, rank, step, peer_flags, c, 3));
  B200_REQUIRE(inout && peer_bufs && flags_local && n >= 0, B200REC_ERR_ARG, "bad argument");
  B200_TRY(use_device(m->device));
  PeerF bufs, outs;
  for (int p = 0; p < world; ++p) {
    bufs.p[p] = (float*)peer_bufs[p];
    outs.p[p] = peer_out ? (float*)peer_out[p] : nullptr;
  }
  return p2p_allreduce(n, inout, flags_local, c, bufs, peer_out ? &outs : nullptr, m->scal.as<int>() + 6,
                       stream ? (cudaStream_t)stream : m->stream);
  B200_GUARD_END
}

// ---- side streams of the sort workspaces, for work the caller wants off the main stream ----------------
int b200rec_model_side_stream(b200rec_model_t m, int ws, void** stream_out) {
  B200_REQUIRE(m && stream_out && ws >= 0 && ws <= 2, B200REC_ERR_ARG, "bad argument");
  *stream_out = (void*)m->side_of(ws);
  return B200REC_OK;
}

int b200rec_side_fork_dev(b200rec_model_t m, int ws, void* stream) {
  B200_REQUIRE(m && ws >= 0 && ws <= 2, B200REC_ERR_ARG, "bad argument");
  B200_TRY(use_device(m->device));
  B200_CUDA(cudaEventRecord(m->fork_of(ws), stream ? (cudaStream_t)stream : m->stream));
  B200_CUDA(cudaStreamWaitEvent(m->side_of(ws), m->fork_of(ws), 0));
  return B200REC_OK;
}

int b200rec_side_rejoin_dev(b200rec_model_t m, int ws) {
  B200_REQUIRE(m && ws >= 0 && ws <= 2, B200REC_ERR_ARG, "bad argument");
  B200_TRY(use_device(m->device));
  B200_CUDA(cudaEventRecord(m->join_of(ws), m->side_of(ws)));
  m->join_pending[ws] = true;
  return B200REC_OK;
}

// ---- user-driven CUDA graph capture of a multi-call step (the sharded step is a sequence of ABI calls) --
int b200rec_capture_begin(b200rec_model_t m, void* stream) {
  B200_REQUIRE(m, B200REC_ERR_ARG, "NULL model");
  B200_REQUIRE(!m->capturing, B200REC_ERR_STATE, "a capture is already open on this model");
  B200_TRY(use_device(m->device));
  m->capture_l0 = g_launches.load();
  B200_CUDA(cudaStreamBeginCapture(stream ? (cudaStream_t)stream : m->stream, cudaStreamCaptureModeThreadLocal));
  m->capturing = true;
  return B200REC_OK;
}

int b200rec_capture_end(b200rec_model_t m, int* graph_id, void* stream) {
  B200_REQUIRE(m && graph_id, B200REC_ERR_ARG, "NULL argument");
  B200_REQUIRE(m->capturing, B200REC_ERR_STATE, "no capture is open on this model");
  m->capturing = false;
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamEndCapture(stream ? (cudaStream_t)stream : m->stream, &graph);
  const int nodes = (int)(g_launches.load() - m->capture_l0);   // recorded, not executed
  g_launches.fetch_sub(nodes);
  cudaGraphExec_t exec = nullptr;
  if (e == cudaSuccess && graph) e = cudaGraphInstantiate(&exec, graph, 0);
  if (graph) cudaGraphDestroy(graph);
  if (e != cudaSuccess || !exec) {
    cudaGetLastError();
    set_error("graph capture failed: %s", cudaGetErrorString(e));
    return B200REC_ERR_CUDA;
  }
  m->user_graphs.push_back({exec, nodes, g_alloc_epoch.load()});
  *graph_id = (int)m->user_graphs.size() - 1;
  return B200REC_OK;
}

int b200rec_graph_launch(b200rec_model_t m, int graph_id, void* stream) {
  B200_REQUIRE(m && graph_id >= 0 && graph_id < (int)m->user_graphs.size(), B200REC_ERR_ARG, "bad graph id");
  B200_TRY(use_device(m->device));
  auto& g = m->user_graphs[graph_id];
  // a workspace of the library was freed since the capture: the graph holds a dangling pointer
  if (!g.exec || g.epoch != g_alloc_epoch.load()) {
    if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
    set_error("graph %d is stale: device buffers were reallocated after its capture; run the step eagerly "
              "once and capture again", graph_id);
    return B200REC_ERR_STATE;
  }
  B200_CUDA(cudaGraphLaunch(g.exec, stream ? (cudaStream_t)stream : m->stream));
  g_launches.fetch_add(g.nodes, std::memory_order_relaxed);
  return B200REC_OK;
}

int b200rec_alloc_epoch(int64_t* epoch) {
  B200_REQUIRE(epoch, B200REC_ERR_ARG, "epoch is NULL");
  *epoch = (int64_t)g_alloc_epoch.load();
  return B200REC_OK;
}

int b200rec_p2p_dispatch_ids_dev(b200rec_model_t m, int64_t nnz, const int* n_dev, int world, int rank,
                                 int64_t period, int cap, int step, const int* feats,
                                 void* const* peer_ids_in, void* const* peer_flags, int* dst,
                                 int* overflow, void* stream) {
  B200_GUARD_BEGIN
  P2P c;
  B200_TRY(fill_p2p(m, world, rank, step, peer_flags, c, 2));
  B200_REQUIRE(peer_ids_in && dst && overflow && (feats || nnz == 0), B200REC_ERR_ARG, "NULL argument");
  B200_TRY(use_device(m->device));
  PeerI ids;
  for (int p = 0; p < world; ++p) ids.p[p] = (int*)peer_ids_in[p];
  return p2p_plan(m->plan, nnz, n_dev, period, cap, feats, dst, overflow, c, ids,
                  stream ? (cudaStream_t)stream : m->stream);
  B200_GUARD_END
}

int b200rec_p2p_gather_dev(b200rec_model_t m, b200rec_table_t t, int world, int rank, int cap, int step,
                           const int* ids_in, void* const* peer_rows_in, void* const* peer_w_in,
                           void* const* peer_flags, void* stream) {
  B200_GUARD_BEGIN
  P2P c;
  B200_TRY(fill_p2p(m, world, rank, step, peer_flags, c));
  B200_REQUIRE(t && ids_in && peer_rows_in && peer_w_in, B200REC_ERR_ARG, "NULL argument");
  B200_TRY(use_device(m->device));
  PeerF rows, w;
  for (int p = 0; p < world; ++p) { rows.p[p] = (float*)peer_rows_in[p]; w.p[p] = (float*)peer_w_in[p]; }
  return p2p_gather(t->rows, t->dim ? t->dim : 4, cap, ids_in, t->dim ? t->emb.as<float>() : nullptr,
                    t->w.as<float>(), c, rows, w, t->err.as<int>(), stream ? (cudaStream_t)stream : m->stream);
  B200_GUARD_END
}

int b200rec_p2p_compose_dst_dev(b200rec_model_t m, int ws, int64_t nnz, const int* dst_unique, int* dst,
                                void* stream) {
  B200_REQUIRE(m && dst_unique && dst && ws >= 0 && ws <= 2, B200REC_ERR_ARG, "bad argument");
  B200_TRY(use_device(m->device));
  return p2p_compose(m->ws_of(ws), nnz, dst_unique, dst, stream ? (cudaStream_t)stream : m->stream);
}

int b200rec_p2p_push_grads_dev(b200rec_model_t m, int64_t nnz, const int* n_dev, int world, int rank,
                               int cap, int step, const int* dst, const float* emb_grad,
                               const float* w_grad,
                               void* const* peer_grad_in, void* const* peer_gw_in,
                               void* const* peer_flags, void* stream) {
  B200_GUARD_BEGIN
  P2P c;
  B200_TRY(fill_p2p(m, world, rank, step, peer_flags, c));
  B200_REQUIRE(dst && w_grad && peer_grad_in && peer_gw_in, B200REC_ERR_ARG, "NULL argument");
  B200_TRY(use_device(m->device));
  PeerF g, gw;
  for (int p = 0; p < world; ++p) { g.p[p] = (float*)peer_grad_in[p]; gw.p[p] = (float*)peer_gw_in[p]; }
  return p2p_push_grads(nnz, n_dev, m->kind == B200REC_LR ? 4 : m->K, cap, dst, m->kind == B200REC_LR ? nullptr : emb_grad,
                        w_grad, c, g, gw, stream ? (cudaStream_t)stream : m->stream);
  B200_GUARD_END
}

// local in-order pre-reduce per distinct id FUSED with the gradient push: the sums are stored straight into
// the owners' grad_in / gw_in slots and the kernel raises the phase-2 flags (csrc/segsum.cu, SegPush)
int b200rec_p2p_reduce_push_dev(b200rec_model_t m, int ws, int64_t nnz, int key_bits, const int* feats,
                                const float* emb_grad, const float* w_grad, int* unique, int* n_unique_dev,
                                int world, int rank, int cap, int step, const int* dst_unique,
                                void* const* peer_grad_in, void* const* peer_gw_in, void* const* peer_flags,
                                void* stream) {
  B200_GUARD_BEGIN
  SegPush sp;
  B200_TRY(fill_p2p(m, world, rank, step, peer_flags, sp.c));
  B200_REQUIRE(dst_unique && w_grad && peer_grad_in && peer_gw_in && cap > 0, B200REC_ERR_ARG, "bad argument");
  B200_REQUIRE(ws >= 0 && ws <= 2, B200REC_ERR_ARG, "workspace must be 0, 1 or 2");
  const int K = m->kind == B200REC_LR ? 4 : m->K;
  B200_REQUIRE(K == 4 || K == 8 || K == 16 || K == 32 || K == 64, B200REC_ERR_ARG,
               "p2p exchange supports embeddingDim in {4,8,16,32,64}, got %d", K);
  SegSum a;
  B200_TRY(fill_segsum(m, K, nnz, key_bits, 0, feats, m->kind == B200REC_LR ? nullptr : emb_grad, w_grad, unique,
                       nullptr, nullptr, n_unique_dev, a));
  B200_TRY(use_device(m->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : m->stream;
  for (int p = 0; p < world; ++p) { sp.grad_in.p[p] = (float*)peer_grad_in[p]; sp.gw_in.p[p] = (float*)peer_gw_in[p]; }
  sp.dst = dst_unique;
  sp.cap = cap;
  a.push = &sp;
  if (m->join_pending[ws]) B200_CUDA(cudaStreamWaitEvent(st, m->join_of(ws), 0));
  m->join_pending[ws] = false;
  ProfTag tag("p2p_push_grads");
  return segsum_reduce(m->ws_of(ws), a, st);
  B200_GUARD_END
}

int b200rec_table_apply_sgd_dev(b200rec_table_t t, int64_t n_unique_cap, const int* n_unique,
                                const int* unique, const float* emb_grad, const float* w_grad,
                                float lr, void* stream) {
  B200_REQUIRE(t && n_unique && unique, B200REC_ERR_ARG, "NULL argument");
  B200_TRY(use_device(t->device));
  return apply_sgd(t->dim ? t->dim : 1, t->rows, n_unique_cap, n_unique, unique, t->dim ? emb_grad : nullptr,
                   w_grad, lr, t->emb.as<float>(), t->w.as<float>(), t->err.as<int>(),
                   stream ? (cudaStream_t)stream : t->stream);
}

// ---- optimizer step -----------------------------------------------------------------------------------
static int zeroed(DevBuf& b, size_t bytes, cudaStream_t st) {
  if (b.p && b.cap >= bytes) return B200REC_OK;
  B200_TRY(b.reserve(bytes));
  B200_CUDA(cudaMemsetAsync(b.p, 0, b.cap, st));
  return B200REC_OK;
}

static int table_apply_optimizer(b200rec_table_t t, int optimizer, float lr, float p1, float p2,
                                 int64_t step, const int* step_dev, int64_t n_unique_cap,
                                 const int* n_unique, const int* unique, const float* emb_grad,
                                 const float* w_grad, void* stream) {
  B200_GUARD_BEGIN
  B200_REQUIRE(t && n_unique && unique, B200REC_ERR_ARG, "NULL argument");
  B200_REQUIRE(optimizer >= B200REC_OPT_SGD && optimizer <= B200REC_OPT_ADAM, B200REC_ERR_ARG,
               "unknown optimizer %d (OptimUtils.scala:6-11 raises MatchError)", optimizer);
  B200_TRY(use_device(t->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : t->stream;
  const size_t ne = (size_t)t->rows * (t->dim ? t->dim : 1) * sizeof(float), nw = (size_t)t->rows * sizeof(float);
  if (optimizer != B200REC_OPT_SGD) {
    B200_TRY(zeroed(t->s1e, ne, st));
    B200_TRY(zeroed(t->s1w, nw, st));
  }
  if (optimizer == B200REC_OPT_ADAM) {
    B200_TRY(zeroed(t->s2e, ne, st));
    B200_TRY(zeroed(t->s2w, nw, st));
  }
  return opt_rows(optimizer, t->dim ? t->dim : 1, t->rows, n_unique_cap, n_unique, unique,
                  t->dim ? emb_grad : nullptr, w_grad, lr, p1, p2, step, step_dev, t->emb.as<float>(),
                  t->w.as<float>(), t->s1e.as<float>(), t->s2e.as<float>(), t->s1w.as<float>(),
                  t->s2w.as<float>(), t->err.as<int>(), st);
  B200_GUARD_END
}

int b200rec_table_apply_optimizer_dev(b200rec_table_t t, int optimizer, float lr, float p1, float p2,
                                      int64_t step, int64_t n_unique_cap, const int* n_unique,
                                      const int* unique, const float* emb_grad, const float* w_grad,
                                      void* stream) {
  return table_apply_optimizer(t, optimizer, lr, p1, p2, step, nullptr, n_unique_cap, n_unique, unique,
                               emb_grad, w_grad, stream);
}

int b200rec_table_apply_optimizer_stepdev_dev(b200rec_table_t t, int optimizer, float lr, float p1, float p2,
                                              const int* step_dev, int64_t n_unique_cap,
                                              const int* n_unique, const int* unique,
                                              const float* emb_grad, const float* w_grad, void* stream) {
  B200_REQUIRE(step_dev, B200REC_ERR_ARG, "step_dev is NULL");
  return table_apply_optimizer(t, optimizer, lr, p1, p2, 0, step_dev, n_unique_cap, n_unique, unique,
                               emb_grad, w_grad, stream);
}

// The device status word of the table (ids outside the table seen by a lookup / gather / optimizer
// kernel since the last call); reset != 0 clears it.  Synchronises `stream` (NULL: the table's).
int b200rec_table_status(b200rec_table_t t, int reset, void* stream) {
  B200_REQUIRE(t, B200REC_ERR_ARG, "NULL table");
  B200_TRY(use_device(t->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : t->stream;
  int word = 0;
  B200_CUDA(cudaMemcpyAsync(&word, t->err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  if (reset) B200_CUDA(cudaMemsetAsync(t->err.p, 0, sizeof(int), st));
  B200_CUDA(cudaStreamSynchronize(st));
  return dev_status(word, 0, t->rows);
}

// The model's device step counter (advanced by b200rec_p2p_begin_step_dev): lets an optimizer call
// inside a replayed CUDA graph read the update count from the device.
int b200rec_model_step_counter(b200rec_model_t m, int** step_dev) {
  B200_REQUIRE(m && step_dev, B200REC_ERR_ARG, "NULL argument");
  B200_TRY(use_device(m->device));
  if (!m->p2p_ctr.p) {
    B200_TRY(m->p2p_ctr.reserve(16));
    B200_CUDA(cudaMemset(m->p2p_ctr.p, 0, 16));
    B200_CUDA(cudaDeviceSynchronize());
  }
  *step_dev = m->p2p_ctr.as<int>() + 1;
  return B200REC_OK;
}

static int model_apply_optimizer(b200rec_model_t m, int optimizer, float lr, float p1, float p2,
                                 int64_t step, const int* step_dev, void* stream) {
  B200_GUARD_BEGIN
  B200_REQUIRE(m, B200REC_ERR_ARG, "NULL model");
  B200_REQUIRE(m->params_set && m->last_B > 0, B200REC_ERR_STATE, "no step has produced gradients yet");
  B200_REQUIRE(optimizer >= B200REC_OPT_SGD && optimizer <= B200REC_OPT_ADAM, B200REC_ERR_ARG,
               "unknown optimizer %d", optimizer);
  B200_TRY(use_device(m->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : m->stream;
  const size_t n = (size_t)(m->mats_len + 1) * sizeof(float);
  if (optimizer != B200REC_OPT_SGD) B200_TRY(zeroed(m->s1m, n, st));
  if (optimizer == B200REC_OPT_ADAM) B200_TRY(zeroed(m->s2m, n, st));
  if (m->mats_len)
    B200_TRY(opt_dense(optimizer, m->mats_len, m->gmats.as<float>(), lr, p1, p2, step, step_dev,
                       m->p_mats.as<float>(), m->s1m.as<float>(), m->s2m.as<float>(), st));
  // bias: gradient at gmats[mats_len], slots at s*m[mats_len]
  return opt_dense(optimizer, 1, m->gmats.as<float>() + m->mats_len, lr, p1, p2, step, step_dev,
                   m->p_bias.as<float>(), m->s1m.p ? m->s1m.as<float>() + m->mats_len : nullptr,
                   m->s2m.p ? m->s2m.as<float>() + m->mats_len : nullptr, st);
  B200_GUARD_END
}

int b200rec_model_apply_optimizer_dev(b200rec_model_t m, int optimizer, float lr, float p1, float p2,
                                      int64_t step, void* stream) {
  return model_apply_optimizer(m, optimizer, lr, p1, p2, step, nullptr, stream);
}

int b200rec_model_apply_optimizer_stepdev_dev(b200rec_model_t m, int optimizer, float lr, float p1, float p2,
                                              const int* step_dev, void* stream) {
  B200_REQUIRE(step_dev, B200REC_ERR_ARG, "step_dev is NULL");
  return model_apply_optimizer(m, optimizer, lr, p1, p2, 0, step_dev, stream);
}

// ---- the reference's own BigDL modules -------------------------------------------------------------
// Host arrays in/out; one call = upload, kernel(s), download on the calling thread's stream.
struct OpCtx {
  cudaStream_t st = cudaStreamPerThread;
  ScopedBuf err, pack;
  PackScope pack_scope{&pack};
  int init() {
    B200_TRY(err.reserve(16));
    B200_CUDA(cudaMemsetAsync(err.p, 0, 16, st));
    return B200REC_OK;
  }
  int finish(int batch_size) {
    int word[2] = {0, 0};
    B200_CUDA(cudaMemcpyAsync(word, err.p, 8, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    return dev_status(word[0], batch_size, 0);
  }
};

// ---- encoders: HigherOrderEncoder / CINEncoder / CrossEncoder / ProductEncoder .forward / .backward ------
// The reference's encoders take the [B, F*K] embedding tensor and return the [B,1] branch output
// ([B, fcDims.head] for ProductEncoder); backward returns gradInput and copies the parameter gradients
// over `mats` at the parameters' offsets (BackwardUtil.linearBackward).  The handle's kind selects the
// encoder: DEEPFM -> HigherOrderEncoder, XDEEPFM -> CINEncoder, DCN -> CrossEncoder, PNN -> ProductEncoder.
static long long encoder_mats_len(const Model* m) {
  if (m->kind == B200REC_PNN) {
    const long long P = (long long)m->F * (m->F - 1) / 2, O = m->fc[0];
    return (long long)m->D * O + P * O + 1;      // [W_z][W_p][c]  (ProductEncoder.scala:72-108)
  }
  return m->mats_len;
}

// what: 0 forward, 1 gradInput, 2 accGradParameters, 3 backward (gradInput + grads over mats)
static int encoder_call(b200rec_model_t m, int want_kind, const char* name, int what, int B,
                        const float* input, const float* mats_in, float* mats_out, const float* grad_output,
                        float scale, float* output, float* grad_input, float* grad_mats) {
  B200_GUARD_BEGIN
  B200_REQUIRE(m, B200REC_ERR_ARG, "NULL model");
  B200_REQUIRE(m->kind == want_kind, B200REC_ERR_ARG, "%s needs a handle created with kind %d (this one is kind %d)",
               name, want_kind, m->kind);
  B200_REQUIRE(B > 0, B200REC_ERR_ARG, "batchSize must be positive");
  B200_REQUIRE(input && mats_in, B200REC_ERR_ARG, "NULL input / mats");
  if (what == 0) B200_REQUIRE(output, B200REC_ERR_ARG, "NULL output");
  if (what >= 1) B200_REQUIRE(grad_output, B200REC_ERR_ARG, "NULL gradOutput");
  if (what == 1 || what == 3) B200_REQUIRE(grad_input, B200REC_ERR_ARG, "NULL gradInput");
  if (what == 2) B200_REQUIRE(grad_mats, B200REC_ERR_ARG, "NULL grad_mats");
  B200_TRY(use_device(m->device));
  cudaStream_t st = m->stream;
  const size_t f = sizeof(float);
  const long long n_in = (long long)B * m->D, n_mats = encoder_mats_len(m);
  const long long n_out = m->kind == B200REC_PNN ? (long long)B * m->fc[0] : B;
  B200_TRY(upload(m->X, input, (size_t)n_in * f, st));
  B200_TRY(m->stage_a.reserve((size_t)(m->mats_len > 0 ? m->mats_len : 1) * f));
  B200_CUDA(cudaMemcpyAsync(m->stage_a.p, mats_in, (size_t)n_mats * f, cudaMemcpyHostToDevice, st));
  B200_TRY(m->enc_o.reserve((size_t)n_out * f));
  RunArgs a;
  a.B = B; a.nnz = (long long)B * m->F;
  a.emb = m->X.as<float>();
  a.mats = m->stage_a.as<float>();
  a.encoder_only = true;
  a.enc_out = m->enc_o.as<float>();
  a.gmats_out = m->gmats.as<float>();
  if (what >= 1) {
    B200_TRY(upload(m->enc_go, grad_output, (size_t)n_out * f, st));
    B200_TRY(m->enc_dx.reserve((size_t)n_in * f));
    a.enc_grad = m->enc_go.as<float>();
    a.enc_dx = m->enc_dx.as<float>();
  }
  B200_TRY(m->run(a, st));
  if (what == 0) B200_TRY(download(output, m->enc_o.p, (size_t)n_out * f, st));
  if (what == 1 || what == 3) B200_TRY(download(grad_input, m->enc_dx.p, (size_t)n_in * f, st));
  std::vector<float> g;
  if (what == 2) {
    g.resize((size_t)n_mats);
    B200_TRY(download(g.data(), m->gmats.p, (size_t)n_mats * f, st));
  }
  if (what == 3) B200_TRY(download(mats_out, m->gmats.p, (size_t)n_mats * f, st));
  B200_TRY(download(m->h_scal, m->scal.p, 8 * f, st));
  B200_CUDA(cudaStreamSynchronize(st));
  B200_TRY(dev_status(((int*)m->h_scal)[4], B, 0));
  if (what == 2)   // accGradParameters ACCUMULATES (BigDL: gradWeight += scale * ...)
    for (long long i = 0; i < n_mats; ++i) grad_mats[i] += scale * g[(size_t)i];
  return B200REC_OK;
  B200_GUARD_END
}

int b200rec_encoder_mats_len(b200rec_model_t m, int64_t* len) {
  B200_REQUIRE(m && len, B200REC_ERR_ARG, "NULL argument");
  *len = encoder_mats_len(m);
  return B200REC_OK;
}

#define B200_ENCODER_ABI(NAME, KIND)                                                                             \
  int b200rec_##NAME##_update_output(b200rec_model_t m, int batch_size, const float* input, const float* mats,   \
                                     float* output) {                                                            \
    return encoder_call(m, KIND, #NAME, 0, batch_size, input, mats, nullptr, nullptr, 0.f, output, nullptr,      \
                        nullptr);                                                                                \
  }                                                                                                              \
  int b200rec_##NAME##_update_grad_input(b200rec_model_t m, int batch_size, const float* input,                  \
                                         const float* mats, const float* grad_output, float* grad_input) {       \
    return encoder_call(m, KIND, #NAME, 1, batch_size, input, mats, nullptr, grad_output, 0.f, nullptr,          \
                        grad_input, nullptr);                                                                    \
  }                                                                                                              \
  int b200rec_##NAME##_acc_grad_parameters(b200rec_model_t m, int batch_size, const float* input,                \
                                           const float* mats, const float* grad_output, float scale,             \
                                           float* grad_mats) {                                                   \
    return encoder_call(m, KIND, #NAME, 2, batch_size, input, mats, nullptr, grad_output, scale, nullptr,        \
                        nullptr, grad_mats);                                                                     \
  }                                                                                                              \
  int b200rec_##NAME##_backward(b200rec_model_t m, int batch_size, const float* input, float* mats,              \
                                const float* grad_output, float* grad_input) {                                   \
    return encoder_call(m, KIND, #NAME, 3, batch_size, input, mats, mats, grad_output, 0.f, nullptr, grad_input, \
                        nullptr);                                                                                \
  }
B200_ENCODER_ABI(higher_order, B200REC_DEEPFM)
B200_ENCODER_ABI(cin, B200REC_XDEEPFM)
B200_ENCODER_ABI(cross, B200REC_DCN)
B200_ENCODER_ABI(product, B200REC_PNN)
#undef B200_ENCODER_ABI

// nn/DuplicateTable.scala:13-56: the fan-out container.  Forward hands the SAME input to every branch
// (nothing to compute: the model kernels read the tensor in place); its backward is the sum of the
// branches' input gradients, added in branch order into a zeroed tensor (:22-33; the reference forgets to
// re-zero a reused gradInput, SURVEY B-9 -- zeroed here).  grad_outputs: n_branches x len, branch-major.
int b200rec_duplicate_table_update_grad_input(int device, int n_branches, int64_t len,
                                              const float* grad_outputs, float* grad_input) {
  B200_GUARD_BEGIN
  B200_REQUIRE(n_branches >= 0 && len >= 0, B200REC_ERR_ARG, "bad sizes");
  B200_REQUIRE((grad_outputs && grad_input) || len == 0 || n_branches == 0, B200REC_ERR_ARG, "NULL argument");
  B200_TRY(use_device(device));
  if (len == 0) return B200REC_OK;
  OpCtx c;
  B200_TRY(c.init());
  ScopedBuf d_g, d_o;
  B200_TRY(upload(d_g, grad_outputs, (size_t)n_branches * len * sizeof(float), c.st));
  B200_TRY(d_o.reserve((size_t)len * sizeof(float)));
  B200_CUDA(cudaMemsetAsync(d_o.p, 0, (size_t)len * sizeof(float), c.st));
  for (int i = 0; i < n_branches; ++i) B200_TRY(axpy(len, d_g.as<float>() + (size_t)i * len, d_o.as<float>(), c.st));
  B200_TRY(download(grad_input, d_o.p, (size_t)len * sizeof(float), c.st));
  return c.finish(0);
  B200_GUARD_END
}

int b200rec_scatter_update_output(int device, int batch_size, int n_output, int64_t n,
                                  const float* input, const int* index, float* output) {
  B200_GUARD_BEGIN
  B200_REQUIRE(batch_size >= 0 && n_output > 0 && n >= 0, B200REC_ERR_ARG, "bad sizes");
  B200_REQUIRE((input && index) || n == 0, B200REC_ERR_ARG, "NULL input");
  B200_REQUIRE(output || batch_size == 0, B200REC_ERR_ARG, "NULL output");
  B200_TRY(use_device(device));
  OpCtx c;
  B200_TRY(c.init());
  ScopedBuf d_in, d_idx, d_out;
  B200_TRY(upload(d_in, input, (size_t)n * n_output * sizeof(float), c.st));
  B200_TRY(upload(d_idx, index, (size_t)n * sizeof(int), c.st));
  B200_TRY(d_out.reserve((size_t)batch_size * n_output * sizeof(float) + 4));
  B200_TRY(scatter_fwd(batch_size, n_output, n, d_in.as<float>(), d_idx.as<int>(), d_out.as<float>(),
                       c.err.as<int>(), c.st));
  B200_TRY(download(output, d_out.p, (size_t)batch_size * n_output * sizeof(float), c.st));
  return c.finish(batch_size);
  B200_GUARD_END
}

int b200rec_scatter_update_grad_input(int device, int batch_size, int n_output, int64_t n,
                                      const int* index, const float* grad_output,
                                      float* grad_input) {
  B200_GUARD_BEGIN
  B200_REQUIRE(batch_size >= 0 && n_output > 0 && n >= 0, B200REC_ERR_ARG, "bad sizes");
  B200_REQUIRE((index && grad_input) || n == 0, B200REC_ERR_ARG, "NULL argument");
  B200_REQUIRE(grad_output || batch_size == 0, B200REC_ERR_ARG, "NULL grad_output");
  B200_TRY(use_device(device));
  OpCtx c;
  B200_TRY(c.init());
  ScopedBuf d_idx, d_go, d_gi;
  B200_TRY(upload(d_idx, index, (size_t)n * sizeof(int), c.st));
  B200_TRY(upload(d_go, grad_output, (size_t)batch_size * n_output * sizeof(float), c.st));
  B200_TRY(d_gi.reserve((size_t)n * n_output * sizeof(float) + 4));
  B200_TRY(scatter_bwd(batch_size, n_output, n, d_idx.as<int>(), d_go.as<float>(), d_gi.as<float>(),
                       c.err.as<int>(), c.st));
  B200_TRY(download(grad_input, d_gi.p, (size_t)n * n_output * sizeof(float), c.st));
  return c.finish(batch_size);
  B200_GUARD_END
}

static int check_pairs(const int* rows, const int* cols, int P, int F) {
  for (int p = 0; p < P; ++p)
    B200_REQUIRE(rows[p] >= 0 && rows[p] < F && cols[p] >= 0 && cols[p] < F, B200REC_ERR_INDEX,
                 "pair %d = (%d,%d) outside [0,%d)", p, rows[p], cols[p], F);
  return B200REC_OK;
}

int b200rec_gather_update_output(int device, int batch_size, int n_fields, int n_pairs, int dim,
                                 const float* input, const int* rows, const int* cols,
                                 float* row_out, float* col_out) {
  B200_GUARD_BEGIN
  B200_REQUIRE(batch_size >= 0 && n_fields > 0 && n_pairs >= 0 && dim > 0, B200REC_ERR_ARG, "bad sizes");
  B200_REQUIRE(input && rows && cols && row_out && col_out, B200REC_ERR_ARG, "NULL argument");
  B200_TRY(check_pairs(rows, cols, n_pairs, n_fields));
  B200_TRY(use_device(device));
  OpCtx c;
  B200_TRY(c.init());
  const size_t n_in = (size_t)batch_size * n_fields * dim, n_out = (size_t)batch_size * n_pairs * dim;
  if (n_out == 0) return B200REC_OK;
  ScopedBuf d_in, d_r, d_c, d_ro, d_co;
  B200_TRY(upload(d_in, input, n_in * sizeof(float), c.st));
  B200_TRY(upload(d_r, rows, (size_t)n_pairs * sizeof(int), c.st));
  B200_TRY(upload(d_c, cols, (size_t)n_pairs * sizeof(int), c.st));
  B200_TRY(d_ro.reserve(n_out * sizeof(float)));
  B200_TRY(d_co.reserve(n_out * sizeof(float)));
  B200_TRY(pair_gather_fwd(batch_size, n_fields, n_pairs, dim, d_in.as<float>(), d_r.as<int>(),
                           d_c.as<int>(), d_ro.as<float>(), d_co.as<float>(), c.st));
  B200_TRY(download(row_out, d_ro.p, n_out * sizeof(float), c.st));
  B200_TRY(download(col_out, d_co.p, n_out * sizeof(float), c.st));
  return c.finish(batch_size);
  B200_GUARD_END
}

int b200rec_gather_update_grad_input(int device, int batch_size, int n_fields, int n_pairs, int dim,
                                     const int* rows, const int* cols, const float* g_row,
                                     const float* g_col, float* grad_input) {
  B200_GUARD_BEGIN
  B200_REQUIRE(batch_size >= 0 && n_fields > 0 && n_pairs >= 0 && dim > 0, B200REC_ERR_ARG, "bad sizes");
  B200_REQUIRE(rows && cols && g_row && g_col && grad_input, B200REC_ERR_ARG, "NULL argument");
  B200_TRY(check_pairs(rows, cols, n_pairs, n_fields));
  B200_TRY(use_device(device));
  OpCtx c;
  B200_TRY(c.init());
  const size_t n_in = (size_t)batch_size * n_fields * dim, n_out = (size_t)batch_size * n_pairs * dim;
  if (n_in == 0) return B200REC_OK;
  ScopedBuf d_r, d_c, d_gr, d_gc, d_gi;
  B200_TRY(upload(d_r, rows, (size_t)n_pairs * sizeof(int), c.st));
  B200_TRY(upload(d_c, cols, (size_t)n_pairs * sizeof(int), c.st));
  B200_TRY(upload(d_gr, g_row, n_out * sizeof(float), c.st));
  B200_TRY(upload(d_gc, g_col, n_out * sizeof(float), c.st));
  B200_TRY(d_gi.reserve(n_in * sizeof(float)));
  B200_TRY(pair_gather_bwd(batch_size, n_fields, n_pairs, dim, d_r.as<int>(), d_c.as<int>(),
                           d_gr.as<float>(), d_gc.as<float>(), d_gi.as<float>(), c.st));
  B200_TRY(download(grad_input, d_gi.p, n_in * sizeof(float), c.st));
  return c.finish(batch_size);
  B200_GUARD_END
}

int b200rec_dotproduct2_update_output(int device, int64_t n_rows, int dim, const float* a,
                                      const float* b, float* out) {
  B200_GUARD_BEGIN
  B200_REQUIRE(n_rows >= 0 && dim > 0, B200REC_ERR_ARG, "bad sizes");
  B200_REQUIRE((a && b && out) || n_rows == 0, B200REC_ERR_ARG, "NULL argument");
  B200_TRY(use_device(device));
  if (n_rows == 0) return B200REC_OK;
  OpCtx c;
  B200_TRY(c.init());
  ScopedBuf d_a, d_b, d_o;
  B200_TRY(upload(d_a, a, (size_t)n_rows * dim * sizeof(float), c.st));
  B200_TRY(upload(d_b, b, (size_t)n_rows * dim * sizeof(float), c.st));
  B200_TRY(d_o.reserve((size_t)n_rows * sizeof(float)));
  B200_TRY(dot2_fwd(n_rows, dim, d_a.as<float>(), d_b.as<float>(), d_o.as<float>(), c.st));
  B200_TRY(download(out, d_o.p, (size_t)n_rows * sizeof(float), c.st));
  return c.finish(0);
  B200_GUARD_END
}

int b200rec_dotproduct2_update_grad_input(int device, int64_t n_rows, int dim, const float* a,
                                          const float* b, const float* grad_output, float* ga,
                                          float* gb) {
  B200_GUARD_BEGIN
  B200_REQUIRE(n_rows >= 0 && dim > 0, B200REC_ERR_ARG, "bad sizes");
  B200_REQUIRE((a && b && grad_output && ga && gb) || n_rows == 0, B200REC_ERR_ARG, "NULL argument");
  B200_TRY(use_device(device));
  if (n_rows == 0) return B200REC_OK;
  OpCtx c;
  B200_TRY(c.init());
  const size_t nb = (size_t)n_rows * dim * sizeof(float);
  ScopedBuf d_a, d_b, d_g, d_ga, d_gb;
  B200_TRY(upload(d_a, a, nb, c.st));
  B200_TRY(upload(d_b, b, nb, c.st));
  B200_TRY(upload(d_g, grad_output, (size_t)n_rows * sizeof(float), c.st));
  B200_TRY(d_ga.reserve(nb));
  B200_TRY(d_gb.reserve(nb));
  B200_TRY(dot2_bwd(n_rows, dim, d_a.as<float>(), d_b.as<float>(), d_g.as<float>(), d_ga.as<float>(),
                    d_gb.as<float>(), c.st));
  B200_TRY(download(ga, d_ga.p, nb, c.st));
  B200_TRY(download(gb, d_gb.p, nb, c.st));
  return c.finish(0);
  B200_GUARD_END
}

int b200rec_second_order_update_output(int device, int batch_size, int n_fields, int dim,
                                       const float* embedding, float* out) {
  B200_GUARD_BEGIN
  B200_REQUIRE(batch_size >= 0 && n_fields > 0 && dim > 0, B200REC_ERR_ARG, "bad sizes");
  B200_REQUIRE((embedding && out) || batch_size == 0, B200REC_ERR_ARG, "NULL argument");
  B200_TRY(use_device(device));
  if (batch_size == 0) return B200REC_OK;
  OpCtx c;
  B200_TRY(c.init());
  ScopedBuf d_e, d_o;
  B200_TRY(upload(d_e, embedding, (size_t)batch_size * n_fields * dim * sizeof(float), c.st));
  B200_TRY(d_o.reserve((size_t)batch_size * sizeof(float)));
  SparseFwd s;
  s.B = batch_size; s.F = n_fields; s.K = dim; s.emb_in = d_e.as<float>(); s.second = d_o.as<float>();
  B200_TRY(sparse_fwd(s, c.st));
  B200_TRY(download(out, d_o.p, (size_t)batch_size * sizeof(float), c.st));
  return c.finish(batch_size);
  B200_GUARD_END
}

int b200rec_second_order_update_grad_input(int device, int batch_size, int n_fields, int dim,
                                           const float* embedding, const float* grad_output,
                                           float* grad_input) {
  B200_GUARD_BEGIN
  B200_REQUIRE(batch_size >= 0 && n_fields > 0 && dim > 0, B200REC_ERR_ARG, "bad sizes");
  B200_REQUIRE((embedding && grad_output && grad_input) || batch_size == 0, B200REC_ERR_ARG, "NULL argument");
  B200_TRY(use_device(device));
  if (batch_size == 0) return B200REC_OK;
  OpCtx c;
  B200_TRY(c.init());
  const size_t nb = (size_t)batch_size * n_fields * dim * sizeof(float);
  ScopedBuf d_e, d_g, d_S, d_o;
  B200_TRY(upload(d_e, embedding, nb, c.st));
  B200_TRY(upload(d_g, grad_output, (size_t)batch_size * sizeof(float), c.st));
  B200_TRY(d_S.reserve((size_t)batch_size * dim * sizeof(float)));
  B200_TRY(d_o.reserve(nb));
  SparseFwd s;
  s.B = batch_size; s.F = n_fields; s.K = dim; s.emb_in = d_e.as<float>(); s.S = d_S.as<float>();
  B200_TRY(sparse_fwd(s, c.st));
  SparseBwd b;
  b.B = batch_size; b.F = n_fields; b.K = dim; b.X = d_e.as<float>(); b.S = d_S.as<float>();
  b.dlogit = d_g.as<float>(); b.dE = d_o.as<float>();
  B200_TRY(sparse_bwd(b, c.st));
  B200_TRY(download(grad_input, d_o.p, nb, c.st));
  return c.finish(batch_size);
  B200_GUARD_END
}

int b200rec_linear_update_output(int device, int batch_size, int in_dim, int out_dim, const float* x,
                                 const float* w, const float* bias, int relu, float* y) {
  B200_GUARD_BEGIN
  B200_REQUIRE(batch_size >= 0 && in_dim > 0 && out_dim > 0, B200REC_ERR_ARG, "bad sizes");
  B200_REQUIRE(w && ((x && y) || batch_size == 0), B200REC_ERR_ARG, "NULL argument");
  B200_TRY(use_device(device));
  if (batch_size == 0) return B200REC_OK;
  OpCtx c;
  B200_TRY(c.init());
  ScopedBuf d_x, d_w, d_b, d_y;
  B200_TRY(upload(d_x, x, (size_t)batch_size * in_dim * sizeof(float), c.st));
  B200_TRY(upload(d_w, w, (size_t)out_dim * in_dim * sizeof(float), c.st));
  if (bias) B200_TRY(upload(d_b, bias, (size_t)out_dim * sizeof(float), c.st));
  B200_TRY(d_y.reserve((size_t)batch_size * out_dim * sizeof(float)));
  B200_TRY(linear_fwd(batch_size, out_dim, in_dim, d_x.as<float>(), d_w.as<float>(),
                      bias ? d_b.as<float>() : nullptr, relu != 0, d_y.as<float>(), c.st, g_default_gemm_mode));
  B200_TRY(download(y, d_y.p, (size_t)batch_size * out_dim * sizeof(float), c.st));
  return c.finish(batch_size);
  B200_GUARD_END
}

int b200rec_linear_update_grad_input(int device, int batch_size, int in_dim, int out_dim,
                                     const float* gy, const float* w, float* gx) {
  B200_GUARD_BEGIN
  B200_REQUIRE(batch_size >= 0 && in_dim > 0 && out_dim > 0, B200REC_ERR_ARG, "bad sizes");
  B200_REQUIRE(w && ((gy && gx) || batch_size == 0), B200REC_ERR_ARG, "NULL argument");
  B200_TRY(use_device(device));
  if (batch_size == 0) return B200REC_OK;
  OpCtx c;
  B200_TRY(c.init());
  ScopedBuf d_g, d_w, d_x;
  B200_TRY(upload(d_g, gy, (size_t)batch_size * out_dim * sizeof(float), c.st));
  B200_TRY(upload(d_w, w, (size_t)out_dim * in_dim * sizeof(float), c.st));
  B200_TRY(d_x.reserve((size_t)batch_size * in_dim * sizeof(float)));
  B200_TRY(linear_bwd_input(batch_size, out_dim, in_dim, d_g.as<float>(), d_w.as<float>(), nullptr,
                            d_x.as<float>(), false, c.st, g_default_gemm_mode));
  B200_TRY(download(gx, d_x.p, (size_t)batch_size * in_dim * sizeof(float), c.st));
  return c.finish(batch_size);
  B200_GUARD_END
}

int b200rec_linear_acc_grad_parameters(int device, int batch_size, int in_dim, int out_dim,
                                       const float* x, const float* gy, float scale, float* grad_w,
                                       float* grad_b) {
  B200_GUARD_BEGIN
  B200_REQUIRE(batch_size >= 0 && in_dim > 0 && out_dim > 0, B200REC_ERR_ARG, "bad sizes");
  B200_REQUIRE(grad_w && ((x && gy) || batch_size == 0), B200REC_ERR_ARG, "NULL argument");
  B200_TRY(use_device(device));
  if (batch_size == 0) return B200REC_OK;
  OpCtx c;
  B200_TRY(c.init());
  ScopedBuf d_x, d_g, d_gw, d_gb, scratch;
  B200_TRY(upload(d_x, x, (size_t)batch_size * in_dim * sizeof(float), c.st));
  B200_TRY(upload(d_g, gy, (size_t)batch_size * out_dim * sizeof(float), c.st));
  B200_TRY(upload(d_gw, grad_w, (size_t)out_dim * in_dim * sizeof(float), c.st));
  if (grad_b) B200_TRY(upload(d_gb, grad_b, (size_t)out_dim * sizeof(float), c.st));
  B200_TRY(linear_bwd_params(batch_size, out_dim, in_dim, d_x.as<float>(), d_g.as<float>(), scale,
                             true, d_gw.as<float>(), grad_b ? d_gb.as<float>() : nullptr, scratch,
                             c.st, g_default_gemm_mode));
  B200_TRY(download(grad_w, d_gw.p, (size_t)out_dim * in_dim * sizeof(float), c.st));
  if (grad_b) B200_TRY(download(grad_b, d_gb.p, (size_t)out_dim * sizeof(float), c.st));
  return c.finish(batch_size);
  B200_GUARD_END
}

}  // extern "C"
