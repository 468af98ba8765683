// CPU BASELINE of the reference's gather / scatter-add (TEST + BENCH INFRASTRUCTURE ONLY; the product
// never links or loads this file).  BASELINE.md section 3: "hash-map gather N*K lookups and hash-map
// scatter-add N*K addTo's, with a C++ open-addressing map standing in for fastutil".
//
// Reference (paths relative to /root/reference/src/main/scala/io/yaochi/recommendation):
//   model/ParRecModel.scala:279-284  makeWeights       buf(i) = weight.get(feats(i))
//   model/ParRecModel.scala:300-306  makeEmbeddings    buf(i*K+j) = embeddings(j).get(feats(i))
//   model/ParRecModel.scala:293-298  makeWeightsGrad   grad.addTo(feats(i), buf(i))
//   model/ParRecModel.scala:316-328  makeEmbeddingGrad grads(j).addTo(feats(i), buf(i*K+j)), i ascending
//   model/ParRecModel.scala:337-345  distinctIntIndices IntOpenHashSet.add(cols(i))
// The pulled vectors are K hash-backed sparse vectors (Angel IntFloatVector over an
// Int2FloatOpenHashMap, one per embedding dimension), the gradient maps are fastutil 8.2.2
// Int2FloatOpenHashMap (pom.xml:20-25; third-party, not vendored).  The map below restates fastutil's
// published algorithm: power-of-two table sized arraySize(expected, 0.75), key slot =
// mix(key) & mask with mix(x) = (h = x * 0x9E3779B9) ^ (h >>> 16), linear probing, key 0 kept apart
// (containsNullKey), rehash at 3/4 fill.
//
// `threads` > 1 runs the K per-dimension maps in parallel (they are independent objects in the
// reference): the "generous, all host cores" variant.  threads == 1 is the reference's own loop order.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include <omp.h>

namespace {

inline uint32_t mix(uint32_t x) {
  const uint32_t h = x * 0x9E3779B9u;
  return h ^ (h >> 16);
}

inline uint64_t next_pow2(uint64_t x) {
  if (x <= 2) return 2;
  --x;
  x |= x >> 1; x |= x >> 2; x |= x >> 4; x |= x >> 8; x |= x >> 16; x |= x >> 32;
  return x + 1;
}

struct Int2FloatMap {
  std::vector<int32_t> key;
  std::vector<float> val;
  uint64_t mask = 0, n = 0, size = 0, max_fill = 0;
  bool has_null = false;   // fastutil keeps key 0 in the extra last slot
  float null_val = 0.f;

  explicit Int2FloatMap(uint64_t expected = 16) { init(expected); }
  void init(uint64_t expected) {
    n = next_pow2((uint64_t)((double)expected / 0.75 + 0.999999));
    mask = n - 1;
    max_fill = std::min<uint64_t>((uint64_t)((double)n * 0.75 + 0.999999), n - 1);
    key.assign(n + 1, 0);
    val.assign(n + 1, 0.f);
    size = 0;
    has_null = false;
  }
  void rehash(uint64_t new_n) {
    std::vector<int32_t> ok;
    std::vector<float> ov;
    ok.swap(key);
    ov.swap(val);
    const uint64_t old_n = n;
    n = new_n;
    mask = n - 1;
    max_fill = std::min<uint64_t>((uint64_t)((double)n * 0.75 + 0.999999), n - 1);
    key.assign(n + 1, 0);
    val.assign(n + 1, 0.f);
    for (uint64_t i = 0; i < old_n; ++i) {
      if (ok[i] == 0) continue;
      uint64_t pos = mix((uint32_t)ok[i]) & mask;
      while (key[pos] != 0) pos = (pos + 1) & mask;
      key[pos] = ok[i];
      val[pos] = ov[i];
    }
  }
  inline float get(int32_t k) const {
    if (k == 0) return has_null ? null_val : 0.f;
    uint64_t pos = mix((uint32_t)k) & mask;
    for (;;) {
      const int32_t c = key[pos];
      if (c == 0) return 0.f;
      if (c == k) return val[pos];
      pos = (pos + 1) & mask;
    }
  }
  inline void add_to(int32_t k, float incr) {
    if (k == 0) {
      if (has_null) { null_val += incr; return; }
      has_null = true;
      null_val = incr;   // defRetValue (0) + incr
      return;
    }
    uint64_t pos = mix((uint32_t)k) & mask;
    for (;;) {
      const int32_t c = key[pos];
      if (c == 0) break;
      if (c == k) { val[pos] += incr; return; }
      pos = (pos + 1) & mask;
    }
    key[pos] = k;
    val[pos] = incr;
    if (size++ >= max_fill) rehash(n * 2);
  }
  inline void put(int32_t k, float v) {
    if (k == 0) { has_null = true; null_val = v; return; }
    uint64_t pos = mix((uint32_t)k) & mask;
    for (;;) {
      const int32_t c = key[pos];
      if (c == 0) break;
      if (c == k) { val[pos] = v; return; }
      pos = (pos + 1) & mask;
    }
    key[pos] = k;
    val[pos] = v;
    if (size++ >= max_fill) rehash(n * 2);
  }
  uint64_t count() const { return size + (has_null ? 1 : 0); }
};

struct Pulled {   // what pullEmbeddings / pullWeights hand to make*: K + 1 hash-backed sparse vectors
  int K = 0;
  std::vector<Int2FloatMap> emb;
  Int2FloatMap w;
};

}  // namespace

extern "C" {

// ParRecModel.pullEmbeddings :174-177 (outside the timed path): rows[U*K] / w[U] of the distinct ids
// become K sparse vectors + 1.  Returns an opaque handle.
void* cb_pull(int K, int64_t U, const int32_t* ids, const float* rows, const float* w, int threads) {
  Pulled* p = new Pulled();
  p->K = K;
  p->emb.resize(K);
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(static)
  for (int j = 0; j < K; ++j) {
    p->emb[j].init((uint64_t)U);
    for (int64_t u = 0; u < U; ++u) p->emb[j].put(ids[u], rows[u * K + j]);
  }
  p->w.init((uint64_t)U);
  if (w)
    for (int64_t u = 0; u < U; ++u) p->w.put(ids[u], w[u]);
  return p;
}
void cb_pull_free(void* h) { delete static_cast<Pulled*>(h); }

// makeEmbeddings :300-306.  threads == 1: the reference's loop nest (i outer, j inner).
void cb_make_embeddings(void* h, int64_t N, const int32_t* feats, float* buf, int threads) {
  const Pulled* p = static_cast<Pulled*>(h);
  const int K = p->K;
  if (threads <= 1) {
    for (int64_t i = 0; i < N; ++i)
      for (int j = 0; j < K; ++j) buf[i * K + j] = p->emb[j].get(feats[i]);
    return;
  }
#pragma omp parallel for num_threads(threads) schedule(static)
  for (int64_t i = 0; i < N; ++i)
    for (int j = 0; j < K; ++j) buf[i * K + j] = p->emb[j].get(feats[i]);
}

// makeWeights :279-284
void cb_make_weights(void* h, int64_t N, const int32_t* feats, float* buf) {
  const Pulled* p = static_cast<Pulled*>(h);
  for (int64_t i = 0; i < N; ++i) buf[i] = p->w.get(feats[i]);
}

// distinctIntIndices :337-345 -> sorted here (hash order is unspecified); returns U
int64_t cb_distinct(int64_t N, const int32_t* feats, int32_t* out) {
  Int2FloatMap set(16);   // IntOpenHashSet() default capacity
  for (int64_t i = 0; i < N; ++i) set.add_to(feats[i], 0.f);
  int64_t u = 0;
  if (set.has_null) out[u++] = 0;
  for (uint64_t s = 0; s < set.n; ++s)
    if (set.key[s] != 0) out[u++] = set.key[s];
  std::sort(out, out + u);
  return u;
}

// makeEmbeddingGrad :316-328 + makeWeightsGrad :293-298: K (+1) maps sized `expected` (= the pulled
// vectors' size()), addTo in nnz order.  Results are emitted for the SORTED distinct ids `ids[U]` so that
// they can be compared with the GPU path: out_emb[U*K], out_w[U].  The emission is outside the
// reference's work and is not timed by the callers (cb_scatter_add_timed_ns reports the addTo part).
static thread_local double g_last_addto_s = 0.0;
double cb_last_addto_seconds(void) { return g_last_addto_s; }

void cb_scatter_add(int K, int64_t N, const int32_t* feats, const float* emb_grad, const float* w_grad,
                    int64_t expected, int64_t U, const int32_t* ids, float* out_emb, float* out_w,
                    int threads) {
  std::vector<Int2FloatMap> grads((size_t)K);
  Int2FloatMap gw(16);
  const double t0 = omp_get_wtime();
  if (threads <= 1) {
    for (int j = 0; j < K; ++j) grads[j].init((uint64_t)expected);
    for (int64_t i = 0; i < N; ++i)
      for (int j = 0; j < K; ++j) grads[j].add_to(feats[i], emb_grad[i * K + j]);
  } else {
#pragma omp parallel for num_threads(threads) schedule(static)
    for (int j = 0; j < K; ++j) {
      grads[j].init((uint64_t)expected);
      for (int64_t i = 0; i < N; ++i) grads[j].add_to(feats[i], emb_grad[i * K + j]);
    }
  }
  if (w_grad) {
    gw.init((uint64_t)expected);
    for (int64_t i = 0; i < N; ++i) gw.add_to(feats[i], w_grad[i]);
  }
  g_last_addto_s = omp_get_wtime() - t0;
  if (out_emb)
    for (int64_t u = 0; u < U; ++u)
      for (int j = 0; j < K; ++j) out_emb[u * K + j] = grads[j].get(ids[u]);
  if (out_w && w_grad)
    for (int64_t u = 0; u < U; ++u) out_w[u] = gw.get(ids[u]);
}

int cb_max_threads(void) { return omp_get_max_threads(); }

}  // extern "C"
