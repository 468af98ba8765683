// Micro-benchmark: what can a random row gather reach on this GPU?  The ceiling the gather / FM / scatter
// kernels are measured against should be the one the ACCESS PATTERN allows, next to the streaming-copy peak.
//   rows of 64 B (K = 16 fp32) picked by uniform or power-law ids from a table much larger than L2, read by
//   4 lanes x 128 bit each, with U independent row loads in flight per lane; the rows are summed and one
//   float4 per warp-row-slot is written (negligible) or the rows are copied out ([N,16] coalesced, "copy").
//   variants: separate 4-byte weight gather (the product's layout), or the weight in the row's own 128-B line.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 scripts/ubench/gather_rate.cu -o scripts/ubench/gather_rate
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

template <int U, bool COPY, int WMODE>   // WMODE 0: no weight, 1: separate wtable[id], 2: weight at row + 16 floats (128-B stride rows)
__global__ void __launch_bounds__(256) gather_kernel(long long n, const int* __restrict__ ids, const float* __restrict__ table,
                                                     const float* __restrict__ wtable, int stride, float* __restrict__ out,
                                                     float* __restrict__ wout) {
  const int lane = threadIdx.x & 31, sub = lane & 3, slot = lane >> 2;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float wacc = 0.f;
  for (long long base = warp * 8 * U; base < n; base += nwarps * 8 * U) {
    long long id[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = base + u * 8 + slot;
      id[u] = i < n ? ids[i] : -1;
    }
    float4 v[U];
    float w[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      w[u] = 0.f;
      if (id[u] >= 0) {
        v[u] = __ldg(reinterpret_cast<const float4*>(table + id[u] * stride + sub * 4));
        if (WMODE == 1 && sub == 0) w[u] = __ldg(wtable + id[u]);
        if (WMODE == 2 && sub == 0) w[u] = __ldg(table + id[u] * stride + 16);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = base + u * 8 + slot;
      if (COPY && i < n) {
        *reinterpret_cast<float4*>(out + i * 16 + sub * 4) = v[u];
        if (WMODE && sub == 0) wout[i] = w[u];
      }
      acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
      wacc += w[u];
    }
  }
  if (!COPY) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    out[t] = acc.x + acc.y + acc.z + acc.w + wacc;
  }
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventElapsedTime(&ms, a, b); return ms; }

template <int U, bool COPY, int WMODE>
static void run(const char* name, long long n, const int* ids_sets, int n_sets, const float* table, const float* wtable,
                int stride, float* out, float* wout, int blocks) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f, sum = 0.f;
  for (int it = 0; it < n_sets; ++it) {   // a different id set every launch: nothing of the last launch helps
    const int* ids = ids_sets + (size_t)it * n;
    cudaEventRecord(e0);
    gather_kernel<U, COPY, WMODE><<<blocks, 256>>>(n, ids, table, wtable, stride, out, wout);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    const float ms = time_ms(e0, e1);
    if (it >= 2) { best = ms < best ? ms : best; sum += ms; }
  }
  const double avg = sum / (n_sets - 2);
  const double bytes_rows = (double)n * 64, bytes_all = (double)n * (64 + 4 + (WMODE ? 4 : 0) + (COPY ? 64 + (WMODE ? 4 : 0) : 0));
  printf("%-44s n=%8lld blocks=%5d  avg %7.2f us  best %7.2f us   rows-only %6.0f GB/s   all-bytes %6.0f GB/s (avg)\n", name, n,
         blocks, avg * 1e3, best * 1e3, bytes_rows / (avg * 1e-3) / 1e9, bytes_all / (avg * 1e-3) / 1e9);
}

int main(int argc, char** argv) {
  const long long rows = 39LL << 18;          // 10 223 616 rows, as the benchmark's table
  const int n_sets = 12;
  const long long n_small = 8192LL * 39, n_big = 4LL << 20;
  float *table, *table128, *wtable, *out, *wout;
  cudaMalloc(&table, rows * 64);
  cudaMalloc(&table128, rows * 128);
  cudaMalloc(&wtable, rows * 4);
  cudaMalloc(&out, n_big * 64);
  cudaMalloc(&wout, n_big * 4);
  cudaMemset(table, 0, rows * 64); cudaMemset(table128, 0, rows * 128); cudaMemset(wtable, 0, rows * 4);
  for (int dist = 0; dist < 2; ++dist) {
    // dist 0: uniform ids; dist 1: Criteo-shaped: field f of 39 owns rows [f * 2^18, (f+1) * 2^18), power-law inside
    std::vector<int> h((size_t)n_big * n_sets);
    unsigned long long s = 88172645463325252ULL;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    for (size_t i = 0; i < h.size(); ++i) {
      if (dist == 0) h[i] = (int)(rnd() % (unsigned long long)rows);
      else {
        const int f = (int)(i % 39);
        const double u = (double)(rnd() >> 11) / 9007199254740992.0;
        const long long r = (long long)std::floor(std::pow((double)(1 << 18), u)) - 1;   // log-uniform: heavy head
        h[i] = (int)((long long)f * (1 << 18) + (r < 0 ? 0 : r));
      }
    }
    int* ids;
    cudaMalloc(&ids, h.size() * 4);
    cudaMemcpy(ids, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    printf("== ids: %s\n", dist == 0 ? "uniform over the table" : "39 fields, log-uniform (power-law) inside each field");
    for (long long n : {n_small, n_big}) {
      const int blocks_full = (int)((n / 8 + 7) / 8);   // one warp-slot row per 4 lanes, U = 1 equivalent
      const int cap = 148 * 8;
      const int b1 = blocks_full < cap ? blocks_full : cap;
      run<1, false, 0>("sum  U=1  rows only", n, ids, n_sets, table, wtable, 16, out, wout, blocks_full);
      run<5, false, 0>("sum  U=5  rows only", n, ids, n_sets, table, wtable, 16, out, wout, (blocks_full + 4) / 5);
      run<8, false, 0>("sum  U=8  rows only (capped grid)", n, ids, n_sets, table, wtable, 16, out, wout, b1);
      run<5, false, 1>("sum  U=5  rows + separate weights", n, ids, n_sets, table, wtable, 16, out, wout, (blocks_full + 4) / 5);
      run<5, false, 2>("sum  U=5  rows + weight in the 128-B line", n, ids, n_sets, table128, wtable, 32, out, wout, (blocks_full + 4) / 5);
      run<5, true, 0>("copy U=5  rows only", n, ids, n_sets, table, wtable, 16, out, wout, (blocks_full + 4) / 5);
      run<5, true, 1>("copy U=5  rows + separate weights", n, ids, n_sets, table, wtable, 16, out, wout, (blocks_full + 4) / 5);
      run<5, true, 2>("copy U=5  rows + weight in the 128-B line", n, ids, n_sets, table128, wtable, 32, out, wout, (blocks_full + 4) / 5);
    }
    cudaFree(ids);
  }
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error\n"); return 1; }
  return 0;
}
