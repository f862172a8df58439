// L2 -> SM bandwidth on B200 for the access patterns of the gather kernels: what a row-gather formulation can reach.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/l2_gather_bench scripts/micro/l2_gather_bench.cu
//   run  : scripts/micro/l2_gather_bench > gpurun_out/l2_gather.json
// Patterns (all reads, fp32 sums kept so nothing is optimised away; times by cudaEvent over 20 launches after 3 warm-ups):
//   stream  : every thread reads consecutive float4s of a buffer (coalesced 512 B per warp instruction)
//   gather R: rows of R bytes at random row indices (index array read sequentially, 4 B per row); R / 16 lanes share a
//             row, so a warp instruction touches 512 / R rows -- the pattern of conv_mma(q)_kernel (R = 64), of the
//             vector loads of conv_dw_mma_kernel (R = 64 / 128) and of conv_tc_kernel (R = 64 pieces of wider rows)
// Buffer sizes: 32 MB (L2 resident on either die), 96 MB (fits the 126 MB L2 as a whole), 1024 MB (HBM).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int kThreads = 256, kUnroll = 8;

__global__ void __launch_bounds__(kThreads) stream_kernel(const float4* __restrict__ buf, int64_t n4, float* __restrict__ out) {
  float acc = 0.f;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += stride * kUnroll) {
    float4 v[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t j = i + u * stride;
      v[u] = j < n4 ? __ldg(buf + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
  }
  if (acc == 123.456f) out[0] = acc;
}

// LPR lanes per row (row bytes = 16 * LPR); a warp instruction reads 32 / LPR rows
template <int LPR>
__global__ void __launch_bounds__(kThreads) gather_kernel(const float4* __restrict__ buf, const int32_t* __restrict__ idx,
                                                          int64_t n_gathers, float* __restrict__ out) {
  constexpr int RPW = 32 / LPR;   // rows per warp instruction
  const int lane = threadIdx.x & 31, sub = lane % LPR, rsel = lane / LPR;
  const int64_t warp = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * kThreads) >> 5;
  float acc = 0.f;
  for (int64_t g = warp * RPW; g < n_gathers; g += n_warps * RPW * kUnroll) {
    int r[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t j = g + (int64_t)u * n_warps * RPW + rsel;
      r[u] = j < n_gathers ? __ldg(idx + j) : -1;
    }
    float4 v[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) v[u] = r[u] >= 0 ? __ldg(buf + (int64_t)r[u] * LPR + sub) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
  }
  if (acc == 123.456f) out[0] = acc;
}

template <typename F>
static float time_ms(F launch) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) launch();
  CK(cudaEventRecord(e0));
  for (int i = 0; i < 20; ++i) launch();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / 20.f;
}

int main() {
  const int grid = 148 * 8;
  float* out;
  CK(cudaMalloc(&out, 4));
  const int64_t n_gathers = 8 << 20;   // 8 M row reads per launch
  std::vector<int32_t> h((size_t)n_gathers);
  int32_t* idx;
  CK(cudaMalloc(&idx, n_gathers * 4));
  printf("[\n");
  bool first = true;
  for (int64_t mb : {32, 96, 1024}) {
    const int64_t bytes = mb << 20, n4 = bytes / 16;
    float4* buf;
    CK(cudaMalloc(&buf, bytes));
    CK(cudaMemset(buf, 0, bytes));
    const int reps = mb <= 96 ? (int)(512 / mb) : 1;   // stream: several passes over an L2-resident buffer per timing
    float ms = time_ms([&] { for (int r = 0; r < reps; ++r) stream_kernel<<<grid, kThreads>>>(buf, n4, out); });
    printf("%s {\"pattern\": \"stream\", \"buffer_mb\": %lld, \"gbps\": %.0f}", first ? "" : ",\n", (long long)mb, bytes * (double)reps / ms / 1e6);
    first = false;
    for (int lpr : {4, 8, 16}) {
      const int64_t rows = bytes / (16 * lpr);
      uint64_t s = 0x9E3779B97F4A7C15ull;
      for (auto& v : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; v = (int32_t)(s % (uint64_t)rows); }
      CK(cudaMemcpy(idx, h.data(), n_gathers * 4, cudaMemcpyHostToDevice));
      if (lpr == 4) ms = time_ms([&] { gather_kernel<4><<<grid, kThreads>>>(buf, idx, n_gathers, out); });
      else if (lpr == 8) ms = time_ms([&] { gather_kernel<8><<<grid, kThreads>>>(buf, idx, n_gathers, out); });
      else ms = time_ms([&] { gather_kernel<16><<<grid, kThreads>>>(buf, idx, n_gathers, out); });
      printf(",\n {\"pattern\": \"gather\", \"row_bytes\": %d, \"buffer_mb\": %lld, \"row_gbps\": %.0f, \"rows_per_us\": %.0f}", 16 * lpr,
             (long long)mb, n_gathers * 16.0 * lpr / ms / 1e6, n_gathers / ms / 1e3);
    }
    CK(cudaFree(buf));
  }
  printf("\n]\n");
  CK(cudaGetLastError());
  return 0;
}
