// Flat-kernel mean shift on learned embeddings (the clustering that the shipped paper settings I, IV, V
// actually run on the embedding head).
//
// Replaces sklearn.cluster.MeanShift(bandwidth=0.6, bin_seeding=True).fit(X) as called by
// torch_points3d/utils/meanshift_cluster.py:9-18 (per scene, from cluster_single :72-123; call sites
// models/panoptic/PointGroup3heads.py:235,276,323,376 and pointgroupembed.py:491,532).  The algorithm restated from
// scikit-learn 1.9 (sklearn/cluster/_mean_shift.py:108-132 single-seed iteration, :247-297 bin seeding, :470-560 fit):
//   seed s:  repeat { P = points within `bandwidth` of the mean; if P is empty: drop the seed;
//                     old = mean; mean = average(P);  stop when |mean - old| <= 1e-3 * bandwidth or after max_iter }
//            -> (mean, |P| of the last query)
// then (host, a few thousand centres) sort by (|P|, mean) descending, greedily drop centres within `bandwidth` of a kept
// one, and label every point with its nearest kept centre.
//
// pgs_ms_iterate: one CTA per seed, all points streamed from L2 every iteration (n <= a few 100 k rows of D <= 8 floats:
// 2-10 MB, L2 resident), distances and sums in fp64 like the KD-tree query of the reference (points and means are fp32
// values, so the in / out decision of every point is the reference's up to the rounding of the mean itself).
// pgs_ms_assign: nearest centre per point, fp64, ties to the lower centre index.
#include "common.cuh"

namespace pgs {

constexpr int kMsThreads = 256;
constexpr int kMsMaxD = 8;

template <int D>
__global__ void __launch_bounds__(kMsThreads) ms_iterate_kernel(const float* __restrict__ X, int64_t n,
                                                                const float* __restrict__ seeds, float bandwidth,
                                                                int max_iter, float* __restrict__ centers,
                                                                int32_t* __restrict__ counts, int32_t* __restrict__ iters) {
  __shared__ double s_sum[kMsThreads / 32][D];
  __shared__ int s_cnt[kMsThreads / 32];
  __shared__ float s_mean[D];
  __shared__ int s_state;   // 0 = continue, 1 = stop
  const int seed = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < D) s_mean[tid] = seeds[(size_t)seed * D + tid];
  if (tid == 0) s_state = 0;
  __syncthreads();
  const double h2 = (double)bandwidth * (double)bandwidth;
  const double stop = 1e-3 * (double)bandwidth;
  int it = 0, last_count = 0;
  for (;;) {
    double m[D], sum[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      m[d] = (double)s_mean[d];
      sum[d] = 0.0;
    }
    int cnt = 0;
    for (int64_t i = tid; i < n; i += kMsThreads) {
      double x[D], d2 = 0.0;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        x[d] = (double)__ldg(&X[i * D + d]);
        const double t = x[d] - m[d];
        d2 = fma(t, t, d2);
      }
      if (d2 <= h2) {
        ++cnt;
#pragma unroll
        for (int d = 0; d < D; ++d) sum[d] += x[d];
      }
    }
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) {
      cnt += __shfl_xor_sync(0xffffffffu, cnt, sft);
#pragma unroll
      for (int d = 0; d < D; ++d) sum[d] += __shfl_xor_sync(0xffffffffu, sum[d], sft);
    }
    if (lane == 0) {
      s_cnt[warp] = cnt;
#pragma unroll
      for (int d = 0; d < D; ++d) s_sum[warp][d] = sum[d];
    }
    __syncthreads();
    if (tid == 0) {
      int c = 0;
      double tot[D];
#pragma unroll
      for (int d = 0; d < D; ++d) tot[d] = 0.0;
      for (int w = 0; w < kMsThreads / 32; ++w) {
        c += s_cnt[w];
#pragma unroll
        for (int d = 0; d < D; ++d) tot[d] += s_sum[w][d];
      }
      s_cnt[0] = c;
      if (c == 0) {
        s_state = 1;   // nothing within the bandwidth: the seed is dropped (count 0)
      } else {
        double mv = 0.0;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const float nm = (float)(tot[d] / (double)c);   // the reference's means are fp32 arrays
          const double t = (double)nm - (double)s_mean[d];
          mv = fma(t, t, mv);
          s_mean[d] = nm;
        }
        if (sqrt(mv) <= stop || it == max_iter) s_state = 1;
      }
    }
    __syncthreads();
    last_count = s_cnt[0];
    if (s_state) break;
    ++it;
    __syncthreads();
  }
  if (tid < D) centers[(size_t)seed * D + tid] = s_mean[tid];
  if (tid == 0) {
    counts[seed] = last_count;
    iters[seed] = it;
  }
}

template <int D>
__global__ void __launch_bounds__(kMsThreads) ms_assign_kernel(const float* __restrict__ X, int64_t n,
                                                               const float* __restrict__ centers, int n_centers,
                                                               int32_t* __restrict__ labels, double* __restrict__ dist) {
  extern __shared__ float s_c[];   // [tile][D]
  const int64_t i = (int64_t)blockIdx.x * kMsThreads + threadIdx.x;
  double x[D];
#pragma unroll
  for (int d = 0; d < D; ++d) x[d] = (i < n) ? (double)X[i * D + d] : 0.0;
  double best = 1e300;
  int arg = -1;
  constexpr int TILE = 1024;
  for (int c0 = 0; c0 < n_centers; c0 += TILE) {
    const int m = (n_centers - c0 < TILE) ? (n_centers - c0) : TILE;
    __syncthreads();
    for (int e = threadIdx.x; e < m * D; e += kMsThreads) s_c[e] = centers[(size_t)c0 * D + e];
    __syncthreads();
    for (int c = 0; c < m; ++c) {
      double d2 = 0.0;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const double t = x[d] - (double)s_c[c * D + d];
        d2 = fma(t, t, d2);
      }
      if (d2 < best) {
        best = d2;
        arg = c0 + c;
      }
    }
  }
  if (i < n) {
    labels[i] = arg;
    if (dist) dist[i] = sqrt(best);
  }
}

}  // namespace pgs

using namespace pgs;

extern "C" {

int pgs_ms_iterate(const float* X, int64_t n, int32_t D, const float* seeds, int64_t n_seeds, float bandwidth,
                   int32_t max_iter, float* centers, int32_t* counts, int32_t* iters, void* stream) {
  PGS_CHECK_ARG(D >= 1 && D <= kMsMaxD, "1..8 feature dimensions are supported");
  PGS_CHECK_ARG(bandwidth > 0.f && max_iter >= 0, "bandwidth must be positive");
  PGS_CHECK_ARG(n_seeds < (1ll << 31), "too many seeds");
  if (n_seeds == 0) return PGS_OK;
  cudaStream_t s = (cudaStream_t)stream;
#define PGS_MS(DD)                                                                                              \
  case DD:                                                                                                      \
    ms_iterate_kernel<DD><<<(unsigned)n_seeds, kMsThreads, 0, s>>>(X, n, seeds, bandwidth, max_iter, centers, counts, iters); \
    break;
  switch (D) {
    PGS_MS(1) PGS_MS(2) PGS_MS(3) PGS_MS(4) PGS_MS(5) PGS_MS(6) PGS_MS(7) PGS_MS(8)
  }
#undef PGS_MS
  count_launch();
  PGS_CHECK_LAUNCH();
  return PGS_OK;
}

int pgs_ms_assign(const float* X, int64_t n, int32_t D, const float* centers, int32_t n_centers, int32_t* labels,
                  double* dist, void* stream) {
  PGS_CHECK_ARG(D >= 1 && D <= kMsMaxD, "1..8 feature dimensions are supported");
  PGS_CHECK_ARG(n_centers >= 1, "at least one centre is required");
  if (n == 0) return PGS_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned g = (unsigned)((n + kMsThreads - 1) / kMsThreads);
  const size_t sm = (size_t)1024 * D * sizeof(float);
#define PGS_MS(DD)                                                                                   \
  case DD:                                                                                           \
    ms_assign_kernel<DD><<<g, kMsThreads, sm, s>>>(X, n, centers, n_centers, labels, dist);          \
    break;
  switch (D) {
    PGS_MS(1) PGS_MS(2) PGS_MS(3) PGS_MS(4) PGS_MS(5) PGS_MS(6) PGS_MS(7) PGS_MS(8)
  }
#undef PGS_MS
  count_launch();
  PGS_CHECK_LAUNCH();
  return PGS_OK;
}

}  // extern "C"
