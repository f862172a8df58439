// Fused BatchNorm (+ ReLU) over the rows of a sparse tensor's feature matrix.
//
// Replaces ME.MinkowskiBatchNorm + ME.MinkowskiReLU as the reference stacks them after every sparse convolution
// (torch_points3d/modules/MinkowskiEngine/api_modules.py:40-41,53-54,269-270: conv -> BN -> ReLU; the 1x1
// shortcut branch is conv -> BN without ReLU).  MinkowskiBatchNorm is nn.BatchNorm1d on F [N, C]: batch
// statistics over all active rows, biased variance for normalisation, unbiased for the running estimate.
//
// Two passes each way instead of torch's four-plus kernels per layer and direction:
//   forward : (1) per-channel sum / sum of squares   (2) normalise + affine + ReLU, save mean / invstd
//   backward: (1) per-channel sum(g), sum(g * xhat) with g = dY * [Y > 0]   (2) dX
// HBM-bound elementwise work: rows are read with 16-byte loads, a warp covers 128 contiguous channels-bytes.
// Per-channel partial sums are accumulated per thread in fp32 over <= 64 rows, across the CTA in shared memory
// and across CTAs with fp64 atomics (so the variance does not suffer from cancellation).
#include "common.cuh"

namespace pgs {

constexpr int kBnThreads = 256;
constexpr int kBnRowsPerBlock = 512;

// Thread block for the two reduction kernels: x <-> channel quad (C / 4 threads, so a thread always sees the
// same four channels and a row is one coalesced C*4-byte read), y <-> row slot (floor(256 / (C/4)) rows side by
// side).  Per-thread fp32 partials over <= kBnRowsPerBlock / rows_y rows, summed over y through shared memory,
// then ONE fp64 atomicAdd per channel and block.
//
// sums[0..C) = sum x, sums[C..2C) = sum x^2   (fp64, zeroed by the caller)
__global__ void __launch_bounds__(kBnThreads) bn_stats_kernel(const float* __restrict__ X, int64_t n, int C,
                                                               double* __restrict__ sums) {
  extern __shared__ float bn_sm[];   // [rows_y][2C]
  const int cq = threadIdx.x, ry = threadIdx.y, R = blockDim.y;
  const int64_t row_begin = (int64_t)blockIdx.x * kBnRowsPerBlock;
  const int64_t row_end = (row_begin + kBnRowsPerBlock < n) ? row_begin + kBnRowsPerBlock : n;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
  for (int64_t r = row_begin + ry; r < row_end; r += R) {
    const float4 v = __ldg((const float4*)(X + r * C) + cq);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
  }
  float* mine = bn_sm + (size_t)ry * 2 * C;
  *(float4*)&mine[4 * cq] = s;
  *(float4*)&mine[C + 4 * cq] = q;
  __syncthreads();
  for (int i = ry * blockDim.x + cq; i < 2 * C; i += blockDim.x * R) {
    float t = 0.f;
    for (int y = 0; y < R; ++y) t += bn_sm[(size_t)y * 2 * C + i];
    atomicAdd(&sums[i], (double)t);
  }
}

// mean / invstd from the sums (training) or from the running estimates (eval); running update by block 0
__global__ void __launch_bounds__(kBnThreads) bn_apply_kernel(
    const float* __restrict__ X, int64_t n, int C, const double* __restrict__ sums, const float* __restrict__ weight,
    const float* __restrict__ bias, float* __restrict__ running_mean, float* __restrict__ running_var, int training,
    float momentum, float eps, int relu, float* __restrict__ save_mean, float* __restrict__ save_invstd,
    float* __restrict__ Y) {
  extern __shared__ float bn_sm[];   // [2][C]: scale, shift
  for (int c = threadIdx.x; c < C; c += kBnThreads) {
    float mean, invstd;
    if (training) {
      const double m = sums[c] / (double)n;
      double var = sums[C + c] / (double)n - m * m;
      if (var < 0.0) var = 0.0;
      mean = (float)m;
      invstd = (float)(1.0 / sqrt(var + (double)eps));
      if (blockIdx.x == 0) {
        save_mean[c] = mean;
        save_invstd[c] = invstd;
        if (running_mean) {
          const double unbiased = n > 1 ? var * (double)n / (double)(n - 1) : var;
          running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
          running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
        }
      }
    } else {
      mean = running_mean[c];
      invstd = rsqrtf(running_var[c] + eps);
      if (blockIdx.x == 0) {
        save_mean[c] = mean;
        save_invstd[c] = invstd;
      }
    }
    const float w = weight ? weight[c] : 1.f, b = bias ? bias[c] : 0.f;
    bn_sm[c] = invstd * w;
    bn_sm[C + c] = b - mean * invstd * w;
  }
  __syncthreads();
  const int c4 = C / 4;
  const int64_t total = n * c4;
  for (int64_t e = (int64_t)blockIdx.x * kBnThreads + threadIdx.x; e < total; e += (int64_t)gridDim.x * kBnThreads) {
    const int cq = (int)(e % c4);
    const float4 v = __ldg((const float4*)X + e);
    float4 y;
    y.x = fmaf(v.x, bn_sm[4 * cq + 0], bn_sm[C + 4 * cq + 0]);
    y.y = fmaf(v.y, bn_sm[4 * cq + 1], bn_sm[C + 4 * cq + 1]);
    y.z = fmaf(v.z, bn_sm[4 * cq + 2], bn_sm[C + 4 * cq + 2]);
    y.w = fmaf(v.w, bn_sm[4 * cq + 3], bn_sm[C + 4 * cq + 3]);
    if (relu) {
      y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f);
    }
    ((float4*)Y)[e] = y;
  }
}

// sums[0..C) = sum g, sums[C..2C) = sum g * xhat,  g = dY * [Y > 0] (relu) or dY   (same block shape as bn_stats)
__global__ void __launch_bounds__(kBnThreads) bn_bwd_stats_kernel(
    const float* __restrict__ X, const float* __restrict__ Y, const float* __restrict__ dY, int64_t n, int C,
    const float* __restrict__ save_mean, const float* __restrict__ save_invstd, int relu, double* __restrict__ sums) {
  extern __shared__ float bn_sm[];   // [rows_y][2C]
  const int cq = threadIdx.x, ry = threadIdx.y, R = blockDim.y;
  const float4 mean = *(const float4*)&save_mean[4 * cq];
  const float4 inv = *(const float4*)&save_invstd[4 * cq];
  const int64_t row_begin = (int64_t)blockIdx.x * kBnRowsPerBlock;
  const int64_t row_end = (row_begin + kBnRowsPerBlock < n) ? row_begin + kBnRowsPerBlock : n;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
  for (int64_t r = row_begin + ry; r < row_end; r += R) {
    const float4 x = __ldg((const float4*)(X + r * C) + cq);
    float4 g = __ldg((const float4*)(dY + r * C) + cq);
    if (relu) {
      const float4 y = __ldg((const float4*)(Y + r * C) + cq);
      g.x = y.x > 0.f ? g.x : 0.f; g.y = y.y > 0.f ? g.y : 0.f; g.z = y.z > 0.f ? g.z : 0.f; g.w = y.w > 0.f ? g.w : 0.f;
    }
    s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
    q.x = fmaf(g.x, (x.x - mean.x) * inv.x, q.x);
    q.y = fmaf(g.y, (x.y - mean.y) * inv.y, q.y);
    q.z = fmaf(g.z, (x.z - mean.z) * inv.z, q.z);
    q.w = fmaf(g.w, (x.w - mean.w) * inv.w, q.w);
  }
  float* mine = bn_sm + (size_t)ry * 2 * C;
  *(float4*)&mine[4 * cq] = s;
  *(float4*)&mine[C + 4 * cq] = q;
  __syncthreads();
  for (int i = ry * blockDim.x + cq; i < 2 * C; i += blockDim.x * R) {
    float t = 0.f;
    for (int y = 0; y < R; ++y) t += bn_sm[(size_t)y * 2 * C + i];
    atomicAdd(&sums[i], (double)t);
  }
}

// training: dX = w * invstd * (g - mean(g) - xhat * mean(g * xhat));  eval: dX = w * invstd * g
// dweight = sum(g * xhat), dbias = sum(g)  (written by block 0)
__global__ void __launch_bounds__(kBnThreads) bn_bwd_apply_kernel(
    const float* __restrict__ X, const float* __restrict__ Y, const float* __restrict__ dY, int64_t n, int C,
    const float* __restrict__ save_mean, const float* __restrict__ save_invstd, const float* __restrict__ weight,
    const double* __restrict__ sums, int training, int relu, int accumulate, float* __restrict__ dX,
    float* __restrict__ dweight, float* __restrict__ dbias) {
  extern __shared__ float bn_sm[];   // [4][C]: mean, invstd * w, mean_g, mean_gh * invstd... see below
  float* a_mean = bn_sm;
  float* a_inv = bn_sm + C;
  float* a_k1 = bn_sm + 2 * C;   // w * invstd
  float* a_mg = bn_sm + 3 * C;   // mean(g)
  float* a_mgh = bn_sm + 4 * C;  // mean(g * xhat)
  for (int c = threadIdx.x; c < C; c += kBnThreads) {
    const float w = weight ? weight[c] : 1.f;
    a_mean[c] = save_mean[c];
    a_inv[c] = save_invstd[c];
    a_k1[c] = w * save_invstd[c];
    a_mg[c] = training ? (float)(sums[c] / (double)n) : 0.f;
    a_mgh[c] = training ? (float)(sums[C + c] / (double)n) : 0.f;
    if (blockIdx.x == 0) {
      if (dbias) dbias[c] = (accumulate ? dbias[c] : 0.f) + (float)sums[c];
      if (dweight) dweight[c] = (accumulate ? dweight[c] : 0.f) + (float)sums[C + c];
    }
  }
  __syncthreads();
  const int c4 = C / 4;
  const int64_t total = n * c4;
  for (int64_t e = (int64_t)blockIdx.x * kBnThreads + threadIdx.x; e < total; e += (int64_t)gridDim.x * kBnThreads) {
    const int cq = (int)(e % c4);
    const float4 x = __ldg((const float4*)X + e);
    float4 g = __ldg((const float4*)dY + e);
    if (relu) {
      const float4 y = __ldg((const float4*)Y + e);
      g.x = y.x > 0.f ? g.x : 0.f; g.y = y.y > 0.f ? g.y : 0.f; g.z = y.z > 0.f ? g.z : 0.f; g.w = y.w > 0.f ? g.w : 0.f;
    }
    const float xv[4] = {x.x, x.y, x.z, x.w}, gv[4] = {g.x, g.y, g.z, g.w};
    float o[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int c = 4 * cq + t;
      const float h = (xv[t] - a_mean[c]) * a_inv[c];
      o[t] = a_k1[c] * (gv[t] - a_mg[c] - h * a_mgh[c]);
    }
    ((float4*)dX)[e] = make_float4(o[0], o[1], o[2], o[3]);
  }
}


// ---------------------------------------------------------------------------------------------
// feature-matrix glue of the U-Net: residual add, channel concatenation and its split (ME.cat / "+" on SparseTensors,
// api_modules.py:76-82,306-311) -- float4 grid-stride kernels, rows are multiples of 4 channels
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) add2_kernel(const float4* __restrict__ a, const float4* __restrict__ b,
                                                   float4* __restrict__ y, int64_t n4) {
  for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < n4; e += (int64_t)gridDim.x * 256) {
    const float4 u = __ldg(a + e), v = __ldg(b + e);
    y[e] = make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w);
  }
}
// y[r] = [a[r] | b[r]]  (split = 0)   or   a[r], b[r] = halves of y[r]  (split = 1)
__global__ void __launch_bounds__(256) cat2_kernel(float4* __restrict__ a, int ca4, float4* __restrict__ b, int cb4,
                                                   float4* __restrict__ y, int64_t n, int split) {
  const int c4 = ca4 + cb4;
  const int64_t total = n * c4;
  for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (int64_t)gridDim.x * 256) {
    const int64_t r = e / c4;
    const int c = (int)(e - r * c4);
    float4* src = (c < ca4) ? (a + r * ca4 + c) : (b + r * cb4 + (c - ca4));
    if (split)
      *src = y[e];
    else
      y[e] = *src;
  }
}

static inline int bn_apply_grid(int64_t total4) {
  int64_t g = (total4 + kBnThreads - 1) / kBnThreads;
  const int64_t cap = (int64_t)kNumSM * 8;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace pgs

using namespace pgs;

extern "C" {

int pgs_bn_forward(const float* X, int64_t n, int32_t C, const float* weight, const float* bias, float* running_mean,
                   float* running_var, int32_t training, float momentum, float eps, int32_t relu, double* sums,
                   float* save_mean, float* save_invstd, float* Y, void* stream) {
  return pgs_bn_forward_ex(X, n, C, weight, bias, running_mean, running_var, training, momentum, eps, relu, 0, sums,
                           save_mean, save_invstd, Y, stream);
}

int pgs_bn_forward_ex(const float* X, int64_t n, int32_t C, const float* weight, const float* bias, float* running_mean,
                      float* running_var, int32_t training, float momentum, float eps, int32_t relu, int32_t flags,
                      double* sums, float* save_mean, float* save_invstd, float* Y, void* stream) {
  PGS_CHECK_ARG(C >= 4 && C % 4 == 0 && C <= 1024, "channel count must be a multiple of 4, at most 1024");
  PGS_CHECK_ARG(training || (running_mean && running_var), "eval mode needs running statistics");
  if (n == 0) return PGS_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (training) {
    if (!(flags & PGS_BN_SUMS_ZEROED)) PGS_CUDA(cudaMemsetAsync(sums, 0, 2 * C * sizeof(double), s));
    const unsigned g = (unsigned)((n + kBnRowsPerBlock - 1) / kBnRowsPerBlock);
    const dim3 blk(C / 4, kBnThreads / (C / 4));
    bn_stats_kernel<<<g, blk, (size_t)blk.y * 2 * C * sizeof(float), s>>>(X, n, C, sums);
    count_launch();
  }
  bn_apply_kernel<<<bn_apply_grid(n * (C / 4)), kBnThreads, 2 * C * sizeof(float), s>>>(
      X, n, C, sums, weight, bias, running_mean, running_var, training, momentum, eps, relu, save_mean, save_invstd, Y);
  count_launch();
  PGS_CHECK_LAUNCH();
  return PGS_OK;
}

int pgs_bn_backward(const float* X, const float* Y, const float* dY, int64_t n, int32_t C, const float* weight,
                    const float* save_mean, const float* save_invstd, int32_t training, int32_t relu, double* sums,
                    float* dX, float* dweight, float* dbias, void* stream) {
  return pgs_bn_backward_ex(X, Y, dY, n, C, weight, save_mean, save_invstd, training, relu, 0, sums, dX, dweight, dbias,
                            stream);
}

int pgs_bn_backward_ex(const float* X, const float* Y, const float* dY, int64_t n, int32_t C, const float* weight,
                       const float* save_mean, const float* save_invstd, int32_t training, int32_t relu, int32_t flags,
                       double* sums, float* dX, float* dweight, float* dbias, void* stream) {
  PGS_CHECK_ARG(C >= 4 && C % 4 == 0 && C <= 1024, "channel count must be a multiple of 4, at most 1024");
  PGS_CHECK_ARG(!relu || Y != nullptr, "the ReLU mask needs the forward output");
  if (n == 0) return PGS_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (!(flags & PGS_BN_SUMS_ZEROED)) PGS_CUDA(cudaMemsetAsync(sums, 0, 2 * C * sizeof(double), s));
  const unsigned g = (unsigned)((n + kBnRowsPerBlock - 1) / kBnRowsPerBlock);
  const dim3 blk(C / 4, kBnThreads / (C / 4));
  bn_bwd_stats_kernel<<<g, blk, (size_t)blk.y * 2 * C * sizeof(float), s>>>(X, Y, dY, n, C, save_mean, save_invstd, relu,
                                                                          sums);
  bn_bwd_apply_kernel<<<bn_apply_grid(n * (C / 4)), kBnThreads, 5 * C * sizeof(float), s>>>(
      X, Y, dY, n, C, save_mean, save_invstd, weight, sums, training, relu, (flags & PGS_BN_ACCUMULATE_PARAM_GRADS) ? 1 : 0,
      dX, dweight, dbias);
  count_launch(2);
  PGS_CHECK_LAUNCH();
  return PGS_OK;
}

int pgs_add2(const float* a, const float* b, float* y, int64_t n_elems, void* stream) {
  PGS_CHECK_ARG(n_elems % 4 == 0, "element count must be a multiple of 4");
  if (n_elems == 0) return PGS_OK;
  add2_kernel<<<bn_apply_grid(n_elems / 4), 256, 0, (cudaStream_t)stream>>>((const float4*)a, (const float4*)b, (float4*)y,
                                                                            n_elems / 4);
  count_launch();
  PGS_CHECK_LAUNCH();
  return PGS_OK;
}

int pgs_cat2(float* a, int32_t ca, float* b, int32_t cb, float* y, int64_t n, int32_t split, void* stream) {
  PGS_CHECK_ARG(ca % 4 == 0 && cb % 4 == 0 && ca > 0 && cb > 0, "channel counts must be positive multiples of 4");
  if (n == 0) return PGS_OK;
  cat2_kernel<<<bn_apply_grid(n * ((ca + cb) / 4)), 256, 0, (cudaStream_t)stream>>>((float4*)a, ca / 4, (float4*)b, cb / 4,
                                                                                   (float4*)y, n, split);
  count_launch();
  PGS_CHECK_LAUNCH();
  return PGS_OK;
}

}  // extern "C"
