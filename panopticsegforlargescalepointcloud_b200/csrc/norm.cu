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
constexpr int kBnUnroll = 8;            // rows a thread has in flight in the reduction kernels
constexpr int kBnMaxRowsPerThread = 32;   // fp32 partial sums never run over more rows than this

// scale / shift of the normalise + affine step; spelled with explicit roundings because the backward kernels
// recompute the ReLU mask [y > 0] from x and must land on the forward's bits
__device__ __forceinline__ void bn_scale_shift(float mean, float invstd, float w, float b, float& scale, float& shift) {
  scale = __fmul_rn(invstd, w);
  shift = __fsub_rn(b, __fmul_rn(__fmul_rn(mean, invstd), w));
}
__device__ __forceinline__ float bn_affine(float x, float scale, float shift) { return __fmaf_rn(x, scale, shift); }

// Thread block for the two reduction kernels: x <-> channel quad (C / 4 threads, so a thread always sees the
// same four channels and a row is one coalesced C*4-byte read), y <-> row slot (floor(256 / (C/4)) rows side by
// side).  A block owns rows_per_block = rows_per_thread * blockDim.y rows; a thread keeps kBnUnroll row loads in
// flight, sums its <= 32 rows in fp32, the block sums the row slots in fp64 through shared memory, then ONE fp64
// atomicAdd per channel and block.  (Round 2: 8 rows per thread and one load in flight ran these two kernels at
// ~2 TB/s; the block count is now chosen so that the grid still covers the SMs twice.)
//
// sums[0..C) = sum x, sums[C..2C) = sum x^2   (fp64, zeroed by the caller)
__global__ void __launch_bounds__(kBnThreads) bn_stats_kernel(const float* __restrict__ X, int64_t n, int C,
                                                               int rows_per_block, double* __restrict__ sums) {
  extern __shared__ float bn_sm[];   // [rows_y][2C]
  const int cq = threadIdx.x, ry = threadIdx.y, R = blockDim.y;
  const int64_t row_begin = (int64_t)blockIdx.x * rows_per_block;
  const int64_t row_end = (row_begin + rows_per_block < n) ? row_begin + rows_per_block : n;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
  for (int64_t r = row_begin + ry; r < row_end; r += (int64_t)R * kBnUnroll) {
    float4 v[kBnUnroll];
#pragma unroll
    for (int u = 0; u < kBnUnroll; ++u) {
      const int64_t rr = r + (int64_t)u * R;
      v[u] = (rr < row_end) ? __ldg((const float4*)(X + rr * C) + cq) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < kBnUnroll; ++u) {
      s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w;
      q.x = fmaf(v[u].x, v[u].x, q.x); q.y = fmaf(v[u].y, v[u].y, q.y);
      q.z = fmaf(v[u].z, v[u].z, q.z); q.w = fmaf(v[u].w, v[u].w, q.w);
    }
  }
  float* mine = bn_sm + (size_t)ry * 2 * C;
  *(float4*)&mine[4 * cq] = s;
  *(float4*)&mine[C + 4 * cq] = q;
  __syncthreads();
  for (int i = ry * blockDim.x + cq; i < 2 * C; i += blockDim.x * R) {
    double t = 0.0;
    for (int y = 0; y < R; ++y) t += (double)bn_sm[(size_t)y * 2 * C + i];
    atomicAdd(&sums[i], t);
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
    bn_scale_shift(mean, invstd, weight ? weight[c] : 1.f, bias ? bias[c] : 0.f, bn_sm[c], bn_sm[C + c]);
  }
  __syncthreads();
  const int c4 = C / 4;
  const int64_t total = n * c4;
  for (int64_t e = (int64_t)blockIdx.x * kBnThreads + threadIdx.x; e < total; e += (int64_t)gridDim.x * kBnThreads) {
    const int cq = (int)(e % c4);
    const float4 v = __ldg((const float4*)X + e);
    float4 y;
    y.x = bn_affine(v.x, bn_sm[4 * cq + 0], bn_sm[C + 4 * cq + 0]);
    y.y = bn_affine(v.y, bn_sm[4 * cq + 1], bn_sm[C + 4 * cq + 1]);
    y.z = bn_affine(v.z, bn_sm[4 * cq + 2], bn_sm[C + 4 * cq + 2]);
    y.w = bn_affine(v.w, bn_sm[4 * cq + 3], bn_sm[C + 4 * cq + 3]);
    if (relu) {
      y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f);
    }
    ((float4*)Y)[e] = y;
  }
}

// sums[0..C) = sum g, sums[C..2C) = sum g * xhat,  g = dY * [y > 0] (relu) or dY   (same block shape as bn_stats).
// The mask comes from the forward output Y when the caller passes it, else it is recomputed from x (one array less
// to read: y = fma(x, scale, shift) with the forward's own roundings).
__global__ void __launch_bounds__(kBnThreads) bn_bwd_stats_kernel(
    const float* __restrict__ X, const float* __restrict__ Y, const float* __restrict__ dY, int64_t n, int C,
    const float* __restrict__ save_mean, const float* __restrict__ save_invstd, const float* __restrict__ weight,
    const float* __restrict__ bias, int relu, int rows_per_block, double* __restrict__ sums) {
  extern __shared__ float bn_sm[];   // [rows_y][2C]
  constexpr int UB = kBnUnroll / 2;
  const int cq = threadIdx.x, ry = threadIdx.y, R = blockDim.y;
  const float4 mean = *(const float4*)&save_mean[4 * cq];
  const float4 inv = *(const float4*)&save_invstd[4 * cq];
  float sc[4], sh[4];
  {
    const float mv[4] = {mean.x, mean.y, mean.z, mean.w}, iv[4] = {inv.x, inv.y, inv.z, inv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      bn_scale_shift(mv[j], iv[j], weight ? weight[4 * cq + j] : 1.f, bias ? bias[4 * cq + j] : 0.f, sc[j], sh[j]);
  }
  const bool mask_y = relu && Y != nullptr, mask_x = relu && Y == nullptr;
  const int64_t row_begin = (int64_t)blockIdx.x * rows_per_block;
  const int64_t row_end = (row_begin + rows_per_block < n) ? row_begin + rows_per_block : n;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
  for (int64_t r = row_begin + ry; r < row_end; r += (int64_t)R * UB) {
    float4 x[UB], g[UB], y[UB];
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const int64_t rr = r + (int64_t)u * R;
      const bool ok = rr < row_end;
      x[u] = ok ? __ldg((const float4*)(X + rr * C) + cq) : make_float4(0.f, 0.f, 0.f, 0.f);
      g[u] = ok ? __ldg((const float4*)(dY + rr * C) + cq) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (mask_y) y[u] = ok ? __ldg((const float4*)(Y + rr * C) + cq) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      if (mask_x) {
        y[u] = make_float4(bn_affine(x[u].x, sc[0], sh[0]), bn_affine(x[u].y, sc[1], sh[1]),
                           bn_affine(x[u].z, sc[2], sh[2]), bn_affine(x[u].w, sc[3], sh[3]));
      }
      if (relu) {
        g[u].x = y[u].x > 0.f ? g[u].x : 0.f; g[u].y = y[u].y > 0.f ? g[u].y : 0.f;
        g[u].z = y[u].z > 0.f ? g[u].z : 0.f; g[u].w = y[u].w > 0.f ? g[u].w : 0.f;
      }
      s.x += g[u].x; s.y += g[u].y; s.z += g[u].z; s.w += g[u].w;
      q.x = fmaf(g[u].x, (x[u].x - mean.x) * inv.x, q.x);
      q.y = fmaf(g[u].y, (x[u].y - mean.y) * inv.y, q.y);
      q.z = fmaf(g[u].z, (x[u].z - mean.z) * inv.z, q.z);
      q.w = fmaf(g[u].w, (x[u].w - mean.w) * inv.w, q.w);
    }
  }
  float* mine = bn_sm + (size_t)ry * 2 * C;
  *(float4*)&mine[4 * cq] = s;
  *(float4*)&mine[C + 4 * cq] = q;
  __syncthreads();
  for (int i = ry * blockDim.x + cq; i < 2 * C; i += blockDim.x * R) {
    double t = 0.0;
    for (int y = 0; y < R; ++y) t += (double)bn_sm[(size_t)y * 2 * C + i];
    atomicAdd(&sums[i], t);
  }
}

// training: dX = w * invstd * (g - mean(g) - xhat * mean(g * xhat));  eval: dX = w * invstd * g
// dweight = sum(g * xhat), dbias = sum(g)  (written by block 0)
__global__ void __launch_bounds__(kBnThreads) bn_bwd_apply_kernel(
    const float* __restrict__ X, const float* __restrict__ Y, const float* __restrict__ dY, int64_t n, int C,
    const float* __restrict__ save_mean, const float* __restrict__ save_invstd, const float* __restrict__ weight,
    const float* __restrict__ bias, const double* __restrict__ sums, int training, int relu, int accumulate,
    float* __restrict__ dX, float* __restrict__ dweight, float* __restrict__ dbias) {
  extern __shared__ float bn_sm[];   // [7][C]
  float* a_mean = bn_sm;
  float* a_inv = bn_sm + C;
  float* a_k1 = bn_sm + 2 * C;   // w * invstd
  float* a_mg = bn_sm + 3 * C;   // mean(g)
  float* a_mgh = bn_sm + 4 * C;  // mean(g * xhat)
  float* a_sc = bn_sm + 5 * C;   // forward scale / shift (ReLU mask from x)
  float* a_sh = bn_sm + 6 * C;
  for (int c = threadIdx.x; c < C; c += kBnThreads) {
    const float w = weight ? weight[c] : 1.f;
    a_mean[c] = save_mean[c];
    a_inv[c] = save_invstd[c];
    a_k1[c] = w * save_invstd[c];
    a_mg[c] = training ? (float)(sums[c] / (double)n) : 0.f;
    a_mgh[c] = training ? (float)(sums[C + c] / (double)n) : 0.f;
    bn_scale_shift(save_mean[c], save_invstd[c], w, bias ? bias[c] : 0.f, a_sc[c], a_sh[c]);
    if (blockIdx.x == 0) {
      if (dbias) dbias[c] = (accumulate ? dbias[c] : 0.f) + (float)sums[c];
      if (dweight) dweight[c] = (accumulate ? dweight[c] : 0.f) + (float)sums[C + c];
    }
  }
  __syncthreads();
  const bool mask_y = relu && Y != nullptr, mask_x = relu && Y == nullptr;
  const int c4 = C / 4;
  const int64_t total = n * c4;
  for (int64_t e = (int64_t)blockIdx.x * kBnThreads + threadIdx.x; e < total; e += (int64_t)gridDim.x * kBnThreads) {
    const int cq = (int)(e % c4);
    const float4 x = __ldg((const float4*)X + e);
    float4 g = __ldg((const float4*)dY + e);
    const float xv[4] = {x.x, x.y, x.z, x.w};
    float gv[4] = {g.x, g.y, g.z, g.w};
    if (mask_y) {
      const float4 y = __ldg((const float4*)Y + e);
      gv[0] = y.x > 0.f ? gv[0] : 0.f; gv[1] = y.y > 0.f ? gv[1] : 0.f;
      gv[2] = y.z > 0.f ? gv[2] : 0.f; gv[3] = y.w > 0.f ? gv[3] : 0.f;
    } else if (mask_x) {
#pragma unroll
      for (int t = 0; t < 4; ++t) gv[t] = bn_affine(xv[t], a_sc[4 * cq + t], a_sh[4 * cq + t]) > 0.f ? gv[t] : 0.f;
    }
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

// rows per block of the reduction kernels: as many rows per thread as still leave >= 2 blocks per SM
static inline int bn_rows_per_block(int64_t n, int rows_y) {
  int rpt = kBnMaxRowsPerThread;
  while (rpt > 4 && (n + (int64_t)rpt * rows_y - 1) / ((int64_t)rpt * rows_y) < 2 * kNumSM) rpt >>= 1;
  return rpt * rows_y;
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
    const dim3 blk(C / 4, kBnThreads / (C / 4));
    const int rpb = bn_rows_per_block(n, blk.y);
    const unsigned g = (unsigned)((n + rpb - 1) / rpb);
    bn_stats_kernel<<<g, blk, (size_t)blk.y * 2 * C * sizeof(float), s>>>(X, n, C, rpb, sums);
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
  PGS_CHECK_ARG(!relu || Y != nullptr, "the ReLU mask needs the forward output");
  return pgs_bn_backward_ex(X, Y, dY, n, C, weight, nullptr, save_mean, save_invstd, training, relu, 0, sums, dX, dweight,
                            dbias, stream);
}

int pgs_bn_backward_ex(const float* X, const float* Y, const float* dY, int64_t n, int32_t C, const float* weight,
                       const float* bias, const float* save_mean, const float* save_invstd, int32_t training, int32_t relu,
                       int32_t flags,
                       double* sums, float* dX, float* dweight, float* dbias, void* stream) {
  PGS_CHECK_ARG(C >= 4 && C % 4 == 0 && C <= 1024, "channel count must be a multiple of 4, at most 1024");
  if (n == 0) return PGS_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (!(flags & PGS_BN_SUMS_ZEROED)) PGS_CUDA(cudaMemsetAsync(sums, 0, 2 * C * sizeof(double), s));
  const dim3 blk(C / 4, kBnThreads / (C / 4));
  const int rpb = bn_rows_per_block(n, blk.y);
  const unsigned g = (unsigned)((n + rpb - 1) / rpb);
  bn_bwd_stats_kernel<<<g, blk, (size_t)blk.y * 2 * C * sizeof(float), s>>>(X, Y, dY, n, C, save_mean, save_invstd, weight,
                                                                          bias, relu, rpb, sums);
  bn_bwd_apply_kernel<<<bn_apply_grid(n * (C / 4)), kBnThreads, 7 * C * sizeof(float), s>>>(
      X, Y, dY, n, C, save_mean, save_invstd, weight, bias, sums, training, relu,
      (flags & PGS_BN_ACCUMULATE_PARAM_GRADS) ? 1 : 0, dX, dweight, dbias);
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
