// Sparse convolution forward / backward-input (one kernel) and backward-weight.
//
// Replaces MinkowskiEngine's ConvolutionForward/BackwardKernelGPU on the reference hot path
// (reference call sites: torch_points3d/modules/MinkowskiEngine/api_modules.py:26-55 ResBlock,
//  :244-270 ResNetDown.conv_in, :293 ResNetUp -> MinkowskiConvolutionTranspose).
//
// Formulation: OUTPUT STATIONARY gather-GEMM.  One CTA owns TM consecutive output rows and a TN
// wide slice of output channels, walks the K kernel offsets, gathers the (<= TM) input rows the
// gather table names for that offset into shared memory, multiplies by the W[k] tile and keeps
// the accumulators in registers; the output row is written exactly once (no atomics, no
// scatter-add, deterministic).  Offsets for which no row of the tile has a neighbour are
// skipped with one block-wide vote.  fp32 FFMA math: the parity bar is 1e-4 against an fp32
// oracle, which rules out single-pass tf32/bf16 tensor-core math.
#include "common.cuh"

namespace pgs {

constexpr int kConvThreads = 256;
constexpr int kTM = 128;  // output rows per CTA
constexpr int kKC = 16;   // input channels per smem stage

template <int TN>
__global__ void __launch_bounds__(kConvThreads) conv_fwd_kernel(
    const float* __restrict__ X, const float* __restrict__ W, const int32_t* __restrict__ nbr,
    int64_t n_q, int K, int c_in, int c_out, int mirror, int w_transposed, float* __restrict__ Y) {
  constexpr int TX = TN / 4;                // threads across the channel tile
  constexpr int TY = kConvThreads / TX;     // threads down the row tile
  constexpr int RM = kTM / TY;              // rows per thread
  constexpr int LDA = kTM + 4;
  __shared__ __align__(16) float As[kKC][LDA];
  __shared__ __align__(16) float Bs[kKC][TN];
  __shared__ int idx_s[kTM];

  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int64_t row0 = (int64_t)blockIdx.x * kTM;
  const int col0 = blockIdx.y * TN;
  const bool vec_in = (c_in & 3) == 0;

  float acc[RM][4];
#pragma unroll
  for (int i = 0; i < RM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k = 0; k < K; ++k) {
    const int tk = mirror ? (K - 1 - k) : k;
    int has = 0;
    if (tid < kTM) {
      const int64_t r = row0 + tid;
      int v = -1;
      if (r < n_q) v = nbr ? __ldg(&nbr[(int64_t)tk * n_q + r]) : (int)r;
      idx_s[tid] = v;
      has = v >= 0;
    }
    if (!__syncthreads_or(has)) continue;  // also publishes idx_s

    const float* Wk = W + (size_t)k * c_in * c_out;
    for (int c0 = 0; c0 < c_in; c0 += kKC) {
      // ---- gather A: kTM rows x kKC channels, stored channel-major ----
      if (vec_in) {
#pragma unroll
        for (int it = 0; it < (kTM * kKC / 4) / kConvThreads; ++it) {
          const int e = tid + it * kConvThreads;
          const int r = e / (kKC / 4), q4 = e % (kKC / 4);
          const int src = idx_s[r];
          const int c = c0 + q4 * 4;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (src >= 0 && c < c_in) v = __ldg((const float4*)(X + (size_t)src * c_in + c));
          As[q4 * 4 + 0][r] = v.x;
          As[q4 * 4 + 1][r] = v.y;
          As[q4 * 4 + 2][r] = v.z;
          As[q4 * 4 + 3][r] = v.w;
        }
      } else {
        for (int e = tid; e < kTM * kKC; e += kConvThreads) {
          const int r = e / kKC, cc = e % kKC;
          const int src = idx_s[r];
          const int c = c0 + cc;
          As[cc][r] = (src >= 0 && c < c_in) ? __ldg(X + (size_t)src * c_in + c) : 0.f;
        }
      }
      // ---- W tile: kKC x TN ----
      if (!w_transposed) {
        for (int e = tid; e < kKC * TN; e += kConvThreads) {
          const int kk = e / TN, j = e % TN;
          const int c = c0 + kk, o = col0 + j;
          Bs[kk][j] = (c < c_in && o < c_out) ? __ldg(Wk + (size_t)c * c_out + o) : 0.f;
        }
      } else {  // stored [K][c_out][c_in]
        for (int e = tid; e < kKC * TN; e += kConvThreads) {
          const int j = e / kKC, kk = e % kKC;
          const int c = c0 + kk, o = col0 + j;
          Bs[kk][j] = (c < c_in && o < c_out) ? __ldg(Wk + (size_t)o * c_in + c) : 0.f;
        }
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < kKC; ++kk) {
        float a[RM];
#pragma unroll
        for (int i = 0; i < RM; ++i) a[i] = As[kk][ty * RM + i];
        const float4 b = *(const float4*)&Bs[kk][tx * 4];
#pragma unroll
        for (int i = 0; i < RM; ++i) {
          acc[i][0] = fmaf(a[i], b.x, acc[i][0]);
          acc[i][1] = fmaf(a[i], b.y, acc[i][1]);
          acc[i][2] = fmaf(a[i], b.z, acc[i][2]);
          acc[i][3] = fmaf(a[i], b.w, acc[i][3]);
        }
      }
      __syncthreads();
    }
  }

  const int col = col0 + tx * 4;
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    const int64_t r = row0 + ty * RM + i;
    if (r >= n_q) continue;
    float* y = Y + (size_t)r * c_out + col;
    if ((c_out & 3) == 0 && col + 3 < c_out) {
      *(float4*)y = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (col + j < c_out) y[j] = acc[i][j];
    }
  }
}

// dW[k] (64x64 tile) += sum over a chunk of the pairs of table offset tk
constexpr int kWT = 64;
constexpr int kPK = 16;
constexpr int kPairChunk = 2048;

__global__ void __launch_bounds__(kConvThreads) conv_bwd_weight_kernel(
    const float* __restrict__ X, const float* __restrict__ dY, const int32_t* __restrict__ in_idx,
    const int32_t* __restrict__ out_idx, const int32_t* __restrict__ offs, int64_t n_identity, int K,
    int mirror, int c_in, int c_out, int tiles_co, float* __restrict__ dW) {
  __shared__ __align__(16) float Xs[kPK][kWT + 4];
  __shared__ __align__(16) float Ds[kPK][kWT + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int ci0 = (blockIdx.x / tiles_co) * kWT, co0 = (blockIdx.x % tiles_co) * kWT;
  const int tk = blockIdx.z;
  const int64_t p_begin = offs ? offs[tk] : 0, p_end = offs ? offs[tk + 1] : n_identity;
  const int64_t p0 = p_begin + (int64_t)blockIdx.y * kPairChunk;
  const int64_t p1 = (p0 + kPairChunk < p_end) ? p0 + kPairChunk : p_end;
  if (p0 >= p1) return;
  float* dWk = dW + (size_t)(mirror ? (K - 1 - tk) : tk) * c_in * c_out;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t pb = p0; pb < p1; pb += kPK) {
    for (int e = tid; e < kPK * kWT; e += kConvThreads) {
      const int pp = e / kWT, c = e % kWT;
      const int64_t p = pb + pp;
      float xv = 0.f, dv = 0.f;
      if (p < p1) {
        const int64_t ri = in_idx ? in_idx[p] : p;
        const int64_t ro = out_idx ? out_idx[p] : p;
        if (ci0 + c < c_in) xv = __ldg(X + ri * c_in + ci0 + c);
        if (co0 + c < c_out) dv = __ldg(dY + ro * c_out + co0 + c);
      }
      Xs[pp][c] = xv;
      Ds[pp][c] = dv;
    }
    __syncthreads();
#pragma unroll
    for (int pp = 0; pp < kPK; ++pp) {
      const float4 a = *(const float4*)&Xs[pp][ty * 4];
      const float4 b = *(const float4*)&Ds[pp][tx * 4];
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ci = ci0 + ty * 4 + i;
    if (ci >= c_in) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + tx * 4 + j;
      if (co < c_out) atomicAdd(dWk + (size_t)ci * c_out + co, acc[i][j]);
    }
  }
}

}  // namespace pgs

using namespace pgs;

extern "C" {

int pgs_conv_fwd(const float* X, const float* W, const int32_t* nbr, int64_t n_q, int32_t K, int32_t c_in,
                 int32_t c_out, int32_t mirror, int32_t w_transposed, float* Y, void* stream) {
  PGS_CHECK_ARG(K >= 1 && c_in >= 1 && c_out >= 1, "bad shape");
  PGS_CHECK_ARG(nbr != nullptr || K == 1, "nbr == NULL requires K == 1");
  if (n_q == 0) return PGS_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned gx = (unsigned)((n_q + kTM - 1) / kTM);
  if (c_out <= 16) {
    conv_fwd_kernel<16><<<dim3(gx, 1), kConvThreads, 0, s>>>(X, W, nbr, n_q, K, c_in, c_out, mirror,
                                                             w_transposed, Y);
  } else if (c_out <= 32) {
    conv_fwd_kernel<32><<<dim3(gx, 1), kConvThreads, 0, s>>>(X, W, nbr, n_q, K, c_in, c_out, mirror,
                                                             w_transposed, Y);
  } else {
    conv_fwd_kernel<64><<<dim3(gx, (c_out + 63) / 64), kConvThreads, 0, s>>>(X, W, nbr, n_q, K, c_in, c_out,
                                                                             mirror, w_transposed, Y);
  }
  count_launch();
  PGS_CHECK_LAUNCH();
  return PGS_OK;
}

int pgs_conv_bwd_weight(const float* X, const float* dY, const int32_t* in_idx, const int32_t* out_idx,
                        const int32_t* offs, int64_t max_pairs, int32_t K, int32_t c_in, int32_t c_out,
                        int32_t mirror, float* dW, void* stream) {
  PGS_CHECK_ARG(K >= 1 && c_in >= 1 && c_out >= 1, "bad shape");
  PGS_CHECK_ARG(offs != nullptr || K == 1, "offs == NULL requires K == 1 (identity pairs)");
  if (max_pairs <= 0) return PGS_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int tiles_ci = (c_in + kWT - 1) / kWT, tiles_co = (c_out + kWT - 1) / kWT;
  const unsigned chunks = (unsigned)((max_pairs + kPairChunk - 1) / kPairChunk);
  conv_bwd_weight_kernel<<<dim3(tiles_ci * tiles_co, chunks, K), kConvThreads, 0, s>>>(
      X, dY, in_idx, out_idx, offs, max_pairs, K, mirror, c_in, c_out, tiles_co, dW);
  count_launch();
  PGS_CHECK_LAUNCH();
  return PGS_OK;
}

}  // extern "C"
