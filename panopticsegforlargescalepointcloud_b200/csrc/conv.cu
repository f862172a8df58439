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
#include <cstdlib>

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

__device__ __forceinline__ void dw_cp_async16(void* dst_smem, const void* src, bool pred) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
  const uint32_t n = pred ? 16u : 0u;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}

// ------------------------------------------------------------------------------------------
// Small-channel forward / input-gradient kernel (c_out in {16, 32}): output stationary, cp.async ring.
//
// Why not the tensor cores here: an M = 128 tcgen05.mma costs >= 97 cycles whatever N (profiles/
// r1_mma_issue_rate.txt), so at N = 16 the 3-pass tf32 path is MMA-issue bound at 162 instructions per 128-row
// tile (131 us per 200k-row layer) while the same layer is only ~0.9 GFLOP of fp32 FMA work.
//
// One CTA owns 128 output rows.  Per pipeline step (one non-empty kernel offset x one slice of <= 32 input
// channels) the gathered input rows land in shared memory by 16-byte cp.async (missing neighbours zero-filled
// by src-size 0) together with the matching [slice x c_out] weight block; thread <-> (row, half of the output
// channels) keeps c_out / 2 accumulators in registers and reads x by broadcast, w by 16-byte broadcast loads.
// ------------------------------------------------------------------------------------------
constexpr int kSfThreads = 256;
constexpr int kSfM = 128;
constexpr int kSfCS = 32;       // input-channel slice per step
constexpr int kSfXS = kSfCS + 4;  // row stride in floats: +16 B so that 8 consecutive rows hit 8 different bank groups
constexpr int kSfStages = 3;
constexpr int kSfMaxK = 27;

template <int COUT>
__global__ void __launch_bounds__(kSfThreads) conv_small_kernel(
    const float* __restrict__ X, const float* __restrict__ W, const int32_t* __restrict__ nbr, int64_t n_q, int K,
    int c_in, int mirror, int w_transposed, float* __restrict__ Y) {
  constexpr int CPT = COUT / 2;            // output channels per thread
  extern __shared__ __align__(16) uint8_t sf_smem[];
  float (*xs)[kSfM][kSfXS] = (float (*)[kSfM][kSfXS])sf_smem;                       // [stage][row][ci], padded
  float (*ws)[kSfCS][COUT] = (float (*)[kSfCS][COUT])(&xs[kSfStages][0][0]);       // [stage][ci][co]
  __shared__ int idx_all[kSfMaxK][kSfM];
  __shared__ int klist[kSfMaxK];
  __shared__ unsigned kmask_s;

  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t row0 = (int64_t)blockIdx.x * kSfM;
  if (tid == 0) kmask_s = 0u;
  __syncthreads();
  {
    unsigned mine = 0u;
    for (int e = tid; e < K * kSfM; e += kSfThreads) {
      const int k = e / kSfM, r = e - k * kSfM;
      const int64_t row = row0 + r;
      int v = -1;
      if (row < n_q) v = nbr ? __ldg(&nbr[(int64_t)k * n_q + row]) : (int)row;
      idx_all[k][r] = v;
      if (v >= 0) mine |= 1u << k;
    }
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) mine |= __shfl_xor_sync(0xffffffffu, mine, sft);
    if (lane == 0 && mine) atomicOr(&kmask_s, mine);
  }
  __syncthreads();
  int n_off = 0;
  {
    const unsigned km = kmask_s;
    for (int k = 0; k < K; ++k) {
      const int tk = mirror ? (K - 1 - k) : k;
      if (km & (1u << tk)) {
        if (tid == 0) klist[n_off] = k;
        ++n_off;
      }
    }
  }
  __syncthreads();
  const int n_slices = (c_in + kSfCS - 1) / kSfCS;
  const int total = n_off * n_slices;

  auto issue = [&](int step) {
    if (step < total) {
      const int o = step / n_slices, sl = step - o * n_slices;
      const int k = klist[o];
      const int tk = mirror ? (K - 1 - k) : k;
      const int buf = step % kSfStages;
      const int c0 = sl * kSfCS;
      const int cw = (c_in - c0 < kSfCS) ? (c_in - c0) : kSfCS;   // multiple of 4
      const int xq = cw / 4;
      for (int e = tid; e < kSfM * xq; e += kSfThreads) {
        const int r = e / xq, q = e - r * xq;
        const int src = idx_all[tk][r];
        dw_cp_async16(&xs[buf][r][4 * q], X + (size_t)(src < 0 ? 0 : src) * c_in + c0 + 4 * q, src >= 0);
      }
      if (!w_transposed) {   // W [K][c_in][COUT]: rows of the slice are contiguous
        const float* wk = W + ((size_t)k * c_in + c0) * COUT;
        for (int e = tid; e < cw * (COUT / 4); e += kSfThreads)
          dw_cp_async16(&ws[buf][0][0] + 4 * e, wk + 4 * e, true);
      } else {               // W [K][COUT][c_in]: transpose on the way in (4-byte copies)
        const float* wk = W + (size_t)k * COUT * c_in + c0;
        for (int e = tid; e < cw * COUT; e += kSfThreads) {
          const int co = e / cw, ci = e - co * cw;
          const uint32_t d = (uint32_t)__cvta_generic_to_shared(&ws[buf][ci][co]);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(wk + (size_t)co * c_in + ci) : "memory");
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  const int r = tid >> 1, ch = (tid & 1) * CPT;
  float acc[CPT];
#pragma unroll
  for (int j = 0; j < CPT; ++j) acc[j] = 0.f;

  issue(0);
  issue(1);
  for (int step = 0; step < total; ++step) {
    issue(step + 2);
    asm volatile("cp.async.wait_group 2;" ::: "memory");
    __syncthreads();
    const int buf = step % kSfStages;
    const int sl = step % n_slices;
    const int cw = (c_in - sl * kSfCS < kSfCS) ? (c_in - sl * kSfCS) : kSfCS;
    for (int c = 0; c < cw; c += 4) {
      const float4 xv = *(const float4*)&xs[buf][r][c];
      const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
#pragma unroll
        for (int j = 0; j < CPT / 4; ++j) {
          const float4 w = *(const float4*)&ws[buf][c + t][ch + 4 * j];
          acc[4 * j + 0] = fmaf(xa[t], w.x, acc[4 * j + 0]);
          acc[4 * j + 1] = fmaf(xa[t], w.y, acc[4 * j + 1]);
          acc[4 * j + 2] = fmaf(xa[t], w.z, acc[4 * j + 2]);
          acc[4 * j + 3] = fmaf(xa[t], w.w, acc[4 * j + 3]);
        }
      }
    }
    __syncthreads();
  }
  const int64_t row = row0 + r;
  if (row < n_q) {
    float4* y = (float4*)(Y + (size_t)row * COUT + ch);
#pragma unroll
    for (int j = 0; j < CPT / 4; ++j) y[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
  }
}

// ------------------------------------------------------------------------------------------
// weight gradient:  dW[k][ci][co] += sum over the pairs (in -> out) of offset k of X[in][ci] * dY[out][co]
//
// One CTA owns a run of kDwChunk consecutive pairs of one offset and a TCI x TCO tile of dW[k].  The pair rows
// (the tile's slice of X[in] and dY[out]) stream through a 3-stage cp.async ring in shared memory, so the bytes
// in flight do not cost registers (the register-staged predecessor sat at 17 % warp occupancy stalled on the
// long scoreboard, ncu profiles/r1_dw_warp_ncu.txt; the 64 x 64 tile kernel before it left 15/16 of its threads
// idle at 16 channels: 570 us per launch whatever the shape).  Compute: lane <-> input channel, registers <-> the
// TCO output channels; a warp takes every 8th pair of a stage, with 32 / TCI pairs side by side in its sub-groups.
// Partial tiles are summed across the 8 warps in shared memory and flushed with ONE atomicAdd per element and CTA.
// Pairs are sorted by output row inside an offset, so the dY reads stream.
// ------------------------------------------------------------------------------------------
constexpr int kDwThreads = 256;
constexpr int kDwChunk = 1024;   // pairs per CTA
constexpr int kDwSP = 64;        // pairs per pipeline stage
constexpr int kDwStages = 3;


template <int TCI, int TCO>
__global__ void __launch_bounds__(kDwThreads) conv_dw_kernel(
    const float* __restrict__ X, const float* __restrict__ dY, const int32_t* __restrict__ in_idx,
    const int32_t* __restrict__ out_idx, const int32_t* __restrict__ offs, int64_t n_identity, int K, int mirror,
    int c_in, int c_out, int n_co_tiles, float* __restrict__ dW) {
  constexpr int G = 32 / TCI;            // pairs side by side in one warp
  constexpr int XQ = TCI / 4;            // 16-byte chunks per staged X row slice
  constexpr int DQ = TCO / 4;
  // dynamic shared memory: [s_in][s_out][xs ring][ds ring]; the reduction buffer aliases the rings afterwards
  extern __shared__ __align__(16) uint8_t dw_smem[];
  int* s_in = (int*)dw_smem;
  int* s_out = s_in + kDwChunk;
  float (*xs)[kDwSP][TCI] = (float (*)[kDwSP][TCI])(s_out + kDwChunk);
  float (*ds)[kDwSP][TCO] = (float (*)[kDwSP][TCO])(&xs[kDwStages][0][0]);
  float (*red)[TCI][TCO + 1] = (float (*)[TCI][TCO + 1])(&xs[0][0][0]);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane / TCI, ci_l = lane % TCI;
  const int ci0 = (blockIdx.y / n_co_tiles) * TCI, co0 = (blockIdx.y % n_co_tiles) * TCO;
  const int tk = blockIdx.z;
  const int64_t p_begin = offs ? offs[tk] : 0, p_end = offs ? offs[tk + 1] : n_identity;
  const int64_t start = p_begin + (int64_t)blockIdx.x * kDwChunk;
  if (start >= p_end) return;
  const int n_pairs = (int)((p_end - start < kDwChunk) ? (p_end - start) : kDwChunk);

  for (int i = tid; i < n_pairs; i += kDwThreads) {
    s_in[i] = in_idx ? __ldg(&in_idx[start + i]) : (int)(start + i);
    s_out[i] = out_idx ? __ldg(&out_idx[start + i]) : (int)(start + i);
  }
  __syncthreads();

  const int n_stages = (n_pairs + kDwSP - 1) / kDwSP;
  auto issue = [&](int st) {
    if (st < n_stages) {
      const int buf = st % kDwStages, p0 = st * kDwSP;
      for (int e = tid; e < kDwSP * XQ; e += kDwThreads) {
        const int pp = e / XQ, q = e % XQ;
        const bool ok = p0 + pp < n_pairs && ci0 + 4 * q < c_in;
        const int64_t row = ok ? s_in[p0 + pp] : 0;
        dw_cp_async16(&xs[buf][pp][4 * q], X + row * c_in + ci0 + 4 * q, ok);
      }
      for (int e = tid; e < kDwSP * DQ; e += kDwThreads) {
        const int pp = e / DQ, q = e % DQ;
        const bool ok = p0 + pp < n_pairs && co0 + 4 * q < c_out;
        const int64_t row = ok ? s_out[p0 + pp] : 0;
        dw_cp_async16(&ds[buf][pp][4 * q], dY + row * c_out + co0 + 4 * q, ok);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  float acc[TCO];
#pragma unroll
  for (int j = 0; j < TCO; ++j) acc[j] = 0.f;

  issue(0);
  issue(1);
  for (int st = 0; st < n_stages; ++st) {
    issue(st + 2);
    asm volatile("cp.async.wait_group 2;" ::: "memory");
    __syncthreads();
    const int buf = st % kDwStages;
    // warp w takes pair slots w*G + sub, (w + 8)*G + sub, ... of the stage (zero-filled slots add nothing)
#pragma unroll
    for (int t = 0; t < kDwSP / (8 * G); ++t) {
      const int pp = (t * 8 + warp) * G + sub;
      const float xv = xs[buf][pp][ci_l];
#pragma unroll
      for (int j = 0; j < DQ; ++j) {
        const float4 d = *(const float4*)&ds[buf][pp][4 * j];
        acc[4 * j + 0] = fmaf(xv, d.x, acc[4 * j + 0]);
        acc[4 * j + 1] = fmaf(xv, d.y, acc[4 * j + 1]);
        acc[4 * j + 2] = fmaf(xv, d.z, acc[4 * j + 2]);
        acc[4 * j + 3] = fmaf(xv, d.w, acc[4 * j + 3]);
      }
    }
    __syncthreads();   // the ring slot is refilled by the issue() of the next iteration
  }
  // fold the sub-groups (same ci, different pairs), then the 8 warps, then one atomicAdd per element
#pragma unroll
  for (int m = TCI; m < 32; m <<= 1)
#pragma unroll
    for (int j = 0; j < TCO; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], m);
  if (sub == 0) {
#pragma unroll
    for (int j = 0; j < TCO; ++j) red[warp][ci_l][j] = acc[j];
  }
  __syncthreads();
  float* dWk = dW + (size_t)(mirror ? (K - 1 - tk) : tk) * c_in * c_out;
  for (int e = tid; e < TCI * TCO; e += kDwThreads) {
    const int ci = e / TCO, j = e % TCO;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kDwThreads / 32; ++w) v += red[w][ci][j];
    if (ci0 + ci < c_in && co0 + j < c_out) atomicAdd(dWk + (size_t)(ci0 + ci) * c_out + co0 + j, v);
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
  if ((c_out == 16 || c_out == 32) && c_in % 4 == 0 && K <= kSfMaxK) {
    const unsigned g = (unsigned)((n_q + kSfM - 1) / kSfM);
    const size_t sm = (size_t)kSfStages * (kSfM * kSfXS + kSfCS * c_out) * sizeof(float);
    static bool attr = false;
    if (!attr) {
      PGS_CUDA(cudaFuncSetAttribute(conv_small_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      PGS_CUDA(cudaFuncSetAttribute(conv_small_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      attr = true;
    }
    if (c_out == 16)
      conv_small_kernel<16><<<g, kSfThreads, sm, s>>>(X, W, nbr, n_q, K, c_in, mirror, w_transposed, Y);
    else
      conv_small_kernel<32><<<g, kSfThreads, sm, s>>>(X, W, nbr, n_q, K, c_in, mirror, w_transposed, Y);
    count_launch();
    PGS_CHECK_LAUNCH();
    return PGS_OK;
  }
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
  PGS_CHECK_ARG(c_in % 4 == 0 && c_out % 4 == 0, "channel counts must be multiples of 4 (16-byte row chunks)");
  static int use_mma = -1;
  if (use_mma < 0) {
    const char* e = getenv("PGS_DW_IMPL");   // "ffma": always the fp32 FFMA kernel below
    use_mma = !(e && e[0] == 'f');
  }
  if (use_mma && pgs_conv_dw_mma_supported(c_in, c_out))   // tensor-core version (conv_dw_mma.cu)
    return pgs_conv_bwd_weight_mma(X, dY, in_idx, out_idx, offs, max_pairs, K, c_in, c_out, mirror, dW, stream);
  const int tci = (c_in % 32 == 0) ? 32 : (c_in % 16 == 0 ? 16 : 4);
  const int tco = (c_out % 32 == 0) ? 32 : 16;
  const int n_ci = (c_in + tci - 1) / tci, n_co = (c_out + tco - 1) / tco;
  const unsigned gx = (unsigned)((max_pairs + kDwChunk - 1) / kDwChunk);
  const dim3 grid(gx, n_ci * n_co, K);
#define PGS_DW(TCI, TCO)                                                                                          \
  do {                                                                                                            \
    const size_t ring = (size_t)kDwStages * kDwSP * (TCI + TCO) * 4, redb = (size_t)8 * TCI * (TCO + 1) * 4;       \
    const size_t sm = (size_t)2 * kDwChunk * 4 + (ring > redb ? ring : redb);                                     \
    static bool attr = false;                                                                                     \
    if (!attr) {                                                                                                  \
      PGS_CUDA(cudaFuncSetAttribute(conv_dw_kernel<TCI, TCO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); \
      attr = true;                                                                                                \
    }                                                                                                             \
    conv_dw_kernel<TCI, TCO><<<grid, kDwThreads, sm, s>>>(X, dY, in_idx, out_idx, offs, max_pairs, K, mirror, c_in, c_out, \
                                                          n_co, dW);                                              \
  } while (0)
  if (tci == 32 && tco == 32) PGS_DW(32, 32);
  else if (tci == 32) PGS_DW(32, 16);
  else if (tci == 16 && tco == 32) PGS_DW(16, 32);
  else if (tci == 16) PGS_DW(16, 16);
  else if (tco == 32) PGS_DW(4, 32);
  else PGS_DW(4, 16);
#undef PGS_DW
  count_launch();
  PGS_CHECK_LAUNCH();
  return PGS_OK;
}

}  // extern "C"
