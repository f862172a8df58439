// Weight re-arrangement shared by the conv kernels and the batched prep entry point.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pgs {

constexpr int kPrepKC = 16;   // == kTcKC (conv_tc.cu): input channels per tcgen05 pipeline step

__device__ __forceinline__ uint32_t prep_tf32_rn(float x) {
  const uint32_t u = __float_as_uint(x);
  return (u + 0x00000FFFu + ((u >> 13) & 1u)) & 0xFFFFE000u;
}

__device__ __forceinline__ uint32_t prep_pack_bf16(uint32_t first_bits, uint32_t second_bits);

// tcgen05 layout:  Wp[k][j][plane][q][n][4] = split_plane(B_k[n][j*16 + q*4 .. +3]); e enumerates the 2*K*C*N OUTPUT
// floats.  plane 0: hi = rn_tf32(w).  plane 1, corr16 == 0: lo = rn_tf32(w - hi) (w - hi is exact in fp32);
// corr16 != 0: the B operand of the bf16 correction MMA -- the 16 bytes of (q, n) hold 8 bf16 contraction slots
// [hi(4q) hi(4q+1) hi(4q+2) hi(4q+3) lo(4q) .. lo(4q+3)], which meet [x_lo .. | x_hi ..] of the same channels in A.
// Splitting here, once per step, instead of in the conv kernel at every pipeline step lets the kernel fetch a chunk
// with one TMA copy.
__device__ __forceinline__ float prep_tc_elem(const float* __restrict__ W, int K, int C, int N, int w_transposed, int corr16,
                                              int64_t e) {
  const int t = (int)(e & 3);
  int64_t r = e >> 2;
  const int n = (int)(r % N);
  r /= N;
  const int q = (int)(r & 3);
  r >>= 2;
  const int plane = (int)(r & 1);
  r >>= 1;
  const int J = C / kPrepKC;
  const int j = (int)(r % J);
  const int k = (int)(r / J);
  // !w_transposed: W stored [K][C][N];  w_transposed: W stored [K][N][C]
  auto weight = [&](int c) { return w_transposed ? W[((int64_t)k * N + n) * C + c] : W[((int64_t)k * C + c) * N + n]; };
  if (plane == 1 && corr16) {   // float slot t = bf16 pair: t < 2 -> hi of channels 2t, 2t+1; t >= 2 -> lo of 2(t-2), 2(t-2)+1
    const int c0 = j * kPrepKC + q * 4 + 2 * (t & 1);
    const float w0 = weight(c0), w1 = weight(c0 + 1);
    const uint32_t h0 = prep_tf32_rn(w0), h1 = prep_tf32_rn(w1);
    if (t < 2) return __uint_as_float(prep_pack_bf16(h0, h1));
    return __uint_as_float(prep_pack_bf16(__float_as_uint(w0 - __uint_as_float(h0)), __float_as_uint(w1 - __uint_as_float(h1))));
  }
  const float w = weight(j * kPrepKC + q * 4 + t);
  const float hi = __uint_as_float(prep_tf32_rn(w));
  return plane == 0 ? hi : __uint_as_float(prep_tf32_rn(w - hi));
}
__device__ __forceinline__ uint32_t prep_pack_bf16(uint32_t first_bits, uint32_t second_bits) {   // first -> low half
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(second_bits)), "f"(__uint_as_float(first_bits)));
  return d;
}
// mma.sync fragment layout (conv_mma.cu): e enumerates float4 fragments [k][j][n][lane]
__device__ __forceinline__ float4 prep_mma_frag(const float* __restrict__ W, int K, int C, int N, int w_transposed, int64_t e) {
  const int J = C / 8, NT = N / 8;
  const int lane = (int)(e & 31);
  int64_t r = e >> 5;
  const int n = (int)(r % NT);
  r /= NT;
  const int j = (int)(r % J);
  const int k = (int)(r / J);
  const int g = lane >> 2, t = lane & 3;
  const int co = 16 * (n >> 1) + 4 * (g >> 1) + 2 * (n & 1) + (g & 1);
  float v[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {   // position p = t + 4 h
    const int ci = 16 * (j >> 1) + 4 * t + 2 * (j & 1) + h;
    v[h] = w_transposed ? W[((int64_t)k * N + co) * C + ci] : W[((int64_t)k * C + ci) * N + co];
  }
  const uint32_t h0 = prep_tf32_rn(v[0]), h1 = prep_tf32_rn(v[1]);
  const uint32_t l0 = __float_as_uint(v[0] - __uint_as_float(h0)), l1 = __float_as_uint(v[1] - __uint_as_float(h1));
  // (b0_hi, b1_hi) tf32 for the main MMA; bf16 pairs (hi, hi) and (lo, lo) = B fragment of the correction MMA
  return make_float4(__uint_as_float(h0), __uint_as_float(h1), __uint_as_float(prep_pack_bf16(h0, h1)),
                     __uint_as_float(prep_pack_bf16(l0, l1)));
}

}  // namespace pgs
