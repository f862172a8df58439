// Weight gradient of the sparse convolution on the tensor cores (register-operand mma.sync tf32).
//
// Reference call site: the backward of MinkowskiConvolution / MinkowskiConvolutionTranspose
// (torch_points3d/modules/MinkowskiEngine/api_modules.py:26-55,244-270,293 -> ME ConvolutionBackwardKernelGPU):
//     dW[k][ci][co] = sum over the pairs (i, o) of kernel offset k of  X[i][ci] * dY[o][co]
// i.e. per offset a GEMM  X_k^T [Cin x P_k] * dY_k [P_k x Cout]  whose contraction runs over the PAIRS.  With
// m16n8k8 fragments (lane = 4 g + t):  A[m = ci][k = pair]: a0 = X[in(t)][g], a1 = X[in(t)][g+8], a2 = X[in(t+4)][g],
// a3 = X[in(t+4)][g+8];  B[k = pair][n = co]: b0 = dY[out(t)][g], b1 = dY[out(t+4)][g] -- every 4-byte load
// instruction of a warp touches 4 pair rows x 32 contiguous bytes (full sectors), nothing is staged in shared memory
// and there is no barrier in the main loop.  The fp32 FFMA kernel this replaces (conv_dw_kernel, conv.cu) spent
// 64 FFMA + ~40 shared-memory loads per 8 pairs at 16 x 16; this one 6 HMMA + 8 loads + 24 integer ops.
//
// Precision: the same 3-product tf32 split as the forward kernels (hi*hi in one accumulator, lo*hi + hi*lo in a
// second one); a warp accumulates at most 128 pairs before its partial tile is summed with fp32 adds.
#include <cstdlib>

#include "common.cuh"

namespace pgs {

constexpr int kDmThreads = 256;
constexpr int kDmChunk = 1024;   // pairs per CTA (128 per warp)

__device__ __forceinline__ void dm_split(float v, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(v) + 0x00001000u) & 0xFFFFE000u;   // round-half-up to tf32; the tensor core truncates lo
  lo = __float_as_uint(v - __uint_as_float(hi));
}
__device__ __forceinline__ void dm_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// correction products lo*hi + hi*lo in ONE bf16 m16n8k16 MMA (see mma_corr in conv_mma.cu): contraction slots 2t / 2t+1
// hold (x_lo, d_hi) of the lane's pairs t / t+4, slots 2t+8 / 2t+9 hold (x_hi, d_lo) of the same pairs
__device__ __forceinline__ uint32_t dm_pack(uint32_t first_bits, uint32_t second_bits) {   // first -> low half
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(second_bits)), "f"(__uint_as_float(first_bits)));
  return d;
}
__device__ __forceinline__ void dm_corr(float (&d)[4], const uint32_t (&ahi)[4], const uint32_t (&alo)[4],
                                        const uint32_t (&bhi)[2], const uint32_t (&blo)[2]) {
  const uint32_t a0 = dm_pack(alo[0], alo[2]), a1 = dm_pack(alo[1], alo[3]);
  const uint32_t a2 = dm_pack(ahi[0], ahi[2]), a3 = dm_pack(ahi[1], ahi[3]);
  const uint32_t b0 = dm_pack(bhi[0], bhi[1]), b1 = dm_pack(blo[0], blo[1]);
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Row pieces of a pair arrive as ONE vector load per lane: lane (g, t) reads the 2 MT (A) / NT (B) consecutive
// floats  X[in(p)][ci0 + 2 MT g ..]  /  dY[out(p)][co0 + NT g ..]  of its pairs p = t and t + 4, so the 8 lanes of a
// pair cover one contiguous 64 MT / 32 NT-byte piece of the row and a warp-level load touches 4 rows (4 L1 wavefronts
// moving 256 .. 512 bytes) instead of one 4-byte load per fragment register (4 rows x 32 bytes per instruction, 8
// instructions per 8 pairs at 16 x 16 and 16 at 32 x 32).  A quarter of the load instructions, the same time
// (146 vs 150 us at 16 x 16 x 5.3 M pairs): the sectors, not the instructions, are the limit (see the kernel's
// header).  The tile's channels are therefore permuted inside the MMA (dW is written back through the same
// permutation):
//     A m-tile m, fragment row g / g + 8   <->  input channel  ci0 + 2 MT g + 2 m / + 2 m + 1
//     B n-tile n, fragment column g        <->  output channel co0 + NT g + n
template <int NF>
__device__ __forceinline__ void dm_load(const float* __restrict__ p, bool ok, float (&v)[NF]) {
  if constexpr (NF % 4 == 0) {
#pragma unroll
    for (int i = 0; i < NF / 4; ++i) {
      const float4 q = ok ? __ldg((const float4*)p + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
    }
  } else {
    static_assert(NF % 2 == 0, "pieces are whole float2s");
#pragma unroll
    for (int i = 0; i < NF / 2; ++i) {
      const float2 q = ok ? __ldg((const float2*)p + i) : make_float2(0.f, 0.f);
      v[2 * i] = q.x; v[2 * i + 1] = q.y;
    }
  }
}

// MT m-tiles of 16 input channels, NT n-tiles of 8 output channels per CTA tile; grid: x = chunk of kDmChunk pairs of
// one kernel offset, y = channel tile, z = kernel offset.
// (Tried and dropped, profiles/r2_dw_variants.json: walking the output rows in blocks of 4096 with the kernel offset
// as the fastest grid index, so that a block's X / dY rows stay in L2 across the 27 offsets -- 0 .. 20 % SLOWER.  The
// kernel is not waiting for HBM: it moves ~16 bytes per clock and SM of scattered 32-byte sectors from L2, 4.6 TB/s
// over the chip, whatever the order, the load width or the math pipe -- the FFMA kernel takes the same time at
// 16 x 16 -- which is the L2 -> SM bandwidth this access pattern gets.)
template <int MT, int NT>
__global__ void __launch_bounds__(kDmThreads) conv_dw_mma_kernel(
    const float* __restrict__ X, const float* __restrict__ dY, const int32_t* __restrict__ in_idx,
    const int32_t* __restrict__ out_idx, const int32_t* __restrict__ offs, int64_t n_identity, int K, int mirror,
    int c_in, int c_out, int n_co_tiles, float* __restrict__ dW) {
  constexpr int TCI = 16 * MT, TCO = 8 * NT, AF = 2 * MT;
  __shared__ int s_in[kDmChunk], s_out[kDmChunk];
  __shared__ float red[kDmThreads / 32][TCI][TCO + 1];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int ci0 = (blockIdx.y / n_co_tiles) * TCI, co0 = (blockIdx.y % n_co_tiles) * TCO;
  const int tk = blockIdx.z;
  const int64_t p_begin = offs ? offs[tk] : 0, p_end = offs ? offs[tk + 1] : n_identity;
  const int64_t start = p_begin + (int64_t)blockIdx.x * kDmChunk;
  if (start >= p_end) return;
  const int n_pairs = (int)((p_end - start < kDmChunk) ? (p_end - start) : kDmChunk);
  for (int i = tid; i < kDmChunk; i += kDmThreads) {
    const bool ok = i < n_pairs;
    s_in[i] = ok ? (in_idx ? __ldg(&in_idx[start + i]) : (int)(start + i)) : -1;
    s_out[i] = ok ? (out_idx ? __ldg(&out_idx[start + i]) : (int)(start + i)) : -1;
  }
  __syncthreads();

  float accm[MT][NT][4], accc[MT][NT][4];
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
      for (int c = 0; c < 4; ++c) accm[m][n][c] = accc[m][n][c] = 0.f;

  // warp w owns pairs [128 w, 128 w + 128) of the chunk: 16 contraction steps of 8 pairs
  const int wbase = warp * (kDmChunk / (kDmThreads / 32));
  const float* Xc = X + ci0 + AF * g;
  const float* Dc = dY + co0 + NT * g;
  // U contraction steps are loaded together before their MMAs (latency: index -> row -> mma)
  constexpr int U = (MT * NT <= 2) ? 4 : 2;
  constexpr int STEPS = kDmChunk / (kDmThreads / 32) / 8;
  for (int step = 0; step < STEPS; step += U) {
    if (wbase + 8 * step >= n_pairs) break;
    float av[U][2][AF], bv[U][2][NT];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = wbase + 8 * (step + u);   // < kDmChunk: entries past n_pairs hold -1
      const int i0 = s_in[p + t], i1 = s_in[p + t + 4], o0 = s_out[p + t], o1 = s_out[p + t + 4];
      dm_load<AF>(Xc + (size_t)(i0 >= 0 ? i0 : 0) * c_in, i0 >= 0, av[u][0]);
      dm_load<AF>(Xc + (size_t)(i1 >= 0 ? i1 : 0) * c_in, i1 >= 0, av[u][1]);
      dm_load<NT>(Dc + (size_t)(o0 >= 0 ? o0 : 0) * c_out, o0 >= 0, bv[u][0]);
      dm_load<NT>(Dc + (size_t)(o1 >= 0 ? o1 : 0) * c_out, o1 >= 0, bv[u][1]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      uint32_t ahi[MT][4], alo[MT][4], bhi[NT][2], blo[NT][2];
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        dm_split(av[u][0][2 * m], ahi[m][0], alo[m][0]);
        dm_split(av[u][0][2 * m + 1], ahi[m][1], alo[m][1]);
        dm_split(av[u][1][2 * m], ahi[m][2], alo[m][2]);
        dm_split(av[u][1][2 * m + 1], ahi[m][3], alo[m][3]);
      }
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        dm_split(bv[u][0][n], bhi[n][0], blo[n][0]);
        dm_split(bv[u][1][n], bhi[n][1], blo[n][1]);
      }
#pragma unroll
      for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int n = 0; n < NT; ++n) {
          dm_mma(accm[m][n], ahi[m], bhi[n][0], bhi[n][1]);
          dm_corr(accc[m][n], ahi[m], alo[m], bhi[n], blo[n]);
        }
    }
  }
  // C fragment: c0 = C[g][2t], c1 = C[g][2t+1], c2 = C[g+8][2t], c3 = C[g+8][2t+1]  (tile channels permuted, see above)
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      const int r0 = AF * g + 2 * m, q0 = NT * (2 * t) + n, q1 = NT * (2 * t + 1) + n;
      red[warp][r0][q0] = accm[m][n][0] + accc[m][n][0];
      red[warp][r0][q1] = accm[m][n][1] + accc[m][n][1];
      red[warp][r0 + 1][q0] = accm[m][n][2] + accc[m][n][2];
      red[warp][r0 + 1][q1] = accm[m][n][3] + accc[m][n][3];
    }
  __syncthreads();
  float* dWk = dW + (size_t)(mirror ? (K - 1 - tk) : tk) * c_in * c_out;
  for (int e = tid; e < TCI * TCO; e += kDmThreads) {
    const int ci = e / TCO, j = e % TCO;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kDmThreads / 32; ++w) v += red[w][ci][j];
    atomicAdd(dWk + (size_t)(ci0 + ci) * c_out + co0 + j, v);
  }
}

}  // namespace pgs

using namespace pgs;

extern "C" {

int pgs_conv_dw_mma_supported(int32_t c_in, int32_t c_out) { return c_in % 16 == 0 && c_out % 16 == 0 && c_in >= 16 && c_out >= 16; }

int pgs_conv_bwd_weight_mma(const float* X, const float* dY, const int32_t* in_idx, const int32_t* out_idx,
                            const int32_t* offs, int64_t max_pairs, int32_t K, int32_t c_in, int32_t c_out,
                            int32_t mirror, float* dW, void* stream) {
  PGS_CHECK_ARG(pgs_conv_dw_mma_supported(c_in, c_out), "channel counts must be multiples of 16");
  PGS_CHECK_ARG(offs != nullptr || K == 1, "offs == NULL requires K == 1 (identity pairs)");
  if (max_pairs <= 0) return PGS_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int nt = (c_out % 32 == 0) ? 4 : 2;
  const int mt = (c_in % 32 == 0) ? 2 : ((c_in % 48 == 0 && nt == 2) ? 3 : 1);   // 48 x 48: one X tile instead of three
  const int n_ci = c_in / (16 * mt), n_co = c_out / (8 * nt);
  const unsigned gx = (unsigned)((max_pairs + kDmChunk - 1) / kDmChunk);
  const dim3 grid(gx, n_ci * n_co, K);
  if (mt == 3)
    conv_dw_mma_kernel<3, 2><<<grid, kDmThreads, 0, s>>>(X, dY, in_idx, out_idx, offs, max_pairs, K, mirror, c_in, c_out, n_co, dW);
  else if (mt == 2 && nt == 4)
    conv_dw_mma_kernel<2, 4><<<grid, kDmThreads, 0, s>>>(X, dY, in_idx, out_idx, offs, max_pairs, K, mirror, c_in, c_out, n_co, dW);
  else if (mt == 2)
    conv_dw_mma_kernel<2, 2><<<grid, kDmThreads, 0, s>>>(X, dY, in_idx, out_idx, offs, max_pairs, K, mirror, c_in, c_out, n_co, dW);
  else if (nt == 4)
    conv_dw_mma_kernel<1, 4><<<grid, kDmThreads, 0, s>>>(X, dY, in_idx, out_idx, offs, max_pairs, K, mirror, c_in, c_out, n_co, dW);
  else
    conv_dw_mma_kernel<1, 2><<<grid, kDmThreads, 0, s>>>(X, dY, in_idx, out_idx, offs, max_pairs, K, mirror, c_in, c_out, n_co, dW);
  count_launch();
  PGS_CHECK_LAUNCH();
  return PGS_OK;
}

}  // extern "C"
