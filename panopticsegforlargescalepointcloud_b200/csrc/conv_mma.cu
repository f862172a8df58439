// Sparse convolution forward / backward-input for the narrow layers (<= 64 channels, many rows): warp-level
// register-operand tensor-core kernel.
//
// Same contract as pgs_conv_fwd / pgs_conv_fwd_tc (reference call sites: torch_points3d/modules/MinkowskiEngine/
// api_modules.py:26-55,244-270,293 -> MinkowskiEngine ConvolutionForward/BackwardKernelGPU):
//     Y[q] = sum_k X[nbr[tk(k)][q]] * W[k]            (or W[k]^T for the input gradient)
//
// Why a second tensor-core kernel: the tcgen05 path (conv_tc.cu) stages the gathered rows in shared memory and
// issues M = 128 MMAs from one thread; at N = Cout <= 32 every such MMA is bound by its 4 KB A-operand fetch from
// shared memory (profiles/r1_mma_issue_rate.txt) and the CTA-wide barrier per 16-channel step, and the 16 -> 16
// layers on the 200 k-row level run at 145 us (5 % of the HBM roofline).  Here the gathered rows never touch shared
// memory: the 4 lanes that own a fragment row load 64 contiguous bytes of the feature row straight into the
// mma.sync A-fragment registers (the contraction index is permuted so that one 16-byte load feeds two K = 8
// steps; the weights are pre-arranged with the same permutation), so warps run independently, one barrier per
// kernel offset (for the shared weight stage) instead of one per 16 channels, and empty (16-row tile, offset)
// pairs are skipped.
//
// Precision: operands split a = hi + lo (hi = round-to-nearest tf32); hi*hi accumulates in one fp32 fragment (tf32
// MMA), lo*hi + hi*lo in a second one (one bf16 m16n8k16 MMA, see mma_corr), summed in the epilogue.
#include <cstdlib>

#include "common.cuh"
#include "conv_prep.cuh"

namespace pgs {

constexpr int kMmWarps = 4;   // 4-warp CTAs (64 / 128 rows): measured 2-13 % faster than 8 warps (finer barrier domains, better tail balance on the 25-30 k-row levels)
constexpr int kMmThreads = kMmWarps * 32;
constexpr int kMmMaxK = 27;

__device__ __forceinline__ uint32_t tf32_rn_bits(float x) {
  const uint32_t u = __float_as_uint(x);
  return (u + 0x00000FFFu + ((u >> 13) & 1u)) & 0xFFFFE000u;
}
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  hi = tf32_rn_bits(v);
  lo = tf32_rn_bits(v - __uint_as_float(hi));
}
// gathered-row split on the hot path: hi = round-half-up to tf32 (2 integer ops), lo = v - hi (exact in fp32); the
// tensor core ignores the low 13 mantissa bits of lo itself (|lo| <= 2^-11 |v|, so that truncation is <= 2^-21 |v|,
// the same order as the dropped lo*lo term; its sign follows lo, not v, so it does not bias the sum)
__device__ __forceinline__ void split_tf32_fast(float v, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(v) + 0x00001000u) & 0xFFFFE000u;
  lo = __float_as_uint(v - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// Correction products on the bf16 path.  lo*hi + hi*lo are 2^-12 of the main product, so 8 mantissa bits are enough
// for them (every rounding below contributes <= 2^-21 |x||w|, the size of the lo*lo term that is dropped anyway;
// numpy emulation: max error 3e-6 at output scale 4 vs 3e-7, next to ~4e-6 from the tensor core's own accumulation).
// ONE m16n8k16 bf16 MMA does both: contraction slots 0..7 hold (x_lo, w_hi), slots 8..15 hold (x_hi, w_lo) of the same
// 8 channels -- slot 2t / 2t+1 <-> the lane's own channels c(j,t) / c(j,t+4), so no data moves between lanes.
__device__ __forceinline__ uint32_t pack_bf16(uint32_t first_bits, uint32_t second_bits) {   // first -> low half
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(second_bits)), "f"(__uint_as_float(first_bits)));
  return d;
}
__device__ __forceinline__ void mma_corr(float (&d)[4], const uint32_t (&ahi)[4], const uint32_t (&alo)[4], uint32_t bh,
                                         uint32_t bl) {
  const uint32_t a0 = pack_bf16(alo[0], alo[2]), a1 = pack_bf16(alo[1], alo[3]);
  const uint32_t a2 = pack_bf16(ahi[0], ahi[2]), a3 = pack_bf16(ahi[1], ahi[3]);
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(bh), "r"(bl));
}
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// Fragment-order weights.  m16n8k8 tf32 fragments, lane = 4 g + t:
//   A: a0 = A[g][t], a1 = A[g+8][t], a2 = A[g][t+4], a3 = A[g+8][t+4]      B: b0 = B[t][g], b1 = B[t+4][g]
//   C: c0 = C[g][2t], c1 = C[g][2t+1], c2 = C[g+8][2t], c3 = C[g+8][2t+1]
// Contraction permutation (so that lane t's float4 #(t + 4i) of a feature row = channels 16i + 4t .. +3 feeds
// K-steps 2i and 2i+1):   K-step j, position p  <->  channel 16 (j/2) + 4 (p%4) + 2 (j%2) + p/4
// Output permutation (so that a lane's c0,c1 of n-tiles 2m and 2m+1 are 4 consecutive channels):
//   n-tile n, position q  <->  channel 16 (n/2) + 4 (q/2) + 2 (n%2) + q%2
// Wf[k][j][n][lane] = float4(b0_hi, b1_hi, bf16x2(b0_hi, b1_hi), bf16x2(b0_lo, b1_lo))
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv_mma_prep_weights_kernel(const float* __restrict__ W, int K, int C, int N,
                                                                     int w_transposed, float* __restrict__ Wf) {
  const int64_t total = (int64_t)K * (C / 8) * (N / 8) * 32;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
    ((float4*)Wf)[e] = prep_mma_frag(W, K, C, N, w_transposed, e);
}

// All weight re-arrangements of a pass in ONE launch: blockIdx.y = descriptor (8 x int64: W, dst, K, C, N,
// w_transposed, layout (0 = tcgen05, 1 = mma fragments), unused)
__global__ void __launch_bounds__(256) conv_prep_batch_kernel(const int64_t* __restrict__ desc, int tc_corr16) {
  const int64_t* d = desc + 8 * (int64_t)blockIdx.y;
  const float* W = (const float*)d[0];
  float* dst = (float*)d[1];
  const int K = (int)d[2], C = (int)d[3], N = (int)d[4], wt = (int)d[5], layout = (int)d[6];
  if (layout == 0) {
    const int64_t total = 2 * (int64_t)K * C * N;   // hi + lo planes (conv_prep.cuh)
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
      dst[e] = prep_tc_elem(W, K, C, N, wt, tc_corr16, e);
  } else {
    const int64_t total = (int64_t)K * (C / 8) * (N / 8) * 32;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
      ((float4*)dst)[e] = prep_mma_frag(W, K, C, N, wt, e);
  }
}

// WS = 0: weights of one kernel offset staged in shared memory (cp.async double buffer, one barrier per offset)
// WS = 1: weight fragments read straight from global memory (L1-resident for the small shapes), no barrier
template <int CIN, int COUT, int MT, int WS>
__global__ void __launch_bounds__(kMmThreads, (CIN * COUT * MT <= 32 * 32 ? 4 : 2)) conv_mma_kernel(const float* __restrict__ X, const float4* __restrict__ Wf,
                                                               const int32_t* __restrict__ nbr, int64_t n_q, int K,
                                                               int mirror, const int32_t* __restrict__ order,
                                                               float* __restrict__ Y) {
  constexpr int R = kMmWarps * 16 * MT;   // rows per CTA
  constexpr int J = CIN / 8, NT = COUT / 8, F4 = CIN / 16;
  constexpr int WSTAGE = J * NT * 32;     // float4 elements of one offset's fragments
  extern __shared__ __align__(16) uint8_t smem_raw[];
  int* idx = (int*)smem_raw;                                         // [K][R]
  float4* wst = (float4*)(smem_raw + (size_t)kMmMaxK * R * 4);       // [2][WSTAGE]  (WS == 0 only)
  __shared__ int klist[kMmMaxK];
  __shared__ unsigned kmask_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int64_t row0 = (int64_t)blockIdx.x * R;
  if (tid == 0) kmask_s = 0u;
  __syncthreads();
  {
    unsigned mine = 0u;
    for (int e = tid; e < K * R; e += kMmThreads) {
      const int k = e / R, r = e - k * R;
      const int64_t row = row0 + r;
      int v = -1;
      if (row < n_q) v = __ldg(&nbr[(int64_t)k * n_q + row]);
      idx[e] = v;
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
    for (int k = 0; k < K; ++k) {   // weight index order; table offset tk = mirror ? K-1-k : k
      const int tk = mirror ? (K - 1 - k) : k;
      if (km & (1u << tk)) {
        if (tid == 0) klist[n_off] = k;
        ++n_off;
      }
    }
  }
  __syncthreads();

  float accm[MT][NT][4], accc[MT][NT][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
      for (int c = 0; c < 4; ++c) accm[mt][n][c] = accc[mt][n][c] = 0.f;

  float4 bufA[MT][2][F4], bufB[MT][2][F4];   // ping-pong: rows of the current / the next kernel offset
  unsigned vA = 0u, vB = 0u;   // bit mt: some row of m-tile mt has a neighbour at this offset (warp-uniform)

  auto gather = [&](int o, float4 (&dst)[MT][2][F4], unsigned& valid) {
    const int k = klist[o];
    const int tk = mirror ? (K - 1 - k) : k;
    const int* col = idx + tk * R + warp * 16 * MT + g;
    valid = 0u;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int s0 = col[mt * 16], s1 = col[mt * 16 + 8];
#pragma unroll
      for (int i = 0; i < F4; ++i) {
        dst[mt][0][i] = (s0 >= 0) ? __ldg((const float4*)(X + (size_t)s0 * CIN) + t + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
        dst[mt][1][i] = (s1 >= 0) ? __ldg((const float4*)(X + (size_t)s1 * CIN) + t + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (__any_sync(0xffffffffu, (s0 >= 0) | (s1 >= 0))) valid |= 1u << mt;
    }
  };
  auto stage_w = [&](int o) {
    if (WS == 0) {
      const float4* src = Wf + (size_t)klist[o] * WSTAGE;
      float4* dst = wst + (o & 1) * WSTAGE;
      for (int e = tid; e < WSTAGE; e += kMmThreads) cp_async16(dst + e, src + e);
      cp_async_commit();
    }
  };
  auto compute = [&](int o, const float4 (&cur)[MT][2][F4], unsigned vcur) {
    if (vcur == 0u) return;
    const float4* wb = (WS == 0) ? (wst + (o & 1) * WSTAGE) : (Wf + (size_t)klist[o] * WSTAGE);
#pragma unroll
    for (int j = 0; j < J; ++j) {
      uint32_t ahi[MT][4], alo[MT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const float4 v0 = cur[mt][0][j >> 1], v1 = cur[mt][1][j >> 1];
        split_tf32_fast((j & 1) ? v0.z : v0.x, ahi[mt][0], alo[mt][0]);
        split_tf32_fast((j & 1) ? v1.z : v1.x, ahi[mt][1], alo[mt][1]);
        split_tf32_fast((j & 1) ? v0.w : v0.y, ahi[mt][2], alo[mt][2]);
        split_tf32_fast((j & 1) ? v1.w : v1.y, ahi[mt][3], alo[mt][3]);
      }
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        float4 b;
        if (WS == 0)
          b = wb[(j * NT + n) * 32 + lane];
        else
          b = __ldg(wb + (j * NT + n) * 32 + lane);
        const uint32_t bh0 = __float_as_uint(b.x), bh1 = __float_as_uint(b.y);
        const uint32_t bl0 = __float_as_uint(b.z), bl1 = __float_as_uint(b.w);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          if (MT > 1 && !(vcur & (1u << mt))) continue;
          mma_tf32(accm[mt][n], ahi[mt], bh0, bh1);
          mma_corr(accc[mt][n], ahi[mt], alo[mt], bl0, bl1);
        }
      }
    }
  };
  auto turn = [&](int o, const float4 (&cur)[MT][2][F4], unsigned vcur, float4 (&nxt)[MT][2][F4], unsigned& vnxt) {
    if (WS == 0) {
      cp_async_wait_all();
      __syncthreads();   // stage o has landed for everyone; everyone is done reading the other buffer
    }
    if (o + 1 < n_off) {
      stage_w(o + 1);
      gather(o + 1, nxt, vnxt);
    }
    compute(o, cur, vcur);
  };

  if (n_off > 0) {
    stage_w(0);
    gather(0, bufA, vA);
  }
  for (int o = 0; o < n_off; o += 2) {
    turn(o, bufA, vA, bufB, vB);
    if (o + 1 < n_off) turn(o + 1, bufB, vB, bufA, vA);
  }

  // epilogue: lane (g, t) owns channels 16 m + 4 t .. +3 of rows g and g + 8 of each m-tile
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int64_t t0 = row0 + warp * 16 * MT + mt * 16 + g, t1 = t0 + 8;   // rows of the (sorted) table
    const bool ok0 = t0 < n_q, ok1 = t1 < n_q;
    const int64_t r0 = (ok0 && order) ? (int64_t)__ldg(&order[t0]) : t0;   // output rows
    const int64_t r1 = (ok1 && order) ? (int64_t)__ldg(&order[t1]) : t1;
#pragma unroll
    for (int m = 0; m < NT / 2; ++m) {
      const int c = 16 * m + 4 * t;
      if (ok0)
        *(float4*)(Y + (size_t)r0 * COUT + c) =
            make_float4(accm[mt][2 * m][0] + accc[mt][2 * m][0], accm[mt][2 * m][1] + accc[mt][2 * m][1],
                        accm[mt][2 * m + 1][0] + accc[mt][2 * m + 1][0], accm[mt][2 * m + 1][1] + accc[mt][2 * m + 1][1]);
      if (ok1)
        *(float4*)(Y + (size_t)r1 * COUT + c) =
            make_float4(accm[mt][2 * m][2] + accc[mt][2 * m][2], accm[mt][2 * m][3] + accc[mt][2 * m][3],
                        accm[mt][2 * m + 1][2] + accc[mt][2 * m + 1][2], accm[mt][2 * m + 1][3] + accc[mt][2 * m + 1][3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Queued variant (used for the narrowest shapes, CIN * COUT <= 512, where it measured 20-30 % faster; from 32 x 32 on
// the register version is ahead by ~10 %): the gathered rows travel global -> shared memory with cp.async into a per-THREAD ring of
// D kernel offsets (each lane later reads back exactly the 16-byte pieces it copied, so no cross-lane
// synchronisation is needed for them), the weight fragments of an offset ride in the same cp.async group.  Compared
// with conv_mma_kernel (register double buffer, depth 1) this frees the prefetch registers -- 16 x 16 drops to one
// m-tile per warp at <= 64 registers, 4 CTAs = 32 warps per SM instead of 16 -- and keeps D - 1 offsets of loads in
// flight per thread: ncu of the register version showed 23 % of warp slots active and long-scoreboard (gather
// latency) as the top stall at 24 % tensor-pipe activity (profiles/r1_conv_mma_ncu.txt).
// Missing neighbours are zero-filled by cp.async itself (src-size 0).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16_zfill(void* dst_smem, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src),
               "r"(sz)
               : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int CIN, int COUT, int D, int MINB>
__global__ void __launch_bounds__(kMmThreads, MINB) conv_mmaq_kernel(const float* __restrict__ X,
                                                                      const float4* __restrict__ Wf,
                                                                      const int32_t* __restrict__ nbr, int64_t n_q, int K,
                                                                      int mirror, const int32_t* __restrict__ order,
                                                                      float* __restrict__ Y) {
  constexpr int R = kMmWarps * 16;        // rows per CTA (one 16-row m-tile per warp)
  constexpr int J = CIN / 8, NT = COUT / 8, F4 = CIN / 16;
  constexpr int WSTAGE = J * NT * 32;     // float4 elements of one offset's weight fragments
  constexpr int QSLOT = 2 * F4;           // float4 pieces per thread and offset (rows g and g + 8)
  extern __shared__ __align__(16) uint8_t smem_raw[];
  int* idx = (int*)smem_raw;                                                     // [K][R]
  float4* wst = (float4*)(smem_raw + (size_t)kMmMaxK * R * 4);                   // [D][WSTAGE]
  float4* que = wst + (size_t)D * WSTAGE;                                        // [D][QSLOT][kMmThreads]
  __shared__ int klist[kMmMaxK];
  __shared__ unsigned kmask_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int64_t row0 = (int64_t)blockIdx.x * R;
  if (tid == 0) kmask_s = 0u;
  __syncthreads();
  {
    unsigned mine = 0u;
    for (int e = tid; e < K * R; e += kMmThreads) {
      const int k = e / R, r = e - k * R;
      const int64_t row = row0 + r;
      int v = -1;
      if (row < n_q) v = __ldg(&nbr[(int64_t)k * n_q + row]);
      idx[e] = v;
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
    for (int k = 0; k < K; ++k) {   // weight index order; table offset tk = mirror ? K-1-k : k
      const int tk = mirror ? (K - 1 - k) : k;
      if (km & (1u << tk)) {
        if (tid == 0) klist[n_off] = k;
        ++n_off;
      }
    }
  }
  __syncthreads();

  float accm[NT][4], accc[NT][4];
#pragma unroll
  for (int n = 0; n < NT; ++n)
#pragma unroll
    for (int c = 0; c < 4; ++c) accm[n][c] = accc[n][c] = 0.f;

  // group G_o = rows + weights of the o-th non-empty offset (an empty group past the end keeps the counting uniform)
  auto issue = [&](int o) {
    if (o < n_off) {
      const int k = klist[o];
      const int tk = mirror ? (K - 1 - k) : k;
      const int st = o % D;
      const int* col = idx + tk * R + warp * 16 + g;
      const int s0 = col[0], s1 = col[8];
      float4* q = que + (size_t)st * QSLOT * kMmThreads + tid;
      const float4* x0 = (const float4*)(X + (size_t)(s0 >= 0 ? s0 : 0) * CIN) + t;
      const float4* x1 = (const float4*)(X + (size_t)(s1 >= 0 ? s1 : 0) * CIN) + t;
#pragma unroll
      for (int i = 0; i < F4; ++i) {
        cp_async16_zfill(q + (size_t)(2 * i) * kMmThreads, x0 + 4 * i, s0 >= 0);
        cp_async16_zfill(q + (size_t)(2 * i + 1) * kMmThreads, x1 + 4 * i, s1 >= 0);
      }
      const float4* wsrc = Wf + (size_t)k * WSTAGE;
      float4* wdst = wst + (size_t)st * WSTAGE;
      for (int e = tid; e < WSTAGE; e += kMmThreads) cp_async16(wdst + e, wsrc + e);
    }
    cp_async_commit();
  };

#pragma unroll
  for (int o = 0; o < D - 1; ++o) issue(o);
  for (int o = 0; o < n_off; ++o) {
    cp_async_wait<D - 2>();   // G_o has landed (at most the D - 2 younger groups are still in flight)
    __syncthreads();          // ... for every thread; and everyone is done with the slot G_{o+D-1} overwrites
    issue(o + D - 1);
    const int k = klist[o];
    const int tk = mirror ? (K - 1 - k) : k;
    const int* col = idx + tk * R + warp * 16 + g;
    if (!__any_sync(0xffffffffu, (col[0] >= 0) | (col[8] >= 0))) continue;   // this warp's 16 rows: no neighbour
    const int st = o % D;
    const float4* q = que + (size_t)st * QSLOT * kMmThreads + tid;
    const float4* wb = wst + (size_t)st * WSTAGE + lane;
#pragma unroll
    for (int i = 0; i < F4; ++i) {
      const float4 v0 = q[(size_t)(2 * i) * kMmThreads], v1 = q[(size_t)(2 * i + 1) * kMmThreads];
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        uint32_t ahi[4], alo[4];
        split_tf32_fast(jj ? v0.z : v0.x, ahi[0], alo[0]);
        split_tf32_fast(jj ? v1.z : v1.x, ahi[1], alo[1]);
        split_tf32_fast(jj ? v0.w : v0.y, ahi[2], alo[2]);
        split_tf32_fast(jj ? v1.w : v1.y, ahi[3], alo[3]);
#pragma unroll
        for (int n = 0; n < NT; ++n) {
          const float4 b = wb[((2 * i + jj) * NT + n) * 32];
          const uint32_t bh0 = __float_as_uint(b.x), bh1 = __float_as_uint(b.y);
          mma_tf32(accm[n], ahi, bh0, bh1);
          mma_corr(accc[n], ahi, alo, __float_as_uint(b.z), __float_as_uint(b.w));
        }
      }
    }
  }
  cp_async_wait<0>();

  // epilogue: lane (g, t) owns channels 16 m + 4 t .. +3 of rows g and g + 8
  const int64_t t0 = row0 + warp * 16 + g, t1 = t0 + 8;   // rows of the (sorted) table
  const bool ok0 = t0 < n_q, ok1 = t1 < n_q;
  const int64_t r0 = (ok0 && order) ? (int64_t)__ldg(&order[t0]) : t0;   // output rows
  const int64_t r1 = (ok1 && order) ? (int64_t)__ldg(&order[t1]) : t1;
#pragma unroll
  for (int m = 0; m < NT / 2; ++m) {
    const int c = 16 * m + 4 * t;
    if (ok0)
      *(float4*)(Y + (size_t)r0 * COUT + c) =
          make_float4(accm[2 * m][0] + accc[2 * m][0], accm[2 * m][1] + accc[2 * m][1],
                      accm[2 * m + 1][0] + accc[2 * m + 1][0], accm[2 * m + 1][1] + accc[2 * m + 1][1]);
    if (ok1)
      *(float4*)(Y + (size_t)r1 * COUT + c) =
          make_float4(accm[2 * m][2] + accc[2 * m][2], accm[2 * m][3] + accc[2 * m][3],
                      accm[2 * m + 1][2] + accc[2 * m + 1][2], accm[2 * m + 1][3] + accc[2 * m + 1][3]);
  }
}

template <int CIN, int COUT>
static int launch_mmaq(const float* X, const float* Wf, const int32_t* nbr, const int32_t* order, int64_t n_q, int K,
                       int mirror, float* Y, cudaStream_t s) {
  // ring depth and CTAs per SM from the shared memory of one ring stage (weight fragments + per-thread row pieces)
  constexpr int STAGE = CIN * COUT * 8 + CIN * 2 * kMmThreads;
  constexpr int D = (STAGE <= 8 * 1024) ? 4 : (STAGE <= 20 * 1024 ? 3 : 2);
  constexpr int FIT = (224 * 1024) / (kMmMaxK * kMmWarps * 16 * 4 + D * STAGE + 1024);
  constexpr int RCAP = COUT <= 16 ? 8 : (COUT <= 32 ? 6 : 4);   // 2 * COUT accumulator registers per thread, 128-thread CTAs
  constexpr int MINB = FIT < 1 ? 1 : (FIT > RCAP ? RCAP : FIT);
  constexpr int R = kMmWarps * 16;
  constexpr size_t smem = (size_t)kMmMaxK * R * 4 +
                          (size_t)D * ((size_t)(CIN / 8) * (COUT / 8) * 32 * 16 + (size_t)2 * (CIN / 16) * kMmThreads * 16);
  static bool attr_set = false;
  if (!attr_set) {
    PGS_CUDA(cudaFuncSetAttribute(conv_mmaq_kernel<CIN, COUT, D, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    PGS_CUDA(cudaFuncSetAttribute(conv_mmaq_kernel<CIN, COUT, D, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  cudaSharedmemCarveoutMaxShared));
    attr_set = true;
  }
  const unsigned gx = (unsigned)((n_q + R - 1) / R);
  conv_mmaq_kernel<CIN, COUT, D, MINB><<<gx, kMmThreads, smem, s>>>(X, (const float4*)Wf, nbr, n_q, K, mirror, order, Y);
  return PGS_OK;
}

// ---------------------------------------------------------------------------------------------
// Few-row layers (coarse U-Net levels: 50 .. a few thousand rows, 64 .. 192 channels).  These launches are latency
// bound, not throughput bound: the whole layer is a few MFLOP but the weights are up to 4 MB.  Work item of one
// WARP = (16-row tile, 16 output channels, one part of the kernel offsets); no shared memory, no barrier; the
// weight fragments come straight from L2 / L1 (items that share a weight slice are adjacent in launch order), the
// partial sums of the offset parts meet in Y by atomicAdd (Y zeroed by the caller when ksplit > 1).
// ---------------------------------------------------------------------------------------------
constexpr int kSplitWarps = 4;

__global__ void __launch_bounds__(kSplitWarps * 32) conv_mma_split_kernel(const float* __restrict__ X,
                                                                          const float4* __restrict__ Wf,
                                                                          const int32_t* __restrict__ nbr, int64_t n_q,
                                                                          int K, int c_in, int c_out, int mirror,
                                                                          int ksplit, int mtiles,
                                                                          const int32_t* __restrict__ order,
                                                                          float* __restrict__ Y) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int64_t item = (int64_t)blockIdx.x * kSplitWarps + (threadIdx.x >> 5);
  const int nchunks = c_out >> 4;
  if (item >= (int64_t)mtiles * nchunks * ksplit) return;
  const int mt = (int)(item % mtiles);
  const int nc = (int)((item / mtiles) % nchunks);
  const int kp = (int)(item / ((int64_t)mtiles * nchunks));
  const int k_begin = (int)(((long long)K * kp) / ksplit), k_end = (int)(((long long)K * (kp + 1)) / ksplit);
  const int J = c_in >> 3, NT = c_out >> 3, CC = c_in >> 4;
  const int64_t r0 = (int64_t)mt * 16 + g, r1 = r0 + 8;

  float accm[2][4], accc[2][4];
#pragma unroll
  for (int n = 0; n < 2; ++n)
#pragma unroll
    for (int c = 0; c < 4; ++c) accm[n][c] = accc[n][c] = 0.f;
  bool touched = false;

  for (int k = k_begin; k < k_end; ++k) {
    const int tk = mirror ? (K - 1 - k) : k;
    const int s0 = (r0 < n_q) ? __ldg(&nbr[(int64_t)tk * n_q + r0]) : -1;
    const int s1 = (r1 < n_q) ? __ldg(&nbr[(int64_t)tk * n_q + r1]) : -1;
    if (!__any_sync(0xffffffffu, (s0 >= 0) | (s1 >= 0))) continue;
    touched = true;
    const float4* x0 = (const float4*)(X + (size_t)(s0 >= 0 ? s0 : 0) * c_in) + t;
    const float4* x1 = (const float4*)(X + (size_t)(s1 >= 0 ? s1 : 0) * c_in) + t;
    const float4* wk = Wf + ((size_t)k * J * NT + 2 * nc) * 32 + lane;
#pragma unroll 2
    for (int cc = 0; cc < CC; ++cc) {
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 v0 = (s0 >= 0) ? __ldg(x0 + 4 * cc) : z;
      const float4 v1 = (s1 >= 0) ? __ldg(x1 + 4 * cc) : z;
      float4 b[2][2];
#pragma unroll
      for (int jj = 0; jj < 2; ++jj)
#pragma unroll
        for (int n = 0; n < 2; ++n) b[jj][n] = __ldg(wk + ((size_t)(2 * cc + jj) * NT + n) * 32);
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        uint32_t ahi[4], alo[4];
        split_tf32(jj ? v0.z : v0.x, ahi[0], alo[0]);
        split_tf32(jj ? v1.z : v1.x, ahi[1], alo[1]);
        split_tf32(jj ? v0.w : v0.y, ahi[2], alo[2]);
        split_tf32(jj ? v1.w : v1.y, ahi[3], alo[3]);
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          const uint32_t bh0 = __float_as_uint(b[jj][n].x), bh1 = __float_as_uint(b[jj][n].y);
          const uint32_t bl0 = __float_as_uint(b[jj][n].z), bl1 = __float_as_uint(b[jj][n].w);
          mma_tf32(accm[n], ahi, bh0, bh1);
          mma_corr(accc[n], ahi, alo, bl0, bl1);
        }
      }
    }
  }
  const int c = 16 * nc + 4 * t;
  const bool ok0 = r0 < n_q, ok1 = r1 < n_q;
  const int64_t y0 = (ok0 && order) ? (int64_t)__ldg(&order[r0]) : r0;   // output rows
  const int64_t y1 = (ok1 && order) ? (int64_t)__ldg(&order[r1]) : r1;
  const float o0[4] = {accm[0][0] + accc[0][0], accm[0][1] + accc[0][1], accm[1][0] + accc[1][0], accm[1][1] + accc[1][1]};
  const float o1[4] = {accm[0][2] + accc[0][2], accm[0][3] + accc[0][3], accm[1][2] + accc[1][2], accm[1][3] + accc[1][3]};
  if (ksplit == 1) {
    if (ok0) *(float4*)(Y + (size_t)y0 * c_out + c) = make_float4(o0[0], o0[1], o0[2], o0[3]);
    if (ok1) *(float4*)(Y + (size_t)y1 * c_out + c) = make_float4(o1[0], o1[1], o1[2], o1[3]);
  } else if (touched) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (ok0) atomicAdd(Y + (size_t)y0 * c_out + c + i, o0[i]);
      if (ok1) atomicAdd(Y + (size_t)y1 * c_out + c + i, o1[i]);
    }
  }
}

template <int CIN, int COUT, int MT, int WS>
static int launch_mma(const float* X, const float* Wf, const int32_t* nbr, const int32_t* order, int64_t n_q, int K,
                      int mirror, float* Y, cudaStream_t s) {
  constexpr int R = kMmWarps * 16 * MT;
  constexpr size_t smem = (size_t)kMmMaxK * R * 4 + (WS == 0 ? 2 * (size_t)(CIN / 8) * (COUT / 8) * 32 * 16 : 0);
  static bool attr_set = false;
  if (!attr_set) {
    PGS_CUDA(cudaFuncSetAttribute(conv_mma_kernel<CIN, COUT, MT, WS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    attr_set = true;
  }
  const unsigned gx = (unsigned)((n_q + R - 1) / R);
  conv_mma_kernel<CIN, COUT, MT, WS><<<gx, kMmThreads, smem, s>>>(X, (const float4*)Wf, nbr, n_q, K, mirror, order, Y);
  return PGS_OK;
}

template <int CIN, int COUT>
static int launch_mma_shape(const float* X, const float* Wf, const int32_t* nbr, const int32_t* order, int64_t n_q,
                            int K, int mirror, float* Y, int ws, int mt, cudaStream_t s) {
  // two m-tiles per warp (weight fragments reused) where they fit 128 registers; 32 x 32 needs 166 -> one m-tile
  constexpr bool kTwo = (CIN * COUT <= 32 * 16);
  if (kTwo && mt != 1) {
    return launch_mma<CIN, COUT, (kTwo ? 2 : 1), 0>(X, Wf, nbr, order, n_q, K, mirror, Y, s);
  }
  return launch_mma<CIN, COUT, 1, 0>(X, Wf, nbr, order, n_q, K, mirror, Y, s);
}

}  // namespace pgs

using namespace pgs;

extern "C" {

int pgs_conv_mma_supported(int32_t c_in, int32_t c_out) {
  return c_in % 16 == 0 && c_out % 16 == 0 && c_in >= 16 && c_in <= 64 && c_out >= 16 && c_out <= 64;
}

size_t pgs_conv_mma_scratch_bytes(int32_t K, int32_t c_in, int32_t c_out) {
  return align_up((size_t)K * c_in * c_out * 2 * sizeof(float), 256);
}

int pgs_conv_fwd_mma(const float* X, const float* W, const int32_t* nbr, const int32_t* order, int64_t n_q, int32_t K,
                     int32_t c_in,
                     int32_t c_out, int32_t mirror, int32_t w_transposed, float* Y, void* scratch,
                     size_t scratch_bytes, void* stream) {
  PGS_CHECK_ARG(K >= 1 && K <= kMmMaxK, "kernel volume must be in 1..27 for the mma path");
  PGS_CHECK_ARG(pgs_conv_mma_supported(c_in, c_out), "channel counts not supported by the mma path");
  PGS_CHECK_ARG(nbr != nullptr, "the mma path needs a gather table");
  PGS_CHECK_ARG(scratch_bytes >= pgs_conv_mma_scratch_bytes(K, c_in, c_out), "scratch too small");
  if (n_q == 0) return PGS_OK;
  cudaStream_t s = (cudaStream_t)stream;
  float* Wf = (float*)scratch;
  static int ws = -1, mt = -1, queued = 1;
  if (ws < 0) {
    const char* q = getenv("PGS_MMA_QUEUE");  // "0": register double-buffer version (conv_mma_kernel)
    queued = !(q && q[0] == '0');
    const char* e = getenv("PGS_MMA_WSRC");   // "ldg": weight fragments straight from global / L1 (experiment)
    ws = (e && e[0] == 'l') ? 1 : 0;
    const char* m = getenv("PGS_MMA_MT");     // "1": one m-tile per warp everywhere (experiment)
    mt = (m && m[0] == '1') ? 1 : 0;
  }
  const int64_t total = (int64_t)K * (c_in / 8) * (c_out / 8) * 32;
  int pg = (int)((total + 255) / 256);
  if (pg > kNumSM * 8) pg = kNumSM * 8;
  if (W != nullptr)   // W == NULL: scratch already holds the arranged weights (pgs_conv_prep_weights_batch)
    conv_mma_prep_weights_kernel<<<pg, 256, 0, s>>>(W, K, c_in, c_out, w_transposed, Wf);
  int rc = PGS_ERR_INVALID;
#define PGS_MMA_CASE(CI, CO)                                                                   \
  case CI * 1000 + CO:                                                                         \
    rc = (queued && CI * CO <= 32 * 16) ? launch_mmaq<CI, CO>(X, Wf, nbr, order, n_q, K, mirror, Y, s) \
                : launch_mma_shape<CI, CO>(X, Wf, nbr, order, n_q, K, mirror, Y, ws, mt, s);     \
    break;
  switch (c_in * 1000 + c_out) {
    PGS_MMA_CASE(16, 16) PGS_MMA_CASE(16, 32) PGS_MMA_CASE(16, 48) PGS_MMA_CASE(16, 64)
    PGS_MMA_CASE(32, 16) PGS_MMA_CASE(32, 32) PGS_MMA_CASE(32, 48) PGS_MMA_CASE(32, 64)
    PGS_MMA_CASE(48, 16) PGS_MMA_CASE(48, 32) PGS_MMA_CASE(48, 48) PGS_MMA_CASE(48, 64)
    PGS_MMA_CASE(64, 16) PGS_MMA_CASE(64, 32) PGS_MMA_CASE(64, 48) PGS_MMA_CASE(64, 64)
    default:
      set_error("pgs_conv_fwd_mma: unsupported shape %d -> %d", c_in, c_out);
      return PGS_ERR_INVALID;
  }
#undef PGS_MMA_CASE
  if (rc != PGS_OK) return rc;
  count_launch(W != nullptr ? 2 : 1);
  PGS_CHECK_LAUNCH();
  return PGS_OK;
}


int pgs_conv_mma_split_supported(int32_t c_in, int32_t c_out) {
  return c_in % 16 == 0 && c_out % 16 == 0 && c_in >= 16 && c_out >= 16;
}

/* few-row variant: warp items (16 rows x 16 output channels x part of the offsets), see conv_mma_split_kernel */
int pgs_conv_fwd_mma_split(const float* X, const float* W, const int32_t* nbr, const int32_t* order, int64_t n_q,
                           int32_t K, int32_t c_in,
                           int32_t c_out, int32_t mirror, int32_t w_transposed, float* Y, void* scratch,
                           size_t scratch_bytes, void* stream) {
  PGS_CHECK_ARG(K >= 1 && K <= kMmMaxK, "kernel volume must be in 1..27 for the mma path");
  PGS_CHECK_ARG(pgs_conv_mma_split_supported(c_in, c_out), "channel counts must be multiples of 16");
  PGS_CHECK_ARG(nbr != nullptr, "the mma path needs a gather table");
  PGS_CHECK_ARG(scratch_bytes >= pgs_conv_mma_scratch_bytes(K, c_in, c_out), "scratch too small");
  if (n_q == 0) return PGS_OK;
  cudaStream_t s = (cudaStream_t)stream;
  float* Wf = (float*)scratch;
  const int64_t total = (int64_t)K * (c_in / 8) * (c_out / 8) * 32;
  int pg = (int)((total + 255) / 256);
  if (pg > kNumSM * 8) pg = kNumSM * 8;
  if (W != nullptr)   // W == NULL: scratch already holds the arranged weights (pgs_conv_prep_weights_batch)
    conv_mma_prep_weights_kernel<<<pg, 256, 0, s>>>(W, K, c_in, c_out, w_transposed, Wf);
  const int mtiles = (int)((n_q + 15) / 16);
  const int64_t base = (int64_t)mtiles * (c_out / 16);
  static int target = -1;
  if (target < 0) {
    const char* e = getenv("PGS_MMA_SPLIT_WARPS");   // warps to aim for (default: 32 per SM)
    target = e ? atoi(e) : kNumSM * 32;
  }
  int ksplit = 1;
  if (K == 27) {
    const int64_t want = (target + base - 1) / base;
    ksplit = want >= 27 ? 27 : want >= 9 ? 9 : want >= 3 ? 3 : 1;
  }
  if (ksplit > 1) PGS_CUDA(cudaMemsetAsync(Y, 0, (size_t)n_q * c_out * sizeof(float), s));
  const int64_t items = base * ksplit;
  const unsigned gx = (unsigned)((items + kSplitWarps - 1) / kSplitWarps);
  conv_mma_split_kernel<<<gx, kSplitWarps * 32, 0, s>>>(X, (const float4*)Wf, nbr, n_q, K, c_in, c_out, mirror, ksplit,
                                                        mtiles, order, Y);
  count_launch(W != nullptr ? 2 : 1);
  PGS_CHECK_LAUNCH();
  return PGS_OK;
}

int pgs_conv_prep_weights_batch(const int64_t* desc, int32_t n_desc, int64_t max_elems, void* stream) {
  if (n_desc <= 0) return PGS_OK;
  PGS_CHECK_ARG(desc != nullptr && max_elems > 0, "bad descriptor table");
  int gx = (int)((max_elems + 255) / 256);
  if (gx > 64) gx = 64;
  conv_prep_batch_kernel<<<dim3(gx, n_desc), 256, 0, (cudaStream_t)stream>>>(desc, tc_corr16());
  count_launch();
  PGS_CHECK_LAUNCH();
  return PGS_OK;
}

}  // extern "C"
