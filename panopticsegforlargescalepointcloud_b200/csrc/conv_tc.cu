// Sparse convolution forward / backward-input on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// Same contract as conv_fwd_kernel in conv.cu (reference call sites: torch_points3d/modules/MinkowskiEngine/
// api_modules.py:26-55,244-270,293 -> MinkowskiEngine ConvolutionForward/BackwardKernelGPU):
//     Y[q] = sum_k X[nbr[tk(k)][q]] * W[k]            (or W[k]^T for the input gradient)
//
// Formulation: output stationary gather-GEMM.  One CTA owns 128 output rows (UMMA M = 128, cta_group::1) and
// ALL output channels (UMMA N = Cout, 16..192), so the fp32 accumulator tile D[128 x Cout] lives in tensor
// memory for the whole walk over the K kernel offsets and Cin/16 channel chunks and every output row is
// written exactly once (no atomics, deterministic).  Per step the CTA's threads gather 128 rows x 16 input
// channels through the rulebook table with 16-byte loads, and copy the matching pre-arranged weight chunk;
// both go to shared memory in the canonical K-major no-swizzle UMMA layout [chunk of 4 floats][row][4].
// One elected thread issues tcgen05.mma (kind::tf32); tcgen05.commit on an mbarrier recycles the 2-stage ring.
//
// Precision: the parity bar is 1e-4 against an fp32 oracle; single-pass tf32 (10-bit mantissa) cannot meet
// it, so operands are split a = hi + lo (hi = rn_tf32(a), lo = rn_tf32(a - hi)) on the way into shared memory
// and three MMAs accumulate hi*hi + hi*lo + lo*hi into the same TMEM tile (error ~2^-21 per product).
#include "common.cuh"

namespace pgs {

constexpr int kTcThreads = 256;
constexpr int kTcM = 128;   // rows per CTA == UMMA M
constexpr int kTcKC = 16;   // input channels per pipeline step (two K=8 tf32 MMAs)
constexpr int kTcStages = 2;

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, no-swizzle shared memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start >> 4 | [16,30) leading byte offset >> 4 (between the two 16-byte K chunks of one MMA)
//   [32,46) stride byte offset >> 4 (between 8-row groups) | [46,48) version = 1 | [61,64) layout = 0
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

// cute::UMMA::InstrDescriptor: c_format F32 (1) @4, a/b_format TF32 (2) @7/@10, K-major A/B, N>>3 @17, M>>4 @24
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// round-to-nearest tf32 (low 13 mantissa bits cleared); unbiased, unlike the tensor core's own truncation
__device__ __forceinline__ float to_tf32_rn(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// a = hi + lo with hi = rn_tf32(a), lo = rn_tf32(a - hi)   (a - hi is exact in fp32)
__device__ __forceinline__ void split_store(float4 v, float4* hi_dst, float4* lo_dst) {
  float4 h, l;
  h.x = to_tf32_rn(v.x);
  h.y = to_tf32_rn(v.y);
  h.z = to_tf32_rn(v.z);
  h.w = to_tf32_rn(v.w);
  l.x = to_tf32_rn(v.x - h.x);
  l.y = to_tf32_rn(v.y - h.y);
  l.z = to_tf32_rn(v.z - h.z);
  l.w = to_tf32_rn(v.w - h.w);
  *hi_dst = h;
  *lo_dst = l;
}

// ---------------------------------------------------------------------------------------------
// weight pre-arrangement:  Wp[k][j][q][n][4] = B_k[n][j*16 + q*4 .. +3]   with  B_k = W[k]^T (forward: n = cout,
// contraction over cin) or B_k = W[k] (input gradient: n = cin, contraction over cout)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv_tc_prep_weights_kernel(const float* __restrict__ W, int K, int c_in,
                                                                    int c_out, int w_transposed,
                                                                    float* __restrict__ Wp) {
  // kernel-side naming: contraction length C (= c_in of the launch), N output channels (= c_out of the launch)
  const int C = c_in, N = c_out;
  const int64_t total = (int64_t)K * C * N;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    // e enumerates the OUTPUT layout: ((((k * J + j) * 4 + q) * N + n) * 4 + t)
    const int t = (int)(e & 3);
    int64_t r = e >> 2;
    const int n = (int)(r % N);
    r /= N;
    const int q = (int)(r & 3);
    r >>= 2;
    const int J = C / kTcKC;
    const int j = (int)(r % J);
    const int k = (int)(r / J);
    const int c = j * kTcKC + q * 4 + t;
    // !w_transposed: W stored [K][C][N];  w_transposed: W stored [K][N][C]
    const float v = w_transposed ? W[((int64_t)k * N + n) * C + c] : W[((int64_t)k * C + c) * N + n];
    Wp[e] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTcThreads) conv_tc_kernel(const float* __restrict__ X, const float* __restrict__ Wp,
                                                              const int32_t* __restrict__ nbr, int64_t n_q, int K,
                                                              int c_in, int c_out, int mirror, uint32_t tmem_cols,
                                                              float* __restrict__ Y) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_empty[kTcStages];
  __shared__ uint64_t bar_done;
  __shared__ uint32_t tmem_base_s;
  __shared__ int idx_s[kTcM];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t row0 = (int64_t)blockIdx.x * kTcM;
  const int N = c_out;
  const uint32_t a_bytes = kTcM * kTcKC * 4;        // one of hi / lo
  const uint32_t b_bytes = (uint32_t)N * kTcKC * 4;
  const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
  // stage layout: [A_hi][A_lo][B_hi][B_lo]

  if (tid == 0) {
    mbar_init(&bar_empty[0], 1);
    mbar_init(&bar_empty[1], 1);
    mbar_init(&bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t idesc = make_idesc_tf32(kTcM, N);
  const int J = c_in / kTcKC;
  const int nB4 = 4 * N;  // float4 elements of one B chunk

  // gather mapping (see header): 8 consecutive rows x 4 channel quads per warp pass
  int g_r[2], g_q;
  g_q = (lane >> 3) & 3;
#pragma unroll
  for (int i = 0; i < 2; ++i) g_r[i] = (lane & 7) + 8 * (warp + 8 * i);

  int it = 0;
  for (int k = 0; k < K; ++k) {
    const int tk = mirror ? (K - 1 - k) : k;
    int has = 0;
    if (tid < kTcM) {
      const int64_t r = row0 + tid;
      int v = -1;
      if (r < n_q) v = nbr ? __ldg(&nbr[(int64_t)tk * n_q + r]) : (int)r;
      idx_s[tid] = v;
      has = v >= 0;
    }
    if (!__syncthreads_or(has)) continue;
    const int src0 = idx_s[g_r[0]], src1 = idx_s[g_r[1]];

    for (int j = 0; j < J; ++j, ++it) {
      const int s = it & 1;
      uint8_t* st = smem + (size_t)s * stage_bytes;
      // global loads first (latency overlaps the wait for the stage to drain)
      float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
      if (src0 >= 0) a0 = __ldg((const float4*)(X + (size_t)src0 * c_in + j * kTcKC + g_q * 4));
      if (src1 >= 0) a1 = __ldg((const float4*)(X + (size_t)src1 * c_in + j * kTcKC + g_q * 4));
      const float4* wsrc = (const float4*)(Wp + ((size_t)k * J + j) * (size_t)N * kTcKC);
      float4 b[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int e = tid + i * kTcThreads;
        b[i] = (e < nB4) ? __ldg(wsrc + e) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (it >= kTcStages) {
        mbar_wait(&bar_empty[s], (uint32_t)(((it - kTcStages) >> 1) & 1));
        tc_fence_after();
      }
      float4* Ahi = (float4*)st;
      float4* Alo = (float4*)(st + a_bytes);
      float4* Bhi = (float4*)(st + 2 * a_bytes);
      float4* Blo = (float4*)(st + 2 * a_bytes + b_bytes);
      split_store(a0, &Ahi[g_q * kTcM + g_r[0]], &Alo[g_q * kTcM + g_r[0]]);
      split_store(a1, &Ahi[g_q * kTcM + g_r[1]], &Alo[g_q * kTcM + g_r[1]]);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int e = tid + i * kTcThreads;
        if (e < nB4) split_store(b[i], &Bhi[e], &Blo[e]);
      }
      fence_proxy_async();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        const uint32_t sa = smem_u32(st);
        const uint32_t a_lbo = kTcM * 16, b_lbo = (uint32_t)N * 16;
#pragma unroll
        for (int kk = 0; kk < kTcKC / 8; ++kk) {
          const uint64_t ahi = make_smem_desc(sa + kk * 2 * a_lbo, a_lbo, 128);
          const uint64_t alo = make_smem_desc(sa + a_bytes + kk * 2 * a_lbo, a_lbo, 128);
          const uint64_t bhi = make_smem_desc(sa + 2 * a_bytes + kk * 2 * b_lbo, b_lbo, 128);
          const uint64_t blo = make_smem_desc(sa + 2 * a_bytes + b_bytes + kk * 2 * b_lbo, b_lbo, 128);
          // main products and the two correction products accumulate in SEPARATE tensor-memory tiles: the
          // accumulator add inside the tensor core truncates, so fewer adds into the large sum = less bias
          const uint32_t first = (it > 0 || kk > 0) ? 1u : 0u;
          umma_tf32(tmem_base, ahi, bhi, idesc, first);
          umma_tf32(tmem_base + (uint32_t)N, alo, bhi, idesc, first);
          umma_tf32(tmem_base + (uint32_t)N, ahi, blo, idesc, 1u);
        }
        umma_commit(&bar_empty[s]);
      }
    }
  }

  if (tid == 0) umma_commit(&bar_done);
  if (it > 0) {
    mbar_wait(&bar_done, 0);
    tc_fence_after();
  }
  // epilogue: warp w reads TMEM lanes 32*(w%4).., columns of half (w/4)
  {
    const int lq = warp & 3, half = warp >> 2;
    const int64_t r = row0 + lq * 32 + lane;
    const int c_begin = half * (N / 2), c_end = c_begin + N / 2;   // N is a multiple of 16
    for (int c = c_begin; c < c_end; c += 8) {
      uint32_t v[8], u[8];
      if (it > 0) {
        tmem_ld8(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)c, v);
        tmem_ld8(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(N + c), u);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(u[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0u;
      }
      if (r < n_q) {
        float4* y = (float4*)(Y + (size_t)r * c_out + c);
        y[0] = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
        y[1] = make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

static inline uint32_t tmem_cols_for(int n) {
  uint32_t c = 32;
  while ((int)c < n) c <<= 1;
  return c;
}

}  // namespace pgs

using namespace pgs;

extern "C" {

int pgs_conv_tc_supported(int32_t c_in, int32_t c_out) {
  return (c_in % kTcKC == 0) && (c_out % 16 == 0) && c_out >= 16 && c_out <= 192 && c_in >= kTcKC;
}

size_t pgs_conv_tc_scratch_bytes(int32_t K, int32_t c_in, int32_t c_out) {
  return align_up((size_t)K * c_in * c_out * sizeof(float), 256);
}

int pgs_conv_fwd_tc(const float* X, const float* W, const int32_t* nbr, int64_t n_q, int32_t K, int32_t c_in,
                    int32_t c_out, int32_t mirror, int32_t w_transposed, float* Y, void* scratch,
                    size_t scratch_bytes, void* stream) {
  PGS_CHECK_ARG(K >= 1, "bad shape");
  PGS_CHECK_ARG(pgs_conv_tc_supported(c_in, c_out), "channel counts not supported by the tcgen05 path");
  PGS_CHECK_ARG(nbr != nullptr || K == 1, "nbr == NULL requires K == 1");
  PGS_CHECK_ARG(scratch_bytes >= pgs_conv_tc_scratch_bytes(K, c_in, c_out), "scratch too small");
  if (n_q == 0) return PGS_OK;
  cudaStream_t s = (cudaStream_t)stream;
  float* Wp = (float*)scratch;
  const int64_t total = (int64_t)K * c_in * c_out;
  int pg = (int)((total + 255) / 256);
  if (pg > kNumSM * 8) pg = kNumSM * 8;
  conv_tc_prep_weights_kernel<<<pg, 256, 0, s>>>(W, K, c_in, c_out, w_transposed, Wp);
  const size_t smem = (size_t)kTcStages * (2 * kTcM * kTcKC * 4 + 2 * (size_t)c_out * kTcKC * 4);
  static bool attr_set = false;
  if (!attr_set) {
    PGS_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr_set = true;
  }
  const unsigned gx = (unsigned)((n_q + kTcM - 1) / kTcM);
  conv_tc_kernel<<<gx, kTcThreads, smem, s>>>(X, Wp, nbr, n_q, K, c_in, c_out, mirror, tmem_cols_for(2 * c_out), Y);
  count_launch(2);
  PGS_CHECK_LAUNCH();
  return PGS_OK;
}

}  // extern "C"
