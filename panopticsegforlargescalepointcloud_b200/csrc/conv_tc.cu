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
// both go to shared memory in the K-major SWIZZLE_128B UMMA layout (128-byte rows, 16-byte chunks XOR-swizzled).
// Warp-specialised: 8 producer warps gather / split / store through a register prefetch queue, a ninth warp's
// elected lane issues tcgen05.mma (kind::tf32); tcgen05.commit on an mbarrier recycles the 3-stage ring.
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

// K-major SWIZZLE_128B descriptor: rows of 128 bytes (32 tf32), 16-byte chunk c of row r stored at chunk
// position c ^ (r & 7); 8-row groups 1024 bytes apart (stride byte offset); layout type 2 @ [61,64).
// The tile base must be 1024-byte aligned; advancing K by 8 tf32 = +32 bytes on the start address.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}

// cute::UMMA::InstrDescriptor: c_format F32 (1) @4, a/b_format TF32 (2) @7/@10, K-major A/B, N>>3 @17, M>>4 @24
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// round-to-nearest-even to tf32 (low 13 mantissa bits cleared) with integer ops: cvt.rna.tf32.f32 runs on the
// slow conversion pipe (measured: 370 cycles per step for 32 conversions per thread), these run at full rate.
// Unbiased, unlike the tensor core's own truncation of fp32 operands.
__device__ __forceinline__ float to_tf32_rn(float x) {
  const uint32_t u = __float_as_uint(x);
  return __uint_as_float((u + 0x00000FFFu + ((u >> 13) & 1u)) & 0xFFFFE000u);
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
constexpr int kTcMaxK = 27;
constexpr int kTcStepK = 32;                      // contraction elements per pipeline step (= one 128-byte swizzle row)
constexpr int kTcProducers = 256;                 // warps 0..7 gather / split / store; they also run the epilogue
constexpr int kTcThreadsWS = kTcProducers + 32;   // warp 8 issues the MMAs
constexpr uint32_t kTcABytes = kTcM * kTcStepK * 4;   // 16 KB: one of A_hi / A_lo

__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// The contraction runs over the concatenated (kernel offset, input channel) axis in steps of 32 elements =
// two "halves" of 16 channels of one offset each (c_in is a multiple of 16), so a 16-channel layer packs two
// offsets into one step.  Operands sit in shared memory in the K-major SWIZZLE_128B UMMA layout.
//
// Measured on B200 (clock64 traces): with the no-swizzle (interleaved) layout each
// tcgen05.mma M128 x N96 x K8 took ~240 cycles (floor: 48) -- the tensor core's operand fetch from
// un-swizzled shared memory was the bottleneck, not the gather; a single thread group doing gather + MMA
// issue added another ~800 cycles of instruction issue per step on the critical path.
//
// NB = float4 weight elements per producer thread and half: ceil(4 * N / 256); RING = smem stages; PF = steps of
// global loads in flight per producer thread (register queue).
template <int NB, int RING, int PF>
__global__ void __launch_bounds__(kTcThreadsWS) conv_tc_kernel(const float* __restrict__ X,
                                                                const float* __restrict__ Wp,
                                                                const int32_t* __restrict__ nbr, int64_t n_q, int K,
                                                                int c_in, int c_out, int mirror, uint32_t tmem_cols,
                                                                float* __restrict__ Y) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_empty[RING];
  __shared__ uint64_t bar_done;
  __shared__ uint32_t tmem_base_s;
  __shared__ int idx_all[kTcMaxK][kTcM];
  __shared__ int klist[kTcMaxK];
  __shared__ unsigned kmask_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t row0 = (int64_t)blockIdx.x * kTcM;
  const int N = c_out;
  const uint32_t b_bytes = (uint32_t)N * kTcStepK * 4;
  const uint32_t stage_bytes = 2 * kTcABytes + 2 * b_bytes;   // [A_hi][A_lo][B_hi][B_lo], all 1024-byte aligned

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < RING; ++s) mbar_init(&bar_empty[s], 1);
    mbar_init(&bar_done, 1);
    fence_barrier_init();
    kmask_s = 0u;
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int J = c_in / 16;   // halves per offset

  // rulebook columns of this tile for all offsets at once; which offsets have any neighbour in the tile
  if (tid < kTcProducers) {
    unsigned mine = 0u;
    for (int e = tid; e < K * kTcM; e += kTcProducers) {
      const int k = e / kTcM, r = e - k * kTcM;
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
    for (int k = 0; k < K; ++k) {   // weight index order; table offset tk = mirror ? K-1-k : k
      const int tk = mirror ? (K - 1 - k) : k;
      if (km & (1u << tk)) {
        if (tid == 0) klist[n_off] = k;
        ++n_off;
      }
    }
  }
  __syncthreads();
  const int halves = n_off * J;
  const int total = (halves + 1) >> 1;   // steps; the last one may hold a single half (the other is zero)

  if (warp == kTcProducers / 32) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = make_idesc_tf32(kTcM, N);
    const uint32_t sbase = smem_u32(smem);
    for (int it = 0; it < total; ++it) {
      const int s = it % RING;
      named_bar_sync(1 + s, kTcThreadsWS);   // producers have filled stage s (and fenced it to the async proxy)
      if (lane == 0) {
        tc_fence_after();
        const uint32_t sa = sbase + (uint32_t)s * stage_bytes;
        const uint64_t ahi = make_smem_desc_sw128(sa), alo = make_smem_desc_sw128(sa + kTcABytes);
        const uint64_t bhi = make_smem_desc_sw128(sa + 2 * kTcABytes);
        const uint64_t blo = make_smem_desc_sw128(sa + 2 * kTcABytes + b_bytes);
#pragma unroll
        for (int kk = 0; kk < kTcStepK / 8; ++kk) {
          // +32 bytes per K = 8 step on the (>> 4 encoded) start address
          const uint64_t adv = (uint64_t)(kk * 2);
          // main products and the two correction products accumulate in SEPARATE tensor-memory tiles: the
          // accumulator add inside the tensor core truncates, so fewer adds into the large sum = less bias
          const uint32_t first = (it > 0 || kk > 0) ? 1u : 0u;
          umma_tf32(tmem_base, ahi + adv, bhi + adv, idesc, first);
          umma_tf32(tmem_base + (uint32_t)N, alo + adv, bhi + adv, idesc, first);
          umma_tf32(tmem_base + (uint32_t)N, ahi + adv, blo + adv, idesc, 1u);
        }
        umma_commit(&bar_empty[s]);
      }
      __syncwarp();
    }
    if (lane == 0) umma_commit(&bar_done);
  } else {
    // ===================== producers =====================
    const int nB4 = 4 * N;  // float4 elements of one B half ([q][n][4] in global memory)
    // gather mapping: per warp pass 8 consecutive rows x 4 channel quads (64 contiguous bytes of a feature row)
    const int g_q = (lane >> 3) & 3;
    const int g_r0 = (lane & 7) + 8 * warp, g_r1 = g_r0 + 64;
    // swizzled 16-byte slots of my A elements for half 0 / half 1 (rows g_r0 and g_r1 share r & 7)
    const uint32_t a_off0[2] = {(uint32_t)g_r0 * 128u + (uint32_t)(((0 + g_q) ^ (g_r0 & 7)) << 4),
                                (uint32_t)g_r0 * 128u + (uint32_t)(((4 + g_q) ^ (g_r0 & 7)) << 4)};
    const uint32_t a_off1[2] = {a_off0[0] + 64u * 128u, a_off0[1] + 64u * 128u};
    uint32_t b_off[NB][2];
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const int e = tid + i * kTcProducers;
      const int q = e / N, n = e - q * N;
      b_off[i][0] = (uint32_t)n * 128u + (uint32_t)(((0 + q) ^ (n & 7)) << 4);
      b_off[i][1] = (uint32_t)n * 128u + (uint32_t)(((4 + q) ^ (n & 7)) << 4);
    }

    float4 qa[PF][2][2], qb[PF][2][NB];
    int po[PF], pj[PF];   // (offset ordinal, 16-channel chunk) of the FIRST half of the step in each queue slot
    auto load_half = [&](int o, int j, float4* a, float4* b) {
      a[0] = make_float4(0.f, 0.f, 0.f, 0.f);
      a[1] = a[0];
      if (o < n_off) {
        const int k = klist[o];
        const int tk = mirror ? (K - 1 - k) : k;
        const int src0 = idx_all[tk][g_r0], src1 = idx_all[tk][g_r1];
        if (src0 >= 0) a[0] = __ldg((const float4*)(X + (size_t)src0 * c_in + j * 16 + g_q * 4));
        if (src1 >= 0) a[1] = __ldg((const float4*)(X + (size_t)src1 * c_in + j * 16 + g_q * 4));
        const float4* wsrc = (const float4*)(Wp + ((size_t)k * J + j) * (size_t)N * 16);
#pragma unroll
        for (int i = 0; i < NB; ++i) {
          const int e = tid + i * kTcProducers;
          b[i] = (e < nB4) ? __ldg(wsrc + e) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      } else {
#pragma unroll
        for (int i = 0; i < NB; ++i) b[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto load_step = [&](int o, int j, int p) {
      load_half(o, j, qa[p][0], qb[p][0]);
      int o1 = o, j1 = j + 1;
      if (j1 == J) {
        j1 = 0;
        ++o1;
      }
      load_half(o1, j1, qa[p][1], qb[p][1]);
    };
    {
      int o = 0, j = 0;
#pragma unroll
      for (int p = 0; p < PF; ++p) {
        po[p] = o;
        pj[p] = j;
        if (p < total) load_step(o, j, p);
        j += 2;
        while (j >= J) {
          j -= J;
          ++o;
        }
      }
    }
    for (int i0 = 0; i0 < total; i0 += PF) {
#pragma unroll
      for (int p = 0; p < PF; ++p) {
        const int it = i0 + p;
        if (it < total) {
          const int s = it % RING;
          uint8_t* st = smem + (size_t)s * stage_bytes;
          if (it >= RING) {   // MMAs of step it - RING must have drained this stage
            mbar_wait(&bar_empty[s], (uint32_t)(((it / RING) - 1) & 1));
            tc_fence_after();
          }
          uint8_t* Ahi = st;
          uint8_t* Alo = st + kTcABytes;
          uint8_t* Bhi = st + 2 * kTcABytes;
          uint8_t* Blo = Bhi + b_bytes;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            split_store(qa[p][h][0], (float4*)(Ahi + a_off0[h]), (float4*)(Alo + a_off0[h]));
            split_store(qa[p][h][1], (float4*)(Ahi + a_off1[h]), (float4*)(Alo + a_off1[h]));
#pragma unroll
            for (int i = 0; i < NB; ++i)
              if (tid + i * kTcProducers < nB4)
                split_store(qb[p][h][i], (float4*)(Bhi + b_off[i][h]), (float4*)(Blo + b_off[i][h]));
          }
          fence_proxy_async();
          named_bar_arrive(1 + s, kTcThreadsWS);
          // refill the queue slot with the step PF ahead (division-free bookkeeping)
          int o = po[p], j = pj[p] + 2 * PF;
          while (j >= J) {
            j -= J;
            ++o;
          }
          po[p] = o;
          pj[p] = j;
          if (it + PF < total) load_step(o, j, p);
        }
      }
    }
    if (total > 0) {
      mbar_wait(&bar_done, 0);
      tc_fence_after();
    }
    // epilogue: warp w reads TMEM lanes 32*(w%4).., columns of half (w/4); main + correction tiles summed in fp32
    const int lq = warp & 3, half = warp >> 2;
    const int64_t r = row0 + lq * 32 + lane;
    const int c_begin = half * (N / 2), c_end = c_begin + N / 2;   // N is a multiple of 16
    for (int c = c_begin; c < c_end; c += 8) {
      uint32_t v[8], u[8];
      if (total > 0) {
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
  PGS_CHECK_ARG(K >= 1 && K <= kTcMaxK, "kernel volume must be in 1..27 for the tcgen05 path");
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
  const size_t stage = 2 * (size_t)kTcABytes + 2 * (size_t)c_out * kTcStepK * 4;
  static bool attr_set = false;
  if (!attr_set) {
    PGS_CUDA(cudaFuncSetAttribute(conv_tc_kernel<1, 3, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    PGS_CUDA(cudaFuncSetAttribute(conv_tc_kernel<2, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    PGS_CUDA(cudaFuncSetAttribute(conv_tc_kernel<3, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  const unsigned gx = (unsigned)((n_q + kTcM - 1) / kTcM);
  const uint32_t cols = tmem_cols_for(2 * c_out);
  if (c_out <= 64)
    conv_tc_kernel<1, 3, 3><<<gx, kTcThreadsWS, 3 * stage, s>>>(X, Wp, nbr, n_q, K, c_in, c_out, mirror, cols, Y);
  else if (c_out <= 128)
    conv_tc_kernel<2, 2, 2><<<gx, kTcThreadsWS, 2 * stage, s>>>(X, Wp, nbr, n_q, K, c_in, c_out, mirror, cols, Y);
  else
    conv_tc_kernel<3, 2, 2><<<gx, kTcThreadsWS, 2 * stage, s>>>(X, Wp, nbr, n_q, K, c_in, c_out, mirror, cols, Y);
  count_launch(2);
  PGS_CHECK_LAUNCH();
  return PGS_OK;
}

}  // extern "C"
