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
// channels through the rulebook table with 16-byte loads, split them into tf32 hi / lo and store them to shared
// memory in the canonical K-major no-swizzle UMMA layout [chunk of 4 floats][row][4] (2-stage ring).  The matching
// weight chunk -- pre-arranged AND pre-split into hi / lo planes once per step by the prep kernel -- is a contiguous
// 128 * Cout-byte tile that ONE thread fetches with a TMA tensor copy (cp.async.bulk.tensor.2d -> UTMALDG) into its
// own 2..4-deep ring, `ring - 2` steps ahead, completing on an mbarrier (expect-tx).  One elected thread issues
// tcgen05.mma (kind::tf32); tcgen05.commit on an mbarrier recycles both rings.
//
// Measured on B200 (scripts/micro/mma_bench.cu, profiles/r1_mma_issue_rate.txt): with both operands in shared
// memory (SS mode) an M = 128 tcgen05.mma costs >= 97 cycles whatever N <= 96, the layout (interleaved or
// SWIZZLE_128B) or the element type -- the A-operand fetch (128 rows x 32 B) is exposed.  This kernel is
// therefore MMA-issue bound at 3 x (Cin / 8) x 97 cycles per kernel offset and 128-row tile; a warp-specialised
// SWIZZLE_128B variant (kept as scripts/micro/conv_tc_v5_swizzle128_ws.cu.txt) measured no faster.  The way out
// is to feed A from tensor memory (TS mode, floor M*N/256 cycles).
//
// Precision: the parity bar is 1e-4 against an fp32 oracle; single-pass tf32 (10-bit mantissa) cannot meet
// it, so operands are split a = hi + lo (hi = rn_tf32(a)) on the way into shared memory: hi*hi accumulates in one
// TMEM tile (tf32 MMA), the correction lo*hi + hi*lo in a second one -- as ONE bf16 K = 16 MMA per 8 channels whose
// 16 contraction slots are [x_lo x 4 | x_hi x 4] . [w_hi x 4 | w_lo x 4] twice (error ~2^-20 per product; the
// kernel is MMA-issue bound, so 2 instead of 3 MMAs per 8 channels is the lever; PGS_TC_CORR=tf32 selects the older
// two-tf32-MMA correction).
#include <cuda.h>   // CUtensorMap (types only: cuTensorMapEncodeTiled is fetched with cudaGetDriverEntryPoint)

#include <cstdlib>

#include "common.cuh"
#include "conv_prep.cuh"

namespace pgs {

constexpr int kTcThreads = 256;
constexpr int kTcM = 128;   // rows per CTA == UMMA M
constexpr int kTcKC = 16;   // input channels per pipeline step (two K=8 tf32 MMAs)
static_assert(kTcKC == kPrepKC, "conv_prep.cuh must use the same chunk size");
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
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA: one box of the 2-D weight tensor (coordinates {0, row}) -> shared memory, bytes counted on `bar`
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* tmap, int32_t c0, int32_t c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst_smem)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
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
// kind::f16 with bf16 operands (K = 16 per instruction = the same 32 bytes per operand row as a tf32 K = 8 MMA)
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
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
// the same with a/b_format BF16 (1)
__device__ __forceinline__ uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// round-to-nearest-even to tf32 (low 13 mantissa bits cleared) with integer ops: cvt.rna.tf32.f32 runs on the
// slow conversion pipe (clock64 trace: 370 cycles per step for 32 conversions per thread), these run at full
// rate.  Unbiased, unlike the tensor core's own truncation of fp32 operands.
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
// bf16-correction form: hi as above; the second 16 bytes are the 8 bf16 contraction slots [lo x 4 | hi x 4] of the
// correction MMA (they meet [w_hi x 4 | w_lo x 4] of the same channels, conv_prep.cuh).  lo*hi + hi*lo is 2^-11 of the
// main product, so 8 mantissa bits keep every term's error at the 2^-20 level of the dropped lo*lo product (the
// mma.sync kernels use the same split, conv_mma.cu:mma_corr).
__device__ __forceinline__ void split_store16(float4 v, float4* hi_dst, float4* corr_dst) {
  float4 h;
  h.x = to_tf32_rn(v.x);
  h.y = to_tf32_rn(v.y);
  h.z = to_tf32_rn(v.z);
  h.w = to_tf32_rn(v.w);
  uint4 c;
  c.x = prep_pack_bf16(__float_as_uint(v.x - h.x), __float_as_uint(v.y - h.y));
  c.y = prep_pack_bf16(__float_as_uint(v.z - h.z), __float_as_uint(v.w - h.w));
  c.z = prep_pack_bf16(__float_as_uint(h.x), __float_as_uint(h.y));
  c.w = prep_pack_bf16(__float_as_uint(h.z), __float_as_uint(h.w));
  *hi_dst = h;
  *(uint4*)corr_dst = c;
}

// ---------------------------------------------------------------------------------------------
// weight pre-arrangement:  Wp[k][j][plane][q][n][4] = split_plane(B_k[n][j*16 + q*4 .. +3])   with  B_k = W[k]^T
// (forward: n = cout, contraction over cin) or B_k = W[k] (input gradient: n = cin, contraction over cout);
// plane 0 = tf32 hi, plane 1 = tf32 lo (conv_prep.cuh).  One (k, j) chunk = 128 * N contiguous bytes = one TMA box.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv_tc_prep_weights_kernel(const float* __restrict__ W, int K, int c_in,
                                                                    int c_out, int w_transposed, int corr16,
                                                                    float* __restrict__ Wp) {
  // kernel-side naming: contraction length C (= c_in of the launch), N output channels (= c_out of the launch)
  const int64_t total = 2 * (int64_t)K * c_in * c_out;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
    Wp[e] = prep_tc_elem(W, K, c_in, c_out, w_transposed, corr16, e);
}

// ---------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------
constexpr int kTcMaxK = 27;
constexpr int kTcPrefetch = 4;   // steps of global loads in flight per thread (register queue)

constexpr int kTcMaxBRing = 4;

__global__ void __launch_bounds__(kTcThreads) conv_tc_kernel(const float* __restrict__ X,
                                                              const __grid_constant__ CUtensorMap w_map,
                                                              const int32_t* __restrict__ nbr, int64_t n_q, int K,
                                                              int c_in, int c_out, int mirror, uint32_t tmem_cols,
                                                              int ksplit, int bring, int corr16,
                                                              const int32_t* __restrict__ order, float* __restrict__ Y) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_empty[kTcStages];
  __shared__ uint64_t bar_full[kTcMaxBRing];
  __shared__ uint64_t bar_done;
  __shared__ uint32_t tmem_base_s;
  __shared__ int idx_all[kTcMaxK][kTcM];
  __shared__ int klist[kTcMaxK];
  __shared__ unsigned kmask_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t row0 = (int64_t)blockIdx.x * kTcM;
  const int N = c_out;
  const uint32_t a_bytes = kTcM * kTcKC * 4;        // one of hi / lo
  const uint32_t b_bytes = (uint32_t)N * kTcKC * 4;
  const uint32_t stage_bytes = 2 * a_bytes;         // A ring: [A_hi][A_lo]
  uint8_t* const bsm = smem + (size_t)kTcStages * stage_bytes;   // B ring: `bring` x [B_hi][B_lo], filled by TMA
  const int look = bring - 2;                       // the weight tile of step it + look is requested at step it

  if (tid == 0) {
    mbar_init(&bar_empty[0], 1);
    mbar_init(&bar_empty[1], 1);
    for (int b = 0; b < kTcMaxBRing; ++b) mbar_init(&bar_full[b], 1);
    mbar_init(&bar_done, 1);
    fence_barrier_init();
    kmask_s = 0u;
    asm volatile("prefetch.tensormap [%0];" ::"l"(&w_map) : "memory");
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t idesc = make_idesc_tf32(kTcM, N), idesc16 = make_idesc_bf16(kTcM, N);
  const int J = c_in / kTcKC;

  // rulebook columns of this tile for all offsets at once; which offsets have any neighbour in the tile
  {
    unsigned mine = 0u;
    for (int e = tid; e < K * kTcM; e += kTcThreads) {
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
  // split-K over kernel offsets for layers with too few row tiles to fill the GPU: CTA (x, y) walks offsets
  // [off_begin, off_end) of the tile's non-empty list and adds its partial tile to Y with atomics (Y is zeroed)
  const int off_begin = (int)(((long long)n_off * blockIdx.y) / ksplit);
  const int off_end = (int)(((long long)n_off * (blockIdx.y + 1)) / ksplit);
  const int total = (off_end - off_begin) * J;

  // gather mapping: per warp pass 8 consecutive rows (conflict-free 128 B in shared memory) x 4 channel quads
  // (64 contiguous bytes of a feature row in global memory)
  const int g_q = (lane >> 3) & 3;
  const int g_r0 = (lane & 7) + 8 * warp, g_r1 = g_r0 + 64;

  float4 qa0[kTcPrefetch], qa1[kTcPrefetch];
  auto load_step = [&](int step, float4& a0, float4& a1) {
    const int o = step / J, j = step - o * J;
    const int k = klist[off_begin + o];
    const int tk = mirror ? (K - 1 - k) : k;
    const int src0 = idx_all[tk][g_r0], src1 = idx_all[tk][g_r1];
    a0 = make_float4(0.f, 0.f, 0.f, 0.f);
    a1 = a0;
    if (src0 >= 0) a0 = __ldg((const float4*)(X + (size_t)src0 * c_in + j * kTcKC + g_q * 4));
    if (src1 >= 0) a1 = __ldg((const float4*)(X + (size_t)src1 * c_in + j * kTcKC + g_q * 4));
  };
  // weight tile of pipeline step `step` -> B ring slot step % bring  (thread 0 only): rows [(k*J + j) * N/2, +N/2) of
  // the 2-D tensor [K*J*N/2][64 floats] that the prep kernel wrote = the contiguous 128*N-byte [hi | lo] chunk
  auto request_weights = [&](int step) {
    const int o = step / J, j = step - o * J;
    const int k = klist[off_begin + o];
    const int slot = step % bring;
    mbar_arrive_expect_tx(&bar_full[slot], 2 * b_bytes);
    tma_load_2d(bsm + (size_t)slot * 2 * b_bytes, &w_map, 0, (k * J + j) * (N / 2), &bar_full[slot]);
  };

#pragma unroll
  for (int p = 0; p < kTcPrefetch; ++p)
    if (p < total) load_step(p, qa0[p], qa1[p]);
  if (tid == 0)
    for (int p = 0; p < look && p < total; ++p) request_weights(p);

  for (int i0 = 0; i0 < total; i0 += kTcPrefetch) {
#pragma unroll
    for (int p = 0; p < kTcPrefetch; ++p) {
      const int it = i0 + p;
      if (it < total) {
        const int s = it & 1;
        uint8_t* st = smem + (size_t)s * stage_bytes;
        if (it >= kTcStages) {   // MMAs of step it-2 must have drained this stage
          mbar_wait(&bar_empty[s], (uint32_t)(((it - kTcStages) >> 1) & 1));
          tc_fence_after();
        }
        // slot (it + look) % bring held the weights of step it + look - bring = it - 2, whose MMAs have drained
        if (tid == 0 && it + look < total) request_weights(it + look);
        float4* Ahi = (float4*)st;
        float4* Alo = (float4*)(st + a_bytes);
        if (corr16) {
          split_store16(qa0[p], &Ahi[g_q * kTcM + g_r0], &Alo[g_q * kTcM + g_r0]);
          split_store16(qa1[p], &Ahi[g_q * kTcM + g_r1], &Alo[g_q * kTcM + g_r1]);
        } else {
          split_store(qa0[p], &Ahi[g_q * kTcM + g_r0], &Alo[g_q * kTcM + g_r0]);
          split_store(qa1[p], &Ahi[g_q * kTcM + g_r1], &Alo[g_q * kTcM + g_r1]);
        }
        if (it + kTcPrefetch < total) load_step(it + kTcPrefetch, qa0[p], qa1[p]);   // refill the queue slot
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
          const int slot = it % bring;
          mbar_wait(&bar_full[slot], (uint32_t)((it / bring) & 1));   // TMA bytes of this step's weight tile landed
          tc_fence_after();
          const uint32_t sa = smem_u32(st);
          const uint32_t sb = smem_u32(bsm + (size_t)slot * 2 * b_bytes);
          const uint32_t a_lbo = kTcM * 16, b_lbo = (uint32_t)N * 16;
#pragma unroll
          for (int kk = 0; kk < kTcKC / 8; ++kk) {
            const uint64_t ahi = make_smem_desc(sa + kk * 2 * a_lbo, a_lbo, 128);
            const uint64_t alo = make_smem_desc(sa + a_bytes + kk * 2 * a_lbo, a_lbo, 128);
            const uint64_t bhi = make_smem_desc(sb + kk * 2 * b_lbo, b_lbo, 128);
            const uint64_t blo = make_smem_desc(sb + b_bytes + kk * 2 * b_lbo, b_lbo, 128);
            // main products and the two correction products accumulate in SEPARATE tensor-memory tiles: the
            // accumulator add inside the tensor core truncates, so fewer adds into the large sum = less bias
            const uint32_t first = (it > 0 || kk > 0) ? 1u : 0u;
            umma_tf32(tmem_base, ahi, bhi, idesc, first);
            if (corr16) {   // [x_lo | x_hi] . [w_hi | w_lo] of these 8 channels in one K = 16 bf16 MMA
              umma_bf16(tmem_base + (uint32_t)N, alo, blo, idesc16, first);
            } else {
              umma_tf32(tmem_base + (uint32_t)N, alo, bhi, idesc, first);
              umma_tf32(tmem_base + (uint32_t)N, ahi, blo, idesc, 1u);
            }
          }
          umma_commit(&bar_empty[s]);
        }
      }
    }
  }

  if (tid == 0) umma_commit(&bar_done);
  if (total > 0) {
    mbar_wait(&bar_done, 0);
    tc_fence_after();
  }
  // epilogue: warp w reads TMEM lanes 32*(w%4).., columns of half (w/4); main + correction tiles summed in fp32
  {
    const int lq = warp & 3, half = warp >> 2;
    const int64_t rt = row0 + lq * 32 + lane;   // row of the (possibly occupancy-sorted) table
    const bool r_ok = rt < n_q;
    const int64_t r = (r_ok && order) ? (int64_t)__ldg(&order[rt]) : rt;   // output row
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
      if (r_ok) {
        if (ksplit == 1) {
          float4* y = (float4*)(Y + (size_t)r * c_out + c);
          y[0] = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
          y[1] = make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
        } else if (total > 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) atomicAdd(Y + (size_t)r * c_out + c + i, __uint_as_float(v[i]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_weight_map(CUtensorMap* map, void* base, cuuint64_t rows, uint32_t box_rows) {
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
        qres != cudaDriverEntryPointSuccess) {
      set_error("pgs_conv_fwd_tc: cuTensorMapEncodeTiled is not available from this driver");
      return PGS_ERR_CUDA;
    }
    encode = (EncodeTiledFn)fn;
  }
  const cuuint64_t gdim[2] = {64, rows};          // innermost first: 64 floats = 256 B per row
  const cuuint64_t gstride[1] = {256};            // bytes between rows
  const cuuint32_t box[2] = {64, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstride, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("pgs_conv_fwd_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return PGS_ERR_CUDA;
  }
  return PGS_OK;
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
  return align_up(2 * (size_t)K * c_in * c_out * sizeof(float), 256);   // tf32 hi + lo planes
}

int pgs_conv_fwd_tc(const float* X, const float* W, const int32_t* nbr, const int32_t* order, int64_t n_q, int32_t K,
                    int32_t c_in,
                    int32_t c_out, int32_t mirror, int32_t w_transposed, float* Y, void* scratch,
                    size_t scratch_bytes, void* stream) {
  PGS_CHECK_ARG(K >= 1 && K <= kTcMaxK, "kernel volume must be in 1..27 for the tcgen05 path");
  PGS_CHECK_ARG(pgs_conv_tc_supported(c_in, c_out), "channel counts not supported by the tcgen05 path");
  PGS_CHECK_ARG(nbr != nullptr || K == 1, "nbr == NULL requires K == 1");
  PGS_CHECK_ARG(scratch_bytes >= pgs_conv_tc_scratch_bytes(K, c_in, c_out), "scratch too small");
  PGS_CHECK_ARG(((uintptr_t)scratch & 15) == 0, "scratch must be 16-byte aligned (TMA source)");
  if (n_q == 0) return PGS_OK;
  cudaStream_t s = (cudaStream_t)stream;
  float* Wp = (float*)scratch;
  const int64_t total = 2 * (int64_t)K * c_in * c_out;
  int pg = (int)((total + 255) / 256);
  if (pg > kNumSM * 8) pg = kNumSM * 8;
  if (W != nullptr)   // W == NULL: scratch already holds the arranged weights (pgs_conv_prep_weights_batch)
    conv_tc_prep_weights_kernel<<<pg, 256, 0, s>>>(W, K, c_in, c_out, w_transposed, tc_corr16(), Wp);
  // tensor map of the arranged weights: [K * J * N/2 rows][64 floats], one box = N/2 rows = the 128*N-byte (k, j) chunk
  CUtensorMap w_map;
  {
    const cuuint64_t rows = (cuuint64_t)K * (c_in / kTcKC) * (c_out / 2);
    int rc = make_weight_map(&w_map, Wp, rows, (uint32_t)(c_out / 2));
    if (rc != PGS_OK) return rc;
  }
  // weight ring depth (2..4 tiles of 128 * Cout bytes next to the 32 KB A ring): the deepest ring that does not cost a
  // resident CTA on layers with enough row tiles to need them (a 4-deep ring at Cout = 64 drops 3 -> 2 CTAs per SM and
  // measured 7 % slower over the step); few-tile layers take the deepest ring that fits 100 KB.
  const size_t a_ring = (size_t)kTcStages * 2 * kTcM * kTcKC * 4, b_tile = 2 * (size_t)c_out * kTcKC * 4;
  const size_t static_smem = sizeof(int) * kTcMaxK * kTcM + 1024;
  auto ctas_per_sm = [&](int ring) {
    const size_t per = a_ring + (size_t)ring * b_tile + static_smem;
    size_t c = (size_t)227 * 1024 / per;
    return (int)(c > 4 ? 4 : c);   // 64 registers x 256 threads: at most 4 CTAs per SM
  };
  const unsigned gx0 = (unsigned)((n_q + kTcM - 1) / kTcM);
  const int needed = (int)((gx0 + kNumSM - 1) / kNumSM);
  const int target = ctas_per_sm(2) < needed ? ctas_per_sm(2) : needed;
  int bring = kTcMaxBRing;
  while (bring > 2 && (a_ring + (size_t)bring * b_tile > 100 * 1024 || ctas_per_sm(bring) < target)) --bring;
  const size_t smem = a_ring + (size_t)bring * b_tile;
  const unsigned gx = (unsigned)((n_q + kTcM - 1) / kTcM);
  static bool attr_set = false;
  if (!attr_set) {
    PGS_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr_set = true;
  }
  const uint32_t cols = tmem_cols_for(2 * c_out);
  // few row tiles: spread the kernel offsets of a tile over several CTAs (partial tiles meet in Y by atomicAdd)
  int ksplit = 1;
  if (K > 1 && gx * 2 <= (unsigned)kNumSM) {   // (splitting mid-size layers too measured slower: atomics + memset)
    ksplit = (int)((2 * kNumSM) / gx);
    if (ksplit > 9) ksplit = 9;
  }
  if (ksplit > 1) PGS_CUDA(cudaMemsetAsync(Y, 0, (size_t)n_q * c_out * sizeof(float), s));
  const dim3 grid(gx, ksplit);
  conv_tc_kernel<<<grid, kTcThreads, smem, s>>>(X, w_map, nbr, n_q, K, c_in, c_out, mirror, cols, ksplit, bring,
                                                tc_corr16(), order, Y);
  count_launch(W != nullptr ? 2 : 1);
  PGS_CHECK_LAUNCH();
  return PGS_OK;
}

}  // extern "C"
