// Micro-benchmark: issue rate of tcgen05.mma kind::tf32 (M=128) from shared-memory descriptors.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench mma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void umma_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint64_t desc_none(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__global__ void bench(int N, int mode, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  const int tid = threadIdx.x;
  for (int i = tid; i < 48 * 1024 / 4; i += blockDim.x) ((float*)smem)[i] = 1.0f;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tbase;
  if (tid == 0) {
    const uint32_t sa = smem_u32(smem);
    const bool f16 = (mode & 4) != 0;
    // tf32: a/b format 2 ; f16 kind with bf16: format 1
    const uint32_t fmt = f16 ? 1u : 2u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint64_t a, b;
    if (mode & 1) { a = desc_sw128(sa); b = desc_sw128(sa + 16384); }
    else { a = desc_none(sa, 128 * 16, 128); b = desc_none(sa + 16384, (uint32_t)N * 16, 128); }
    const bool alt = (mode & 2) != 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t d = tm + ((alt && (i & 1)) ? (uint32_t)N : 0u);
      if (f16) umma_f16(d, a, b, idesc, 1u); else umma_tf32(d, a, b, idesc, 1u);
    }
    long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("{\n.reg .pred p;\nWL:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DN;\nbra WL;\nDN:\n}\n" ::"r"(smem_u32(&bar)), "r"(0u) : "memory");
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}
int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 256;
  printf("mode bits: 1 = swizzle128 (else interleaved), 2 = alternate two accumulators, 4 = kind::f16 (bf16) instead of tf32\n");
  for (int mode = 0; mode < 8; ++mode)
    for (int N : {16, 32, 64, 96, 128, 192}) {
      long long h[2];
      bench<<<1, 128, 64 * 1024>>>(N, mode, iters, d);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("mode %d N %3d: issue %.1f cyc/mma, complete %.1f cyc/mma (floor %d) %s\n", mode, N, h[0] / (double)iters, h[1] / (double)iters, 128 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
