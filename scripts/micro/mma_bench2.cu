// Micro-benchmark 2: does tcgen05.mma (M=128, SS) throughput scale with the number of issuing warps per CTA
// and with the number of CTAs per SM?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench2 mma_bench2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void umma_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t desc_none(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// issuers = number of warps that each issue `iters` MMAs into their own accumulator (columns w*N)
__global__ void bench(int N, int issuers, int iters, uint32_t cols, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[8];
  __shared__ uint32_t tbase;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 24 * 1024 / 4; i += blockDim.x) ((float*)smem)[i] = 1.0f;
  if (tid == 0) {
    for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tbase;
  long long t0 = clock64();
  if (warp < issuers && lane == 0) {
    const uint32_t sa = smem_u32(smem);
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t a = desc_none(sa, 128 * 16, 128), b = desc_none(sa + 8192, (uint32_t)N * 16, 128);
    for (int i = 0; i < iters; ++i) umma_tf32(tm + (uint32_t)(warp * N), a, b, idesc, 1u);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[warp])) : "memory");
    asm volatile("{\n.reg .pred p;\nWL:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DN;\nbra WL;\nDN:\n}\n" ::"r"(smem_u32(&bar[warp])), "r"(0u) : "memory");
  }
  __syncthreads();
  long long t1 = clock64();
  if (tid == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(cols) : "memory");
}
int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
  const int iters = 512;
  for (int N : {16, 64})
    for (int ctas_per_sm : {1, 2, 4})
      for (int issuers : {1, 2, 4}) {
        if (issuers * N > 128) continue;
        long long h;
        bench<<<148 * ctas_per_sm, 128, 32 * 1024>>>(N, issuers, iters, 128, d);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("N %3d ctas/SM %d issuers/CTA %d: %.1f cycles per MMA per SM (CTA0 span %lld cyc, %d MMAs on the SM) %s\n", N, ctas_per_sm, issuers,
               h / (double)(iters * issuers * ctas_per_sm), h, iters * issuers * ctas_per_sm, e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  return 0;
}
