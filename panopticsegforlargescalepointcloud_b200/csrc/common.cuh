// Shared helpers for the sm_100a kernels behind include/pgs_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/pgs_b200.h"

namespace pgs {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define PGS_CHECK_ARG(cond, msg)                 \
  do {                                           \
    if (!(cond)) {                               \
      pgs::set_error("%s: %s", __func__, msg);   \
      return PGS_ERR_INVALID;                    \
    }                                            \
  } while (0)

#define PGS_CHECK_LAUNCH()                                                        \
  do {                                                                            \
    cudaError_t e__ = cudaGetLastError();                                         \
    if (e__ != cudaSuccess) {                                                     \
      pgs::set_error("%s: CUDA error %s", __func__, cudaGetErrorString(e__));     \
      return PGS_ERR_CUDA;                                                        \
    }                                                                             \
  } while (0)

#define PGS_CUDA(call)                                                            \
  do {                                                                            \
    cudaError_t e__ = (call);                                                     \
    if (e__ != cudaSuccess) {                                                     \
      pgs::set_error("%s: %s -> %s", __func__, #call, cudaGetErrorString(e__));   \
      return PGS_ERR_CUDA;                                                        \
    }                                                                             \
  } while (0)

constexpr uint64_t kEmptyKey = ~0ull;
constexpr int kNumSM = 148;  // B200

// (b,x,y,z) -> 64-bit key, 16 bits per field, spatial fields biased so negatives sort/pack.
__host__ __device__ inline bool pack_key(int b, int x, int y, int z, uint64_t* key) {
  const unsigned ux = (unsigned)(x + 32768), uy = (unsigned)(y + 32768), uz = (unsigned)(z + 32768);
  const bool ok = ((unsigned)b < 65536u) & (ux < 65536u) & (uy < 65536u) & (uz < 65536u);
  *key = ((uint64_t)(unsigned)b << 48) | ((uint64_t)ux << 32) | ((uint64_t)uy << 16) | (uint64_t)uz;
  return ok;
}

// murmur3 fmix64
__host__ __device__ inline uint64_t hash64(uint64_t k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return k;
}

__host__ __device__ inline int floor_div(int a, int b) {  // b > 0
  int q = a / b;
  return (a % b != 0 && a < 0) ? q - 1 : q;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// tcgen05 conv: 1 = the lo*hi + hi*lo correction products run as ONE bf16 MMA per 8 channels (default), 0 = as two
// tf32 MMAs (PGS_TC_CORR=tf32); decides the layout of the arranged weights, so prep and kernel ask the same function
int tc_corr16();

// exclusive scan of int32 flags/counts (device-wide, 3 launches); out has n+1 entries,
// out[n] = total.  scratch: scan_scratch_bytes(n).
size_t scan_scratch_bytes(int64_t n);
int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, void* scratch, cudaStream_t s);

}  // namespace pgs
