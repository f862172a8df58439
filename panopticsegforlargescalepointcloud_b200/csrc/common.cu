// Error reporting, launch accounting and the device-wide exclusive scan used by the
// coordinate-map / rulebook / clustering kernels.
#include <stdarg.h>
#include <string.h>

#include <cstdlib>

#include "common.cuh"

namespace pgs {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ------------------------------------------------------------------------------------------
// exclusive scan: 256 threads x 8 items per block
// ------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += t;
  }
  return v;
}

// block-wide exclusive scan of one value per thread; returns exclusive prefix, *total = block sum
template <int THREADS>
__device__ __forceinline__ int block_excl_scan(int v, int* total, int* smem /* THREADS/32 + 1 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = warp_incl_scan(v, lane);
  if (lane == 31) smem[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = (lane < THREADS / 32) ? smem[lane] : 0;
    int wi = warp_incl_scan(w, lane);
    if (lane < THREADS / 32) smem[lane] = wi - w;  // exclusive warp offsets
    if (lane == 31) smem[THREADS / 32] = wi;
  }
  __syncthreads();
  int res = incl - v + smem[warp];
  *total = smem[THREADS / 32];
  __syncthreads();
  return res;
}

__global__ void __launch_bounds__(kScanThreads) scan_block_sums_kernel(const int32_t* __restrict__ in,
                                                                        int64_t n,
                                                                        int32_t* __restrict__ sums) {
  __shared__ int sm[kScanThreads / 32 + 1];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    int64_t j = base + i;
    if (j < n) s += in[j];
  }
  int total;
  block_excl_scan<kScanThreads>(s, &total, sm);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) scan_sums_kernel(int32_t* __restrict__ sums, int nb) {
  __shared__ int sm[1024 / 32 + 1];
  int carry = 0;
  for (int base = 0; base < nb; base += 1024) {
    int i = base + threadIdx.x;
    int v = (i < nb) ? sums[i] : 0;
    int total;
    int ex = block_excl_scan<1024>(v, &total, sm);
    if (i < nb) sums[i] = ex + carry;
    carry += total;
  }
  if (threadIdx.x == 0) sums[nb] = carry;
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const int32_t* __restrict__ in,
                                                                   int64_t n,
                                                                   const int32_t* __restrict__ sums,
                                                                   int nb, int32_t* __restrict__ out) {
  __shared__ int sm[kScanThreads / 32 + 1];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int v[kScanItems];
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    int64_t j = base + i;
    v[i] = (j < n) ? in[j] : 0;
    s += v[i];
  }
  int total;
  int ex = block_excl_scan<kScanThreads>(s, &total, sm) + sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    int64_t j = base + i;
    if (j < n) out[j] = ex;
    ex += v[i];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = sums[nb];
}

size_t scan_scratch_bytes(int64_t n) {
  int64_t nb = (n + kScanTile - 1) / kScanTile;
  if (nb < 1) nb = 1;
  return align_up((size_t)(nb + 1) * sizeof(int32_t), 256);
}

int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, void* scratch, cudaStream_t s) {
  int32_t* sums = (int32_t*)scratch;
  int nb = (int)((n + kScanTile - 1) / kScanTile);
  if (nb < 1) nb = 1;
  scan_block_sums_kernel<<<nb, kScanThreads, 0, s>>>(in, n, sums);
  scan_sums_kernel<<<1, 1024, 0, s>>>(sums, nb);
  scan_apply_kernel<<<nb, kScanThreads, 0, s>>>(in, n, sums, nb, out);
  count_launch(3);
  PGS_CHECK_LAUNCH();
  return PGS_OK;
}

int tc_corr16() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PGS_TC_CORR");
    v = (e && e[0] == 't') ? 0 : 1;
  }
  return v;
}

}  // namespace pgs

extern "C" {
int pgs_version(void) { return 100; }
const char* pgs_last_error(void) { return pgs::g_err; }
int64_t pgs_launch_count(void) { return pgs::g_launches.load(); }
}
