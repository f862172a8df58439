// Proposal bookkeeping after clustering, on the device:
//
//   pgs_prop_gt_iou     IoU of every proposal with every ground-truth instance of its scene
//                       (replaces torch_points_kernels.instance_iou; reference call sites
//                        torch_points3d/core/losses/panoptic_losses.py:37,126 and every panoptic tracker)
//   pgs_prop_cross_nms  proposal x proposal intersections + greedy non-maximum suppression in descending score order
//                       (replaces the dense [n_prop, N] mask matmul and the numpy loop of
//                        torch_points3d/models/panoptic/structure_3heads.py:6-17,40-61)
//
// Proposals arrive in CSR form: flat point ids (int64) + offsets (int32 [n_prop + 1]).  Integer counting with atomics,
// one fp32 division per matrix entry; HBM/L2-bound gathers, nothing here is a GEMM.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace pgs {

constexpr int kPT = 256;

__device__ __forceinline__ int prop_of(const int32_t* __restrict__ offs, int n_prop, int64_t e) {
  int lo = 0, hi = n_prop;   // last p with offs[p] <= e
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if ((int64_t)offs[mid] <= e) lo = mid; else hi = mid;
  }
  return lo;
}

// one thread per proposal member: count it under (proposal, ground-truth instance of that point)
__global__ void __launch_bounds__(kPT) prop_gt_count_kernel(const int64_t* __restrict__ flat, const int32_t* __restrict__ offs,
                                                             int n_prop, int64_t n_flat, const int32_t* __restrict__ gt_id,
                                                             int total_gt, int32_t* __restrict__ inter) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_flat) return;
  const int g = gt_id[flat[e]];
  if (g < 0) return;
  atomicAdd(&inter[(int64_t)prop_of(offs, n_prop, e) * total_gt + g], 1);
}

__global__ void __launch_bounds__(kPT) prop_gt_iou_kernel(const int32_t* __restrict__ inter, const int32_t* __restrict__ offs,
                                                           int n_prop, int total_gt, const int32_t* __restrict__ gt_size,
                                                           const int32_t* __restrict__ gt_scene,
                                                           const int32_t* __restrict__ prop_scene, float* __restrict__ iou) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n_prop * total_gt) return;
  const int p = (int)(i / total_gt), g = (int)(i - (int64_t)p * total_gt);
  float v = 0.f;
  if (prop_scene[p] == gt_scene[g]) {   // a proposal only competes with the instances of its own scene
    const int it = inter[i];
    v = (float)it / (float)((offs[p + 1] - offs[p]) + gt_size[g] - it);
  }
  iou[i] = v;
}

// rows sorted by point id: every run of equal ids lists the proposals that contain the point; count all pairs of the run
__global__ void __launch_bounds__(kPT) prop_pair_count_kernel(const int64_t* __restrict__ spoint, const int32_t* __restrict__ sprop,
                                                               int64_t n_flat, int n_prop, int32_t* __restrict__ inter) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_flat) return;
  const int64_t pt = spoint[e];
  const int a = sprop[e];
  for (int64_t f = e + 1; f < n_flat && spoint[f] == pt; ++f) {
    const int b = sprop[f];
    atomicAdd(&inter[(int64_t)a * n_prop + b], 1);
    atomicAdd(&inter[(int64_t)b * n_prop + a], 1);
  }
}

// one CTA: walk the proposals in descending score order; a picked proposal suppresses every later one whose cross IoU
// with it exceeds the threshold (structure_3heads.py:6-17).  keep[i] = 1 for picked proposals.
__global__ void __launch_bounds__(1024) prop_nms_kernel(const int32_t* __restrict__ inter, const int32_t* __restrict__ offs,
                                                         const int32_t* __restrict__ rank_order, int n_prop, float threshold,
                                                         uint8_t* __restrict__ keep) {
  extern __shared__ uint8_t alive[];
  for (int i = threadIdx.x; i < n_prop; i += blockDim.x) {
    alive[i] = 1;
    keep[i] = 0;
  }
  __syncthreads();
  for (int r = 0; r < n_prop; ++r) {
    const int i = rank_order[r];
    if (alive[i]) {   // (uniform: every thread reads the same flag after the barrier)
      const int ni = offs[i + 1] - offs[i];
      for (int j = threadIdx.x; j < n_prop; j += blockDim.x) {
        if (j == i || !alive[j]) continue;
        const int it = inter[(int64_t)i * n_prop + j];
        const float iou = (float)it / (float)(ni + (offs[j + 1] - offs[j]) - it);
        if (iou > threshold) alive[j] = 0;
      }
      if (threadIdx.x == 0) keep[i] = 1;   // (alive[i] stays set: rank r is never visited again, and clearing it here
                                           //  would race with the other threads' read of the flag above)
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kPT) prop_expand_kernel(const int32_t* __restrict__ offs, int n_prop, int64_t n_flat,
                                                           int32_t* __restrict__ pid) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n_flat) pid[e] = prop_of(offs, n_prop, e);
}

static inline unsigned pblocks(int64_t n) { return (unsigned)((n + kPT - 1) / kPT); }

}  // namespace pgs

using namespace pgs;

extern "C" {

int pgs_prop_gt_iou(const int64_t* flat, const int32_t* offs, int32_t n_prop, int64_t n_flat, const int32_t* gt_id,
                    int32_t total_gt, const int32_t* gt_size, const int32_t* gt_scene, const int32_t* prop_scene,
                    int32_t* inter, float* iou, void* stream) {
  PGS_CHECK_ARG(n_prop >= 0 && total_gt >= 0 && n_flat >= 0, "negative size");
  if (n_prop == 0 || total_gt == 0) return PGS_OK;
  cudaStream_t s = (cudaStream_t)stream;
  PGS_CUDA(cudaMemsetAsync(inter, 0, sizeof(int32_t) * (size_t)n_prop * total_gt, s));
  if (n_flat) prop_gt_count_kernel<<<pblocks(n_flat), kPT, 0, s>>>(flat, offs, n_prop, n_flat, gt_id, total_gt, inter);
  prop_gt_iou_kernel<<<pblocks((int64_t)n_prop * total_gt), kPT, 0, s>>>(inter, offs, n_prop, total_gt, gt_size, gt_scene,
                                                                          prop_scene, iou);
  count_launch(2);
  PGS_CHECK_LAUNCH();
  return PGS_OK;
}

size_t pgs_prop_nms_scratch_bytes(int64_t n_flat) {
  size_t cb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cb, (const int64_t*)nullptr, (int64_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, (int)(n_flat < 1 ? 1 : n_flat));
  const size_t n = (size_t)(n_flat < 1 ? 1 : n_flat);
  return align_up(cb, 256) + align_up(8 * n, 256) + 2 * align_up(4 * n, 256);
}

int pgs_prop_cross_nms(const int64_t* flat, const int32_t* offs, int32_t n_prop, int64_t n_flat,
                       const int32_t* rank_order, float threshold, int32_t* inter, uint8_t* keep, void* scratch,
                       size_t scratch_bytes, void* stream) {
  PGS_CHECK_ARG(n_prop >= 0 && n_flat >= 0 && n_flat < (1ll << 31), "bad sizes");
  PGS_CHECK_ARG(n_prop <= 48 * 1024, "more than 49152 proposals in one call");
  PGS_CHECK_ARG(scratch_bytes >= pgs_prop_nms_scratch_bytes(n_flat), "scratch too small");
  if (n_prop == 0) return PGS_OK;
  cudaStream_t s = (cudaStream_t)stream;
  char* p = (char*)scratch;
  size_t cb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cb, (const int64_t*)nullptr, (int64_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, (int)(n_flat < 1 ? 1 : n_flat));
  void* cub_ws = p;
  p += align_up(cb, 256);
  const size_t n = (size_t)(n_flat < 1 ? 1 : n_flat);
  int64_t* spoint = (int64_t*)p;
  p += align_up(8 * n, 256);
  int32_t* pid = (int32_t*)p;
  p += align_up(4 * n, 256);
  int32_t* sprop = (int32_t*)p;
  PGS_CUDA(cudaMemsetAsync(inter, 0, sizeof(int32_t) * (size_t)n_prop * n_prop, s));
  if (n_flat) {
    prop_expand_kernel<<<pblocks(n_flat), kPT, 0, s>>>(offs, n_prop, n_flat, pid);
    PGS_CUDA(cub::DeviceRadixSort::SortPairs(cub_ws, cb, flat, spoint, pid, sprop, (int)n_flat, 0, 64, s));
    prop_pair_count_kernel<<<pblocks(n_flat), kPT, 0, s>>>(spoint, sprop, n_flat, n_prop, inter);
    count_launch(2);
  }
  prop_nms_kernel<<<1, 1024, (size_t)n_prop, s>>>(inter, offs, rank_order, n_prop, threshold, keep);
  count_launch();
  PGS_CHECK_LAUNCH();
  return PGS_OK;
}

}  // extern "C"
