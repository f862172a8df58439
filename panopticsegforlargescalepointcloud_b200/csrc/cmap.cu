// Coordinate hash map, strided maps and kernel maps (rulebooks).
//
// Replaces MinkowskiEngine's CoordinateManager on the reference hot path
// (reference call sites: torch_points3d/applications/minkowski.py:121-122,
//  torch_points3d/modules/MinkowskiEngine/api_modules.py:26-55,244-270,293).
//
// All kernels here are integer HBM/L2-bound work: one thread per row (or per row x offset),
// coalesced row reads, open-addressing probes that hit L2.  Grids are sized for whole waves on
// 148 SMs; nothing is reshaped into a GEMM.
#include "common.cuh"

namespace pgs {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) table_clear_kernel(uint64_t* __restrict__ keys,
                                                                int32_t* __restrict__ vals,
                                                                int64_t cap) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap;
       i += (int64_t)gridDim.x * blockDim.x) {
    keys[i] = kEmptyKey;
    vals[i] = 0x7fffffff;
  }
}

// insert key(floor(c/ts)*ts) for every row; table value = smallest row that carries the key
__global__ void __launch_bounds__(kThreads) cmap_insert_kernel(
    const int4* __restrict__ coords, int64_t n, int ts, uint64_t* __restrict__ keys,
    int32_t* __restrict__ vals, uint64_t mask, int32_t* __restrict__ slot_of_row,
    uint32_t* __restrict__ status) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int4 c = coords[i];
  if (ts > 1) {
    c.y = floor_div(c.y, ts) * ts;
    c.z = floor_div(c.z, ts) * ts;
    c.w = floor_div(c.w, ts) * ts;
  }
  uint64_t key;
  if (!pack_key(c.x, c.y, c.z, c.w, &key)) {
    atomicOr(status, PGS_STATUS_COORD_RANGE);
    slot_of_row[i] = -1;
    return;
  }
  uint64_t slot = hash64(key) & mask;
  for (uint64_t probe = 0; probe <= mask; ++probe) {
    unsigned long long prev =
        atomicCAS((unsigned long long*)&keys[slot], (unsigned long long)kEmptyKey, (unsigned long long)key);
    if (prev == kEmptyKey || prev == key) {
      atomicMin(&vals[slot], (int32_t)i);
      slot_of_row[i] = (int32_t)slot;
      return;
    }
    slot = (slot + 1) & mask;
  }
  atomicOr(status, PGS_STATUS_TABLE_FULL);
  slot_of_row[i] = -1;
}

__global__ void __launch_bounds__(kThreads) cmap_flag_kernel(const int32_t* __restrict__ vals,
                                                              const int32_t* __restrict__ slot_of_row,
                                                              int64_t n, int32_t* __restrict__ flags) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = slot_of_row[i];
  flags[i] = (s >= 0 && vals[s] == (int32_t)i) ? 1 : 0;
}

// in2out[i] = pos[first row of i's key]; first rows emit the (quantised) coordinate
__global__ void __launch_bounds__(kThreads) cmap_emit_kernel(
    const int4* __restrict__ coords, int64_t n, int ts, const int32_t* __restrict__ vals,
    const int32_t* __restrict__ slot_of_row, const int32_t* __restrict__ pos,
    int4* __restrict__ out_coords, int32_t* __restrict__ in2out, int32_t* __restrict__ n_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *n_out = pos[n];
  if (i >= n) return;
  const int s = slot_of_row[i];
  if (s < 0) {
    in2out[i] = -1;
    return;
  }
  const int first = vals[s];
  const int o = pos[first];
  in2out[i] = o;
  if (first == (int32_t)i) {
    int4 c = coords[i];
    if (ts > 1) {
      c.y = floor_div(c.y, ts) * ts;
      c.z = floor_div(c.z, ts) * ts;
      c.w = floor_div(c.w, ts) * ts;
    }
    out_coords[o] = c;
  }
}

// after every reader of the first-row ids is done: table value := row id in the NEW map
__global__ void __launch_bounds__(kThreads) cmap_relabel_kernel(const int32_t* __restrict__ flags,
                                                                 const int32_t* __restrict__ slot_of_row,
                                                                 const int32_t* __restrict__ pos,
                                                                 int64_t n, int32_t* __restrict__ vals) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (flags[i]) vals[slot_of_row[i]] = pos[i];
}

__device__ __forceinline__ int table_lookup(const uint64_t* __restrict__ keys,
                                            const int32_t* __restrict__ vals, uint64_t mask,
                                            uint64_t key) {
  uint64_t slot = hash64(key) & mask;
  for (uint64_t probe = 0; probe <= mask; ++probe) {
    const uint64_t k = __ldg(&keys[slot]);
    if (k == key) return __ldg(&vals[slot]);
    if (k == kEmptyKey) return -1;
    slot = (slot + 1) & mask;
  }
  return -1;
}

// one thread per (query row, kernel offset); k-major output so a warp writes 128 contiguous bytes.
__global__ void __launch_bounds__(kThreads) kmap_build_kernel(
    const int4* __restrict__ q, int64_t n_q, const uint64_t* __restrict__ keys,
    const int32_t* __restrict__ vals, uint64_t mask, int step, int ksize, int32_t* __restrict__ nbr) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (r >= n_q) return;
  const int4 c = q[r];
  const int half = (ksize & 1) ? ksize / 2 : 0;  // even kernels are non-centred {0..k-1} (ME)
  const int dx = (k % ksize) - half, dy = ((k / ksize) % ksize) - half, dz = (k / (ksize * ksize)) - half;
  uint64_t key;
  int res = -1;
  if (pack_key(c.x, c.y + dx * step, c.z + dy * step, c.w + dz * step, &key))
    res = table_lookup(keys, vals, mask, key);
  nbr[(int64_t)k * n_q + r] = res;
}

__global__ void __launch_bounds__(kThreads) nonneg_flag_kernel(const int32_t* __restrict__ v, int64_t n,
                                                                int32_t* __restrict__ flags) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = v[i] >= 0;
}

__global__ void __launch_bounds__(kThreads) pairs_emit_kernel(
    const int32_t* __restrict__ nbr, const int32_t* __restrict__ pos, int64_t n_q, int K,
    int32_t* __restrict__ in_idx, int32_t* __restrict__ out_idx, int32_t* __restrict__ offs) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = n_q * K;
  if (i <= K) offs[i] = pos[i < K ? i * n_q : total];
  if (i >= total) return;
  const int v = nbr[i];
  if (v >= 0) {
    const int p = pos[i];
    in_idx[p] = v;
    out_idx[p] = (int32_t)(i % n_q);
  }
}


// ---------------------------------------------------------------------------------------------
// Occupancy-sorted gather tables.  The tensor-core conv kernels skip every (16-row tile, kernel offset) pair whose
// rows have no neighbour at that offset; with rows in arbitrary order almost no pair is empty (each row has 2..13
// of 27 neighbours but a 16-row union has ~26).  Sorting the rows of a table by their K-bit occupancy mask (rarest
// offset = most significant bit) makes the rows of a tile share their empty offsets: the union drops to ~11 of 27
// for the stride-1 maps and to ~2.4 for the transposed (up-sampling) maps, whose masks are parity classes.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) kmap_masks_kernel(const int32_t* __restrict__ nbr, int64_t n_q, int K,
                                                              int ksize, int32_t* __restrict__ masks) {
  // bit position of offset k: offsets ordered by (number of non-zero components, k) -- centre = bit 0, then the 6
  // faces, the 12 edges, the 8 corners (the rarest neighbours) in the most significant bits; measured as good as
  // ordering by the actual pair counts (profiles/r1_mask_order.txt) without a counting pass
  __shared__ int bitpos[32];
  if (threadIdx.x < K) {
    const int half = (ksize & 1) ? ksize / 2 : 0;
    auto nz = [&](int k) {
      const int ix = k % ksize - half, iy = (k / ksize) % ksize - half, iz = k / (ksize * ksize) - half;
      return (ix != 0) + (iy != 0) + (iz != 0);
    };
    const int k = threadIdx.x, mine = nz(k);
    int pos = 0;
    for (int j = 0; j < K; ++j) {
      const int o = nz(j);
      pos += (o < mine) || (o == mine && j < k);
    }
    bitpos[k] = pos;
  }
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_q) return;
  uint32_t m = 0;
  for (int k = 0; k < K; ++k)
    if (nbr[(int64_t)k * n_q + i] >= 0) m |= 1u << bitpos[k];
  masks[i] = (int32_t)m;
}

__global__ void __launch_bounds__(kThreads) kmap_permute_kernel(const int32_t* __restrict__ nbr, int64_t n_q,
                                                                const int32_t* __restrict__ order,
                                                                int32_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_q) return;
  const int k = blockIdx.y;
  out[(int64_t)k * n_q + i] = nbr[(int64_t)k * n_q + order[i]];
}

static inline int grid_for(int64_t n) { return (int)((n + kThreads - 1) / kThreads); }

}  // namespace pgs

using namespace pgs;

extern "C" {

int64_t pgs_cmap_capacity(int64_t n) {
  int64_t cap = 1024;
  while (cap < 2 * n) cap <<= 1;
  return cap;
}

size_t pgs_cmap_build_scratch_bytes(int64_t n) {
  // slot_of_row[n] + flags[n] + pos[n+1] + scan scratch
  return align_up((size_t)n * 4, 256) * 2 + align_up((size_t)(n + 1) * 4, 256) + scan_scratch_bytes(n);
}

int pgs_cmap_build(const int32_t* coords, int64_t n, int32_t ts, uint64_t* tkeys, int32_t* tvals,
                   int64_t cap, int32_t* out_coords, int32_t* in2out, int32_t* n_out,
                   uint32_t* status, void* scratch, size_t scratch_bytes, void* stream) {
  PGS_CHECK_ARG(n >= 0 && n < (1ll << 31), "row count out of range");
  PGS_CHECK_ARG(ts >= 1, "tensor stride must be >= 1");
  PGS_CHECK_ARG(cap >= 2 * n && (cap & (cap - 1)) == 0, "capacity must be a power of two >= 2n");
  PGS_CHECK_ARG(scratch_bytes >= pgs_cmap_build_scratch_bytes(n), "scratch too small");
  cudaStream_t s = (cudaStream_t)stream;
  char* p = (char*)scratch;
  int32_t* slot_of_row = (int32_t*)p;
  p += align_up((size_t)n * 4, 256);
  int32_t* flags = (int32_t*)p;
  p += align_up((size_t)n * 4, 256);
  int32_t* pos = (int32_t*)p;
  p += align_up((size_t)(n + 1) * 4, 256);
  void* scan_ws = p;

  int cgrid = (int)((cap + kThreads - 1) / kThreads);
  if (cgrid > kNumSM * 16) cgrid = kNumSM * 16;
  table_clear_kernel<<<cgrid, kThreads, 0, s>>>(tkeys, tvals, cap);
  count_launch();
  if (n == 0) {
    PGS_CUDA(cudaMemsetAsync(n_out, 0, sizeof(int32_t), s));
    PGS_CHECK_LAUNCH();
    return PGS_OK;
  }
  const int g = grid_for(n);
  cmap_insert_kernel<<<g, kThreads, 0, s>>>((const int4*)coords, n, ts, tkeys, tvals, (uint64_t)(cap - 1),
                                            slot_of_row, status);
  cmap_flag_kernel<<<g, kThreads, 0, s>>>(tvals, slot_of_row, n, flags);
  count_launch(2);
  int rc = exclusive_scan_i32(flags, pos, n, scan_ws, s);
  if (rc) return rc;
  cmap_emit_kernel<<<g, kThreads, 0, s>>>((const int4*)coords, n, ts, tvals, slot_of_row, pos,
                                          (int4*)out_coords, in2out, n_out);
  cmap_relabel_kernel<<<g, kThreads, 0, s>>>(flags, slot_of_row, pos, n, tvals);
  count_launch(2);
  PGS_CHECK_LAUNCH();
  return PGS_OK;
}

int pgs_kmap_build(const int32_t* q_coords, int64_t n_q, const uint64_t* tkeys, const int32_t* tvals,
                   int64_t cap, int32_t step, int32_t sign, int32_t ksize, int32_t* nbr, void* stream) {
  PGS_CHECK_ARG(ksize >= 1 && ksize <= 5, "kernel size must be in 1..5");
  PGS_CHECK_ARG(sign == 1 || sign == -1, "sign must be +1 or -1");
  PGS_CHECK_ARG((cap & (cap - 1)) == 0, "capacity must be a power of two");
  if (n_q == 0) return PGS_OK;
  const int K = ksize * ksize * ksize;
  dim3 grid(grid_for(n_q), K);
  kmap_build_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>((const int4*)q_coords, n_q, tkeys, tvals,
                                                                 (uint64_t)(cap - 1), step * sign, ksize, nbr);
  count_launch();
  PGS_CHECK_LAUNCH();
  return PGS_OK;
}

size_t pgs_kmap_pairs_scratch_bytes(int64_t n_q, int32_t K) {
  const int64_t t = n_q * K;
  return align_up((size_t)t * 4, 256) + align_up((size_t)(t + 1) * 4, 256) + scan_scratch_bytes(t);
}

int pgs_kmap_pairs(const int32_t* nbr, int64_t n_q, int32_t K, int32_t* in_idx, int32_t* out_idx,
                   int32_t* offs, void* scratch, size_t scratch_bytes, void* stream) {
  PGS_CHECK_ARG(scratch_bytes >= pgs_kmap_pairs_scratch_bytes(n_q, K), "scratch too small");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t t = n_q * K;
  PGS_CHECK_ARG(t < (1ll << 31), "K * n_q must fit int32");
  if (t == 0) {
    PGS_CUDA(cudaMemsetAsync(offs, 0, sizeof(int32_t) * (K + 1), s));
    return PGS_OK;
  }
  char* p = (char*)scratch;
  int32_t* flags = (int32_t*)p;
  p += align_up((size_t)t * 4, 256);
  int32_t* pos = (int32_t*)p;
  p += align_up((size_t)(t + 1) * 4, 256);
  nonneg_flag_kernel<<<grid_for(t), kThreads, 0, s>>>(nbr, t, flags);
  count_launch();
  int rc = exclusive_scan_i32(flags, pos, t, p, s);
  if (rc) return rc;
  pairs_emit_kernel<<<grid_for(t + 1), kThreads, 0, s>>>(nbr, pos, n_q, K, in_idx, out_idx, offs);
  count_launch();
  PGS_CHECK_LAUNCH();
  return PGS_OK;
}

int pgs_kmap_row_masks(const int32_t* nbr, int64_t n_q, int32_t K, int32_t* masks, void* stream) {
  PGS_CHECK_ARG(K >= 1 && K <= 31, "kernel volume must be in 1..31");
  if (n_q == 0) return PGS_OK;
  int ksize = 1;
  while (ksize * ksize * ksize < K) ++ksize;
  PGS_CHECK_ARG(ksize * ksize * ksize == K, "kernel volume must be a cube");
  kmap_masks_kernel<<<grid_for(n_q), kThreads, 0, (cudaStream_t)stream>>>(nbr, n_q, K, ksize, masks);
  count_launch();
  PGS_CHECK_LAUNCH();
  return PGS_OK;
}

int pgs_kmap_permute(const int32_t* nbr, int64_t n_q, int32_t K, const int32_t* order, int32_t* nbr_sorted,
                     void* stream) {
  if (n_q == 0) return PGS_OK;
  kmap_permute_kernel<<<dim3(grid_for(n_q), K), kThreads, 0, (cudaStream_t)stream>>>(nbr, n_q, order, nbr_sorted);
  count_launch();
  PGS_CHECK_LAUNCH();
  return PGS_OK;
}

}  // extern "C"
