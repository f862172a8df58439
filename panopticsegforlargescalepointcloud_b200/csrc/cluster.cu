// Radius ball query on a voxel-hash grid and region growing by min-ancestor label propagation.
//
// Replaces torch-points-kernels 0.7.0 `ball_query(mode="PARTIAL_DENSE")` + `region_grow`
// (un-vendored dependency of the reference; call sites:
//  torch_points3d/models/panoptic/PointGroup3heads.py:166-174,185-202,296-304,340-357,
//  torch_points3d/models/panoptic/pointgroup.py:141-149,160-177,
//  torch_points3d/core/spatial_ops/neighbour_finder.py:35-37,164).
//
// Definition that is reproduced bit-exactly: for query q the neighbour list is the
// FIRST `nsample` support points, in ascending index order, of the same group (semantic class x scene)
// with d2 = fma(dz,dz, fma(dy,dy, dx*dx)) <= r*r in fp32.  The upstream kernel finds them with an
// O(n^2) scan per query; here points are bucketed in cells of edge >= r (cell key sorted, points inside
// a cell in ascending index), one warp owns one query, lanes 0..26 own the 27 neighbouring cells and the
// warp walks all 27 ascending lists in lock step by index windows, so the scan stops after `nsample`
// hits exactly like the reference scan does.  Region growing is the fixpoint of
// label[j] = min(label[j], label[q]) over directed edges q -> j plus pointer jumping (SURVEY App. C
// theorem: sequential seeded BFS == min-ancestor labelling).
//
// Integer / compare work, L2- and HBM-bound; nothing here is a GEMM.
#include <cub/device/device_radix_sort.cuh>

#include <cstdlib>

#include "common.cuh"

namespace pgs {

constexpr int kCT = 256;
constexpr uint64_t kInvalidKey = ~0ull;
constexpr int kWin = 8;  // elements a lane looks ahead per round

__device__ __forceinline__ bool cell_of(float x, float y, float z, double inv_cell, int* cx, int* cy, int* cz) {
  // double: p / cell is exact enough that |dx| <= r  =>  cells differ by at most one
  const double fx = floor((double)x * inv_cell), fy = floor((double)y * inv_cell), fz = floor((double)z * inv_cell);
  const bool ok = fabs(fx) < 32766.0 && fabs(fy) < 32766.0 && fabs(fz) < 32766.0;
  *cx = (int)fx + 32768;
  *cy = (int)fy + 32768;
  *cz = (int)fz + 32768;
  return ok;
}

__device__ __forceinline__ uint64_t cell_key(int gid, int cx, int cy, int cz) {
  return ((uint64_t)(unsigned)gid << 48) | ((uint64_t)(unsigned)cz << 32) | ((uint64_t)(unsigned)cy << 16) |
         (uint64_t)(unsigned)cx;
}

__global__ void __launch_bounds__(kCT) grid_key_kernel(const float* __restrict__ pos, const int32_t* __restrict__ gid,
                                                        int64_t n, double inv_cell, uint64_t* __restrict__ keys,
                                                        int32_t* __restrict__ ids, uint32_t* __restrict__ status) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  ids[i] = (int32_t)i;
  const int g = gid[i];
  if (g < 0) {
    keys[i] = kInvalidKey;
    return;
  }
  const float x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
  int cx, cy, cz;
  const bool finite = (x == x) && (y == y) && (z == z);
  if (!finite || !cell_of(x, y, z, inv_cell, &cx, &cy, &cz) || g >= 32767) {
    atomicOr(status, PGS_STATUS_COORD_RANGE);
    keys[i] = kInvalidKey;
    return;
  }
  keys[i] = cell_key(g, cx, cy, cz);
}

// sorted order -> packed (x, y, z, idx) rows, cell-head flags
__global__ void __launch_bounds__(kCT) grid_pack_kernel(const float* __restrict__ pos,
                                                         const uint64_t* __restrict__ skeys,
                                                         const int32_t* __restrict__ sids, int64_t n,
                                                         float4* __restrict__ spos, int32_t* __restrict__ heads) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t k = skeys[i];
  const int id = sids[i];
  spos[i] = make_float4(pos[3 * (int64_t)id], pos[3 * (int64_t)id + 1], pos[3 * (int64_t)id + 2], __int_as_float(id));
  heads[i] = (k != kInvalidKey) && (i == 0 || skeys[i - 1] != k);
}

// queries that are not the support set: (x, y, z, row) + cell key, in input order
__global__ void __launch_bounds__(kCT) query_pack_kernel(const float* __restrict__ pos, const int32_t* __restrict__ gid,
                                                          int64_t n, double inv_cell, float4* __restrict__ qpos,
                                                          uint64_t* __restrict__ qkeys) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
  qpos[i] = make_float4(x, y, z, __int_as_float((int)i));
  const int g = gid[i];
  int cx, cy, cz;
  const bool finite = (x == x) && (y == y) && (z == z);
  // a query outside the packable range has no support point within r either (supports are range-checked)
  qkeys[i] = (g >= 0 && g < 32767 && finite && cell_of(x, y, z, inv_cell, &cx, &cy, &cz)) ? cell_key(g, cx, cy, cz)
                                                                                         : kInvalidKey;
}

// one thread per sorted row: cell heads register (key -> cell ordinal) and write cell_start[ordinal];
// the last valid row closes the list with cell_start[n_cells] = n_valid.
__global__ void __launch_bounds__(kCT) grid_cells_kernel(const uint64_t* __restrict__ skeys,
                                                          const int32_t* __restrict__ heads,
                                                          const int32_t* __restrict__ ord, int64_t n,
                                                          uint64_t* __restrict__ tkeys, int32_t* __restrict__ tvals,
                                                          uint64_t mask, int32_t* __restrict__ cell_start,
                                                          int32_t* __restrict__ meta /* [n_valid, n_cells] */) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t k = skeys[i];
  if (k == kInvalidKey) return;
  if (i == n - 1 || skeys[i + 1] == kInvalidKey) {
    const int ncell = ord[i] + heads[i];
    cell_start[ncell] = (int32_t)(i + 1);
    meta[0] = (int32_t)(i + 1);
    meta[1] = ncell;
  }
  if (!heads[i]) return;
  const int c = ord[i];
  cell_start[c] = (int32_t)i;
  uint64_t slot = hash64(k) & mask;
  for (;;) {
    unsigned long long prev = atomicCAS((unsigned long long*)&tkeys[slot], (unsigned long long)kEmptyKey,
                                        (unsigned long long)k);
    if (prev == kEmptyKey) {
      tvals[slot] = c;
      return;
    }
    slot = (slot + 1) & mask;
  }
}

__device__ __forceinline__ int warp_min_i(int v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, d));
  return v;
}

__device__ __forceinline__ int warp_incl_scan_i(int v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += t;
  }
  return v;
}

// One warp per query (queries taken in cell-sorted order so that neighbouring warps touch the same
// cells).  nbr row layout: int32 [n, nsample], first cnt[q] entries valid, order unspecified unless
// the row overflowed (then it holds exactly the nsample smallest indices).
__global__ void __launch_bounds__(kCT) ball_query_kernel(
    const float4* __restrict__ spos, const float4* __restrict__ qpos, const uint64_t* __restrict__ qkeys,
    int64_t n_q, const uint64_t* __restrict__ tkeys, const int32_t* __restrict__ tvals, uint64_t mask,
    const int32_t* __restrict__ cell_start, float r2, int nsample, int32_t* __restrict__ nbr,
    int32_t* __restrict__ cnt, float* __restrict__ dist /* nullable, same layout as nbr */) {
  __shared__ int s_idx[kCT / 32][27 * kWin];
  __shared__ float s_d2[kCT / 32][27 * kWin];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= n_q) return;
  const uint64_t qk = qkeys[w];
  if (qk == kInvalidKey) return;  // ignored / out-of-range query: cnt stays 0
  const float4 qp = qpos[w];
  const int q = __float_as_int(qp.w);

  int cur = 0, end = 0;
  if (lane < 27) {
    const int dx = lane % 3 - 1, dy = (lane / 3) % 3 - 1, dz = lane / 9 - 1;
    // 16-bit fields never wrap: cells are confined to [2, 65534]
    const uint64_t k = qk + (uint64_t)(int64_t)dx + ((uint64_t)(int64_t)dy << 16) + ((uint64_t)(int64_t)dz << 32);
    uint64_t slot = hash64(k) & mask;
    for (;;) {
      const uint64_t tk = __ldg(&tkeys[slot]);
      if (tk == k) {
        const int c = __ldg(&tvals[slot]);
        cur = __ldg(&cell_start[c]);
        end = __ldg(&cell_start[c + 1]);
        break;
      }
      if (tk == kEmptyKey) break;
      slot = (slot + 1) & mask;
    }
  }

  int32_t* row = nbr + (int64_t)q * nsample;
  float* drow = dist ? dist + (int64_t)q * nsample : nullptr;
  int hits = 0;
  for (;;) {
    const int rem = end - cur;
    const int look = rem < kWin ? rem : kWin;
    int prop = 0x7fffffff;
    if (look > 0) prop = __float_as_int(__ldg(&spos[cur + look - 1]).w);
    const int T = warp_min_i(prop);
    if (T == 0x7fffffff) break;
    int c_idx[kWin];
    float c_d2[kWin];
    unsigned hm = 0;  // bit t: candidate t of my window is a hit
    int adv = 0;
#pragma unroll
    for (int t = 0; t < kWin; ++t) {
      c_idx[t] = 0;
      c_d2[t] = 0.f;
      if (t < look) {
        const float4 p = __ldg(&spos[cur + t]);
        const int id = __float_as_int(p.w);
        if (id <= T) {  // the window is ascending: everything <= T is consumed this round
          adv = t + 1;
          const float dx = p.x - qp.x, dy = p.y - qp.y, dz = p.z - qp.z;
          const float d2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
          c_idx[t] = id;
          c_d2[t] = d2;
          if (d2 <= r2) hm |= 1u << t;
        }
      }
    }
    cur += adv;
    const int nh = __popc(hm);
    const int incl = warp_incl_scan_i(nh, lane);
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    const int off = incl - nh;
    if (hits + total <= nsample) {
#pragma unroll
      for (int t = 0; t < kWin; ++t)
        if (hm & (1u << t)) {
          const int o = hits + off + __popc(hm & ((1u << t) - 1u));
          row[o] = c_idx[t];
          if (drow) drow[o] = c_d2[t];
        }
      hits += total;
      if (hits == nsample) break;
    } else {
      // overflow inside this round: keep the (nsample - hits) smallest indices of the round
      const int need = nsample - hits;
#pragma unroll
      for (int t = 0; t < kWin; ++t)
        if (hm & (1u << t)) {
          const int o = off + __popc(hm & ((1u << t) - 1u));
          s_idx[wib][o] = c_idx[t];
          s_d2[wib][o] = c_d2[t];
        }
      __syncwarp();
      for (int e = lane; e < total; e += 32) {
        const int mine = s_idx[wib][e];
        int rank = 0;
        for (int j = 0; j < total; ++j) rank += s_idx[wib][j] < mine;
        if (rank < need) {
          row[hits + rank] = mine;
          if (drow) drow[hits + rank] = s_d2[wib][e];
        }
      }
      hits = nsample;
      break;
    }
  }
  if (lane == 0) cnt[q] = hits;
}

// reference layout: idx int64 [n, nsample] ascending index, -1 padded; dist2 fp32, -1 padded.
// One warp per row; rows are short (<= nsample), rank sort through shared memory in chunks.
__global__ void __launch_bounds__(kCT) ball_query_export_kernel(const int32_t* __restrict__ nbr,
                                                                 const float* __restrict__ dist,
                                                                 const int32_t* __restrict__ cnt, int64_t n,
                                                                 int nsample, int64_t* __restrict__ idx_out,
                                                                 float* __restrict__ dist_out) {
  const int lane = threadIdx.x & 31;
  const int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= n) return;
  const int c = cnt[q];
  const int32_t* row = nbr + q * nsample;
  const float* drow = dist + q * nsample;
  for (int e = lane; e < c; e += 32) {
    const int mine = row[e];
    int rank = 0;
    for (int j = 0; j < c; ++j) rank += row[j] < mine;
    idx_out[q * nsample + rank] = mine;
    dist_out[q * nsample + rank] = drow[e];
  }
  for (int e = c + lane; e < nsample; e += 32) {
    idx_out[q * nsample + e] = -1;
    dist_out[q * nsample + e] = -1.f;
  }
}

// ------------------------------------------------------------------------------------------
// region growing: min-ancestor labels
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCT) rg_init_kernel(const int32_t* __restrict__ gid, int64_t n,
                                                       int32_t* __restrict__ label) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) label[i] = gid[i] >= 0 ? (int32_t)i : -1;
}

// LANES lanes per source row (rows hold ~10-30 entries; a whole warp per row leaves most lanes idle and makes the launch
// 6.4 M threads for 200 k points): push my (freshest) label along my out-edges.  PGS_RG_LANES selects 8 (default) / 32.
template <int LANES>
__global__ void __launch_bounds__(kCT) rg_push_kernel(const int32_t* __restrict__ nbr, const int32_t* __restrict__ cnt,
                                                       const int32_t* __restrict__ gid, int64_t n, int nsample,
                                                       int32_t* __restrict__ label, int32_t* __restrict__ pushed,
                                                       int32_t* __restrict__ changed) {
  const int lane = threadIdx.x & (LANES - 1);
  const int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  if (q >= n || gid[q] < 0) return;
  // the LANES lanes of a row agree on ONE snapshot of the row's label (lane 0 reads, the group shuffles): they all push the
  // same value and take the same frontier decision
  const unsigned gmask = LANES == 32 ? 0xffffffffu : (((1u << LANES) - 1u) << ((threadIdx.x & 31) / LANES * LANES));
  int mine = 0, prev = -1;
  if (lane == 0) {
    mine = *(volatile int32_t*)&label[q];
    if (pushed) prev = pushed[q];
  }
  mine = __shfl_sync(gmask, mine, 0, LANES);
  prev = __shfl_sync(gmask, prev, 0, LANES);
  // frontier: a row whose label has not moved since it last pushed has nothing new to tell its neighbours (atomicMin is
  // monotone, so label[j] <= pushed[q] already holds for every out-edge); later sweeps only touch the rows that changed
  if (pushed) {
    if (prev == mine) return;
    if (lane == 0) pushed[q] = mine;
  }
  const int c = cnt[q];
  const int32_t* row = nbr + q * nsample;
  bool any = false;
  for (int e = lane; e < c; e += LANES) {
    const int j = row[e];
    if (*(volatile int32_t*)&label[j] > mine) {
      atomicMin(&label[j], mine);
      any = true;
    }
  }
  // one flag for the whole grid: read before write, so that after the first few writers nobody stores any more
  if (any && *(volatile int32_t*)changed == 0) *changed = 1;
}

// pointer jumping: label[v] = label[label[v]] until stable (valid because "reaches" is transitive)
__global__ void __launch_bounds__(kCT) rg_jump_kernel(int64_t n, int32_t* __restrict__ label,
                                                       int32_t* __restrict__ changed) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int l = label[i];
  if (l < 0) return;
  int r = l;
  for (;;) {
    const int p = *(volatile int32_t*)&label[r];
    if (p >= r) break;
    r = p;
  }
  if (r < l) {
    atomicMin(&label[i], r);
    *changed = 1;
  }
}

static inline int grid_for_c(int64_t n) { return (int)((n + kCT - 1) / kCT); }

struct GridLayout {
  uint64_t *keys, *skeys;
  int32_t *ids, *sids, *heads, *ord;
  void* scan_ws;
  void* cub_ws;
  size_t cub_bytes;
  size_t total;
};

static GridLayout grid_layout(int64_t n, void* base) {
  GridLayout L;
  char* p = (char*)base;
  auto take = [&](size_t bytes) {
    char* r = p;
    p += align_up(bytes ? bytes : 1, 256);
    return r;
  };
  L.keys = (uint64_t*)take((size_t)n * 8);
  L.skeys = (uint64_t*)take((size_t)n * 8);
  L.ids = (int32_t*)take((size_t)n * 4);
  L.sids = (int32_t*)take((size_t)n * 4);
  L.heads = (int32_t*)take((size_t)n * 4);
  L.ord = (int32_t*)take((size_t)(n + 1) * 4);
  L.scan_ws = take(scan_scratch_bytes(n));
  L.cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, L.cub_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)n);
  L.cub_ws = take(L.cub_bytes);
  L.total = (size_t)(p - (char*)base);
  return L;
}

// ---------------------------------------------------------------------------------------------
// Nearest support point per query (k = 1) on the same voxel-hash grid: eval-time back-projection of block / subsampled
// predictions onto the full cloud (replaces torch_geometric.nn.knn(x, y, k=1) in
// torch_points3d/metrics/panoptic_tracker_pointgroup_npm3d.py:384,592).
// One thread per query walks the cube of cells around it ring by ring (Chebyshev radius R = 1, 2, ...): after ring R every
// unvisited support point is farther than R * cell, so the search stops as soon as the best distance is <= R * cell.
// Best = smallest fp32 d2 = fma(dz,dz, fma(dy,dy, dx*dx)), ties to the smaller support index (a brute-force scan in
// index order with a strict `<` keeps exactly that one).  Queries with nothing inside max_ring rings scan all rows.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void nn1_scan_cell(const float4* __restrict__ spos, int cur, int end, float qx, float qy,
                                              float qz, float* best_d2, int* best_id) {
  for (int i = cur; i < end; ++i) {
    const float4 p = __ldg(&spos[i]);
    const float dx = p.x - qx, dy = p.y - qy, dz = p.z - qz;
    const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    const int id = __float_as_int(p.w);
    if (d2 < *best_d2 || (d2 == *best_d2 && id < *best_id)) {
      *best_d2 = d2;
      *best_id = id;
    }
  }
}

__global__ void __launch_bounds__(kCT) nn1_kernel(const float4* __restrict__ spos, const float4* __restrict__ qpos,
                                                   const uint64_t* __restrict__ qkeys, int64_t n_q,
                                                   const uint64_t* __restrict__ tkeys, const int32_t* __restrict__ tvals,
                                                   uint64_t mask, const int32_t* __restrict__ cell_start,
                                                   const int32_t* __restrict__ meta, float cell, int max_ring,
                                                   int32_t* __restrict__ idx_out, float* __restrict__ d2_out) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_q) return;
  const float4 qp = qpos[w];
  const int q = __float_as_int(qp.w);
  const uint64_t qk = qkeys[w];
  float best_d2 = __int_as_float(0x7f800000);   // +inf
  int best_id = 0x7fffffff;
  bool done = false;
  if (qk != kInvalidKey) {
    const int cx = (int)(qk & 0xffff), cy = (int)((qk >> 16) & 0xffff), cz = (int)((qk >> 32) & 0xffff);
    const uint64_t ghi = qk & 0xffff000000000000ull;
    for (int R = 1; R <= max_ring && !done; ++R) {
      // ring R = cells at Chebyshev distance exactly R (R == 1 also takes the centre cube: distance 0)
      for (int dz = -R; dz <= R; ++dz)
        for (int dy = -R; dy <= R; ++dy) {
          const bool edge_zy = (dz == -R || dz == R || dy == -R || dy == R);
          const int step = (edge_zy || R == 1) ? 1 : 2 * R;   // interior of the slab: only the two x faces
          for (int dx = -R; dx <= R; dx += step) {
            const int x = cx + dx, y = cy + dy, z = cz + dz;
            if (x < 2 || y < 2 || z < 2 || x > 65534 || y > 65534 || z > 65534) continue;
            const uint64_t k = ghi | ((uint64_t)(unsigned)z << 32) | ((uint64_t)(unsigned)y << 16) | (uint64_t)(unsigned)x;
            uint64_t slot = hash64(k) & mask;
            for (;;) {
              const uint64_t tk = __ldg(&tkeys[slot]);
              if (tk == k) {
                const int c = __ldg(&tvals[slot]);
                nn1_scan_cell(spos, __ldg(&cell_start[c]), __ldg(&cell_start[c + 1]), qp.x, qp.y, qp.z, &best_d2, &best_id);
                break;
              }
              if (tk == kEmptyKey) break;
              slot = (slot + 1) & mask;
            }
          }
        }
      const float reach = (float)R * cell * 0.9999f;   // (margin for the fp32 rounding of d2)
      if (best_d2 <= reach * reach) done = true;
    }
  }
  if (!done) {   // sparse neighbourhood, or a query outside the packable cell range: exhaustive scan of the support rows
    const int n_valid = meta[0];   // (single-group contract: callers pass one group; merging.py does)
    best_d2 = __int_as_float(0x7f800000);
    best_id = 0x7fffffff;
    for (int i = 0; i < n_valid; ++i) {
      const float4 p = __ldg(&spos[i]);
      const float dx = p.x - qp.x, dy = p.y - qp.y, dz = p.z - qp.z;
      const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
      const int id = __float_as_int(p.w);
      if (d2 < best_d2 || (d2 == best_d2 && id < best_id)) {
        best_d2 = d2;
        best_id = id;
      }
    }
  }
  idx_out[q] = best_id == 0x7fffffff ? -1 : best_id;
  d2_out[q] = best_d2;
}

}  // namespace pgs

using namespace pgs;

extern "C" {

size_t pgs_bq_grid_scratch_bytes(int64_t n) { return grid_layout(n < 1 ? 1 : n, nullptr).total; }

int pgs_bq_grid_build(const float* pos, const int32_t* gid, int64_t n, float cell, uint64_t* tkeys, int32_t* tvals,
                      int64_t cap, float* spos, uint64_t* skeys_out, int32_t* cell_start, int32_t* meta,
                      uint32_t* status, void* scratch, size_t scratch_bytes, void* stream) {
  PGS_CHECK_ARG(n >= 0 && n < (1ll << 31), "row count out of range");
  PGS_CHECK_ARG(cell > 0.f, "cell edge must be positive");
  PGS_CHECK_ARG(cap >= 2 * n && (cap & (cap - 1)) == 0, "capacity must be a power of two >= 2n");
  PGS_CHECK_ARG(scratch_bytes >= pgs_bq_grid_scratch_bytes(n), "scratch too small");
  cudaStream_t s = (cudaStream_t)stream;
  PGS_CUDA(cudaMemsetAsync(meta, 0, 2 * sizeof(int32_t), s));
  PGS_CUDA(cudaMemsetAsync(tkeys, 0xff, (size_t)cap * sizeof(uint64_t), s));
  if (n == 0) return PGS_OK;
  GridLayout L = grid_layout(n, scratch);
  const int g = grid_for_c(n);
  grid_key_kernel<<<g, kCT, 0, s>>>(pos, gid, n, 1.0 / (double)cell, L.keys, L.ids, status);
  count_launch();
  size_t cb = L.cub_bytes;
  PGS_CUDA(cub::DeviceRadixSort::SortPairs(L.cub_ws, cb, L.keys, skeys_out, L.ids, L.sids, (int)n, 0, 64, s));
  grid_pack_kernel<<<g, kCT, 0, s>>>(pos, skeys_out, L.sids, n, (float4*)spos, L.heads);
  count_launch();
  int rc = exclusive_scan_i32(L.heads, L.ord, n, L.scan_ws, s);
  if (rc) return rc;
  grid_cells_kernel<<<g, kCT, 0, s>>>(skeys_out, L.heads, L.ord, n, tkeys, tvals, (uint64_t)(cap - 1), cell_start,
                                      meta);
  count_launch();
  PGS_CHECK_LAUNCH();
  return PGS_OK;
}

int pgs_bq_pack_queries(const float* pos, const int32_t* gid, int64_t n, float cell, float* qpos, uint64_t* qkeys,
                        void* stream) {
  PGS_CHECK_ARG(cell > 0.f, "cell edge must be positive");
  if (n == 0) return PGS_OK;
  query_pack_kernel<<<grid_for_c(n), kCT, 0, (cudaStream_t)stream>>>(pos, gid, n, 1.0 / (double)cell, (float4*)qpos,
                                                                     qkeys);
  count_launch();
  PGS_CHECK_LAUNCH();
  return PGS_OK;
}

int pgs_bq_query(const float* spos, const float* qpos, const uint64_t* qkeys, int64_t n_q, const uint64_t* tkeys,
                 const int32_t* tvals, int64_t cap, const int32_t* cell_start, float radius, int32_t nsample,
                 int32_t* nbr, int32_t* cnt, float* dist, void* stream) {
  PGS_CHECK_ARG(nsample >= 1, "nsample must be >= 1");
  PGS_CHECK_ARG((cap & (cap - 1)) == 0, "capacity must be a power of two");
  if (n_q == 0) return PGS_OK;
  cudaStream_t s = (cudaStream_t)stream;
  PGS_CUDA(cudaMemsetAsync(cnt, 0, (size_t)n_q * sizeof(int32_t), s));
  const float r2 = radius * radius;
  const int64_t threads = n_q * 32;
  ball_query_kernel<<<(unsigned)((threads + kCT - 1) / kCT), kCT, 0, s>>>(
      (const float4*)spos, (const float4*)qpos, qkeys, n_q, tkeys, tvals, (uint64_t)(cap - 1), cell_start, r2,
      nsample, nbr, cnt, dist);
  count_launch();
  PGS_CHECK_LAUNCH();
  return PGS_OK;
}

int pgs_bq_export(const int32_t* nbr, const float* dist, const int32_t* cnt, int64_t n, int32_t nsample, int64_t* idx_out, float* dist_out, void* stream) {
  if (n == 0) return PGS_OK;
  const int64_t threads = n * 32;
  ball_query_export_kernel<<<(unsigned)((threads + kCT - 1) / kCT), kCT, 0, (cudaStream_t)stream>>>(
      nbr, dist, cnt, n, nsample, idx_out, dist_out);
  count_launch();
  PGS_CHECK_LAUNCH();
  return PGS_OK;
}

int pgs_rg_init(const int32_t* gid, int64_t n, int32_t* label, void* stream) {
  if (n == 0) return PGS_OK;
  rg_init_kernel<<<grid_for_c(n), kCT, 0, (cudaStream_t)stream>>>(gid, n, label);
  count_launch();
  PGS_CHECK_LAUNCH();
  return PGS_OK;
}

int pgs_rg_propagate(const int32_t* nbr, const int32_t* cnt, const int32_t* gid, int64_t n, int32_t nsample,
                     int32_t rounds, int32_t* label, int32_t* pushed, int32_t* changed, void* stream) {
  if (n == 0) return PGS_OK;
  cudaStream_t s = (cudaStream_t)stream;
  PGS_CUDA(cudaMemsetAsync(changed, 0, sizeof(int32_t), s));
  static int lanes = 0;
  if (lanes == 0) {
    const char* e = getenv("PGS_RG_LANES");
    lanes = (e && atoi(e) == 32) ? 32 : 8;
  }
  const int64_t threads = n * lanes;
  for (int r = 0; r < rounds; ++r) {
    if (lanes == 32)
      rg_push_kernel<32><<<(unsigned)((threads + kCT - 1) / kCT), kCT, 0, s>>>(nbr, cnt, gid, n, nsample, label, pushed, changed);
    else
      rg_push_kernel<8><<<(unsigned)((threads + kCT - 1) / kCT), kCT, 0, s>>>(nbr, cnt, gid, n, nsample, label, pushed, changed);
    rg_jump_kernel<<<grid_for_c(n), kCT, 0, s>>>(n, label, changed);
    count_launch(2);
  }
  PGS_CHECK_LAUNCH();
  return PGS_OK;
}

int pgs_nn1_query(const float* spos, const float* qpos, const uint64_t* qkeys, int64_t n_q, const uint64_t* tkeys,
                  const int32_t* tvals, int64_t cap, const int32_t* cell_start, const int32_t* meta, float cell,
                  int32_t max_ring, int32_t* idx_out, float* d2_out, void* stream) {
  PGS_CHECK_ARG((cap & (cap - 1)) == 0, "capacity must be a power of two");
  PGS_CHECK_ARG(cell > 0.f && max_ring >= 1, "cell edge and ring bound must be positive");
  if (n_q == 0) return PGS_OK;
  nn1_kernel<<<grid_for_c(n_q), kCT, 0, (cudaStream_t)stream>>>((const float4*)spos, (const float4*)qpos, qkeys, n_q,
                                                                tkeys, tvals, (uint64_t)(cap - 1), cell_start, meta,
                                                                cell, max_ring, idx_out, d2_out);
  count_launch();
  PGS_CHECK_LAUNCH();
  return PGS_OK;
}

}  // extern "C"
