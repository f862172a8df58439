// HDBSCAN front half on the device: k-NN core distances and the exact mutual-reachability minimum
// spanning tree (Boruvka), float64 arithmetic.
//
// Replaces hdbscan 0.8.27's KD-tree / dual-tree Boruvka (Cython, CPU) that the reference reaches through
// torch_points3d/utils/hdbscan_cluster.py:8-13,117-167 (and pointgroupembed.py:240-245,704;
// pointgroup.py:208-212).  Published algorithm: SURVEY App. D steps 1-3; scikit-learn's statement of the
// same steps: sklearn/cluster/_hdbscan/hdbscan.py:343-360, _linkage.pyx:111-223.
//
// Layout: points are sorted by a D-dimensional Morton code and cut into leaf blocks of 32 consecutive
// points with an axis-aligned box each.  One warp owns one query block (lane = query point) and sweeps ALL
// candidate blocks outward from its own position: 32 boxes are tested per step (one per lane, box-to-box
// lower bound against the warp's current pruning bound, plus "same component" and "minimum core distance"
// rejections in the Boruvka rounds), surviving blocks are staged through shared memory and evaluated
// 32 x 32.  No tree, no stack: the sweep is a coalesced stream over nb * 2D floats that lives in L2.
//
// Determinism: edge weights are float64 with one rounding per operation in the order
// d = sqrt(((t0^2 + t1^2) + t2^2) + ...), w = max(core_a, core_b, d / alpha); edges are strictly ordered by
// (w, min(a,b), max(a,b)) with ORIGINAL row ids, so the MST is unique and bit-identical to the oracle's Prim.
#include <cub/device/device_radix_sort.cuh>

#include <cfloat>

#include "common.cuh"

namespace pgs {

constexpr int kHB = 32;        // points per leaf block
constexpr int kHT = 256;       // threads per CTA (8 warps)
constexpr int kHdbParLanes = 6;   // <= this many interested query lanes: candidates are evaluated lane-parallel
constexpr uint64_t kU64Max = ~0ull;

__device__ __forceinline__ unsigned f2ord(float f) {
  const unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// mm[0..D) = ordered-uint min, mm[D..2D) = ordered-uint max  (caller: min init 0xffffffff, max init 0)
__global__ void __launch_bounds__(kHT) hdb_bbox_kernel(const float* __restrict__ X, int64_t n, int D,
                                                        unsigned* __restrict__ mm, uint32_t* __restrict__ status) {
  for (int d = 0; d < D; ++d) {
    unsigned lo = 0xffffffffu, hi = 0u;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      const float x = X[i * D + d];
      if (!(fabsf(x) <= FLT_MAX)) atomicOr(status, PGS_STATUS_COORD_RANGE);  // NaN / inf
      const unsigned o = f2ord(x);
      lo = min(lo, o);
      hi = max(hi, o);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, s));
      hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, s));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(&mm[d], lo);
      atomicMax(&mm[D + d], hi);
    }
  }
}

__global__ void __launch_bounds__(kHT) hdb_morton_kernel(const float* __restrict__ X, int64_t n, int D,
                                                          const unsigned* __restrict__ mm,
                                                          uint64_t* __restrict__ keys, int32_t* __restrict__ ids) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int bits = min(60 / D, 21);
  uint64_t key = 0;
  for (int d = 0; d < D; ++d) {
    const float lo = ord2f(mm[d]), hi = ord2f(mm[D + d]);
    const float span = hi - lo;
    float t = span > 0.f ? (X[i * D + d] - lo) / span : 0.f;
    t = fminf(fmaxf(t, 0.f), 1.f);
    const uint32_t q = (uint32_t)(t * (float)((1u << bits) - 1u));
    for (int b = 0; b < bits; ++b) key |= (uint64_t)((q >> b) & 1u) << (b * D + d);
  }
  keys[i] = key;
  ids[i] = (int32_t)i;
}

// sorted rows P [n, D] fp32, original ids, inverse permutation
__global__ void __launch_bounds__(kHT) hdb_gather_kernel(const float* __restrict__ X, const int32_t* __restrict__ sids,
                                                          int64_t n, int D, float* __restrict__ P,
                                                          int32_t* __restrict__ inv) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int id = sids[i];
  inv[id] = (int32_t)i;
  for (int d = 0; d < D; ++d) P[i * D + d] = X[(int64_t)id * D + d];
}

// one warp per leaf block: box lo/hi per dimension
template <int D>
__global__ void __launch_bounds__(kHT) hdb_block_box_kernel(const float* __restrict__ P, int64_t n, int nb,
                                                             float* __restrict__ blo, float* __restrict__ bhi) {
  const int lane = threadIdx.x & 31;
  const int b = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (b >= nb) return;
  const int64_t i = (int64_t)b * kHB + lane;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    float lo = FLT_MAX, hi = -FLT_MAX;
    if (i < n) lo = hi = P[i * D + d];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, s));
      hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, s));
    }
    if (lane == 0) {
      blo[(int64_t)b * D + d] = lo;
      bhi[(int64_t)b * D + d] = hi;
    }
  }
}

// Second level of the box hierarchy: one box per GROUP of 32 consecutive leaf blocks (1024 Morton-consecutive points).
// The sweeps test a group box first and only look at the 32 leaf boxes of groups that survive: ~nb/32 + survivors steps
// per warp instead of nb/32 steps of 32 leaf tests each (measured at 480 k points: every warp walked all 15 k leaf boxes
// in every round, 10.7 M box steps of ~1.5 k cycles).
template <int D>
__global__ void __launch_bounds__(kHT) hdb_group_box_kernel(const float* __restrict__ blo, const float* __restrict__ bhi,
                                                             int nb, int ng, float* __restrict__ glo,
                                                             float* __restrict__ ghi) {
  const int lane = threadIdx.x & 31;
  const int g = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (g >= ng) return;
  const int b = g * 32 + lane;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    float lo = FLT_MAX, hi = -FLT_MAX;
    if (b < nb) {
      lo = blo[(int64_t)b * D + d];
      hi = bhi[(int64_t)b * D + d];
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, s));
      hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, s));
    }
    if (lane == 0) {
      glo[(int64_t)g * D + d] = lo;
      ghi[(int64_t)g * D + d] = hi;
    }
  }
}

// per group: minimum core distance (once) and, per round, the component shared by ALL its blocks (-1: mixed)
__global__ void __launch_bounds__(kHT) hdb_group_meta_kernel(const double* __restrict__ bmincore,
                                                              const int32_t* __restrict__ bcomp, int nb, int ng,
                                                              double* __restrict__ gmincore, int32_t* __restrict__ gcomp) {
  const int lane = threadIdx.x & 31;
  const int g = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (g >= ng) return;
  const int b = g * 32 + lane;
  if (gmincore) {
    double c = b < nb ? bmincore[b] : INFINITY;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) c = fmin(c, __shfl_xor_sync(0xffffffffu, c, s));
    if (lane == 0) gmincore[g] = c;
  }
  if (gcomp) {
    const int c = b < nb ? bcomp[b] : -2;           // -2: past the end (neutral), -1: the block itself is mixed
    const int c0 = __shfl_sync(0xffffffffu, c, 0);
    const bool uni = __all_sync(0xffffffffu, (c == c0 && c >= 0) || c == -2);
    if (lane == 0) gcomp[g] = (uni && c0 >= 0) ? c0 : -1;
  }
}

__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, s));
  return v;
}

// squared distance, one rounding per operation (no FMA contraction), dimension order 0..D-1
template <int D>
__device__ __forceinline__ double sqdist_rn(const double* q, const float* p) {
  double s = 0.0;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const double t = __dsub_rn(q[d], (double)p[d]);
    s = __dadd_rn(s, __dmul_rn(t, t));
  }
  return s;
}

// lower bound of the squared distance between two boxes; every term is <= the matching term of any
// point pair inside the boxes and rounding is monotone, so lb <= sqdist_rn of every such pair.
template <int D>
__device__ __forceinline__ double box_box_lb2(const float* qlo, const float* qhi, const float* __restrict__ clo,
                                              const float* __restrict__ chi) {
  double s = 0.0;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const double g1 = __dsub_rn((double)clo[d], (double)qhi[d]);
    const double g2 = __dsub_rn((double)qlo[d], (double)chi[d]);
    const double g = fmax(fmax(g1, g2), 0.0);
    s = __dadd_rn(s, __dmul_rn(g, g));
  }
  return s;
}

// lower bound of the squared distance from a point to a box (same monotonicity argument as box_box_lb2)
template <int D>
__device__ __forceinline__ double point_box_lb2(const double* q, const float* __restrict__ clo,
                                                const float* __restrict__ chi) {
  double s = 0.0;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const double g1 = __dsub_rn((double)clo[d], q[d]);
    const double g2 = __dsub_rn(q[d], (double)chi[d]);
    const double g = fmax(fmax(g1, g2), 0.0);
    s = __dadd_rn(s, __dmul_rn(g, g));
  }
  return s;
}

// candidate block visited at step t of the outward sweep from block qb
__device__ __forceinline__ int sweep_block(int qb, int t) { return (t & 1) ? qb + ((t + 1) >> 1) : qb - (t >> 1); }

// ------------------------------------------------------------------------------------------
// core distances: k-th smallest distance counting the point itself
// ------------------------------------------------------------------------------------------
template <int D, int KMAX>
__global__ void __launch_bounds__(kHT) hdb_knn_kernel(const float* __restrict__ P, const int32_t* __restrict__ sids,
                                                       int64_t n, int nb, const float* __restrict__ blo,
                                                       const float* __restrict__ bhi, const float* __restrict__ glo,
                                                       const float* __restrict__ ghi, int k,
                                                       double* __restrict__ core_sorted,
                                                       double* __restrict__ core_orig) {
  __shared__ float s_pts[kHT / 32][kHB * D];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int qb = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (qb >= nb) return;
  const int64_t a = (int64_t)qb * kHB + lane;
  const bool valid = a < n;
  double q[D];
#pragma unroll
  for (int d = 0; d < D; ++d) q[d] = valid ? (double)P[a * D + d] : 0.0;
  float qlo[D], qhi[D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    qlo[d] = blo[(int64_t)qb * D + d];
    qhi[d] = bhi[(int64_t)qb * D + d];
  }
  double best[KMAX];
#pragma unroll
  for (int j = 0; j < KMAX; ++j) best[j] = (valid && j < k) ? INFINITY : -1.0;
  // best[0..k) ascending; entries >= k are -1 and never move
  double kth = valid ? INFINITY : -1.0;
  double wb = INFINITY;

  const int qg = qb >> 5, ng = (nb + 31) >> 5;
  const int gspan = 2 * max(qg, ng - 1 - qg) + 1;
  for (int gbase = 0; gbase < gspan; gbase += 32) {
    const int cg = sweep_block(qg, gbase + lane);
    bool gpass = cg >= 0 && cg < ng;
    if (gpass) gpass = box_box_lb2<D>(qlo, qhi, glo + (int64_t)cg * D, ghi + (int64_t)cg * D) <= wb;
    unsigned gmask = __ballot_sync(0xffffffffu, gpass);
    while (gmask) {
      const int gsrc = __ffs(gmask) - 1;
      gmask &= gmask - 1;
      const int grp = __shfl_sync(0xffffffffu, cg, gsrc);
      const int cb = grp * 32 + lane;
      bool pass = cb < nb;
      if (pass) pass = box_box_lb2<D>(qlo, qhi, blo + (int64_t)cb * D, bhi + (int64_t)cb * D) <= wb;
      unsigned mask = __ballot_sync(0xffffffffu, pass);
      while (mask) {
        const int src = __ffs(mask) - 1;
        mask &= mask - 1;
        const int blk = __shfl_sync(0xffffffffu, cb, src);
        const int64_t p0 = (int64_t)blk * kHB;
        const int cnt = (int)min((int64_t)kHB, n - p0);
        // per lane: only points whose own k-th distance still reaches the candidate box look at it (see hdb_search_kernel)
        const bool want = valid && point_box_lb2<D>(q, blo + (int64_t)blk * D, bhi + (int64_t)blk * D) <= kth;
        if (!__any_sync(0xffffffffu, want)) continue;
        __syncwarp();
        for (int e = lane; e < cnt * D; e += 32) s_pts[wib][e] = __ldg(&P[p0 * D + e]);
        __syncwarp();
        if (want) {
          for (int j = 0; j < cnt; ++j) {
            double x = sqdist_rn<D>(q, &s_pts[wib][j * D]);
            if (x < kth) {
#pragma unroll
              for (int m = 0; m < KMAX; ++m)
                if (m < k && x < best[m]) {
                  const double t = best[m];
                  best[m] = x;
                  x = t;
                }
#pragma unroll
              for (int m = 0; m < KMAX; ++m)
                if (m == k - 1) kth = best[m];
            }
          }
        }
        wb = warp_max_d(kth);
      }
    }
  }
  if (valid) {
    const double c = __dsqrt_rn(kth);
    core_sorted[a] = c;
    core_orig[sids[a]] = c;
  }
}

// ------------------------------------------------------------------------------------------
// Boruvka rounds
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kHT) hdb_block_core_kernel(const double* __restrict__ core_sorted, int64_t n, int nb,
                                                              double* __restrict__ bmincore) {
  const int lane = threadIdx.x & 31;
  const int b = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (b >= nb) return;
  const int64_t i = (int64_t)b * kHB + lane;
  double c = i < n ? core_sorted[i] : INFINITY;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) c = fmin(c, __shfl_xor_sync(0xffffffffu, c, s));
  if (lane == 0) bmincore[b] = c;
}

__global__ void __launch_bounds__(kHT) hdb_round_init_kernel(const int32_t* __restrict__ comp, int64_t n, int nb,
                                                              int32_t* __restrict__ bcomp, uint64_t* __restrict__ U,
                                                              uint64_t* __restrict__ E) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    U[i] = kU64Max;
    E[i] = kU64Max;
  }
  const int lane = threadIdx.x & 31;
  const int64_t b = i >> 5;
  if (b < nb) {
    const int c = i < n ? comp[i] : -2;
    const int c0 = __shfl_sync(0xffffffffu, c, 0);
    const bool same = (c == c0) || (c == -2);
    const bool uni = __all_sync(0xffffffffu, same);
    if (lane == 0) bcomp[b] = uni ? c0 : -1;
  }
}

// search statistics of the last pgs_hdb_mst call (diagnostics: pgs_hdb_search_stats): per round
// {box-test steps (group + leaf level), candidate blocks evaluated, point pairs evaluated, warps}
__device__ unsigned long long g_hdb_stats[64 * 4];
__device__ unsigned long long g_hdb_clk[64 * 4];   // per round: sum / max of warp cycles, sum / max of cycles in point evaluation

__device__ __forceinline__ double u2d(uint64_t u) { return u == kU64Max ? INFINITY : __longlong_as_double((long long)u); }

template <int D>
__global__ void __launch_bounds__(kHT) hdb_search_kernel(
    const float* __restrict__ P, const int32_t* __restrict__ sids, const double* __restrict__ core_sorted,
    const int32_t* __restrict__ comp, int64_t n, int nb, const float* __restrict__ blo, const float* __restrict__ bhi,
    const double* __restrict__ bmincore, const int32_t* __restrict__ bcomp, const float* __restrict__ glo,
    const float* __restrict__ ghi, const double* __restrict__ gmincore, const int32_t* __restrict__ gcomp, double alpha,
    uint64_t* __restrict__ U, double* __restrict__ bestw, int32_t* __restrict__ bestp, int round) {
  __shared__ float s_pts[kHT / 32][kHB * D];
  __shared__ double s_core[kHT / 32][kHB];
  __shared__ int s_comp[kHT / 32][kHB];
  __shared__ int s_oid[kHT / 32][kHB];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int qb = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (qb >= nb) return;
  const int64_t a = (int64_t)qb * kHB + lane;
  const bool valid = a < n;
  double q[D];
#pragma unroll
  for (int d = 0; d < D; ++d) q[d] = valid ? (double)P[a * D + d] : 0.0;
  float qlo[D], qhi[D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    qlo[d] = blo[(int64_t)qb * D + d];
    qhi[d] = bhi[(int64_t)qb * D + d];
  }
  const int ca = valid ? comp[a] : -3;
  const double core_a = valid ? core_sorted[a] : INFINITY;
  const int oa = valid ? sids[a] : 0;
  const int qc = bcomp[qb];
  const double qmin = bmincore[qb];

  double bw = INFINITY;  // best weight
  int bj = -1, blo_id = 0x7fffffff, bhi_id = 0x7fffffff;
  double published = INFINITY;
  unsigned st_steps = 0, st_blocks = 0, st_pairs = 0;
  double ucache = INFINITY;   // last value read from U[ca]
  const long long t_begin = clock64();
  long long t_eval = 0;

  const int qg = qb >> 5, ng = (nb + 31) >> 5;
  const int gspan = 2 * max(qg, ng - 1 - qg) + 1;
  // pruning bound of a lane: its best so far and whatever its component has already published (a stale value of U is
  // only a weaker, still valid, upper bound); -1 once nothing at this lane can win any more (w >= core_a)
  auto lane_bound = [&]() {
    ucache = valid ? u2d(*(volatile uint64_t*)&U[ca]) : INFINITY;
    double b = valid ? fmin(bw, ucache) : -1.0;
    if (core_a > b) b = -1.0;
    return b;
  };
  for (int gbase = 0; gbase < gspan; gbase += 32) {
    double bnd = lane_bound();
    double wb = warp_max_d(bnd);
    if (wb < 0.0) break;           // every lane of the block is settled
    ++st_steps;
    double wba = wb * alpha;                              // bound on the raw distance (w >= d / alpha)
    double wb2 = wba * wba * (1.0 + 8.0 * DBL_EPSILON);  // squared-space bound, rounded up
    const int cg = sweep_block(qg, gbase + lane);
    bool gpass = cg >= 0 && cg < ng;
    if (gpass)
      gpass = !(qc >= 0 && gcomp[cg] == qc) && fmax(gmincore[cg], qmin) <= wb &&
              box_box_lb2<D>(qlo, qhi, glo + (int64_t)cg * D, ghi + (int64_t)cg * D) <= wb2;
    unsigned gmask = __ballot_sync(0xffffffffu, gpass);
    const long long t_e0 = clock64();
    while (gmask) {
      const int gsrc = __ffs(gmask) - 1;
      gmask &= gmask - 1;
      const int grp = __shfl_sync(0xffffffffu, cg, gsrc);
      // the bound may have tightened since the group test
      bnd = lane_bound();
      wb = warp_max_d(bnd);
      if (wb < 0.0) break;
      ++st_steps;
      wba = wb * alpha;
      wb2 = wba * wba * (1.0 + 8.0 * DBL_EPSILON);
      const int cb = grp * 32 + lane;
      bool pass = cb < nb;
      if (pass) {
        const int bc = bcomp[cb];
        pass = !(qc >= 0 && bc == qc) && fmax(bmincore[cb], qmin) <= wb &&
               box_box_lb2<D>(qlo, qhi, blo + (int64_t)cb * D, bhi + (int64_t)cb * D) <= wb2;
      }
      unsigned mask = __ballot_sync(0xffffffffu, pass);
      while (mask) {
        const int src = __ffs(mask) - 1;
        mask &= mask - 1;
        const int blk = __shfl_sync(0xffffffffu, cb, src);
        const int64_t p0 = (int64_t)blk * kHB;
        const int cnt = (int)min((int64_t)kHB, n - p0);
        // The box-to-box tests use the LOOSEST bound of the 32 lanes: one lane with a far best candidate (a block
        // straddling two clusters, an outlier) would drag the whole warp through every block point by point (measured:
        // a single warp spending 135 M cycles = the entire 71 ms of a late round).  Per lane: point-to-box bound against
        // the lane's OWN pruning bound; the block is staged only if some lane still wants it.
        bool want = false;
        if (bnd >= 0.0 && fmax(core_a, bmincore[blk]) <= bnd) {
          const double ba = bnd * alpha;
          want = point_box_lb2<D>(q, blo + (int64_t)blk * D, bhi + (int64_t)blk * D) <= ba * ba * (1.0 + 8.0 * DBL_EPSILON);
        }
        if (!__any_sync(0xffffffffu, want)) continue;
        ++st_blocks;
        __syncwarp();
        for (int e = lane; e < cnt * D; e += 32) s_pts[wib][e] = __ldg(&P[p0 * D + e]);
        if (lane < cnt) {
          s_core[wib][lane] = core_sorted[p0 + lane];
          s_comp[wib][lane] = comp[p0 + lane];
          s_oid[wib][lane] = sids[p0 + lane];
        }
        __syncwarp();
        const unsigned wmask = __ballot_sync(0xffffffffu, want);
        if (__popc(wmask) <= kHdbParLanes) {
          // Few lanes still want this block (the usual case in the late rounds: a far, isolated component has a large
          // bound and its warps visit ~1000 blocks each with 1-3 interested lanes).  Turn the loop around: for each
          // interested query lane, the 32 lanes evaluate the block's 32 candidates in parallel and a warp argmin under the
          // strict (w, min id, max id) order hands the winner back -- ~32x fewer serial fp64 chains than lane-per-query.
          unsigned rem = wmask;
          while (rem) {
            const int L = __ffs(rem) - 1;
            rem &= rem - 1;
            double ql[D];
#pragma unroll
            for (int d = 0; d < D; ++d) ql[d] = __shfl_sync(0xffffffffu, q[d], L);
            const double coreL = __shfl_sync(0xffffffffu, core_a, L), bndL = __shfl_sync(0xffffffffu, bnd, L);
            const int caL = __shfl_sync(0xffffffffu, ca, L), oaL = __shfl_sync(0xffffffffu, oa, L);
            double w = INFINITY;
            int lo = 0x7fffffff, hi = 0x7fffffff, idx = lane;
            if (lane < cnt && s_comp[wib][lane] != caL) {
              double ww = fmax(coreL, s_core[wib][lane]);
              if (ww <= bndL) {
                ++st_pairs;
                const double d2 = sqdist_rn<D>(ql, &s_pts[wib][lane * D]);
                const double ba = bndL * alpha;
                if (d2 <= ba * ba * (1.0 + 8.0 * DBL_EPSILON)) {
                  ww = fmax(ww, __ddiv_rn(__dsqrt_rn(d2), alpha));
                  const int oj = s_oid[wib][lane];
                  w = ww;
                  lo = min(oaL, oj);
                  hi = max(oaL, oj);
                }
              }
            }
#pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1) {
              const double w2 = __shfl_xor_sync(0xffffffffu, w, sft);
              const int lo2 = __shfl_xor_sync(0xffffffffu, lo, sft), hi2 = __shfl_xor_sync(0xffffffffu, hi, sft);
              const int idx2 = __shfl_xor_sync(0xffffffffu, idx, sft);
              if (w2 < w || (w2 == w && (lo2 < lo || (lo2 == lo && hi2 < hi)))) {
                w = w2;
                lo = lo2;
                hi = hi2;
                idx = idx2;
              }
            }
            if (lane == L && w < INFINITY) {
              const bool better = (w < bw) || (w == bw && (lo < blo_id || (lo == blo_id && hi < bhi_id)));
              if (better) {
                bw = w;
                bj = (int)(p0 + idx);
                blo_id = lo;
                bhi_id = hi;
                if (w < bnd) bnd = w;
              }
            }
          }
        } else if (want) {
          for (int j = 0; j < cnt; ++j) {
            if (s_comp[wib][j] == ca) continue;
            double w = fmax(core_a, s_core[wib][j]);
            if (w > bnd) continue;
            ++st_pairs;
            const double d2 = sqdist_rn<D>(q, &s_pts[wib][j * D]);
            const double ba = bnd * alpha;
            if (d2 > ba * ba * (1.0 + 8.0 * DBL_EPSILON)) continue;
            const double dist = __ddiv_rn(__dsqrt_rn(d2), alpha);
            w = fmax(w, dist);
            const int oj = s_oid[wib][j];
            const int lo = min(oa, oj), hi = max(oa, oj);
            const bool better = (w < bw) || (w == bw && (lo < blo_id || (lo == blo_id && hi < bhi_id)));
            if (better) {
              bw = w;
              bj = (int)(p0 + j);
              blo_id = lo;
              bhi_id = hi;
              if (w < bnd) bnd = w;
            }
          }
        }
      }
      if (valid && bw < published) {
        atomicMin((unsigned long long*)&U[ca], (unsigned long long)__double_as_longlong(bw));
        published = bw;
      }
    }
    t_eval += clock64() - t_e0;
  }
  if (valid) {
    bestw[a] = bw;
    bestp[a] = bj;
    if (bw < published) atomicMin((unsigned long long*)&U[ca], (unsigned long long)__double_as_longlong(bw));
  }
  if (round < 64) {
    unsigned pr = st_pairs;
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) pr += __shfl_xor_sync(0xffffffffu, pr, sft);
    if (lane == 0) {
      atomicAdd(&g_hdb_stats[round * 4 + 0], (unsigned long long)st_steps);
      atomicAdd(&g_hdb_stats[round * 4 + 1], (unsigned long long)st_blocks);
      atomicAdd(&g_hdb_stats[round * 4 + 2], (unsigned long long)pr);
      atomicAdd(&g_hdb_stats[round * 4 + 3], 1ull);   // warps that ran
      const unsigned long long tt = (unsigned long long)(clock64() - t_begin);
      atomicAdd(&g_hdb_clk[round * 4 + 0], tt);
      atomicMax(&g_hdb_clk[round * 4 + 1], tt);
      atomicAdd(&g_hdb_clk[round * 4 + 2], (unsigned long long)t_eval);
      atomicMax(&g_hdb_clk[round * 4 + 3], (unsigned long long)t_eval);
    }
  }
}

// among the points that hold their component's minimum weight, the smallest (min id, max id) wins
__global__ void __launch_bounds__(kHT) hdb_select_kernel(const int32_t* __restrict__ sids,
                                                          const int32_t* __restrict__ comp,
                                                          const double* __restrict__ bestw,
                                                          const int32_t* __restrict__ bestp, int64_t n,
                                                          const uint64_t* __restrict__ U, uint64_t* __restrict__ E) {
  const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n) return;
  const int j = bestp[a];
  if (j < 0) return;
  const int c = comp[a];
  if ((uint64_t)__double_as_longlong(bestw[a]) != U[c]) return;
  const unsigned oa = (unsigned)sids[a], oj = (unsigned)sids[j];
  const uint64_t e = ((uint64_t)min(oa, oj) << 32) | (uint64_t)max(oa, oj);
  atomicMin((unsigned long long*)&E[c], (unsigned long long)e);
}

// one thread per component representative: emit its edge (once per mutual pair) and hook
__global__ void __launch_bounds__(kHT) hdb_merge_kernel(const int32_t* __restrict__ comp, const int32_t* __restrict__ inv,
                                                         int64_t n, const uint64_t* __restrict__ U,
                                                         const uint64_t* __restrict__ E, int32_t* __restrict__ next,
                                                         uint64_t* __restrict__ edge_uv, double* __restrict__ edge_w,
                                                         int32_t* __restrict__ n_edges) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  if (comp[c] != (int32_t)c) return;
  next[c] = (int32_t)c;
  const uint64_t e = E[c];
  if (e == kU64Max) return;
  const int pa = inv[(int32_t)(e >> 32)], pb = inv[(int32_t)(e & 0xffffffffu)];
  const int ca = comp[pa], cb = comp[pb];
  const int other = (ca == (int32_t)c) ? cb : ca;
  const bool mutual = E[other] == e;
  if (!mutual || (int32_t)c < other) {
    const int slot = atomicAdd(n_edges, 1);
    edge_uv[slot] = e;
    edge_w[slot] = u2d(U[c]);
  }
  if (!(mutual && (int32_t)c < other)) next[c] = other;
}

__global__ void __launch_bounds__(kHT) hdb_relabel_kernel(int32_t* __restrict__ comp, const int32_t* __restrict__ next,
                                                           int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c = comp[i];
  for (;;) {
    const int p = next[c];
    if (p == c) break;
    c = p;
  }
  comp[i] = c;
}

__global__ void __launch_bounds__(kHT) hdb_iota_kernel(int32_t* __restrict__ v, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = (int32_t)i;
}

__global__ void __launch_bounds__(kHT) hdb_unpack_kernel(const uint64_t* __restrict__ uv, int64_t m,
                                                          int32_t* __restrict__ u, int32_t* __restrict__ v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  u[i] = (int32_t)(uv[i] >> 32);
  v[i] = (int32_t)(uv[i] & 0xffffffffu);
}

struct HdbLayout {
  unsigned* mm;
  uint32_t* status;
  int32_t* n_edges;
  uint64_t *keys, *skeys, *U, *E, *edge_uv, *edge_uv2;
  int32_t *ids, *sids, *inv, *comp, *next, *bcomp, *bestp, *gcomp;
  float *P, *blo, *bhi, *glo, *ghi;
  double *core_sorted, *bmincore, *bestw, *edge_w, *edge_w2, *gmincore;
  void* cub_ws;
  size_t cub_bytes, total;
};

static HdbLayout hdb_layout(int64_t n, int D, void* base) {
  HdbLayout L;
  char* p = (char*)base;
  auto take = [&](size_t bytes) {
    char* r = p;
    p += align_up(bytes ? bytes : 1, 256);
    return r;
  };
  const int64_t nb = (n + kHB - 1) / kHB;
  L.mm = (unsigned*)take(sizeof(unsigned) * 2 * D);
  L.status = (uint32_t*)take(sizeof(uint32_t));
  L.n_edges = (int32_t*)take(sizeof(int32_t));
  L.keys = (uint64_t*)take(8 * n);
  L.skeys = (uint64_t*)take(8 * n);
  L.U = (uint64_t*)take(8 * n);
  L.E = (uint64_t*)take(8 * n);
  L.edge_uv = (uint64_t*)take(8 * n);
  L.edge_uv2 = (uint64_t*)take(8 * n);
  L.ids = (int32_t*)take(4 * n);
  L.sids = (int32_t*)take(4 * n);
  L.inv = (int32_t*)take(4 * n);
  L.comp = (int32_t*)take(4 * n);
  L.next = (int32_t*)take(4 * n);
  L.bcomp = (int32_t*)take(4 * nb);
  L.bestp = (int32_t*)take(4 * n);
  L.P = (float*)take(4 * n * D);
  L.blo = (float*)take(4 * nb * D);
  L.bhi = (float*)take(4 * nb * D);
  L.core_sorted = (double*)take(8 * n);
  L.bmincore = (double*)take(8 * nb);
  {
    const int64_t ng = (nb + 31) / 32;
    L.glo = (float*)take(4 * ng * D);
    L.ghi = (float*)take(4 * ng * D);
    L.gmincore = (double*)take(8 * ng);
    L.gcomp = (int32_t*)take(4 * ng);
  }
  L.bestw = (double*)take(8 * n);
  L.edge_w = (double*)take(8 * n);
  L.edge_w2 = (double*)take(8 * n);
  size_t b1 = 0, b2 = 0, b3 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, b1, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, (int)n);
  cub::DeviceRadixSort::SortPairs(nullptr, b2, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const double*)nullptr,
                                  (double*)nullptr, (int)n);
  cub::DeviceRadixSort::SortPairs(nullptr, b3, (const double*)nullptr, (double*)nullptr, (const uint64_t*)nullptr,
                                  (uint64_t*)nullptr, (int)n);
  L.cub_bytes = b1 > b2 ? (b1 > b3 ? b1 : b3) : (b2 > b3 ? b2 : b3);
  L.cub_ws = take(L.cub_bytes);
  L.total = (size_t)(p - (char*)base);
  return L;
}

static inline unsigned blocks_for(int64_t threads) { return (unsigned)((threads + kHT - 1) / kHT); }

template <int D>
static int hdb_mst_impl(const float* X, int64_t n, int k, double alpha, double* core, int32_t* u, int32_t* v, double* w,
                        int32_t* rounds_host, HdbLayout& L, cudaStream_t s) {
  const int nb = (int)((n + kHB - 1) / kHB);
  const unsigned gp = blocks_for(n), gw = blocks_for((int64_t)nb * 32);
  PGS_CUDA(cudaMemsetAsync(L.mm, 0xff, sizeof(unsigned) * D, s));
  PGS_CUDA(cudaMemsetAsync(L.mm + D, 0x00, sizeof(unsigned) * D, s));
  PGS_CUDA(cudaMemsetAsync(L.status, 0, sizeof(uint32_t), s));
  PGS_CUDA(cudaMemsetAsync(L.n_edges, 0, sizeof(int32_t), s));
  hdb_bbox_kernel<<<min(gp, (unsigned)(kNumSM * 8)), kHT, 0, s>>>(X, n, D, L.mm, L.status);
  hdb_morton_kernel<<<gp, kHT, 0, s>>>(X, n, D, L.mm, L.keys, L.ids);
  count_launch(2);
  size_t cb = L.cub_bytes;
  PGS_CUDA(cub::DeviceRadixSort::SortPairs(L.cub_ws, cb, L.keys, L.skeys, L.ids, L.sids, (int)n, 0, 64, s));
  hdb_gather_kernel<<<gp, kHT, 0, s>>>(X, L.sids, n, D, L.P, L.inv);
  hdb_block_box_kernel<D><<<gw, kHT, 0, s>>>(L.P, n, nb, L.blo, L.bhi);
  const int ng = (nb + 31) / 32;
  const unsigned gg = blocks_for((int64_t)ng * 32);
  hdb_group_box_kernel<D><<<gg, kHT, 0, s>>>(L.blo, L.bhi, nb, ng, L.glo, L.ghi);
  count_launch(3);
  if (k <= 8)
    hdb_knn_kernel<D, 8><<<gw, kHT, 0, s>>>(L.P, L.sids, n, nb, L.blo, L.bhi, L.glo, L.ghi, k, L.core_sorted, core);
  else
    hdb_knn_kernel<D, 32><<<gw, kHT, 0, s>>>(L.P, L.sids, n, nb, L.blo, L.bhi, L.glo, L.ghi, k, L.core_sorted, core);
  hdb_block_core_kernel<<<gw, kHT, 0, s>>>(L.core_sorted, n, nb, L.bmincore);
  hdb_group_meta_kernel<<<gg, kHT, 0, s>>>(L.bmincore, nullptr, nb, ng, L.gmincore, nullptr);
  hdb_iota_kernel<<<gp, kHT, 0, s>>>(L.comp, n);
  count_launch(4);
  PGS_CHECK_LAUNCH();
  uint32_t status_h = 0;
  PGS_CUDA(cudaMemcpyAsync(&status_h, L.status, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  PGS_CUDA(cudaStreamSynchronize(s));
  if (status_h) {
    set_error("pgs_hdb_mst: input contains NaN or infinity");
    return PGS_ERR_RANGE;
  }
  int32_t n_edges = 0, rounds = 0;
  {
    void* sp = nullptr;
    PGS_CUDA(cudaGetSymbolAddress(&sp, g_hdb_stats));
    PGS_CUDA(cudaMemsetAsync(sp, 0, sizeof(unsigned long long) * 64 * 4, s));
    PGS_CUDA(cudaGetSymbolAddress(&sp, g_hdb_clk));
    PGS_CUDA(cudaMemsetAsync(sp, 0, sizeof(unsigned long long) * 64 * 4, s));
  }
  while (n_edges < n - 1) {
    if (rounds >= 64) {
      set_error("pgs_hdb_mst: Boruvka did not converge (%d of %lld edges)", n_edges, (long long)(n - 1));
      return PGS_ERR_CUDA;
    }
    hdb_round_init_kernel<<<blocks_for((int64_t)nb * 32), kHT, 0, s>>>(L.comp, n, nb, L.bcomp, L.U, L.E);
    hdb_group_meta_kernel<<<gg, kHT, 0, s>>>(nullptr, L.bcomp, nb, ng, nullptr, L.gcomp);
    hdb_search_kernel<D><<<gw, kHT, 0, s>>>(L.P, L.sids, L.core_sorted, L.comp, n, nb, L.blo, L.bhi, L.bmincore,
                                            L.bcomp, L.glo, L.ghi, L.gmincore, L.gcomp, alpha, L.U, L.bestw, L.bestp,
                                            rounds);
    count_launch();
    hdb_select_kernel<<<gp, kHT, 0, s>>>(L.sids, L.comp, L.bestw, L.bestp, n, L.U, L.E);
    hdb_merge_kernel<<<gp, kHT, 0, s>>>(L.comp, L.inv, n, L.U, L.E, L.next, L.edge_uv, L.edge_w, L.n_edges);
    hdb_relabel_kernel<<<gp, kHT, 0, s>>>(L.comp, L.next, n);
    count_launch(5);
    PGS_CHECK_LAUNCH();
    const int32_t before = n_edges;
    PGS_CUDA(cudaMemcpyAsync(&n_edges, L.n_edges, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    PGS_CUDA(cudaStreamSynchronize(s));
    ++rounds;
    if (n_edges == before) {
      set_error("pgs_hdb_mst: a Boruvka round added no edge");
      return PGS_ERR_CUDA;
    }
  }
  if (rounds_host) *rounds_host = rounds;
  const int m = (int)(n - 1);
  // strict total order (w, min id, max id): stable sort by the packed ids, then by the weight
  cb = L.cub_bytes;
  PGS_CUDA(cub::DeviceRadixSort::SortPairs(L.cub_ws, cb, L.edge_uv, L.edge_uv2, L.edge_w, L.edge_w2, m, 0, 64, s));
  cb = L.cub_bytes;
  PGS_CUDA(cub::DeviceRadixSort::SortPairs(L.cub_ws, cb, L.edge_w2, w, L.edge_uv2, L.edge_uv, m, 0, 64, s));
  hdb_unpack_kernel<<<blocks_for(m), kHT, 0, s>>>(L.edge_uv, m, u, v);
  count_launch();
  PGS_CHECK_LAUNCH();
  return PGS_OK;
}

}  // namespace pgs

using namespace pgs;

extern "C" {

int pgs_hdb_morton_rank(const void* scratch, int64_t n, int32_t D, int32_t* rank_out, void* stream) {
  PGS_CHECK_ARG(scratch != nullptr && rank_out != nullptr && n >= 1 && D >= 1 && D <= 8, "bad arguments");
  HdbLayout L = hdb_layout(n, D, const_cast<void*>(scratch));
  PGS_CUDA(cudaMemcpyAsync(rank_out, L.inv, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return PGS_OK;
}

int pgs_hdb_search_stats(int64_t* out_host, int32_t max_rounds) {
  unsigned long long h[64 * 4];
  PGS_CUDA(cudaMemcpyFromSymbol(h, g_hdb_stats, sizeof(h)));
  const int nr = max_rounds < 64 ? max_rounds : 64;
  for (int i = 0; i < 4 * nr; ++i) out_host[i] = (int64_t)h[i];
  if (max_rounds < 0) {   // (negative: also return the cycle counters after the 4 * 64 statistics -- profiling scripts)
    PGS_CUDA(cudaMemcpyFromSymbol(h, g_hdb_clk, sizeof(h)));
    for (int i = 0; i < 256; ++i) out_host[i] = (int64_t)h[i];
  }
  return PGS_OK;
}

size_t pgs_hdb_scratch_bytes(int64_t n, int32_t D) { return hdb_layout(n < 1 ? 1 : n, D < 1 ? 1 : D, nullptr).total; }

int pgs_hdb_mst(const float* X, int64_t n, int32_t D, int32_t min_samples, double alpha, double* core, int32_t* u,
                int32_t* v, double* w, int32_t* rounds_host, void* scratch, size_t scratch_bytes, void* stream) {
  PGS_CHECK_ARG(n >= 2 && n < (1ll << 31), "need 2 <= n < 2^31 samples");
  PGS_CHECK_ARG(D >= 1 && D <= 8, "dimension must be in 1..8");
  PGS_CHECK_ARG(min_samples >= 1 && min_samples <= 32 && min_samples <= n, "min_samples must be in 1..min(32, n)");
  PGS_CHECK_ARG(alpha > 0.0, "alpha must be positive");
  PGS_CHECK_ARG(scratch_bytes >= pgs_hdb_scratch_bytes(n, D), "scratch too small");
  HdbLayout L = hdb_layout(n, D, scratch);
  cudaStream_t s = (cudaStream_t)stream;
  switch (D) {
    case 1: return hdb_mst_impl<1>(X, n, min_samples, alpha, core, u, v, w, rounds_host, L, s);
    case 2: return hdb_mst_impl<2>(X, n, min_samples, alpha, core, u, v, w, rounds_host, L, s);
    case 3: return hdb_mst_impl<3>(X, n, min_samples, alpha, core, u, v, w, rounds_host, L, s);
    case 4: return hdb_mst_impl<4>(X, n, min_samples, alpha, core, u, v, w, rounds_host, L, s);
    case 5: return hdb_mst_impl<5>(X, n, min_samples, alpha, core, u, v, w, rounds_host, L, s);
    case 6: return hdb_mst_impl<6>(X, n, min_samples, alpha, core, u, v, w, rounds_host, L, s);
    case 7: return hdb_mst_impl<7>(X, n, min_samples, alpha, core, u, v, w, rounds_host, L, s);
    default: return hdb_mst_impl<8>(X, n, min_samples, alpha, core, u, v, w, rounds_host, L, s);
  }
}

}  // extern "C"
