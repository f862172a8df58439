/* ORACLE (test infrastructure, not product code).
 *
 * CPU restatement of torch-points-kernels 0.7.0 `ball_query(mode="PARTIAL_DENSE")` and the numba
 * `_grow_proximity_core` BFS behind `region_grow` (un-vendored dependency of the reference, absent from
 * /root/reference and from this image => PARITY UNPINNED; semantics frozen in SURVEY.md App. C).
 * Reference call sites: torch_points3d/models/panoptic/PointGroup3heads.py:166-174,296-304;
 * torch_points3d/core/spatial_ops/neighbour_finder.py:35-37.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (see oracle/build_oracle.py).  fmaf() is called
 * explicitly so that the squared distance is the value the upstream CUDA kernel computes
 * (`dist += d*d` contracted to FMA by nvcc): d2 = fma(dz,dz, fma(dy,dy, dx*dx)).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline float sqdist(const float* a, const float* b) {
  const float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
  return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

/* Literal scan: for every query (= support) point of scene s, walk the scene's points in ascending index
 * and keep the first nsample with d2 <= r*r.  ptr[s]..ptr[s+1] delimit scene s.  idx/dist are -1 filled. */
void ref_ball_query_scan(const float* x, const int64_t* ptr, int64_t n_scenes, float radius, int64_t nsample,
                         int64_t* idx, float* dist) {
  const float r2 = radius * radius;
  const int64_t n = ptr[n_scenes];
  for (int64_t i = 0; i < n * nsample; ++i) {
    idx[i] = -1;
    if (dist) dist[i] = -1.f;
  }
  for (int64_t s = 0; s < n_scenes; ++s)
    for (int64_t q = ptr[s]; q < ptr[s + 1]; ++q) {
      int64_t count = 0;
      for (int64_t p = ptr[s]; p < ptr[s + 1] && count < nsample; ++p) {
        const float d = sqdist(x + 3 * p, x + 3 * q);
        if (d <= r2) {
          idx[q * nsample + count] = p;
          if (dist) dist[q * nsample + count] = d;
          ++count;
        }
      }
    }
}

/* Same result through a uniform grid (cell edge = 1.0001 r): the CPU baseline that bench.py times.
 * Cells keep their points in ascending index; a query merges its 27 cells by index (27-way head scan). */
typedef struct { int64_t key; int64_t id; } kv_t;
static int kv_cmp(const void* a, const void* b) {
  const kv_t* x = (const kv_t*)a; const kv_t* y = (const kv_t*)b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  return x->id < y->id ? -1 : (x->id > y->id);
}
static int64_t lower_bound(const kv_t* a, int64_t n, int64_t key) {
  int64_t lo = 0, hi = n;
  while (lo < hi) { int64_t m = (lo + hi) >> 1; if (a[m].key < key) lo = m + 1; else hi = m; }
  return lo;
}
int ref_ball_query_grid(const float* x, const int64_t* ptr, int64_t n_scenes, float radius, int64_t nsample,
                        int64_t* idx, float* dist) {
  const float r2 = radius * radius;
  const double inv = 1.0 / ((double)radius * 1.0001);
  const int64_t n = ptr[n_scenes];
  for (int64_t i = 0; i < n * nsample; ++i) { idx[i] = -1; if (dist) dist[i] = -1.f; }
  kv_t* kv = (kv_t*)malloc(sizeof(kv_t) * (size_t)(n > 0 ? n : 1));
  int64_t* cell = (int64_t*)malloc(sizeof(int64_t) * 3 * (size_t)(n > 0 ? n : 1));
  if (!kv || !cell) return 1;
  for (int64_t s = 0; s < n_scenes; ++s) {
    const int64_t a = ptr[s], b = ptr[s + 1], m = b - a;
    for (int64_t i = a; i < b; ++i) {
      for (int d = 0; d < 3; ++d) cell[3 * i + d] = (int64_t)floor((double)x[3 * i + d] * inv) + (1 << 20);
      kv[i - a].key = (cell[3 * i + 2] << 42) | (cell[3 * i + 1] << 21) | cell[3 * i];
      kv[i - a].id = i;
    }
    qsort(kv, (size_t)m, sizeof(kv_t), kv_cmp);
    for (int64_t q = a; q < b; ++q) {
      int64_t cur[27], end[27];
      int c = 0;
      for (int dz = -1; dz <= 1; ++dz) for (int dy = -1; dy <= 1; ++dy) for (int dx = -1; dx <= 1; ++dx, ++c) {
        const int64_t key = ((cell[3 * q + 2] + dz) << 42) | ((cell[3 * q + 1] + dy) << 21) | (cell[3 * q] + dx);
        cur[c] = lower_bound(kv, m, key);
        end[c] = lower_bound(kv, m, key + 1);
      }
      int64_t count = 0;
      while (count < nsample) {
        int best = -1; int64_t best_id = INT64_MAX;
        for (c = 0; c < 27; ++c) if (cur[c] < end[c] && kv[cur[c]].id < best_id) { best_id = kv[cur[c]].id; best = c; }
        if (best < 0) break;
        ++cur[best];
        const float d = sqdist(x + 3 * best_id, x + 3 * q);
        if (d <= r2) { idx[q * nsample + count] = best_id; if (dist) dist[q * nsample + count] = d; ++count; }
      }
    }
  }
  free(kv); free(cell);
  return 0;
}

/* numba `_grow_proximity_core`: seeds in ascending index, LIFO stack, rows end at the first -1.
 * Writes kept clusters (size >= min_cluster_size) back to back into `members` in discovery order and their
 * sizes into `sizes`; returns the number of kept clusters. */
int64_t ref_grow_proximity(const int64_t* nbr, int64_t n, int64_t nsample, int64_t min_cluster_size,
                           int64_t* members, int64_t* sizes) {
  uint8_t* visited = (uint8_t*)calloc((size_t)(n > 0 ? n : 1), 1);
  int64_t* stack = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
  int64_t* cluster = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
  int64_t n_kept = 0, out = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (visited[i]) continue;
    int64_t sp = 0, cs = 0;
    visited[i] = 1; stack[sp++] = i; cluster[cs++] = i;
    while (sp) {
      const int64_t k = stack[--sp];
      const int64_t* row = nbr + k * nsample;
      for (int64_t e = 0; e < nsample; ++e) {
        const int64_t j = row[e];
        if (j == -1) break;
        if (!visited[j]) { visited[j] = 1; stack[sp++] = j; cluster[cs++] = j; }
      }
    }
    if (cs >= min_cluster_size) {
      memcpy(members + out, cluster, sizeof(int64_t) * (size_t)cs);
      out += cs; sizes[n_kept++] = cs;
    }
  }
  free(visited); free(stack); free(cluster);
  return n_kept;
}
