/* ORACLE (test infrastructure, not product code).
 *
 * O(n^2) CPU restatement of the numeric front half of HDBSCAN as the reference calls it
 * (torch_points3d/utils/hdbscan_cluster.py:8-13 -> hdbscan.HDBSCAN(min_cluster_size=15, min_samples=5,
 * cluster_selection_epsilon=0.006).fit_predict; hdbscan 0.8.27 is un-vendored and absent => PARITY UNPINNED
 * against it; pinned instead against scikit-learn's HDBSCAN, which is present in this image:
 * sklearn/cluster/_hdbscan/hdbscan.py:343-360, _linkage.pyx:111-223).
 *
 * All arithmetic is float64 with one rounding per operation (build with -ffp-contract=off), in the operation
 * order of sklearn's euclidean rdist/dist:  d = sqrt(((t0^2 + t1^2) + t2^2) + ...).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static inline double sqdist_d(const double* a, const double* b, int64_t D) {
  double s = 0.0;
  for (int64_t j = 0; j < D; ++j) { const double t = a[j] - b[j]; s += t * t; }
  return s;
}

/* core[i] = distance to the k-th nearest sample counting i itself (sklearn kneighbors(X, k)[:, -1]) */
void ref_core_distances(const double* X, int64_t n, int64_t D, int64_t k, double* core) {
  double* best = (double*)malloc(sizeof(double) * (size_t)k);
  for (int64_t i = 0; i < n; ++i) {
    int64_t m = 0;
    for (int64_t j = 0; j < n; ++j) {
      const double d2 = sqdist_d(X + i * D, X + j * D, D);
      if (m < k) {
        int64_t p = m++;
        while (p > 0 && best[p - 1] > d2) { best[p] = best[p - 1]; --p; }
        best[p] = d2;
      } else if (d2 < best[k - 1]) {
        int64_t p = k - 1;
        while (p > 0 && best[p - 1] > d2) { best[p] = best[p - 1]; --p; }
        best[p] = d2;
      }
    }
    core[i] = sqrt(best[k - 1]);
  }
  free(best);
}

/* Exact MST of the complete mutual-reachability graph, w(a,b) = max(core_a, core_b, d(a,b)/alpha), under the
 * STRICT TOTAL ORDER (w, min(a,b), max(a,b)) -- the canonical tie-break this project freezes:
 * with a strict order the MST is unique, so any correct algorithm (Prim here, Boruvka on the GPU) returns
 * the same edge set.  Output edges are sorted by that order, u < v. */
typedef struct { double w; int64_t a, b; } edge_t;
static inline int key_less(double w1, int64_t a1, int64_t b1, double w2, int64_t a2, int64_t b2) {
  if (w1 != w2) return w1 < w2;
  if (a1 != a2) return a1 < a2;
  return b1 < b2;
}
static int edge_cmp(const void* x, const void* y) {
  const edge_t* e = (const edge_t*)x; const edge_t* f = (const edge_t*)y;
  if (key_less(e->w, e->a, e->b, f->w, f->a, f->b)) return -1;
  if (key_less(f->w, f->a, f->b, e->w, e->a, e->b)) return 1;
  return 0;
}
int ref_mst_total_order(const double* X, const double* core, int64_t n, int64_t D, double alpha,
                        int64_t* u, int64_t* v, double* w) {
  if (n < 2) return 0;
  uint8_t* in_tree = (uint8_t*)calloc((size_t)n, 1);
  double* bw = (double*)malloc(sizeof(double) * (size_t)n);
  int64_t* ba = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
  int64_t* bb = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
  edge_t* edges = (edge_t*)malloc(sizeof(edge_t) * (size_t)(n - 1));
  if (!in_tree || !bw || !ba || !bb || !edges) return 1;
  for (int64_t j = 0; j < n; ++j) { bw[j] = INFINITY; ba[j] = bb[j] = INT64_MAX; }
  int64_t cur = 0;
  for (int64_t i = 0; i < n - 1; ++i) {
    in_tree[cur] = 1;
    int64_t next = -1;
    for (int64_t j = 0; j < n; ++j) {
      if (in_tree[j]) continue;
      double d = sqrt(sqdist_d(X + cur * D, X + j * D, D)) / alpha;
      double m = core[cur] > core[j] ? core[cur] : core[j];
      if (d > m) m = d;
      const int64_t a = cur < j ? cur : j, b = cur < j ? j : cur;
      if (key_less(m, a, b, bw[j], ba[j], bb[j])) { bw[j] = m; ba[j] = a; bb[j] = b; }
      if (next < 0 || key_less(bw[j], ba[j], bb[j], bw[next], ba[next], bb[next])) next = j;
    }
    edges[i].w = bw[next]; edges[i].a = ba[next]; edges[i].b = bb[next];
    cur = next;
  }
  qsort(edges, (size_t)(n - 1), sizeof(edge_t), edge_cmp);
  for (int64_t i = 0; i < n - 1; ++i) { u[i] = edges[i].a; v[i] = edges[i].b; w[i] = edges[i].w; }
  free(in_tree); free(bw); free(ba); free(bb); free(edges);
  return 0;
}
