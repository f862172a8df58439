// HDBSCAN back half: sorted mutual-reachability MST -> single-linkage dendrogram -> condensed tree ->
// stability / excess-of-mass selection (+ cluster_selection_epsilon) -> labels.
//
// Replaces the Cython tree code of hdbscan 0.8.27 that the reference reaches through
// torch_points3d/utils/hdbscan_cluster.py:8-13 (HDBSCAN(...).fit_predict).  The algorithm is the one
// scikit-learn ships as sklearn/cluster/_hdbscan/_linkage.pyx:226-273 (make_single_linkage) and
// _tree.pyx:122-238 (_condense_tree), :240-280 (_compute_stability), :578-642 (epsilon_search),
// :644+ (_get_clusters, "eom"), :433-513 (_do_labelling); oracle/hdbscan_ref.py restates it independently.
//
// This stage is an O(n alpha(n)) pointer-chasing pass over 2n-1 tree nodes with a strictly sequential
// dependency (union-find over edges in weight order); it runs on the host, between the device MST
// (csrc/hdbscan.cu) and the label upload, on pinned buffers.  It is part of the product design (DESIGN.md
// "HDBSCAN stages"), not a fallback: there is no device twin to fall back from.
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

#include "../../include/pgs_b200.h"

namespace {

struct UF {
  std::vector<int32_t> parent;
  explicit UF(size_t n) : parent(n) {
    for (size_t i = 0; i < n; ++i) parent[i] = (int32_t)i;
  }
  int32_t find(int32_t x) {
    int32_t r = x;
    while (parent[r] != r) r = parent[r];
    while (parent[x] != r) {
      const int32_t nx = parent[x];
      parent[x] = r;
      x = nx;
    }
    return r;
  }
};

}  // namespace

extern "C" int pgs_hdb_labels_host(const int32_t* u_host, const int32_t* v_host, const double* w_host, int64_t n,
                                   int32_t min_cluster_size, double cluster_selection_epsilon,
                                   int32_t* labels_host, int32_t* n_clusters_host) {
  if (n_clusters_host) *n_clusters_host = 0;
  if (n <= 0) return PGS_OK;
  for (int64_t i = 0; i < n; ++i) labels_host[i] = -1;
  if (n < 2 || min_cluster_size < 2) return n < 2 ? PGS_OK : PGS_ERR_INVALID;
  const int32_t N = (int32_t)n;
  const int32_t n_int = N - 1;  // internal dendrogram nodes N .. 2N-2

  // ---- single linkage (left = component of min(a,b), right = component of max(a,b)) ----
  std::vector<int32_t> left(n_int), right(n_int), size(n_int);
  {
    UF uf((size_t)2 * N - 1);
    std::vector<int32_t> sz((size_t)2 * N - 1, 1);
    for (int32_t i = 0; i < n_int; ++i) {
      const int32_t a = uf.find(u_host[i]), b = uf.find(v_host[i]);
      if (a == b) return PGS_ERR_INVALID;  // not a tree
      left[i] = a;
      right[i] = b;
      size[i] = sz[a] + sz[b];
      uf.parent[a] = uf.parent[b] = N + i;
      sz[N + i] = size[i];
    }
  }
  auto count_of = [&](int32_t node) { return node >= N ? size[node - N] : 1; };

  // ---- condensed tree (rows in upstream order: BFS over the dendrogram from the root) ----
  std::vector<int32_t> row_parent, row_child, row_size;
  std::vector<double> row_lambda;
  row_parent.reserve((size_t)N + 64);
  row_child.reserve((size_t)N + 64);
  row_size.reserve((size_t)N + 64);
  row_lambda.reserve((size_t)N + 64);
  std::vector<int32_t> relabel((size_t)2 * N - 1, -1);
  const int32_t root = 2 * N - 2;
  relabel[root] = N;
  int32_t next_label = N + 1;
  std::vector<int32_t> queue, sub;
  queue.reserve(N);
  queue.push_back(root);
  auto fall_out = [&](int32_t start, int32_t cluster, double lam) {
    sub.clear();
    sub.push_back(start);
    for (size_t h = 0; h < sub.size(); ++h) {
      const int32_t x = sub[h];
      if (x < N) {
        row_parent.push_back(cluster);
        row_child.push_back(x);
        row_lambda.push_back(lam);
        row_size.push_back(1);
      } else {
        sub.push_back(left[x - N]);
        sub.push_back(right[x - N]);
      }
    }
  };
  for (size_t h = 0; h < queue.size(); ++h) {
    const int32_t node = queue[h];
    if (node < N) continue;
    const int32_t l = left[node - N], r = right[node - N];
    const double d = w_host[node - N];
    const double lam = d > 0.0 ? 1.0 / d : std::numeric_limits<double>::infinity();
    const int32_t lc = count_of(l), rc = count_of(r);
    const int32_t me = relabel[node];
    if (lc >= min_cluster_size && rc >= min_cluster_size) {
      relabel[l] = next_label++;
      row_parent.push_back(me); row_child.push_back(relabel[l]); row_lambda.push_back(lam); row_size.push_back(lc);
      relabel[r] = next_label++;
      row_parent.push_back(me); row_child.push_back(relabel[r]); row_lambda.push_back(lam); row_size.push_back(rc);
      queue.push_back(l);
      queue.push_back(r);
    } else {
      if (lc < min_cluster_size) fall_out(l, me, lam);
      if (rc < min_cluster_size) fall_out(r, me, lam);
      if (lc >= min_cluster_size) { relabel[l] = me; queue.push_back(l); }
      if (rc >= min_cluster_size) { relabel[r] = me; queue.push_back(r); }
    }
  }
  // BFS order note: upstream walks the FULL level-order list and skips ignored nodes; enqueueing only the
  // surviving children visits the same nodes in the same relative order.  But a surviving left child must be
  // enqueued before a surviving right child even when the other one fell out -- handled above.

  const int32_t n_clusters_all = next_label - N;  // cluster ids N .. next_label-1, N = root
  const size_t n_rows = row_parent.size();
  std::vector<double> births(n_clusters_all, 0.0), stability(n_clusters_all, 0.0);
  std::vector<int32_t> cparent(n_clusters_all, -1);
  std::vector<int32_t> child0(n_clusters_all, -1), child1(n_clusters_all, -1);
  for (size_t i = 0; i < n_rows; ++i)
    if (row_size[i] > 1) {
      const int32_t c = row_child[i] - N, p = row_parent[i] - N;
      births[c] = row_lambda[i];
      cparent[c] = p;
      if (child0[p] < 0) child0[p] = c; else child1[p] = c;
    }
  births[0] = 0.0;
  for (size_t i = 0; i < n_rows; ++i) {
    const int32_t p = row_parent[i] - N;
    // one rounding per operation, in row order (matches the oracle bit for bit)
    const double term = (row_lambda[i] - births[p]) * (double)row_size[i];
    stability[p] = stability[p] + term;
  }

  // ---- excess of mass (root excluded: allow_single_cluster=False) ----
  std::vector<uint8_t> is_cluster(n_clusters_all, 1);
  is_cluster[0] = 0;
  std::vector<int32_t> stack;
  for (int32_t c = n_clusters_all - 1; c >= 1; --c) {
    double sub_stab = 0.0;
    if (child0[c] >= 0) sub_stab = sub_stab + stability[child0[c]];
    if (child1[c] >= 0) sub_stab = sub_stab + stability[child1[c]];
    if (sub_stab > stability[c]) {
      is_cluster[c] = 0;
      stability[c] = sub_stab;
    } else {
      stack.clear();
      if (child0[c] >= 0) stack.push_back(child0[c]);
      if (child1[c] >= 0) stack.push_back(child1[c]);
      while (!stack.empty()) {
        const int32_t x = stack.back();
        stack.pop_back();
        is_cluster[x] = 0;
        if (child0[x] >= 0) stack.push_back(child0[x]);
        if (child1[x] >= 0) stack.push_back(child1[x]);
      }
    }
  }

  // ---- cluster_selection_epsilon (ascending cluster id = canonical iteration order) ----
  if (cluster_selection_epsilon != 0.0 && n_clusters_all > 1) {
    const double eps = cluster_selection_epsilon;
    std::vector<uint8_t> picked(n_clusters_all, 0), processed(n_clusters_all, 0);
    for (int32_t c = 1; c < n_clusters_all; ++c) {
      if (!is_cluster[c]) continue;
      if (1.0 / births[c] < eps) {
        if (processed[c]) continue;
        int32_t node = c, top;
        for (;;) {
          const int32_t par = cparent[node];
          if (par == 0) { top = node; break; }
          if (1.0 / births[par] > eps) { top = par; break; }
          node = par;
        }
        picked[top] = 1;
        stack.clear();
        if (child0[top] >= 0) stack.push_back(child0[top]);
        if (child1[top] >= 0) stack.push_back(child1[top]);
        while (!stack.empty()) {
          const int32_t x = stack.back();
          stack.pop_back();
          processed[x] = 1;
          if (child0[x] >= 0) stack.push_back(child0[x]);
          if (child1[x] >= 0) stack.push_back(child1[x]);
        }
      } else {
        picked[c] = 1;
      }
    }
    is_cluster.swap(picked);
  }

  // ---- labels: lowest selected ancestor-or-self of the cluster a point fell out of ----
  std::vector<int32_t> label_of(n_clusters_all, -1), owner(n_clusters_all, -1);
  int32_t n_sel = 0;
  for (int32_t c = 1; c < n_clusters_all; ++c)
    if (is_cluster[c]) label_of[c] = n_sel++;
  for (int32_t c = 1; c < n_clusters_all; ++c) owner[c] = is_cluster[c] ? c : owner[cparent[c]];
  for (size_t i = 0; i < n_rows; ++i)
    if (row_size[i] == 1 && row_child[i] < N) {
      const int32_t o = owner[row_parent[i] - N];
      labels_host[row_child[i]] = o >= 0 ? label_of[o] : -1;
    }
  if (n_clusters_host) *n_clusters_host = n_sel;
  return PGS_OK;
}
