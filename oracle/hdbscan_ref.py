"""ORACLE (test infrastructure, not product code): HDBSCAN as the reference calls it, restated on the CPU.

Reference call site: torch_points3d/utils/hdbscan_cluster.py:8-13
    hdbscan.HDBSCAN(min_cluster_size=15, min_samples=5, core_dist_n_jobs=1, cluster_selection_epsilon=0.006)
        .fit_predict(X)                       (X float32 [n, D], D = 5 embeddings or 3 xyz)
and its per-scene fan-out `cluster_single` (:117-167), restated in `cluster_single` below.

PARITY UNPINNED against hdbscan==0.8.27 itself (un-vendored, conda-installed dependency, README.md:66; absent
from /root/reference and from this image; the reference ships no tests or golden vectors).  The pipeline is
the published HDBSCAN* algorithm and is pinned against scikit-learn 1.9's implementation,
which IS in this image (tests/test_oracle_hdbscan.py): identical float64 core distances, identical MST
weight multiset, identical label partitions on generic data, sklearn's own known-answer tests.
Core-distance rank: hdbscan 0.8.27 does not count the sample itself (rank min_samples + 1), scikit-learn does
(rank min_samples) -- `core_k()`; the default follows hdbscan, the sklearn pinning passes core_includes_self=True.

Canonical tie-break frozen by this project (DESIGN.md "HDBSCAN determinism"): edges of the mutual-reachability
graph are strictly ordered by (weight, min(a,b), max(a,b)); the MST under a strict order is unique; the
dendrogram merges edges in that order with left = component of min(a,b), right = component of max(a,b).
Upstream libraries sort MST edges with an unstable argsort, so on exact weight ties their result is
implementation-defined; any difference is confined to single bridge points at a split.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
import ctypes

import numpy as np

from . import build_oracle

_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build_oracle.build())
        P, I64, F64 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_double
        lib.ref_core_distances.argtypes = [P, I64, I64, I64, P]
        lib.ref_core_distances.restype = None
        lib.ref_mst_total_order.argtypes = [P, P, I64, I64, F64, P, P, P]
        lib.ref_mst_total_order.restype = ctypes.c_int
        lib.big_core_distances.argtypes = [P, I64, I64, I64, P, ctypes.c_int]
        lib.big_core_distances.restype = ctypes.c_int
        lib.big_mst_total_order.argtypes = [P, P, I64, I64, F64, P, P, P, ctypes.c_int]
        lib.big_mst_total_order.restype = ctypes.c_int
        _LIB = lib
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def core_distances(X, min_samples, threads=None):
    """Distance to the `min_samples`-th nearest sample COUNTING the sample itself (sklearn/cluster/_hdbscan/
    hdbscan.py:343-358: kneighbors(X, min_samples)[:, -1]); see core_k() for the rank each library uses."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    core = np.empty(X.shape[0], np.float64)
    if threads:      # oracle/c/hdbscan_big.c: the same loop with the rows split over threads (bit-identical)
        assert _lib().big_core_distances(_p(X), X.shape[0], X.shape[1], int(min_samples), _p(core), int(threads)) == 0
    else:
        _lib().ref_core_distances(_p(X), X.shape[0], X.shape[1], int(min_samples), _p(core))
    return core


def mst(X, core, alpha=1.0, threads=None):
    """Exact mutual-reachability MST under the canonical strict order.  -> u, v (u < v), w; sorted.
    threads: parallel Prim of oracle/c/hdbscan_big.c (same unique tree: the order is strict)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    n = X.shape[0]
    u = np.empty(max(n - 1, 0), np.int64)
    v = np.empty(max(n - 1, 0), np.int64)
    w = np.empty(max(n - 1, 0), np.float64)
    core = np.ascontiguousarray(core, dtype=np.float64)
    if threads:
        rc = _lib().big_mst_total_order(_p(X), _p(core), n, X.shape[1], float(alpha), _p(u), _p(v), _p(w), int(threads))
    else:
        rc = _lib().ref_mst_total_order(_p(X), _p(core), n, X.shape[1], float(alpha), _p(u), _p(v), _p(w))
    assert rc == 0
    return u, v, w


def single_linkage(u, v, w, n):
    """_linkage.pyx:226-273 make_single_linkage: rows (left, right, distance, size); new node ids n, n+1, ..."""
    parent = np.arange(2 * n - 1)
    size = np.ones(2 * n - 1, np.int64)

    def find(x):
        r = x
        while parent[r] != r:
            r = parent[r]
        while parent[x] != r:
            parent[x], x = r, parent[x]
        return r

    left = np.empty(n - 1, np.int64)
    right = np.empty(n - 1, np.int64)
    sz = np.empty(n - 1, np.int64)
    for i in range(n - 1):
        a, b = find(int(u[i])), find(int(v[i]))
        left[i], right[i] = a, b
        sz[i] = size[a] + size[b]
        parent[a] = parent[b] = n + i
        size[n + i] = sz[i]
    return left, right, np.asarray(w, np.float64), sz


def _bfs(left, right, n, root):
    out, queue = [], [root]
    while queue:
        out.extend(queue)
        nxt = []
        for x in queue:
            if x >= n:
                nxt.append(int(left[x - n]))
                nxt.append(int(right[x - n]))
        queue = nxt
    return out


def condense_tree(left, right, dist, size, min_cluster_size):
    """_tree.pyx:122-238 _condense_tree.  -> rows [(parent, child, lambda, child_size)] in upstream order."""
    n = len(left) + 1
    root = 2 * n - 2
    relabel = {root: n}
    next_label = n + 1
    ignore = np.zeros(2 * n - 1, bool)
    rows = []
    for node in _bfs(left, right, n, root):
        if ignore[node] or node < n:
            continue
        l, r, d = int(left[node - n]), int(right[node - n]), float(dist[node - n])
        lam = 1.0 / d if d > 0.0 else np.inf
        lc = int(size[l - n]) if l >= n else 1
        rc = int(size[r - n]) if r >= n else 1
        if lc >= min_cluster_size and rc >= min_cluster_size:
            relabel[l] = next_label
            next_label += 1
            rows.append((relabel[node], relabel[l], lam, lc))
            relabel[r] = next_label
            next_label += 1
            rows.append((relabel[node], relabel[r], lam, rc))
        else:
            for side, cnt, other in ((l, lc, r), (r, rc, l)):
                if cnt < min_cluster_size:
                    for sub in _bfs(left, right, n, side):
                        if sub < n:
                            rows.append((relabel[node], sub, lam, 1))
                        ignore[sub] = True
            if lc >= min_cluster_size:
                relabel[l] = relabel[node]
            if rc >= min_cluster_size:
                relabel[r] = relabel[node]
    return rows


def select_and_label(rows, n, cluster_selection_epsilon=0.0, allow_single_cluster=False):
    """_tree.pyx:240-280 (_compute_stability), :644+ (_get_clusters, EOM), :578-642 (epsilon_search),
    :433-513 (_do_labelling).  Returns int labels [n], -1 = noise, clusters numbered by ascending node id."""
    if allow_single_cluster:
        raise NotImplementedError("allow_single_cluster=True is not used by the reference")
    labels = np.full(n, -1, np.int64)
    if not rows:
        return labels
    root = min(r[0] for r in rows)
    births = {root: 0.0}
    for p, c, lam, s in rows:
        births[c] = lam
    births[root] = 0.0
    stability = {}
    for p, c, lam, s in rows:
        stability.setdefault(p, 0.0)
    for p, c, lam, s in rows:
        stability[p] += (lam - births[p]) * s
    ctree = [(p, c, lam, s) for p, c, lam, s in rows if s > 1]
    children = {}
    parent_of = {}
    for p, c, lam, s in ctree:
        children.setdefault(p, []).append(c)
        parent_of[c] = p
    node_list = sorted(stability.keys(), reverse=True)[:-1]
    is_cluster = {c: True for c in node_list}

    def descendants(c):
        out, q = [], list(children.get(c, []))
        while q:
            out.extend(q)
            q = [g for x in q for g in children.get(x, [])]
        return out

    for node in node_list:
        sub = 0.0
        for c in children.get(node, []):
            sub = sub + stability[c]
        if sub > stability[node]:
            is_cluster[node] = False
            stability[node] = sub
        else:
            for d in descendants(node):
                is_cluster[d] = False
    selected = {c for c in is_cluster if is_cluster[c]}

    if cluster_selection_epsilon != 0.0 and ctree:
        eps = float(cluster_selection_epsilon)
        picked, processed = [], set()
        for leaf in sorted(selected):  # canonical: ascending node id (upstream iterates a python set)
            if 1.0 / births[leaf] < eps:
                if leaf not in processed:
                    node = leaf
                    while True:
                        par = parent_of[node]
                        if par == root:
                            top = node
                            break
                        if 1.0 / births[par] > eps:
                            top = par
                            break
                        node = par
                    picked.append(top)
                    processed.update(descendants(top))
            else:
                picked.append(leaf)
        selected = set(picked)

    cluster_map = {c: i for i, c in enumerate(sorted(selected))}
    # a point carries the label of its lowest selected ancestor (union-find formulation upstream)
    owner = {}

    def lowest_selected(c):
        path = []
        x = c
        while x not in owner:
            if x in selected:
                owner[x] = x
                break
            if x == root:
                owner[x] = -1
                break
            path.append(x)
            x = parent_of[x]
        res = owner[x]
        for y in path:
            owner[y] = res
        return res

    for p, c, lam, s in rows:
        if s == 1 and c < n:
            o = lowest_selected(p)
            labels[c] = cluster_map[o] if o >= 0 else -1
    return labels


def core_k(n, min_samples, core_includes_self=False):
    """Rank (counting the sample itself) of the neighbour whose distance is the core distance.
    hdbscan 0.8.27 (the library the reference imports; recalled from its source, unverifiable here): `hdbscan()`
    clamps min_samples to [1, n-1], and every MST front end (KDTreeBoruvkaAlgorithm._compute_bounds:
    tree.query(k=min_samples+1)[:, min_samples]; _hdbscan_prims_kdtree: query(k=min_samples+1)[:, -1]; generic
    mutual_reachability: partition(d, min_points)[:, min_points]) takes the min_samples-th neighbour NOT counting
    the sample => rank min_samples + 1.  scikit-learn's HDBSCAN counts the sample (kneighbors(X, min_samples)[:, -1])
    => rank min_samples; `core_includes_self=True` selects that convention (used to pin against sklearn)."""
    if core_includes_self:
        if min_samples > n:
            raise ValueError("min_samples must be at most the number of samples")
        return int(min_samples)
    return max(min(n - 1, int(min_samples)), 1) + 1


def fit_predict(X, min_cluster_size=15, min_samples=5, cluster_selection_epsilon=0.006, alpha=1.0, return_parts=False,
                core_includes_self=False, threads=None):
    X = np.ascontiguousarray(X, dtype=np.float64)
    n = X.shape[0]
    if n < 2:
        raise ValueError("HDBSCAN requires more than one sample")
    core = core_distances(X, core_k(n, min_samples, core_includes_self), threads)
    u, v, w = mst(X, core, alpha, threads)
    left, right, dist, size = single_linkage(u, v, w, n)
    rows = condense_tree(left, right, dist, size, min_cluster_size)
    labels = select_and_label(rows, n, cluster_selection_epsilon)
    if return_parts:
        return labels, dict(core=core, u=u, v=v, w=w, rows=rows)
    return labels


def cluster_single(embeds, unique_in_batch, label_batch, local_ind, cluster_type, **kw):
    """torch_points3d/utils/hdbscan_cluster.py:117-167: per scene with > 3 points run HDBSCAN on the raw
    (un-normalised) block, emit local_ind[labels == l] for l != -1 in ascending l; scenes ascending."""
    out, types = [], []
    for s in unique_in_batch:
        mask = label_batch == s
        if mask.sum() > 3:
            ind = np.asarray(local_ind)[mask]
            lab = fit_predict(np.asarray(embeds)[mask], **kw)
            for l in np.unique(lab):
                if l == -1:
                    continue
                out.append(ind[lab == l])
                types.append(cluster_type)
    return out, types
