"""Pins oracle/hdbscan_ref.py against scikit-learn's HDBSCAN (present in this image; the un-vendored hdbscan
0.8.27 the reference imports is not): float64 core distances, MST weight multiset, label partitions, and
sklearn's own known-answer test (sklearn/cluster/tests/test_hdbscan.py:27-43: make_blobs(200, random_state=10)
=> 3 clusters).  CPU only."""
import numpy as np
import pytest
from sklearn.cluster import HDBSCAN
from sklearn.cluster._hdbscan._linkage import mst_from_data_matrix
from sklearn.datasets import make_blobs
from sklearn.metrics import DistanceMetric
from sklearn.neighbors import NearestNeighbors
from sklearn.utils import shuffle

from oracle import hdbscan_ref as hr


def _same_partition(a, b):
    """Equal as partitions (cluster numbering may differ), noise == noise."""
    a, b = np.asarray(a), np.asarray(b)
    if not np.array_equal(a == -1, b == -1):
        return False
    m = {}
    for x, y in zip(a, b):
        if m.setdefault(x, y) != y:
            return False
    return len(set(m.values())) == len(m)


def _bridge_points(parts):
    """Points whose membership is tie-order dependent: they sit on >= 2 MST edges of exactly equal weight
    (oracle/hdbscan_ref.py header: upstream sorts tied edges with an unstable argsort)."""
    u, v, w = parts["u"], parts["v"], parts["w"]
    tied = np.zeros(len(w), bool)
    tied[1:] |= w[1:] == w[:-1]
    tied[:-1] |= w[:-1] == w[1:]
    pts = np.concatenate([u[tied], v[tied]])
    ids, cnt = np.unique(pts, return_counts=True)
    return set(ids[cnt >= 2].tolist())


def _same_partition_up_to_ties(got, ref, parts, max_bridge_diffs=3):
    """Every point that is not a tie-bridge must agree exactly; among bridge points at most a handful may
    land on the other side of a split."""
    got, ref = np.asarray(got), np.asarray(ref)
    bridge = np.zeros(len(got), bool)
    bridge[list(_bridge_points(parts))] = True
    if not _same_partition(got[~bridge], ref[~bridge]):
        return False
    m = {-1: -1}
    for x, y in zip(got[~bridge], ref[~bridge]):
        m[x] = y
    bad = sum(1 for x, y in zip(got[bridge], ref[bridge]) if m.get(x, -2) != y)
    return bad <= max_bridge_diffs


def _blobs(seed, n=600, d=5, centers=6, std=0.15, spread=3.0):
    rng = np.random.default_rng(seed)
    mu = rng.normal(0, spread, (centers, d))
    X = mu[rng.integers(0, centers, n)] + rng.normal(0, std, (n, d))
    X[: n // 20] = rng.uniform(-2 * spread, 2 * spread, (n // 20, d))  # background noise
    return X.astype(np.float32)


def test_sklearn_known_answer():
    X, y = make_blobs(n_samples=200, random_state=10)
    X, y = shuffle(X, y, random_state=7)
    from sklearn.preprocessing import StandardScaler
    X = StandardScaler().fit_transform(X)
    ref = HDBSCAN().fit_predict(X)                     # defaults: min_cluster_size=5, min_samples=None -> 5
    got = hr.fit_predict(X, min_cluster_size=5, min_samples=5, cluster_selection_epsilon=0.0, core_includes_self=True)
    assert len(set(got) - {-1}) == 3
    assert _same_partition(got, ref)


@pytest.mark.parametrize("seed,d", [(0, 5), (1, 5), (2, 3), (3, 2)])
def test_core_and_mst_weights_match_sklearn(seed, d):
    X = _blobs(seed, d=d).astype(np.float64)
    core = hr.core_distances(X, 5)
    ref_core = NearestNeighbors(n_neighbors=5, algorithm="kd_tree").fit(X).kneighbors(X, 5)[0][:, -1]
    assert np.array_equal(core, ref_core)
    u, v, w = hr.mst(X, core)
    ref = mst_from_data_matrix(X, np.ascontiguousarray(ref_core), DistanceMetric.get_metric("euclidean"), 1.0)
    assert np.array_equal(np.sort(w), np.sort(ref["distance"]))   # MST weight multiset is tie-independent
    assert np.all(u < v) and np.all(np.diff(w) >= 0)
    # it is a spanning tree
    parent = np.arange(len(X))
    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x
    for a, b in zip(u, v):
        ra, rb = find(a), find(b)
        assert ra != rb
        parent[ra] = rb


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("eps", [0.0, 0.006, 0.3])
def test_labels_match_sklearn(seed, eps):
    X = _blobs(10 + seed, n=500 + 40 * seed, d=5 if seed % 2 == 0 else 3)
    ref = HDBSCAN(min_cluster_size=15, min_samples=5, cluster_selection_epsilon=eps, algorithm="kd_tree").fit_predict(
        X.astype(np.float64))
    got, parts = hr.fit_predict(X, 15, 5, eps, return_parts=True, core_includes_self=True)
    assert len(set(ref) - {-1}) >= 2
    assert _same_partition_up_to_ties(got, ref, parts)


def test_degenerate_inputs():
    X = np.zeros((40, 3), np.float32)                 # all duplicates: distances 0, lambda = inf
    got = hr.fit_predict(X, 15, 5, 0.006, core_includes_self=True)
    ref = HDBSCAN(min_cluster_size=15, min_samples=5, cluster_selection_epsilon=0.006).fit_predict(X.astype(np.float64))
    assert _same_partition(got, ref)
    X = _blobs(3, n=20)                               # fewer points than any cluster could hold
    assert (hr.fit_predict(X, 15, 5, 0.006) == -1).all()


def test_cluster_single_contract():
    X = _blobs(5, n=400)
    batch = np.sort(np.random.default_rng(0).integers(0, 2, 400))
    local = np.arange(1000, 1400)
    out, types = hr.cluster_single(X, [0, 1], batch, local, 7)
    assert len(out) >= 2 and types == [7] * len(out)
    for c in out:
        assert len(set(batch[c - 1000])) == 1 and len(c) >= 15


def test_core_rank_convention():
    """hdbscan 0.8.27 (what the reference imports) takes the min_samples-th neighbour NOT counting the sample;
    scikit-learn counts it.  Default = hdbscan's rank, which is scikit-learn's with min_samples + 1."""
    X = _blobs(21, n=700)
    n = len(X)
    assert hr.core_k(n, 5) == 6 and hr.core_k(n, 5, core_includes_self=True) == 5
    assert hr.core_k(4, 5) == 4            # hdbscan(): min_samples = min(n - 1, min_samples), then + 1 for the rank
    lab, parts = hr.fit_predict(X, 15, 5, 0.006, return_parts=True)
    ref_core = NearestNeighbors(n_neighbors=6, algorithm="kd_tree").fit(X.astype(np.float64)).kneighbors(
        X.astype(np.float64), 6)[0][:, -1]          # 5 other samples + the sample itself
    assert np.array_equal(parts["core"], ref_core)
    assert np.array_equal(lab, hr.fit_predict(X, 15, 6, 0.006, core_includes_self=True))
    ref = HDBSCAN(min_cluster_size=15, min_samples=6, cluster_selection_epsilon=0.006, algorithm="kd_tree").fit_predict(
        X.astype(np.float64))
    lab2, parts2 = hr.fit_predict(X, 15, 6, 0.006, return_parts=True, core_includes_self=True)
    assert _same_partition_up_to_ties(lab2, ref, parts2)


def test_threaded_oracle_is_bit_identical():
    X = _blobs(31, n=3000, centers=9)
    a, pa = hr.fit_predict(X, 15, 5, 0.006, return_parts=True)
    b, pb = hr.fit_predict(X, 15, 5, 0.006, return_parts=True, threads=4)
    assert np.array_equal(a, b)
    for k in ("core", "u", "v", "w"):
        assert np.array_equal(pa[k], pb[k])


def test_benchmark_size_golden_50k():
    """tests/golden/hdbscan_big_50k.npz (scripts/make_golden_hdbscan_big.py) is what the oracle computes today, and the
    integer-only input generator reproduces the frozen inputs bit for bit on this machine."""
    import os, sys
    sys.path.insert(0, os.path.dirname(__file__))
    import hdb_big_inputs as inp
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "hdbscan_big_50k.npz"))
    X, owner = inp.make("50k")
    assert inp.digest(X) == str(g["x_sha"])
    lab, parts = hr.fit_predict(X, 15, 5, 0.006, return_parts=True, threads=os.cpu_count())
    assert np.array_equal(lab, g["labels"].astype(np.int64))
    assert inp.digest(parts["core"]) == str(g["core_sha"])
    assert inp.digest(parts["u"].astype(np.int32), parts["v"].astype(np.int32), parts["w"]) == str(g["mst_sha"])
    # the recovered clusters are the planted trees
    ids = [l for l in np.unique(lab) if l >= 0]
    pure = []
    for l in ids:
        own = owner[(lab == l) & (owner >= 0)]
        pure.append(np.bincount(own).max() / len(own) if len(own) else 0.0)    # (a cluster of background points: 0)
    assert len(ids) >= 20 and np.mean(pure) > 0.9 and np.median(pure) > 0.99
