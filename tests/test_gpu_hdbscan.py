"""GPU parity: HDBSCAN through the C ABI (device core distances + MST, host tree stage) against the CPU oracle
(oracle/hdbscan_ref.py) -- bit-identical float64 core distances and MST (edges, weights, order), identical
labels -- and against scikit-learn's HDBSCAN up to tie-bridge points."""
import numpy as np
import pytest
import torch

from oracle import hdbscan_ref as hr

pytestmark = pytest.mark.gpu


def _hdb():
    from panopticsegforlargescalepointcloud_b200 import hdbscan
    return hdbscan


def _blobs(seed, n=600, d=5, centers=6, std=0.15, spread=3.0):
    rng = np.random.default_rng(seed)
    mu = rng.normal(0, spread, (centers, d))
    X = mu[rng.integers(0, centers, n)] + rng.normal(0, std, (n, d))
    X[: n // 20] = rng.uniform(-2 * spread, 2 * spread, (n // 20, d))
    return X.astype(np.float32)


@pytest.mark.parametrize("n,d,k", [(700, 5, 5), (2500, 5, 5), (1500, 3, 5), (900, 2, 3), (400, 8, 12), (65, 5, 5), (33, 1, 2)])
def test_core_and_mst_bit_exact(cuda_device, n, d, k):
    hdb = _hdb()
    X = _blobs(n + d, n=n, d=d)
    ref_lab, parts = hr.fit_predict(X, 15, k, 0.006, return_parts=True)
    m = hdb.HDBSCAN(min_cluster_size=15, min_samples=k, cluster_selection_epsilon=0.006)
    lab = m.fit_predict(torch.from_numpy(X).to(cuda_device))
    assert np.array_equal(m.core_distances_.cpu().numpy(), parts["core"])
    u, v, w = (t.cpu().numpy() for t in m.mst_)
    assert np.array_equal(w, parts["w"])
    assert np.array_equal(u, parts["u"]) and np.array_equal(v, parts["v"])
    assert np.array_equal(lab.cpu().numpy(), ref_lab)


def test_duplicates_ties_and_grid_data(cuda_device):
    """Quantised coordinates: massive exact ties in distances -- the strict (w, min, max) order must still
    give the oracle's tree."""
    hdb = _hdb()
    rng = np.random.default_rng(3)
    X = (rng.integers(0, 12, (1200, 3)) * 0.25).astype(np.float32)   # many duplicates and equal distances
    ref_lab, parts = hr.fit_predict(X, 10, 4, 0.0, return_parts=True)
    m = hdb.HDBSCAN(min_cluster_size=10, min_samples=4, cluster_selection_epsilon=0.0)
    lab = m.fit_predict(torch.from_numpy(X).to(cuda_device))
    u, v, w = (t.cpu().numpy() for t in m.mst_)
    assert np.array_equal(w, parts["w"]) and np.array_equal(u, parts["u"]) and np.array_equal(v, parts["v"])
    assert np.array_equal(lab.cpu().numpy(), ref_lab)
    Z = np.zeros((50, 5), np.float32)
    assert np.array_equal(hdb.HDBSCAN(15, 5, 0.006).fit_predict(Z), hr.fit_predict(Z, 15, 5, 0.006))


@pytest.mark.parametrize("eps", [0.0, 0.006, 0.3])
def test_labels_match_sklearn_up_to_ties(cuda_device, eps):
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    from test_oracle_hdbscan import _same_partition_up_to_ties
    from sklearn.cluster import HDBSCAN as SK
    hdb = _hdb()
    X = _blobs(41, n=4000, d=5, centers=12)
    ref = SK(min_cluster_size=15, min_samples=5, cluster_selection_epsilon=eps, algorithm="kd_tree", copy=True).fit_predict(
        X.astype(np.float64))
    m = hdb.HDBSCAN(min_cluster_size=15, min_samples=5, cluster_selection_epsilon=eps, core_includes_self=True)
    got = m.fit_predict(X)                      # numpy in -> numpy out, like upstream
    assert isinstance(got, np.ndarray)
    u, v, w = (t.cpu().numpy() for t in m.mst_)
    assert _same_partition_up_to_ties(got, ref, dict(u=u, v=v, w=w))


def test_cluster_single_contract(cuda_device):
    hdb = _hdb()
    X = _blobs(5, n=900)
    batch = np.sort(np.random.default_rng(0).integers(0, 2, 900))
    local = np.arange(1000, 1900)
    want, wt = hr.cluster_single(X, [0, 1], batch, local, 7)
    got, gt = hdb.cluster_single(torch.from_numpy(X).to(cuda_device), torch.tensor([0, 1], device=cuda_device),
                                 torch.from_numpy(batch).to(cuda_device), torch.from_numpy(local).to(cuda_device), 7)
    assert gt == wt and len(got) == len(want)
    for a, b in zip(got, want):
        assert np.array_equal(a.cpu().numpy(), b)   # same clusters, same order, members ascending


def test_errors(cuda_device):
    hdb = _hdb()
    with pytest.raises(ValueError):
        hdb.HDBSCAN().fit_predict(torch.zeros(1, 5, device=cuda_device))
    with pytest.raises(Exception):
        hdb.HDBSCAN().fit_predict(torch.full((10, 3), float("nan"), device=cuda_device))
    with pytest.raises(Exception):
        hdb.HDBSCAN().fit_predict(torch.zeros(10, 3))          # CPU tensor: no CPU path


def test_scaling_property_20k(cuda_device):
    """Size the O(n^2) oracle cannot reach quickly: MST is a spanning tree, weights sorted, w >= both cores,
    total weight equals scikit-learn's Prim on a 20k subsample-free run."""
    hdb = _hdb()
    X = _blobs(77, n=20000, d=5, centers=40)
    m = hdb.HDBSCAN(min_cluster_size=15, min_samples=5, cluster_selection_epsilon=0.006)
    lab = m.fit_predict(torch.from_numpy(X).to(cuda_device)).cpu().numpy()
    u, v, w = (t.cpu().numpy() for t in m.mst_)
    core = m.core_distances_.cpu().numpy()
    assert np.all(np.diff(w) >= 0) and np.all(u < v)
    assert np.all(w >= np.maximum(core[u], core[v]))
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
    from sklearn.cluster._hdbscan._linkage import mst_from_data_matrix
    from sklearn.metrics import DistanceMetric
    ref = mst_from_data_matrix(X.astype(np.float64), core, DistanceMetric.get_metric("euclidean"), 1.0)
    assert np.array_equal(np.sort(ref["distance"]), w)
    assert len(set(lab) - {-1}) >= 30


def test_c3_shape_embeddings_100k(cuda_device):
    """FOR-instance-shaped case (configs[2], reduced to 100k thing points so the test stays short): synthetic 5-D
    embeddings of a forest cylinder.  Size-independent properties: spanning tree, sorted weights, w >= cores,
    every recovered cluster is dominated by one ground-truth tree."""
    hdb = _hdb()
    from panopticsegforlargescalepointcloud_b200 import scenes
    s = scenes.make_scene("forest", 120000, 0.06, 8.0, seed=2)
    _, emb, _ = scenes.synthetic_head_outputs(s, seed=2)
    thing = s.instance_mask
    X = emb[thing][:100000]
    inst = s.instance_labels[thing][:100000]
    m = hdb.HDBSCAN(min_cluster_size=15, min_samples=5, cluster_selection_epsilon=0.006)
    lab = m.fit_predict(torch.from_numpy(X).to(cuda_device)).cpu().numpy()
    u, v, w = (t.cpu().numpy() for t in m.mst_)
    core = m.core_distances_.cpu().numpy()
    assert len(w) == len(X) - 1 and np.all(np.diff(w) >= 0) and np.all(u < v)
    assert np.all(w >= np.maximum(core[u], core[v]))
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
    ids = [l for l in np.unique(lab) if l >= 0]
    assert len(ids) >= 5
    pure = [np.bincount(inst[lab == l]).max() / (lab == l).sum() for l in ids]
    assert np.mean(pure) > 0.95
    assert m.boruvka_rounds_ <= 32


@pytest.mark.parametrize("name", ["50k", "350k"])
def test_partition_exact_at_benchmark_size(cuda_device, name):
    """north_star: bit-exact instance partitions at the C3 size (~350 k thing points x 5-D per FOR-instance cylinder).
    The CPU oracle's answer (multithreaded exact kNN + Prim, oracle/c/hdbscan_big.c; ~5 min at 350 k) is frozen in
    tests/golden/hdbscan_big_*.npz by scripts/make_golden_hdbscan_big.py; inputs come from an integer-only generator
    and are bit-identical on every machine (checked by digest).  Core distances, the canonical MST (edges, weights,
    order) and the labels must all be EXACTLY the oracle's."""
    import os, sys
    sys.path.insert(0, os.path.dirname(__file__))
    import hdb_big_inputs as inp
    hdb = _hdb()
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "hdbscan_big_%s.npz" % name))
    X, owner = inp.make(name)
    assert inp.digest(X) == str(g["x_sha"])
    m = hdb.HDBSCAN(min_cluster_size=15, min_samples=5, cluster_selection_epsilon=0.006)
    lab = m.fit_predict(torch.from_numpy(X).to(cuda_device)).cpu().numpy()
    u, v, w = (t.cpu().numpy() for t in m.mst_)
    assert inp.digest(m.core_distances_.cpu().numpy()) == str(g["core_sha"])
    assert float(w.max()) == float(g["w_max"])
    assert inp.digest(u.astype(np.int32), v.astype(np.int32), w) == str(g["mst_sha"])
    assert np.array_equal(lab, g["labels"].astype(np.int64))
    assert m.n_clusters_ == int(g["n_clusters"])


def test_core_rank_convention(cuda_device):
    """Default = hdbscan 0.8.27's rank (sample not counted) == scikit-learn's convention with min_samples + 1."""
    hdb = _hdb()
    X = torch.from_numpy(_blobs(5, n=1500)).to(cuda_device)
    a = hdb.HDBSCAN(15, 5, 0.006)
    b = hdb.HDBSCAN(15, 6, 0.006, core_includes_self=True)
    la, lb = a.fit_predict(X), b.fit_predict(X)
    assert torch.equal(la, lb) and torch.equal(a.core_distances_, b.core_distances_)
    c = hdb.HDBSCAN(15, 5, 0.006, core_includes_self=True)
    c.fit_predict(X)
    assert bool((c.core_distances_ <= a.core_distances_).all()) and not torch.equal(c.core_distances_, a.core_distances_)
