"""Pins oracle/tpk_ref.py (CPU restatement of tpk ball_query + region_grow) against independent definitions:
scipy cKDTree radius search, scipy connected components (untruncated case) and the min-ancestor theorem of
SURVEY App. C that the CUDA path relies on.  CPU only."""
import numpy as np
import pytest
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components
from scipy.spatial import cKDTree

from oracle import tpk_ref


def _cloud(seed, n=2500, scenes=2, clumps=True):
    rng = np.random.default_rng(seed)
    pos = rng.random((n, 3)).astype(np.float32) * 2
    if clumps:  # dense clumps like offset-shifted instances: many more than nsample points within r
        c = rng.random((6, 3)).astype(np.float32) * 2
        k = n // 3
        pos[:k] = c[rng.integers(0, 6, k)] + rng.normal(0, 0.01, (k, 3)).astype(np.float32)
        pos = pos[rng.permutation(n)]
    batch = np.sort(rng.integers(0, scenes, n))
    return pos, batch


@pytest.mark.parametrize("nsample", [4, 16, 200])
def test_scan_equals_grid_and_kdtree(nsample):
    pos, batch = _cloud(0)
    r = 0.12
    a, da = tpk_ref.ball_query(r, nsample, pos, batch, "scan")
    b, db = tpk_ref.ball_query(r, nsample, pos, batch, "grid")
    assert np.array_equal(a, b) and np.array_equal(da, db)
    # independent: KD-tree candidates (slightly larger radius), exact fp32 predicate, first nsample by index
    r2 = np.float32(r) * np.float32(r)
    for s in np.unique(batch):
        ids = np.nonzero(batch == s)[0]
        tree = cKDTree(pos[ids].astype(np.float64))
        for q in ids[:: max(len(ids) // 150, 1)]:
            cand = ids[np.array(sorted(tree.query_ball_point(pos[q].astype(np.float64), r * 1.01)), dtype=int)]
            d = (pos[cand] - pos[q]).astype(np.float32)
            d2 = np.array([np.float32(np.float64(x[2]) * x[2] + np.float32(np.float64(x[1]) * x[1] + np.float32(x[0] * x[0])))
                           for x in d], np.float32)
            hits = cand[d2 <= r2][:nsample]
            row = a[q][a[q] >= 0]
            assert np.array_equal(row, hits)
    assert (a >= 0).sum(1).max() == min(nsample, (a >= 0).sum(1).max())
    assert np.all(a[np.arange(len(a)), 0] <= np.arange(len(a)))  # self (d=0) is a hit unless truncated before it


def test_untruncated_grow_is_connected_components():
    pos, batch = _cloud(1, clumps=False)
    r, ns = 0.11, 64
    nbr, _ = tpk_ref.ball_query(r, ns, pos, batch, "grid")
    assert (nbr[:, -1] == -1).all(), "case must be untruncated"
    rows, cols = np.nonzero(nbr >= 0)
    g = coo_matrix((np.ones(len(rows)), (rows, nbr[rows, cols])), shape=(len(pos),) * 2)
    _, comp = connected_components(g, directed=False)
    want = {}
    for i, c in enumerate(comp):
        want.setdefault(c, []).append(i)
    want = sorted(tuple(v) for v in want.values() if len(v) >= 5)
    got = sorted(tpk_ref.partition_key(tpk_ref.grow_proximity(pos, batch, ns, r, 5, "grid")))
    assert got == want


@pytest.mark.parametrize("nsample", [3, 8])
def test_min_ancestor_theorem(nsample):
    """Seeded sequential BFS over TRUNCATED (directed) lists == grouping by the smallest index that reaches v."""
    pos, batch = _cloud(2, n=1500)
    r = 0.15
    nbr, _ = tpk_ref.ball_query(r, nsample, pos, batch, "grid")
    assert (nbr[:, -1] >= 0).any(), "case must be truncated"
    n = len(pos)
    label = np.arange(n)
    while True:  # literal fixpoint of label[j] = min(label[j], label[q]) over edges q -> j
        new = label.copy()
        for q in range(n):
            row = nbr[q][nbr[q] >= 0]
            np.minimum.at(new, row, label[q])
        if np.array_equal(new, label):
            break
        label = new
    groups = {}
    for i, l in enumerate(label):
        groups.setdefault(l, []).append(i)
    want = sorted(tuple(v) for v in groups.values() if len(v) >= 4)
    got = sorted(tpk_ref.partition_key(tpk_ref.grow_proximity(pos, batch, nsample, r, 4, "grid")))
    assert got == want


def test_region_grow_order_and_ignore():
    pos, batch = _cloud(3)
    labels = np.random.default_rng(0).integers(-1, 4, len(pos))
    cl = tpk_ref.region_grow(pos, labels, batch, ignore_labels=[-1, 0], nsample=16, radius=0.12, min_cluster_size=6)
    assert len(cl) > 3
    cls = [int(labels[c[0]]) for c in cl]
    assert cls == sorted(cls) and set(cls) <= {1, 2, 3}
    for c in cl:
        assert len(set(labels[c])) == 1 and len(set(batch[c])) == 1 and len(c) >= 6
    seeds = [(int(labels[c[0]]), int(c.min())) for c in cl]
    assert seeds == sorted(seeds)
    assert all(int(c[0]) == int(c.min()) for c in cl)  # a cluster is discovered from its smallest member
