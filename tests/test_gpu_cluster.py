"""GPU parity: ball_query(PARTIAL_DENSE) and region_grow through the C ABI against the CPU oracle
(oracle/tpk_ref.py).  Bars: bit-exact neighbour tables and squared distances, identical instance partitions."""
import numpy as np
import pytest
import torch

from oracle import tpk_ref

pytestmark = pytest.mark.gpu


def _tpk():
    from panopticsegforlargescalepointcloud_b200 import tpk
    return tpk


def _cloud(seed, n=6000, scenes=2, clumps=True, spread=2.0):
    rng = np.random.default_rng(seed)
    pos = (rng.random((n, 3)).astype(np.float32) - 0.5) * spread
    if clumps:
        c = (rng.random((8, 3)).astype(np.float32) - 0.5) * spread
        k = n // 3
        pos[:k] = c[rng.integers(0, 8, k)] + rng.normal(0, 0.01, (k, 3)).astype(np.float32)
        pos = pos[rng.permutation(n)]
    batch = np.sort(rng.integers(0, scenes, n))
    return pos, batch


@pytest.mark.parametrize("nsample,radius", [(16, 0.1), (200, 0.1), (5, 0.2), (1, 0.05), (64, 0.02)])
def test_ball_query_bit_exact(cuda_device, nsample, radius):
    tpk = _tpk()
    pos, batch = _cloud(nsample)
    ref_idx, ref_d = tpk_ref.ball_query(radius, nsample, pos, batch, "grid")
    p = torch.from_numpy(pos).to(cuda_device)
    b = torch.from_numpy(batch).to(cuda_device)
    idx, d2 = tpk.ball_query(radius, nsample, p, p, mode="PARTIAL_DENSE", batch_x=b, batch_y=b)
    assert idx.dtype == torch.int64 and tuple(idx.shape) == (len(pos), nsample)
    assert np.array_equal(idx.cpu().numpy(), ref_idx)
    assert np.array_equal(d2.cpu().numpy(), ref_d)


def test_ball_query_separate_queries_and_empty(cuda_device):
    tpk = _tpk()
    pos, batch = _cloud(7, n=3000)
    rng = np.random.default_rng(1)
    qy = (rng.random((500, 3)).astype(np.float32) - 0.5) * 2.4   # some queries have no neighbour at all
    by = np.sort(rng.integers(0, 2, 500))
    p, b = torch.from_numpy(pos).to(cuda_device), torch.from_numpy(batch).to(cuda_device)
    idx, d2 = tpk.ball_query(0.15, 8, p, torch.from_numpy(qy).to(cuda_device), mode="PARTIAL_DENSE", batch_x=b,
                             batch_y=torch.from_numpy(by).to(cuda_device))
    idx = idx.cpu().numpy()
    r2 = np.float32(0.15) * np.float32(0.15)
    for q in range(500):
        ids = np.nonzero(batch == by[q])[0]
        d = pos[ids] - qy[q]
        dd = np.float32(np.float64(d[:, 2]) * d[:, 2] + np.float32(np.float64(d[:, 1]) * d[:, 1] + d[:, 0] * d[:, 0]))
        hits = ids[dd <= r2][:8]
        assert np.array_equal(idx[q][idx[q] >= 0], hits)
    e = torch.zeros((0, 3), device=cuda_device)
    i0, d0 = tpk.ball_query(0.1, 4, p, e, mode="PARTIAL_DENSE", batch_x=b, batch_y=torch.zeros(0, dtype=torch.long, device=cuda_device))
    assert tuple(i0.shape) == (0, 4)


def _check_partition(got, want):
    got_k = [tuple(int(i) for i in c.cpu().numpy()) for c in got]
    for c in got_k:
        assert list(c) == sorted(c)
    want_k = tpk_ref.partition_key(want)
    assert len(got_k) == len(want_k)
    assert got_k == want_k   # same clusters in the same (class, seed) order


@pytest.mark.parametrize("nsample,radius,mcs", [(16, 0.08, 5), (200, 0.08, 10), (4, 0.1, 3), (2, 0.3, 2)])
def test_region_grow_partition(cuda_device, nsample, radius, mcs):
    tpk = _tpk()
    pos, batch = _cloud(11 + nsample, n=8000, scenes=3)
    labels = np.random.default_rng(nsample).integers(-1, 5, len(pos))
    want = tpk_ref.region_grow(pos, labels, batch, [-1, 0, 3], nsample, radius, mcs, method="grid")
    got = tpk.region_grow(torch.from_numpy(pos).to(cuda_device), torch.from_numpy(labels).to(cuda_device),
                          torch.from_numpy(batch).to(cuda_device), ignore_labels=[-1, 0, 3], nsample=nsample,
                          radius=radius, min_cluster_size=mcs)
    assert len(want) > 0
    _check_partition(got, want)


def test_region_grow_edge_cases(cuda_device):
    tpk = _tpk()
    dev = cuda_device
    # everything ignored / empty input
    pos = torch.rand(100, 3, device=dev)
    lab = torch.zeros(100, dtype=torch.long, device=dev)
    bat = torch.zeros(100, dtype=torch.long, device=dev)
    assert tpk.region_grow(pos, lab, bat, ignore_labels=[0], radius=0.5, min_cluster_size=1) == []
    assert tpk.region_grow(pos[:0], lab[:0], bat[:0], radius=0.5) == []
    # all points identical: one cluster, lists truncated to the first nsample indices
    pos = torch.ones(500, 3, device=dev)
    out = tpk.region_grow(pos, lab[:1].repeat(500), bat[:1].repeat(500), nsample=16, radius=0.1, min_cluster_size=1)
    want = tpk_ref.region_grow(np.ones((500, 3), np.float32), np.zeros(500, int), np.zeros(500, int), [], 16, 0.1, 1)
    _check_partition(out, want)
    # far from the origin (coarser grid fallback) and negative coordinates
    p = (np.random.default_rng(0).random((2000, 3)).astype(np.float32) - 0.5) * 3 + np.float32(9000.0)
    want = tpk_ref.region_grow(p, np.zeros(2000, int), np.zeros(2000, int), [], 16, 0.2, 4, method="grid")
    got = tpk.region_grow(torch.from_numpy(p).to(dev), lab[:1].repeat(2000), bat[:1].repeat(2000), nsample=16,
                          radius=0.2, min_cluster_size=4)
    _check_partition(got, want)


def test_region_grow_scene_shape(cuda_device):
    """C1-shaped case: synthetic 50k cylinder, offset-shifted coordinates (dense clumps, nsample=200 truncation)."""
    tpk = _tpk()
    from panopticsegforlargescalepointcloud_b200 import scenes
    s = scenes.make_scene("urban", 50000, 0.2, 8.0, seed=1)
    off, _, logits = scenes.synthetic_head_outputs(s, seed=1)
    shifted = (s.pos + off).astype(np.float32)
    pred = logits.argmax(1)
    ignore = [-1] + list(scenes.stuff_classes("urban"))
    want = tpk_ref.region_grow(shifted, pred, s.batch, ignore, 200, 0.3, 10, method="grid")
    got = tpk.region_grow(torch.from_numpy(shifted).to(cuda_device), torch.from_numpy(pred).to(cuda_device),
                          torch.from_numpy(s.batch).to(cuda_device), ignore_labels=ignore, nsample=200, radius=0.3,
                          min_cluster_size=10)
    assert len(want) >= 5
    _check_partition(got, want)
    # raw-position call site omits nsample => 16 (PointGroup3heads.py:185-192)
    want = tpk_ref.region_grow(s.pos, pred, s.batch, ignore, 16, 0.3, 10, method="grid")
    got = tpk.region_grow(torch.from_numpy(s.pos).to(cuda_device), torch.from_numpy(pred).to(cuda_device),
                          torch.from_numpy(s.batch).to(cuda_device), ignore_labels=ignore, radius=0.3,
                          min_cluster_size=10)
    _check_partition(got, want)


def test_region_grow_full_size_c2(cuda_device):
    """BASELINE configs[1] size: 200k-voxel NPM3D-shape cylinder, shifted coordinates, nsample=200, r=0.18 --
    the C grid oracle still finishes in seconds, so this is exact partition parity at the benchmark size."""
    tpk = _tpk()
    from panopticsegforlargescalepointcloud_b200 import scenes
    s = scenes.make_scene("urban", 200000, 0.12, 16.0, seed=0)
    off, _, logits = scenes.synthetic_head_outputs(s, seed=0)
    shifted = (s.pos + off).astype(np.float32)
    pred = logits.argmax(1)
    ignore = [-1] + list(scenes.stuff_classes("urban"))
    want = tpk_ref.region_grow(shifted, pred, s.batch, ignore, 200, 0.18, 10, method="grid")
    got = tpk.region_grow(torch.from_numpy(shifted).to(cuda_device), torch.from_numpy(pred).to(cuda_device),
                          torch.from_numpy(s.batch).to(cuda_device), ignore_labels=ignore, nsample=200, radius=0.18,
                          min_cluster_size=10)
    assert len(want) >= 40
    _check_partition(got, want)
    # size-independent properties: clusters are disjoint, single-class, at least min_cluster_size
    allm = torch.cat(got)
    assert allm.unique().numel() == allm.numel()
    for c in got:
        assert c.numel() >= 10 and len(set(pred[c.cpu().numpy()])) == 1
