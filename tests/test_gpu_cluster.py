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


def test_instance_iou_matches_definition(cuda_device):
    """tpk.instance_iou (core/losses/panoptic_losses.py:37): IoU of every proposal against every ground-truth instance of
    ITS scene, columns = instances 1..M_s of scene 0, then of scene 1, ...; other scenes' columns are 0."""
    from panopticsegforlargescalepointcloud_b200 import tpk
    rng = np.random.default_rng(4)
    n_per, M = [900, 1200, 700], [5, 8, 3]
    batch = np.concatenate([np.full(n, s) for s, n in enumerate(n_per)])
    inst = np.concatenate([rng.integers(0, m + 1, n) for n, m in zip(n_per, M)])       # 0 = no instance
    for s, m in enumerate(M):                                                           # every id 1..M_s occurs
        inst[np.nonzero(batch == s)[0][:m]] = np.arange(1, m + 1)
    starts = np.concatenate([[0], np.cumsum(n_per)])
    props = []
    for s in range(3):
        for _ in range(6):
            k = int(rng.integers(5, 200))
            props.append(np.sort(rng.choice(np.arange(starts[s], starts[s + 1]), k, replace=False)))
    want = np.zeros((len(props), sum(M)), np.float32)
    col0 = np.concatenate([[0], np.cumsum(M)])
    for p, idx in enumerate(props):
        s = batch[idx[0]]
        for g in range(1, M[s] + 1):
            G = np.nonzero((batch == s) & (inst == g))[0]
            inter = len(np.intersect1d(idx, G))
            want[p, col0[s] + g - 1] = inter / float(len(idx) + len(G) - inter)
    got = tpk.instance_iou([torch.from_numpy(p).to(cuda_device) for p in props], torch.from_numpy(inst).to(cuda_device),
                           torch.from_numpy(batch).to(cuda_device))
    assert tuple(got.shape) == want.shape and np.allclose(got.cpu().numpy(), want, atol=1e-6)


def test_get_instances_matches_reference_nms(cuda_device):
    """PanopticResults.get_instances (models/panoptic/structure_3heads.py:28-71): dense-mask cross IoU + greedy NMS in
    descending score order + size / score filters, restated with numpy; the product uses a sparse incidence matrix."""
    from panopticsegforlargescalepointcloud_b200 import panoptic
    rng = np.random.default_rng(5)
    N, n_prop = 6000, 40
    clusters = []
    for i in range(n_prop):
        c0 = int(rng.integers(0, N - 600))
        clusters.append(np.sort(rng.choice(np.arange(c0, c0 + 600), int(rng.integers(50, 400)), replace=False)))
    scores = rng.permutation(n_prop).astype(np.float32) / n_prop          # distinct: no tie ambiguity in the argsort
    masks = np.zeros((n_prop, N), np.float32)
    for i, c in enumerate(clusters):
        masks[i, c] = 1
    inter = masks @ masks.T
    num = masks.sum(1)
    cross = inter / (num[:, None] + num[None, :] - inter)
    ixs = list(np.argsort(scores)[::-1])
    pick = []
    while ixs:
        i = ixs.pop(0)
        pick.append(i)
        ixs = [j for j in ixs if not cross[i, j] > 0.3]
    want = [i for i in pick if len(clusters[i]) > 100 and scores[i] > 0.5]
    res = panoptic.PanopticResults(semantic_logits=torch.zeros(N, 2, device=cuda_device), offset_logits=None, embed_logits=None,
                                   cluster_scores=torch.from_numpy(scores).to(cuda_device), mask_scores=None,
                                   clusters=[torch.from_numpy(c).to(cuda_device) for c in clusters], cluster_type=None)
    ids, out = res.get_instances(nms_threshold=0.3, min_cluster_points=100, min_score=0.5)
    assert [int(i) for i in ids] == [int(i) for i in want]
    for o, i in zip(out, want):
        assert np.array_equal(o.cpu().numpy(), clusters[i])
