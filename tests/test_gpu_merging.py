"""GPU parity of the eval-time stitching: nearest-neighbour back-projection and block merging on the device
(merging.py -> pgs_nn1_query) against the numpy restatement of the reference tracker (oracle/merge_ref.py;
torch_points3d/metrics/panoptic_tracker_pointgroup_npm3d.py:384,397-451,592-633).  Integer results: exact equality."""
import numpy as np
import pytest
import torch

from oracle import merge_ref as mr

pytestmark = pytest.mark.gpu


def _pkg():
    from panopticsegforlargescalepointcloud_b200 import merging, scenes
    return merging, scenes


@pytest.mark.parametrize("n_s,n_q,spread", [(20000, 5000, 1.0), (3000, 4000, 40.0), (1, 100, 1.0), (257, 1000, 0.01)])
def test_nearest_bit_exact(cuda_device, n_s, n_q, spread):
    """Dense, very sparse (ring expansion and the exhaustive fallback), single-support and tightly packed clouds."""
    merging, _ = _pkg()
    rng = np.random.default_rng(n_s + n_q)
    s = (rng.uniform(-5, 5, (n_s, 3)) * np.array([1, 1, 0.05])).astype(np.float32)
    q = (rng.uniform(-5 * spread, 5 * spread, (n_q, 3)) * np.array([1, 1, 0.05])).astype(np.float32)
    idx, d2 = merging.nearest(torch.from_numpy(s).to(cuda_device), torch.from_numpy(q).to(cuda_device))
    ridx, rd2 = mr.nearest(s, q)
    assert np.array_equal(idx.cpu().numpy(), ridx)
    assert np.array_equal(d2.cpu().numpy(), rd2)


def test_nearest_ties_go_to_the_smaller_index(cuda_device):
    merging, _ = _pkg()
    g = np.stack(np.meshgrid(np.arange(20), np.arange(20), [0]), -1).reshape(-1, 3).astype(np.float32)
    q = g[:-21] + np.array([0.5, 0.5, 0.0], np.float32)          # cell centres: four equidistant corners each
    idx, _ = merging.nearest(torch.from_numpy(g).to(cuda_device), torch.from_numpy(q).to(cuda_device))
    ridx, _ = mr.nearest(g, q)
    assert np.array_equal(idx.cpu().numpy(), ridx)
    k = merging.knn(torch.from_numpy(g).to(cuda_device), torch.from_numpy(q).to(cuda_device), k=1)
    assert k.shape == (2, len(q)) and torch.equal(k[1], idx)


def test_block_merging_sequence_matches_reference_loop(cuda_device):
    """A test area stitched from overlapping cylinder blocks, as the tracker does at eval: every block's voxelised points
    carry fresh instance ids; the running full-cloud labelling and max_instance must follow the reference exactly, block
    after block (each call sees the state the previous ones left, including the IoU > 0.1 joins)."""
    merging, scenes = _pkg()
    sc = scenes.make_scene("urban", 60000, 0.12, 12.0, seed=3)
    pos = sc.pos.astype(np.float32)
    inst = sc.instance_labels
    n = len(pos)
    rng = np.random.default_rng(1)
    ref_state = np.full(n, -1, np.int64)
    dev_state = torch.full((n,), -1, dtype=torch.long, device=cuda_device)
    pos_d = torch.from_numpy(pos).to(cuda_device)
    ref_max = dev_max = 0
    centres = [(-5, -4), (0, 0), (5, 3), (-2, 5), (4, -5), (0, 0)]
    joined = 0
    for bi, (cx, cy) in enumerate(centres):
        originids = np.nonzero((pos[:, 0] - cx) ** 2 + (pos[:, 1] - cy) ** 2 <= 6.0 ** 2)[0]
        sub = originids[rng.random(len(originids)) < 0.35]                       # the block's "voxelised" points
        ids, remap = np.unique(inst[sub], return_inverse=True)
        perm = rng.permutation(len(ids))
        pre = perm[remap].astype(np.int64)
        pre[inst[sub] == 0] = -1                                                  # stuff: no instance
        if bi == 3:
            pre[pre == pre.max()] = -1                                            # leave an unused id below the maximum ... 
            pre[0] = pre.max() + 2                                                # ... and a gap above it
        before = ref_max
        ref_state, ref_max = mr.block_merging(pos, originids, sub, pre, ref_state, ref_max)
        dev_state, dev_max = merging.block_merging(pos_d, torch.from_numpy(originids).to(cuda_device),
                                                   torch.from_numpy(sub).to(cuda_device),
                                                   torch.from_numpy(pre).to(cuda_device), dev_state, dev_max)
        assert dev_max == ref_max, bi
        assert np.array_equal(dev_state.cpu().numpy(), ref_state), bi
        joined += int(ref_max - before < len(ids))
    assert ref_max > 20 and joined >= 2            # later blocks did merge into instances of earlier ones
    # an all-unlabelled block changes nothing
    s2, m2 = merging.block_merging(pos_d, torch.arange(100, device=cuda_device), torch.arange(100, device=cuda_device),
                                   torch.full((100,), -1, device=cuda_device), dev_state.clone(), dev_max)
    assert m2 == dev_max and torch.equal(s2, dev_state)


def test_back_projection_matches_reference(cuda_device):
    merging, scenes = _pkg()
    sc = scenes.make_scene("urban", 50000, 0.12, 11.0, seed=9)
    pos = sc.pos.astype(np.float32)
    rng = np.random.default_rng(4)
    ins = np.where((rng.random(len(pos)) < 0.3) & (pos[:, 0] < 6.0), sc.instance_labels.astype(np.int64) - 1, -1)
    sem = sc.y.astype(np.int64)
    want = mr.back_project(pos, ins, sem, list(scenes.stuff_classes("urban")))
    got = merging.back_project(torch.from_numpy(pos).to(cuda_device), torch.from_numpy(ins).to(cuda_device),
                               torch.from_numpy(sem).to(cuda_device), list(scenes.stuff_classes("urban")))
    assert np.array_equal(got.cpu().numpy(), want)
    assert (want >= 0).sum() > 1000 and (want[pos[:, 0] > 7.5] == -1).all()     # beyond 1 m of any prediction: dropped
