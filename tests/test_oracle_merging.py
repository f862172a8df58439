"""CPU checks of the eval-time stitching oracle (oracle/merge_ref.py): nearest() against scipy's cKDTree, block_merging on
hand-made cases that exercise each branch of the reference loop (metrics/panoptic_tracker_pointgroup_npm3d.py:397-451)."""
import numpy as np
from scipy.spatial import cKDTree

from oracle import merge_ref as mr


def test_nearest_matches_ckdtree():
    rng = np.random.default_rng(0)
    s = rng.uniform(-5, 5, (3000, 3)).astype(np.float32)
    q = rng.uniform(-6, 6, (1000, 3)).astype(np.float32)
    idx, d2 = mr.nearest(s, q)
    d, j = cKDTree(s.astype(np.float64)).query(q.astype(np.float64), k=1)
    assert np.array_equal(idx, j)                       # generic positions: no ties
    assert np.allclose(np.sqrt(d2), d, rtol=1e-5, atol=1e-6)
    # exact ties go to the smaller index
    s2 = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0]], np.float32)
    assert mr.nearest(s2, np.zeros((1, 3), np.float32))[0][0] == 0


def _line(n):
    return np.stack([np.arange(n, dtype=np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)], 1)


def test_block_merging_branches():
    pos = _line(40)
    ids = np.arange(40)
    fresh = np.full(40, -1, np.int64)
    # (a) nothing labelled yet: labels shifted by max_instance, which grows by the number of clusters
    pre = np.array([0] * 10 + [1] * 10 + [-1] * 20)
    out, mx = mr.block_merging(pos, ids, ids, pre, fresh, 5)
    assert mx == 7 and set(out[:10]) == {5} and set(out[10:20]) == {6} and set(out[20:]) == {-1}
    # (b) everything labelled: unchanged
    full = np.arange(40) % 3
    out2, mx2 = mr.block_merging(pos, ids, ids, pre, full, 9)
    assert np.array_equal(out2, full) and mx2 == 9
    # (c) partial overlap: cluster 0 overlaps old label 5 with IoU 5/15 > 0.1 -> its unlabelled points join 5;
    #     cluster 1 has no labelled point -> new label max_instance + 1 (note: + 1, unlike branch (a))
    state = fresh.copy()
    state[:5] = 5
    out3, mx3 = mr.block_merging(pos, ids, ids, np.array([0] * 15 + [1] * 10 + [-1] * 15), state, 5)
    assert set(out3[:15]) == {5} and set(out3[15:25]) == {6} and mx3 == 6 and set(out3[25:]) == {-1}
    # (d) overlap below the hard-coded 0.1: 1 of 30 points -> a new instance for the rest
    state = fresh.copy()
    state[0] = 2
    out4, mx4 = mr.block_merging(pos, ids, ids, np.array([0] * 30 + [-1] * 10), state, 2)
    assert out4[0] == 2 and set(out4[1:30]) == {3} and mx4 == 3


def test_back_project_filters():
    pos = _line(60)
    ins = np.full(60, -1, np.int64)
    ins[0:20:2] = 1            # every second point of the first 20 carries instance 1
    ins[30] = 2                # a lone labelled point: its instance stays below min_size
    sem = np.ones(60, np.int64)
    sem[5] = 0                 # stuff
    out = mr.back_project(pos, ins, sem, [0], max_dist=1.0, min_size=10)
    assert out[5] == -1 and set(out[np.r_[0:5, 6:20]]) == {1}
    assert set(out[29:32]) == {-1}                     # instance 2 covers 3 points < 10
    assert set(out[40:]) == {-1}                       # farther than 1 m from any labelled point
