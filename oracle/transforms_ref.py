"""CPU restatement of GridSampling3D / CylinderSampling -- TEST INFRASTRUCTURE ONLY.

Follows torch_points3d/core/data_transform/grid_transform.py:24-31,33-100,152-198 and transforms.py:385-435; the
un-vendored pieces (torch_cluster.grid_cluster, torch_geometric voxel_grid / consecutive_cluster) are restated from
memory: PARITY UNPINNED against those packages.  Checked against an independent dictionary-based definition in
tests/test_oracle_transforms.py."""
import numpy as np


def grid_sample(pos, size, batch=None):
    """-> (unique_pos_indices [M] in ascending voxel-id order = LAST row of each voxel, cluster [N], coords [N,3])"""
    coords = np.round(pos / np.float32(size))              # np.round == torch.round: half to even
    c = coords.astype(np.int64)
    if batch is not None:
        c = np.concatenate([c, np.asarray(batch, np.int64)[:, None]], 1)
    lo = c.min(0)
    nv = c.max(0) - lo + 1
    stride = np.concatenate([[1], np.cumprod(nv[:-1])]).astype(np.int64)
    vid = ((c - lo) * stride).sum(1)
    uniq, cluster = np.unique(vid, return_inverse=True)
    last = np.full(len(uniq), -1, np.int64)
    np.maximum.at(last, cluster, np.arange(len(pos)))
    return last, cluster, coords


def cylinder(pos, centre_xy, radius):
    d = pos[:, :2].astype(np.float64) - np.asarray(centre_xy, np.float32).astype(np.float64)
    return np.nonzero((d * d).sum(1) <= radius * radius)[0]
