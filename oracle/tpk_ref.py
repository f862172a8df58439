"""ORACLE (test infrastructure, not product code): torch-points-kernels 0.7.0 `ball_query(PARTIAL_DENSE)` and
`region_grow`, restated on the CPU (numpy + oracle/c/tpk_ref.c).

PARITY UNPINNED: torch-points-kernels is an un-vendored dependency of the reference (poetry.lock:2430-2431),
absent from /root/reference and from this image; the reference ships no tests for it.  Semantics follow
SURVEY.md App. C: neighbour lists = first `nsample` same-scene points in ascending index with
fma-accumulated squared distance <= r^2 (the upstream CUDA kernel), clusters = the sequential seeded stack-BFS
over those (directed, truncated) lists.  Independent checks in tests/test_oracle_cluster.py: scipy cKDTree
radius search and scipy connected_components for the untruncated case.

Reference call sites: torch_points3d/models/panoptic/PointGroup3heads.py:166-174,185-202,296-304,340-357.
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
        P, I64, F32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_float
        lib.ref_ball_query_scan.argtypes = [P, P, I64, F32, I64, P, P]
        lib.ref_ball_query_scan.restype = None
        lib.ref_ball_query_grid.argtypes = [P, P, I64, F32, I64, P, P]
        lib.ref_ball_query_grid.restype = ctypes.c_int
        lib.ref_grow_proximity.argtypes = [P, I64, I64, I64, P, P]
        lib.ref_grow_proximity.restype = I64
        _LIB = lib
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _ptr_from_batch(batch):
    batch = np.asarray(batch, dtype=np.int64)
    if batch.size and np.any(np.diff(batch) < 0):
        raise ValueError("batch must be sorted (PARTIAL_DENSE converts it to cumulative scene offsets)")
    n_scenes = int(batch.max()) + 1 if batch.size else 0
    ptr = np.zeros(n_scenes + 1, np.int64)
    np.cumsum(np.bincount(batch, minlength=n_scenes), out=ptr[1:])
    return ptr, n_scenes


def ball_query(radius, nsample, x, batch_x, method="scan", with_dist=True):
    """x == y (the only form on the hot path).  -> idx int64 [n, nsample] (-1 padded), dist2 f32 (-1 padded)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    ptr, n_scenes = _ptr_from_batch(batch_x)
    n = x.shape[0]
    idx = np.empty((n, nsample), np.int64)
    dist = np.empty((n, nsample), np.float32) if with_dist else None
    fn = _lib().ref_ball_query_scan if method == "scan" else _lib().ref_ball_query_grid
    fn(_p(x), _p(ptr), n_scenes, np.float32(radius), nsample, _p(idx), _p(dist) if with_dist else None)
    return idx, dist


def grow_proximity(pos, batch, nsample=16, radius=0.02, min_cluster_size=32, method="scan"):
    nbr, _ = ball_query(radius, nsample, pos, batch, method=method, with_dist=False)
    n = nbr.shape[0]
    members = np.empty(max(n, 1), np.int64)
    sizes = np.empty(max(n, 1), np.int64)
    k = _lib().ref_grow_proximity(_p(nbr), n, nsample, min_cluster_size, _p(members), _p(sizes))
    return np.split(members[:int(sizes[:k].sum())], np.cumsum(sizes[:k])[:-1]) if k else []


def region_grow(pos, labels, batch, ignore_labels=(), nsample=16, radius=0.02, min_cluster_size=32, method="scan"):
    """-> list of int64 index arrays (class ascending, then seed ascending; members in discovery order)."""
    pos = np.asarray(pos, dtype=np.float32)
    labels = np.asarray(labels)
    batch = np.asarray(batch)
    clusters = []
    ind = np.arange(pos.shape[0])
    for l in np.unique(labels):
        if l in ignore_labels:
            continue
        mask = labels == l
        local_ind = ind[mask]
        label_batch = batch[mask]
        _, remapped = np.unique(label_batch, return_inverse=True)
        for c in grow_proximity(pos[mask], remapped.reshape(-1), nsample, radius, min_cluster_size, method):
            clusters.append(local_ind[c])
    return clusters


def partition_key(clusters):
    """Order-insensitive-within-cluster canonical form used for partition equality."""
    return [tuple(sorted(int(i) for i in c)) for c in clusters]
