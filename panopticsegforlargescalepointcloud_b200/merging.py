"""Eval-time stitching on the device: nearest-neighbour back-projection and block merging.

Mirror of what the reference's tracker does on the CPU between two forward passes at test time
(torch_points3d/metrics/panoptic_tracker_pointgroup_npm3d.py):

  :384            knn(pos[origin_sub_ids], pos[originids], k=1)        block's voxelised points -> its original points
  :339-452        block_merging(...)                                  merge a block's instances into the running full-cloud
                                                                      labelling (numpy union1d / intersect1d per cluster)
  :592-633        knn(pos[has_prediction], pos, k=1) + stuff / distance (> 1 m) / size (< 10) filters

Here the per-point work (1-NN search on the voxel-hash grid of csrc/cluster.cu, contingency counts between new clusters and
old labels, label scatter) runs on the GPU; the strictly sequential decision loop over a block's clusters (a few hundred
iterations over a few candidates each -- later clusters see the labels earlier ones were given) stays on the host over
those counts.  Results are identical to the reference loop (tests/test_gpu_merging.py against oracle/merge_ref.py).
"""
import numpy as np
import torch

from . import _lib
from . import tpk
from ._lib import check, ptr, stream_ptr

MERGE_IOU = 0.1      # the reference ignores its th_merge argument and hard-codes 0.1 (:443)


def nearest(support_pos, query_pos, cell=None, max_ring=48):
    """-> (idx int64 [M], d2 fp32 [M]): nearest support point of every query (ties: smaller index)."""
    lib = _lib.load()
    xs, ys = tpk._as_f32_pos(support_pos), tpk._as_f32_pos(query_pos)
    dev = xs.device
    n, m = xs.shape[0], ys.shape[0]
    if n == 0:
        raise ValueError("nearest() needs at least one support point")
    if cell is None:
        # about two support points per occupied cell for surface-like clouds: extent / sqrt(n), bounded below
        ext = float((xs.max(0).values - xs.min(0).values).max())
        cell = max(ext / max(n, 1) ** 0.5 * 1.5, 1e-4)
    gid = torch.zeros(n, dtype=torch.int32, device=dev)
    grid = tpk._Grid(xs, gid, cell / 1.0001)       # (_Grid multiplies the radius by 1.0001)
    qpos = torch.empty((max(m, 1), 4), dtype=torch.float32, device=dev)
    qkeys = torch.empty(max(m, 1), dtype=torch.int64, device=dev)
    gq = torch.zeros(m, dtype=torch.int32, device=dev)
    check(lib.pgs_bq_pack_queries(ptr(ys), ptr(gq), m, grid.cell, ptr(qpos), ptr(qkeys), stream_ptr()))
    idx = torch.empty(max(m, 1), dtype=torch.int32, device=dev)
    d2 = torch.empty(max(m, 1), dtype=torch.float32, device=dev)
    check(lib.pgs_nn1_query(ptr(grid.spos), ptr(qpos), ptr(qkeys), m, ptr(grid.tkeys), ptr(grid.tvals), grid.cap,
                            ptr(grid.cell_start), ptr(grid.meta[0:2]), float(grid.cell), int(max_ring), ptr(idx), ptr(d2),
                            stream_ptr()))
    return idx[:m].long(), d2[:m]


def knn(x, y, k=1, batch_x=None, batch_y=None):
    """torch_geometric.nn.knn(x, y, k=1) shape: -> LongTensor [2, M] = (query ids, support ids)."""
    if k != 1 or batch_x is not None or batch_y is not None:
        raise NotImplementedError("only k=1 without batch vectors is on the reference path")
    idx, _ = nearest(x, y)
    return torch.stack([torch.arange(y.shape[0], device=idx.device), idx])


def block_merging(pos, originids, origin_sub_ids, pre_sub_ins, all_pre_ins, max_instance):
    """One call of the reference's block_merging (:339-452) with the full-cloud state on the device.
      pos            f32 [N_full, 3]   full cloud
      originids      i64 [n_o]         full-cloud ids of the block's ORIGINAL points
      origin_sub_ids i64 [n_s]         full-cloud ids of the block's voxelised points (the model's input rows)
      pre_sub_ins    i64 [n_s]         predicted instance id per voxelised point, -1 = none
      all_pre_ins    i64 [N_full]      running labelling, -1 = none -- updated IN PLACE
    -> (all_pre_ins, max_instance)."""
    dev = pos.device
    pre_sub_ins = torch.as_tensor(pre_sub_ins, device=dev).long()
    if not bool((pre_sub_ins != -1).any()):
        return all_pre_ins, max_instance
    originids = torch.as_tensor(originids, device=dev).long()
    origin_sub_ids = torch.as_tensor(origin_sub_ids, device=dev).long()
    nn_idx, _ = nearest(pos[origin_sub_ids], pos[originids])
    pre_ins = pre_sub_ins[nn_idx]                                   # new label of every original point of the block
    t_num = int(pre_ins.max()) + 1
    old = all_pre_ins[originids]
    has_old = old != -1
    n_has = int(has_old.sum())
    if n_has == 0:                                                  # nothing labelled yet in this block
        valid = pre_ins != -1
        all_pre_ins[originids[valid]] = pre_ins[valid] + max_instance
        return all_pre_ins, max_instance + t_num
    if n_has == originids.shape[0]:                                 # everything labelled already
        return all_pre_ins, max_instance
    # ---- counts the sequential loop needs (device) ----
    valid = pre_ins >= 0
    new_v, old_v = pre_ins[valid], old[valid]
    size_new = torch.bincount(new_v, minlength=t_num)                                   # |cluster ii|
    n_not_old = torch.bincount(new_v[old_v == -1], minlength=t_num)                     # its unlabelled points
    both = old_v != -1
    old_ids, old_inv = torch.unique(old[has_old], return_inverse=True)                  # old labels present in the block
    size_old = torch.bincount(old_inv, minlength=old_ids.shape[0])                      # |old label g| inside the block
    pair_key = new_v[both] * old_ids.shape[0] + torch.searchsorted(old_ids, old_v[both])
    pk, pc = torch.unique(pair_key, return_counts=True)                                 # contingency (ii, g) -> count
    size_new, n_not_old, old_ids_h, size_old_h, pk, pc = (t.cpu().numpy() for t in
                                                          (size_new, n_not_old, old_ids, size_old, pk, pc))
    n_old = len(old_ids_h)
    size_g = {int(g): int(c) for g, c in zip(old_ids_h, size_old_h)}
    starts = np.searchsorted(pk // max(n_old, 1), np.arange(t_num + 1))
    # ---- the reference's loop over the block's clusters, on counts instead of index sets ----
    assign = np.full(t_num, -2, np.int64)        # label given to the unlabelled points of cluster ii (-2: untouched)
    for ii in range(t_num):
        lo, hi = starts[ii], starts[ii + 1]
        if size_new[ii] == 0:                     # an unused id below the maximum: the reference's empty index sets take
            max_instance += 1                     # the "no old label" branch and still consume an instance number
            continue
        if hi == lo:                              # no point of the cluster carries an old label
            max_instance += 1
            assign[ii] = max_instance
            size_g[max_instance] = int(n_not_old[ii])
            continue
        if n_not_old[ii] == 0:
            continue
        best_iou, best_g = 0.0, 0
        for e in range(lo, hi):                   # ascending old label, strict '>' keeps the first maximum
            g = int(old_ids_h[pk[e] % n_old])
            inter = int(pc[e])
            iou = float(inter) / float(size_g[g] + int(size_new[ii]) - inter)
            if iou > best_iou:
                best_iou, best_g = iou, g
        if best_iou > MERGE_IOU:
            assign[ii] = best_g
            size_g[best_g] += int(n_not_old[ii])  # later clusters see these points under the old label
        else:
            max_instance += 1
            assign[ii] = max_instance
            size_g[max_instance] = int(n_not_old[ii])
    # ---- scatter (device) ----
    assign_d = torch.from_numpy(assign).to(dev)
    target = (old == -1) & valid
    lab = assign_d[pre_ins.clamp_min(0)]
    target &= lab != -2
    all_pre_ins[originids[target]] = lab[target]
    return all_pre_ins, max_instance


def back_project(pos, ins_pre, sem_pred, stuff_classes, max_dist=1.0, min_size=10):
    """:592-633: instance label of every full-cloud point = label of its nearest point that has a prediction; -1 for stuff
    classes, for points farther than `max_dist` from that neighbour, and for instances smaller than `min_size`."""
    dev = pos.device
    has = ins_pre != -1
    if not bool(has.any()):
        return torch.full_like(ins_pre, -1)
    src = torch.nonzero(has).squeeze(1)
    idx, d2 = nearest(pos[src], pos)
    full = ins_pre[src][idx].clone()
    stuff = torch.as_tensor(list(stuff_classes), device=dev, dtype=sem_pred.dtype)
    full[torch.isin(sem_pred, stuff)] = -1
    full[torch.sqrt(d2) > max_dist] = -1
    ok = full >= 0
    if bool(ok.any()):
        size = torch.bincount(full[ok])
        small = torch.zeros_like(ok)
        small[ok] = size[full[ok]] < min_size
        full[small] = -1
    return full
