"""ORACLE (test infrastructure, not product code): the reference tracker's eval-time stitching, restated in numpy.

  torch_points3d/metrics/panoptic_tracker_pointgroup_npm3d.py:384   knn(pos[origin_sub_ids], pos[originids], k=1)
  torch_points3d/metrics/panoptic_tracker_pointgroup_npm3d.py:397-451 block_merging (the label logic; the PLY dumps and the
                                                                     `viz/` directory of :341-395 are not restated)
  torch_points3d/metrics/panoptic_tracker_pointgroup_npm3d.py:592-633 back-projection onto the full cloud + filters

torch_geometric's knn is absent from this image (PARITY UNPINNED against it); nearest() is the definition frozen here:
smallest fp32 d2 = fma(dz,dz,fma(dy,dy,dx*dx)), ties to the smaller support index, pinned to scipy cKDTree in the tests.
Only tests/ may import this module.
"""
import numpy as np


def nearest(support, query, chunk=2048):
    """Brute force, float64 emulation of the fp32 fma chain is not needed: products of fp32 differences are exact in
    float64, so the fp32 fma result is the float64 sum rounded once per fma -- reproduced with float32 casts."""
    s = np.asarray(support, np.float32)
    q = np.asarray(query, np.float32)
    idx = np.empty(len(q), np.int64)
    d2o = np.empty(len(q), np.float32)
    for a in range(0, len(q), chunk):
        d = (s[None, :, :].astype(np.float32) - q[a:a + chunk, None, :]).astype(np.float32).astype(np.float64)
        t = (d[..., 0] * d[..., 0]).astype(np.float32).astype(np.float64)          # dx*dx rounded to fp32
        t = (d[..., 1] * d[..., 1] + t).astype(np.float32).astype(np.float64)      # fma(dy, dy, .)
        t = (d[..., 2] * d[..., 2] + t).astype(np.float32)                         # fma(dz, dz, .)
        j = np.argmin(t, axis=1)                                                   # first minimum = smallest index
        idx[a:a + chunk] = j
        d2o[a:a + chunk] = t[np.arange(len(j)), j]
    return idx, d2o


def block_merging(pos, originids, origin_sub_ids, pre_sub_ins, all_pre_ins, max_instance):
    """:397-451, index sets and all (np.union1d / np.intersect1d per candidate), mutating a copy of all_pre_ins."""
    all_pre_ins = np.array(all_pre_ins, copy=True)
    pre_sub_ins = np.asarray(pre_sub_ins)
    if not np.any(pre_sub_ins != -1):
        return all_pre_ins, max_instance
    x_idx, _ = nearest(pos[origin_sub_ids], pos[originids])
    pre_ins = pre_sub_ins[x_idx]
    t_num_clusters = int(np.max(pre_ins)) + 1
    idx = np.argwhere(all_pre_ins[originids] != -1)
    idx2 = np.argwhere(all_pre_ins[originids] == -1)
    if len(idx) == 0:
        mask_valid = pre_ins != -1
        all_pre_ins[originids[mask_valid]] = pre_ins[mask_valid] + max_instance
        max_instance = max_instance + t_num_clusters
    elif len(idx2) == 0:
        return all_pre_ins, max_instance
    else:
        new_label = pre_ins.reshape(-1)
        for ii in range(t_num_clusters):
            members = originids[np.argwhere(new_label == ii).reshape(-1)]
            has_old = members[np.argwhere(all_pre_ins[members] != -1)]
            not_old = members[np.argwhere(all_pre_ins[members] == -1)]
            if len(has_old) == 0:
                all_pre_ins[not_old] = max_instance + 1
                max_instance = max_instance + 1
            elif len(not_old) == 0:
                continue
            else:
                un = np.unique(all_pre_ins[has_old])
                best_iou, best_label = 0, 0
                for g in un:
                    idx_old_all = originids[np.argwhere(all_pre_ins[originids] == g).reshape(-1)]
                    union = np.union1d(idx_old_all, members)
                    inter = np.intersect1d(idx_old_all, members)
                    iou = float(inter.size) / float(union.size)
                    if iou > best_iou:
                        best_iou, best_label = iou, g
                if best_iou > 0.1:
                    all_pre_ins[not_old] = best_label
                else:
                    all_pre_ins[not_old] = max_instance + 1
                    max_instance = max_instance + 1
    return all_pre_ins, max_instance


def back_project(pos, ins_pre, sem_pred, stuff_classes, max_dist=1.0, min_size=10):
    """:592-633."""
    ins_pre = np.asarray(ins_pre)
    has = ins_pre != -1
    if not has.any():
        return np.full_like(ins_pre, -1)
    x_idx, d2 = nearest(pos[has], pos)
    full = ins_pre[has][x_idx].copy()
    for l in stuff_classes:
        full[np.asarray(sem_pred) == l] = -1
    full[np.sqrt(d2) > max_dist] = -1
    for l in np.unique(full):
        if l == -1:
            continue
        m = full == l
        if m.sum() < min_size:
            full[m] = -1
    return full
