"""Panoptic quality as the reference's final evaluation defines it
(torch_points3d/datasets/panoptic/npm3d.py:241-348 `final_eval`): per thing class a predicted instance is a TP
when its IoU with its best ground-truth instance is >= 0.5; RQ = 2PR/(P+R), SQ = mean IoU over TPs, PQ = SQ*RQ;
stuff classes count as one segment each (IoU >= 0.5); the mean runs over classes present in the ground truth.
numpy, host side: it is the measurement, not the path."""
import numpy as np


def panoptic_quality(pred_sem, clusters, gt_sem, gt_inst, num_classes, thing_classes):
    pred_sem, gt_sem, gt_inst = np.asarray(pred_sem), np.asarray(gt_sem), np.asarray(gt_inst)
    n = len(gt_sem)
    pred_inst = np.zeros(n, np.int64)
    for i, c in enumerate(clusters):
        pred_inst[np.asarray(c)] = i + 1
    pq, sq, rq = [], [], []
    for cls in range(num_classes):
        gmask = gt_sem == cls
        if not gmask.any():
            continue
        if cls in thing_classes:
            gids = np.unique(gt_inst[gmask & (gt_inst > 0)])
            # predicted instances of this class: majority semantic prediction of the cluster
            pids = [i + 1 for i, c in enumerate(clusters) if len(c) and np.bincount(pred_sem[np.asarray(c)]).argmax() == cls]
            tp, iou_sum = 0, 0.0
            matched = set()
            for p in pids:
                pm = pred_inst == p
                cand, cnt = np.unique(gt_inst[pm & gmask], return_counts=True)
                best, best_iou = None, 0.0
                for g, inter in zip(cand, cnt):
                    if g == 0:
                        continue
                    union = pm.sum() + ((gt_inst == g) & gmask).sum() - inter
                    iou = inter / union
                    if iou > best_iou:
                        best, best_iou = g, iou
                if best is not None and best_iou >= 0.5 and best not in matched:
                    matched.add(best)
                    tp += 1
                    iou_sum += best_iou
            fp, fn = len(pids) - tp, len(gids) - tp
        else:
            pm = pred_sem == cls
            inter = (pm & gmask).sum()
            union = (pm | gmask).sum()
            iou = inter / union if union else 0.0
            tp, iou_sum = (1, iou) if iou >= 0.5 else (0, 0.0)
            fp = fn = 1 - tp
        s = iou_sum / tp if tp else 0.0
        r = 2 * tp / (2 * tp + fp + fn) if (tp + fp + fn) else 0.0
        sq.append(s), rq.append(r), pq.append(s * r)
    return {"PQ": float(np.mean(pq)), "SQ": float(np.mean(sq)), "RQ": float(np.mean(rq))}
