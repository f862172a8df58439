"""Panoptic losses of the reference, on the device, without torch_scatter / numba / `.cuda()` literals.

Mirror of torch_points3d/core/losses/panoptic_losses.py:
  offset_loss            :7-23      L1 norm + negative cosine, normalised by the number of instance points
  instance_iou_loss      :92-114    BCE against the soft IoU target clip((iou - lo) / (hi - lo), 0, 1)
  discriminative_loss    :203-343   per scene: pull (L1, delta_v=0.5) + push (L1, 2*delta_d=3.0) + 0.001 * reg
These are the consumers at the edge of the hot path (autograd through the backbone starts here); they are
plain tensor code, not kernels (SURVEY 2.1 row `core/losses/panoptic_losses.py`).
"""
import torch


def offset_loss(pred_offsets, gt_offsets, total_instance_points):
    pt_dist = torch.sum(torch.abs(pred_offsets - gt_offsets), dim=-1)
    offset_norm_loss = torch.sum(pt_dist) / (total_instance_points + 1e-6)
    gt_ = gt_offsets / (torch.norm(gt_offsets, p=2, dim=1).unsqueeze(-1) + 1e-8)
    pr_ = pred_offsets / (torch.norm(pred_offsets, p=2, dim=1).unsqueeze(-1) + 1e-8)
    offset_dir_loss = torch.sum(-(gt_ * pr_).sum(-1)) / (total_instance_points + 1e-6)
    return {"offset_norm_loss": offset_norm_loss, "offset_dir_loss": offset_dir_loss}


def instance_iou_loss(ious, predicted_clusters, cluster_scores, instance_labels=None, batch=None,
                      min_iou_threshold=0.25, max_iou_threshold=0.75):
    assert len(predicted_clusters) == cluster_scores.shape[0]
    best = ious.max(1)[0]
    shat = ((best - min_iou_threshold) / (max_iou_threshold - min_iou_threshold)).clamp(0.0, 1.0)
    shat = torch.where(best < min_iou_threshold, torch.zeros_like(shat), shat)
    shat = torch.where(best > max_iou_threshold, torch.ones_like(shat), shat)
    return torch.nn.functional.binary_cross_entropy(cluster_scores, shat.detach())


def discriminative_loss_single(prediction, correct_label, feature_dim, delta_v=0.5, delta_d=1.5, param_var=1.0,
                               param_dist=1.0, param_reg=0.001):
    pred = prediction.reshape(-1, feature_dim)
    zero = pred.new_zeros(())
    uniq, uid, counts = torch.unique(correct_label, return_inverse=True, return_counts=True)
    k = uniq.numel()
    if k == 0:
        return zero, zero, zero, zero
    seg = pred.new_zeros((k, feature_dim)).index_add_(0, uid, pred)
    mu = seg / (counts.reshape(-1, 1) + 1e-8)
    dist = torch.norm(pred - mu[uid], p=1, dim=1)
    dist = torch.square(torch.clamp(dist - delta_v, min=0.0))
    l_var = pred.new_zeros(k).index_add_(0, uid, dist) / (counts + 1e-8)
    l_var = l_var.sum() / float(k)
    if k > 1:
        diff = mu.unsqueeze(1) - mu.unsqueeze(0)                      # [k, k, D]
        off = ~torch.eye(k, dtype=torch.bool, device=pred.device)
        mu_norm = torch.norm(diff[off], p=1, dim=1)
        l_dist = torch.square(torch.clamp(2.0 * delta_d - mu_norm, min=0.0)).mean()
    else:
        l_dist = zero
    l_reg = torch.norm(mu, p=1, dim=1).mean()
    l_var, l_dist, l_reg = param_var * l_var, param_dist * l_dist, param_reg * l_reg
    return l_var + l_dist + l_reg, l_var, l_dist, l_reg


def discriminative_loss(embedding_logits, instance_labels, batch, feature_dim):
    parts = [[], [], [], []]
    for s in torch.unique(batch):
        m = batch == s
        for lst, v in zip(parts, discriminative_loss_single(embedding_logits[m], instance_labels[m], feature_dim)):
            lst.append(v)
    if not parts[0]:
        z = embedding_logits.new_zeros(())
        return {"ins_loss": z, "ins_var_loss": z, "ins_dist_loss": z, "ins_reg_loss": z}
    names = ("ins_loss", "ins_var_loss", "ins_dist_loss", "ins_reg_loss")
    return {n: torch.stack(p).mean() for n, p in zip(names, parts)}
