"""ORACLE (test infrastructure, not product code): the reference's CPU algorithm for the backbone, restated.

Sparse convolution in MinkowskiEngine's own CPU formulation -- per kernel offset: gather rows, GEMM,
scatter-add (SURVEY 2.4 / 8d "CPU baseline") -- with torch CPU tensors so that autograd provides the backward
pass, wired into the ResUNet exactly as the reference wires it:

  torch_points3d/modules/MinkowskiEngine/api_modules.py:26-82   ResBlock (conv3 BN ReLU conv3 BN ReLU + shortcut)
  torch_points3d/modules/MinkowskiEngine/api_modules.py:251-285 ResNetDown (conv_in BN ReLU, N blocks)
  torch_points3d/modules/MinkowskiEngine/api_modules.py:288-311 ResNetUp (cat skip, transposed convs)
  torch_points3d/applications/minkowski.py:160-196              MinkowskiUnet.forward (skip stack)
  torch_points3d/models/panoptic/PointGroup3heads.py:69-81,103-108 heads

It is written functionally over a state_dict (no nn.Module shared with the product) so that the wiring is an
independent restatement.  PARITY UNPINNED against MinkowskiEngine itself (absent, see sparse_ref.py header).

Used by: tests/ (parity of the CUDA path), __graft_entry__.smoke(), bench.py cpu_baseline / --impl reference.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import sparse_ref as sr


class CpuMaps(sr.Maps):
    """sparse_ref.Maps + cached torch pair lists per kernel map."""

    def __init__(self, coords):
        super().__init__(coords)
        self._pairs = {}

    def pair_lists(self, nbr):
        key = id(nbr)
        if key not in self._pairs:
            i, o, offs = sr.pairs(nbr)
            self._pairs[key] = (torch.from_numpy(i.astype(np.int64)), torch.from_numpy(o.astype(np.int64)),
                                offs.tolist(), nbr)
        return self._pairs[key]


def sparse_conv(X, W, maps, ts, ksize, stride, transpose):
    """-> (Y, ts_out).  Per-offset gather -> mm -> index_add_ (autograd-able)."""
    nbr, mirror, _, _, ts_out = maps.conv_maps(ts, ksize, stride, transpose)
    if nbr is None:
        return X @ W.reshape(W.shape[-2], W.shape[-1]), ts_out
    in_idx, out_idx, offs, _ = maps.pair_lists(nbr)
    K = W.shape[0]
    Y = torch.zeros(nbr.shape[1], W.shape[2], dtype=X.dtype)
    for tk in range(K):
        a, b = offs[tk], offs[tk + 1]
        if b > a:
            k = K - 1 - tk if mirror else tk
            Y = Y.index_add(0, out_idx[a:b], X.index_select(0, in_idx[a:b]) @ W[k])
    return Y, ts_out


def _bn(x, sd, prefix, training, eps=1e-5):
    if training:
        return F.batch_norm(x, None, None, sd[prefix + ".bn.weight"], sd[prefix + ".bn.bias"], True, 0.0, eps)
    return F.batch_norm(x, sd[prefix + ".bn.running_mean"], sd[prefix + ".bn.running_var"],
                        sd[prefix + ".bn.weight"], sd[prefix + ".bn.bias"], False, 0.0, eps)


def _res_block(x, ts, sd, p, maps, transpose, training):
    y, _ = sparse_conv(x, sd[p + ".block.0.kernel"], maps, ts, 3, 1, transpose)
    y = torch.relu(_bn(y, sd, p + ".block.1", training))
    y, _ = sparse_conv(y, sd[p + ".block.3.kernel"], maps, ts, 3, 1, transpose)
    y = torch.relu(_bn(y, sd, p + ".block.4", training))
    if (p + ".downsample.0.kernel") in sd:
        s, _ = sparse_conv(x, sd[p + ".downsample.0.kernel"], maps, ts, 1, 1, transpose)
        s = _bn(s, sd, p + ".downsample.1", training)
    else:
        s = x
    return y + s


def _res_level(x, ts, sd, p, maps, ksize, stride, n_blocks, transpose, training):
    y, ts = sparse_conv(x, sd[p + ".conv_in.0.kernel"], maps, ts, ksize, stride, transpose)
    y = torch.relu(_bn(y, sd, p + ".conv_in.1", training))
    for j in range(n_blocks):
        y = _res_block(y, ts, sd, "%s.blocks.%d" % (p, j), maps, transpose, training)
    return y, ts


def _per_level(v, i):
    return v[i] if isinstance(v, (list, tuple)) else v


def unet_forward(sd, cfg, x, coords, training=True, prefix=""):
    """MinkowskiUnet.forward over a state_dict.  cfg = resolved compact config (ints).  -> [N, C_out]."""
    maps = CpuMaps(np.asarray(coords))
    down, up = cfg["down_conv"], cfg["up_conv"]
    nd, nu = len(down["down_conv_nn"]), len(up["up_conv_nn"])
    ts = 1
    stack = []
    for i in range(nd):
        x, ts = _res_level(x, ts, sd, "%sdown_modules.%d" % (prefix, i), maps, _per_level(down["kernel_size"], i),
                           _per_level(down["stride"], i), _per_level(down["N"], i), False, training)
        stack.append((x, ts) if i < nd - 1 else None)
    for i in range(nu):
        skip = stack.pop()
        if skip is not None:
            assert skip[1] == ts
            x = torch.cat([x, skip[0]], dim=1)
        x, ts = _res_level(x, ts, sd, "%sup_modules.%d" % (prefix, i), maps, _per_level(up["kernel_size"], i),
                           _per_level(up["stride"], i), _per_level(up["N"], i), True, training)
    assert ts == 1
    return x


def mlp_head(x, sd, p, training, final_bias=True):
    """Seq(MLP([C,C], bias=False), Linear) with FastBatchNorm1d + LeakyReLU(0.2)
    (core/common_modules/base_modules.py:35-45; PointGroup3heads.py:69-81)."""
    y = x @ sd[p + ".0.0.0.weight"].t()
    bnp = p + ".0.0.1.batch_norm"
    if training:
        y = F.batch_norm(y, None, None, sd[bnp + ".weight"], sd[bnp + ".bias"], True, 0.0, 1e-5)
    else:
        y = F.batch_norm(y, sd[bnp + ".running_mean"], sd[bnp + ".running_var"], sd[bnp + ".weight"],
                         sd[bnp + ".bias"], False, 0.0, 1e-5)
    y = F.leaky_relu(y, 0.2)
    return y @ sd[p + ".1.weight"].t() + sd[p + ".1.bias"]


def resolve_cfg(cfg, input_nc):
    """Evaluate the "2*in_feat" style strings (model_definition_resolver.py:29-58)."""
    consts = {"FEAT": input_nc}
    for k, v in (cfg.get("define_constants") or {}).items():
        consts[k] = eval(v, {}, consts) if isinstance(v, str) else v

    def ev(o):
        if isinstance(o, str):
            try:
                return eval(o, {"__builtins__": {}}, consts)
            except Exception:
                return o
        if isinstance(o, (list, tuple)):
            return [ev(v) for v in o]
        if isinstance(o, dict):
            return {k: ev(v) for k, v in o.items()}
        return o

    return ev(cfg)


# ------------------------------------------------------------------------------------------------
# whole step on the CPU: heads, losses, clustering  (reference: models/panoptic/PointGroup3heads.py:101-157,
# 552-639; core/losses/panoptic_losses.py:7-23,203-343) -- used as the parity oracle of the model-level tests
# and as the CPU arm that bench.py times next to the GPU path.
# ------------------------------------------------------------------------------------------------
def heads_forward(sd, feats, training=True, has_offset=True, has_embed=True):
    sem = torch.log_softmax(mlp_head(feats, sd, "Semantic", training), dim=-1)
    off = mlp_head(feats, sd, "Offset", training) if has_offset else None
    emb = mlp_head(feats, sd, "Embed", training) if has_embed else None
    return sem, off, emb


def offset_loss_ref(pred, gt, n_inst):
    """panoptic_losses.py:7-23."""
    norm = (pred - gt).abs().sum(-1).sum() / (n_inst + 1e-6)
    g = gt / (gt.norm(dim=1, keepdim=True) + 1e-8)
    p = pred / (pred.norm(dim=1, keepdim=True) + 1e-8)
    return norm, (-(g * p).sum(-1)).sum() / (n_inst + 1e-6)


def discriminative_loss_ref(emb, inst, batch, delta_v=0.5, delta_d=1.5, reg=0.001):
    """panoptic_losses.py:203-343, written per scene / per instance with explicit python loops."""
    totals = []
    for s in torch.unique(batch):
        e, l = emb[batch == s], inst[batch == s]
        ids = torch.unique(l)
        mus, l_var = [], 0.0
        for i in ids:
            pts = e[l == i]
            mu = pts.sum(0) / (pts.shape[0] + 1e-8)
            mus.append(mu)
            l_var = l_var + (torch.clamp((pts - mu).abs().sum(1) - delta_v, min=0.0) ** 2).sum() / (pts.shape[0] + 1e-8)
        k = len(mus)
        l_var = l_var / k
        l_dist = 0.0
        if k > 1:
            acc = 0.0
            for a in range(k):
                for b in range(k):
                    if a != b:
                        acc = acc + torch.clamp(2 * delta_d - (mus[a] - mus[b]).abs().sum(), min=0.0) ** 2
            l_dist = acc / (k * (k - 1))
        l_reg = sum(m.abs().sum() for m in mus) / k
        totals.append(l_var + l_dist + reg * l_reg)
    return sum(totals) / len(totals)


def step_loss(sd, cfg, batch, weights, training=True, has_offset=True, has_embed=True):
    """Forward + the epoch <= prepare_epoch loss of PointGroup3heads._compute_loss on CPU tensors.
    `batch`: object with x, coords, batch, y, instance_labels, instance_mask, vote_label (CPU torch tensors)."""
    coords = np.concatenate([batch.batch.numpy()[:, None], batch.coords.numpy()], 1).astype(np.int32)
    feats = unet_forward(sd, cfg, batch.x, coords, training=training, prefix="Backbone.")
    sem, off, emb = heads_forward(sd, feats, training, has_offset, has_embed)
    loss = weights["semantic"] * F.nll_loss(sem, batch.y.long(), ignore_index=-1)
    im = batch.instance_mask
    if has_offset:
        a, b = offset_loss_ref(off[im], batch.vote_label[im], im.sum())
        loss = loss + weights["offset_norm_loss"] * a + weights["offset_dir_loss"] * b
    if has_embed:
        loss = loss + weights["embedding_loss"] * discriminative_loss_ref(emb[im], batch.instance_labels[im],
                                                                          batch.batch[im])
    return loss, sem, off, emb


# ------------------------------------------------------------------------------------------------
# ScoreNet + score loss  (reference: models/panoptic/PointGroup3heads.py:393-454 `_compute_score`, scorer_type
# "unet"; core/losses/panoptic_losses.py:25-37 `instance_ious` -> torch_points_kernels.instance_iou, :92-114
# `instance_iou_loss`; combined at PointGroup3heads.py:593-623).  Written with the reference's explicit per-proposal
# loops -- the product batches them.
# ------------------------------------------------------------------------------------------------
def score_forward(sd, scorer_cfg, backbone_features, coords_xyz, clusters, training=True):
    """-> (cluster_scores [n_prop], per-row scorer features).  Proposal i becomes batch id i of a second sparse tensor
    built from the proposal's voxel coordinates (PointGroup3heads.py:402-416), run through ScorerUnet, max-pooled per
    proposal (scatter max, :436-438) and scored by Linear + Sigmoid (ScorerHead, :51,440)."""
    xs, cs = [], []
    for i, c in enumerate(clusters):
        c = torch.as_tensor(c, dtype=torch.long)
        xs.append(backbone_features[c])
        cc = np.asarray(coords_xyz)[c.numpy()]
        cs.append(np.concatenate([np.full((len(c), 1), i, np.int32), cc.astype(np.int32)], 1))
    x = torch.cat(xs)
    coords = np.concatenate(cs, 0)
    out = unet_forward(sd, scorer_cfg, x, coords, training=training, prefix="ScorerUnet.")
    feats, start = [], 0
    for c in clusters:
        feats.append(out[start:start + len(c)].max(0)[0])
        start += len(c)
    feats = torch.stack(feats)
    scores = torch.sigmoid(feats @ sd["ScorerHead.0.weight"].t() + sd["ScorerHead.0.bias"]).squeeze(-1)
    return scores, out


def instance_iou_ref(clusters, instance_labels, batch):
    """torch_points_kernels.instance_iou as the reference uses it (panoptic_losses.py:37): IoU of every proposal with
    every ground-truth instance (labels 1..M_s per scene, 0 = none) of ITS scene, scenes concatenated; 0 elsewhere."""
    il, b = np.asarray(instance_labels), np.asarray(batch)
    nb = int(b.max()) + 1
    per_scene = [int(il[b == s].max()) if (b == s).any() else 0 for s in range(nb)]
    offs = np.concatenate([[0], np.cumsum(per_scene)])
    out = np.zeros((len(clusters), int(offs[-1])), np.float32)
    for p, c in enumerate(clusters):
        c = np.asarray(c)
        s = int(b[c[0]])
        for inst in range(1, per_scene[s] + 1):
            gt = (b == s) & (il == inst)
            inter = int(gt[c].sum())
            out[p, offs[s] + inst - 1] = inter / float(len(c) + int(gt.sum()) - inter)
    return out


def score_loss_ref(ious, scores, lo=0.25, hi=0.75):
    """panoptic_losses.py:92-114 with its three masks."""
    best = torch.as_tensor(ious).max(1)[0]
    shat = torch.zeros_like(best)
    mid = (best >= lo) & (best <= hi)
    shat[best > hi] = 1.0
    shat[mid] = (best[mid] - lo) / (hi - lo)
    return F.binary_cross_entropy(scores, shat)
