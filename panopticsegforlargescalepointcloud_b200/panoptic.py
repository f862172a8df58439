"""PointGroup-style panoptic models with the torch_points3d model API, on the B200 kernels.

Host-side mirror of
  torch_points3d/models/base_model.py:23-67,259-297        BaseModel (set_input / forward / backward /
                                                            optimize_parameters2 / get_output / get_current_losses)
  torch_points3d/models/panoptic/PointGroup3heads.py:21-639 PointGroup3heads (semantic + offset + embed heads)
  torch_points3d/models/panoptic/pointgroup.py:20-358      PointGroup       (semantic + offset heads)
  torch_points3d/models/panoptic/pointgroupembed.py:29-1019 PointGroupEmbed  (semantic + embed heads)
  torch_points3d/models/panoptic/structure_3heads.py:19-79 PanopticResults / PanopticLabels
  torch_points3d/core/common_modules/base_modules.py:35-45,128-164  MLP / FastBatchNorm1d / Seq

Module names follow the reference (Backbone, ScorerUnet, ScorerHead, Offset, Embed, Semantic and the nesting
inside Seq/MLP) so state_dict keys are interchangeable.  Constructor: (option, model_type, dataset, modules)
exactly as model_factory.py:27-44 calls it; `option` is any mapping with attribute access (AttrDict here,
OmegaConf in the reference).

Differences by design (DESIGN.md "host path"): the batch lives on the device for the whole step (the
reference keeps it on the CPU and bounces proposals through host memory, PointGroup3heads.py:97,416),
proposals are gathered with one index op instead of a python loop, and there is no multiprocessing.Pool.
cluster_type values with MeanShift on the embeddings (3-6 of PointGroup3heads, 7-8 of PointGroupEmbed: paper settings I,
IV, V) run on meanshift.py; the random feature-subset loops (`cluster_loop*`) are not reproduced.
"""
from collections import OrderedDict
from typing import List, NamedTuple

import torch
import torch.nn as nn

from . import hdbscan as _hdbscan
from . import meanshift as _meanshift
from . import losses as L
from . import tpk
from . import me as _me
from .backbone import Minkowski

IGNORE_LABEL = -1


class AttrDict(dict):
    """dict with attribute access and .get, recursively -- all the models use of OmegaConf nodes."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        for k, v in list(self.items()):
            if isinstance(v, dict) and not isinstance(v, AttrDict):
                self[k] = AttrDict(v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            return None  # OmegaConf.set_struct(cfg, False): missing keys read as None

    def __setattr__(self, k, v):
        self[k] = v


class PanopticLabels(NamedTuple):
    center_label: torch.Tensor
    y: torch.Tensor
    num_instances: torch.Tensor
    instance_labels: torch.Tensor
    instance_mask: torch.Tensor
    vote_label: torch.Tensor


class PanopticResults(NamedTuple):
    semantic_logits: torch.Tensor
    offset_logits: torch.Tensor
    embed_logits: torch.Tensor
    cluster_scores: torch.Tensor
    mask_scores: torch.Tensor
    clusters: List[torch.Tensor]
    cluster_type: torch.Tensor

    def get_instances(self, nms_threshold=0.3, min_cluster_points=100, min_score=0.5):
        """Proposal NMS (structure_3heads.py:28-71): cross IoU from sorted point -> proposal lists and the greedy loop in
        one device pass (tpk.proposal_nms -> pgs_prop_cross_nms), then the size / score filters; one read-back of the
        picked ids (the caller needs a Python list of index tensors)."""
        if not self.clusters:
            return [], []
        if self.cluster_scores is None:
            return None, self.clusters
        dev = self.semantic_logits.device
        keep, order = tpk.proposal_nms(self.clusters, self.cluster_scores, nms_threshold)
        sizes = torch.tensor([c.shape[0] for c in self.clusters], device=dev)
        ok = keep & (sizes > min_cluster_points) & (self.cluster_scores.detach() > min_score)
        ids = order[ok[order]].tolist()                 # picked proposals in descending-score order, like the reference
        return ids, [self.clusters[i] for i in ids]


# --------------------------------------------------------------------------------------------
# common modules (core/common_modules/base_modules.py)
# --------------------------------------------------------------------------------------------
class FastBatchNorm1d(nn.Module):
    def __init__(self, num_features, momentum=0.1, **kwargs):
        super().__init__()
        self.batch_norm = nn.BatchNorm1d(num_features, momentum=momentum, **kwargs)

    def forward(self, x):
        return self.batch_norm(x)


def MLP(channels, activation=None, bn_momentum=0.1, bias=True):
    activation = activation if activation is not None else nn.LeakyReLU(0.2)
    return nn.Sequential(*[
        nn.Sequential(nn.Linear(channels[i - 1], channels[i], bias=bias),
                      FastBatchNorm1d(channels[i], momentum=bn_momentum), activation)
        for i in range(1, len(channels))
    ])


class Seq(nn.Sequential):
    def __init__(self):
        super().__init__()
        self._num_modules = 0

    def append(self, module):
        self.add_module(str(self._num_modules), module)
        self._num_modules += 1
        return self


class _Batch:
    """Device-resident view of the keys the hot path reads from a collated batch."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def __getitem__(self, k):
        return getattr(self, k)

    def to(self, device):
        return _Batch(**{k: (v.to(device) if torch.is_tensor(v) else v) for k, v in self.__dict__.items()})


def _nll_loss_mean(logp, target, ignore_index):
    """F.nll_loss(logp, target, ignore_index=..., reduction="mean") (PointGroup3heads.py:554-557) as gather + masked mean:
    torch's nll_loss reduces [N, C] with a single thread block (0.2 ms forward + 0.1 ms backward per 200 k rows)."""
    valid = target != ignore_index
    picked = logp.gather(1, target.clamp_min(0).unsqueeze(1)).squeeze(1)
    return -(picked * valid).sum() / valid.sum()


def _get(data, key):
    v = data[key] if not hasattr(data, key) else getattr(data, key)
    return torch.as_tensor(v) if not torch.is_tensor(v) else v


# --------------------------------------------------------------------------------------------
# BaseModel (the slice of models/base_model.py the trainer calls)
# --------------------------------------------------------------------------------------------
class BaseModel(nn.Module):
    __REQUIRED_DATA__: List[str] = []
    __REQUIRED_LABELS__: List[str] = []

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.loss_names = []
        self.output = None
        self._conv_type = "SPARSE"
        self._optimizer = None
        self._lr_scheduler = None
        self._grad_clip = -1
        self._num_epochs = self._num_batches = self._num_samples = 0
        self._grad_hook = None   # set by parallel.DataParallelStep: called between backward and the optimizer

    @property
    def conv_type(self):
        return self._conv_type

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def optimizer(self):
        return self._optimizer

    def get_input(self):
        return getattr(self, "input", None)

    def get_output(self):
        return self.output

    def get_labels(self):
        return getattr(self, "labels", None)

    def get_current_losses(self):
        out = OrderedDict()
        for name in self.loss_names:
            if hasattr(self, name):
                try:
                    out[name] = float(getattr(self, name))
                except Exception:
                    out[name] = None
        return out

    def instantiate_optimizers(self, config=None, cuda_enabled=False):
        """Adam lr 1e-3 + ExponentialLR, the recipe of conf/training/*.yaml + conf/lr_scheduler/exponential.yaml."""
        tr = (config or {}).get("training", config or {}) if hasattr(config or {}, "get") else {}
        opt_cfg = (tr.get("optim", {}) or {}).get("optimizer", {}) if hasattr(tr, "get") else {}
        params = dict((opt_cfg.get("params", {}) or {})) if hasattr(opt_cfg, "get") else {}
        params.setdefault("lr", 1e-3)
        cls = getattr(torch.optim, (opt_cfg.get("class", "Adam") if hasattr(opt_cfg, "get") else "Adam"))
        plist = list(self.parameters())
        if (cls in (torch.optim.Adam, torch.optim.AdamW) and "fused" not in params and "foreach" not in params
                and plist and all(p.is_cuda and p.dtype == torch.float32 for p in plist)):
            params["fused"] = True   # same update rule in 2-3 multi-tensor kernels instead of ~30 (0.6 ms + 1 ms of host time)
        self._optimizer = cls(plist, **params)
        self._lr_scheduler = torch.optim.lr_scheduler.ExponentialLR(self._optimizer, gamma=0.9885)
        return self._optimizer

    def optimize_parameters2(self, epoch, step, batch_size):
        """base_model.py:259-285: forward, zero_grad, backward, (clip), optimizer step."""
        self.forward(epoch=epoch, step=step, is_training=True)
        if getattr(self, "_zero_grad_hook", None) is not None:
            self._zero_grad_hook()
        else:
            self._optimizer.zero_grad(set_to_none=False)
        self.backward(epoch)
        _me.join_side_stream()   # weight-gradient kernels run on a side stream (me.DW_DIRECT)
        if self._grad_hook is not None:
            self._grad_hook()
        if self._grad_clip > 0:
            torch.nn.utils.clip_grad_value_(self.parameters(), self._grad_clip)
        self._optimizer.step()
        if self._lr_scheduler is not None and epoch != self._num_epochs:
            self._lr_scheduler.step()
        self._num_epochs = epoch
        self._num_batches += 1
        self._num_samples += batch_size

    def optimize_parameters(self, epoch, batch_size):
        """base_model.py: the older two-argument entry point (no step index)."""
        return self.optimize_parameters2(epoch, self._num_batches, batch_size)


# --------------------------------------------------------------------------------------------
# panoptic models
# --------------------------------------------------------------------------------------------
class _PanopticBase(BaseModel):
    __REQUIRED_DATA__ = ["pos"]
    __REQUIRED_LABELS__ = list(PanopticLabels._fields)
    HAS_OFFSET = True
    HAS_EMBED = True

    def __init__(self, option, model_type, dataset, modules=None):
        super().__init__(option)
        backbone_options = option.get("backbone", {"architecture": "unet"})
        self.Backbone = Minkowski(backbone_options.get("architecture", "unet"), input_nc=dataset.feature_dimension,
                                  num_layers=4, config=backbone_options.get("config", {}))
        nc = self.Backbone.output_nc
        self._scorer_type = option.get("scorer_type", None)
        self.use_score_net = option.get("use_score_net", True)
        if option.get("scorer_unet", None) is not None:
            self.ScorerUnet = Minkowski("unet", input_nc=nc, num_layers=4, config=option.scorer_unet)
            self.ScorerHead = Seq().append(nn.Linear(self.ScorerUnet.output_nc, 1)).append(nn.Sigmoid())
        else:
            self.ScorerUnet = None
            self.ScorerHead = None
        if self.HAS_OFFSET:
            self.Offset = Seq().append(MLP([nc, nc], bias=False))
            self.Offset.append(nn.Linear(nc, 3))
        if self.HAS_EMBED:
            self.Embed = Seq().append(MLP([nc, nc], bias=False))
            self.Embed.append(nn.Linear(nc, option.get("embed_dim", 5)))
        self.Semantic = (Seq().append(MLP([nc, nc], bias=False)).append(nn.Linear(nc, dataset.num_classes))
                         .append(nn.LogSoftmax(dim=-1)))
        self.loss_names = ["loss", "offset_norm_loss", "offset_dir_loss", "ins_loss", "ins_var_loss", "ins_dist_loss",
                           "ins_reg_loss", "semantic_loss", "score_loss", "mask_loss"]
        stuff = torch.as_tensor(list(dataset.stuff_classes)).long()
        self._stuff_classes = torch.cat([torch.tensor([IGNORE_LABEL]), stuff])

    def get_opt_mergeTh(self):
        return self.opt.block_merge_th if self.opt.block_merge_th else 0.01

    def set_input(self, data, device):
        """Host -> device for everything the step reads (reference: PointGroup3heads.py:95-99 keeps `data` on the
        CPU; here one device-resident batch feeds backbone, clustering, scoring and losses)."""
        keys = ["pos", "coords", "x", "batch"] + self.__REQUIRED_LABELS__
        self.input = _Batch(**{k: _get(data, k).to(device, non_blocking=True) for k in keys})
        pm = getattr(data, "coordinate_manager", None)     # handle from prefetch_maps (backbone.PrebuiltMaps), optional
        if pm is not None:
            self.input.coordinate_manager = pm
        self.raw_pos = self.input.pos
        self.labels = PanopticLabels(**{l: self.input[l] for l in self.__REQUIRED_LABELS__})

    def prefetch_maps(self, data, stream=None, wait_event=None):
        """Coordinate maps of the NEXT batch built ahead of time (backbone.BaseMinkowski.prefetch_maps): call it with the
        device-resident batch after the current step has been launched, attach the result to that batch as
        `coordinate_manager` before `set_input`.  The forward pass then has no host read-back left."""
        return self.Backbone.prefetch_maps(_get(data, "batch"), _get(data, "coords"), stream=stream, wait_event=wait_event)

    # ---- clustering recipes ----
    def _region_grow(self, pos, predicted_labels, nsample=None):
        kw = dict(ignore_labels=self._stuff_classes.to(self.device), radius=self.opt.cluster_radius_search,
                  min_cluster_size=10)
        if nsample is not None:
            kw["nsample"] = nsample   # raw-position call sites omit it => tpk default 300 (PointGroup3heads.py:185-192)
        return tpk.region_grow(pos, predicted_labels, self.input.batch, **kw)

    def _thing_mask(self, predicted_labels):
        return ~torch.isin(predicted_labels, self._stuff_classes.to(self.device))

    def _hdbscan(self, feats, predicted_labels, ctype):
        mask = self._thing_mask(predicted_labels)
        local_ind = torch.nonzero(mask).squeeze(1)
        label_batch = self.input.batch[mask]
        return _hdbscan.cluster_single(feats[mask].detach(), torch.unique(label_batch), label_batch, local_ind, ctype)

    def _meanshift(self, feats, predicted_labels, ctype):
        """Embedding branch of the reference's _cluster3.._cluster6 / pointgroupembed._cluster7/8: thing points only,
        per scene, sklearn-equivalent mean shift with opt.bandwidth (meanshift_cluster.cluster_single)."""
        mask = self._thing_mask(predicted_labels)
        local_ind = torch.nonzero(mask).squeeze(1)
        label_batch = self.input.batch[mask]
        bw = self.opt.bandwidth if self.opt.bandwidth is not None else 0.6
        return _meanshift.cluster_single(feats[mask].detach(), torch.unique(label_batch), label_batch, local_ind, ctype, bw)

    def _cluster_ms(self, semantic_logits, offset_logits, embed_logits, raw, votes, ms_type):
        """region_grow on raw positions (nsample default) and / or shifted positions (nsample 200), plus mean shift on
        the embeddings: PointGroup3heads._cluster3 (-,-), _cluster4 (raw), _cluster5 (votes), _cluster6 (raw + votes);
        pointgroupembed._cluster7 (-,-), _cluster8 (raw)."""
        pred = torch.max(semantic_logits, 1)[1]
        clusters, types = [], []
        if raw:
            c = self._region_grow(self.raw_pos, pred)
            clusters += c
            types += [0] * len(c)
        if votes:
            c = self._region_grow(self.raw_pos + offset_logits.detach(), pred, nsample=200)
            types += [1 if raw else 0] * len(c)
            clusters += c
        c, t = self._meanshift(embed_logits, pred, ms_type)
        return clusters + c, torch.tensor(types + t, dtype=torch.uint8, device=self.device)

    def _cluster_votes(self, semantic_logits, offset_logits):                      # PointGroup3heads._cluster
        pred = torch.max(semantic_logits, 1)[1]
        clusters = self._region_grow(self.raw_pos + offset_logits.detach(), pred, nsample=200)
        return clusters, torch.zeros(len(clusters), dtype=torch.uint8, device=self.device)

    def _cluster_pos_and_votes(self, semantic_logits, offset_logits):              # PointGroup3heads._cluster2
        pred = torch.max(semantic_logits, 1)[1]
        c_pos = self._region_grow(self.raw_pos, pred)
        c_vote = self._region_grow(self.raw_pos + offset_logits.detach(), pred, nsample=200)
        ctype = torch.zeros(len(c_pos) + len(c_vote), dtype=torch.uint8, device=self.device)
        if len(c_pos):   # upstream marks the vote clusters only when raw-position growing found something
            ctype[len(c_pos):] = 1
        return c_pos + c_vote, ctype

    def _cluster_hdbscan_embed(self, semantic_logits, embed_logits):               # pointgroupembed._cluster14
        pred = torch.max(semantic_logits, 1)[1]
        clusters, _ = self._hdbscan(embed_logits, pred, 0)
        return clusters, torch.zeros(len(clusters), dtype=torch.uint8, device=self.device)

    def _cluster_hdbscan_xyz_embed(self, semantic_logits, embed_logits):           # pointgroupembed._cluster
        pred = torch.max(semantic_logits, 1)[1]
        c0, t0 = self._hdbscan(self.raw_pos, pred, 0)
        c1, t1 = self._hdbscan(embed_logits, pred, 1)
        return c0 + c1, torch.tensor(t0 + t1, dtype=torch.uint8, device=self.device)

    def _cluster_hdbscan_xyz_shifted(self, semantic_logits, offset_logits):        # pointgroup._cluster3
        pred = torch.max(semantic_logits, 1)[1]
        c0, t0 = self._hdbscan(self.raw_pos, pred, 0)
        c1, t1 = self._hdbscan(self.raw_pos + offset_logits.detach(), pred, 1)
        return c0 + c1, torch.tensor(t0 + t1, dtype=torch.uint8, device=self.device)

    def _do_cluster(self, semantic_logits, offset_logits, embed_logits):
        raise NotImplementedError

    # ---- scoring ----
    def _compute_score(self, epoch, all_clusters, backbone_features, semantic_logits):
        """PointGroup3heads.py:393-454, batched: proposal i becomes batch id i of one sparse tensor."""
        dev = self.device
        sizes = torch.tensor([c.shape[0] for c in all_clusters], device=dev)
        flat = torch.cat(all_clusters)
        pid = torch.repeat_interleave(torch.arange(len(all_clusters), device=dev), sizes)
        if self._scorer_type:
            if self._scorer_type in ("MLP", "encoder"):
                raise NotImplementedError("scorer_type %r: the shipped configs use the U-Net scorer" % self._scorer_type)
            bc = _Batch(x=backbone_features[flat], coords=self.input.coords[flat], batch=pid, pos=None)
            out = self.ScorerUnet(bc).x
            feats = torch.full((len(all_clusters), out.shape[1]), float("-inf"), device=dev, dtype=out.dtype)
            feats = feats.scatter_reduce(0, pid.unsqueeze(1).expand(-1, out.shape[1]), out, reduce="amax",
                                         include_self=True)
            return self.ScorerHead(feats).squeeze(-1), None
        with torch.no_grad():
            sem = torch.zeros((len(all_clusters), semantic_logits.shape[1]), device=dev).index_add_(
                0, pid, semantic_logits[flat]) / sizes.unsqueeze(1)
            return torch.max(torch.exp(sem), 1)[0], None

    def forward(self, epoch=-1, **kwargs):
        backbone_features = self.Backbone(self.input).x
        semantic_logits = self.Semantic(backbone_features)
        offset_logits = self.Offset(backbone_features) if self.HAS_OFFSET else None
        embed_logits = self.Embed(backbone_features) if self.HAS_EMBED else None
        cluster_scores = mask_scores = all_clusters = cluster_type = None
        if self.use_score_net:
            if epoch > self.opt.prepare_epoch:
                all_clusters, cluster_type = self._do_cluster(semantic_logits, offset_logits, embed_logits)
                if len(all_clusters):
                    cluster_scores, mask_scores = self._compute_score(epoch, all_clusters, backbone_features,
                                                                      semantic_logits)
        else:
            with torch.no_grad():
                all_clusters, cluster_type = self._do_cluster(semantic_logits, offset_logits, embed_logits)
        self.output = PanopticResults(semantic_logits=semantic_logits, offset_logits=offset_logits,
                                      embed_logits=embed_logits, clusters=all_clusters, cluster_scores=cluster_scores,
                                      mask_scores=mask_scores, cluster_type=cluster_type)
        return self.output

    def _compute_loss(self, epoch):
        """PointGroup3heads.py:552-634 (mask loss omitted: mask_supervise is False in every shipped config)."""
        w = self.opt.loss_weights
        inp, out = self.input, self.output
        self.semantic_loss = _nll_loss_mean(out.semantic_logits, inp.y.to(torch.int64), IGNORE_LABEL)
        self.loss = w["semantic"] * self.semantic_loss
        im = inp.instance_mask
        if self.HAS_OFFSET:
            for name, v in L.offset_loss(out.offset_logits[im], inp.vote_label[im], torch.sum(im)).items():
                setattr(self, name, v)
                self.loss = self.loss + w[name] * v
        if self.HAS_EMBED:
            for name, v in L.discriminative_loss(out.embed_logits[im], inp.instance_labels[im], inp.batch[im],
                                                 self.opt.get("embed_dim", 5)).items():
                setattr(self, name, v)
                if name == "ins_loss":
                    self.loss = self.loss + w["embedding_loss"] * v
        if out.cluster_scores is not None and self._scorer_type and epoch > self.opt.prepare_epoch and self.use_score_net:
            ious = tpk.instance_iou(out.clusters, inp.instance_labels, inp.batch)
            self.score_loss = L.instance_iou_loss(ious, out.clusters, out.cluster_scores, inp.instance_labels, inp.batch,
                                                  min_iou_threshold=self.opt.min_iou_threshold,
                                                  max_iou_threshold=self.opt.max_iou_threshold)
            self.loss = self.loss + self.score_loss * w["score_loss"]

    def backward(self, epoch=-1):
        self._compute_loss(epoch)
        self.loss.backward()


class PointGroup3heads(_PanopticBase):
    """models/panoptic/PointGroup3heads.py (paper settings IV / V)."""

    def _do_cluster(self, sem, off, emb):
        ct = self.opt.cluster_type
        if ct == 1:
            return self._cluster_votes(sem, off)
        if ct == 2:
            return self._cluster_pos_and_votes(sem, off)
        if ct == 3:
            return self._cluster_ms(sem, off, emb, False, False, 0)   # PointGroup3heads._cluster3
        if ct == 4:
            return self._cluster_ms(sem, off, emb, True, False, 1)    # _cluster4
        if ct == 5:
            return self._cluster_ms(sem, off, emb, False, True, 1)    # _cluster5 (paper setting IV)
        if ct == 6:
            return self._cluster_ms(sem, off, emb, True, True, 2)     # _cluster6 (paper setting V)
        if ct == 14:
            return self._cluster_hdbscan_embed(sem, emb)
        raise NotImplementedError("cluster_type %r" % ct)


class PointGroup(_PanopticBase):
    """models/panoptic/pointgroup.py (paper settings II / III): semantic + offset heads."""
    HAS_EMBED = False

    def _do_cluster(self, sem, off, emb):
        ct = self.opt.cluster_type
        if ct == 1:
            return self._cluster_votes(sem, off)
        if ct == 2:
            return self._cluster_pos_and_votes(sem, off)
        if ct == 3:
            return self._cluster_hdbscan_xyz_shifted(sem, off)
        raise NotImplementedError("cluster_type %r" % ct)


class PointGroupEmbed(_PanopticBase):
    """models/panoptic/pointgroupembed.py: semantic + embedding heads, HDBSCAN recipes 1 and 14."""
    HAS_OFFSET = False

    def _do_cluster(self, sem, off, emb):
        ct = self.opt.cluster_type
        if ct == 1:
            return self._cluster_hdbscan_xyz_embed(sem, emb)
        if ct == 7:
            return self._cluster_ms(sem, off, emb, False, False, 0)   # pointgroupembed._cluster7 (paper setting I)
        if ct == 8:
            return self._cluster_ms(sem, off, emb, True, False, 1)    # pointgroupembed._cluster8
        if ct == 14:
            return self._cluster_hdbscan_embed(sem, emb)
        raise NotImplementedError("cluster_type %r (random feature-subset loops, SURVEY App. E)" % ct)


def paper_options(kind="urban", cluster_type=1, grid=0.12, use_score_net=True, prepare_epoch=30, scorer=True,
                  backbone="paper"):
    """conf/models/panoptic/area4_ablation_3heads_5.yaml:63-174 as a resolved AttrDict."""
    from . import backbone as bb
    cfg = {"paper": bb.paper_backbone_config, "two_level": bb.two_level_config}[backbone](16)
    return AttrDict(
        backbone=AttrDict(architecture="unet", config=cfg),
        scorer_unet=bb.scorer_unet_config(16) if scorer else None,
        scorer_type="unet" if scorer else None,
        use_score_net=use_score_net, prepare_epoch=prepare_epoch, cluster_type=cluster_type, embed_dim=5,
        cluster_radius_search=1.5 * grid, min_iou_threshold=0.25, max_iou_threshold=0.75, bandwidth=0.6,
        loss_weights=AttrDict(semantic=1, offset_norm_loss=0.1, offset_dir_loss=0.1, embedding_loss=1, score_loss=1,
                              mask_loss=1),
    )


class DatasetProperties:
    """The three attributes the model constructors read from a dataset (models/model_factory.py:8-45)."""

    def __init__(self, kind="urban", feature_dimension=4):
        from . import scenes
        self.feature_dimension = feature_dimension
        self.num_classes = scenes.num_classes(kind)
        self.stuff_classes = list(scenes.stuff_classes(kind))
