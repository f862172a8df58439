"""torch_points_kernels-shaped ops backed by the sm_100a kernels in libpgs_b200.so (csrc/cluster.cu).

Host-side mirror of the part of torch-points-kernels 0.7.0 the reference's hot path calls:

  region_grow(pos, labels, batch, ignore_labels, radius, nsample, min_cluster_size) -> List[LongTensor]
      torch_points3d/models/panoptic/PointGroup3heads.py:166-174,185-202,250-257,296-304,340-357
      torch_points3d/models/panoptic/pointgroup.py:141-149,160-177
  ball_query(radius, nsample, x, y, mode="PARTIAL_DENSE", batch_x=, batch_y=) -> (idx, dist2)
      torch_points3d/core/spatial_ops/neighbour_finder.py:35-37,164

Install as a drop-in with `sys.modules["torch_points_kernels"] = panopticsegforlargescalepointcloud_b200.tpk`.
Semantics frozen in DESIGN.md: neighbour lists are the first `nsample` same-scene points in ascending index
(the upstream CUDA kernel's scan order); clusters are the sequential seeded BFS partition, computed as the
min-ancestor labelling, emitted class-ascending then seed-ascending with members ascending.
There is no CPU path.
"""
from typing import List

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr

_MAX_GID = 32767


class _Grid:
    """Device voxel-hash grid over the support points of one call."""

    def __init__(self, pos, gid, radius):
        lib = _lib.load()
        n = pos.shape[0]
        dev = pos.device
        self.n, self.dev = n, dev
        self.pos, self.gid = pos, gid
        self.cap = lib.pgs_cmap_capacity(n)
        self.tkeys = torch.empty(self.cap, dtype=torch.int64, device=dev)
        self.tvals = torch.empty(self.cap, dtype=torch.int32, device=dev)
        self.spos = torch.empty((max(n, 1), 4), dtype=torch.float32, device=dev)
        self.skeys = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
        self.cell_start = torch.empty(n + 1, dtype=torch.int32, device=dev)
        self.meta = torch.zeros(3, dtype=torch.int32, device=dev)  # n_support, n_cells, status
        nb = lib.pgs_bq_grid_scratch_bytes(n)
        scratch = torch.empty(max(nb, 1), dtype=torch.uint8, device=dev)
        cell = float(radius) * 1.0001
        for _ in range(6):
            self.meta.zero_()
            check(lib.pgs_bq_grid_build(ptr(pos), ptr(gid), n, cell, ptr(self.tkeys), ptr(self.tvals), self.cap,
                                        ptr(self.spos), ptr(self.skeys), ptr(self.cell_start), ptr(self.meta[0:2]),
                                        ptr(self.meta[2:3]), ptr(scratch), nb, stream_ptr()))
            self.cell = cell
            # the status word is read lazily by the caller together with its own first sync
            self._status_checked = False
            if not self._range_error():
                break
            cell *= 8.0  # coordinates further than 32766 cells from the origin: coarser cells stay exact
        else:
            raise ValueError("positions are not finite or too far from the origin for the neighbour grid")

    def _range_error(self):
        return bool(int(self.meta[2]) & 1)


def _as_f32_pos(pos):
    if pos.dim() != 2 or pos.shape[1] != 3:
        raise ValueError("positions must be [N, 3]")
    if not pos.is_cuda:
        raise _lib.PgsError("positions must be on a CUDA device (no CPU path)")
    return pos.detach().to(torch.float32).contiguous()


def _query(grid, qpos, qkeys, n_q, n_rows, radius, nsample, want_dist):
    lib = _lib.load()
    dev = grid.dev
    nbr = torch.empty((max(n_rows, 1), nsample), dtype=torch.int32, device=dev)
    cnt = torch.zeros(max(n_rows, 1), dtype=torch.int32, device=dev)
    dist = torch.empty((max(n_rows, 1), nsample), dtype=torch.float32, device=dev) if want_dist else None
    check(lib.pgs_bq_query(ptr(grid.spos), ptr(qpos), ptr(qkeys), n_q, ptr(grid.tkeys), ptr(grid.tvals), grid.cap,
                           ptr(grid.cell_start), float(radius), int(nsample), ptr(nbr), ptr(cnt), ptr(dist),
                           stream_ptr()))
    return nbr, cnt, dist


def _gid_from_batch(batch, n, dev):
    if batch is None:
        return torch.zeros(n, dtype=torch.int32, device=dev)
    if batch.shape[0] != n:
        raise ValueError("batch must have one entry per point")
    return batch.to(device=dev, dtype=torch.int32).contiguous()


def ball_query(radius, nsample, x, y, mode="dense", batch_x=None, batch_y=None, sort=False):
    """PARTIAL_DENSE radius search: x support [N,3], y queries [M,3] -> (idx int64 [M,nsample], dist2 f32).
    Rows hold the first `nsample` hits in ascending support index, padded with -1 (idx and dist2)."""
    if str(mode).lower() != "partial_dense":
        raise NotImplementedError("only mode='PARTIAL_DENSE' is on the reference hot path")
    lib = _lib.load()
    xs, ys = _as_f32_pos(x), _as_f32_pos(y)
    dev = xs.device
    gx = _gid_from_batch(batch_x, xs.shape[0], dev)
    gy = _gid_from_batch(batch_y, ys.shape[0], dev)
    grid = _Grid(xs, gx, radius)
    m = ys.shape[0]
    qpos = torch.empty((max(m, 1), 4), dtype=torch.float32, device=dev)
    qkeys = torch.empty(max(m, 1), dtype=torch.int64, device=dev)
    check(lib.pgs_bq_pack_queries(ptr(ys), ptr(gy), m, grid.cell, ptr(qpos), ptr(qkeys), stream_ptr()))
    nbr, cnt, dist = _query(grid, qpos, qkeys, m, m, radius, nsample, True)
    idx = torch.empty((m, nsample), dtype=torch.int64, device=dev)
    d2 = torch.empty((m, nsample), dtype=torch.float32, device=dev)
    check(lib.pgs_bq_export(ptr(nbr), ptr(dist), ptr(cnt), m, nsample, ptr(idx), ptr(d2), stream_ptr()))
    return idx, d2


def grow_labels(pos, gid, radius, nsample):
    """Min-ancestor label per point (int32 [N]; -1 where gid < 0) and the neighbour table used."""
    lib = _lib.load()
    n = pos.shape[0]
    dev = pos.device
    grid = _Grid(pos, gid, radius)
    nbr, cnt, _ = _query(grid, grid.spos, grid.skeys, n, n, radius, nsample, False)
    label = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    changed = torch.zeros(1, dtype=torch.int32, device=dev)
    pushed = torch.full((max(n, 1),), -1, dtype=torch.int32, device=dev)     # frontier: label at a row's last push
    check(lib.pgs_rg_init(ptr(gid), n, ptr(label), stream_ptr()))
    for _ in range(4096):
        check(lib.pgs_rg_propagate(ptr(nbr), ptr(cnt), ptr(gid), n, nsample, 2, ptr(label), ptr(pushed), ptr(changed),
                                   stream_ptr()))
        if int(changed) == 0:
            break
    else:
        raise _lib.PgsError("region growing did not converge")
    return label[:n], nbr, cnt


def region_grow(pos, labels, batch, ignore_labels=[], nsample=300, radius=0.03, min_cluster_size=50) -> List[torch.Tensor]:
    """PointGroup region growing (tpk signature).  Returns index tensors (int64, on pos.device) into `pos`:
    classes in ascending order, inside a class clusters by ascending seed (= smallest member), members ascending.
    Defaults are those of torch-points-kernels 0.7.0 `region_grow` (nsample=300, radius=0.03, min_cluster_size=50;
    16 / 0.02 / 32 are the defaults of its `grow_proximity` helper, not of this function): the reference's
    raw-position call sites omit nsample (PointGroup3heads.py:185-192,250-257,340-347) and therefore run with 300."""
    if labels.dim() != 1 or pos.dim() != 2 or pos.shape[0] != labels.shape[0]:
        raise ValueError("pos [N,3] and labels [N] are required")
    p = _as_f32_pos(pos)
    dev = p.device
    n = p.shape[0]
    if n == 0:
        return []
    labels = labels.to(dev)
    batch = batch.to(dev)
    ign = torch.as_tensor(list(ignore_labels) if not torch.is_tensor(ignore_labels) else ignore_labels,
                          device=dev, dtype=labels.dtype).reshape(-1)
    valid = ~torch.isin(labels, ign) if ign.numel() else torch.ones(n, dtype=torch.bool, device=dev)
    lmin = int(labels.min())
    nb = int(batch.max()) + 1
    nl = int(labels.max()) - lmin + 1
    if nb * nl >= _MAX_GID:
        raise ValueError("too many (class, scene) groups for one region_grow call: %d" % (nb * nl))
    gid = torch.where(valid, (labels - lmin) * nb + batch, torch.full_like(labels, -1)).to(torch.int32).contiguous()
    with _lib.nvtx_range("pgs.region_grow.grow_labels"):
        root, _, _ = grow_labels(p, gid, radius, nsample)
    rootl = root.long()
    size = torch.bincount(rootl[valid], minlength=n)
    keep = valid & (size[rootl.clamp_min(0)] >= int(min_cluster_size))
    members = torch.nonzero(keep).squeeze(1)                      # ascending
    if members.numel() == 0:
        return []
    key = (labels[members].long() - lmin) * n + rootl[members]    # class-major, then seed
    order = torch.sort(key, stable=True).indices
    members = members[order]
    _, counts = torch.unique_consecutive(key[order], return_counts=True)
    return list(torch.split(members, counts.tolist()))


def _csr(instance_idx, dev):
    """List of index tensors -> (flat int64, offs int32 [n + 1]) on `dev`."""
    sizes = [int(c.shape[0]) for c in instance_idx]
    offs = torch.zeros(len(sizes) + 1, dtype=torch.int32)
    if sizes:
        offs[1:] = torch.cumsum(torch.tensor(sizes, dtype=torch.int64), 0).to(torch.int32)
    flat = torch.cat([c.reshape(-1) for c in instance_idx]).to(device=dev, dtype=torch.int64) if sizes else \
        torch.zeros(0, dtype=torch.int64, device=dev)
    return flat.contiguous(), offs.to(dev), sizes


def instance_iou(instance_idx: List[torch.Tensor], instance_labels: torch.Tensor, batch=None) -> torch.Tensor:
    """IoU of every proposal against every ground-truth instance (tpk signature; reference call sites:
    torch_points3d/core/losses/panoptic_losses.py:37, every panoptic tracker).
    instance_labels: 0 = no instance, 1..M per scene.  -> f32 [n_proposals, sum_s M_s], scenes in order.
    Counting and division run in csrc/proposals.cu (pgs_prop_gt_iou); the label bookkeeping (instances per scene, their
    sizes) is a handful of reductions over the label vector."""
    lib = _lib.load()
    dev = instance_labels.device
    if not instance_labels.is_cuda:
        raise _lib.PgsError("instance_iou needs CUDA tensors (no CPU path)")
    if batch is None:
        batch = torch.zeros_like(instance_labels)
    n_prop = len(instance_idx)
    nb = int(batch.max()) + 1 if batch.numel() else 0
    per_scene = torch.zeros(nb, dtype=torch.long, device=dev).scatter_reduce(
        0, batch, instance_labels, reduce="amax", include_self=True) if nb else torch.zeros(0, dtype=torch.long, device=dev)
    offs_gt = torch.cumsum(per_scene, 0) - per_scene
    total = int(per_scene.sum()) if nb else 0
    ious = torch.zeros((n_prop, total), dtype=torch.float32, device=dev)
    if n_prop == 0 or total == 0:
        return ious
    gid = torch.where(instance_labels > 0, offs_gt[batch] + instance_labels - 1, torch.full_like(instance_labels, -1))
    gt_size = torch.bincount(gid[gid >= 0], minlength=total).to(torch.int32)
    gt_scene = torch.repeat_interleave(torch.arange(nb, device=dev), per_scene).to(torch.int32)
    flat, offs, sizes = _csr(instance_idx, dev)
    if min(sizes) == 0:
        raise ValueError("empty proposal")
    prop_scene = batch[flat[offs[:-1].long()]].to(torch.int32)
    inter = torch.empty((n_prop, total), dtype=torch.int32, device=dev)
    check(lib.pgs_prop_gt_iou(ptr(flat), ptr(offs), n_prop, flat.shape[0], ptr(gid.to(torch.int32).contiguous()), total,
                              ptr(gt_size), ptr(gt_scene), ptr(prop_scene.contiguous()), ptr(inter), ptr(ious), stream_ptr()))
    return ious


def proposal_nms(instance_idx: List[torch.Tensor], scores: torch.Tensor, threshold: float):
    """Greedy non-maximum suppression over proposals by cross IoU in descending score order
    (models/panoptic/structure_3heads.py:6-17,40-61) -> (keep bool [n_prop], order int64 [n_prop] = proposals by descending
    score).  Intersections are counted from the point -> proposals lists (sorted CSR), not from dense masks."""
    lib = _lib.load()
    dev = scores.device
    n_prop = len(instance_idx)
    flat, offs, sizes = _csr(instance_idx, dev)
    order = torch.argsort(scores.detach().float(), descending=True, stable=True)
    inter = torch.empty((n_prop, n_prop), dtype=torch.int32, device=dev)
    keep = torch.empty(n_prop, dtype=torch.uint8, device=dev)
    nb = lib.pgs_prop_nms_scratch_bytes(flat.shape[0])
    scratch = torch.empty(nb, dtype=torch.uint8, device=dev)
    check(lib.pgs_prop_cross_nms(ptr(flat), ptr(offs), n_prop, flat.shape[0], ptr(order.to(torch.int32).contiguous()),
                                 float(threshold), ptr(inter), ptr(keep), ptr(scratch), nb, stream_ptr()))
    return keep.bool(), order
