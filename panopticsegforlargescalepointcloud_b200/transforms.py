"""Device versions of the two dataset transforms that sit immediately before the hot path.

Host-side mirror of
  torch_points3d/core/data_transform/grid_transform.py:24-31,33-100,152-198   shuffle_data / group_data / GridSampling3D
  torch_points3d/core/data_transform/transforms.py:385-435                    CylinderSampling
which the reference runs on the CPU inside dataloader workers (torch_cluster.grid_cluster / voxel_grid +
torch_geometric consecutive_cluster, sklearn KDTree.query_radius).  Here every per-point attribute stays on the GPU:

  GridSampling3D : coords = round(pos / size) (half to even, like torch.round in the reference); voxel id as
                   torch_geometric.nn.voxel_grid builds it (x fastest, then y, z, batch; offsets from the per-axis
                   minimum); one row per occupied voxel IN ASCENDING VOXEL-ID ORDER, represented by the LAST point of the
                   voxel in the current row order (consecutive_cluster's scatter_ on the CPU keeps the last writer;
                   `_process` forces mode "last", grid_transform.py:189); the shuffle of mode "last" is an explicit
                   permutation here (argument `perm`, default torch.randperm on the device).
  CylinderSampling: rows with (x - cx)^2 + (y - cy)^2 <= r^2 in ascending row order (KDTree.query_radius returns an
                   unordered set; every consumer is order-free), optional re-centring of x, y.

Voxel de-duplication uses the hot path's own coordinate hash (csrc/cmap.cu, pgs_cmap_build); the remaining sort over the
occupied voxels (torch.argsort: library radix sort) only establishes the reference's output order.  There is no CPU path:
CPU tensors raise.  Semantics of the un-vendored torch_cluster / torch_geometric pieces are restated from memory (SURVEY App. B note):
PARITY UNPINNED against those packages; pinned against oracle/transforms_ref.py, which is checked against an independent
dictionary-based definition.
"""
import torch

from . import _lib


def _need_cuda(t, what):
    if not torch.is_tensor(t) or not t.is_cuda:
        raise _lib.PgsError("%s must be a CUDA tensor: this backend has no CPU path" % what)


def _items(data):
    return list(data.items()) if hasattr(data, "items") else list(vars(data).items())


def _set(data, k, v):
    if hasattr(data, "items"):
        data[k] = v
    else:
        setattr(data, k, v)


def _get(data, k, default=None):
    if hasattr(data, "items"):
        return data.get(k, default)
    return getattr(data, k, default)


def voxel_ids(coords, batch=None):
    """torch_geometric.nn.voxel_grid(coords, batch, size=1) / torch_cluster.grid_cluster: linear cell index with the first
    axis fastest and the batch as the slowest axis, offsets from the per-axis minimum."""
    c = coords.to(torch.int64)
    if batch is not None:
        c = torch.cat([c, batch.to(torch.int64).unsqueeze(1)], 1)
    lo = c.min(0).values
    nv = c.max(0).values - lo + 1
    stride = torch.cumprod(torch.cat([torch.ones(1, dtype=torch.int64, device=c.device), nv[:-1]]), 0)
    return ((c - lo) * stride).sum(1)


def grid_sample_indices(pos, size, batch=None):
    """-> (unique_pos_indices int64 [M] in ascending voxel-id order, cluster int64 [N] consecutive voxel ids,
    coords float [N, 3] = round(pos / size)).

    De-duplication runs on the coordinate-hash kernel of the hot path (pgs_cmap_build: one atomicCAS insert per point,
    first-occurrence voxel numbers), so only the M occupied voxels -- not the N points -- go through a sort, which is
    there for the ORDER contract alone (consecutive_cluster numbers voxels by ascending linear index).  Grids wider than
    65535 cells per axis (7.8 km at 0.12 m) fall back to sorting the N linear indices."""
    _need_cuda(pos, "pos")
    # the reference divides on the CPU (grid_transform.py:185, true fp32 division); torch's CUDA kernel turns `tensor / python
    # scalar` into a multiplication by the reciprocal, which rounds differently near .5 -- divide by a device scalar instead
    coords = torch.round(pos / torch.full((), float(size), dtype=pos.dtype, device=pos.device))
    n = pos.shape[0]
    dev = pos.device
    ci = coords.to(torch.int64)
    lo, hi = ci.min(0).values, ci.max(0).values
    nb = (int(batch.max()) + 1) if batch is not None else 1
    ar = torch.arange(n, device=dev)
    if n > 0 and int((hi - lo).max()) < 65535 and nb < 65536:
        from . import me
        c4 = torch.cat([(batch.to(torch.int64) if batch is not None else torch.zeros(n, dtype=torch.int64, device=dev))
                        .unsqueeze(1), ci - lo - 32767], 1).to(torch.int32).contiguous()
        m, in2out = me.build_coordinate_map(c4, 1)
        first = in2out.long()                                            # voxel number by first occurrence
        vid = voxel_ids(m.coords[:, 1:4], m.coords[:, 0] if batch is not None else None)   # same offsets: min is shared
        order = torch.argsort(vid)                                       # M keys
        rank = torch.empty_like(order)
        rank[order] = torch.arange(order.shape[0], device=dev)
        cluster = rank[first]
        n_vox = order.shape[0]
    else:
        vid = voxel_ids(coords, batch)
        uniq, cluster = torch.unique(vid, sorted=True, return_inverse=True)
        n_vox = uniq.shape[0]
    last = torch.full((n_vox,), -1, dtype=torch.int64, device=dev)
    last.scatter_reduce_(0, cluster, ar, reduce="amax", include_self=True)
    return last, cluster, coords


class GridSampling3D:
    """grid_transform.py:152-212 on device-resident data (dict or attribute container of tensors)."""

    def __init__(self, size, quantize_coords=False, mode="mean", verbose=False, return_inverse=False):
        if mode not in ("mean", "last"):
            raise ValueError("mode must be 'mean' or 'last'")
        self._grid_size = size
        self._quantize_coords = quantize_coords
        self._mode = mode
        self.return_inverse = return_inverse

    def _process(self, data, perm=None):
        pos = _get(data, "pos")
        _need_cuda(pos, "data.pos")
        n = pos.shape[0]
        per_point = [k for k, v in _items(data) if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == n]
        if self._mode == "last":                       # shuffle_data: the representative of a voxel is a random member
            if perm is None:
                perm = torch.randperm(n, device=pos.device)
            for k in per_point:
                _set(data, k, _get(data, k)[perm])
            pos = _get(data, "pos")
        self._mode = "last"                            # grid_transform.py:189 (sic): grouping is always "last"
        idx, cluster, coords = grid_sample_indices(pos, self._grid_size, _get(data, "batch"))
        for k in per_point:
            _set(data, k, _get(data, k)[idx])
        if self._quantize_coords:
            _set(data, "coords", coords[idx].int())
        if self.return_inverse:
            _set(data, "inverse_indices", cluster)
        _set(data, "grid_size", torch.tensor([self._grid_size]))
        return data

    def __call__(self, data, perm=None):
        if isinstance(data, list):
            return [self._process(d, perm) for d in data]
        return self._process(data, perm)

    def __repr__(self):
        return "{}(grid_size={}, quantize_coords={}, mode={})".format(self.__class__.__name__, self._grid_size,
                                                                      self._quantize_coords, self._mode)


class CylinderSampling:
    """transforms.py:385-435: rows inside the vertical cylinder of `radius` around `cylinder_centre` (x, y[, z])."""

    def __init__(self, radius, cylinder_centre, align_origin=True):
        c = torch.as_tensor(cylinder_centre, dtype=torch.float32).reshape(-1)
        self._centre = c[:2]
        self._radius = float(radius)
        self._align_origin = align_origin

    def __call__(self, data):
        pos = _get(data, "pos")
        _need_cuda(pos, "data.pos")
        n = pos.shape[0]
        c = self._centre.to(pos.device)
        d = pos[:, :2].double() - c.double()            # KDTree.query_radius compares in float64
        ind = torch.nonzero((d * d).sum(1) <= self._radius * self._radius).squeeze(1)
        out = type(data)() if not hasattr(data, "items") else {}
        for k, v in _items(data):
            if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == n:
                v = v[ind]
                if self._align_origin and k == "pos":
                    v = v.clone()
                    v[:, :2] -= c.to(v.dtype)
            elif torch.is_tensor(v):
                v = v.clone()
            _set(out, k, v)
        return out

    def __repr__(self):
        return "{}(radius={}, center={}, align_origin={})".format(self.__class__.__name__, self._radius,
                                                                  self._centre.tolist(), self._align_origin)
