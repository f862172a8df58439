"""sklearn-shaped MeanShift backed by libpgs_b200.so (csrc/meanshift.cu) + the reference's per-scene fan-out.

Host-side mirror of
  sklearn.cluster.MeanShift(bandwidth=0.6, bin_seeding=True).fit(X).labels_     (scikit-learn, pinned 0.24.2 in the
                                                                                 reference's poetry.lock:2028-2029)
  torch_points3d/utils/meanshift_cluster.py:9-18    meanshift_cluster(prediction, bandwidth)
  torch_points3d/utils/meanshift_cluster.py:72-123  cluster_single(embeds, unique_in_batch, label_batch, local_ind, type, bandwidth)

Stages (restated from sklearn/cluster/_mean_shift.py, see csrc/meanshift.cu): grid-bin seeding (tensor ops), mode
seeking of every seed (pgs_ms_iterate, device), removal of near-duplicate centres (host, a few thousand centres, strictly
sequential like the reference), nearest-centre labels (pgs_ms_assign, device).  `fit` accepts a numpy array (returns
numpy attributes, like sklearn) or a CUDA tensor (returns CUDA tensors).  There is no CPU implementation of the device
stages in this package.
"""
import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


def select_centres(centres, counts, bandwidth):
    """Host stage (_mean_shift.py:513-545): converged (centre, points-within-bandwidth) pairs -> cluster centres.
    Seeds with no point nearby are dropped, the rest sorted by (intensity, centre tuple) descending, and every centre
    within `bandwidth` of an earlier kept one is removed.  A few thousand rows, strictly sequential, numpy."""
    keep = counts > 0
    if not keep.any():
        raise ValueError("No point was within bandwidth=%f of any seed." % bandwidth)
    c, k = centres[keep], counts[keep]
    D = c.shape[1]
    order = np.lexsort(tuple(c[:, d] for d in range(D - 1, -1, -1)) + (k,))[::-1]
    c = c[order]
    c64 = c.astype(np.float64)
    unique = np.ones(len(c), dtype=bool)
    h2 = float(bandwidth) * float(bandwidth)
    for i in range(len(c)):
        if unique[i]:
            d2 = ((c64 - c64[i]) ** 2).sum(1)
            unique[d2 <= h2] = False
            unique[i] = True
    return np.ascontiguousarray(c[unique])


class MeanShift:
    def __init__(self, *, bandwidth=None, seeds=None, bin_seeding=False, min_bin_freq=1, cluster_all=True, n_jobs=None,
                 max_iter=300, device=None):
        if bandwidth is None:
            raise NotImplementedError("estimate_bandwidth is not on the reference hot path: pass bandwidth")
        if seeds is not None:
            raise NotImplementedError("explicit seeds are not used by the reference")
        self.bandwidth = float(bandwidth)
        self.bin_seeding = bool(bin_seeding)
        self.min_bin_freq = int(min_bin_freq)
        self.cluster_all = bool(cluster_all)
        self.max_iter = int(max_iter)
        self.device = device
        self.labels_ = self.cluster_centers_ = self.n_iter_ = None

    # ---- stages ----
    def _seeds(self, X):
        """get_bin_seeds (_mean_shift.py:247-297): occupied cells of a grid with spacing `bandwidth`."""
        if not self.bin_seeding:
            return X
        h = torch.tensor(self.bandwidth, dtype=torch.float32, device=X.device)
        bins, freq = torch.unique(torch.round(X / h), dim=0, return_counts=True)   # round half to even, like np.round
        bins = bins[freq >= self.min_bin_freq]
        if bins.shape[0] == X.shape[0]:
            return X           # "Binning data failed ... using data points as seeds"
        return (bins * h).contiguous()

    def _run(self, X):
        lib = _lib.load()
        n, D = X.shape
        dev = X.device
        if n == 0:
            raise ValueError("MeanShift requires at least one sample")
        if not 1 <= D <= 8:
            raise NotImplementedError("1..8 feature dimensions are supported (the reference uses 5)")
        seeds = self._seeds(X).contiguous()
        ns = seeds.shape[0]
        centers = torch.empty((ns, D), dtype=torch.float32, device=dev)
        counts = torch.empty(ns, dtype=torch.int32, device=dev)
        iters = torch.empty(ns, dtype=torch.int32, device=dev)
        check(lib.pgs_ms_iterate(ptr(X), n, D, ptr(seeds), ns, self.bandwidth, self.max_iter, ptr(centers), ptr(counts),
                                 ptr(iters), stream_ptr()))
        self.n_iter_ = int(iters.max())
        cc = select_centres(centers.cpu().numpy(), counts.cpu().numpy(), self.bandwidth)
        cdev = torch.from_numpy(cc).to(dev)
        labels = torch.empty(n, dtype=torch.int32, device=dev)
        dist = None if self.cluster_all else torch.empty(n, dtype=torch.float64, device=dev)
        check(lib.pgs_ms_assign(ptr(X), n, D, ptr(cdev), cc.shape[0], ptr(labels), ptr(dist), stream_ptr()))
        labels = labels.long()
        if not self.cluster_all:
            labels = torch.where(dist <= self.bandwidth, labels, torch.full_like(labels, -1))
        return labels, cdev

    # ---- sklearn surface ----
    def fit(self, X, y=None):
        is_np = not torch.is_tensor(X)
        Xt = torch.as_tensor(np.ascontiguousarray(X, dtype=np.float32)) if is_np else X
        if Xt.dim() != 2:
            raise ValueError("X must be [n_samples, n_features]")
        if not Xt.is_cuda:
            if not torch.cuda.is_available():
                raise _lib.PgsError("MeanShift needs a CUDA device: this backend has no CPU path")
            Xt = Xt.to(self.device or "cuda")
        Xt = Xt.detach().to(torch.float32).contiguous()
        labels, centers = self._run(Xt)
        if is_np:
            self.labels_, self.cluster_centers_ = labels.cpu().numpy(), centers.cpu().numpy()
        else:
            self.labels_, self.cluster_centers_ = labels, centers
        return self

    def fit_predict(self, X, y=None):
        return self.fit(X).labels_


def meanshift_cluster(prediction, bandwidth):
    """utils/meanshift_cluster.py:9-18 -> label tensor (on the input's device for tensors)."""
    labels = MeanShift(bandwidth=bandwidth, bin_seeding=True).fit(prediction).labels_
    return labels if torch.is_tensor(labels) else torch.from_numpy(labels)


def cluster_single(embed_logits_logits_u, unique_in_batch, label_batch, local_ind, type, bandwidth):
    """utils/meanshift_cluster.py:72-123 without the multiprocessing.Pool / CPU round trip: per scene with more than 3
    thing points, mean shift on the embeddings; clusters in ascending label order, members as `local_ind` entries."""
    final_result, cluster_type = [], []
    for s in unique_in_batch.tolist():
        m = label_batch == s
        if int(m.sum()) > 3:
            idx = local_ind[m]
            labels = meanshift_cluster(embed_logits_logits_u[m].detach(), bandwidth)
            order = torch.sort(labels, stable=True).indices
            _, counts = torch.unique_consecutive(labels[order], return_counts=True)
            for part in torch.split(idx[order], counts.tolist()):
                final_result.append(part)
                cluster_type.append(type)
    return final_result, cluster_type
