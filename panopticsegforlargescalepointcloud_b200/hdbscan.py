"""hdbscan-shaped class backed by libpgs_b200.so (csrc/hdbscan.cu on the device, csrc/hdbscan_tree.cpp on the host).

Host-side mirror of what the reference imports from the un-vendored `hdbscan` 0.8.27 package:

  hdbscan.HDBSCAN(min_cluster_size=15, min_samples=5, core_dist_n_jobs=1, cluster_selection_epsilon=0.006)
      .fit_predict(X)                                torch_points3d/utils/hdbscan_cluster.py:8-13
  hdbscan_cluster.cluster_single(embeds, unique_in_batch, label_batch, local_ind, type)
                                                     torch_points3d/utils/hdbscan_cluster.py:117-167

Install as a drop-in with `sys.modules["hdbscan"] = panopticsegforlargescalepointcloud_b200.hdbscan`.
`fit_predict` accepts a numpy array (returns numpy, like upstream) or a CUDA tensor (returns a CUDA tensor).
Stages: k-NN core distances + exact mutual-reachability MST on the device in float64 (pgs_hdb_mst), then the
sequential tree stage on the host over pinned buffers (pgs_hdb_labels_host), labels copied back.
Semantics frozen in DESIGN.md ("HDBSCAN determinism"): core distance = min_samples-th neighbour not counting the
sample itself (hdbscan 0.8.27; `core_includes_self=True` gives scikit-learn's convention), edges strictly
ordered by (weight, min id, max id), EOM selection, allow_single_cluster=False.  No CPU implementation of the
device stages exists in this package.
"""
import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


class HDBSCAN:
    def __init__(self, min_cluster_size=5, min_samples=None, cluster_selection_epsilon=0.0, alpha=1.0,
                 metric="euclidean", core_dist_n_jobs=None, cluster_selection_method="eom",
                 allow_single_cluster=False, approx_min_span_tree=True, algorithm="best", leaf_size=40,
                 core_includes_self=False, **kwargs):
        if metric != "euclidean":
            raise NotImplementedError("only the euclidean metric is on the reference hot path")
        if cluster_selection_method != "eom":
            raise NotImplementedError("only cluster_selection_method='eom' is on the reference hot path")
        if allow_single_cluster:
            raise NotImplementedError("allow_single_cluster=True is not used by the reference")
        self.min_cluster_size = int(min_cluster_size)
        self.min_samples = self.min_cluster_size if min_samples is None else int(min_samples)
        self.cluster_selection_epsilon = float(cluster_selection_epsilon)
        self.alpha = float(alpha)
        # hdbscan 0.8.27 takes the min_samples-th neighbour NOT counting the sample itself (all of its MST front ends
        # query k = min_samples + 1 and read column min_samples); scikit-learn's HDBSCAN counts it.  The default is
        # the library the reference imports; True = scikit-learn's convention (what the sklearn goldens pin).
        self.core_includes_self = bool(core_includes_self)
        self.labels_ = None
        self.mst_ = None
        self.core_distances_ = None
        self.boruvka_rounds_ = None
        self.device = kwargs.get("device", None)

    def _run(self, X):
        lib = _lib.load()
        n, D = X.shape
        dev = X.device
        if n < 2:
            raise ValueError("HDBSCAN requires more than one sample")
        if not 1 <= D <= 8:
            raise NotImplementedError("1..8 feature dimensions are supported (the reference uses 3 and 5)")
        # hdbscan.hdbscan_(): min_samples = min(n - 1, min_samples), at least 1; k = rank counting the sample itself
        k = max(min(n - 1, self.min_samples), 1)
        if not self.core_includes_self:
            k += 1
        if k > 32:
            raise NotImplementedError("min_samples > 31 is not supported by the device k-NN sweep")
        core = torch.empty(n, dtype=torch.float64, device=dev)
        u = torch.empty(n - 1, dtype=torch.int32, device=dev)
        v = torch.empty(n - 1, dtype=torch.int32, device=dev)
        w = torch.empty(n - 1, dtype=torch.float64, device=dev)
        nb = lib.pgs_hdb_scratch_bytes(n, D)
        scratch = torch.empty(nb, dtype=torch.uint8, device=dev)
        rounds = np.zeros(1, np.int32)
        with _lib.nvtx_range("pgs.hdbscan.mst"):
            check(lib.pgs_hdb_mst(ptr(X), n, D, k, self.alpha, ptr(core), ptr(u), ptr(v), ptr(w),
                                  rounds.ctypes.data, ptr(scratch), nb, stream_ptr()))
        self.boruvka_rounds_ = int(rounds[0])
        # host tree stage over pinned buffers.  The endpoints go to the host as MORTON RANKS (each edge keeps its (u, v)
        # order, so the dendrogram is the same tree with relabelled leaves): spatial neighbours become index neighbours
        # and the sequential union-find / condensed-tree walk stays in cache; labels come back per rank.
        rank = torch.empty(n, dtype=torch.int32, device=dev)
        check(lib.pgs_hdb_morton_rank(ptr(scratch), n, D, ptr(rank), stream_ptr()))
        rl = rank.long()
        u_h = torch.empty(n - 1, dtype=torch.int32, pin_memory=True)
        v_h = torch.empty(n - 1, dtype=torch.int32, pin_memory=True)
        w_h = torch.empty(n - 1, dtype=torch.float64, pin_memory=True)
        u_h.copy_(rank[u.long()], non_blocking=True)
        v_h.copy_(rank[v.long()], non_blocking=True)
        w_h.copy_(w, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        labels_r = torch.empty(n, dtype=torch.int32, pin_memory=True)
        ncl = np.zeros(1, np.int32)
        with _lib.nvtx_range("pgs.hdbscan.tree_host"):
            check(lib.pgs_hdb_labels_host(u_h.data_ptr(), v_h.data_ptr(), w_h.data_ptr(), n, self.min_cluster_size,
                                          self.cluster_selection_epsilon, labels_r.data_ptr(), ncl.ctypes.data))
        labels_d = labels_r.to(dev, non_blocking=True)[rl]            # label of input row i = label of its rank
        self.core_distances_ = core
        self.mst_ = (u, v, w)
        self.n_clusters_ = int(ncl[0])
        return labels_d

    def fit(self, X, y=None):
        self.fit_predict(X)
        return self

    def fit_predict(self, X, y=None):
        as_numpy = not torch.is_tensor(X)
        if as_numpy:
            if not torch.cuda.is_available():
                raise _lib.PgsError("HDBSCAN needs a CUDA device: this backend has no CPU path")
            Xt = torch.as_tensor(np.ascontiguousarray(X, dtype=np.float32)).to(self.device or "cuda")
        else:
            if not X.is_cuda:
                raise _lib.PgsError("HDBSCAN input tensor must be on a CUDA device (no CPU path)")
            Xt = X.detach().to(torch.float32).contiguous()
        if Xt.dim() != 2:
            raise ValueError("X must be [n_samples, n_features]")
        labels_d = self._run(Xt)
        if as_numpy:
            self.labels_ = labels_d.cpu().numpy().astype(np.int64)
        else:
            self.labels_ = labels_d.long()
        return self.labels_


def hdbscan_cluster(prediction, min_cluster_size=15, min_samples=5, cluster_selection_epsilon=0.006):
    """utils/hdbscan_cluster.py:8-13."""
    return HDBSCAN(min_cluster_size=min_cluster_size, min_samples=min_samples, core_dist_n_jobs=1,
                   cluster_selection_epsilon=cluster_selection_epsilon).fit_predict(prediction)


def cluster_single(embed_logits_logits_u, unique_in_batch, label_batch, local_ind, type, **kw):
    """utils/hdbscan_cluster.py:117-167 without the CPU round trip and the per-call process pool: per scene with
    more than 3 points, HDBSCAN on the raw block; one index tensor per non-noise label, label ascending.
    -> (List[LongTensor] on the input device, List[type])."""
    final_result, cluster_type = [], []
    for s in unique_in_batch.tolist() if torch.is_tensor(unique_in_batch) else list(unique_in_batch):
        mask = label_batch == s
        sample_local = local_ind[mask]
        if sample_local.shape[0] > 3:
            labels = hdbscan_cluster(embed_logits_logits_u[mask], **kw)
            order = torch.sort(labels, stable=True)
            lab_sorted = order.values
            keep = lab_sorted >= 0
            if not bool(keep.any()):
                continue
            idx = sample_local[order.indices[keep]]
            _, counts = torch.unique_consecutive(lab_sorted[keep], return_counts=True)
            parts = torch.split(idx, counts.tolist())
            final_result.extend(parts)
            cluster_type.extend([type] * len(parts))
    return final_result, cluster_type
