"""Device mean shift (csrc/meanshift.cu behind meanshift.py) against the CPU oracle and scikit-learn itself: identical
partitions and label order, centres to 1e-5 (the device averages in fp64 and rounds to fp32, numpy averages in fp32)."""
import numpy as np
import pytest
import torch
from sklearn.cluster import MeanShift as SkMeanShift

from oracle import meanshift_ref as mr

pytestmark = pytest.mark.gpu


def _ms():
    from panopticsegforlargescalepointcloud_b200 import meanshift
    return meanshift


@pytest.mark.parametrize("n,D,k,seed,h", [(600, 5, 7, 0, 0.6), (1500, 3, 12, 1, 0.6), (400, 5, 3, 2, 1.0), (300, 2, 5, 3, 0.3),
                                          (20000, 5, 60, 4, 0.6)])
def test_labels_equal_oracle_and_sklearn(cuda_device, n, D, k, seed, h):
    X, _ = mr.blobs(n, D, k, seed)
    got = _ms().MeanShift(bandwidth=h, bin_seeding=True).fit(torch.from_numpy(X).to(cuda_device))
    labels, centres = mr.mean_shift(X, h) if n <= 2000 else (None, None)
    sk = SkMeanShift(bandwidth=h, bin_seeding=True).fit(X)
    g = got.labels_.cpu().numpy()
    assert g.dtype == np.int64 and np.array_equal(g, sk.labels_)
    assert np.allclose(got.cluster_centers_.cpu().numpy(), sk.cluster_centers_, atol=1e-5)
    if labels is not None:
        assert np.array_equal(g, labels)


def test_numpy_in_numpy_out_and_edge_cases(cuda_device):
    ms = _ms()
    X, _ = mr.blobs(500, 5, 4, 9)
    a = ms.MeanShift(bandwidth=0.6, bin_seeding=True).fit(X)              # numpy in -> numpy attributes, like sklearn
    assert isinstance(a.labels_, np.ndarray) and np.array_equal(a.labels_, SkMeanShift(bandwidth=0.6, bin_seeding=True).fit(X).labels_)
    same = np.ones((50, 5), np.float32)                                   # all points identical: one cluster
    assert set(ms.MeanShift(bandwidth=0.6, bin_seeding=True).fit(same).labels_.tolist()) == {0}
    far = (np.arange(12, dtype=np.float32).reshape(6, 2) * 10)            # binning "fails": points seed themselves
    with pytest.warns(UserWarning):
        ref = SkMeanShift(bandwidth=0.6, bin_seeding=True).fit(far).labels_
    assert np.array_equal(ms.MeanShift(bandwidth=0.6, bin_seeding=True).fit(far).labels_, ref)
    with pytest.raises(Exception):
        ms.MeanShift(bandwidth=0.6, bin_seeding=True).fit(torch.zeros(4, 9, device=cuda_device))   # D > 8
    nc = ms.MeanShift(bandwidth=0.6, bin_seeding=True, cluster_all=False).fit(X)
    assert np.array_equal(nc.labels_, SkMeanShift(bandwidth=0.6, bin_seeding=True, cluster_all=False).fit(X).labels_)


def test_cluster_single_matches_reference_fan_out(cuda_device):
    """utils/meanshift_cluster.py:72-123 restated with sklearn on the CPU: same clusters in the same order."""
    ms = _ms()
    rng = np.random.default_rng(3)
    parts, batch = [], []
    for s, (n, k) in enumerate([(900, 6), (3, 1), (1200, 9)]):            # the 3-point scene is skipped (needs > 3)
        X, _ = mr.blobs(n, 5, k, 20 + s)
        parts.append(X); batch.append(np.full(n, s))
    E = np.concatenate(parts); b = np.concatenate(batch)
    perm = rng.permutation(len(E)); E, b = E[perm], b[perm]
    local = np.sort(rng.choice(10 * len(E), len(E), replace=False))       # indices into the full cloud
    want = []
    for s in np.unique(b):
        m = b == s
        if m.sum() > 3:
            lab = SkMeanShift(bandwidth=0.6, bin_seeding=True).fit(E[m]).labels_
            for l in np.unique(lab):
                want.append(local[m][lab == l])
    dev = cuda_device
    got, types = ms.cluster_single(torch.from_numpy(E).to(dev), torch.unique(torch.from_numpy(b)).to(dev), torch.from_numpy(b).to(dev),
                                   torch.from_numpy(local).to(dev), 1, 0.6)
    assert len(got) == len(want) and types == [1] * len(want)
    for g, w in zip(got, want):
        assert np.array_equal(g.cpu().numpy(), w)
