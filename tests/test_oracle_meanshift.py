"""Pins oracle/meanshift_ref.py against scikit-learn's MeanShift (the implementation the reference imports)."""
import numpy as np
import pytest
from sklearn.cluster import MeanShift

from oracle import meanshift_ref as mr


@pytest.mark.parametrize("n,D,k,seed,h", [(600, 5, 7, 0, 0.6), (1500, 3, 12, 1, 0.6), (400, 5, 3, 2, 1.0), (300, 2, 5, 3, 0.3)])
def test_oracle_equals_sklearn(n, D, k, seed, h):
    X, _ = mr.blobs(n, D, k, seed)
    ms = MeanShift(bandwidth=h, bin_seeding=True).fit(X)
    labels, centres = mr.mean_shift(X, h)
    assert np.array_equal(labels, ms.labels_)
    assert centres.shape == ms.cluster_centers_.shape and np.allclose(centres, ms.cluster_centers_, atol=1e-6)


def test_oracle_when_binning_fails():
    """Every point in its own bin -> sklearn seeds with the points themselves."""
    X = (np.arange(12, dtype=np.float32).reshape(6, 2) * 10).astype(np.float32)
    with pytest.warns(UserWarning):
        ms = MeanShift(bandwidth=0.6, bin_seeding=True).fit(X)
    labels, centres = mr.mean_shift(X, 0.6)
    assert np.array_equal(labels, ms.labels_) and np.allclose(centres, ms.cluster_centers_)


def test_product_host_stage_selects_sklearn_centres():
    """meanshift.select_centres (the product's host stage, pure numpy) fed with the oracle's converged seeds gives
    scikit-learn's cluster centres in scikit-learn's order."""
    from panopticsegforlargescalepointcloud_b200 import meanshift
    X, _ = mr.blobs(1200, 5, 9, 7)
    seeds = mr.bin_seeds(X, 0.6)
    X64 = X.astype(np.float64)
    res = [mr.single_seed(s, X64, X, 0.6, 300) for s in seeds]
    centres = np.array([r[0] for r in res], np.float32)
    counts = np.array([r[1] for r in res], np.int32)
    got = meanshift.select_centres(centres, counts, 0.6)
    want = MeanShift(bandwidth=0.6, bin_seeding=True).fit(X).cluster_centers_
    assert got.shape == want.shape and np.allclose(got, want, atol=1e-6)
    with pytest.raises(ValueError):
        meanshift.select_centres(centres, np.zeros_like(counts), 0.6)
