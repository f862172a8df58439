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
