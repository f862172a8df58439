"""The CUDA path replayed on the committed golden fixtures (tests/golden/*.npz: scikit-learn MeanShift / HDBSCAN labels,
torch dense conv3d outputs, scipy cKDTree radius sets -- scripts/make_golden.py).  Independent of the oracle and of the
libraries installed on the GPU box."""
import numpy as np
import pytest
import torch

from test_golden import load, neighbour_rows
from test_oracle_hdbscan import _same_partition_up_to_ties

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("i", [0, 1, 2])
def test_meanshift_vs_golden(cuda_device, i):
    from panopticsegforlargescalepointcloud_b200 import meanshift
    g = load("meanshift")
    m = meanshift.MeanShift(bandwidth=float(g["h%d" % i]), bin_seeding=True).fit(torch.from_numpy(g["X%d" % i]).to(cuda_device))
    assert np.array_equal(m.labels_.cpu().numpy(), g["labels%d" % i])
    assert np.allclose(m.cluster_centers_.cpu().numpy(), g["centres%d" % i], atol=1e-5)


@pytest.mark.parametrize("i", [0, 1, 2])
def test_hdbscan_vs_golden(cuda_device, i):
    from panopticsegforlargescalepointcloud_b200 import hdbscan
    g = load("hdbscan")
    m = hdbscan.HDBSCAN(min_cluster_size=15, min_samples=5, cluster_selection_epsilon=0.006, core_includes_self=True)
    got = m.fit_predict(g["X%d" % i])                      # numpy in -> numpy out, like upstream
    u, v, w = (t.cpu().numpy() for t in m.mst_)
    assert _same_partition_up_to_ties(got, g["labels%d" % i], dict(u=u, v=v, w=w))


@pytest.mark.parametrize("impl", ["auto", "mma", "tc", "split", "ffma"])
def test_conv_vs_golden(cuda_device, monkeypatch, impl):
    from panopticsegforlargescalepointcloud_b200 import me
    g = load("conv_dense")
    monkeypatch.setattr(me, "CONV_IMPL", impl)
    monkeypatch.setattr(me, "SORT_MIN_ROWS", 0)
    mgr = me.CoordinateManager(torch.from_numpy(g["coords"]).to(cuda_device))
    km = mgr.kernel_map(1, 1, 1, 1, 3)
    X = torch.from_numpy(g["X"]).to(cuda_device)
    W = torch.from_numpy(g["W"]).to(cuda_device)
    Y = me._conv_fwd_raw(X, W, km, km.n_q, 0, 0)
    Yt = me._conv_fwd_raw(X, W, km, km.n_q, 1, 0)          # transposed stride-1 conv == mirrored offsets
    assert np.allclose(Y.cpu().numpy(), g["Y"], atol=1e-4, rtol=1e-4)
    assert np.allclose(Yt.cpu().numpy(), g["Y_transposed"], atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize("nsample", [200, 8])
def test_ball_query_vs_golden(cuda_device, nsample):
    from panopticsegforlargescalepointcloud_b200 import tpk
    g = load("ball_query")
    p = torch.from_numpy(g["pos"]).to(cuda_device)
    b = torch.from_numpy(g["batch"]).to(cuda_device)
    idx, _ = tpk.ball_query(float(g["radius"]), nsample, p, p, mode="PARTIAL_DENSE", batch_x=b, batch_y=b)
    idx = idx.cpu().numpy()
    for row, want in zip(idx, neighbour_rows(g)):
        w = want[:nsample]
        assert np.array_equal(row[:len(w)], w) and np.all(row[len(w):] == -1)
