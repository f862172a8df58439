"""CPU-side checks of the drop-in boundary: libpgs_b200.so loads without a GPU and exports every symbol that
include/pgs_b200.h declares; the ctypes table covers the header; argument validation works without a device;
the host (tree) stage of HDBSCAN -- the only entry point that computes on the host -- matches the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pgs_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pgs_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    from panopticsegforlargescalepointcloud_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "missing export " + n
        assert n in _lib.SIGNATURES, "no ctypes signature for " + n
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.pgs_version() >= 100


def test_size_queries_and_argument_errors_without_gpu():
    from panopticsegforlargescalepointcloud_b200 import _lib
    lib = _lib.load()
    assert lib.pgs_cmap_capacity(1000) == 2048 and lib.pgs_cmap_capacity(0) == 1024
    assert lib.pgs_cmap_build_scratch_bytes(1000) > 3 * 4000
    assert lib.pgs_bq_grid_scratch_bytes(1000) > 0 and lib.pgs_hdb_scratch_bytes(1000, 5) > 0
    # invalid arguments are rejected before any device work
    rc = lib.pgs_kmap_build(None, 10, None, None, 1000, 1, 1, 3, None, None)   # capacity not a power of two
    assert rc == 1 and b"power of two" in lib.pgs_last_error()
    rc = lib.pgs_hdb_mst(None, 1, 5, 5, 1.0, None, None, None, None, None, None, 0, None)
    assert rc == 1 and b"samples" in lib.pgs_last_error()
    rc = lib.pgs_conv_fwd(None, None, None, 10, 27, 4, 4, 0, 0, None, None)      # nbr == NULL with K != 1
    assert rc == 1


def test_ops_fail_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from panopticsegforlargescalepointcloud_b200 import me, tpk, hdbscan, _lib
    with pytest.raises(_lib.PgsError):
        me.SparseTensor(torch.zeros(4, 4), coordinates=torch.zeros(4, 4, dtype=torch.int32))
    with pytest.raises(_lib.PgsError):
        tpk.region_grow(torch.zeros(4, 3), torch.zeros(4, dtype=torch.long), torch.zeros(4, dtype=torch.long))
    with pytest.raises(_lib.PgsError):
        hdbscan.HDBSCAN().fit_predict(np.zeros((10, 3), np.float32))


@pytest.mark.parametrize("seed,eps", [(0, 0.0), (1, 0.006), (2, 0.4)])
def test_host_tree_stage_matches_oracle(seed, eps):
    from oracle import hdbscan_ref as hr
    from panopticsegforlargescalepointcloud_b200 import _lib
    rng = np.random.default_rng(seed)
    mu = rng.normal(0, 3, (7, 5))
    X = (mu[rng.integers(0, 7, 900)] + rng.normal(0, 0.15, (900, 5))).astype(np.float32)
    ref, parts = hr.fit_predict(X, 15, 5, eps, return_parts=True)
    u = parts["u"].astype(np.int32)
    v = parts["v"].astype(np.int32)
    w = parts["w"].astype(np.float64)
    labels = np.empty(len(X), np.int32)
    ncl = np.zeros(1, np.int32)
    rc = _lib.load().pgs_hdb_labels_host(u.ctypes.data, v.ctypes.data, w.ctypes.data, len(X), 15, eps,
                                         labels.ctypes.data, ncl.ctypes.data)
    assert rc == 0
    assert np.array_equal(labels, ref)
    assert ncl[0] == len(set(ref.tolist()) - {-1}) >= 3
