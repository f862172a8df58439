"""Freeze the CPU oracle's HDBSCAN answers at the benchmark sizes (50 k and 350 k x 5-D) under tests/golden/.

    python scripts/make_golden_hdbscan_big.py [50k] [350k]

Oracle = oracle/hdbscan_ref.py with the multithreaded exact kNN / Prim of oracle/c/hdbscan_big.c (reference call
site torch_points3d/utils/hdbscan_cluster.py:8-13: HDBSCAN(15, 5, eps=0.006), hdbscan's core-distance rank).
Stored: labels (int16, compressed) and sha256 digests of the float64 core distances and of the canonical MST
(u int32, v int32, w float64, sorted by the strict order) -- the arrays themselves would be 7 MB.
350 k takes ~5 min on 8 cores.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hdb_big_inputs as inp  # noqa: E402
from oracle import hdbscan_ref as hr  # noqa: E402

for name in (sys.argv[1:] or ["50k", "350k"]):
    X, owner = inp.make(name)
    t = time.time()
    labels, parts = hr.fit_predict(X, 15, 5, 0.006, return_parts=True, threads=os.cpu_count())
    dt = time.time() - t
    ncl = len(set(labels.tolist()) - {-1})
    assert ncl < 32000
    out = os.path.join(ROOT, "tests", "golden", "hdbscan_big_%s.npz" % name)
    np.savez_compressed(out, labels=labels.astype(np.int16), n_clusters=ncl,
                        x_sha=inp.digest(X), core_sha=inp.digest(parts["core"]),
                        mst_sha=inp.digest(parts["u"].astype(np.int32), parts["v"].astype(np.int32), parts["w"]),
                        w_sum=float(parts["w"].sum()), w_max=float(parts["w"].max()))
    print(name, "n=%d clusters=%d noise=%d  %.1f s  -> %s (%d bytes)" % (
        len(X), ncl, int((labels < 0).sum()), dt, out, os.path.getsize(out)))
