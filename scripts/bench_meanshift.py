"""Mean shift on synthetic instance embeddings (mu_inst ~ N(0, 3^2 I_5), sigma 0.15, bandwidth 0.6): device
time per call (CUDA events, median of 5) next to scikit-learn's MeanShift on the host (one core, like the reference's
per-scene call), identical labels checked.

    python scripts/bench_meanshift.py [out.json]
"""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from sklearn.cluster import MeanShift as SkMeanShift
from panopticsegforlargescalepointcloud_b200 import meanshift
from oracle import meanshift_ref as mr   # only its synthetic-embedding generator

dev = torch.device("cuda:0")
out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "meanshift_bench.json")
rows = []
for n, k, cpu in ((20000, 60, True), (50000, 90, True), (100000, 90, False), (350000, 60, False)):
    X, _ = mr.blobs(n, 5, k, 0)
    Xd = torch.from_numpy(X).to(dev)
    f = lambda: meanshift.MeanShift(bandwidth=0.6, bin_seeding=True).fit(Xd)
    f(); f()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); got = f(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    r = {"n": n, "D": 5, "instances": k, "clusters": int(got.cluster_centers_.shape[0]), "iterations": got.n_iter_,
         "b200_ms": round(float(np.median(ts)), 2)}
    if cpu:
        t = time.time(); sk = SkMeanShift(bandwidth=0.6, bin_seeding=True).fit(X); r["sklearn_s_1core"] = round(time.time() - t, 2)
        r["labels_equal"] = bool(np.array_equal(got.labels_.cpu().numpy(), sk.labels_))
        r["speedup"] = round(r["sklearn_s_1core"] * 1e3 / r["b200_ms"], 1)
    print(json.dumps(r), flush=True)
    rows.append(r)
json.dump(rows, open(out_path, "w"), indent=1)
