"""Per-launch durations of the HDBSCAN device kernels at the C3 size (torch.profiler / CUPTI).
    python scripts/hdbscan_kernels.py > gpurun_out/hdbscan_kernels.json"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
import bench
from panopticsegforlargescalepointcloud_b200 import hdbscan, scenes

dev = torch.device("cuda:0")
b = bench.make_inputs(0, n=500000, kind="forest", grid=0.04, radius=8.0)
thing = ~np.isin(b.syn_pred, [-1] + list(scenes.stuff_classes("forest")))
X = torch.from_numpy(b.syn_embed[thing]).to(dev).contiguous()
m = hdbscan.HDBSCAN(min_cluster_size=15, min_samples=5, cluster_selection_epsilon=0.006)
m.fit_predict(X[:20000])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    m.fit_predict(X)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type.name == "CUDA"]
ev.sort(key=lambda e: e.time_range.start)
rows = [{"kernel": e.name[:60], "us": round(e.device_time if hasattr(e, "device_time") else e.cuda_time, 1)} for e in ev]
agg = {}
for r in rows:
    a = agg.setdefault(r["kernel"], [0, 0.0]); a[0] += 1; a[1] += r["us"]
print(json.dumps({"n": int(X.shape[0]), "by_kernel": {k: {"launches": v[0], "total_ms": round(v[1] / 1e3, 3)} for k, v in
                  sorted(agg.items(), key=lambda kv: -kv[1][1])},
                  "search_launches_us": [r["us"] for r in rows if "hdb_search" in r["kernel"]],
                  "knn_us": [r["us"] for r in rows if "hdb_knn" in r["kernel"]]}, indent=1))
