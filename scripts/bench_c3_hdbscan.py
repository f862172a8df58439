"""Secondary measurement (BASELINE configs[2], C3): FOR-instance-shape cylinder, embedding head + HDBSCAN(15, 5, eps=0.006)
on the thing points.  Prints one JSON line: GPU seconds per scene at the named size, stage split, and the CPU
stand-in (scikit-learn HDBSCAN, kd_tree, all cores) on a bounded subsample."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from panopticsegforlargescalepointcloud_b200 import scenes, hdbscan, _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
s = scenes.make_scene("forest", n, 0.04, 8.0, seed=0)
_, emb, _ = scenes.synthetic_head_outputs(s, seed=0)
X = emb[s.instance_mask]
dev = torch.device("cuda:0")
Xd = torch.from_numpy(X).to(dev)
m = hdbscan.HDBSCAN(min_cluster_size=15, min_samples=5, cluster_selection_epsilon=0.006)
m.fit_predict(Xd[:20000])                      # warm-up (module load, allocator)
torch.cuda.synchronize()
l0 = _lib.launch_count()
t = time.time(); lab = m.fit_predict(Xd); torch.cuda.synchronize(); dt = time.time() - t
launches = _lib.launch_count() - l0
lab = lab.cpu().numpy()
inst = s.instance_labels[s.instance_mask]
ids = [l for l in np.unique(lab) if l >= 0]
pur = float(np.mean([np.bincount(inst[lab == l]).max() / (lab == l).sum() for l in ids])) if ids else 0.0
out = {"workload": "C3: forest cylinder R=8 m, grid 0.04 m, %d voxels, %d thing points x 5-D embeddings" % (len(s.pos), len(X)),
       "gpu_seconds_per_scene": dt, "scenes_per_s": 1.0 / dt, "boruvka_rounds": m.boruvka_rounds_, "clusters": len(ids),
       "noise_fraction": float((lab < 0).mean()), "mean_cluster_purity": pur, "gpu_launches": launches}
try:
    from sklearn.cluster import HDBSCAN as SK
    k = 30000
    sub = X[np.random.default_rng(0).permutation(len(X))[:k]].astype(np.float64)
    t = time.time(); SK(min_cluster_size=15, min_samples=5, cluster_selection_epsilon=0.006, algorithm="kd_tree", n_jobs=os.cpu_count(), copy=True).fit_predict(sub); ct = time.time() - t
    out["cpu_sklearn"] = {"n_sample": k, "seconds": ct, "cores": os.cpu_count(),
                          "note": "exact Prim MST is O(n^2): extrapolated to the full size = %.0f s" % (ct * (len(X) / k) ** 2)}
except Exception as e:
    out["cpu_sklearn"] = {"error": str(e)}
print(json.dumps(out))
