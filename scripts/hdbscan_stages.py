"""Stage times of HDBSCAN at the C3 size (thing points of a 500 k-voxel FOR-instance cylinder, 5-D embeddings):
device MST (k-NN + Boruvka + edge sort), device->host copy, host tree stage, labels back.
    python scripts/hdbscan_stages.py > gpurun_out/hdbscan_stages.json"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from panopticsegforlargescalepointcloud_b200 import _lib, hdbscan, scenes
from panopticsegforlargescalepointcloud_b200._lib import check, ptr, stream_ptr

dev = torch.device("cuda:0")
b = bench.make_inputs(0, n=500000, kind="forest", grid=0.04, radius=8.0)
thing = ~np.isin(b.syn_pred, [-1] + list(scenes.stuff_classes("forest")))
X = torch.from_numpy(b.syn_embed[thing]).to(dev).contiguous()
n, D = X.shape
lib = _lib.load()
res = {"n": n, "D": D}
m = hdbscan.HDBSCAN(min_cluster_size=15, min_samples=5, cluster_selection_epsilon=0.006)
m.fit_predict(X[:20000])
for rep in range(3):
    core = torch.empty(n, dtype=torch.float64, device=dev)
    u = torch.empty(n - 1, dtype=torch.int32, device=dev); v = torch.empty(n - 1, dtype=torch.int32, device=dev)
    w = torch.empty(n - 1, dtype=torch.float64, device=dev)
    nb = lib.pgs_hdb_scratch_bytes(n, D)
    scratch = torch.empty(nb, dtype=torch.uint8, device=dev)
    rounds = np.zeros(1, np.int32)
    torch.cuda.synchronize(); t = time.perf_counter()
    check(lib.pgs_hdb_mst(ptr(X), n, D, 6, 1.0, ptr(core), ptr(u), ptr(v), ptr(w), rounds.ctypes.data, ptr(scratch), nb, stream_ptr()))
    torch.cuda.synchronize(); t_mst = time.perf_counter() - t
    t = time.perf_counter()
    u_h = torch.empty(n - 1, dtype=torch.int32, pin_memory=True); v_h = torch.empty(n - 1, dtype=torch.int32, pin_memory=True)
    w_h = torch.empty(n - 1, dtype=torch.float64, pin_memory=True)
    t_alloc = time.perf_counter() - t
    t = time.perf_counter()
    u_h.copy_(u, non_blocking=True); v_h.copy_(v, non_blocking=True); w_h.copy_(w, non_blocking=True)
    torch.cuda.synchronize(); t_d2h = time.perf_counter() - t
    labels_h = torch.empty(n, dtype=torch.int32, pin_memory=True)
    ncl = np.zeros(1, np.int32)
    t = time.perf_counter()
    check(lib.pgs_hdb_labels_host(u_h.data_ptr(), v_h.data_ptr(), w_h.data_ptr(), n, 15, 0.006, labels_h.data_ptr(), ncl.ctypes.data))
    t_tree = time.perf_counter() - t
    t = time.perf_counter(); lab = labels_h.to(dev).long(); torch.cuda.synchronize(); t_h2d = time.perf_counter() - t
    # the same stage on Morton-rank endpoints (what hdbscan.py does): same tree, cache-local leaves
    rank = torch.empty(n, dtype=torch.int32, device=dev)
    check(lib.pgs_hdb_morton_rank(ptr(scratch), n, D, ptr(rank), stream_ptr()))
    ur = rank[u.long()].cpu().pin_memory(); vr = rank[v.long()].cpu().pin_memory()
    labels_r = torch.empty(n, dtype=torch.int32, pin_memory=True)
    t = time.perf_counter()
    check(lib.pgs_hdb_labels_host(ur.data_ptr(), vr.data_ptr(), w_h.data_ptr(), n, 15, 0.006, labels_r.data_ptr(), ncl.ctypes.data))
    t_tree_rank = time.perf_counter() - t
    same = bool(torch.equal(labels_r.to(dev)[rank.long()].long(), lab))
    res["rep%d" % rep] = {"mst_ms": t_mst * 1e3, "pinned_alloc_ms": t_alloc * 1e3, "d2h_ms": t_d2h * 1e3, "tree_host_ms": t_tree * 1e3, "tree_host_morton_rank_ms": t_tree_rank * 1e3, "same_labels": same,
                          "labels_h2d_ms": t_h2d * 1e3, "rounds": int(rounds[0]), "clusters": int(ncl[0])}
st = np.zeros((16, 4), np.int64)
check(lib.pgs_hdb_search_stats(st.ctypes.data, 16))
res["search_stats_per_round [steps, blocks, pairs, full_sweeps]"] = st[:int(rounds[0])].tolist()
res["leaf_blocks"] = (n + 31) // 32
ck = np.zeros((64, 4), np.int64)
check(lib.pgs_hdb_search_stats(ck.ctypes.data, -1))
res["warp_cycles_per_round [sum, max, eval_sum, eval_max]"] = ck[:int(rounds[0])].tolist()
torch.cuda.synchronize(); t = time.perf_counter(); m.fit_predict(X); torch.cuda.synchronize()
res["fit_predict_ms"] = (time.perf_counter() - t) * 1e3
print(json.dumps(res, indent=1))
