"""One pass over the non-convolution kernels of the path on C2 / C3-shaped inputs, for
    ncu --set full --clock-control none -k regex:'cmap_insert|kmap_build|kmap_masks|ball_query|rg_push|hdb_|ms_iterate|ms_assign|bn_stats|bn_bwd_stats' -c 40 -o gpurun_out/other python scripts/ncu_other_kernels.py
(coordinate hash build, rulebook build, occupancy masks, ball query, label propagation, HDBSCAN kNN + Boruvka search,
mean shift, BatchNorm statistics)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from panopticsegforlargescalepointcloud_b200 import me, scenes, tpk, hdbscan, meanshift, _lib
from panopticsegforlargescalepointcloud_b200._lib import ptr, check, stream_ptr
from oracle import meanshift_ref as mr   # synthetic embedding generator only

dev = torch.device("cuda:0")
b = bench.make_inputs(0)
coords = torch.cat([torch.zeros(len(b.coords), 1, dtype=torch.int32), torch.as_tensor(b.coords).int()], 1).to(dev)
mgr = me.CoordinateManager(coords)                      # cmap_insert / flag / emit
km = mgr.kernel_map(1, 1, 1, 1, 3)                      # kmap_build
km.sorted()                                             # kmap_masks + sort + permute
ignore = [-1] + list(scenes.stuff_classes("urban"))
d = {k: torch.as_tensor(getattr(b, k)).to(dev) for k in ("syn_shifted", "syn_pred", "batch")}
tpk.region_grow(d["syn_shifted"], d["syn_pred"], d["batch"], ignore_labels=ignore, nsample=200, radius=1.5 * bench.GRID,
                min_cluster_size=10)                    # ball_query + rg_push / rg_jump
X, _ = mr.blobs(50000, 5, 90, 0)
Xd = torch.from_numpy(X).to(dev)
hdbscan.HDBSCAN(min_cluster_size=15, min_samples=5, cluster_selection_epsilon=0.006).fit_predict(Xd)   # hdb_knn / hdb_search
meanshift.MeanShift(bandwidth=0.6, bin_seeding=True).fit(Xd)                                            # ms_iterate / ms_assign
lib = _lib.load()
n, C = 200000, 16
F = torch.randn(n, C, device=dev); Y = torch.empty_like(F); dY = torch.randn(n, C, device=dev); dX = torch.empty_like(F)
w = torch.ones(C, device=dev); bb = torch.zeros(C, device=dev); rm = torch.zeros(C, device=dev); rv = torch.ones(C, device=dev)
sums = torch.empty(2 * C, dtype=torch.float64, device=dev); st = torch.empty(2, C, device=dev); dwb = torch.empty(2, C, device=dev)
check(lib.pgs_bn_forward(ptr(F), n, C, ptr(w), ptr(bb), ptr(rm), ptr(rv), 1, 0.1, 1e-5, 1, ptr(sums), ptr(st[0]), ptr(st[1]), ptr(Y), stream_ptr()))
check(lib.pgs_bn_backward(ptr(F), ptr(Y), ptr(dY), n, C, ptr(w), ptr(st[0]), ptr(st[1]), 1, 1, ptr(sums), ptr(dX), ptr(dwb[0]), ptr(dwb[1]), stream_ptr()))
torch.cuda.synchronize()
