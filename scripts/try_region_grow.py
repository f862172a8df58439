"""Time tpk.region_grow on the C2 scene's synthetic head outputs (the call bench.py makes every step)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from panopticsegforlargescalepointcloud_b200 import scenes, tpk
dev = torch.device("cuda:0")
b = bench.make_inputs(0)
d = {k: torch.as_tensor(getattr(b, k)).to(dev) for k in ("syn_shifted", "syn_pred", "batch")}
ignore = [-1] + list(scenes.stuff_classes("urban"))
f = lambda: tpk.region_grow(d["syn_shifted"], d["syn_pred"], d["batch"], ignore_labels=ignore, nsample=200, radius=1.5 * bench.GRID, min_cluster_size=10)
for _ in range(3): c = f()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(10): c = f()
e1.record(); torch.cuda.synchronize()
print("PGS_RG_LANES=%s region_grow %.3f ms per call, %d clusters, %d points" % (os.environ.get("PGS_RG_LANES", "8"), e0.elapsed_time(e1) / 10, len(c), sum(x.numel() for x in c)))
