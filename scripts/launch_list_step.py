"""One bench step (C2 workload, inputs resident) between cudaProfilerStart/Stop, for
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python scripts/launch_list_step.py [cylinders per step]
Same model / step function as bench.py (device-resident arm)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from panopticsegforlargescalepointcloud_b200 import panoptic, parallel, scenes, tpk
dev = torch.device("cuda:0")
torch.manual_seed(2022)
opt = panoptic.paper_options("urban", cluster_type=1, grid=bench.GRID, use_score_net=True, prepare_epoch=30, scorer=False)
model = panoptic.PointGroup(opt, "dummy", panoptic.DatasetProperties("urban"), None).to(dev)
model.instantiate_optimizers({}); model.train()
dp = parallel.DataParallelStep(model)
ignore = [-1] + list(scenes.stuff_classes("urban"))
data = []
SPR = int(sys.argv[1]) if len(sys.argv) > 1 else bench.REF_BATCH_SIZE      # cylinders per step (bench.py default)
for i in range(2):
    b = bench.make_inputs(list(range(i * SPR, (i + 1) * SPR)))
    data.append({k: torch.as_tensor(getattr(b, k)).to(dev) for k in bench.HOST_KEYS})
class View:
    def __init__(self, d): self.__dict__.update(d)
    def __getitem__(self, k): return self.__dict__[k]
def step(i):
    d = data[i % 2]
    dp.step(View(d), epoch=1, step=i, batch_size=SPR)
    return tpk.region_grow(d["syn_shifted"], d["syn_pred"], d["batch"], ignore_labels=ignore, nsample=200,
                           radius=1.5 * bench.GRID, min_cluster_size=10)
for i in range(3):
    step(i)
torch.cuda.synchronize()
torch.cuda.profiler.start()
step(3)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
