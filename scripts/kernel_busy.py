"""GPU busy time vs wall time of the C2 step (torch.profiler / CUPTI): sum of kernel durations per step, by kernel, and the
idle share -- tells whether the step is bound by kernels or by launch gaps / host syncs.
    python scripts/kernel_busy.py [scenes_per_step] > gpurun_out/kernel_busy.json"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from panopticsegforlargescalepointcloud_b200 import panoptic, parallel, scenes, tpk

spr = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda:0")
torch.manual_seed(2022)
opt = panoptic.paper_options("urban", cluster_type=1, grid=bench.GRID, use_score_net=True, prepare_epoch=30, scorer=False)
model = panoptic.PointGroup(opt, "dummy", panoptic.DatasetProperties("urban"), None).to(dev)
model.instantiate_optimizers({})
model.train()
dp = parallel.DataParallelStep(model)
ignore = [-1] + list(scenes.stuff_classes("urban"))
pool = []
for i in range(3):
    b = bench.make_inputs(list(range(i * spr, (i + 1) * spr)))
    pool.append({k: torch.as_tensor(getattr(b, k)).to(dev) for k in bench.HOST_KEYS})


class View:
    def __init__(self, d):
        self.__dict__.update(d)

    def __getitem__(self, k):
        return self.__dict__[k]


def step(i):
    d = pool[i % 3]
    dp.step(View(d), epoch=1, step=i, batch_size=spr)
    tpk.region_grow(d["syn_shifted"], d["syn_pred"], d["batch"], ignore_labels=ignore, nsample=200, radius=1.5 * bench.GRID,
                    min_cluster_size=10)


for i in range(4):
    step(i)
torch.cuda.synchronize()
reps = 4
t0 = time.perf_counter()
for i in range(reps):
    step(i)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) * 1e3 / reps
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(reps):
        step(i)
    torch.cuda.synchronize()
rows = []
busy = 0.0
for e in prof.key_averages():
    t = getattr(e, "device_time_total", None)
    if t is None:
        t = getattr(e, "cuda_time_total", 0.0)
    if t and e.device_type.name == "CUDA" if hasattr(e, "device_type") else t:
        rows.append((e.key[:90], e.count / reps, t / 1e3 / reps))
rows.sort(key=lambda r: -r[2])
busy = sum(r[2] for r in rows)
print(json.dumps({"scenes_per_step": spr, "wall_ms_per_step_unprofiled": wall, "kernel_busy_ms_per_step": busy,
                  "idle_share": 1 - busy / wall, "kernels_per_step": sum(r[1] for r in rows),
                  "top": [{"kernel": r[0], "per_step": r[1], "ms_per_step": round(r[2], 4)} for r in rows[:45]]}, indent=1))
