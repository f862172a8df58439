"""Where does the C2 step spend its time?  Per phase: host enqueue time (no sync) and device time (sync after the phase).
    python scripts/phase_profile.py [n_voxels] > gpurun_out/phase_profile.json"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from panopticsegforlargescalepointcloud_b200 import _lib, me, panoptic, parallel, scenes, tpk, fastpath

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
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
    b = bench.make_inputs(i, n=n)
    pool.append({k: torch.as_tensor(getattr(b, k)).to(dev) for k in bench.HOST_KEYS})


class View:
    def __init__(self, d):
        self.__dict__.update(d)

    def __getitem__(self, k):
        return self.__dict__[k]


def phases(d, sync):
    out = {}
    def mark(name, t0):
        if sync:
            torch.cuda.synchronize()
        out[name] = (time.perf_counter() - t0) * 1e3
        return time.perf_counter()
    t = time.perf_counter()
    model.set_input(View(d), dev)
    t = mark("set_input", t)
    feats = model.Backbone(model.input).x
    t = mark("backbone_fwd", t)
    sem = model.Semantic(feats); off = model.Offset(feats)
    model.output = panoptic.PanopticResults(semantic_logits=sem, offset_logits=off, embed_logits=None, clusters=None,
                                            cluster_scores=None, mask_scores=None, cluster_type=None)
    t = mark("heads_fwd", t)
    model._zero_grad_hook()
    model._compute_loss(1)
    t = mark("loss", t)
    model.loss.backward()
    me.join_side_stream()
    t = mark("backward", t)
    model._grad_hook()
    model._optimizer.step()
    t = mark("optimizer", t)
    tpk.region_grow(d["syn_shifted"], d["syn_pred"], d["batch"], ignore_labels=ignore, nsample=200, radius=1.5 * bench.GRID,
                    min_cluster_size=10)
    t = mark("region_grow", t)
    return out


res = {}
for sync in (True, False):
    for i in range(3):
        phases(pool[i % 3], sync)
    torch.cuda.synchronize()
    acc = {}
    reps = 6
    t0 = time.perf_counter()
    for i in range(reps):
        for k, v in phases(pool[i % 3], sync).items():
            acc[k] = acc.get(k, 0.0) + v / reps
    torch.cuda.synchronize()
    acc["TOTAL_wall_per_step"] = (time.perf_counter() - t0) * 1e3 / reps
    res["synced (device time per phase)" if sync else "async (host enqueue time per phase)"] = acc
# forward split: maps (pass 1) vs launches, from the executor's own marks
print(json.dumps(res, indent=1))
