"""Step time of the C2 training step with B cylinders collated into one batch (BASELINE config C4's per-GPU view:
"NPM3D-shape batch=8 cylinders"), one GPU, inputs resident.  Shows how far the per-scene cost drops once the host's
per-launch overhead and the latency-bound coarse levels are shared by several scenes.

    python scripts/batch_scaling.py [out.json] [B ...]
"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from panopticsegforlargescalepointcloud_b200 import panoptic, parallel, scenes, tpk

dev = torch.device("cuda:0")
out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "batch_scaling.json")
Bs = [int(a) for a in sys.argv[2:]] or [1, 2, 4, 8]
torch.manual_seed(2022)
opt = panoptic.paper_options("urban", cluster_type=1, grid=bench.GRID, use_score_net=True, prepare_epoch=30, scorer=False)
model = panoptic.PointGroup(opt, "dummy", panoptic.DatasetProperties("urban"), None).to(dev)
model.instantiate_optimizers({}); model.train()
dp = parallel.DataParallelStep(model)
ignore = [-1] + list(scenes.stuff_classes("urban"))


def batch(seeds):
    ss = [scenes.make_scene("urban", bench.N_POINTS, bench.GRID, bench.RADIUS, seed=s) for s in seeds]
    b = scenes.collate(ss)
    sh, pr = [], []
    for s, seed in zip(ss, seeds):
        off, _, logits = scenes.synthetic_head_outputs(s, seed=seed)
        sh.append((s.pos + off).astype(np.float32)); pr.append(logits.argmax(1).astype(np.int64))
    b.syn_shifted, b.syn_pred = np.concatenate(sh), np.concatenate(pr)
    return {k: torch.as_tensor(getattr(b, k)).to(dev) for k in bench.HOST_KEYS}


class View:
    def __init__(self, d): self.__dict__.update(d)
    def __getitem__(self, k): return self.__dict__[k]


rows = []
for B in Bs:
    pool = [batch(list(range(i * B, (i + 1) * B))) for i in range(2)]
    def step(i):
        d = pool[i % 2]
        dp.step(View(d), epoch=1, step=i, batch_size=B)
        return tpk.region_grow(d["syn_shifted"], d["syn_pred"], d["batch"], ignore_labels=ignore, nsample=200,
                               radius=1.5 * bench.GRID, min_cluster_size=10)
    for i in range(8): step(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    K = 8
    for i in range(K): step(i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    r = {"scenes_per_step": B, "rows_per_step": B * bench.N_POINTS, "ms_per_step": round(ms, 2), "scenes_per_s": round(1000.0 * B / ms, 1),
         "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 2)}
    print(json.dumps(r), flush=True)
    rows.append(r)
    del pool
    torch.cuda.empty_cache()
json.dump(rows, open(out_path, "w"), indent=1)
