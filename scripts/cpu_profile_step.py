"""cProfile of the host side of bench steps (where does the CPU time go?)."""
import cProfile, pstats, sys, os, io, time
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
b = bench.make_inputs(0)
d = {k: torch.as_tensor(getattr(b, k)).to(dev) for k in bench.HOST_KEYS}
class View:
    def __init__(self, d): self.__dict__.update(d)
    def __getitem__(self, k): return self.__dict__[k]
ignore = [-1] + list(scenes.stuff_classes("urban"))
def step():
    dp.step(View(d), epoch=1, step=0, batch_size=1)
    return tpk.region_grow(d["syn_shifted"], d["syn_pred"], d["batch"], ignore_labels=ignore, nsample=200, radius=1.5 * bench.GRID, min_cluster_size=10)
for _ in range(3): step()
torch.cuda.synchronize()
t = time.time()
for _ in range(5): step()
t_nosync = (time.time() - t) / 5
torch.cuda.synchronize()
t_all = (time.time() - t) / 5
print("host enqueue time per step %.1f ms, wall incl. drain %.1f ms" % (t_nosync * 1e3, t_all * 1e3))
pr = cProfile.Profile(); pr.enable()
for _ in range(3): step()
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28); print(s.getvalue()[:6000])
