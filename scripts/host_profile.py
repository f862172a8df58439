"""Host-side cost of one training step, phase by phase, on a SMALL scene (GPU work is negligible, so each phase's
wall time between synchronisations is the host's enqueue cost), followed by a cumulative cProfile.

    python scripts/host_profile.py [n_points]
"""
import cProfile, pstats, sys, os, io, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from panopticsegforlargescalepointcloud_b200 import panoptic, parallel, scenes, tpk, me as _me
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
torch.manual_seed(2022)
opt = panoptic.paper_options("urban", cluster_type=1, grid=bench.GRID, use_score_net=True, prepare_epoch=30, scorer=False)
model = panoptic.PointGroup(opt, "dummy", panoptic.DatasetProperties("urban"), None).to(dev)
model.instantiate_optimizers({}); model.train()
dp = parallel.DataParallelStep(model)
b = bench.make_inputs(0, n=n)
d = {k: torch.as_tensor(getattr(b, k)).to(dev) for k in bench.HOST_KEYS}
class View:
    def __init__(self, d): self.__dict__.update(d)
    def __getitem__(self, k): return self.__dict__[k]
ignore = [-1] + list(scenes.stuff_classes("urban"))
phases = {}
def tick(name, t0):
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    phases[name] = phases.get(name, 0.0) + (t1 - t0)
    return t1
def step(timed=False):
    m = model
    t = time.perf_counter()
    m.set_input(View(d), m.device)
    if timed: t = tick("set_input", t)
    m.forward(epoch=1, step=0, is_training=True)
    if timed: t = tick("forward(+loss)", t)
    m._optimizer.zero_grad(set_to_none=False)
    if timed: t = tick("zero_grad", t)
    m.backward(1)
    if timed: t = tick("backward", t)
    _me.join_side_stream()
    if m._grad_hook is not None:
        m._grad_hook()
    if timed: t = tick("grad_hook", t)
    m._optimizer.step()
    if timed: t = tick("optimizer.step", t)
    c = tpk.region_grow(d["syn_shifted"], d["syn_pred"], d["batch"], ignore_labels=ignore, nsample=200, radius=1.5 * bench.GRID, min_cluster_size=10)
    if timed: t = tick("region_grow", t)
    return c
for _ in range(3): step()
torch.cuda.synchronize()
R = 5
for _ in range(R): step(True)
print("n =", n, " phase wall times per step (ms), sync after each phase:")
for k, v in phases.items():
    print("   %-16s %7.2f" % (k, 1e3 * v / R))
print("   %-16s %7.2f" % ("total", 1e3 * sum(phases.values()) / R))
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(R): step()
torch.cuda.synchronize()
print("untimed loop: %.2f ms / step" % (1e3 * (time.perf_counter() - t) / R))
pr = cProfile.Profile(); pr.enable()
for _ in range(3): step()
torch.cuda.synchronize(); pr.disable()
for key in ("cumulative", "tottime"):
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats(key).print_stats(45); print(s.getvalue()[:9000])
