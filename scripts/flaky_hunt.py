"""Repeat the backbone's forward + backward (paper config, eval mode) on ONE input and report the runs whose gradients
deviate from the others by more than summation-order noise (2e-6): hunts rare races.  PGS_DBG_FAST=1: fused executor.
    python scripts/flaky_hunt.py [runs]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from test_gpu_sparse import _batch, _scene
from panopticsegforlargescalepointcloud_b200 import backbone as bb, fastpath
dev = torch.device("cuda:0")
runs = int(sys.argv[1]) if len(sys.argv) > 1 else 30
torch.manual_seed(7)
net = bb.Minkowski("unet", input_nc=4, config=bb.paper_backbone_config(16)).to(dev)
net.train(os.environ.get("PGS_DBG_TRAIN") == "1")
rng = np.random.default_rng(5)
coords = _scene(9, n=int(os.environ.get("PGS_DBG_N", "14000")), extent=64)
x = rng.standard_normal((len(coords), 4)).astype(np.float32)
g = torch.from_numpy(rng.standard_normal((len(coords), 16)).astype(np.float32)).to(dev)
fastpath.ENABLED = os.environ.get("PGS_DBG_FAST") == "1"
sd0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
names = [n for n, _ in net.named_parameters()]
res = []
for r in range(runs):
    net.load_state_dict(sd0)
    for p in net.parameters():
        p.grad = torch.zeros_like(p)
    xin = _batch(coords, x, dev); xin.x.requires_grad_(True)
    out = net(xin).x
    out.backward(g)
    torch.cuda.synchronize()
    res.append([out.detach().clone()] + [p.grad.detach().clone() for p in net.parameters()])


def diff(a, b):
    worst, where = 0.0, None
    for i, (u, v) in enumerate(zip(a, b)):
        d = float((u - v).abs().max()) / max(float(v.abs().max()), 1e-6)
        if d > worst:
            worst, where = d, ("out" if i == 0 else names[i - 1])
    return worst, where


# reference = the run that agrees with most others
thr = float(os.environ.get("PGS_DBG_THR", "1e-4"))
d0 = [diff(r, res[0]) for r in res]
ref = 0 if sum(d[0] > thr for d in d0) < runs / 2 else next(i for i, d in enumerate(d0) if d[0] > thr)
dr = [diff(r, res[ref]) for r in res]
bad = [(i, round(d[0], 6), d[1]) for i, d in enumerate(dr) if d[0] > thr]
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("PGS_")}, "runs": runs, "ref": ref,
                  "noise": round(max(d[0] for d in dr if d[0] <= thr), 8), "glitched": bad}))
