"""Like flaky_hunt.py, but also fingerprints the integer structures of every run (coordinate maps, gather tables, pair
lists, occupancy orders) and lists EVERY parameter whose gradient deviates -- separates "an activation flipped" (all
layers before it move) from "one weight gradient is wrong" (only that tensor moves)."""
import os, sys, json, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from test_gpu_sparse import _batch, _scene
from panopticsegforlargescalepointcloud_b200 import backbone as bb, fastpath, me
dev = torch.device("cuda:0")
runs = int(sys.argv[1]) if len(sys.argv) > 1 else 12
torch.manual_seed(7)
net = bb.Minkowski("unet", input_nc=4, config=bb.paper_backbone_config(16)).to(dev)
net.train(False)
rng = np.random.default_rng(5)
coords = _scene(9, n=14000, extent=64)
x = rng.standard_normal((len(coords), 4)).astype(np.float32)
g = torch.from_numpy(rng.standard_normal((len(coords), 16)).astype(np.float32)).to(dev)
fastpath.ENABLED = os.environ.get("PGS_DBG_FAST") == "1"
names = [n for n, _ in net.named_parameters()]
mgrs = []
_init = me.CoordinateManager.__init__
def _rec(self, c):
    _init(self, c); mgrs.append(self)
me.CoordinateManager.__init__ = _rec


def h(t):
    return hashlib.md5(t.detach().cpu().numpy().tobytes()).hexdigest()[:10]


def fingerprint(m):
    f = {}
    for ts, cm in sorted(m.maps.items()):
        f["coords%d" % ts] = h(cm.coords)
    for key, km in sorted(m.kmaps.items()):
        f["nbr%s" % (key,)] = h(km.nbr)
        if km._pairs is not None:
            i, o, offs, _ = km._pairs
            P = int(offs[-1])
            f["pairs%s" % (key,)] = h(i[:P]) + h(o[:P]) + h(offs)
        if km._sorted is not None:
            f["sorted%s" % (key,)] = h(km._sorted[0]) + h(km._sorted[1])
    return f


res, fps = [], []
for r in range(runs):
    del mgrs[:]
    for p in net.parameters():
        p.grad = torch.zeros_like(p)
    xin = _batch(coords, x, dev); xin.x.requires_grad_(True)
    out = net(xin).x
    out.backward(g)
    torch.cuda.synchronize()
    res.append([out.detach().clone(), xin.x.grad.detach().clone()] + [p.grad.detach().clone() for p in net.parameters()])
    fps.append(fingerprint(mgrs[0]))
lab = ["out", "dx"] + names
for r in range(1, runs):
    fd = [k for k in fps[0] if fps[r].get(k) != fps[0][k]]
    d = []
    for i, (u, v) in enumerate(zip(res[r], res[0])):
        e = float((u - v).abs().max()) / max(float(v.abs().max()), 1e-6)
        if e > 1e-4:
            d.append((lab[i], round(e, 5)))
    print(json.dumps({"run": r, "struct_diff": fd, "n_tensors_diff": len(d), "tensors": d[:60]}))
