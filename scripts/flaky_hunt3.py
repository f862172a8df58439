"""Which op diverges first?  Records the output of every conv launch (forward and input-gradient) and every BatchNorm
forward / backward of the module path, run after run on one input, and prints the first ops whose result differs from
run 0 by more than 1e-5 of its scale."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from test_gpu_sparse import _batch, _scene
from panopticsegforlargescalepointcloud_b200 import backbone as bb, fastpath, me
dev = torch.device("cuda:0")
runs = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.manual_seed(7)
net = bb.Minkowski("unet", input_nc=4, config=bb.paper_backbone_config(16)).to(dev)
net.train(False)
rng = np.random.default_rng(5)
coords = _scene(9, n=14000, extent=64)
x = rng.standard_normal((len(coords), 4)).astype(np.float32)
g = torch.from_numpy(rng.standard_normal((len(coords), 16)).astype(np.float32)).to(dev)
fastpath.ENABLED = False
log = []
_raw = me._conv_fwd_raw
def conv(X, W3, km, n_q, mirror, wt):
    Y = _raw(X, W3, km, n_q, mirror, wt)
    K = W3.shape[0]
    kind = me._conv_kernel_choice(me._lib.load(), K, X.shape[1], Y.shape[1], n_q, km is not None)
    log.append(("conv%s %dx%d->%d K%d n%d %s" % ("T" if wt else "", X.shape[0], X.shape[1], Y.shape[1], K, n_q, kind), Y.detach().clone()))
    return Y
me._conv_fwd_raw = conv
_bf, _bb = me._BnFn.forward, me._BnFn.backward
def bnf(ctx, X, *a):
    Y = _bf(ctx, X, *a)
    log.append(("bn_fwd %dx%d relu%d" % (X.shape[0], X.shape[1], int(a[-1])), Y.detach().clone()))
    return Y
def bnb(ctx, dY):
    r = _bb(ctx, dY)
    log.append(("bn_bwd %dx%d" % tuple(dY.shape), r[0].detach().clone()))
    return r
me._BnFn.forward = staticmethod(bnf); me._BnFn.backward = staticmethod(bnb)
logs = []
for r in range(runs):
    log = []
    for p in net.parameters():
        p.grad = torch.zeros_like(p)
    xin = _batch(coords, x, dev); xin.x.requires_grad_(True)
    out = net(xin).x
    out.backward(g)
    torch.cuda.synchronize()
    logs.append(log)
print("ops per run", len(logs[0]))
for r in range(1, runs):
    bad = []
    for i, ((na, a), (nb_, b)) in enumerate(zip(logs[r], logs[0])):
        sc = max(float(b.abs().max()), 1e-12)
        d = (a - b).abs()
        e = float(d.max()) / sc
        if e > 1e-5:
            j = int(d.argmax())
            bad.append((i, na, "%.2e" % e, "n_elems_diff=%d" % int((d > 1e-5 * sc).sum()), "at=%d" % j, "vals=%.3e/%.3e" % (float(a.reshape(-1)[j]), float(b.reshape(-1)[j]))))
    print(json.dumps({"run": r, "n_bad": len(bad), "first": bad[:6]}))
