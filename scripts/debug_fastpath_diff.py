"""Fused executor vs per-layer module path on the paper backbone (eval mode): per-parameter gradient difference, and the
run-to-run difference of each path with itself.  Switches are read from the environment (PGS_BN_MASK, PGS_TC_CORR ...).
    python scripts/debug_fastpath_diff.py"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from test_gpu_sparse import _batch, _scene
from panopticsegforlargescalepointcloud_b200 import backbone as bb, fastpath
dev = torch.device("cuda:0")
torch.manual_seed(7)
net = bb.Minkowski("unet", input_nc=4, config=bb.paper_backbone_config(16)).to(dev)
net.train(False)
rng = np.random.default_rng(5)
coords = _scene(9, n=int(os.environ.get("PGS_DBG_N", "14000")), extent=64)
x = rng.standard_normal((len(coords), 4)).astype(np.float32)
with torch.no_grad():
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.uniform_(-0.1, 0.1); m.running_var.uniform_(0.8, 1.2)
g = torch.from_numpy(rng.standard_normal((len(coords), 16)).astype(np.float32)).to(dev)


def run(fast):
    fastpath.ENABLED = fast
    for p in net.parameters():
        p.grad = torch.zeros_like(p)
    xin = _batch(coords, x, dev); xin.x.requires_grad_(True)
    out = net(xin).x
    out.backward(g)
    return out.detach().clone(), {n: p.grad.detach().clone() for n, p in net.named_parameters()}


def worst(ga, gb):
    rows = []
    for k in ga:
        sc = max(float(gb[k].abs().max()), 1e-6)
        rows.append((float((ga[k] - gb[k]).abs().max()) / sc, k))
    rows.sort(reverse=True)
    return [(round(r, 7), k) for r, k in rows[:4]]


if os.environ.get("PGS_DBG_ONLY"):       # for compute-sanitizer: one pass of the chosen path ("m" / "f"), no comparison
    run(os.environ["PGS_DBG_ONLY"] == "f"); torch.cuda.synchronize(); print("done"); sys.exit(0)
om, gm = run(False); om2, gm2 = run(False); of, gf = run(True); of2, gf2 = run(True)
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("PGS_")},
                  "out_fast_vs_module": float((of - om).abs().max() / om.abs().max()),
                  "module_vs_module": worst(gm2, gm), "fast_vs_fast": worst(gf2, gf), "fast_vs_module": worst(gf, gm)}))
