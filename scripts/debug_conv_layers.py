"""Per-conv check inside the paper U-Net backward: every dX / dW from the CUDA kernels vs a float64 torch
recomputation from the same saved inputs (debug aid)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from panopticsegforlargescalepointcloud_b200 import backbone as bb, me
from test_gpu_sparse import _scene, _batch

orig_bwd = me._SparseConvFn.backward
report = []

def ref64(X, W3, nbr, n_out, mirror, transposed_w):
    K = W3.shape[0]
    Xd, Wd = X.double(), W3.double()
    if transposed_w:
        Wd = Wd.transpose(1, 2)
    Y = torch.zeros(n_out, Wd.shape[2], dtype=torch.float64, device=X.device)
    if nbr is None:
        return Xd @ Wd[0]
    for k in range(K):
        tk = K - 1 - k if mirror else k
        idx = nbr[tk].long()
        m = idx >= 0
        Y[m] += Xd[idx[m]] @ Wd[k]
    return Y

def checked_bwd(ctx, dY):
    out = orig_bwd(ctx, dY)
    X, W = ctx.saved_tensors
    W3 = W.reshape(-1, W.shape[-2], W.shape[-1])
    K = W3.shape[0]
    dX, dW = out[0], out[1]
    nbr_b = ctx.km_b.nbr if ctx.km_b is not None else None
    nbr_f = ctx.km_f.nbr if ctx.km_f is not None else None
    if dX is not None:
        r = ref64(dY, W3, nbr_b, X.shape[0], ctx.mirror_b, True)
        ex = float((dX.double() - r).abs().max() / r.abs().max().clamp_min(1e-30))
    else:
        ex = -1
    # dW ref
    Xd, dYd = X.double(), dY.double()
    dWr = torch.zeros_like(W3, dtype=torch.float64)
    if nbr_f is None:
        dWr[0] = Xd.t() @ dYd
    else:
        for k in range(K):
            tk = K - 1 - k if ctx.mirror_f else k
            idx = nbr_f[tk].long(); m = idx >= 0
            dWr[k] = Xd[idx[m]].t() @ dYd[m]
    ew = float((dW.reshape(W3.shape).double() - dWr).abs().max() / dWr.abs().max().clamp_min(1e-30))
    report.append((max(ex, ew), ex, ew, tuple(W.shape), X.shape[0], dY.shape[0], ctx.mirror_f, ctx.mirror_b,
                   float(dY.abs().max()), float(X.abs().max())))
    return out

me._SparseConvFn.backward = staticmethod(checked_bwd)
dev = torch.device("cuda:0")
torch.manual_seed(2022)
cfg = bb.paper_backbone_config(16)
net = bb.Minkowski("unet", input_nc=4, config=cfg).to(dev)
net.eval()
rng = np.random.default_rng(11)
coords = _scene(4, n=15000, extent=64)
x = rng.standard_normal((len(coords), 4)).astype(np.float32)
out = net(_batch(coords, x, dev)).x
g = torch.from_numpy(rng.standard_normal(tuple(out.shape)).astype(np.float32))
out.backward(g.to(dev))
report.sort(reverse=True)
for r in report[:12]:
    print("worst %.2e dX %.2e dW %.2e W%s n_in %d n_out %d mir %s/%s |dY| %.2e |X| %.2e" % r)
print("convs checked", len(report))
