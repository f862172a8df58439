"""Bring-up harness for the tcgen05 conv: compare against the FFMA kernel and the numpy oracle, time both."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from panopticsegforlargescalepointcloud_b200 import me, _lib
from test_gpu_sparse import _scene
from oracle import sparse_ref as sr

dev = torch.device("cuda:0")
shapes = [(16, 16), (32, 48), (64, 64), (192, 80), (96, 112), (160, 192)] if len(sys.argv) < 2 else [tuple(map(int, a.split("x"))) for a in sys.argv[1:]]
NPTS = int(os.environ.get("NPTS", "20000"))
coords = _scene(2, n=NPTS, extent=int(40 * (NPTS / 20000) ** 0.5))
mgr = me.CoordinateManager(torch.from_numpy(coords).to(dev))
km = mgr.kernel_map(1, 1, 1, 1, 3)
n = km.n_q
rng = np.random.default_rng(0)
for cin, cout in shapes:
    X = torch.from_numpy(rng.standard_normal((n, cin)).astype(np.float32)).to(dev)
    W = torch.from_numpy((rng.standard_normal((27, cin, cout)) / np.sqrt(27 * cin)).astype(np.float32)).to(dev)
    for mirror, wt in [(0, 0), (1, 1)]:
        Wk = W if not wt else W.transpose(1, 2).contiguous()      # [K][cout][cin] storage when w_transposed
        me.CONV_IMPL = "ffma"
        Yf = me._conv_fwd_raw(X, Wk, km.nbr, n, mirror, wt)
        me.CONV_IMPL = "tc"
        Yt = me._conv_fwd_raw(X, Wk, km.nbr, n, mirror, wt)
        torch.cuda.synchronize()
        err = float((Yf - Yt).abs().max()); scale = float(Yf.abs().max())
        print("cin %3d cout %3d mirror %d wT %d  max|ffma-tc| %.3e (scale %.2f)" % (cin, cout, mirror, wt, err, scale), flush=True)
    Yr = sr.conv_fwd(X.cpu().numpy(), W.cpu().numpy(), km.nbr.cpu().numpy())
    me.CONV_IMPL = "tc"
    Yt = me._conv_fwd_raw(X, W, km.nbr, n, 0, 0)
    print("   vs fp64 oracle: tc %.3e" % float(np.abs(Yt.cpu().numpy() - Yr).max()))
    for impl in ("ffma", "tc"):
        me.CONV_IMPL = impl
        for _ in range(3):
            me._conv_fwd_raw(X, W, km.nbr, n, 0, 0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            me._conv_fwd_raw(X, W, km.nbr, n, 0, 0)
        e1.record(); torch.cuda.synchronize()
        print("   %s: %.1f us / launch (n=%d)" % (impl, e0.elapsed_time(e1) * 100, n), flush=True)
