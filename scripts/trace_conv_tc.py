import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from panopticsegforlargescalepointcloud_b200 import me, _lib
from test_gpu_sparse import _scene
dev = torch.device("cuda:0")
cin, cout, npts = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
coords = _scene(2, n=npts, batch=1, extent=max(8, int(40 * (npts / 20000) ** 0.5)))
mgr = me.CoordinateManager(torch.from_numpy(coords).to(dev))
km = mgr.kernel_map(1, 1, 1, 1, 3)
n = km.n_q
rng = np.random.default_rng(0)
X = torch.from_numpy(rng.standard_normal((n, cin)).astype(np.float32)).to(dev)
W = torch.from_numpy((rng.standard_normal((27, cin, cout)) / np.sqrt(27 * cin)).astype(np.float32)).to(dev)
me.CONV_IMPL = "tc"
for _ in range(3):
    me._conv_fwd_raw(X, W, km.nbr, n, 0, 0)
torch.cuda.synchronize()
lib = _lib.load()
buf = (ctypes.c_longlong * 4096)()
lib._handle if False else None
raw = ctypes.CDLL(_lib.lib_path())
raw.pgs_debug_trace(buf)
t = np.array(buf[:320]).reshape(40, 8)
m = np.array(buf[2048:2048 + 320]).reshape(40, 8)
print("n rows", n, "cin", cin, "cout", cout)
print("producer tid0: [wait, cvt+tmem st, B store, wait::st, fences+arrive, refill] period | mma lane: [bar wait, issue+commit] period")
for i in range(3, 13):
    r, q = t[i], m[i]
    print(i, [int(r[j + 1] - r[j]) for j in range(6)], int(t[i + 1][0] - r[0]), "|", [int(q[1] - q[0]), int(q[2] - q[1])], int(m[i + 1][0] - q[0]))
