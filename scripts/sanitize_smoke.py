"""One small invocation of every kernel family of libpgs_b200.so, for compute-sanitizer:

    compute-sanitizer --tool memcheck  python scripts/sanitize_smoke.py
    compute-sanitizer --tool racecheck python scripts/sanitize_smoke.py      (shared-memory hazards: mbarrier / TMEM kernels)
    compute-sanitizer --tool synccheck python scripts/sanitize_smoke.py

Sizes are tiny (the tools slow kernels down 10-100x); every conv kernel variant (tcgen05 + TMA, register-operand mma, queued
mma, few-row split, FFMA), both executor directions, BatchNorm, region growing, nearest neighbour, HDBSCAN, mean shift,
proposal IoU / NMS, block merging."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from panopticsegforlargescalepointcloud_b200 import me, tpk, hdbscan, meanshift, merging, scenes, panoptic, backbone as bb

dev = torch.device("cuda:0")
torch.manual_seed(0)
s = scenes.make_scene("urban", 6000, 0.2, 4.0, seed=1)
coords = torch.from_numpy(np.concatenate([np.zeros((len(s.coords), 1), np.int32), s.coords], 1)).to(dev)
mgr = me.CoordinateManager(coords)
mgr.stride(1, 2)
n = coords.shape[0]
for impl, cin, cout in (("tc", 16, 16), ("tc", 64, 96), ("tc", 32, 192), ("mma", 16, 16), ("mma", 32, 32), ("mma", 48, 48),
                        ("split", 64, 64), ("ffma", 4, 16)):
    me.CONV_IMPL = impl
    me.MMA_MIN_ROWS, me.SORT_MIN_ROWS = 0, 0
    for stride, T in ((1, False), (2, False), (1, True)):
        cls = me.MinkowskiConvolutionTranspose if T else me.MinkowskiConvolution
        conv = cls(cin, cout, kernel_size=3, stride=stride, dimension=3).to(dev)
        x = torch.randn(n, cin, device=dev, requires_grad=True)
        y = conv(me.SparseTensor(x, coordinate_manager=mgr, tensor_stride=1)).F
        y.sum().backward()
    print("conv", impl, cin, cout, "ok", flush=True)
me.CONV_IMPL = "auto"
# whole network through the native executor (forward + backward), two-level and a 3-level U-Net
for cfg in (bb.two_level_config(16),):
    net = bb.Minkowski("unet", input_nc=4, config=cfg).to(dev)

    class D:
        pass
    d = D()
    d.batch = torch.zeros(n, dtype=torch.int64, device=dev)
    d.coords = torch.from_numpy(s.coords).to(dev)
    d.x = torch.from_numpy(s.x).to(dev)
    d.pos = torch.from_numpy(s.pos).to(dev)
    net(d).x.sum().backward()
print("executor ok", flush=True)
off, emb, logits = scenes.synthetic_head_outputs(s, seed=1)
pos = torch.from_numpy((s.pos + off).astype(np.float32)).to(dev)
pred = torch.from_numpy(logits.argmax(1)).to(dev)
batch = torch.zeros(n, dtype=torch.long, device=dev)
cl = tpk.region_grow(pos, pred, batch, ignore_labels=[-1, 0, 1, 5], nsample=200, radius=0.3, min_cluster_size=10)
idx, d2 = tpk.ball_query(0.3, 16, pos, pos[:500], mode="PARTIAL_DENSE", batch_x=batch, batch_y=batch[:500])
print("region_grow", len(cl), "ball_query ok", flush=True)
if cl:
    il = torch.from_numpy(s.instance_labels).to(dev)
    tpk.instance_iou(cl, il, batch)
    tpk.proposal_nms(cl, torch.rand(len(cl), device=dev), 0.3)
    print("proposals ok", flush=True)
X = torch.from_numpy(emb[s.instance_mask][:3000]).to(dev)
hdbscan.HDBSCAN(15, 5, 0.006).fit_predict(X)
meanshift.MeanShift(bandwidth=0.6, bin_seeding=True).fit(X)
print("hdbscan / meanshift ok", flush=True)
p = torch.from_numpy(s.pos).to(dev)
merging.nearest(p[::3], p)
state = torch.full((n,), -1, dtype=torch.long, device=dev)
ids = torch.arange(n, device=dev)
lab = torch.from_numpy(s.instance_labels.astype(np.int64) - 1).to(dev)
state, mx = merging.block_merging(p, ids[: n // 2], ids[: n // 2: 2], lab[: n // 2: 2], state, 0)
state, mx = merging.block_merging(p, ids[n // 4:], ids[n // 4:: 2], lab[n // 4:: 2], state, mx)
merging.back_project(p, state, torch.from_numpy(s.y).to(dev), [0, 1, 5])
torch.cuda.synchronize()
print("merging ok; all kernel families ran", flush=True)
