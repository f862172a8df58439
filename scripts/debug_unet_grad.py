"""Per-parameter gradient error of the paper backbone vs the CPU oracle (debug aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import cpu_path
from panopticsegforlargescalepointcloud_b200 import backbone as bb
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_sparse import _scene, _batch

dev = torch.device("cuda:0")
for n, extent, training in [(15000, 64, True), (15000, 64, False), (60000, 200, True)]:
    torch.manual_seed(2022)
    cfg = bb.paper_backbone_config(16)
    net = bb.Minkowski("unet", input_nc=4, config=cfg).to(dev)
    net.train(training)
    rng = np.random.default_rng(11)
    coords = _scene(4, n=n, extent=extent)
    x = rng.standard_normal((len(coords), 4)).astype(np.float32)
    sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    out = net(_batch(coords, x, dev)).x
    g = torch.from_numpy(rng.standard_normal(tuple(out.shape)).astype(np.float32))
    out.backward(g.to(dev))
    sd_ref = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()}
    ref = cpu_path.unet_forward(sd_ref, cpu_path.resolve_cfg(cfg, 4), torch.from_numpy(x), coords, training=training)
    ref.backward(g)
    print("n", len(coords), "train", training, "fwd err", float((out.detach().cpu() - ref.detach()).abs().max()))
    rows = []
    for name, p in net.named_parameters():
        gr = sd_ref[name].grad
        rows.append((float((p.grad.cpu() - gr).abs().max()) / max(float(gr.abs().max()), 1e-12), float(gr.abs().max()), name))
    rows.sort(reverse=True)
    for r in rows[:8]:
        print("  rel %.3e  max|g| %.3e  %s" % r)
    # fp64 CPU reference to see how ill-conditioned the problem is
    sd64 = {k: (v.double() if v.dtype.is_floating_point else v).clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()}
    ref64 = cpu_path.unet_forward(sd64, cpu_path.resolve_cfg(cfg, 4), torch.from_numpy(x).double(), coords, training=training)
    ref64.backward(g.double())
    w_gpu = max(float((p.grad.cpu().double() - sd64[n_].grad).abs().max()) / max(float(sd64[n_].grad.abs().max()), 1e-12) for n_, p in net.named_parameters())
    w_cpu = max(float((sd_ref[n_].grad.double() - sd64[n_].grad).abs().max()) / max(float(sd64[n_].grad.abs().max()), 1e-12) for n_, p in net.named_parameters())
    print("  vs fp64: gpu worst rel %.3e, cpu-fp32 worst rel %.3e" % (w_gpu, w_cpu))
