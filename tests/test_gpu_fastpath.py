"""Fused U-Net executor (fastpath.py) against the per-layer module path it replaces and against the CPU restatement.

Same model object, same inputs: outputs, input gradient, every parameter gradient and the BatchNorm running
statistics must agree (same kernels underneath; the only differences are the summation order of the atomics in the
weight-gradient / few-row kernels and of gradient merges)."""
import numpy as np
import pytest
import torch

from oracle import cpu_path
from test_gpu_sparse import _batch, _scene, TOL

pytestmark = pytest.mark.gpu


def _run(net, coords, x, g, dev, direct_grads):
    for p in net.parameters():
        p.grad = torch.zeros_like(p) if direct_grads else None
    xin = _batch(coords, x, dev)
    xin.x.requires_grad_(True)
    out = net(xin).x
    out.backward(g)
    grads = {n: p.grad.detach().clone() for n, p in net.named_parameters()}
    return out.detach().clone(), xin.x.grad.detach().clone(), grads


@pytest.mark.parametrize("which,training,direct", [("two_level", True, True), ("two_level", False, False),
                                                   ("paper", True, True), ("paper", True, False), ("paper", False, True)])
def test_fastpath_matches_module_path(cuda_device, monkeypatch, which, training, direct):
    from panopticsegforlargescalepointcloud_b200 import backbone as bb, fastpath
    torch.manual_seed(7)
    cfg = bb.two_level_config(16) if which == "two_level" else bb.paper_backbone_config(16)
    net = bb.Minkowski("unet", input_nc=4, config=cfg).to(cuda_device)
    net.train(training)
    assert fastpath.program_for(net) is not None
    rng = np.random.default_rng(5)
    coords = _scene(9, n=14000 if which == "paper" else 7000, extent=64)
    x = rng.standard_normal((len(coords), 4)).astype(np.float32)
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.uniform_(-0.1, 0.1)
                m.running_var.uniform_(0.8, 1.2)
    sd0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
    n_out = cfg["up_conv"]["up_conv_nn"][-1][-1]
    g = torch.from_numpy(rng.standard_normal((len(coords), 16)).astype(np.float32)).to(cuda_device)

    monkeypatch.setattr(fastpath, "ENABLED", False)
    out_m, dx_m, gr_m = _run(net, coords, x, g, cuda_device, direct)
    sd_m = {k: v.detach().clone() for k, v in net.state_dict().items()}
    net.load_state_dict(sd0)
    monkeypatch.setattr(fastpath, "ENABLED", True)
    calls = []
    orig = fastpath._UNetFn.apply
    monkeypatch.setattr(fastpath._UNetFn, "apply", lambda *a: (calls.append(1), orig(*a))[1])
    out_f, dx_f, gr_f = _run(net, coords, x, g, cuda_device, direct)
    assert calls, "the fused executor did not run"
    sd_f = net.state_dict()

    def close(a, b, tol):
        scale = max(float(b.abs().max()), 1e-6)
        return float((a - b).abs().max()) <= tol * scale

    def cos(a, b):
        a, b = a.double().reshape(-1), b.double().reshape(-1)
        return float(torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-300))

    assert close(out_f, out_m, 2e-5)
    if training and which == "paper":
        # batch statistics over the 50-row coarsest levels amplify the summation-order noise of the atomics (and flip
        # ReLU masks of elements within rounding of zero): same bound as the module-path test against the oracle
        assert cos(dx_f, dx_m) >= 0.9999
        ga = torch.cat([gr_f[k].reshape(-1) for k in gr_m])
        gb = torch.cat([gr_m[k].reshape(-1) for k in gr_m])
        assert cos(ga, gb) >= 0.9999
        for name in gr_m:
            assert cos(gr_f[name], gr_m[name]) >= 0.995, name
    else:
        # eval mode has no batch statistics, but the few-row layers still sum by atomicAdd, and among the 10^6..10^7
        # activations of the decoder a handful sit within that rounding of zero: their ReLU masks differ between two
        # runs of the SAME path, and each flip moves the gradients of the layers before it by up to ~3 % of one
        # tensor's scale (scripts/flaky_hunt2.py / flaky_hunt3.py: first divergence is always one element of one
        # BatchNorm backward; frequent on the 7-level net, rare on the 2-level one).  Wrong layouts, masks or stale
        # weights give O(1) differences; value-level precision is pinned by the per-layer fp64 tests.
        assert close(dx_f, dx_m, 5e-3)
        ga = torch.cat([gr_f[k].reshape(-1) for k in gr_m])
        gb = torch.cat([gr_m[k].reshape(-1) for k in gr_m])
        assert 1.0 - cos(ga, gb) <= 1e-5
        for name in gr_m:
            assert close(gr_f[name], gr_m[name], 6e-2) and cos(gr_f[name], gr_m[name]) >= 0.999, name
    for k in sd_m:
        if "running" in k:
            assert close(sd_f[k], sd_m[k], 1e-5), k
        if "num_batches_tracked" in k:
            assert int(sd_f[k]) == int(sd_m[k])

    # and against the functional CPU restatement (forward, 1e-4: north_star)
    net.load_state_dict(sd0)
    with torch.no_grad():
        out_e = net(_batch(coords, x, cuda_device)).x
    ref = cpu_path.unet_forward({k: v.cpu() for k, v in sd0.items()}, cpu_path.resolve_cfg(cfg, 4), torch.from_numpy(x),
                                coords, training=training)
    assert float((out_e.cpu() - ref).abs().max()) <= TOL * max(float(ref.abs().max()), 1.0)


def test_fastpath_falls_back_on_foreign_modules(cuda_device):
    """A module tree the tape compiler does not know runs through the module path (still CUDA kernels)."""
    from panopticsegforlargescalepointcloud_b200 import backbone as bb, fastpath, me
    net = bb.Minkowski("unet", input_nc=4, config=bb.two_level_config(16)).to(cuda_device)
    net.down_modules[0].conv_in[2] = me.MinkowskiLeakyReLU(0.1)
    assert fastpath.program_for(net) is None
    coords = _scene(1, n=3000)
    x = np.random.default_rng(0).standard_normal((len(coords), 4)).astype(np.float32)
    out = net(_batch(coords, x, cuda_device)).x
    assert out.shape == (len(coords), 16) and bool(torch.isfinite(out).all())


def test_backward_after_another_forward_rearranges_its_weights(cuda_device):
    """Round-1 advisor finding: the arranged-weight buffer belongs to the Program and every forward rewrites it.  A second
    forward between a forward and its backward -- another batch whose level sizes cross the kernel-dispatch thresholds, run
    under no_grad so that no backward layouts are written at all -- must not change the first graph's gradients."""
    import numpy as np
    import torch
    from panopticsegforlargescalepointcloud_b200 import backbone as bb, scenes

    def batch(n, radius, seed):
        s = scenes.make_scene("urban", n, 0.2, radius, seed=seed)

        class D:
            pass
        d = D()
        d.batch = torch.zeros(len(s.pos), dtype=torch.int64, device=cuda_device)
        d.coords = torch.from_numpy(s.coords).to(cuda_device)
        d.x = torch.from_numpy(s.x).to(cuda_device)
        d.pos = torch.from_numpy(s.pos).to(cuda_device)
        return d

    torch.manual_seed(5)
    net = bb.Minkowski("unet", input_nc=4, config=bb.paper_backbone_config(16)).to(cuda_device)
    net.eval()
    a, b = batch(12000, 6.0, 1), batch(1500, 2.0, 2)      # 12 k rows: mma / tc kernels; 1.5 k rows: split kernels
    g = None

    def grads(interleave):
        nonlocal g
        net.zero_grad(set_to_none=True)
        out = net(a).x
        if g is None:
            g = torch.randn_like(out)
        if interleave:
            with torch.no_grad():
                net(b)
        out.backward(g)
        return [p.grad.clone() for p in net.parameters()], out.detach().clone()

    clean, out0 = grads(False)
    mixed, out1 = grads(True)
    assert float((out0 - out1).abs().max()) <= 2e-5 * max(1.0, float(out0.abs().max()))   # (few-row layers sum by atomicAdd)
    # run-to-run differences come from the order of fp32 atomic adds alone, amplified by the 3..50-row coarse levels of this
    # small scene (two clean runs differ by up to 4e-3 on single tensors); stale or foreign weight layouts give O(1)
    worst = max(float((x - y).norm() / y.norm().clamp_min(1e-12)) for x, y in zip(mixed, clean))
    ga = torch.cat([x.reshape(-1) for x in mixed]).double()
    gb = torch.cat([x.reshape(-1) for x in clean]).double()
    assert worst <= 2e-2 and 1.0 - float(torch.dot(ga, gb) / (ga.norm() * gb.norm())) <= 1e-6, worst



def test_prefetched_coordinate_maps_are_adopted(cuda_device):
    """backbone.prefetch_maps on a second stream: the forward pass adopts the prebuilt manager (no map is rebuilt) and
    gives the same output and gradients as the inline build."""
    from panopticsegforlargescalepointcloud_b200 import backbone as bb, me
    torch.manual_seed(3)
    net = bb.Minkowski("unet", input_nc=4, config=bb.paper_backbone_config(16)).to(cuda_device)
    net.train(False)
    rng = np.random.default_rng(12)
    coords = _scene(4, n=9000, extent=48)
    x = rng.standard_normal((len(coords), 4)).astype(np.float32)
    g = torch.from_numpy(rng.standard_normal((len(coords), 16)).astype(np.float32)).to(cuda_device)
    out0, dx0, g0 = _run(net, coords, x, g, cuda_device, False)

    side = torch.cuda.Stream(device=cuda_device)
    xin = _batch(coords, x, cuda_device)
    torch.cuda.synchronize()
    pm = net.prefetch_maps(xin.batch, xin.coords, stream=side)
    assert sorted(pm.manager.maps) == [1, 2, 4, 8, 16, 32, 64] and pm.n == len(coords)
    xin.coordinate_manager = pm
    built = []
    orig = me.build_coordinate_map
    try:
        me.build_coordinate_map = lambda *a, **k: (built.append(1), orig(*a, **k))[1]
        for p in net.parameters():
            p.grad = None
        xin.x.requires_grad_(True)
        out1 = net(xin).x
        out1.backward(g)
    finally:
        me.build_coordinate_map = orig
    assert not built, "the forward pass rebuilt a coordinate map although a prebuilt manager was attached"
    assert net.input.coordinate_manager is pm.manager

    def close(a, b, tol):
        return float((a - b).abs().max()) <= tol * max(float(b.abs().max()), 1e-6)

    assert close(out1.detach(), out0, 2e-5) and close(xin.x.grad, dx0, 5e-3)
