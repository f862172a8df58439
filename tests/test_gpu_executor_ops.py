"""C-ABI pieces added for the fused executor, each against a plain torch statement of the same op:
pgs_add2, pgs_cat2 (+ its split), the _ex BatchNorm flags, pgs_conv_prep_weights_batch + W == NULL launches,
and the tensor-core weight gradient against the FFMA kernel / fp64 torch."""
import numpy as np
import pytest
import torch

from test_gpu_sparse import _scene

pytestmark = pytest.mark.gpu


def _lib():
    from panopticsegforlargescalepointcloud_b200 import _lib
    return _lib


@pytest.mark.parametrize("n,ca,cb", [(1, 4, 4), (777, 16, 48), (5000, 96, 32)])
def test_add2_cat2(cuda_device, n, ca, cb):
    L = _lib(); lib = L.load(); P, S = L.ptr, L.stream_ptr
    g = torch.Generator(device="cpu").manual_seed(n)
    a = torch.randn(n, ca, generator=g).to(cuda_device); b = torch.randn(n, cb, generator=g).to(cuda_device)
    a2 = torch.randn(n, ca, generator=g).to(cuda_device)
    y = torch.empty_like(a)
    L.check(lib.pgs_add2(P(a), P(a2), P(y), n * ca, S()))
    assert torch.equal(y, a + a2)
    c = torch.empty(n, ca + cb, device=cuda_device)
    L.check(lib.pgs_cat2(P(a), ca, P(b), cb, P(c), n, 0, S()))
    assert torch.equal(c, torch.cat([a, b], 1))
    ra, rb = torch.empty_like(a), torch.empty_like(b)
    L.check(lib.pgs_cat2(P(ra), ca, P(rb), cb, P(c), n, 1, S()))
    assert torch.equal(ra, a) and torch.equal(rb, b)
    assert lib.pgs_add2(P(a), P(a2), P(y), 3, S()) != 0      # not a multiple of 4: rejected loudly


def test_bn_ex_flags(cuda_device):
    """PGS_BN_SUMS_ZEROED uses caller-zeroed sums; PGS_BN_ACCUMULATE_PARAM_GRADS adds into dweight / dbias."""
    L = _lib(); lib = L.load(); P, S = L.ptr, L.stream_ptr
    n, C = 3001, 48
    g = torch.Generator(device="cpu").manual_seed(1)
    X = torch.randn(n, C, generator=g).to(cuda_device); dY = torch.randn(n, C, generator=g).to(cuda_device)
    w = (torch.rand(C, generator=g) + 0.5).to(cuda_device); b = torch.randn(C, generator=g).to(cuda_device)

    def run(flags, pre, mask_from_x=False):
        rm, rv = torch.zeros(C, device=cuda_device), torch.ones(C, device=cuda_device)
        Y, dX = torch.empty_like(X), torch.empty_like(X)
        st = torch.empty(2, C, device=cuda_device)
        sums = torch.zeros(2 * C, dtype=torch.float64, device=cuda_device) if flags & 2 else \
            torch.full((2 * C,), 7.0, dtype=torch.float64, device=cuda_device)
        L.check(lib.pgs_bn_forward_ex(P(X), n, C, P(w), P(b), P(rm), P(rv), 1, 0.1, 1e-5, 1, flags & 2, P(sums), P(st[0]),
                                      P(st[1]), P(Y), S()))
        if flags & 2:
            sums.zero_()
        dwb = torch.full((2, C), pre, device=cuda_device)
        L.check(lib.pgs_bn_backward_ex(P(X), None if mask_from_x else P(Y), P(dY), n, C, P(w), P(b), P(st[0]), P(st[1]), 1,
                                       1, flags, P(sums), P(dX), P(dwb[0]), P(dwb[1]), S()))
        return Y, dX, dwb, rm, rv

    Y0, dX0, g0, rm0, rv0 = run(0, 123.0)          # plain: memsets inside, gradients overwritten
    Y1, dX1, g1, rm1, rv1 = run(3, 2.5)            # caller-zeroed sums, accumulate onto 2.5
    assert torch.equal(Y0, Y1) and torch.equal(rm0, rm1) and torch.equal(rv0, rv1)
    assert torch.allclose(dX0, dX1, rtol=0, atol=1e-6)
    assert torch.allclose(g1, g0 + 2.5, rtol=1e-6, atol=1e-5)
    # Y == NULL: the ReLU mask is recomputed from x with the forward's roundings -> the same mask, bit for bit
    Y2, dX2, g2, _, _ = run(0, 123.0, mask_from_x=True)
    # (a flipped mask bit would change dX by ~|dY| w invstd; 1e-6 only allows for the order of the fp64 atomics)
    assert torch.allclose(dX2, dX0, rtol=0, atol=1e-6) and torch.allclose(g2, g0, rtol=1e-6, atol=1e-5)
    Xr = X.clone().requires_grad_(True); wr = w.clone().requires_grad_(True); br = b.clone().requires_grad_(True)
    ref = torch.relu(torch.nn.functional.batch_norm(Xr, None, None, wr, br, True, 0.1, 1e-5))
    ref.backward(dY)
    assert torch.allclose(Y0, ref.detach(), atol=1e-5) and torch.allclose(dX0, Xr.grad, atol=1e-5)
    assert torch.allclose(g0[0], wr.grad, rtol=1e-4, atol=1e-3) and torch.allclose(g0[1], br.grad, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("cin,cout,impl", [(16, 16, "mma"), (32, 48, "mma"), (64, 64, "tc"), (96, 112, "tc"), (80, 80, "split")])
def test_batched_weight_prep_matches_inline_prep(cuda_device, monkeypatch, cin, cout, impl):
    """pgs_conv_prep_weights_batch + W == NULL gives bit-identical output to the entry point arranging W itself,
    forward (W) and input-gradient (W^T, mirrored table) form."""
    from panopticsegforlargescalepointcloud_b200 import me
    L = _lib(); lib = L.load()
    monkeypatch.setattr(me, "CONV_IMPL", impl)
    monkeypatch.setattr(me, "SORT_MIN_ROWS", 0)
    coords = _scene(5, n=9000)
    mgr = me.CoordinateManager(torch.from_numpy(coords).to(cuda_device))
    km = mgr.kernel_map(1, 1, 1, 1, 3)
    n = km.n_q
    rng = np.random.default_rng(cin + cout)
    W = torch.from_numpy((rng.standard_normal((27, cin, cout)) / 20).astype(np.float32)).to(cuda_device)
    for wt, (ci, co) in ((0, (cin, cout)), (1, (cout, cin))):
        X = torch.from_numpy(rng.standard_normal((n, ci)).astype(np.float32)).to(cuda_device)
        nb = me._conv_scratch_bytes(lib, 27, ci, co)
        outs = []
        for prepped in (False, True):
            scratch = torch.zeros(nb, dtype=torch.uint8, device=cuda_device)
            if prepped:
                desc = torch.tensor([[W.data_ptr(), scratch.data_ptr(), 27, ci, co, wt, 0 if impl == "tc" else 1, 0]],
                                    dtype=torch.int64).to(cuda_device)
                L.check(lib.pgs_conv_prep_weights_batch(L.ptr(desc), 1, 27 * ci * co, L.stream_ptr()))
            Y = torch.empty(n, co, device=cuda_device)
            kind = me._conv_launch(lib, X.data_ptr(), n, W.data_ptr(), 27, ci, co, km, n, wt, wt, Y.data_ptr(),
                                   scratch.data_ptr(), nb, L.stream_ptr(), prepped=prepped)
            assert kind == impl
            outs.append(Y)
        assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("cin,cout", [(16, 16), (32, 32), (48, 80), (96, 32), (16, 64)])
def test_weight_gradient_mma_matches_fp64(cuda_device, cin, cout):
    from panopticsegforlargescalepointcloud_b200 import me
    L = _lib(); lib = L.load(); P, S = L.ptr, L.stream_ptr
    coords = _scene(6, n=7001)
    mgr = me.CoordinateManager(torch.from_numpy(coords).to(cuda_device))
    km = mgr.kernel_map(1, 1, 1, 1, 3)
    n = km.n_q
    g = torch.Generator(device="cpu").manual_seed(cin * 7 + cout)
    X = torch.randn(n, cin, generator=g).to(cuda_device); dY = torch.randn(n, cout, generator=g).to(cuda_device)
    in_idx, out_idx, offs, max_pairs = km.pairs()
    ref = torch.zeros(27, cin, cout, dtype=torch.float64, device=cuda_device)
    for k in range(27):
        idx = km.nbr[k].long(); m = idx >= 0
        ref[k] = X.double()[idx[m]].t() @ dY.double()[m]
    for mirror in (0, 1):
        dW = torch.zeros(27, cin, cout, device=cuda_device)
        L.check(lib.pgs_conv_bwd_weight_mma(P(X), P(dY), P(in_idx), P(out_idx), P(offs), max_pairs, 27, cin, cout, mirror,
                                            P(dW), S()))
        want = ref.flip(0) if mirror else ref
        assert float((dW.double() - want).abs().max()) <= 2e-5 * float(ref.abs().max())
    # K == 1 identity pairs (1x1 shortcut convolutions)
    dW1 = torch.zeros(1, cin, cout, device=cuda_device)
    L.check(lib.pgs_conv_bwd_weight_mma(P(X), P(dY), None, None, None, n, 1, cin, cout, 0, P(dW1), S()))
    r1 = X.double().t() @ dY.double()
    assert float((dW1[0].double() - r1).abs().max()) <= 2e-5 * float(r1.abs().max())
