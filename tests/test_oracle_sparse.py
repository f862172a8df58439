"""Pins oracle/sparse_ref.py (the CPU restatement of the MinkowskiEngine semantics) against an independent
definition: dense torch.nn.functional.conv3d / conv_transpose3d on a densified toy grid (fp32, 1e-4), and a
literal python-dict rulebook.  CPU only."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import sparse_ref as sr

G = 8  # toy grid edge


def _scene(seed, batch=2, occ=0.35):
    rng = np.random.default_rng(seed)
    rows = []
    for b in range(batch):
        m = rng.random((G, G, G)) < occ
        z, y, x = np.nonzero(m)
        c = np.stack([np.full_like(x, b), x, y, z], 1)
        rows.append(c[rng.permutation(len(c))])
    return np.concatenate(rows).astype(np.int32)


def _densify(coords, X, size=G, div=1):
    B = int(coords[:, 0].max()) + 1
    d = torch.zeros(B, X.shape[1], size, size, size, dtype=torch.float64)
    c = torch.as_tensor(coords).long()
    d[c[:, 0], :, c[:, 3] // div, c[:, 2] // div, c[:, 1] // div] = torch.as_tensor(X, dtype=torch.float64)
    return d


def _sample(dense, coords, div=1):
    c = torch.as_tensor(coords).long()
    return dense[c[:, 0], :, c[:, 3] // div, c[:, 2] // div, c[:, 1] // div]


def _dense_w(W):  # [27,Cin,Cout] -> [Cout,Cin,kz,ky,kx]
    return torch.as_tensor(W, dtype=torch.float64).reshape(3, 3, 3, W.shape[1], W.shape[2]).permute(4, 3, 0, 1, 2)


def test_coordinate_map_first_occurrence_order():
    c = np.array([[0, 5, 5, 5], [0, 1, 1, 1], [0, 4, 5, 4], [1, 1, 1, 1], [0, 0, 0, 1], [0, -1, -2, -3]], np.int32)
    out, in2out = sr.coordinate_map(c, 2)
    assert out.tolist() == [[0, 4, 4, 4], [0, 0, 0, 0], [1, 0, 0, 0], [0, -2, -2, -4]]
    assert in2out.tolist() == [0, 1, 0, 2, 1, 3]
    out1, id1 = sr.coordinate_map(c, 1)
    assert np.array_equal(out1, c) and id1.tolist() == list(range(6))


@pytest.mark.parametrize("case", [(1, 1, 3), (2, 1, 3), (1, -1, 3), (1, 1, 2), (2, -1, 2)])
def test_kernel_map_matches_dict(case):
    step, sign, ks = case
    coords = _scene(3)
    q = sr.coordinate_map(coords, 2)[0] if step == 1 and sign == 1 and ks == 3 else coords
    a = sr.kernel_map(q, coords, step, sign, ks)
    b = sr.kernel_map_dict(q, coords, step, sign, ks)
    assert np.array_equal(a, b)


def test_rulebook_swap_symmetry():
    """(k,i,o) triples of the strided map equal those of the transposed map with in/out swapped (App. B.4)."""
    coords = _scene(5)
    coarse = sr.coordinate_map(coords, 2)[0]
    dn = sr.kernel_map(coarse, coords, 1, +1, 3)   # [k][o] -> i
    up = sr.kernel_map(coords, coarse, 1, -1, 3)   # [k][i] -> o
    a = {(k, int(dn[k, o]), o) for k in range(27) for o in range(dn.shape[1]) if dn[k, o] >= 0}
    b = {(k, i, int(up[k, i])) for k in range(27) for i in range(up.shape[1]) if up[k, i] >= 0}
    assert a == b and len(a) > 0
    i_idx, o_idx, offs = sr.pairs(dn)
    assert offs[-1] == len(a) and np.all(np.diff(offs) >= 0)


@pytest.mark.parametrize("cin,cout", [(4, 16), (16, 8)])
def test_conv_s1_matches_dense(cin, cout):
    rng = np.random.default_rng(0)
    coords = _scene(1)
    X = rng.standard_normal((len(coords), cin)).astype(np.float32)
    W = rng.standard_normal((27, cin, cout)).astype(np.float32) * 0.1
    nbr = sr.kernel_map(coords, coords, 1, 1, 3)
    Y = sr.conv_fwd(X, W, nbr)
    dense = F.conv3d(_densify(coords, X), _dense_w(W), padding=1)
    assert np.allclose(Y, _sample(dense, coords).numpy(), atol=1e-4, rtol=1e-4)
    # transposed stride-1 == mirrored offsets == dense conv_transpose3d
    Yt = sr.conv_fwd(X, W, nbr, mirror=True)
    wt = _dense_w(W).permute(1, 0, 2, 3, 4)  # [Cin,Cout,...]
    dense_t = F.conv_transpose3d(_densify(coords, X), wt, padding=1)
    assert np.allclose(Yt, _sample(dense_t, coords).numpy(), atol=1e-4, rtol=1e-4)


def test_conv_s2_and_transpose_match_dense():
    rng = np.random.default_rng(1)
    coords = _scene(2)
    cin, cout = 8, 12
    X = rng.standard_normal((len(coords), cin)).astype(np.float32)
    W = rng.standard_normal((27, cin, cout)).astype(np.float32) * 0.1
    coarse = sr.coordinate_map(coords, 2)[0]
    dn = sr.kernel_map(coarse, coords, 1, +1, 3)
    Y = sr.conv_fwd(X, W, dn)
    dense = F.conv3d(_densify(coords, X), _dense_w(W), stride=2, padding=1)
    assert np.allclose(Y, _sample(dense, coarse, div=2).numpy(), atol=1e-4, rtol=1e-4)
    # transposed: coarse -> fine on the encoder's fine map
    Xc = rng.standard_normal((len(coarse), cout)).astype(np.float32)
    Wt = rng.standard_normal((27, cout, cin)).astype(np.float32) * 0.1
    up = sr.kernel_map(coords, coarse, 1, -1, 3)
    Yf = sr.conv_fwd(Xc, Wt, up)
    wt = _dense_w(Wt).permute(1, 0, 2, 3, 4)
    dense_t = F.conv_transpose3d(_densify(coarse, Xc, size=G // 2, div=2), wt, stride=2, padding=1, output_padding=1)
    assert np.allclose(Yf, _sample(dense_t, coords).numpy(), atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize("stride", [1, 2])
def test_conv_backward_matches_dense_autograd(stride):
    rng = np.random.default_rng(7)
    coords = _scene(4)
    cin, cout = 6, 10
    X = rng.standard_normal((len(coords), cin)).astype(np.float32)
    W = rng.standard_normal((27, cin, cout)).astype(np.float32) * 0.1
    if stride == 1:
        out_c, nbr, div = coords, sr.kernel_map(coords, coords, 1, 1, 3), 1
    else:
        out_c = sr.coordinate_map(coords, 2)[0]
        nbr, div = sr.kernel_map(out_c, coords, 1, 1, 3), 2
    dY = rng.standard_normal((len(out_c), cout)).astype(np.float32)
    dX, dW = sr.conv_bwd(X, W, dY, nbr)
    Xt = torch.tensor(X, dtype=torch.float64, requires_grad=True)
    Wt = torch.tensor(W, dtype=torch.float64, requires_grad=True)
    B = int(coords[:, 0].max()) + 1
    c = torch.as_tensor(coords).long()
    d = torch.zeros(B, G, G, G, cin, dtype=torch.float64).index_put((c[:, 0], c[:, 3], c[:, 2], c[:, 1]), Xt)
    d = d.permute(0, 4, 1, 2, 3)
    wd = Wt.reshape(3, 3, 3, cin, cout).permute(4, 3, 0, 1, 2)
    out = F.conv3d(d, wd, stride=stride, padding=1)
    (_sample(out, out_c, div=div) * torch.as_tensor(dY, dtype=torch.float64)).sum().backward()
    assert np.allclose(dX, Xt.grad.numpy(), atol=1e-4, rtol=1e-4)
    assert np.allclose(dW, Wt.grad.numpy(), atol=1e-4, rtol=1e-4)


def test_k1_is_plain_matmul():
    rng = np.random.default_rng(2)
    X = rng.standard_normal((50, 6)).astype(np.float32)
    W = rng.standard_normal((6, 9)).astype(np.float32)
    assert np.allclose(sr.conv_fwd(X, W, None), X @ W, atol=1e-5)
    dY = rng.standard_normal((50, 9)).astype(np.float32)
    dX, dW = sr.conv_bwd(X, W, dY, None)
    assert np.allclose(dX, dY @ W.T, atol=1e-4) and np.allclose(dW, X.T @ dY, atol=1e-4)
