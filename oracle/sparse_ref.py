"""ORACLE (test infrastructure, not product code): CPU restatement of the MinkowskiEngine semantics the
reference's hot path relies on -- coordinate maps, kernel maps (rulebooks) and the sparse convolution
forward / backward -- in plain numpy.

PARITY UNPINNED: MinkowskiEngine is an un-vendored dependency of the reference (README.md:41,85-87) that is
absent from /root/reference and from this image, and the reference ships no tests or golden vectors for it.
The definitions below follow SURVEY.md Appendix B (row order B.1, offset enumeration B.2, strided maps B.3,
transposed maps B.4, backward B.7) and are pinned independently against dense torch.nn.functional.conv3d /
conv_transpose3d in tests/test_oracle_sparse.py.

Reference call sites this restates:
  torch_points3d/applications/minkowski.py:121-122            ME.SparseTensor(features, coordinates)
  torch_points3d/modules/MinkowskiEngine/api_modules.py:26-55  ResBlock convs (k3 s1, k1)
  torch_points3d/modules/MinkowskiEngine/api_modules.py:244-270 ResNetDown.conv_in (k3 s2)
  torch_points3d/modules/MinkowskiEngine/api_modules.py:293     ResNetUp -> MinkowskiConvolutionTranspose

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import numpy as np


def pack_keys(coords):
    """(b,x,y,z) int -> int64 key; same packing as csrc/common.cuh pack_key (16 bits per field)."""
    c = np.asarray(coords, dtype=np.int64)
    return (c[:, 0] << 48) | ((c[:, 1] + 32768) << 32) | ((c[:, 2] + 32768) << 16) | (c[:, 3] + 32768)


def coordinate_map(coords, tensor_stride_out=1):
    """Unique floor(c/ts)*ts rows in FIRST-OCCURRENCE input order.

    Returns (out_coords int32 [n_out,4], in2out int32 [n])."""
    c = np.asarray(coords, dtype=np.int64).copy()
    ts = int(tensor_stride_out)
    if ts > 1:
        c[:, 1:] = np.floor_divide(c[:, 1:], ts) * ts
    keys = pack_keys(c)
    uniq, first, inv = np.unique(keys, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")  # unique keys ranked by first occurrence
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    out_coords = c[first[order]].astype(np.int32)
    in2out = rank[inv.reshape(-1)].astype(np.int32)
    return out_coords, in2out


def kernel_offsets(ksize):
    """delta_k with x fastest: k = ix + ksize*iy + ksize^2*iz; odd kernels centred, even {0..k-1}."""
    half = ksize // 2 if ksize % 2 == 1 else 0
    r = np.arange(ksize) - half
    dz, dy, dx = np.meshgrid(r, r, r, indexing="ij")
    return np.stack([dx.reshape(-1), dy.reshape(-1), dz.reshape(-1)], axis=1)  # [K,3]


def kernel_map(q_coords, probe_coords, step, sign=1, ksize=3):
    """nbr[k, q] = row r of probe with c_r == c_q + sign*delta_k*step, else -1   (int32 [K, n_q])."""
    q = np.asarray(q_coords, dtype=np.int64)
    p = np.asarray(probe_coords, dtype=np.int64)
    pk = pack_keys(p)
    order = np.argsort(pk, kind="stable")
    spk = pk[order]
    offs = kernel_offsets(ksize) * int(step) * int(sign)
    K = offs.shape[0]
    nbr = np.full((K, q.shape[0]), -1, dtype=np.int32)
    if p.shape[0] == 0 or q.shape[0] == 0:
        return nbr
    for k in range(K):
        c = q.copy()
        c[:, 1:] += offs[k]
        ok = np.all((c[:, 1:] >= -32768) & (c[:, 1:] < 32768), axis=1)
        kk = pack_keys(c)
        pos = np.searchsorted(spk, kk)
        pos = np.clip(pos, 0, spk.size - 1)
        hit = ok & (spk[pos] == kk)
        nbr[k, hit] = order[pos[hit]]
    return nbr


def kernel_map_dict(q_coords, probe_coords, step, sign=1, ksize=3):
    """Same as kernel_map with a literal python dict (small cases; cross-checks the vectorised form)."""
    table = {tuple(int(v) for v in c): i for i, c in enumerate(np.asarray(probe_coords))}
    offs = kernel_offsets(ksize) * int(step) * int(sign)
    nbr = np.full((offs.shape[0], len(q_coords)), -1, dtype=np.int32)
    for qi, c in enumerate(np.asarray(q_coords)):
        for k, d in enumerate(offs):
            nbr[k, qi] = table.get((int(c[0]), int(c[1] + d[0]), int(c[2] + d[1]), int(c[3] + d[2])), -1)
    return nbr


def pairs(nbr):
    """ME-style rulebook: (in_idx, out_idx, offs[K+1]); grouped by offset, ascending out row."""
    K = nbr.shape[0]
    k_idx, q_idx = np.nonzero(nbr >= 0)  # row-major => k ascending, then q ascending
    in_idx = nbr[k_idx, q_idx].astype(np.int32)
    offs = np.zeros(K + 1, dtype=np.int32)
    np.cumsum(np.bincount(k_idx, minlength=K), out=offs[1:])
    return in_idx, q_idx.astype(np.int32), offs


def conv_fwd(X, W, nbr, mirror=False, dtype=np.float64):
    """Y[q] = sum_k X[nbr[tk(k)][q]] @ W[k];  W [K,Cin,Cout];  nbr None => K==1 identity."""
    X = np.asarray(X, dtype=dtype)
    W = np.asarray(W, dtype=dtype)
    if W.ndim == 2:
        W = W[None]
    K = W.shape[0]
    if nbr is None:
        return (X @ W[0]).astype(np.float32)
    Y = np.zeros((nbr.shape[1], W.shape[2]), dtype=dtype)
    for k in range(K):
        tk = K - 1 - k if mirror else k
        idx = nbr[tk]
        m = idx >= 0
        if m.any():
            Y[m] += X[idx[m]] @ W[k]
    return Y.astype(np.float32)


def conv_bwd(X, W, dY, nbr, mirror=False, dtype=np.float64):
    """(dX, dW) of conv_fwd  (dX[in] += dY[out] W_k^T ; dW_k = X[in]^T dY[out])."""
    X = np.asarray(X, dtype=dtype)
    dY = np.asarray(dY, dtype=dtype)
    W3 = np.asarray(W, dtype=dtype)
    if W3.ndim == 2:
        W3 = W3[None]
    K = W3.shape[0]
    dX = np.zeros_like(X)
    dW = np.zeros_like(W3)
    if nbr is None:
        dX = dY @ W3[0].T
        dW[0] = X.T @ dY
    else:
        for k in range(K):
            tk = K - 1 - k if mirror else k
            idx = nbr[tk]
            m = np.nonzero(idx >= 0)[0]
            if m.size:
                np.add.at(dX, idx[m], dY[m] @ W3[k].T)
                dW[k] = X[idx[m]].T @ dY[m]
    return dX.astype(np.float32), dW.reshape(np.asarray(W).shape).astype(np.float32)


class Maps:
    """The map bookkeeping of one batch, mirroring me.CoordinateManager for the oracle."""

    def __init__(self, coords):
        c, in2out = coordinate_map(coords, 1)
        if c.shape[0] != len(coords):
            raise ValueError("duplicate coordinates")
        self.coords = {1: np.asarray(coords, dtype=np.int32)}
        self.kmaps = {}

    def stride(self, ts_in, ts_out):
        if ts_out not in self.coords:
            self.coords[ts_out] = coordinate_map(self.coords[ts_in], ts_out)[0]
        return self.coords[ts_out]

    def kernel_map(self, ts_q, ts_probe, step, sign, ksize):
        key = (ts_q, ts_probe, step, sign, ksize)
        if key not in self.kmaps:
            self.kmaps[key] = kernel_map(self.coords[ts_q], self.coords[ts_probe], step, sign, ksize)
        return self.kmaps[key]

    def conv_maps(self, ts, ksize, stride, transpose):
        """-> (nbr_fwd, mirror_fwd, nbr_bwd, mirror_bwd, ts_out) exactly as me.MinkowskiConvolutionBase._maps."""
        if stride == 1:
            if ksize == 1:
                return None, False, None, False, ts
            km = self.kernel_map(ts, ts, ts, +1, ksize)
            return (km, True, km, False, ts) if transpose else (km, False, km, True, ts)
        if not transpose:
            ts_out = ts * stride
            self.stride(ts, ts_out)
            return (self.kernel_map(ts_out, ts, ts, +1, ksize), False,
                    self.kernel_map(ts, ts_out, ts, -1, ksize), False, ts_out)
        ts_out = ts // stride
        return (self.kernel_map(ts_out, ts, ts_out, -1, ksize), False,
                self.kernel_map(ts, ts_out, ts_out, +1, ksize), False, ts_out)
