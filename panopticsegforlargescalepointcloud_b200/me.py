"""MinkowskiEngine-shaped namespace backed by the sm_100a kernels in libpgs_b200.so.

This is the host-side mirror of the part of MinkowskiEngine the reference's hot path touches
(SURVEY.md section 8b).  Call sites it must satisfy verbatim:

  torch_points3d/applications/minkowski.py:106-111   ME.MinkowskiConvolution / ME.utils.kaiming_normal_
  torch_points3d/applications/minkowski.py:121-122   ME.SparseTensor(features=, coordinates=, device=)
  torch_points3d/applications/minkowski.py:150,193   .F / .C
  torch_points3d/modules/MinkowskiEngine/api_modules.py:9,26-55,244-270,293,308
                                                     MinkowskiNetwork, MinkowskiConvolution(Transpose),
                                                     MinkowskiBatchNorm, MinkowskiReLU, cat, `a + b`
  torch_points3d/core/schedulers/bn_schedulers.py:7-15  MinkowskiBatchNorm / MinkowskiInstanceNorm symbols

Install as a drop-in with  `sys.modules["MinkowskiEngine"] = panopticsegforlargescalepointcloud_b200.me`
(see INTEGRATION.md).  Semantics frozen in DESIGN.md (row order, kernel-offset enumeration, strided
and transposed maps).  There is no CPU path: tensors must live on a CUDA device.
"""
import math
import types
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib
from ._lib import ptr, check, stream_ptr


# --------------------------------------------------------------------------------------------
# coordinate manager
# --------------------------------------------------------------------------------------------
class CoordinateMap:
    """One level of the hierarchy: unique int32 [n,4] coordinates + their device hash table."""

    __slots__ = ("coords", "n", "tkeys", "tvals", "cap", "tensor_stride")

    def __init__(self, coords, n, tkeys, tvals, cap, tensor_stride):
        self.coords, self.n, self.tkeys, self.tvals, self.cap = coords, n, tkeys, tvals, cap
        self.tensor_stride = tensor_stride


class KernelMap:
    """Gather table nbr[K, n_q] (+ lazily the ME-style pair lists used by the weight gradient and the
    occupancy-sorted copy used by the tensor-core kernels)."""

    __slots__ = ("nbr", "K", "n_q", "_pairs", "_sorted", "_n_pairs")

    def __init__(self, nbr, K, n_q):
        self.nbr, self.K, self.n_q = nbr, K, n_q
        self._pairs = None
        self._sorted = None
        self._n_pairs = None

    def n_pairs(self):
        """Rulebook size P (host value; synchronises once per kernel map -- bench.py's roofline pass only)."""
        if self._n_pairs is None:
            self._n_pairs = int((self.nbr >= 0).sum())
        return self._n_pairs

    def sorted(self):
        """-> (nbr_sorted int32[K, n_q], order int32[n_q]): the table with its rows sorted by their K-bit
        neighbour-occupancy mask, so that the rows of a 16 / 128-row tile share their empty kernel offsets and the
        tensor-core kernels skip them (include/pgs_b200.h, pgs_kmap_row_masks).  Tile row r is output row order[r]."""
        if self._sorted is None:
            lib = _lib.load()
            dev = self.nbr.device
            masks = torch.empty(max(self.n_q, 1), dtype=torch.int32, device=dev)
            check(lib.pgs_kmap_row_masks(ptr(self.nbr), self.n_q, self.K, ptr(masks), stream_ptr()))
            order = torch.sort(masks[:self.n_q], stable=True)[1].to(torch.int32)
            nbr_sorted = torch.empty_like(self.nbr)
            check(lib.pgs_kmap_permute(ptr(self.nbr), self.n_q, self.K, ptr(order), ptr(nbr_sorted), stream_ptr()))
            self._sorted = (nbr_sorted, order)
        return self._sorted

    def pairs(self):
        """(in_idx, out_idx, offs_dev, max_pairs): rulebook grouped by offset, ascending out row."""
        if self._pairs is None:
            lib = _lib.load()
            dev = self.nbr.device
            total = self.K * self.n_q
            in_idx = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
            out_idx = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
            offs = torch.empty(self.K + 1, dtype=torch.int32, device=dev)
            nb = lib.pgs_kmap_pairs_scratch_bytes(self.n_q, self.K)
            scratch = torch.empty(max(nb, 1), dtype=torch.uint8, device=dev)
            check(lib.pgs_kmap_pairs(ptr(self.nbr), self.n_q, self.K, ptr(in_idx), ptr(out_idx), ptr(offs),
                                     ptr(scratch), nb, stream_ptr()))
            # no host sync: the pair count per offset is bounded by n_q (the dW kernel exits on empty chunks)
            self._pairs = (in_idx, out_idx, offs, self.n_q)
        return self._pairs


def build_coordinate_map(coords, ts_out):
    """pgs_cmap_build on int32 [n, 4] (batch, x, y, z) rows -> (CoordinateMap of floor(c / ts_out) * ts_out with rows in
    first-occurrence order, in2out int32 [n]).  Also the voxel de-duplication of transforms.GridSampling3D."""
    lib = _lib.load()
    n = coords.shape[0]
    dev = coords.device
    cap = lib.pgs_cmap_capacity(n)
    tkeys = torch.empty(cap, dtype=torch.int64, device=dev)
    tvals = torch.empty(cap, dtype=torch.int32, device=dev)
    out_coords = torch.empty((max(n, 1), 4), dtype=torch.int32, device=dev)
    in2out = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    meta = torch.zeros(2, dtype=torch.int32, device=dev)  # [n_out, status]
    nb = lib.pgs_cmap_build_scratch_bytes(n)
    scratch = torch.empty(max(nb, 1), dtype=torch.uint8, device=dev)
    check(lib.pgs_cmap_build(ptr(coords), n, ts_out, ptr(tkeys), ptr(tvals), cap, ptr(out_coords),
                             ptr(in2out), ptr(meta[0:1]), ptr(meta[1:2]), ptr(scratch), nb, stream_ptr()))
    n_out, status = meta.tolist()
    if status & 1:
        raise ValueError("coordinate outside the supported range (batch < 65536, |xyz| < 32768)")
    if status & 2:
        raise _lib.PgsError("coordinate hash table overflow")
    return CoordinateMap(out_coords[:n_out], n_out, tkeys, tvals, cap, ts_out), in2out[:n]


class CoordinateManager:
    """Owns the coordinate maps (one per tensor stride) and kernel maps of one batch."""

    def __init__(self, coordinates: torch.Tensor):
        if coordinates.dim() != 2 or coordinates.shape[1] != 4:
            raise ValueError("coordinates must be [N, 1+3] (batch, x, y, z)")
        if not coordinates.is_cuda:
            raise _lib.PgsError("coordinates must be on a CUDA device (no CPU path)")
        self.device = coordinates.device
        self.maps: Dict[int, CoordinateMap] = {}
        self.kmaps: Dict[Tuple, KernelMap] = {}
        self.in2out: Dict[Tuple[int, int], torch.Tensor] = {}
        coords = coordinates.to(torch.int32).contiguous()
        m, in2out = self._build(coords, 1)
        if m.n != coords.shape[0]:
            raise ValueError(
                "duplicate coordinates in SparseTensor input (%d rows, %d unique): the reference feeds one "
                "point per voxel (grid_transform.py:185-194) and indexes outputs positionally" % (coords.shape[0], m.n)
            )
        self.maps[1] = m

    def _build(self, coords, ts_out):
        return build_coordinate_map(coords, ts_out)

    def get_map(self, ts: int) -> CoordinateMap:
        return self.maps[ts]

    def stride(self, ts_in: int, ts_out: int) -> CoordinateMap:
        if ts_out not in self.maps:
            m, in2out = self._build(self.maps[ts_in].coords, ts_out)
            self.maps[ts_out] = m
            self.in2out[(ts_in, ts_out)] = in2out
        return self.maps[ts_out]

    def kernel_map(self, ts_q: int, ts_probe: int, step: int, sign: int, ksize: int) -> KernelMap:
        key = (ts_q, ts_probe, step, sign, ksize)
        km = self.kmaps.get(key)
        if km is None:
            lib = _lib.load()
            q, p = self.maps[ts_q], self.maps[ts_probe]
            K = ksize ** 3
            nbr = torch.empty((K, max(q.n, 1)), dtype=torch.int32, device=self.device)
            check(lib.pgs_kmap_build(ptr(q.coords), q.n, ptr(p.tkeys), ptr(p.tvals), p.cap, step, sign, ksize,
                                     ptr(nbr), stream_ptr()))
            if q.n == 0:
                nbr = nbr[:, :0]
            km = KernelMap(nbr, K, q.n)
            self.kmaps[key] = km
        return km


# --------------------------------------------------------------------------------------------
# SparseTensor
# --------------------------------------------------------------------------------------------
class SparseTensor:
    """Features F [N,C] fp32 over a coordinate map.  Row i of F belongs to row i of C; for the input
    tensor that is the caller's row order (the reference relies on it: applications/minkowski.py:193)."""

    def __init__(self, features: torch.Tensor, coordinates: Optional[torch.Tensor] = None, device=None,
                 tensor_stride: int = 1, coordinate_manager: Optional[CoordinateManager] = None,
                 coordinate_map_key=None, **_ignored):
        if device is not None:
            device = torch.device(device)
            features = features.to(device)
            if coordinates is not None:
                coordinates = coordinates.to(device)
        if not features.is_cuda:
            raise _lib.PgsError("SparseTensor features must live on a CUDA device: this backend has no CPU path")
        if features.dtype != torch.float32:
            features = features.float()
        if coordinate_map_key is not None:
            tensor_stride = int(coordinate_map_key)
        if coordinate_manager is None:
            if coordinates is None:
                raise ValueError("either coordinates or a coordinate_manager is required")
            coordinate_manager = CoordinateManager(coordinates)
            tensor_stride = 1
        self._F = features
        self._lazy = None   # set by MinkowskiBatchNorm: callable(relu: bool) -> features (BN evaluated on demand,
        #                     so that a following MinkowskiReLU can fuse into the same kernel)
        self.coordinate_manager = coordinate_manager
        self.tensor_stride = int(tensor_stride)
        n_map = coordinate_manager.get_map(self.tensor_stride).n
        if features.shape[0] != n_map:
            raise ValueError("features have %d rows but the coordinate map has %d" % (features.shape[0], n_map))

    # ME-compatible accessors
    @property
    def F(self):
        if self._lazy is not None:
            self._F = self._lazy(False)
            self._lazy = None
        if self._F is None:
            raise RuntimeError("this BatchNorm output was consumed by a fused ReLU; use the ReLU's output")
        return self._F

    features = F

    @property
    def C(self):
        return self.coordinate_manager.get_map(self.tensor_stride).coords

    coordinates = C

    @property
    def coordinate_map_key(self):
        return self.tensor_stride

    @property
    def device(self):
        return self._F.device   # (_F of a deferred tensor is the BN input: same device / dtype / shape)

    @property
    def dtype(self):
        return self._F.dtype

    @property
    def D(self):
        return 3

    @property
    def shape(self):
        return self._F.shape

    def size(self, *a):
        return self._F.size(*a)

    def __len__(self):
        return self._F.shape[0]

    def _like(self, F):
        return SparseTensor(F, coordinate_manager=self.coordinate_manager, tensor_stride=self.tensor_stride)

    def _deferred(self, fn):
        """Same map, features = fn(relu) evaluated on first use (see MinkowskiBatchNorm / MinkowskiReLU)."""
        t = SparseTensor(self._F, coordinate_manager=self.coordinate_manager, tensor_stride=self.tensor_stride)
        t._lazy = fn
        return t

    def _check_same_map(self, other):
        if not isinstance(other, SparseTensor):
            return
        if other.coordinate_manager is not self.coordinate_manager or other.tensor_stride != self.tensor_stride:
            raise ValueError("sparse tensors live on different coordinate maps")

    def __add__(self, other):
        self._check_same_map(other)
        return self._like(self.F + (other.F if isinstance(other, SparseTensor) else other))

    __radd__ = __add__

    def __sub__(self, other):
        self._check_same_map(other)
        return self._like(self.F - (other.F if isinstance(other, SparseTensor) else other))

    def __mul__(self, other):
        self._check_same_map(other)
        return self._like(self.F * (other.F if isinstance(other, SparseTensor) else other))

    def __repr__(self):
        return "SparseTensor(n=%d, c=%d, tensor_stride=%d, device=%s)" % (
            self._F.shape[0], self._F.shape[1], self.tensor_stride, self._F.device)


def cat(*tensors):
    """Channel-wise concatenation of tensors that share a coordinate map (api_modules.py:308)."""
    if len(tensors) == 1 and isinstance(tensors[0], (list, tuple)):
        tensors = tuple(tensors[0])
    first = tensors[0]
    for t in tensors[1:]:
        first._check_same_map(t)
    return first._like(torch.cat([t.F for t in tensors], dim=1))


# --------------------------------------------------------------------------------------------
# convolution
# --------------------------------------------------------------------------------------------
# "auto": register-operand mma kernel (csrc/conv_mma.cu) for narrow layers with many rows, tcgen05 gather-GEMM
#         (csrc/conv_tc.cu) where the channel counts allow it, fp32 FFMA kernel (csrc/conv.cu) otherwise;
# "tc" / "mma": prefer that tensor-core kernel wherever it supports the shape;  "ffma": always the FFMA kernel.
# All are CUDA kernels behind the same C ABI.
import os as _os
CONV_IMPL = _os.environ.get("PGS_CONV_IMPL", "auto")
SMALL_COUT = int(_os.environ.get("PGS_SMALL_COUT", "0"))  # measured: the tensor-core path wins even at 16 channels

# bench.py sets this to a list to collect (start_event, end_event, algorithmic_bytes, flops) per conv launch
PROFILE = None


def conv_algorithmic_bytes(n_in, n_out, K, c_in, c_out, has_table):
    """Compulsory HBM bytes of one gather-GEMM launch (DESIGN.md "conv roofline"): read every input row once,
    write every output row once, read the weights once, read the gather table once."""
    return 4 * (n_in * c_in + n_out * c_out) + 4 * K * c_in * c_out + (4 * K * n_out if has_table else 0)


# Rows below which the register-operand mma kernel is not used: few-row layers are latency bound and the tcgen05
# kernel splits their kernel offsets over CTAs.  (csrc/conv_mma.cu; measured per shape in profiles/)
MMA_MIN_ROWS = int(_os.environ.get("PGS_MMA_MIN_ROWS", "8192"))
SPLIT_MAX_ROWS = int(_os.environ.get("PGS_SPLIT_MAX_ROWS", "2048"))
MMA_MAX_CH = int(_os.environ.get("PGS_MMA_MAX_CH", "64"))       # c_in bound
MMA_MAX_COUT = int(_os.environ.get("PGS_MMA_MAX_COUT", "48"))   # measured: tcgen05 wins from 64 output channels on


def _conv_kernel_choice(lib, K, c_in, c_out, n_q, has_table):
    """"mma" (register-operand tensor cores, narrow + tall layers), "tc" (tcgen05) or "ffma"."""
    if CONV_IMPL == "ffma":
        return "ffma"
    split_ok = has_table and K <= 27 and lib.pgs_conv_mma_split_supported(c_in, c_out)
    if split_ok and (CONV_IMPL == "split" or (CONV_IMPL == "auto" and n_q <= SPLIT_MAX_ROWS)):
        return "split"
    mma_ok = (has_table and K <= 27 and max(c_in, c_out) <= MMA_MAX_CH and lib.pgs_conv_mma_supported(c_in, c_out))
    if CONV_IMPL == "mma" and mma_ok:
        return "mma"
    if CONV_IMPL == "auto" and mma_ok and n_q >= MMA_MIN_ROWS and c_out <= MMA_MAX_COUT:
        return "mma"
    if K <= 27 and c_out > SMALL_COUT and lib.pgs_conv_tc_supported(c_in, c_out):
        return "tc"
    return "ffma"


# tensor-core kernels read the occupancy-sorted copy of the gather table (KernelMap.sorted) when there is enough work
# to pay for sorting it once per kernel map
SORT_TABLES = _os.environ.get("PGS_SORT_TABLES", "1") == "1"
SORT_MIN_ROWS = int(_os.environ.get("PGS_SORT_MIN_ROWS", "512"))


def _conv_scratch_bytes(lib, K, c_in, c_out):
    """Scratch that any conv kernel may need for this shape (the re-arranged weights)."""
    nb = 0
    if lib.pgs_conv_tc_supported(c_in, c_out):
        nb = lib.pgs_conv_tc_scratch_bytes(K, c_in, c_out)
    if lib.pgs_conv_mma_split_supported(c_in, c_out):
        nb = max(nb, lib.pgs_conv_mma_scratch_bytes(K, c_in, c_out))
    return nb


def _conv_launch(lib, Xp, n_in, Wp, K, c_in, c_out, km, n_q, mirror, w_transposed, Yp, scratch_p, scratch_bytes, sp,
                 kind=None, prepped=False):
    """Pick the kernel for this shape (or take `kind`) and launch it on raw device pointers (ints); returns the
    kernel kind.  km: KernelMap, raw table tensor or None (K == 1 identity).  prepped: `scratch_p` already holds the
    arranged weights (pgs_conv_prep_weights_batch), so the tensor-core entry points get W == NULL."""
    nbr = km.nbr if isinstance(km, KernelMap) else km
    if kind is None:
        kind = _conv_kernel_choice(lib, K, c_in, c_out, n_q, nbr is not None)
    if prepped and kind != "ffma":
        Wp = None
    nbr_p = order_p = None
    if nbr is not None:
        if kind != "ffma" and SORT_TABLES and isinstance(km, KernelMap) and n_q >= SORT_MIN_ROWS:
            ns, order = km.sorted()
            nbr_p, order_p = ns.data_ptr(), order.data_ptr()
        else:
            nbr_p = nbr.data_ptr()
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    if kind == "tc":
        rc = lib.pgs_conv_fwd_tc(Xp, Wp, nbr_p, order_p, n_q, K, c_in, c_out, int(mirror), int(w_transposed), Yp,
                                 scratch_p, scratch_bytes, sp)
    elif kind == "mma":
        rc = lib.pgs_conv_fwd_mma(Xp, Wp, nbr_p, order_p, n_q, K, c_in, c_out, int(mirror), int(w_transposed), Yp,
                                  scratch_p, scratch_bytes, sp)
    elif kind == "split":
        rc = lib.pgs_conv_fwd_mma_split(Xp, Wp, nbr_p, order_p, n_q, K, c_in, c_out, int(mirror), int(w_transposed), Yp,
                                        scratch_p, scratch_bytes, sp)
    else:
        rc = lib.pgs_conv_fwd(Xp, Wp, nbr_p, n_q, K, c_in, c_out, int(mirror), int(w_transposed), Yp, sp)
    if rc != 0:
        check(rc)
    if PROFILE is not None:
        e1.record()
        if nbr is None:
            pairs = n_q
        elif not PROFILE_COUNT_PAIRS:
            pairs = 0
        else:
            pairs = km.n_pairs() if isinstance(km, KernelMap) else int((nbr >= 0).sum())
        PROFILE.append((e0, e1, conv_algorithmic_bytes(n_in, n_q, K, c_in, c_out, nbr is not None),
                        2 * pairs * c_in * c_out, (n_in, n_q, K, c_in, c_out), kind))
    return kind


def _conv_fwd_raw(X, W3, km, n_q, mirror, w_transposed):
    """Y[q] = sum_k X[nbr[tk(k)][q]] W3[k]  (or with W3[k]^T when w_transposed); km: KernelMap, raw table or None."""
    lib = _lib.load()
    if not (X.is_cuda and W3.is_cuda):
        raise _lib.PgsError("expected CUDA tensors: the B200 path has no CPU implementation")
    K = W3.shape[0]
    c_in, c_out = (W3.shape[2], W3.shape[1]) if w_transposed else (W3.shape[1], W3.shape[2])
    Y = torch.empty((n_q, c_out), dtype=torch.float32, device=X.device)
    nb = _conv_scratch_bytes(lib, K, c_in, c_out)
    scratch = torch.empty(max(nb, 1), dtype=torch.uint8, device=X.device)
    _conv_launch(lib, ptr(X).value, X.shape[0], ptr(W3).value, K, c_in, c_out, km, n_q, mirror, w_transposed,
                 Y.data_ptr(), scratch.data_ptr(), nb, stream_ptr())
    return Y


PROFILE_COUNT_PAIRS = False
PROFILE_DW = None


# Weight gradients are accumulated by the dW kernel (atomicAdd) straight into `param.grad` when that buffer
# exists, on a side stream: no zero-filled temporary, no AccumulateGrad add, and the dW launch overlaps the
# input-gradient chain of the remaining backward pass.  Whoever consumes the gradients (optimizer step, gradient
# all-reduce) must call `join_side_stream()` first -- BaseModel.optimize_parameters2 and parallel.FlatGradBucket do.
DW_DIRECT = _os.environ.get("PGS_DW_DIRECT", "0") == "1"   # measured slower (41.4 vs 35.9 ms per step): off
_SIDE = {}


def _side_stream(device):
    s = _SIDE.get(device.index)
    if s is None:
        s = _SIDE[device.index] = torch.cuda.Stream(device=device)
    return s


def join_side_stream():
    """Make the current stream wait for every weight-gradient kernel launched on the side stream."""
    if torch.cuda.is_available():
        s = _SIDE.get(torch.cuda.current_device())
        if s is not None:
            torch.cuda.current_stream().wait_stream(s)


class _SparseConvFn(torch.autograd.Function):
    """fwd / bwd-input / bwd-weight through the C ABI (pgs_conv_fwd, pgs_conv_bwd_weight)."""

    @staticmethod
    def forward(ctx, X, W, km_f, km_b, mirror_f, mirror_b, n_out):
        X = X.contiguous()
        W3 = W.reshape(-1, W.shape[-2], W.shape[-1]).contiguous()
        Y = _conv_fwd_raw(X, W3, km_f, n_out, mirror_f, False)
        ctx.save_for_backward(X, W)
        ctx.km_f, ctx.km_b, ctx.mirror_f, ctx.mirror_b = km_f, km_b, mirror_f, mirror_b
        ctx.param = W if isinstance(W, nn.Parameter) else None
        return Y

    @staticmethod
    def backward(ctx, dY):
        X, W = ctx.saved_tensors
        dY = dY.contiguous()
        W3 = W.reshape(-1, W.shape[-2], W.shape[-1]).contiguous()
        K, c_in, c_out = W3.shape
        lib = _lib.load()
        dX = dW = None
        if ctx.needs_input_grad[0]:
            dX = _conv_fwd_raw(dY, W3, ctx.km_b, X.shape[0], ctx.mirror_b, True)
        if ctx.needs_input_grad[1]:
            direct = (DW_DIRECT and ctx.param is not None and ctx.param.grad is not None
                      and ctx.param.grad.is_contiguous() and ctx.param.grad.dtype == torch.float32
                      and PROFILE_DW is None)
            if direct:
                # everything the side-stream kernel reads must exist before the side stream forks off the
                # current one, and must be kept from the caching allocator until that kernel has run
                pl = ctx.km_f.pairs() if ctx.km_f is not None else None
                cur = torch.cuda.current_stream()
                side = _side_stream(X.device)
                side.wait_stream(cur)
                sp = _lib.c_void_p(side.cuda_stream)
                g = ctx.param.grad
                X.record_stream(side)
                dY.record_stream(side)
                if pl is not None:
                    in_idx, out_idx, offs, max_pairs = pl
                    for t in (in_idx, out_idx, offs):
                        t.record_stream(side)
                    check(lib.pgs_conv_bwd_weight(ptr(X), ptr(dY), ptr(in_idx), ptr(out_idx), ptr(offs), max_pairs,
                                                  K, c_in, c_out, int(ctx.mirror_f), ptr(g), sp))
                else:
                    check(lib.pgs_conv_bwd_weight(ptr(X), ptr(dY), None, None, None, X.shape[0], 1, c_in, c_out, 0,
                                                  ptr(g), sp))
                return dX, None, None, None, None, None, None
            dW3 = torch.zeros_like(W3)
            if PROFILE_DW is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            if ctx.km_f is not None:
                in_idx, out_idx, offs, max_pairs = ctx.km_f.pairs()
                check(lib.pgs_conv_bwd_weight(ptr(X), ptr(dY), ptr(in_idx), ptr(out_idx), ptr(offs), max_pairs,
                                              K, c_in, c_out, int(ctx.mirror_f), ptr(dW3), stream_ptr()))
            else:
                check(lib.pgs_conv_bwd_weight(ptr(X), ptr(dY), None, None, None, X.shape[0], 1, c_in, c_out, 0,
                                              ptr(dW3), stream_ptr()))
            if PROFILE_DW is not None:
                e1.record()
                PROFILE_DW.append((e0, e1, (X.shape[0], dY.shape[0], K, c_in, c_out)))
            dW = dW3.reshape(W.shape)
        return dX, dW, None, None, None, None, None


def _as_int(v, name):
    if isinstance(v, (list, tuple)):
        if len(set(v)) != 1:
            raise NotImplementedError("anisotropic %s is not supported" % name)
        v = v[0]
    return int(v)


class MinkowskiConvolutionBase(nn.Module):
    TRANSPOSE = False

    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
                 kernel_generator=None, expand_coordinates=False, convolution_mode=None, dimension=None):
        super().__init__()
        if dimension is not None and dimension != 3:
            raise NotImplementedError("only 3-D sparse convolution is supported")
        if expand_coordinates:
            raise NotImplementedError("coordinate expansion is not supported")
        if kernel_generator is not None:      # a plain hyper-cube generator only restates the scalar arguments
            if not isinstance(kernel_generator, KernelGenerator):
                raise NotImplementedError("custom kernel generators are not supported")
            kernel_size, stride, dilation = (kernel_generator.kernel_size, kernel_generator.stride,
                                             kernel_generator.dilation)
        self.in_channels, self.out_channels = int(in_channels), int(out_channels)
        self.kernel_size = _as_int(kernel_size, "kernel_size")
        self.stride = _as_int(stride, "stride")
        self.dilation = _as_int(dilation, "dilation")
        if self.kernel_size not in (1, 2, 3, 5):
            raise NotImplementedError("kernel_size %r is not supported" % (kernel_size,))
        if self.dilation != 1:
            raise NotImplementedError("dilation != 1 is not used by the reference hot path")
        if self.kernel_size == 1 and self.stride != 1:
            raise NotImplementedError("kernel_size 1 with stride > 1")
        self.kernel_volume = self.kernel_size ** 3
        self.dimension = 3
        if self.kernel_volume == 1:
            self.kernel = nn.Parameter(torch.empty(self.in_channels, self.out_channels))
        else:
            self.kernel = nn.Parameter(torch.empty(self.kernel_volume, self.in_channels, self.out_channels))
        self.bias = nn.Parameter(torch.empty(1, self.out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        # ME default: uniform(+-1/sqrt(fan)), fan = Cin*K (conv) or Cout*K (transposed)  [SURVEY App. B.6]
        with torch.no_grad():
            n = (self.out_channels if self.TRANSPOSE else self.in_channels) * self.kernel_volume
            stdv = 1.0 / math.sqrt(n)
            self.kernel.uniform_(-stdv, stdv)
            if self.bias is not None:
                self.bias.uniform_(-stdv, stdv)

    def _maps(self, x: SparseTensor):
        return self.maps_for(x.coordinate_manager, x.tensor_stride, x.F.shape[0])

    def maps_for(self, cm, ts, n_in):
        """-> (km_fwd, km_bwd, mirror_f, mirror_b, out_tensor_stride, n_out) for an input on map `ts` of `cm`"""
        ks = self.kernel_size
        if not self.TRANSPOSE:
            if self.stride == 1:
                if ks == 1:
                    return None, None, False, False, ts, n_in
                km = cm.kernel_map(ts, ts, ts, +1, ks)
                return km, km, False, True, ts, km.n_q
            ts_out = ts * self.stride
            coarse = cm.stride(ts, ts_out)
            km_f = cm.kernel_map(ts_out, ts, ts, +1, ks)
            km_b = cm.kernel_map(ts, ts_out, ts, -1, ks)
            return km_f, km_b, False, False, ts_out, coarse.n
        # transposed
        if self.stride == 1:
            if ks == 1:
                return None, None, False, False, ts, n_in
            km = cm.kernel_map(ts, ts, ts, +1, ks)
            return km, km, True, False, ts, km.n_q
        if ts % self.stride != 0 or (ts // self.stride) not in cm.maps:
            raise NotImplementedError(
                "transposed convolution onto a coordinate map that the encoder did not create (tensor stride %d / %d)"
                % (ts, self.stride))
        ts_out = ts // self.stride
        fine = cm.get_map(ts_out)
        km_f = cm.kernel_map(ts_out, ts, ts_out, -1, ks)
        km_b = cm.kernel_map(ts, ts_out, ts_out, +1, ks)
        return km_f, km_b, False, False, ts_out, fine.n

    def forward(self, x: SparseTensor):
        if x.F.shape[1] != self.in_channels:
            raise ValueError("expected %d input channels, got %d" % (self.in_channels, x.F.shape[1]))
        km_f, km_b, mirror_f, mirror_b, ts_out, n_out = self._maps(x)
        Y = _SparseConvFn.apply(x.F, self.kernel, km_f, km_b, mirror_f, mirror_b, n_out)
        if self.bias is not None:
            Y = Y + self.bias
        return SparseTensor(Y, coordinate_manager=x.coordinate_manager, tensor_stride=ts_out)

    def extra_repr(self):
        return "in=%d, out=%d, kernel_size=%d, stride=%d, dilation=%d" % (
            self.in_channels, self.out_channels, self.kernel_size, self.stride, self.dilation)


class MinkowskiConvolution(MinkowskiConvolutionBase):
    TRANSPOSE = False


class MinkowskiConvolutionTranspose(MinkowskiConvolutionBase):
    TRANSPOSE = True


# --------------------------------------------------------------------------------------------
# pointwise modules
# --------------------------------------------------------------------------------------------
class MinkowskiNetwork(nn.Module):
    def __init__(self, D=3):
        super().__init__()
        self.D = D


class _BnFn(torch.autograd.Function):
    """Fused BatchNorm (+ ReLU) through pgs_bn_forward / pgs_bn_backward."""

    @staticmethod
    def forward(ctx, X, weight, bias, running_mean, running_var, training, momentum, eps, relu):
        lib = _lib.load()
        X = X.contiguous()
        n, C = X.shape
        dev = X.device
        Y = torch.empty_like(X)
        sums = torch.empty(2 * C, dtype=torch.float64, device=dev)
        stats = torch.empty((2, C), dtype=torch.float32, device=dev)
        check(lib.pgs_bn_forward(ptr(X), n, C, ptr(weight), ptr(bias), ptr(running_mean), ptr(running_var),
                                 int(training), float(momentum), float(eps), int(relu), ptr(sums), ptr(stats[0]),
                                 ptr(stats[1]), ptr(Y), stream_ptr()))
        ctx.save_for_backward(X, Y if relu else None, weight, stats)
        ctx.training, ctx.relu = bool(training), bool(relu)
        return Y

    @staticmethod
    def backward(ctx, dY):
        lib = _lib.load()
        X, Y, weight, stats = ctx.saved_tensors
        dY = dY.contiguous()
        n, C = X.shape
        dX = torch.empty_like(X)
        sums = torch.empty(2 * C, dtype=torch.float64, device=X.device)
        dwb = torch.empty((2, C), dtype=torch.float32, device=X.device)
        check(lib.pgs_bn_backward(ptr(X), ptr(Y), ptr(dY), n, C, ptr(weight), ptr(stats[0]), ptr(stats[1]),
                                  int(ctx.training), int(ctx.relu), ptr(sums), ptr(dX), ptr(dwb[0]), ptr(dwb[1]),
                                  stream_ptr()))
        dw = dwb[0] if (weight is not None and ctx.needs_input_grad[1]) else None
        db = dwb[1] if ctx.needs_input_grad[2] else None
        return dX, dw, db, None, None, None, None, None, None


BN_IMPL = _os.environ.get("PGS_BN_IMPL", "fused")   # "torch": nn.BatchNorm1d kernels + separate ReLU


class MinkowskiBatchNorm(nn.Module):
    """nn.BatchNorm1d over all active rows; the inner module is `.bn` (checkpoint key contract).  The arithmetic
    runs in the fused kernels of csrc/norm.cu (and absorbs a directly following MinkowskiReLU)."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum, affine=affine,
                                 track_running_stats=track_running_stats)

    @property
    def momentum(self):   # core/schedulers/bn_schedulers.py pokes .momentum on BN modules
        return self.bn.momentum

    @momentum.setter
    def momentum(self, v):
        self.bn.momentum = v

    def _fused_ok(self, F):
        bn = self.bn
        return (BN_IMPL == "fused" and F.is_cuda and F.dtype == torch.float32 and F.shape[1] % 4 == 0
                and F.shape[1] <= 1024 and F.shape[0] > 0 and bn.momentum is not None
                and (bn.training or bn.track_running_stats))

    def forward(self, x: SparseTensor):
        F = x.F
        if not self._fused_ok(F):
            return x._like(self.bn(F))
        bn = self.bn
        training = bn.training or not bn.track_running_stats
        if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)

        def run(relu):
            return _BnFn.apply(F, bn.weight, bn.bias, bn.running_mean if bn.track_running_stats else None,
                               bn.running_var if bn.track_running_stats else None, training, bn.momentum, bn.eps, relu)

        return x._deferred(run)


class MinkowskiInstanceNorm(nn.Module):
    """Symbol required by core/schedulers/bn_schedulers.py:7-15; not used on the hot path."""

    def __init__(self, num_features):
        super().__init__()
        self.num_features = num_features

    def forward(self, x):
        raise NotImplementedError("MinkowskiInstanceNorm is not on the reference hot path")


class MinkowskiReLU(nn.Module):
    def __init__(self, inplace=False):
        super().__init__()

    def forward(self, x: SparseTensor):
        if x._lazy is not None:          # BN -> ReLU: one fused kernel pair
            fn, x._lazy = x._lazy, None
            y = x._like(fn(True))
            x._F = None          # the un-rectified BN output was never materialised
            return y
        return x._like(torch.relu(x.F))


class MinkowskiLeakyReLU(nn.Module):
    def __init__(self, negative_slope=0.01, inplace=False):
        super().__init__()
        self.negative_slope = negative_slope

    def forward(self, x: SparseTensor):
        return x._like(torch.nn.functional.leaky_relu(x.F, self.negative_slope))


class MinkowskiSigmoid(nn.Module):
    def forward(self, x: SparseTensor):
        return x._like(torch.sigmoid(x.F))


class MinkowskiLinear(nn.Module):
    """nn.Linear on the feature matrix (api_modules.py:129-134, SELayer; not used by the shipped configs)."""

    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.linear = nn.Linear(in_features, out_features, bias=bias)

    def forward(self, x: SparseTensor):
        return x._like(self.linear(x.F))


class RegionType:
    """ME.RegionType: modules/MinkowskiEngine/common.py:53-62 builds lookup tables from these at import time.
    Only HYPER_CUBE kernels exist on the reference hot path."""
    HYPER_CUBE, HYPER_CROSS, CUSTOM = 0, 1, 2


class KernelGenerator:
    """ME.KernelGenerator(kernel_size, stride, dilation, region_type=, dimension=): accepted by the convolutions when it
    describes a plain hyper-cube (modules/MinkowskiEngine/common.py:117-146)."""

    def __init__(self, kernel_size=-1, stride=1, dilation=1, is_transpose=False, region_type=RegionType.HYPER_CUBE,
                 region_offsets=None, axis_types=None, dimension=3, **_ignored):
        if region_type != RegionType.HYPER_CUBE or region_offsets is not None or axis_types is not None:
            raise NotImplementedError("only hyper-cube kernels are on the reference hot path")
        self.kernel_size, self.stride, self.dilation, self.dimension = kernel_size, stride, dilation, dimension
        self.region_type = region_type


def _not_on_hot_path(name):
    class _Stub(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

        def forward(self, *a, **k):
            raise NotImplementedError("%s is not on the reference hot path (SURVEY 2.1: SE / pooling variants are "
                                      "not selected by any shipped panoptic config)" % name)
    _Stub.__name__ = _Stub.__qualname__ = name
    return _Stub


# symbols the reference's model zoo names at import / construction time but never runs on the panoptic path
MinkowskiGlobalPooling = _not_on_hot_path("MinkowskiGlobalPooling")
MinkowskiGlobalMaxPooling = _not_on_hot_path("MinkowskiGlobalMaxPooling")
MinkowskiBroadcastMultiplication = _not_on_hot_path("MinkowskiBroadcastMultiplication")
MinkowskiAvgPooling = _not_on_hot_path("MinkowskiAvgPooling")
MinkowskiSumPooling = _not_on_hot_path("MinkowskiSumPooling")
MinkowskiAvgUnpooling = _not_on_hot_path("MinkowskiAvgUnpooling")


def _fans(tensor):
    if tensor.dim() < 2:
        raise ValueError("fan in / fan out need at least 2 dimensions")
    if tensor.dim() == 2:  # ME treats a [Cin, Cout] (K == 1) kernel like nn.Linear's [out, in]
        return tensor.size(1), tensor.size(0)
    rf = tensor.size(0)
    return tensor.size(1) * rf, tensor.size(2) * rf


def kaiming_normal_(tensor, a=0, mode="fan_in", nonlinearity="leaky_relu"):
    """ME.utils.kaiming_normal_ for [K, Cin, Cout] kernels (applications/minkowski.py:107)."""
    fan_in, fan_out = _fans(tensor)
    fan = fan_in if mode == "fan_in" else fan_out
    std = nn.init.calculate_gain(nonlinearity, a) / math.sqrt(fan)
    with torch.no_grad():
        return tensor.normal_(0, std)


utils = types.SimpleNamespace(kaiming_normal_=kaiming_normal_)

# `import MinkowskiEngine.MinkowskiOps as me` / `import MinkowskiEngine.MinkowskiFunctional as MEF` in the reference's
# stock model zoo (modules/MinkowskiEngine/res16unet.py:5, resunet.py:3): registered as sub-modules by bind.install()
MinkowskiOps = types.ModuleType("MinkowskiEngine.MinkowskiOps")
MinkowskiOps.cat = cat
MinkowskiFunctional = types.ModuleType("MinkowskiEngine.MinkowskiFunctional")
MinkowskiFunctional.relu = lambda x, *a, **k: x._like(torch.relu(x.F))

__version__ = "0.5.4+b200"
