"""`b200` sparse backend for the reference's own plugin hook.

The reference switches sparse-conv backends through `torch_points3d.modules.SparseConv3d.nn.set_backend(name)`
(modules/SparseConv3d/nn/__init__.py:21-52), which imports `torch_points3d.modules.SparseConv3d.nn.<name>` and
re-exports exactly six symbols: cat, Conv3d, Conv3dTranspose, ReLU, SparseTensor, BatchNorm -- with the signatures of
the stock `nn/minkowski.py:5-65` backend.  This module provides those six on the sm_100a kernels (me.py ->
libpgs_b200.so); `bind.install()` registers it under the dotted name the hook imports, and
`bind.enable_sparse_backend()` binds it without touching the reference's `backend_valid` whitelist.
(With `bind.install()` alone the reference's stock "minkowski" backend already runs on this library, because its
`import MinkowskiEngine as ME` resolves to me.py.)
"""
import torch

from .. import me

__all__ = ["cat", "Conv3d", "Conv3dTranspose", "ReLU", "SparseTensor", "BatchNorm"]


def _conv(base):
    class _Conv(base):
        """(in_channels, out_channels, kernel_size=3, stride=1, dilation=1, bias=False); parameter `.kernel`."""

        def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, dilation=1, bias=False):
            base.__init__(self, in_channels, out_channels, kernel_size=kernel_size, stride=stride, dilation=dilation,
                          bias=bias, dimension=3)
    return _Conv


Conv3d = _conv(me.MinkowskiConvolution)
Conv3d.__name__ = Conv3d.__qualname__ = "Conv3d"
Conv3dTranspose = _conv(me.MinkowskiConvolutionTranspose)
Conv3dTranspose.__name__ = Conv3dTranspose.__qualname__ = "Conv3dTranspose"


class BatchNorm(me.MinkowskiBatchNorm):
    """BatchNorm(C); the wrapped nn.BatchNorm1d is `.bn` (checkpoint keys, BN-momentum scheduler)."""

    def __repr__(self):
        return repr(self.bn)


class ReLU(me.MinkowskiReLU):
    def __init__(self, inplace=False):
        super().__init__(inplace=False)


def cat(*tensors):
    return me.cat(*tensors)


def SparseTensor(feats, coordinates, batch, device=None):
    """(feats [N,C], coordinates int [N,3], batch [N] or [N,1], device) -> sparse tensor with .F, .C, `+`.
    The stock backend defaults `device` to the CPU; this backend has no CPU path, so the default is the features' device
    when they already live on a GPU, else the current CUDA device."""
    if batch.dim() == 1:
        batch = batch.unsqueeze(-1)
    if device is None or torch.device(device).type != "cuda":
        device = feats.device if feats.is_cuda else torch.device("cuda", torch.cuda.current_device())
    coords = torch.cat([batch.to(coordinates.device).int(), coordinates.int()], -1)
    return me.SparseTensor(features=feats, coordinates=coords, device=device)
