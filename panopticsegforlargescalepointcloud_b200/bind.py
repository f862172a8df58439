"""The reference-side binding: make the reference's own imports resolve to this package.

The reference has no FFI for its hot path -- the boundary is three `import` statements of un-vendored native packages
(SURVEY 8b).  `install()` is the stub a maintainer adds at the top of train.py / eval.py:

    import panopticsegforlargescalepointcloud_b200.bind as b200; b200.install()

after which, unmodified,
    torch_points3d/modules/MinkowskiEngine/api_modules.py:2        `import MinkowskiEngine as ME`
    torch_points3d/applications/minkowski.py:106-122               ME.MinkowskiConvolution / ME.SparseTensor / ME.utils
    torch_points3d/models/panoptic/PointGroup3heads.py:3           `from torch_points_kernels import region_grow`
    torch_points3d/core/losses/panoptic_losses.py:3                `from torch_points_kernels import instance_iou`
    torch_points3d/utils/hdbscan_cluster.py:4                      `import hdbscan`
    torch_points3d/utils/meanshift_cluster.py:4                    `from sklearn.cluster import MeanShift` (opt-in)
    torch_points3d/modules/SparseConv3d/nn/__init__.py:21-52       set_backend("b200")
run on libpgs_b200.so.  tests/test_reference_binding.py and tests/test_gpu_reference_binding.py drive the reference's
own files through exactly this call.
"""
import importlib
import sys


def install(meanshift=True, sparse_backend=True, fuse_unet=True):
    """Register the drop-in modules under the names the reference imports.  Idempotent.
    meanshift: also route `torch_points3d.utils.meanshift_cluster.MeanShift` to the device mean shift when that module is
    (or gets) imported.  sparse_backend: make `SPARSE_BACKEND=b200` / `sp3d.nn.set_backend("b200")` resolvable.
    fuse_unet: run the reference's own `MinkowskiUnet` / `MinkowskiEncoder` (applications/minkowski.py:129-196) through the
    fused executor (fastpath.py: the whole backbone as one autograd node, one C call per direction) instead of module by
    module -- same kernels, same results, ~500 fewer Python -> C round trips per step."""
    from . import me, tpk, hdbscan
    sys.modules["MinkowskiEngine"] = me
    sys.modules["MinkowskiEngine.MinkowskiOps"] = me.MinkowskiOps
    sys.modules["MinkowskiEngine.MinkowskiFunctional"] = me.MinkowskiFunctional
    sys.modules["torch_points_kernels"] = tpk
    sys.modules["hdbscan"] = hdbscan
    if sparse_backend:
        from .nn import b200
        # modules/SparseConv3d/nn/__init__.py:33-52 imports "torch_points3d.modules.SparseConv3d.nn.<backend>"
        sys.modules["torch_points3d.modules.SparseConv3d.nn.b200"] = b200
    if meanshift:
        _patch_meanshift()
    if fuse_unet:
        _after_import("torch_points3d.applications.minkowski", _fuse_reference_unet)
    return me, tpk, hdbscan


def _after_import(name, fn):
    """Call fn(module) now if `name` is imported already, else right after the reference imports it."""
    mod = sys.modules.get(name)
    if mod is not None:
        fn(mod)
        return

    class _Hook:
        def find_spec(self, fullname, path=None, target=None):
            if fullname != name:
                return None
            sys.meta_path.remove(self)
            spec = importlib.util.find_spec(fullname)
            if spec is None or spec.loader is None:
                return None
            orig = spec.loader.exec_module

            def exec_module(module):
                orig(module)
                fn(module)

            spec.loader.exec_module = exec_module
            return spec

    sys.meta_path.insert(0, _Hook())


def _fuse_reference_unet(mod):
    """Wrap the forward of the reference's MinkowskiUnet / MinkowskiEncoder: try the fused executor on the module tree
    they built (duck-typed tape, fastpath.Program); anything it does not recognise falls back to the original forward."""
    from . import fastpath

    def wrap(cls, make_out):
        orig = cls.forward
        if getattr(orig, "_pgs_fused", False):
            return

        def forward(self, data, *args, **kwargs):
            if fastpath.ENABLED and fastpath.program_for(self) is not None:
                self._set_input(data)
                out = fastpath.run(self, self.input)
                if out is not None:
                    return make_out(self, out)
            return orig(self, data, *args, **kwargs)

        forward._pgs_fused = True
        forward.__doc__ = orig.__doc__
        cls.forward = forward

    def unet_out(self, out):       # applications/minkowski.py:193-196
        res = mod.Data(x=out.F, pos=self.xyz, batch=out.C[:, 0])
        if self.has_mlp_head:
            res.x = self.mlp(res.x)
        return res

    def encoder_out(self, out):    # applications/minkowski.py:150-157
        res = mod.Batch(x=out.F, batch=out.C[:, 0].long().to(out.F.device))
        if not isinstance(self.inner_modules[0], mod.Identity):
            res = self.inner_modules[0](res)
        if self.has_mlp_head:
            res.x = self.mlp(res.x)
        return res

    if hasattr(mod, "MinkowskiUnet"):
        wrap(mod.MinkowskiUnet, unet_out)
    if hasattr(mod, "MinkowskiEncoder"):
        wrap(mod.MinkowskiEncoder, encoder_out)


def _patch_meanshift():
    from . import meanshift
    name = "torch_points3d.utils.meanshift_cluster"
    mod = sys.modules.get(name)
    if mod is not None:
        mod.MeanShift = meanshift.MeanShift
        return

    class _Hook:
        """Patches the module right after the reference imports it (one attribute; the file stays unmodified)."""

        def find_spec(self, fullname, path=None, target=None):
            if fullname != name:
                return None
            sys.meta_path.remove(self)
            try:
                spec = importlib.util.find_spec(fullname)
            finally:
                pass
            if spec is None or spec.loader is None:
                return None
            loader = spec.loader
            orig = loader.exec_module

            def exec_module(module):
                orig(module)
                module.MeanShift = meanshift.MeanShift

            loader.exec_module = exec_module
            return spec

    sys.meta_path.insert(0, _Hook())


def enable_sparse_backend(sp3d_nn):
    """`torch_points3d.modules.SparseConv3d.nn.backend_valid` whitelists {"torchsparse", "minkowski"}
    (nn/__init__.py:21-31); a third backend needs its name added there.  With the reference unmodified, call this with
    the imported `torch_points3d.modules.SparseConv3d.nn` module: it binds the six backend symbols directly."""
    from .nn import b200
    for val in b200.__all__:
        setattr(sp3d_nn, val, getattr(b200, val))
    return sp3d_nn
