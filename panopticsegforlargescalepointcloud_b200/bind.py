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


def install(meanshift=True, sparse_backend=True):
    """Register the drop-in modules under the names the reference imports.  Idempotent.
    meanshift: also route `torch_points3d.utils.meanshift_cluster.MeanShift` to the device mean shift when that module is
    (or gets) imported.  sparse_backend: make `SPARSE_BACKEND=b200` / `sp3d.nn.set_backend("b200")` resolvable."""
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
    return me, tpk, hdbscan


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
