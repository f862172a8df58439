"""The drop-in boundary, proven with the reference's OWN files (CPU part: imports, factories, parameter layout).

baseline/_ref holds the unmodified reference package (scripts/stage_reference.py); tests/ref_binding.py calls the
product's `bind.install()` and stands in for the other absent third-party imports.  No kernels run here: the GPU half is
tests/test_gpu_reference_binding.py.
"""
import importlib
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
import ref_binding as rb  # noqa: E402

pytestmark = pytest.mark.skipif(not rb.available(), reason="baseline/_ref not staged (scripts/stage_reference.py)")


@pytest.fixture(scope="module", autouse=True)
def _bound():
    rb.install()


def test_reference_hot_path_files_import_over_the_aliases():
    from panopticsegforlargescalepointcloud_b200 import me, tpk, hdbscan, meanshift
    mods = {}
    for name in ("modules.MinkowskiEngine.api_modules", "modules.MinkowskiEngine", "applications.minkowski",
                 "models.panoptic.PointGroup3heads", "models.panoptic.pointgroup", "models.panoptic.pointgroupembed",
                 "models.panoptic.structure_3heads", "models.model_factory", "models.base_model", "utils.hdbscan_cluster",
                 "utils.meanshift_cluster", "core.schedulers.bn_schedulers", "core.losses.panoptic_losses",
                 "core.spatial_ops.neighbour_finder", "modules.SparseConv3d.nn", "applications.sparseconv3d"):
        mods[name] = importlib.import_module("torch_points3d." + name)
        assert mods[name].__file__.startswith(rb.REF)          # the reference's file, not a mirror
    api = mods["modules.MinkowskiEngine.api_modules"]
    assert api.ME is me and api.ResNetDown.CONVOLUTION is me.MinkowskiConvolution
    assert api.ResNetUp.CONVOLUTION is me.MinkowskiConvolutionTranspose
    assert mods["models.panoptic.PointGroup3heads"].region_grow is tpk.region_grow
    assert mods["core.losses.panoptic_losses"].instance_iou is tpk.instance_iou
    assert mods["core.spatial_ops.neighbour_finder"].tp is tpk
    assert mods["utils.hdbscan_cluster"].hdbscan is hdbscan
    assert mods["utils.meanshift_cluster"].MeanShift is meanshift.MeanShift
    # the stock model zoo the reference imports alongside (res16unet.py:5, resunet.py:3) resolves too
    assert hasattr(mods["modules.MinkowskiEngine"], "Res16UNet34")
    # BN-momentum scheduler: our BatchNorm is in its module tuple (core/schedulers/bn_schedulers.py:7-17)
    assert me.MinkowskiBatchNorm in mods["core.schedulers.bn_schedulers"].BATCH_NORM_MODULES


def test_reference_backbone_factory_builds_the_same_network():
    """`Minkowski("unet", ...)` of applications/minkowski.py with the reference's ResNetDown / ResNetUp / ResBlock over
    me.py: same parameter names, shapes and -- under the same seed -- the same initial values as the package's mirror
    (backbone.py), i.e. checkpoints are interchangeable."""
    from torch_points3d.applications.minkowski import Minkowski
    from panopticsegforlargescalepointcloud_b200 import backbone as bb, me
    torch.manual_seed(7)
    ref = Minkowski("unet", input_nc=4, num_layers=4, config=rb.config(bb.paper_backbone_config(16)))
    torch.manual_seed(7)
    mine = bb.Minkowski("unet", input_nc=4, config=bb.paper_backbone_config(16))
    a, b = ref.state_dict(), mine.state_dict()
    assert list(a) == list(b) and len(a) == 492
    assert all(torch.equal(a[k], b[k]) for k in a)
    convs = [m for m in ref.modules() if isinstance(m, me.MinkowskiConvolutionBase)]
    assert len(convs) == 82 and sum(isinstance(m, me.MinkowskiConvolutionTranspose) for m in convs) == 41
    assert sum(p.numel() for p in ref.parameters()) == 10413696
    assert type(ref).__module__ == "torch_points3d.applications.minkowski"


def test_reference_model_factory_instantiates_the_shipped_config():
    """models/model_factory.py:8-45 on conf/models/panoptic/area4_ablation_3heads_5.yaml (paper setting IV): class
    lookup by lower-cased name, resolve_model's eval of "2*in_feat" / "1.5 * 0.12", constructor
    (option, "dummy", dataset, modules) -- all the reference's code, none of ours above the three aliases."""
    from torch_points3d.models.model_factory import instantiate_model
    from panopticsegforlargescalepointcloud_b200 import panoptic, me
    cfg = rb.load_run_config("area4_ablation_3heads_5.yaml", "PointGroup-PAPER", 0.12)
    torch.manual_seed(2022)
    model = instantiate_model(cfg, rb.DatasetStub("urban"))
    assert type(model).__module__ == "torch_points3d.models.panoptic.PointGroup3heads"
    assert model.conv_type == "SPARSE" and abs(model.opt.cluster_radius_search - 0.18) < 1e-12
    assert model.opt.cluster_type == 5 and model.get_opt_mergeTh() == 0.01
    assert isinstance(model.Backbone.down_modules[0].conv_in[0], me.MinkowskiConvolution)
    torch.manual_seed(2022)
    mine = panoptic.PointGroup3heads(panoptic.paper_options("urban", cluster_type=5, grid=0.12), "dummy",
                                     panoptic.DatasetProperties("urban"), None)
    a, b = model.state_dict(), mine.state_dict()
    assert set(b) <= set(a)                                   # the reference also builds ScorerEncoder / ScorerMLP (unused)
    assert {k.split(".")[0] for k in set(a) - set(b)} == {"ScorerEncoder", "ScorerMLP"}
    assert all(a[k].shape == b[k].shape for k in b)
    missing, unexpected = mine.load_state_dict(a, strict=False)
    assert not missing and {k.split(".")[0] for k in unexpected} == {"ScorerEncoder", "ScorerMLP"}


def test_sparse_backend_plugin_hook():
    """modules/SparseConv3d/nn/__init__.py:21-52: the stock "minkowski" backend resolves to this library through the alias;
    the `b200` backend module carries the same six symbols and binds without editing the reference's whitelist."""
    import torch_points3d.modules.SparseConv3d.nn as snn
    from panopticsegforlargescalepointcloud_b200 import bind, me
    from panopticsegforlargescalepointcloud_b200.nn import b200
    snn.set_backend("minkowski")
    assert snn.get_backend() == "minkowski" and issubclass(snn.Conv3d, me.MinkowskiConvolution)
    assert issubclass(snn.BatchNorm, me.MinkowskiBatchNorm) and issubclass(snn.Conv3dTranspose, me.MinkowskiConvolutionTranspose)
    assert sorted(b200.__all__) == sorted(snn.__all__)
    assert importlib.import_module("torch_points3d.modules.SparseConv3d.nn.b200") is b200
    bind.enable_sparse_backend(snn)
    assert snn.Conv3d is b200.Conv3d and snn.SparseTensor is b200.SparseTensor and snn.cat is b200.cat
    c = snn.Conv3d(16, 32, kernel_size=3, stride=2)
    assert tuple(c.kernel.shape) == (27, 16, 32) and c.bias is None and c.stride == 2
    ct = snn.Conv3dTranspose(32, 16)
    assert ct.TRANSPOSE and tuple(ct.kernel.shape) == (27, 32, 16)
    assert repr(snn.BatchNorm(8)).startswith("BatchNorm1d(8")
    # the reference's backend-switchable factory (applications/sparseconv3d.py) builds on those symbols
    from torch_points3d.applications.sparseconv3d import SparseConv3d
    net = SparseConv3d("unet", input_nc=4, num_layers=4, config=rb.config({
        "define_constants": {"in_feat": 16},
        "down_conv": {"module_name": "ResNetDown", "block": "ResBlock", "N": [1, 1], "kernel_size": [3, 3], "stride": [1, 2],
                      "down_conv_nn": [["FEAT", "in_feat"], ["in_feat", "2*in_feat"]]},
        "up_conv": {"module_name": "ResNetUp", "block": "ResBlock", "N": [1, 1], "kernel_size": [3, 3], "stride": [2, 1],
                    "up_conv_nn": [["2*in_feat", "in_feat"], ["2*in_feat", "in_feat"]]}}), backend="minkowski")
    assert any(isinstance(m, me.MinkowskiConvolutionBase) for m in net.modules())
