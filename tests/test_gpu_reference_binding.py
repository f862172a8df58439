"""The drop-in boundary on the GPU: the reference's OWN classes and functions (baseline/_ref, unmodified) run on
libpgs_b200.so through `bind.install()`, and agree with the package's device-resident mirrors and with the CPU oracle.

  applications/minkowski.py MinkowskiUnet + modules/MinkowskiEngine/api_modules.py ResNetDown / ResNetUp / ResBlock
  utils/hdbscan_cluster.py cluster_single, utils/meanshift_cluster.py cluster_single
  models/panoptic/structure_3heads.py PanopticResults.get_instances
  models/model_factory.py instantiate_model -> models/panoptic/PointGroup3heads.py set_input / forward / backward
"""
import multiprocessing
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
import ref_binding as rb  # noqa: E402
from oracle import cpu_path, hdbscan_ref  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not rb.available(), reason="baseline/_ref not staged")]
TOL = 1e-4


@pytest.fixture(scope="module", autouse=True)
def _bound():
    rb.install()


def _data(batch, dev, keys=("pos", "coords", "x", "batch", "y", "instance_labels", "instance_mask", "vote_label",
                            "center_label", "num_instances")):
    from torch_geometric.data import Batch          # (the stand-in of tests/ref_binding.py)
    return Batch(**{k: torch.as_tensor(getattr(batch, k)).to(dev) for k in keys})


def _scenes(kind, n, grid, radius, count, seed0):
    from panopticsegforlargescalepointcloud_b200 import scenes
    return scenes.collate([scenes.make_scene(kind, n, grid, radius, seed=seed0 + i) for i in range(count)])


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("which", ["two_level", "paper"])
def test_reference_unet_runs_on_the_kernels(cuda_device, which, fused, monkeypatch):
    """The reference's MinkowskiUnet over me.py -- fused=False: its own forward loop, skip stack and blocks, module by
    module; fused=True: the same object through bind.install(fuse_unet=True), i.e. the tape compiled from ITS module tree
    and walked by pgs_unet_forward / pgs_unet_backward -- == the package's mirror with the same state_dict (same kernels:
    2e-5), == the CPU oracle (1e-4); parameter gradients agree."""
    from torch_points3d.applications.minkowski import Minkowski, MinkowskiUnet
    from panopticsegforlargescalepointcloud_b200 import backbone as bb, _lib, fastpath
    assert getattr(MinkowskiUnet.forward, "_pgs_fused", False)       # bind.install() wrapped the reference class
    cfg = bb.two_level_config(16) if which == "two_level" else bb.paper_backbone_config(16)
    torch.manual_seed(3)
    ref = Minkowski("unet", input_nc=4, num_layers=4, config=rb.config(cfg)).to(cuda_device)
    mine = bb.Minkowski("unet", input_nc=4, config=cfg).to(cuda_device)
    mine.load_state_dict(ref.state_dict())
    ref.eval(); mine.eval()
    b = _scenes("urban", 9000, 0.2, 5.0, 2, 40)
    assert type(ref).__module__ == "torch_points3d.applications.minkowski" and fastpath.program_for(ref) is not None
    before = _lib.launch_count()
    monkeypatch.setattr(fastpath, "ENABLED", fused)
    out_ref = ref(_data(b, cuda_device, ("pos", "coords", "x", "batch")))
    monkeypatch.setattr(fastpath, "ENABLED", True)
    assert _lib.launch_count() > before                              # the reference's modules launched OUR kernels
    assert (out_ref.x.grad_fn is not None) and (("UNetFn" in type(out_ref.x.grad_fn).__name__) == fused)
    out_mine = mine(_data(b, cuda_device, ("pos", "coords", "x", "batch")))
    scale = max(1.0, float(out_mine.x.abs().max()))
    assert float((out_ref.x - out_mine.x).abs().max()) <= 2e-5 * scale
    assert torch.equal(out_ref.batch.long(), torch.as_tensor(b.batch).to(cuda_device).long())
    g = torch.randn_like(out_ref.x)
    out_ref.x.backward(g)
    out_mine.x.backward(g)
    # same kernels on both sides; what differs is the order of the fp32 atomic adds (weight gradients, few-row layers),
    # which the 3..50-row coarse levels of this small scene amplify -- hence a per-tensor bound plus a tight global one
    for (n1, p1), (n2, p2) in zip(ref.named_parameters(), mine.named_parameters()):
        assert n1 == n2
        assert float((p1.grad - p2.grad).norm()) <= 2e-2 * max(float(p2.grad.norm()), 1e-6), n1
    ga = torch.cat([p.grad.reshape(-1) for p in ref.parameters()]).double()
    gb = torch.cat([p.grad.reshape(-1) for p in mine.parameters()]).double()
    assert 1.0 - float(torch.dot(ga, gb) / (ga.norm() * gb.norm())) <= 1e-6
    sd = {k: v.detach().cpu() for k, v in ref.state_dict().items()}
    coords4 = np.concatenate([np.asarray(b.batch)[:, None], np.asarray(b.coords)], 1).astype(np.int32)
    want = cpu_path.unet_forward(sd, cpu_path.resolve_cfg(cfg, 4), torch.as_tensor(b.x), coords4, training=False)
    assert float((out_ref.x.detach().cpu() - want).abs().max()) <= TOL * max(1.0, float(want.abs().max()))


def test_reference_cluster_single_functions(cuda_device, monkeypatch):
    """utils/hdbscan_cluster.py:117-167 and utils/meanshift_cluster.py:72-123, the reference's own fan-out code, with
    `hdbscan.HDBSCAN` / `MeanShift` bound to the device implementations (numpy in, numpy out -- what these files pass)."""
    monkeypatch.setattr(multiprocessing, "Pool", rb.SerialPool)
    from torch_points3d.utils import hdbscan_cluster as ref_h, meanshift_cluster as ref_m
    from sklearn.cluster import MeanShift as SkMeanShift
    rng = np.random.default_rng(5)
    mu = rng.normal(0, 3, (9, 5))
    X = (mu[rng.integers(0, 9, 2400)] + rng.normal(0, 0.15, (2400, 5))).astype(np.float32)
    batch = np.sort(rng.integers(0, 2, 2400))
    local = np.arange(5000, 7400)
    args = (torch.from_numpy(X).to(cuda_device), torch.tensor([0, 1], device=cuda_device),
            torch.from_numpy(batch).to(cuda_device), torch.from_numpy(local).to(cuda_device))
    got, types = ref_h.cluster_single(*args, 3)
    want, wtypes = hdbscan_ref.cluster_single(X, [0, 1], batch, local, 3)
    assert types == wtypes and len(got) == len(want) >= 8
    for a, b in zip(got, want):
        assert np.array_equal(a.numpy(), b)
    got, types = ref_m.cluster_single(*args, 2, 0.6)
    want = []
    for s in (0, 1):
        lab = SkMeanShift(bandwidth=0.6, bin_seeding=True).fit(X[batch == s]).labels_
        want += [local[batch == s][lab == l] for l in np.unique(lab)]
    assert types == [2] * len(want) and len(got) == len(want)
    for a, b in zip(got, want):
        assert np.array_equal(a.numpy(), b)


def test_reference_get_instances(cuda_device):
    """structure_3heads.py:28-71 (dense-mask NMS, hard-coded .cuda()) == the package's sparse-incidence version."""
    from torch_points3d.models.panoptic.structure_3heads import PanopticResults as RefResults
    from panopticsegforlargescalepointcloud_b200 import panoptic
    rng = np.random.default_rng(2)
    n = 5000
    clusters = []
    for i in range(40):
        c0 = int(rng.integers(0, n - 400))
        clusters.append(torch.from_numpy(np.unique(rng.integers(c0, c0 + 400, int(rng.integers(20, 300))))).to(cuda_device))
    scores = torch.from_numpy(rng.random(40).astype(np.float32)).to(cuda_device)
    kw = dict(semantic_logits=torch.zeros(n, 9, device=cuda_device), offset_logits=None, embed_logits=None,
              clusters=clusters, cluster_scores=scores, mask_scores=None, cluster_type=None)
    a = RefResults(**kw).get_instances(nms_threshold=0.3, min_cluster_points=30, min_score=0.2)
    b = panoptic.PanopticResults(**kw).get_instances(nms_threshold=0.3, min_cluster_points=30, min_score=0.2)
    ids_a = a[0].tolist() if torch.is_tensor(a[0]) else list(a[0])
    assert ids_a == list(b[0]) and len(ids_a) >= 3
    for x, y in zip(a[1], b[1]):
        assert torch.equal(x, y)


def test_reference_model_step_through_its_own_factory(cuda_device, monkeypatch):
    """train.py's inner loop on the reference's PointGroup3heads, built by the reference's instantiate_model from the
    shipped YAML (paper setting IV: cluster_type 5 = region_grow on shifted xyz + MeanShift on embeddings + ScoreNet):
    set_input -> forward(epoch > prepare_epoch) -> backward.  Compared with the package's mirror carrying the same
    weights: per-point heads, proposals (as sets), proposal scores, every loss term."""
    monkeypatch.setattr(multiprocessing, "Pool", rb.SerialPool)
    from torch_points3d.models.model_factory import instantiate_model
    from panopticsegforlargescalepointcloud_b200 import panoptic, _lib
    grid = 0.2
    cfg = rb.load_run_config("area4_ablation_3heads_5.yaml", "PointGroup-PAPER", grid)
    torch.manual_seed(2022)
    ref = instantiate_model(cfg, rb.DatasetStub("urban")).to(cuda_device)
    mine = panoptic.PointGroup3heads(panoptic.paper_options("urban", cluster_type=5, grid=grid), "dummy",
                                     panoptic.DatasetProperties("urban"), None).to(cuda_device)
    mine.load_state_dict(ref.state_dict(), strict=False)
    with torch.no_grad():      # an untrained net predicts stuff everywhere and proposes nothing: predict a thing class
        for m in (ref, mine):
            m.Semantic[1].weight.mul_(0.01)
            m.Semantic[1].bias.zero_()
            m.Semantic[1].bias[2] = 3.0
    ref.train(); mine.train()
    b = _scenes("urban", 7000, grid, 4.5, 2, 60)
    data = _data(b, cuda_device)
    before = _lib.launch_count()
    ref.set_input(data, cuda_device)
    ref.forward(epoch=31, step=0, is_training=True)
    ref.backward(31)
    assert _lib.launch_count() - before > 300
    mine.set_input(b, cuda_device)
    mine.forward(epoch=31, step=0, is_training=True)
    mine.backward(31)
    ro, mo = ref.get_output(), mine.get_output()
    for name in ("semantic_logits", "offset_logits", "embed_logits"):
        x, y = getattr(ro, name).detach(), getattr(mo, name).detach()
        assert float((x - y).abs().max()) <= 5e-5 * max(1.0, float(y.abs().max())), name
    rc = sorted(tuple(sorted(c.tolist())) for c in (ro.clusters or []))
    mc = sorted(tuple(sorted(c.tolist())) for c in (mo.clusters or []))
    assert rc == mc
    if rc:
        order_r = sorted(range(len(rc)), key=lambda i: tuple(sorted(ro.clusters[i].tolist())))
        order_m = sorted(range(len(mc)), key=lambda i: tuple(sorted(mo.clusters[i].tolist())))
        sr_, sm_ = ro.cluster_scores.detach()[order_r], mo.cluster_scores.detach()[order_m]
        assert float((sr_ - sm_).abs().max()) <= TOL
    for name in ("loss", "semantic_loss", "offset_norm_loss", "offset_dir_loss", "ins_loss", "score_loss"):
        if hasattr(ref, name) and hasattr(mine, name):
            x, y = float(getattr(ref, name)), float(getattr(mine, name))
            assert abs(x - y) <= TOL * max(1.0, abs(y)), (name, x, y)
    assert ref.get_current_losses().keys() >= {"loss", "semantic_loss"}
