"""GPU parity at the model level: PointGroup-style models behind the torch_points3d API (set_input / forward /
backward / optimize_parameters2 / get_output) against the CPU restatement (oracle/cpu_path.py, oracle/tpk_ref.py,
oracle/hdbscan_ref.py) with the same state_dict and the same synthetic batch."""
import numpy as np
import pytest
import torch

from oracle import cpu_path, tpk_ref, hdbscan_ref

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _pkg():
    from panopticsegforlargescalepointcloud_b200 import panoptic, scenes
    return panoptic, scenes


def _cpu_batch(b):
    class B:
        pass
    o = B()
    for k in ("x", "coords", "batch", "y", "instance_labels", "instance_mask", "vote_label", "pos"):
        setattr(o, k, torch.as_tensor(getattr(b, k)))
    return o


def _make(kind, n, grid, radius, n_scenes, seed0=0):
    panoptic, scenes = _pkg()
    return scenes.collate([scenes.make_scene(kind, n, grid, radius, seed=seed0 + i) for i in range(n_scenes)])


def test_forward_loss_backward_parity(cuda_device):
    """C1-shaped: 2 scenes x 6k voxels, 2-level U-Net, three heads, eval-mode BN for a well-conditioned check."""
    panoptic, scenes = _pkg()
    batch = _make("urban", 6000, 0.2, 4.0, 2)
    torch.manual_seed(2022)
    opt = panoptic.paper_options("urban", cluster_type=1, grid=0.2, use_score_net=True, prepare_epoch=30,
                                 backbone="two_level")
    model = panoptic.PointGroup3heads(opt, "dummy", panoptic.DatasetProperties("urban"), None).to(cuda_device)
    model.eval()
    model.set_input(batch, cuda_device)
    out = model.forward(epoch=1)
    assert out.clusters is None and out.cluster_scores is None          # epoch <= prepare_epoch: no clustering
    model.backward(1)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()}
    cfg = cpu_path.resolve_cfg(opt.backbone.config, 4)
    loss, sem, off, emb = cpu_path.step_loss(sd, cfg, _cpu_batch(batch), opt.loss_weights, training=False)
    loss.backward()
    assert abs(float(model.loss) - float(loss)) <= TOL * max(1.0, abs(float(loss)))
    assert float((out.semantic_logits.detach().cpu() - sem.detach()).abs().max()) <= TOL * 10  # log-softmax range
    assert float((out.offset_logits.detach().cpu() - off.detach()).abs().max()) <= TOL * max(1.0, float(off.abs().max()))
    assert float((out.embed_logits.detach().cpu() - emb.detach()).abs().max()) <= TOL * max(1.0, float(emb.abs().max()))
    ga = torch.cat([p.grad.cpu().reshape(-1) for n, p in model.named_parameters() if p.grad is not None]).double()
    gb = torch.cat([sd[n].grad.reshape(-1) for n, p in model.named_parameters() if p.grad is not None]).double()
    cos = float(torch.dot(ga, gb) / (ga.norm() * gb.norm()))
    assert cos >= 0.9999, cos
    losses = model.get_current_losses()
    assert set(["loss", "semantic_loss", "offset_norm_loss", "offset_dir_loss", "ins_loss"]) <= set(losses)


def test_cluster_and_score_path(cuda_device):
    """epoch > prepare_epoch: shifted-coordinate region growing + ScorerUnet; clusters equal the oracle's on the
    model's own (untrained) head outputs, scores are finite probabilities, the score loss back-propagates."""
    panoptic, scenes = _pkg()
    batch = _make("urban", 8000, 0.2, 5.0, 2, seed0=3)
    torch.manual_seed(2022)
    opt = panoptic.paper_options("urban", cluster_type=2, grid=0.2, prepare_epoch=-1, backbone="two_level")
    model = panoptic.PointGroup3heads(opt, "dummy", panoptic.DatasetProperties("urban"), None).to(cuda_device)
    model.instantiate_optimizers({})
    model.train()
    model.set_input(batch, cuda_device)
    with torch.no_grad():
        # steer the untrained net towards sensible clusters: bias the semantic head with the labels
        pass
    model.optimize_parameters2(epoch=31, step=0, batch_size=2)
    out = model.get_output()
    pred = out.semantic_logits.argmax(1).cpu().numpy()
    ignore = [-1] + list(scenes.stuff_classes("urban"))
    shifted = (model.raw_pos + out.offset_logits.detach()).cpu().numpy()
    want = tpk_ref.region_grow(batch.pos, pred, batch.batch, ignore, 300, 0.3, 10, method="grid") + \
        tpk_ref.region_grow(shifted, pred, batch.batch, ignore, 200, 0.3, 10, method="grid")
    got = out.clusters or []
    assert [tuple(c.cpu().tolist()) for c in got] == tpk_ref.partition_key(want)
    if got:
        assert out.cluster_scores.shape[0] == len(got)
        assert bool(((out.cluster_scores >= 0) & (out.cluster_scores <= 1)).all())
        assert out.cluster_type.shape[0] == len(got)
        assert np.isfinite(model.get_current_losses()["score_loss"])
        assert any(p.grad is not None and float(p.grad.abs().sum()) > 0 for p in model.ScorerUnet.parameters())


def test_region_grow_on_synthetic_heads_matches_oracle_and_pq(cuda_device):
    """Matched PQ: clusters from synthetic 'trained' head outputs, product vs oracle partitions equal
    => identical PQ against the synthetic ground truth."""
    panoptic, scenes = _pkg()
    from panopticsegforlargescalepointcloud_b200 import tpk, metrics
    s = scenes.make_scene("urban", 30000, 0.2, 7.0, seed=5)
    off, emb, logits = scenes.synthetic_head_outputs(s, seed=5)
    pred = logits.argmax(1)
    ignore = [-1] + list(scenes.stuff_classes("urban"))
    shifted = (s.pos + off).astype(np.float32)
    want = tpk_ref.region_grow(shifted, pred, s.batch, ignore, 200, 0.3, 10, method="grid")
    got = tpk.region_grow(torch.from_numpy(shifted).to(cuda_device), torch.from_numpy(pred).to(cuda_device),
                          torch.from_numpy(s.batch).to(cuda_device), ignore_labels=ignore, nsample=200, radius=0.3,
                          min_cluster_size=10)
    got = [c.cpu().numpy() for c in got]
    assert [tuple(c.tolist()) for c in got] == tpk_ref.partition_key(want)
    pq_a = metrics.panoptic_quality(pred, got, s.y, s.instance_labels, scenes.num_classes("urban"), list(scenes.URBAN_THINGS))
    pq_b = metrics.panoptic_quality(pred, [np.sort(c) for c in want], s.y, s.instance_labels, scenes.num_classes("urban"),
                                    list(scenes.URBAN_THINGS))
    assert pq_a == pq_b and pq_a["PQ"] > 0.5


def test_embed_model_hdbscan_path(cuda_device):
    """PointGroupEmbed cluster_type 14 (pointgroupembed.py:683-710): HDBSCAN over the embedding head output, per
    scene; equal to the oracle's cluster_single on the same embeddings."""
    panoptic, scenes = _pkg()
    batch = _make("forest", 3000, 0.2, 4.0, 2, seed0=8)
    torch.manual_seed(2022)
    opt = panoptic.paper_options("forest", cluster_type=14, grid=0.2, use_score_net=False, scorer=False,
                                 backbone="two_level")
    model = panoptic.PointGroupEmbed(opt, "dummy", panoptic.DatasetProperties("forest"), None).to(cuda_device)
    model.eval()
    model.set_input(batch, cuda_device)
    out = model.forward(epoch=1)
    emb = out.embed_logits.detach().cpu().numpy()
    pred = out.semantic_logits.argmax(1).cpu().numpy()
    mask = ~np.isin(pred, [-1] + list(scenes.stuff_classes("forest")))
    local = np.nonzero(mask)[0]
    lb = batch.batch[mask]
    want, _ = hdbscan_ref.cluster_single(emb[mask], np.unique(lb), lb, local, 0)
    got = out.clusters
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert np.array_equal(a.cpu().numpy(), b)


@pytest.mark.parametrize("cls,ct", [("PointGroup3heads", 5), ("PointGroup3heads", 6), ("PointGroupEmbed", 7)])
def test_meanshift_cluster_types(cuda_device, cls, ct):
    """Paper settings IV / V / I (PointGroup3heads._cluster5 / _cluster6, pointgroupembed._cluster7): region growing on
    raw / shifted positions plus mean shift on the embedding head output; the embedding branch equals scikit-learn's
    MeanShift(bandwidth=0.6, bin_seeding=True) run per scene on the same embeddings, cluster types as the reference
    numbers them."""
    from sklearn.cluster import MeanShift
    panoptic, scenes = _pkg()
    batch = _make("forest", 2500, 0.2, 4.0, 2, seed0=3)
    torch.manual_seed(2022)
    opt = panoptic.paper_options("forest", cluster_type=ct, grid=0.2, use_score_net=False, scorer=False, backbone="two_level")
    model = getattr(panoptic, cls)(opt, "dummy", panoptic.DatasetProperties("forest"), None).to(cuda_device)
    model.eval()
    model.set_input(batch, cuda_device)
    out = model.forward(epoch=1)
    emb = out.embed_logits.detach().cpu().numpy()
    pred = out.semantic_logits.argmax(1).cpu().numpy()
    mask = ~np.isin(pred, [-1] + list(scenes.stuff_classes("forest")))
    local = np.nonzero(mask)[0]
    lb = np.asarray(batch.batch)[mask]
    want = []
    for s in np.unique(lb):
        m = lb == s
        if m.sum() > 3:
            lab = MeanShift(bandwidth=0.6, bin_seeding=True).fit(emb[mask][m]).labels_
            want += [local[m][lab == l] for l in np.unique(lab)]
    types = out.cluster_type.cpu().numpy()
    ms_type = {5: 1, 6: 2, 7: 0}[ct]
    n_rg = len(out.clusters) - len(want)
    assert n_rg >= 0 and np.all(types[n_rg:] == ms_type)
    if ct == 6:
        assert set(types[:n_rg].tolist()) <= {0, 1}
    for a, b in zip(out.clusters[n_rg:], want):
        assert np.array_equal(a.cpu().numpy(), b)


def test_scorenet_and_score_loss_parity(cuda_device):
    """a9 (PointGroup3heads.py:393-454 `_compute_score`, panoptic_losses.py:25-37,92-114): the product batches all
    proposals into one second sparse tensor (batch id = proposal id, stride-2 first conv on a fresh coordinate hash);
    the oracle restates the reference's per-proposal loops.  Same proposals (region growing on synthetic head outputs,
    itself covered above), same weights: scores, IoU matrix and score loss must agree, and so must the gradient the
    score loss sends into ScorerUnet."""
    panoptic, scenes = _pkg()
    from panopticsegforlargescalepointcloud_b200 import tpk, backbone as bb
    sc = [scenes.make_scene("urban", 8000, 0.2, 5.0, seed=21 + i) for i in range(2)]
    batch = scenes.collate(sc)
    heads = [scenes.synthetic_head_outputs(s, seed=21 + i) for i, s in enumerate(sc)]
    shifted = np.concatenate([s.pos + h[0] for s, h in zip(sc, heads)]).astype(np.float32)
    pred = np.concatenate([h[2].argmax(1) for h in heads]).astype(np.int64)
    ignore = [-1] + list(scenes.stuff_classes("urban"))
    props = tpk.region_grow(torch.from_numpy(shifted).to(cuda_device), torch.from_numpy(pred).to(cuda_device),
                            torch.as_tensor(batch.batch).to(cuda_device), ignore_labels=ignore, nsample=200, radius=0.3,
                            min_cluster_size=10)
    assert len(props) >= 10
    torch.manual_seed(2022)
    opt = panoptic.paper_options("urban", cluster_type=1, grid=0.2, prepare_epoch=-1, backbone="two_level")
    model = panoptic.PointGroup3heads(opt, "dummy", panoptic.DatasetProperties("urban"), None).to(cuda_device)
    model.eval()                                                   # running statistics: a well-conditioned comparison
    model._do_cluster = lambda sem, off, emb: (props, torch.zeros(len(props), dtype=torch.uint8, device=cuda_device))
    model.set_input(batch, cuda_device)
    out = model.forward(epoch=31)
    model.zero_grad()
    model.backward(31)
    assert out.cluster_scores is not None and out.cluster_scores.shape[0] == len(props)

    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()}
    cb = _cpu_batch(batch)
    coords4 = np.concatenate([np.asarray(batch.batch)[:, None], np.asarray(batch.coords)], 1).astype(np.int32)
    feats = cpu_path.unet_forward(sd, cpu_path.resolve_cfg(opt.backbone.config, 4), cb.x, coords4, training=False,
                                  prefix="Backbone.")
    clusters = [c.cpu().numpy() for c in props]
    scores, _ = cpu_path.score_forward(sd, cpu_path.resolve_cfg(bb.scorer_unet_config(16), 16), feats,
                                       np.asarray(batch.coords), clusters, training=False)
    assert float((out.cluster_scores.detach().cpu() - scores.detach()).abs().max()) <= TOL
    ious = cpu_path.instance_iou_ref(clusters, np.asarray(batch.instance_labels), np.asarray(batch.batch))
    got_iou = tpk.instance_iou(props, model.input.instance_labels, model.input.batch).cpu().numpy()
    assert got_iou.shape == ious.shape and np.allclose(got_iou, ious, rtol=0, atol=1e-7)   # integer counts, one division
    assert ious.max(1).mean() > 0.3                                                 # the proposals do hit instances
    sl = cpu_path.score_loss_ref(ious, scores)
    assert abs(float(model.score_loss) - float(sl)) <= TOL * max(1.0, abs(float(sl)))
    sl.backward()
    ga = torch.cat([p.grad.cpu().reshape(-1) for n, p in model.ScorerUnet.named_parameters()]).double()
    gb = torch.cat([sd["ScorerUnet." + n].grad.reshape(-1) for n, p in model.ScorerUnet.named_parameters()]).double()
    # model.backward adds the other loss terms too, but only the score loss reaches ScorerUnet
    assert float((ga - gb).norm() / gb.norm()) <= 1e-3
