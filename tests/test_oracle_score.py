"""CPU checks of the ScoreNet / score-loss restatement (oracle/cpu_path.py) that the GPU parity test leans on:
known answers for the IoU matrix and the soft-target BCE (reference: core/losses/panoptic_losses.py:25-37,92-114)
and the shape contract of the proposal-batched scorer (models/panoptic/PointGroup3heads.py:393-454)."""
import numpy as np
import torch

from oracle import cpu_path


def test_instance_iou_known_answer():
    inst = np.array([1, 1, 1, 2, 2, 0, 0, 1, 1, 2])
    batch = np.array([0, 0, 0, 0, 0, 0, 0, 1, 1, 1])
    clusters = [np.array([0, 1, 3]), np.array([3, 4]), np.array([7, 8, 9]), np.array([5, 6])]
    iou = cpu_path.instance_iou_ref(clusters, inst, batch)
    assert iou.shape == (4, 4)                                   # 2 instances per scene, scenes concatenated
    want = np.zeros((4, 4), np.float32)
    want[0, 0], want[0, 1] = 2 / 4, 1 / 4                        # |{0,1}| / |{0,1,2,3}| ; |{3}| / |{0,1,3,4}|
    want[1, 1] = 1.0
    want[2, 2], want[2, 3] = 2 / 3, 1 / 3
    assert np.allclose(iou, want)
    assert (iou[3] == 0).all()                                   # a proposal of unlabelled points matches nothing


def test_score_loss_known_answer():
    ious = np.array([[0.1, 0.2], [0.5, 0.3], [0.9, 0.0], [0.25, 0.0], [0.75, 0.1]], np.float32)
    scores = torch.tensor([0.2, 0.6, 0.7, 0.4, 0.9])
    target = torch.tensor([0.0, 0.5, 1.0, 0.0, 1.0])            # clip((iou - 0.25) / 0.5, 0, 1) of the row maxima
    want = torch.nn.functional.binary_cross_entropy(scores, target)
    assert abs(float(cpu_path.score_loss_ref(ious, scores)) - float(want)) < 1e-7


def test_score_forward_contract():
    from panopticsegforlargescalepointcloud_b200 import backbone as bb, panoptic, scenes
    torch.manual_seed(0)
    opt = panoptic.paper_options("urban", backbone="two_level")
    sd = {k: v.detach() for k, v in
          panoptic.PointGroup3heads(opt, "dummy", panoptic.DatasetProperties("urban"), None).state_dict().items()}
    s = scenes.make_scene("urban", 1500, 0.2, 2.5, seed=1)
    feats = torch.randn(len(s.pos), 16)
    clusters = [np.arange(0, 40), np.arange(30, 90), np.arange(200, 215)]      # overlapping proposals are legal
    scores, rows = cpu_path.score_forward(sd, cpu_path.resolve_cfg(bb.scorer_unet_config(16), 16), feats, s.coords,
                                          clusters, training=False)
    assert scores.shape == (3,) and rows.shape == (115, 16)
    assert bool(((scores > 0) & (scores < 1)).all())
    # proposals are independent sparse tensors: scoring one alone gives the same score
    alone, _ = cpu_path.score_forward(sd, cpu_path.resolve_cfg(bb.scorer_unet_config(16), 16), feats, s.coords,
                                      clusters[1:2], training=False)
    assert abs(float(alone[0]) - float(scores[1])) < 1e-6
