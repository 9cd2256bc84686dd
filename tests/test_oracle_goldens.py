"""The oracle against the golden vectors and known-answer tests the reference's own test
suite holds for this path.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import erd_oracle as O


def test_atss_four_priors_golden():
    # reference tests/test_models/test_task_modules/test_assigners/test_atss_assigner.py:12-36
    priors = torch.FloatTensor([[0, 0, 10, 10], [10, 10, 20, 20], [5, 5, 15, 15], [32, 32, 38, 42]])
    gt = torch.FloatTensor([[0, 0, 10, 9], [0, 10, 10, 19]])
    labels = torch.LongTensor([2, 3])
    gt_inds, _, lab = O.atss_assign(priors, [4], gt, labels)
    assert gt_inds.tolist() == [1, 0, 0, 0]
    assert lab.tolist() == [2, -1, -1, -1]


def test_atss_empty_gt_and_empty_priors():
    # test_atss_assigner.py:68-147
    priors = torch.FloatTensor([[0, 0, 10, 10], [10, 10, 20, 20], [5, 5, 15, 15], [32, 32, 38, 42]])
    gt_inds, mo, lab = O.atss_assign(priors, [4], torch.empty(0, 4), torch.empty(0, dtype=torch.long))
    assert gt_inds.tolist() == [0, 0, 0, 0] and lab.tolist() == [-1] * 4 and mo.tolist() == [0.0] * 4
    gt = torch.FloatTensor([[0, 0, 10, 9], [0, 10, 10, 19]])
    gt_inds, _, _ = O.atss_assign(torch.empty(0, 4), [0], gt, torch.LongTensor([2, 3]))
    assert gt_inds.numel() == 0
    gt_inds, _, _ = O.atss_assign(torch.empty(0, 4), [0], torch.empty(0, 4), torch.empty(0, dtype=torch.long))
    assert gt_inds.numel() == 0


def test_giou_golden():
    # tests/test_models/test_task_modules/test_iou2d_calculator.py:84-99
    b1 = torch.FloatTensor([[0, 0, 10, 10], [10, 10, 20, 20], [32, 32, 38, 42]])
    b2 = torch.FloatTensor([[0, 0, 10, 20], [0, 10, 10, 19], [10, 10, 20, 20]])
    g = O.aligned_iou(b1, b2, giou=True, eps=1e-7).numpy().round(4)
    assert np.allclose(g, np.array([0.5000, -0.0500, -0.8214]), rtol=0, atol=1e-7)


def test_iou_shapes_and_range():
    # test_iou2d_calculator.py:11-82 (shape / range contracts)
    a, b = torch.rand(5, 4), torch.rand(7, 4)
    a[:, 2:] += a[:, :2]
    b[:, 2:] += b[:, :2]
    iou = O.pairwise_iou(a, b)
    assert iou.shape == (5, 7) and float(iou.min()) >= 0 and float(iou.max()) <= 1
    assert O.pairwise_iou(torch.empty(0, 4), b).shape == (0, 7)
    assert O.pairwise_iou(a, torch.empty(0, 4)).shape == (5, 0)


def test_qfl_tuple_target_equals_soft_one_hot():
    # tests/test_models/test_losses/test_loss.py:50-67
    pred = torch.rand(6, 4)
    label = torch.LongTensor([0, 3, 4, 1, 2, 4])       # 4 == background
    score = torch.rand(6)
    soft = torch.zeros(6, 4)
    for i, (l, s) in enumerate(zip(label.tolist(), score.tolist())):
        if l < 4:
            soft[i, l] = s
    sig = pred.sigmoid()
    ref = (torch.nn.functional.binary_cross_entropy_with_logits(pred, soft, reduction='none')
           * (soft - sig).abs().pow(2)).sum(1)
    assert torch.allclose(O.qfl_elementwise(pred, label, score), ref, atol=1e-6)


def test_zero_weight_losses_are_zero():
    # test_loss.py:18-27 / 70-107: weight 0 -> loss 0; avg_factor reduction contract
    loss = torch.rand(8)
    assert float(O.reduce_with_avg(loss, torch.zeros(8), 1.0)) == 0.0
    assert torch.isclose(O.reduce_with_avg(loss, torch.ones(8), 4.0), loss.sum() / (4.0 + O.EPS32))


def _head_losses(gt_boxes, gt_labels, seed=0):
    """GFL-head property test of tests/test_models/test_dense_heads/test_gfl_head.py:14-89, on the
    increment head's GT losses: 256x256 image, 4 new classes."""
    g = torch.Generator().manual_seed(seed)
    s = 256
    sizes = [(s // f, s // f) for f in [8, 16, 32, 64, 128]]
    ori, C = 2, 6
    s_cls = [torch.rand(1, C, h, w, generator=g) for h, w in sizes]
    s_box = [torch.rand(1, 68, h, w, generator=g) for h, w in sizes]
    t_cls = [torch.rand(1, ori, h, w, generator=g) for h, w in sizes]
    t_box = [torch.rand(1, 68, h, w, generator=g) for h, w in sizes]
    ci, bi = O.sel_pos(t_cls, t_box)
    return O.loss_by_feat(t_cls, t_box, s_cls, s_box, ci, bi, ori, 1.0, [gt_boxes], [gt_labels], [(s, s)],
                          num_classes=C)


def test_head_empty_gt_properties():
    out = _head_losses(torch.empty(0, 4), torch.empty(0, dtype=torch.long))
    assert float(sum(out['loss_cls'])) > 0
    assert float(sum(out['loss_bbox'])) == 0 and float(sum(out['loss_dfl'])) == 0


def test_head_one_gt_properties():
    out = _head_losses(torch.Tensor([[23.6667, 23.8757, 238.6326, 151.8874]]), torch.LongTensor([2]))
    assert float(sum(out['loss_cls'])) > 0 and float(sum(out['loss_bbox'])) > 0 and float(sum(out['loss_dfl'])) > 0


def test_no_valid_anchor_raises():
    # gfl_head.py:613-617
    anchors = O.level_anchors(2, 2, 8)
    with pytest.raises(ValueError):
        O.image_targets(anchors, torch.zeros(4, dtype=torch.bool), [4], torch.empty(0, 4),
                        torch.empty(0, dtype=torch.long), 80)


def test_nms_matches_torchvision_and_class_offsets():
    tv = pytest.importorskip('torchvision')
    g = torch.Generator().manual_seed(5)
    xy = torch.rand(300, 2, generator=g) * 100
    wh = torch.rand(300, 2, generator=g) * 30 + 1
    boxes = torch.cat([xy, xy + wh], 1)
    scores = torch.rand(300, generator=g)
    ids = torch.randint(0, 5, (300,), generator=g)
    for thr in (0.005, 0.3, 0.6):
        assert torch.equal(O.nms(boxes, scores, thr), tv.ops.nms(boxes, scores, thr))
        _, keep = O.batched_nms(boxes, scores, ids, dict(iou_threshold=thr))
        assert torch.equal(keep, tv.ops.batched_nms(boxes, scores, ids, thr))
    _, keep = O.batched_nms(torch.empty(0, 4), torch.empty(0), torch.empty(0, dtype=torch.long), dict(iou_threshold=0.5))
    assert keep.numel() == 0
