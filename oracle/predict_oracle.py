"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the NEXT row of the hot-path scope table
(SURVEY.md 8(f) rank 2): the GFL inference post-process that turns head outputs into
detections.  No product code exists for it yet; this file and ``tests/golden/predict_*.pt``
pin the behaviour a future CUDA implementation (which reuses the integral decode and the NMS
kernels of the loss path) has to reproduce.

Restates, citing the reference (paths relative to the reference root):
  GFLHead._predict_by_feat_single          mmdet/models/dense_heads/gfl_head.py:408-502
  BaseDenseHead.predict_by_feat            mmdet/models/dense_heads/base_dense_head.py:197-296
  BaseDenseHead._bbox_post_process         mmdet/models/dense_heads/base_dense_head.py:424-486
  filter_scores_and_topk                   mmdet/models/utils/misc.py:308-354
  Integral.forward                         mmdet/models/dense_heads/gfl_head.py:48-62
  distance2bbox (2-D fast path, clamp)     mmdet/structures/bbox/transforms.py:147-182
  get_box_wh                               mmdet/structures/bbox/transforms.py:417-433
``mmcv.ops.batched_nms`` is the same third-party boundary as in erd_oracle.py (PARITY UNPINNED
there); everything else is pinned against the real reference by
tests/test_oracle_predict.py::test_oracle_matches_reference and the golden fixtures.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

from .erd_oracle import STRIDES, batched_nms

TEST_CFG = dict(nms_pre=1000, min_bbox_size=0, score_thr=0.05, iou_threshold=0.6, max_per_img=100)
# configs/_base_/... test_cfg of the gfl configs (configs/gfl_increment/*.py: nms_pre=1000,
# min_bbox_size=0, score_thr=0.05, nms=dict(type='nms', iou_threshold=0.6), max_per_img=100)


def integral(bbox_pred: Tensor, reg_max: int) -> Tensor:
    """(..., 4*(reg_max+1)) -> (M, 4): softmax expectation per side (gfl_head.py:48-62)."""
    p = F.softmax(bbox_pred.reshape(-1, reg_max + 1), dim=1)
    proj = torch.linspace(0, reg_max, reg_max + 1).type_as(p)
    return F.linear(p, proj).reshape(-1, 4)


def level_points(h: int, w: int, stride: int) -> Tensor:
    """anchor_center of the level's priors (gfl_head.py:232-243 over anchor_generator.py:266-301):
    the anchors are centred on (x*s, y*s), index y*W + x."""
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing='ij')
    return torch.stack([xs.reshape(-1), ys.reshape(-1)], 1).float() * float(stride)


def predict_single(cls_scores: Sequence[Tensor], bbox_preds: Sequence[Tensor], img_shape: Tuple[int, int],
                   reg_max: int = 16, cfg: Dict = TEST_CFG) -> Dict[str, Tensor]:
    """One image.  cls_scores[l]: (C, H_l, W_l) logits, bbox_preds[l]: (4*(reg_max+1), H_l, W_l)."""
    mb, ms, ml = [], [], []
    for cls, box, stride in zip(cls_scores, bbox_preds, STRIDES):
        c, h, w = cls.shape
        dist = integral(box.permute(1, 2, 0), reg_max) * stride                     # gfl_head.py:466-467
        scores = cls.permute(1, 2, 0).reshape(-1, c).sigmoid()                      # :469-470
        valid = scores > cfg['score_thr']                                           # misc.py:333
        kept = scores[valid]
        idxs = torch.nonzero(valid)
        k = min(cfg['nms_pre'], idxs.size(0))
        kept, order = kept.sort(descending=True)                                    # misc.py:339
        kept = kept[:k]
        anchor, label = idxs[order[:k]].unbind(dim=1)
        pts = level_points(h, w, stride)[anchor]
        d = dist[anchor]
        b = torch.stack([pts[:, 0] - d[:, 0], pts[:, 1] - d[:, 1], pts[:, 0] + d[:, 2], pts[:, 1] + d[:, 3]], -1)
        b[:, 0::2].clamp_(min=0, max=img_shape[1])                                  # transforms.py:180-181
        b[:, 1::2].clamp_(min=0, max=img_shape[0])
        mb.append(b)
        ms.append(kept)
        ml.append(label)
    boxes, scores, labels = torch.cat(mb), torch.cat(ms), torch.cat(ml)
    if cfg.get('min_bbox_size', -1) >= 0:                                           # base_dense_head.py:470-474
        wv, hv = boxes[:, 2] - boxes[:, 0], boxes[:, 3] - boxes[:, 1]
        ok = (wv > cfg['min_bbox_size']) & (hv > cfg['min_bbox_size'])
        if not ok.all():
            boxes, scores, labels = boxes[ok], scores[ok], labels[ok]
    if boxes.numel() > 0:                                                           # :477-484
        dets, keep = batched_nms(boxes, scores, labels, dict(type='nms', iou_threshold=cfg['iou_threshold']))
        boxes, labels = boxes[keep], labels[keep]
        scores = dets[:, -1]
        m = cfg['max_per_img']
        boxes, scores, labels = boxes[:m], scores[:m], labels[:m]
    return dict(bboxes=boxes, scores=scores, labels=labels)


def predict_by_feat(cls_scores: Sequence[Tensor], bbox_preds: Sequence[Tensor], img_shapes: Sequence[Tuple[int, int]],
                    reg_max: int = 16, cfg: Dict = TEST_CFG) -> List[Dict[str, Tensor]]:
    """Batch form (base_dense_head.py:197-296, rescale=False, with_nms=True)."""
    return [predict_single([t[i] for t in cls_scores], [t[i] for t in bbox_preds], img_shapes[i], reg_max, cfg)
            for i in range(cls_scores[0].size(0))]
