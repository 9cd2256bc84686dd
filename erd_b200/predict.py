"""Host-side mirror of the GFL inference post-process, backed by ``erd_predict`` (csrc/predict.cu).

Reference: ``BaseDenseHead.predict_by_feat`` (mmdet/models/dense_heads/base_dense_head.py:197-296) over
``GFLHead._predict_by_feat_single`` (dense_heads/gfl_head.py:408-502) and ``_bbox_post_process``
(base_dense_head.py:424-486), ``with_nms=True``.  PyTorch is plumbing (allocation, current stream); there is
no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _native as N
from .ops import STRIDES, _check_level_tensors, _ptrs, _stream


class ErdPredictor:
    """``test_cfg`` of the gfl configs: nms_pre=1000, min_bbox_size=0, score_thr=0.05,
    nms=dict(type='nms', iou_threshold=0.6), max_per_img=100."""

    def __init__(self, strides: Sequence[int] = STRIDES, nms_pre: int = 1000, max_per_img: int = 100,
                 score_thr: float = 0.05, iou_threshold: float = 0.6, min_bbox_size: float = 0.0):
        self.lib = N.load()
        self.strides = tuple(int(s) for s in strides)
        self.cfg = N.ErdPredictConfig(int(nms_pre), int(max_per_img), float(score_thr), float(iou_threshold),
                                      float(min_bbox_size))
        self._ws: Dict[tuple, Tuple[N.ErdShape, torch.Tensor]] = {}

    def _plan(self, cls_scores: Sequence[torch.Tensor], reg_max: int):
        t0 = cls_scores[0]
        if not t0.is_cuda:
            raise RuntimeError('erd_b200 runs on CUDA tensors only; there is no CPU fallback')
        shapes = tuple((int(t.shape[2]), int(t.shape[3])) for t in cls_scores)
        key = (int(t0.shape[0]), int(t0.shape[1]), shapes, int(reg_max), t0.device)
        if key not in self._ws:
            sh = N.ErdShape()
            sh.num_imgs, sh.num_levels, sh.num_classes, sh.reg_max = key[0], len(shapes), key[1], int(reg_max)
            for l in range(min(len(shapes), N.MAX_LEVELS)):
                sh.level_h[l], sh.level_w[l] = shapes[l]
                sh.stride[l] = self.strides[l]
            nbytes = C.c_size_t()
            N.check(self.lib.erd_predict_workspace_bytes(C.byref(sh), C.byref(self.cfg), C.byref(nbytes)),
                    'erd_predict_workspace_bytes')
            if len(self._ws) >= 4:
                self._ws.pop(next(iter(self._ws)))
            self._ws[key] = (sh, torch.empty(int(nbytes.value), dtype=torch.uint8, device=t0.device))
        return self._ws[key]

    def predict_by_feat(self, cls_scores: Sequence[torch.Tensor], bbox_preds: Sequence[torch.Tensor],
                        img_shapes: Sequence[Tuple[int, int]], scale_factors: Optional[Sequence[Tuple[float, float]]] = None,
                        reg_max: int = 16) -> List[Dict[str, torch.Tensor]]:
        """Detections per image: dict(bboxes (M,4), scores (M,), labels (M,) int64), M <= max_per_img, in
        descending score order.  ``scale_factors`` ((w, h) per image): rescale=True (base_dense_head.py:458-461)."""
        assert len(cls_scores) == len(bbox_preds)                                   # base_dense_head.py:245
        cls_scores = [t.detach().contiguous() for t in cls_scores]
        bbox_preds = [t.detach().contiguous() for t in bbox_preds]
        sh, ws = self._plan(cls_scores, reg_max)
        n = sh.num_imgs
        shapes = [(int(t.shape[2]), int(t.shape[3])) for t in cls_scores]
        _check_level_tensors('cls_scores', cls_scores, n, sh.num_classes, shapes)
        _check_level_tensors('bbox_preds', bbox_preds, n, 4 * (reg_max + 1), shapes)
        if len(img_shapes) != n:
            raise AssertionError('one img_shape per image expected')
        dev = cls_scores[0].device
        hw = torch.tensor([[int(s[0]), int(s[1])] for s in img_shapes], dtype=torch.int32).to(dev, non_blocking=True)
        inv = None
        if scale_factors is not None:
            inv = torch.tensor([[1 / float(s[0]), 1 / float(s[1])] for s in scale_factors], dtype=torch.float32).to(dev)
        m = int(self.cfg.max_per_img)
        dets = torch.empty(n, m, 5, dtype=torch.float32, device=dev)
        labels = torch.empty(n, m, dtype=torch.int32, device=dev)
        num = torch.empty(n, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            N.check(self.lib.erd_predict(C.byref(sh), C.byref(self.cfg), _ptrs(cls_scores), _ptrs(bbox_preds),
                                         hw.data_ptr(), inv.data_ptr() if inv is not None else None, dets.data_ptr(),
                                         labels.data_ptr(), num.data_ptr(), ws.data_ptr(), _stream()), 'erd_predict')
        counts = num.cpu().tolist()   # the one host sync: results are variable-length
        return [dict(bboxes=dets[i, :k, :4], scores=dets[i, :k, 4], labels=labels[i, :k].long())
                for i, k in enumerate(counts)]
