"""Host-side mirror of the reference's ERD detector for the part on the hot path:
``sel_pos`` (Elastic Response Selection) and the ``loss`` glue.

Reference: ``GFLIncrementERD`` (mmdet/models/detectors/gfl_increment_erd.py:20-220).  Backbone,
neck and teacher construction from a config (:95-122) are out of scope; the checkpoint
surgery that seeds the student from the teacher (:67-93) is ``load_checkpoint_for_new_model``;
the student / teacher networks are injected as modules so that the same ``loss`` sequence
(:202-220) runs: teacher forward, ``sel_pos``, student forward, ``bbox_head.loss``.  This is the
STANDALONE mirror; inside a real mmdet use ``erd_b200/mmdet_plugin.py`` (subclasses of the reference's
own classes, registered under their own names).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch
import torch.nn as nn
from torch import Tensor

from .head import ErsSelection, GFLHeadIncrementERD, fused_sel_pos, fused_teacher_head


class GFLIncrementERD(nn.Module):
    def __init__(self, bbox_head: GFLHeadIncrementERD, ori_num_classes: int, ori_model: Optional[nn.Module] = None,
                 extract_feat: Optional[Callable] = None, top_k: int = 100, dist_loss_weight: float = 1,
                 fuse_teacher_head: bool = False) -> None:
        super().__init__()
        self.bbox_head = bbox_head
        self.ori_num_classes = int(ori_num_classes)
        self.top_k = top_k                      # unused, as in the reference (:49,61)
        self.dist_loss_weight = dist_loss_weight
        self.ori_model = ori_model              # frozen teacher: images -> (cls_scores, bbox_preds)
        self._extract_feat = extract_feat       # student backbone + neck
        # run the teacher's last head convolutions inside the teacher pass.  Needs a teacher with
        # ``extract_feat`` and a ``bbox_head`` that exposes its towers (see ``fused_teacher_head``).
        self.fuse_teacher_head = bool(fuse_teacher_head)
        if ori_model is not None:
            for p in ori_model.parameters():    # :115-116
                p.requires_grad = False

    def load_checkpoint_for_new_model(self, checkpoint, strict: bool = True):
        """gfl_increment_erd.py:67-93 (`_load_checkpoint_for_new_model`): initialise the student from the teacher's
        checkpoint -- a state dict, a ``{'state_dict': ...}`` checkpoint or a file of either -- whose classification
        conv has ``ori_num_classes`` outputs: the rows of the new classes keep the student's own initialisation
        (``gfl_cls.weight/bias[ori_num_classes:]`` are appended to the checkpoint's).  Keys are matched against this
        module (``bbox_head.*`` and, when the student body is a module, ``extract_feat.*``); a ``module.`` prefix is
        stripped as the reference does.  Returns torch's (missing, unexpected) key report."""
        from collections import OrderedDict
        if isinstance(checkpoint, (str, bytes)):
            checkpoint = torch.load(checkpoint, map_location='cpu')
        if isinstance(checkpoint, OrderedDict):
            state = checkpoint
        elif isinstance(checkpoint, dict) and 'state_dict' in checkpoint:
            state = checkpoint['state_dict']
        else:
            raise RuntimeError('No state_dict found in checkpoint')                                       # :75-77
        if list(state.keys())[0].startswith('module.'):                                                   # :79-81
            state = {k[7:]: v for k, v in state.items()}
        state = dict(state)
        head = self.bbox_head
        for name, own in (('bbox_head.gfl_cls.weight', head.gfl_cls.weight), ('bbox_head.gfl_cls.bias', head.gfl_cls.bias)):
            old = state[name]
            if old.shape[0] != self.ori_num_classes:
                raise RuntimeError(f'{name}: checkpoint has {old.shape[0]} classes, ori_num_classes is {self.ori_num_classes}')
            added = own.detach()[self.ori_num_classes:].to(old)                                            # :83-84
            state[name] = torch.cat((old, added), dim=0)                                                  # :85-88
        return self.load_state_dict(state, strict=strict)

    def sel_pos(self, cls_scores: Sequence[Tensor], bbox_preds: Sequence[Tensor]):
        """gfl_increment_erd.py:165-200.  Returns (topk_cls_inds, topk_cls_scores,
        topk_bbox_inds, topk_bbox_preds); the index lists are lazily materialised views of the
        device-side selection, the two gathered-value lists -- which ``loss_by_feat`` never
        reads (:339-340) -- are gathered only on access."""
        head = self.bbox_head
        _, t_cls, t_box, cls_sel, box_sel = fused_sel_pos(head.path, head.num_classes, head.reg_max,
                                                          self.ori_num_classes, cls_scores, bbox_preds)
        return (cls_sel, _LazyGather(cls_sel, t_cls), box_sel, _LazyGather(box_sel, t_box))

    def loss(self, batch_inputs: Tensor, batch_data_samples) -> dict:
        """gfl_increment_erd.py:202-220."""
        if self.fuse_teacher_head:
            head = self.bbox_head
            with torch.no_grad():
                ori_outs, (cls_sel, box_sel), _ = fused_teacher_head(
                    head.path, self.ori_model.bbox_head, self.ori_model.extract_feat(batch_inputs), head.num_classes,
                    head.reg_max)
            sel = (cls_sel, _LazyGather(cls_sel, ori_outs[0]), box_sel, _LazyGather(box_sel, ori_outs[1]))
        else:
            with torch.no_grad():   # the reference relies on requires_grad=False instead (:205)
                ori_outs = self.ori_model(batch_inputs)
            sel = self.sel_pos(*ori_outs)
        new_outs = self.bbox_head(self._extract_feat(batch_inputs))
        return self.bbox_head.loss(ori_outs, new_outs, batch_data_samples, *sel, self.ori_num_classes,
                                   self.dist_loss_weight, self)


class _LazyGather:
    """topk_cls_scores / topk_bbox_preds of sel_pos_single (:152-153,160-161): rows of the
    flattened (A, C) teacher tensors at the selected anchors, built on first access."""

    def __init__(self, sel: ErsSelection, levels: List[Tensor]):
        self.sel, self.levels, self._items = sel, levels, None

    def _materialise(self):
        if self._items is None:
            n = self.levels[0].size(0)
            flat = torch.cat([t.permute(0, 2, 3, 1).reshape(n, -1, t.size(1)) for t in self.levels], dim=1)
            self._items = [flat[i][self.sel[i]] for i in range(n)]
        return self._items

    def __len__(self):
        return len(self.sel)

    def __getitem__(self, i):
        return self._materialise()[i]
