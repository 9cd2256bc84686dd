"""The plugin inside a real mmdet (the reference fork): subclasses of the reference's OWN detector and
head that override only the hot-path methods, registered under their own names.

    custom_imports = dict(imports=['erd_b200.mmdet_plugin'], allow_failed_imports=False)
    model = dict(type='GFLIncrementERDB200', ..., bbox_head=dict(type='GFLHeadIncrementERDB200', ...))

Everything else of the reference config stays as it is: constructor keys, parameter names
(``cls_convs.N.conv.weight`` / ``scales.N.scale`` -- so ``load_checkpoint(strict=True)`` and the
checkpoint surgery of gfl_increment_erd.py:67-122 work), ``predict`` / ``predict_by_feat``, hooks.
Nothing is force-registered over a reference class.

Reference: ``GFLIncrementERD.sel_pos`` (mmdet/models/detectors/gfl_increment_erd.py:165-200) and
``GFLHeadIncrementERD.loss_by_feat`` (mmdet/models/dense_heads/gfl_head_increment_erd.py:334-454).

mmdet / mmcv / mmengine are not installable in the build image (no network, not in the wheelhouse), so
this module is exercised there only up to its import guard; the methods it installs are the same
functions the standalone mirrors (head.py, detector.py) run under the GPU parity tests.
"""
from __future__ import annotations

from .head import ErdPath, _LOSS_DEFAULTS, fused_loss_by_feat, fused_sel_pos, fused_teacher_head   # noqa: F401
from .detector import _LazyGather

try:
    from mmdet.registry import MODELS
    from mmdet.models.dense_heads.gfl_head_increment_erd import GFLHeadIncrementERD as _RefHead
    from mmdet.models.detectors.gfl_increment_erd import GFLIncrementERD as _RefDetector
    HAVE_MMDET = True
except Exception:   # mmdet absent (this image) or not the reference fork
    HAVE_MMDET = False


def path_from_reference_head(head) -> ErdPath:
    """An ErdPath configured like a constructed reference head (strides, anchor scale, loss weights, KD
    temperature); raises for configurations the fused kernels do not implement."""
    gen = head.prior_generator
    strides = tuple(int(s[0]) for s in gen.strides)
    if any(int(s[0]) != int(s[1]) for s in gen.strides):
        raise ValueError('h stride is not equal to w stride!')                  # gfl_head_increment_erd.py:256
    if list(getattr(gen, 'ratios', [1.0])) != [1.0] or len(getattr(gen, 'scales', [8.0])) != 1:
        raise ValueError('erd_b200 implements the single square anchor per location of the GFL configs')
    scale = float(getattr(gen, 'octave_base_scale', None) or gen.scales[0])
    for name, want in (('loss_cls', 'QualityFocalLoss'), ('loss_bbox', 'GIoULoss'),
                       ('loss_dfl', 'DistributionFocalLoss'), ('loss_ld', 'KnowledgeDistillationKLDivLoss')):
        if type(getattr(head, name)).__name__ != want:
            raise ValueError(f'{name}: the fused kernel implements {want}')
    assigner = getattr(head, 'assigner', None)
    if assigner is not None and (type(assigner).__name__ != 'ATSSAssigner' or int(getattr(assigner, 'topk', 9)) != 9):
        raise ValueError('erd_b200 fuses ATSSAssigner(topk=9)')
    weights = (float(head.loss_cls.loss_weight), float(head.loss_bbox.loss_weight), float(head.loss_dfl.loss_weight),
               float(head.loss_ld.loss_weight))
    return ErdPath(strides, scale, 0.005, weights, float(getattr(head.loss_ld, 'T', 10)))


if HAVE_MMDET:

    @MODELS.register_module()
    class GFLHeadIncrementERDB200(_RefHead):
        """The reference head with ``loss_by_feat`` running on the sm_100a kernels."""

        @property
        def path(self) -> ErdPath:
            if getattr(self, '_erd_path', None) is None:
                self._erd_path = path_from_reference_head(self)
            return self._erd_path

        @property
        def strides(self):
            return self.path.strides

        def loss_by_feat(self, ori_outs, new_outs, ori_topk_cls_inds, ori_topk_cls_scores, ori_topk_bbox_inds,
                         ori_topk_bbox_preds, ori_num_classes, dist_loss_weight, model, batch_gt_instances,
                         batch_img_metas, batch_gt_instances_ignore=None) -> dict:
            return fused_loss_by_feat(self, ori_outs, new_outs, ori_topk_cls_inds, ori_topk_bbox_inds,
                                      ori_num_classes, dist_loss_weight, batch_gt_instances, batch_img_metas)

    @MODELS.register_module()
    class GFLIncrementERDB200(_RefDetector):
        """The reference detector with ``sel_pos`` running on the sm_100a kernels."""

        def __init__(self, *args, fuse_teacher_head: bool = False, **kwargs):
            super().__init__(*args, **kwargs)
            # the teacher's last head convolutions inside the teacher pass (tcgen05, TF32)
            self.fuse_teacher_head = bool(fuse_teacher_head)

        def loss(self, batch_inputs, batch_data_samples):
            """gfl_increment_erd.py:202-220; with ``fuse_teacher_head`` the teacher head stops after its towers and
            ``erd_teacher_head_fused`` produces the logits, the teacher cache and the selection in one kernel."""
            if not self.fuse_teacher_head:
                return super().loss(batch_inputs, batch_data_samples)
            import torch
            head = self.bbox_head
            with torch.no_grad():
                ori_outs, (cls_sel, box_sel), _ = fused_teacher_head(
                    head.path, self.ori_model.bbox_head, self.ori_model.extract_feat(batch_inputs), head.num_classes,
                    head.reg_max)
            new_outs = self.bbox_head(self.extract_feat(batch_inputs))
            return self.bbox_head.loss(ori_outs, new_outs, batch_data_samples, cls_sel, _LazyGather(cls_sel, ori_outs[0]),
                                       box_sel, _LazyGather(box_sel, ori_outs[1]), self.ori_num_classes,
                                       self.dist_loss_weight, self)

        def sel_pos(self, cls_scores, bbox_preds):
            head = self.bbox_head
            _, t_cls, t_box, cls_sel, box_sel = fused_sel_pos(head.path, head.num_classes, head.reg_max,
                                                              self.ori_num_classes, cls_scores, bbox_preds)
            return (cls_sel, _LazyGather(cls_sel, t_cls), box_sel, _LazyGather(box_sel, t_box))
