"""Host-side mirror of the reference's registered ERD head: same class name, constructor
keys and ``loss`` / ``loss_by_feat`` signatures, backed by the sm_100a kernels.

Reference: ``GFLHeadIncrementERD`` (mmdet/models/dense_heads/gfl_head_increment_erd.py:57-484)
on top of ``GFLHead`` (dense_heads/gfl_head.py:65-230 for the conv stacks, which stay on
PyTorch/cuDNN as the north star prescribes).  This class is the STANDALONE mirror (no mmdet
needed; flat conv towers).  Inside a real mmdet, ``erd_b200/mmdet_plugin.py`` subclasses the
reference's own head instead and overrides only the hot-path methods with the functions below, so
parameter names, ``predict`` and checkpoint surgery stay the reference's; nothing here replaces a
registered reference class.
"""
from __future__ import annotations

from collections.abc import Sequence as _SequenceABC
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
from torch import Tensor

from .ops import ErdPath, Plan

_LOSS_DEFAULTS = dict(
    loss_cls=dict(type='QualityFocalLoss', use_sigmoid=True, beta=2.0, loss_weight=1.0),
    loss_bbox=dict(type='GIoULoss', loss_weight=2.0),
    loss_dfl=dict(type='DistributionFocalLoss', loss_weight=0.25),
    loss_ld=dict(type='KnowledgeDistillationKLDivLoss', loss_weight=0.25, T=10))


class ErsSelection(_SequenceABC):
    """Per-image index tensors of one ERS list, materialised (one host sync) only when
    somebody reads them; the fused path never does."""

    def __init__(self, plan: Plan, kind: str, generation: int):
        self.plan, self.kind, self.generation = plan, kind, generation
        self._items: Optional[List[Tensor]] = None

    def _materialise(self) -> List[Tensor]:
        if self._items is None:
            p = self.plan
            if p.ers_generation != self.generation:
                raise RuntimeError('ERS selection is stale: sel_pos ran again on this geometry')
            inds, cnt = (p.cls_inds, p.cls_count) if self.kind == 'cls' else (p.box_inds, p.box_count)
            cnt = cnt.cpu()
            self._items = [inds[i, :int(cnt[i])].long() for i in range(p.n)]
        return self._items

    def __len__(self):
        return self.plan.n

    def __getitem__(self, i):
        return self._materialise()[i]


class _ErdLossFn(torch.autograd.Function):
    """losses (3L+2N,) = f(student cls[5], student box[5]); gradients come out of the same
    kernel sweep as the losses.  backward re-derives them on the device only when the
    upstream gradients are not all ones."""

    @staticmethod
    def forward(ctx, head, plan, t_cls, t_box, dist_loss_weight, ers_done, *student):
        s_cls, s_box = list(student[:5]), list(student[5:])
        path: ErdPath = head.path
        _, losses, g_cls, g_box = path.step(t_cls, t_box, s_cls, s_box, None, None, None, head.num_classes,
                                            plan.ori, head.reg_max, dist_loss_weight, targets_set=True,
                                            ers_done=ers_done)
        plan.loss_generation = getattr(plan, 'loss_generation', 0) + 1
        ctx.saved = (path, plan, t_cls, t_box, s_cls, s_box, g_cls, g_box, float(dist_loss_weight),
                     plan.loss_generation)
        return losses

    @staticmethod
    def backward(ctx, grad_losses):
        path, plan, t_cls, t_box, s_cls, s_box, g_cls, g_box, dlw, gen = ctx.saved
        up = grad_losses.contiguous().float()
        if plan.loss_generation == gen:
            scratch = torch.empty_like(up)
            path.loss_fwd_bwd(plan, t_cls, t_box, s_cls, s_box, g_cls, g_box, scratch, dlw, upstream=up,
                              skip_if_unit=True)
        elif bool((up != 1).any()):
            raise RuntimeError('erd_b200: backward with weighted loss terms after the plan was reused')
        grads = [g if ctx.needs_input_grad[6 + i] else None for i, g in enumerate(g_cls + g_box)]
        return (None, None, None, None, None, None, *grads)


def fused_loss_by_feat(head, ori_outs, new_outs, ori_topk_cls_inds, ori_topk_bbox_inds, ori_num_classes,
                       dist_loss_weight, batch_gt_instances, batch_img_metas) -> dict:
    """The hot path behind ``loss_by_feat`` (gfl_head_increment_erd.py:334-454) for any head object that
    carries ``path`` (ErdPath), ``num_classes``, ``reg_max`` and ``strides`` -- the standalone mirror
    below or the subclass of the reference's head in ``mmdet_plugin.py``.  Returns
    dict(loss_cls[5], loss_bbox[5], loss_dfl[5], loss_dist_cls[N], loss_dist_bbox[N]) of 0-dim
    tensors attached to the autograd graph of ``new_outs``.  ``ori_topk_cls_scores``,
    ``ori_topk_bbox_preds`` and ``model`` of the reference signature are unused there too."""
    cls_scores, bbox_preds = new_outs
    t_cls, t_box = ori_outs
    assert len(cls_scores) == len(head.strides)                                  # :374
    num_imgs = cls_scores[0].size(0)
    assert len(batch_img_metas) == num_imgs and len(batch_gt_instances) == num_imgs   # gfl_head.py:517-518
    s_cls = [t.contiguous() for t in cls_scores]
    s_box = [t.contiguous() for t in bbox_preds]
    t_cls = [t[:, :ori_num_classes].detach().contiguous() for t in t_cls]
    t_box = [t.detach().contiguous() for t in t_box]
    plan = head.path.plan(s_cls, head.num_classes, int(ori_num_classes), head.reg_max,
                          max((int(g.bboxes.shape[0]) for g in batch_gt_instances), default=0))
    ers_done = adopt_selection(head.path, plan, t_cls, t_box, ori_topk_cls_inds, ori_topk_bbox_inds)
    plan.set_targets([g.bboxes for g in batch_gt_instances], [g.labels for g in batch_gt_instances],
                     [m['pad_shape'][:2] for m in batch_img_metas])
    vec = _ErdLossFn.apply(head, plan, t_cls, t_box, float(dist_loss_weight), ers_done, *s_cls, *s_box)
    L = len(head.strides)
    parts = vec.split([L, L, L, num_imgs, num_imgs])
    return dict(loss_cls=list(parts[0].unbind()), loss_bbox=list(parts[1].unbind()),
                loss_dfl=list(parts[2].unbind()), loss_dist_cls=list(parts[3].unbind()),
                loss_dist_bbox=list(parts[4].unbind()))


def adopt_selection(path: ErdPath, plan: Plan, t_cls, t_box, cls_inds, box_inds) -> bool:
    """True when the plan already holds ``sel_pos`` results for these teacher tensors.
    Foreign index lists (plain tensors) are honoured: the teacher cache is rebuilt and the
    given rows are loaded into the plan."""
    if (isinstance(cls_inds, ErsSelection) and isinstance(box_inds, ErsSelection)
            and cls_inds.plan is plan and cls_inds.generation == plan.ers_generation
            and box_inds.generation == plan.ers_generation):
        return True
    path.ers_select(plan, t_cls, t_box)
    plan.ers_generation += 1
    plan.load_selection(cls_inds, box_inds)
    return True


def fused_sel_pos(path: ErdPath, num_classes: int, reg_max: int, ori_num_classes: int, cls_scores, bbox_preds):
    """``GFLIncrementERD.sel_pos`` (gfl_increment_erd.py:165-200) on the device.  Returns the plan, the
    teacher tensors as the kernels read them, and the two lazily materialised index lists."""
    assert len(cls_scores) == len(bbox_preds)                                    # :180
    t_cls = [t[:, :ori_num_classes].detach().contiguous() for t in cls_scores]
    t_box = [t.detach().contiguous() for t in bbox_preds]
    plan = path.plan(t_cls, num_classes, ori_num_classes, reg_max)
    path.ers_select(plan, t_cls, t_box)
    gen = plan.ers_generation
    return plan, t_cls, t_box, ErsSelection(plan, 'cls', gen), ErsSelection(plan, 'box', gen)


def fused_teacher_head(path: ErdPath, teacher_head, feats, num_classes: int, reg_max: int):
    """The teacher's head with its LAST convolutions fused into the teacher pass (SURVEY 8(f) rank 1;
    ``gfl_head.py:205-230`` + ``gfl_increment_erd.py:143-200,205`` in one tcgen05 kernel): the towers run in
    PyTorch/cuDNN (channels_last, so their outputs are the NHWC tensors the kernel streams), ``gfl_cls`` /
    ``gfl_reg`` + bias + Scale run inside ``erd_teacher_head_fused``, whose epilogue writes the per-anchor teacher
    cache, the threshold sums and the stash; the logits are still emitted (write-only) because ``loss_by_feat``
    takes them as ``ori_outs``.  ``teacher_head``: a module with ``cls_convs``, ``reg_convs``, ``gfl_cls``,
    ``gfl_reg`` and ``scales`` (the standalone ``GFLHeadIncrementERD`` or mmdet's ``GFLHead``).
    Returns ((cls_scores, bbox_preds), (topk_cls_inds, topk_bbox_inds), plan)."""
    from .ops import TeacherHead
    th = getattr(teacher_head, '_erd_packed', None)
    if th is None or th.packed[0].device != feats[0].device:
        sc = teacher_head.scales
        scales = ([float(v) for v in sc.detach().flatten()] if isinstance(sc, torch.Tensor)
                  else [float(m.scale.detach()) for m in sc])      # mmdet: ModuleList of Scale
        th = TeacherHead(teacher_head.gfl_cls.weight, teacher_head.gfl_cls.bias, teacher_head.gfl_reg.weight,
                         teacher_head.gfl_reg.bias, scales)
        teacher_head._erd_packed = th       # the teacher is frozen: packed once
    def tower(convs, x):   # nn.Sequential (standalone mirror) or mmdet's ModuleList of ConvModule (gfl_head.py:222-227)
        if isinstance(convs, nn.ModuleList):
            for conv in convs:
                x = conv(x)
            return x
        return convs(x)
    cls_f, reg_f = [], []
    for f in feats:
        f = f.contiguous(memory_format=torch.channels_last)
        cls_f.append(tower(teacher_head.cls_convs, f).float().contiguous(memory_format=torch.channels_last))
        reg_f.append(tower(teacher_head.reg_convs, f).float().contiguous(memory_format=torch.channels_last))
    n, ori = int(cls_f[0].shape[0]), th.ori
    plan = path.plan(cls_f, num_classes, ori, reg_max)
    t_cls = [torch.empty(n, ori, h, w, dtype=torch.float32, device=cls_f[0].device) for h, w in plan.shapes]
    t_box = [torch.empty(n, 4 * (reg_max + 1), h, w, dtype=torch.float32, device=cls_f[0].device) for h, w in plan.shapes]
    path.teacher_head_fused(plan, th, cls_f, reg_f, t_cls, t_box)
    path.ers_select_cached(plan)
    gen = plan.ers_generation
    return (t_cls, t_box), (ErsSelection(plan, 'cls', gen), ErsSelection(plan, 'box', gen)), plan


def _cfg_get(cfg, key, default=None):
    return cfg.get(key, default) if cfg is not None else default


class GFLHeadIncrementERD(nn.Module):
    """Drop-in for the reference head.  Constructor keys follow
    configs/gfl_increment/gfl_r50_fpn_1x_coco_first_40_incre_last_40_cats.py:57-90."""

    def __init__(self, num_classes: int, in_channels: int, stacked_convs: int = 4, conv_cfg=None,
                 norm_cfg=dict(type='GN', num_groups=32, requires_grad=True), loss_dfl=None, loss_ld=None,
                 bbox_coder=dict(type='DistancePointBBoxCoder'), reg_max: int = 16, init_cfg=None,
                 feat_channels: int = 256, anchor_generator=None, loss_cls=None, loss_bbox=None,
                 train_cfg=None, test_cfg=None, build_convs: bool = True, **kwargs) -> None:
        super().__init__()
        self.num_classes = self.cls_out_channels = int(num_classes)
        self.in_channels, self.feat_channels = in_channels, feat_channels
        self.stacked_convs, self.reg_max = stacked_convs, int(reg_max)
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        ag = anchor_generator or dict(type='AnchorGenerator', ratios=[1.0], octave_base_scale=8,
                                      scales_per_octave=1, strides=[8, 16, 32, 64, 128])
        if list(ag.get('ratios', [1.0])) != [1.0] or ag.get('scales_per_octave', 1) != 1:
            raise ValueError('erd_b200 implements the single square anchor per location of the GFL configs')
        self.strides = tuple(int(s if not isinstance(s, (tuple, list)) else s[0]) for s in ag['strides'])
        for s in ag['strides']:
            if isinstance(s, (tuple, list)):
                assert s[0] == s[1], 'h stride is not equal to w stride!'     # gfl_head_increment_erd.py:256
        cfgs = {k: dict(_LOSS_DEFAULTS[k], **(v or {})) for k, v in
                dict(loss_cls=loss_cls, loss_bbox=loss_bbox, loss_dfl=loss_dfl, loss_ld=loss_ld).items()}
        for k, v in cfgs.items():
            if v['type'] != _LOSS_DEFAULTS[k]['type']:
                raise ValueError(f'{k}: the fused kernel implements {_LOSS_DEFAULTS[k]["type"]}, got {v["type"]}')
            if v.get('reduction', 'mean') != 'mean':
                raise ValueError(f'{k}: only reduction="mean" is fused')
        if float(cfgs['loss_cls'].get('beta', 2.0)) != 2.0 or not cfgs['loss_cls'].get('use_sigmoid', True):
            raise ValueError('QualityFocalLoss: only use_sigmoid=True, beta=2.0 is fused')
        if float(cfgs['loss_bbox'].get('eps', 1e-6)) != 1e-6:
            raise ValueError('GIoULoss: only eps=1e-6 is fused')
        assigner = _cfg_get(train_cfg, 'assigner', dict(type='ATSSAssigner', topk=9))
        if train_cfg is not None:
            if assigner.get('type') != 'ATSSAssigner' or int(assigner.get('topk', 9)) != 9:
                raise ValueError('erd_b200 fuses ATSSAssigner(topk=9)')
            if _cfg_get(train_cfg, 'allowed_border', -1) >= 0 or _cfg_get(train_cfg, 'pos_weight', -1) > 0:
                raise ValueError('erd_b200 fuses allowed_border=-1, pos_weight=-1 (the gfl_increment configs)')
        self.loss_cfgs = cfgs
        self.path = ErdPath(self.strides, float(ag.get('octave_base_scale', 8)), 0.005,
                            (cfgs['loss_cls']['loss_weight'], cfgs['loss_bbox']['loss_weight'],
                             cfgs['loss_dfl']['loss_weight'], cfgs['loss_ld']['loss_weight']),
                            float(cfgs['loss_ld'].get('T', 10)))
        if build_convs:
            self._init_layers(norm_cfg)

    # ---- conv stacks: plain PyTorch/cuDNN, gfl_head.py:153-230 --------------------------
    def _init_layers(self, norm_cfg):
        groups = int(norm_cfg.get('num_groups', 32)) if norm_cfg else 0

        def tower():
            layers = []
            for i in range(self.stacked_convs):
                cin = self.in_channels if i == 0 else self.feat_channels
                layers += [nn.Conv2d(cin, self.feat_channels, 3, padding=1, bias=not groups)]
                if groups:
                    layers += [nn.GroupNorm(groups, self.feat_channels)]
                layers += [nn.ReLU(inplace=True)]
            return nn.Sequential(*layers)
        self.cls_convs, self.reg_convs = tower(), tower()
        self.gfl_cls = nn.Conv2d(self.feat_channels, self.cls_out_channels, 3, padding=1)
        self.gfl_reg = nn.Conv2d(self.feat_channels, 4 * (self.reg_max + 1), 3, padding=1)
        self.scales = nn.Parameter(torch.ones(len(self.strides)))
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.normal_(m.weight, std=0.01)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
        nn.init.constant_(self.gfl_cls.bias, -4.59511985013459)   # bias_prob=0.01, gfl_head.py:115-123

    def forward(self, x: Sequence[Tensor]) -> Tuple[List[Tensor], List[Tensor]]:
        cls_scores, bbox_preds = [], []
        for lvl, feat in enumerate(x):
            cls_scores.append(self.gfl_cls(self.cls_convs(feat)))
            bbox_preds.append((self.gfl_reg(self.reg_convs(feat)) * self.scales[lvl]).float())
        return cls_scores, bbox_preds

    # ---- the hot path ------------------------------------------------------------------
    def loss_by_feat(self, ori_outs, new_outs, ori_topk_cls_inds, ori_topk_cls_scores, ori_topk_bbox_inds,
                     ori_topk_bbox_preds, ori_num_classes, dist_loss_weight, model, batch_gt_instances,
                     batch_img_metas, batch_gt_instances_ignore=None) -> dict:
        """Same contract as gfl_head_increment_erd.py:334-454 (see ``fused_loss_by_feat``)."""
        return fused_loss_by_feat(self, ori_outs, new_outs, ori_topk_cls_inds, ori_topk_bbox_inds, ori_num_classes,
                                  dist_loss_weight, batch_gt_instances, batch_img_metas)

    def predict_by_feat(self, cls_scores, bbox_preds, score_factors=None, batch_img_metas=None, cfg=None,
                        rescale: bool = False, with_nms: bool = True):
        """base_dense_head.py:197-296 over gfl_head.py:408-502 on the device (``erd_predict``).  Returns one
        object per image with ``bboxes`` (M,4), ``scores`` (M,), ``labels`` (M,) -- the fields of the reference's
        InstanceData."""
        from types import SimpleNamespace
        from .predict import ErdPredictor
        if not with_nms or score_factors is not None:
            raise ValueError('erd_b200 fuses the with_nms=True path of the GFL head (no score_factors)')
        cfg = cfg if cfg is not None else (self.test_cfg or {})
        key = (int(cfg.get('nms_pre', 1000)), int(cfg.get('max_per_img', 100)), float(cfg.get('score_thr', 0.05)),
               float(dict(cfg.get('nms', {})).get('iou_threshold', 0.6)), float(cfg.get('min_bbox_size', 0)))
        if getattr(self, '_predictor_key', None) != key:
            self._predictor, self._predictor_key = ErdPredictor(self.strides, *key), key
        scales = [m['scale_factor'] for m in batch_img_metas] if rescale else None
        out = self._predictor.predict_by_feat(cls_scores, bbox_preds, [m['img_shape'][:2] for m in batch_img_metas],
                                              scales, self.reg_max)
        return [SimpleNamespace(**d) for d in out]

    def loss(self, ori_outs, new_outs, batch_data_samples, topk_cls_inds, topk_cls_scores, topk_bbox_inds,
             topk_bbox_preds, ori_num_classes, dist_loss_weight, model) -> dict:
        """gfl_head_increment_erd.py:457-484 (unpack_gt_instances, models/utils/misc.py:89)."""
        gts, ignored, metas = [], [], []
        for ds in batch_data_samples:
            metas.append(ds.metainfo)
            gts.append(ds.gt_instances)
            ignored.append(getattr(ds, 'ignored_instances', None))
        return self.loss_by_feat(ori_outs, new_outs, topk_cls_inds, topk_cls_scores, topk_bbox_inds,
                                 topk_bbox_preds, ori_num_classes, dist_loss_weight, model, gts, metas,
                                 ignored if any(i is not None for i in ignored) else None)


def parse_losses(losses: dict) -> Tensor:
    """mmengine ``BaseModel.parse_losses`` semantics: sum over keys containing 'loss' of
    ``mean()`` (tensor) or sum of means (list)."""
    total = 0
    for k, v in losses.items():
        if 'loss' in k:
            total = total + (v.mean() if isinstance(v, Tensor) else sum(x.mean() for x in v))
    return total
