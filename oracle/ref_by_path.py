"""TEST INFRASTRUCTURE ONLY -- loads the *real* Hi-FT/ERD hot-path files by path.

The reference is a fork of MMDetection 3.0.0 whose package ``__init__`` files
import ``mmcv``/``mmengine`` (not installable here, no network).  This module
pre-seeds ``sys.modules`` with empty stub packages and a ~100 line shim of the
few mmengine/mmcv symbols the hot path touches, then executes the reference's
own source files *unmodified, where they lie* under ``ERD_REFERENCE_ROOT``
(default ``/root/reference``).  Nothing is copied.

It only works in the build container (``/root/reference`` does not exist on
the GPU box); it is used by ``oracle/make_golden.py`` to generate the fixtures
under ``tests/golden/`` and by ``tests/test_oracle_vs_reference.py`` (skipped
when the reference tree is absent) to pin the portable restatement in
``oracle/erd_oracle.py``.

``mmcv.ops.batched_nms`` is third-party and un-vendored (mmcv>=2.0.0rc4,<2.1.0,
``requirements/mminstall.txt:1``); it is restated in ``erd_oracle.batched_nms``
from mmcv 2.0.x published semantics.  PARITY UNPINNED at that one boundary:
the reference's tests hold no NMS golden vector.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get('ERD_REFERENCE_ROOT', '/root/reference')

_LOADED = None


def available() -> bool:
    return os.path.isfile(
        os.path.join(REF_ROOT, 'mmdet/models/dense_heads/gfl_head_increment_erd.py'))


class _Registry:
    """dict-backed stand-in for mmengine.registry.Registry (mmdet/registry.py:62,102)."""

    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def register_module(self, *args, **kwargs):
        module = kwargs.get('module')
        if module is not None:
            self.module_dict[kwargs.get('name') or module.__name__] = module
            return module

        def deco(cls):
            self.module_dict[cls.__name__] = cls
            return cls
        return deco

    def build(self, cfg, default_args=None):
        cfg = dict(cfg)
        if default_args:
            for k, v in default_args.items():
                cfg.setdefault(k, v)
        typ = cfg.pop('type')
        cls = self.module_dict[typ] if isinstance(typ, str) else typ
        return cls(**cfg)


class InstanceData:
    """Attribute bag with ``in`` support (atss_assigner.py:139 uses it)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def __contains__(self, k):
        return k in self.__dict__

    def __len__(self):
        for v in self.__dict__.values():
            return len(v)
        return 0

    # the inference post-process indexes its results (base_dense_head.py:474,481-484)
    def __getitem__(self, item):
        return InstanceData(**{k: v[item] for k, v in self.__dict__.items()})

    def pop(self, k):
        return self.__dict__.pop(k)


class ConfigDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


class _BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg


class _ConvModule(nn.Module):
    def __init__(self, cin, cout, k, stride=1, padding=0, conv_cfg=None,
                 norm_cfg=None, **kw):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, stride=stride, padding=padding)
        self.gn = nn.GroupNorm(32, cout) if cout % 32 == 0 else nn.Identity()

    def forward(self, x):
        return torch.relu(self.gn(self.conv(x)))


class _Scale(nn.Module):
    def __init__(self, scale=1.0):
        super().__init__()
        self.scale = nn.Parameter(torch.tensor(scale, dtype=torch.float))

    def forward(self, x):
        return x * self.scale


def _pkg(name):
    m = types.ModuleType(name)
    m.__path__ = []
    sys.modules[name] = m
    parent, _, child = name.rpartition('.')
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], child, m)
    return m


def _load(modname, relpath):
    path = os.path.join(REF_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    parent, _, child = modname.rpartition('.')
    setattr(sys.modules[parent], child, mod)
    return mod


def load_reference(batched_nms=None):
    """Execute the reference's hot-path sources; returns a namespace of classes.

    ``batched_nms``: callable installed as ``mmcv.ops.batched_nms`` (defaults to
    the restatement in ``oracle.erd_oracle``).
    """
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not available():
        raise RuntimeError(f'reference tree not found under {REF_ROOT}')
    if 'mmdet' in sys.modules and not getattr(sys.modules['mmdet'], '_erd_shim', False):
        raise RuntimeError('a real mmdet is already imported; shim refuses to shadow it')

    for name in [
            'mmdet', 'mmdet.models', 'mmdet.models.losses', 'mmdet.models.task_modules',
            'mmdet.models.task_modules.assigners', 'mmdet.models.task_modules.coders',
            'mmdet.models.task_modules.samplers',
            'mmdet.models.task_modules.prior_generators', 'mmdet.models.dense_heads',
            'mmdet.models.detectors', 'mmdet.models.utils', 'mmdet.models.test_time_augs',
            'mmdet.structures', 'mmdet.structures.bbox', 'mmdet.structures.mask',
            'mmdet.utils', 'mmdet.registry', 'mmengine', 'mmengine.structures',
            'mmengine.config', 'mmengine.model', 'mmengine.utils', 'mmengine.dist',
            'mmengine.registry', 'mmengine.runner', 'mmengine.runner.checkpoint',
            'mmcv', 'mmcv.ops', 'mmcv.cnn']:
        _pkg(name)
    sm = sys.modules
    sm['mmdet']._erd_shim = True

    MODELS, TASK_UTILS = _Registry('model'), _Registry('task util')
    sm['mmdet.registry'].MODELS = MODELS
    sm['mmdet.registry'].TASK_UTILS = TASK_UTILS
    sm['mmengine.registry'].MODELS = MODELS
    sm['mmengine'].Config = ConfigDict
    sm['mmengine.runner.checkpoint'].load_checkpoint = lambda *a, **k: None
    sm['mmengine.runner.checkpoint'].load_state_dict = lambda *a, **k: None
    sm['mmengine.structures'].InstanceData = InstanceData
    sm['mmengine.config'].ConfigDict = ConfigDict
    sm['mmengine.model'].BaseModule = _BaseModule
    sm['mmengine.model'].constant_init = lambda *a, **k: None
    sm['mmengine.utils'].is_tuple_of = lambda seq, typ: isinstance(seq, tuple) and all(
        isinstance(s, typ) for s in seq)
    sm['mmengine.utils'].digit_version = lambda v: tuple(
        int(x) for x in v.split('+')[0].split('.')[:3] if x.isdigit())
    sm['mmengine.dist'].get_dist_info = lambda: (0, 1)
    sm['mmcv.cnn'].ConvModule = _ConvModule
    sm['mmcv.cnn'].Scale = _Scale
    if batched_nms is None:
        from oracle.erd_oracle import batched_nms as _bn
        batched_nms = _bn
    sm['mmcv.ops'].batched_nms = batched_nms

    u = sm['mmdet.utils']
    for alias in ['ConfigType', 'InstanceList', 'MultiConfig', 'OptConfigType',
                  'OptInstanceList', 'OptMultiConfig']:
        setattr(u, alias, object)
    sm['mmdet.structures'].SampleList = list
    sm['mmdet.structures'].OptSampleList = list
    sb = sm['mmdet.structures.bbox']

    class BaseBoxes:  # never instantiated on this path
        pass

    class HorizontalBoxes(BaseBoxes):
        pass
    sb.BaseBoxes, sb.HorizontalBoxes = BaseBoxes, HorizontalBoxes
    sb.get_box_tensor = lambda b: b
    sb.cat_boxes = lambda data, dim=0: torch.cat(data, dim=dim)
    sb.stack_boxes = lambda data, dim=0: torch.stack(data, dim=dim)
    sb.get_box_type = lambda t: (None, None)
    sb.get_box_wh = sb.scale_boxes = None
    sm['mmdet.structures.mask'].BitmapMasks = type('BitmapMasks', (), {})
    sm['mmdet.structures.mask'].PolygonMasks = type('PolygonMasks', (), {})
    sm['mmdet.models.test_time_augs'].merge_aug_results = None

    class CrossEntropyLoss(nn.Module):  # stray replay loss, gfl_head.py:150-151
        def __init__(self, **kw):
            super().__init__()
    MODELS.register_module(module=CrossEntropyLoss)

    _load('mmdet.utils.util_mixins', 'mmdet/utils/util_mixins.py')
    _load('mmdet.utils.util_random', 'mmdet/utils/util_random.py')
    du = _load('mmdet.utils.dist_utils', 'mmdet/utils/dist_utils.py')
    u.reduce_mean = du.reduce_mean
    bo = _load('mmdet.structures.bbox.bbox_overlaps', 'mmdet/structures/bbox/bbox_overlaps.py')
    sb.bbox_overlaps = bo.bbox_overlaps
    tr = _load('mmdet.structures.bbox.transforms', 'mmdet/structures/bbox/transforms.py')
    sb.distance2bbox, sb.bbox2distance = tr.distance2bbox, tr.bbox2distance
    sb.get_box_wh, sb.scale_boxes = tr.get_box_wh, tr.scale_boxes   # base_dense_head.py:461,471

    ml = 'mmdet.models.losses.'
    _load(ml + 'utils', 'mmdet/models/losses/utils.py')
    gf = _load(ml + 'gfocal_loss', 'mmdet/models/losses/gfocal_loss.py')
    kd = _load(ml + 'kd_loss', 'mmdet/models/losses/kd_loss.py')
    io = _load(ml + 'iou_loss', 'mmdet/models/losses/iou_loss.py')

    ta = 'mmdet.models.task_modules.'
    ar = _load(ta + 'assigners.assign_result', 'mmdet/models/task_modules/assigners/assign_result.py')
    sm[ta + 'assigners'].AssignResult = ar.AssignResult
    _load(ta + 'assigners.base_assigner', 'mmdet/models/task_modules/assigners/base_assigner.py')
    i2 = _load(ta + 'assigners.iou2d_calculator', 'mmdet/models/task_modules/assigners/iou2d_calculator.py')
    at = _load(ta + 'assigners.atss_assigner', 'mmdet/models/task_modules/assigners/atss_assigner.py')
    _load(ta + 'coders.base_bbox_coder', 'mmdet/models/task_modules/coders/base_bbox_coder.py')
    dc = _load(ta + 'coders.distance_point_bbox_coder',
               'mmdet/models/task_modules/coders/distance_point_bbox_coder.py')
    ag = _load(ta + 'prior_generators.anchor_generator',
               'mmdet/models/task_modules/prior_generators/anchor_generator.py')
    pu = _load(ta + 'prior_generators.utils', 'mmdet/models/task_modules/prior_generators/utils.py')
    pg = sm[ta + 'prior_generators']
    pg.AnchorGenerator, pg.anchor_inside_flags = ag.AnchorGenerator, pu.anchor_inside_flags
    pg.SSDAnchorGenerator = getattr(ag, 'SSDAnchorGenerator', None)
    sr = _load(ta + 'samplers.sampling_result', 'mmdet/models/task_modules/samplers/sampling_result.py')
    sm[ta + 'samplers'].SamplingResult = sr.SamplingResult
    _load(ta + 'samplers.base_sampler', 'mmdet/models/task_modules/samplers/base_sampler.py')
    ps = _load(ta + 'samplers.pseudo_sampler', 'mmdet/models/task_modules/samplers/pseudo_sampler.py')
    sm[ta + 'samplers'].PseudoSampler = ps.PseudoSampler

    mi = _load('mmdet.models.utils.misc', 'mmdet/models/utils/misc.py')
    mu = sm['mmdet.models.utils']
    for fn in ['multi_apply', 'unmap', 'images_to_levels', 'unpack_gt_instances',
               'filter_scores_and_topk', 'select_single_mlvl']:
        setattr(mu, fn, getattr(mi, fn))

    dh = 'mmdet.models.dense_heads.'
    _load(dh + 'base_dense_head', 'mmdet/models/dense_heads/base_dense_head.py')
    _load(dh + 'anchor_head', 'mmdet/models/dense_heads/anchor_head.py')
    gh = _load(dh + 'gfl_head', 'mmdet/models/dense_heads/gfl_head.py')
    eh = _load(dh + 'gfl_head_increment_erd', 'mmdet/models/dense_heads/gfl_head_increment_erd.py')

    # The ERD detector: stub only its base class (GFL -> SingleStageDetector ->
    # BaseDetector is backbone/neck plumbing, out of scope); sel_pos / sel_pos_single
    # (gfl_increment_erd.py:143-200) then execute as written.
    gfl_stub = types.ModuleType('mmdet.models.detectors.gfl')
    gfl_stub.GFL = type('GFL', (nn.Module,), {})
    sm['mmdet.models.detectors.gfl'] = gfl_stub
    ed = _load('mmdet.models.detectors.gfl_increment_erd',
               'mmdet/models/detectors/gfl_increment_erd.py')

    ns = types.SimpleNamespace(
        MODELS=MODELS, TASK_UTILS=TASK_UTILS, InstanceData=InstanceData,
        ConfigDict=ConfigDict, GFLHead=gh.GFLHead,
        GFLHeadIncrementERD=eh.GFLHeadIncrementERD, GFLIncrementERD=ed.GFLIncrementERD,
        ATSSAssigner=at.ATSSAssigner, BboxOverlaps2D=i2.BboxOverlaps2D,
        bbox_overlaps=bo.bbox_overlaps, AnchorGenerator=ag.AnchorGenerator,
        DistancePointBBoxCoder=dc.DistancePointBBoxCoder,
        QualityFocalLoss=gf.QualityFocalLoss, DistributionFocalLoss=gf.DistributionFocalLoss,
        GIoULoss=io.GIoULoss, KnowledgeDistillationKLDivLoss=kd.KnowledgeDistillationKLDivLoss,
        quality_focal_loss=gf.quality_focal_loss,
        distribution_focal_loss=gf.distribution_focal_loss, giou_loss=io.giou_loss,
        knowledge_distillation_kl_div_loss=kd.knowledge_distillation_kl_div_loss,
        distance2bbox=tr.distance2bbox, bbox2distance=tr.bbox2distance,
        reduce_mean=du.reduce_mean, multi_apply=mi.multi_apply)
    _LOADED = ns
    return ns


def head_cfg(num_classes=80, reg_max=16):
    """bbox_head dict of configs/gfl_increment/gfl_r50_fpn_1x_coco_first_40_incre_last_40_cats.py:57-90."""
    return dict(
        num_classes=num_classes, in_channels=256, stacked_convs=4, feat_channels=256,
        anchor_generator=dict(type='AnchorGenerator', ratios=[1.0], octave_base_scale=8,
                              scales_per_octave=1, strides=[8, 16, 32, 64, 128]),
        loss_cls=dict(type='QualityFocalLoss', use_sigmoid=True, beta=2.0, loss_weight=1.0),
        loss_dfl=dict(type='DistributionFocalLoss', loss_weight=0.25),
        loss_ld=dict(type='KnowledgeDistillationKLDivLoss', loss_weight=0.25, T=10),
        reg_max=reg_max, loss_bbox=dict(type='GIoULoss', loss_weight=2.0),
        train_cfg=ConfigDict(assigner=dict(type='ATSSAssigner', topk=9), allowed_border=-1,
                             pos_weight=-1, debug=False))


def build_reference_head(num_classes=80, reg_max=16):
    ref = load_reference()
    return ref.GFLHeadIncrementERD(**head_cfg(num_classes, reg_max))


def build_reference_detector_stub(ori_num_classes=40, reg_max=16):
    """A GFLIncrementERD whose __init__ is bypassed (it would build backbones and
    load checkpoints); only the attributes sel_pos reads are provided
    (gfl_increment_erd.py:185,189)."""
    ref = load_reference()
    det = ref.GFLIncrementERD.__new__(ref.GFLIncrementERD)
    nn.Module.__init__(det)
    bbox_head = types.SimpleNamespace(cls_out_channels=ori_num_classes, reg_max=reg_max)
    object.__setattr__(det, 'ori_model', types.SimpleNamespace(bbox_head=bbox_head))
    det.ori_num_classes = ori_num_classes
    det.dist_loss_weight = 1
    return det
